"""Imports the reference's own ``tssep/train/{net,rnnp,feature_extractor_torchaudio}.py`` behind a minimal
stand-in for its absent third-party packages (padertorch==0.0.1, paderbox==0.0.8).

TEST INFRASTRUCTURE ONLY.  The stand-ins carry no arithmetic of the path: ``pt.Configurable`` is an empty mixin,
``pt.ops.sequence.sequence_elementwise(f, x)`` is ``f(x)`` for tensors (the only case ``rnnp.py:161`` meets here),
``paderbox.utils.iterable.zip`` is the builtin ``zip`` (``strict=`` keyword, ``net.py:830``) and the padertorch
``STFT`` base class only stores its constructor arguments (``TorchMFCC.stft_to_feature`` never touches it).
Everything numerical that runs is the reference's own code on torch / torchaudio / einops.

``/root/reference`` exists in the build container only: callers skip when it is absent, and the goldens made with it
(``scripts/make_reference_goldens.py`` -> ``tests/golden/reference_net_goldens.npz``) are what travels.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TSSEP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tssep", "train", "net.py"))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # lets "import a.b.c" resolve through sys.modules
    return m


def _install_stubs():
    import torch

    class Configurable:  # pt.Configurable: factory/config plumbing only, nothing on the numeric path
        pass

    def sequence_elementwise(function, x, *args, **kwargs):
        if isinstance(x, torch.nn.utils.rnn.PackedSequence):
            return torch.nn.utils.rnn.PackedSequence(function(x.data, *args, **kwargs), x.batch_sizes)
        return function(x, *args, **kwargs)

    class STFT:  # padertorch.contrib.cb.feature_extractor.STFT: argument record for TorchMFCC
        def __init__(self, size=1024, shift=256, window_length=None, pad=True, fading=True, output_size=None,
                     window="blackman"):
            super().__init__()
            self.size, self.shift = size, shift
            self.window_length = size if window_length is None else window_length
            self.pad, self.fading, self.window = pad, fading, window
            self.output_size = self._get_output_size(output_size)
            self.frequencies = size // 2 + 1

        def _get_output_size(self, output_size):
            return self.frequencies if output_size is None else output_size

    class _PbSTFT:
        @staticmethod
        def sample_index_to_frame_index(self, sample_index):
            raise NotImplementedError("paderbox is absent; not needed by the parity tests")

    mask_mod = _module("padertorch.ops.sequence.mask", compute_mask=None)
    seq_mod = _module("padertorch.ops.sequence", sequence_elementwise=sequence_elementwise, mask=mask_mod)
    ops_mod = _module("padertorch.ops", sequence=seq_mod)
    fe_mod = _module("padertorch.contrib.cb.feature_extractor", STFT=STFT)
    cb_mod = _module("padertorch.contrib.cb", feature_extractor=fe_mod)
    contrib_mod = _module("padertorch.contrib", cb=cb_mod)
    pt = _module("padertorch", Configurable=Configurable, ops=ops_mod, contrib=contrib_mod)
    it_mod = _module("paderbox.utils.iterable", zip=zip)
    utils_mod = _module("paderbox.utils", iterable=it_mod)
    tr_mod = _module("paderbox.transform", STFT=_PbSTFT)
    pb = _module("paderbox", utils=utils_mod, transform=tr_mod)
    stubs = {
        "padertorch": pt, "padertorch.ops": ops_mod, "padertorch.ops.sequence": seq_mod,
        "padertorch.ops.sequence.mask": mask_mod, "padertorch.contrib": contrib_mod, "padertorch.contrib.cb": cb_mod,
        "padertorch.contrib.cb.feature_extractor": fe_mod,
        "paderbox": pb, "paderbox.utils": utils_mod, "paderbox.utils.iterable": it_mod, "paderbox.transform": tr_mod,
    }
    for k, v in stubs.items():
        sys.modules.setdefault(k, v)


_cache = {}


def load():
    """Returns a namespace with the reference's ``net``, ``rnnp`` and ``feature_extractor_torchaudio`` modules."""
    if "ns" in _cache:
        return _cache["ns"]
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    ns = types.SimpleNamespace(
        rnnp=importlib.import_module("tssep.train.rnnp"),
        net=importlib.import_module("tssep.train.net"),
        mfcc=importlib.import_module("tssep.train.feature_extractor_torchaudio"),
    )
    _cache["ns"] = ns
    return ns


# The cases of tests/test_oracle_vs_reference_net.py and scripts/make_reference_goldens.py: the branches the reference's
# doctests pin by shape only.  Small dims so the goldens stay a few hundred KB.
CASES = {
    "mul_tsvad8_R2": dict(idim=21, odim=17, layers=3, units=6, projs=7, combination="mul", ts_vad=8,
                          aux_net_output_size=17, num_averaged_permutations=2, output_resolution="tf"),
    "mul_tsvad8_R2_t": dict(idim=21, odim=17, layers=3, units=6, projs=7, combination="mul", ts_vad=8,
                            aux_net_output_size=17, num_averaged_permutations=2, output_resolution="t"),
    "cat_tsvad4_R3": dict(idim=21, odim=17, layers=3, units=6, projs=7, combination="cat", ts_vad=4,
                          aux_net_output_size=5, num_averaged_permutations=3, output_resolution="tf"),
    "cat_explicit_vad": dict(idim=17, odim=17, layers=3, units=6, projs=7, combination="cat", ts_vad=False,
                             aux_net_output_size=5, num_averaged_permutations=1, output_resolution="tf",
                             explicit_vad=True),
    "mul_tsvad8_R1": dict(idim=21, odim=17, layers=3, units=6, projs=7, combination="mul", ts_vad=8,
                          aux_net_output_size=17, num_averaged_permutations=1, output_resolution="tf"),
}
T_FRAMES = 23


def case_inputs(name: str, batched: bool):
    """Seeded inputs of a case: features (T, idim) / (2, T, idim) and the K embeddings."""
    import numpy as np
    import torch

    kw = CASES[name]
    K = kw["ts_vad"] if kw["ts_vad"] else 3
    rng = np.random.RandomState(sum(map(ord, name)) + int(batched))
    B = 2 if batched else 1
    xs = torch.tensor(rng.randn(B, T_FRAMES, kw["idim"]).astype(np.float32))
    aux = torch.tensor(rng.rand(B, K, kw["aux_net_output_size"]).astype(np.float32))
    return (xs, aux) if batched else (xs[0], aux[0])


def aux_argument(aux, batched: bool):
    return [[a for a in item] for item in aux] if batched else [a for a in aux]


def load_enhancer():
    """The reference's ``tssep/train/enhancer.py`` (TorchBF is pure torch; pb_bss / paderbox are only needed by the
    numpy beamformers and WPE, which are not exercised)."""
    if "enh" in _cache:
        return _cache["enh"]
    load()
    for name in ("pb_bss", "pb_bss.testing", "pb_bss.testing.random_utils", "pb_bss.extraction", "pb_bss.extraction.beamformer"):
        sys.modules.setdefault(name, _module(name))
    sys.modules["pb_bss"].testing = sys.modules["pb_bss.testing"]
    sys.modules["pb_bss.testing"].random_utils = sys.modules["pb_bss.testing.random_utils"]
    _cache["enh"] = importlib.import_module("tssep.train.enhancer")
    return _cache["enh"]
