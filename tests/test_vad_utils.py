"""stft_vad / istft_vad (tssep/util/utils.py:11-129).

* CPU: the reference's OWN utils.py, imported from /root/reference on top of a stand-in paderbox whose three index
  helpers are the oracle's restatement (paderbox is absent, so the helpers themselves stay parity-unpinned), must agree
  with the oracle's stft_vad / istft_vad -- this pins the control flow the reference holds (run extraction, which bound
  goes through which mapping, half-open fills).  Committed goldens of that run travel to the GPU box.
* GPU: the product kernels (tssep_stft_vad, tssep_segments index_mode 1) against the oracle and the goldens.
"""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from oracle import tssep_oracle as O
from tests import ref_stub as RS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_vad_goldens.npz")
GEOMS = [(1024, 256, True), (1024, 256, False), (64, 32, True), (400, 200, True), (8, 2, False), (8, 8, True)]


def _cases():
    rng = np.random.RandomState(0)
    out = []
    for wl, shift, fading in GEOMS:
        n = 40 * wl // 8 + 37
        v = np.zeros((3, n), dtype=bool)
        for k in range(3):
            for _ in range(4):
                a = rng.randint(0, n)
                v[k, a:a + rng.randint(1, n // 3)] = True
        v[2, :3] = True       # a run touching sample 0
        v[1, -2:] = True      # a run touching the last sample
        out.append((wl, shift, fading, v))
    return out


class _Interval:
    """Minimal stand-in for paderbox.array.interval.ArrayInterval (half-open interval set)."""

    def __init__(self, a=None, shape=None):
        if isinstance(a, _Interval):
            self.intervals, self.shape = list(a.intervals), a.shape
        else:
            self.intervals = [] if a is None else O._runs(a)
            self.shape = shape if a is None else np.shape(a)

    @property
    def normalized_intervals(self):
        iv = sorted(self.intervals)
        merged = []
        for s, e in iv:
            if merged and s <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(merged[-1][1], e))
            elif e > s:
                merged.append((s, e))
        return tuple(merged)

    def __len__(self):
        return self.shape[-1]

    def __getitem__(self, key):
        assert key == (), key  # utils.py walks np.ndindex(shape[:-1]) = [()] for a 1-D interval set
        return self

    def __setitem__(self, key, value):
        assert value is True and key.step is None
        self.intervals.append((int(key.start), int(key.stop)))

    def to_array(self, n):
        out = np.zeros(n, dtype=bool)
        for s, e in self.normalized_intervals:
            out[s:e] = True
        return out


def _reference_utils():
    """The reference's tssep/util/utils.py on a stand-in paderbox."""
    if not RS.available():
        pytest.skip("reference tree not present on this machine")
    interval = types.ModuleType("paderbox.array.interval")
    interval.ArrayInterval = _Interval
    interval.zeros = lambda shape=None: _Interval(shape=(shape,) if shape is not None else None)
    array = types.ModuleType("paderbox.array")
    array.interval = interval
    module_stft = types.ModuleType("paderbox.transform.module_stft")
    module_stft._samples_to_stft_frames = O.pb_samples_to_stft_frames
    module_stft.sample_index_to_stft_frame_index = O.pb_sample_index_to_stft_frame_index
    module_stft.stft_frame_index_to_sample_index = O.pb_stft_frame_index_to_sample_index
    RS.load()  # installs the base stand-ins
    pb = sys.modules["paderbox"]
    pb.array = array
    sys.modules["paderbox.array"] = array
    sys.modules["paderbox.array.interval"] = interval
    sys.modules["paderbox.transform"].module_stft = module_stft
    sys.modules["paderbox.transform.module_stft"] = module_stft
    return importlib.import_module("tssep.util.utils")


def test_reference_utils_agree_with_the_oracle():
    U = _reference_utils()
    for wl, shift, fading, v in _cases():
        T = O.pb_samples_to_stft_frames(v.shape[-1], wl, shift, pad=True, fading=fading)
        want_f = O.stft_vad(v, wl, shift, fading)
        for k in range(v.shape[0]):
            ai = U.stft_vad(_Interval(v[k]), wl, shift, fading)  # ArrayInterval branch (utils.py:29-71)
            assert isinstance(ai, _Interval)
            got = ai.to_array(T)
            assert np.array_equal(got, want_f[k]), (wl, shift, fading, k)
        frames = want_f
        want_s = O.istft_vad(frames, wl, shift, fading)
        got_s = U.istft_vad(frames, wl, shift, fading)
        for k in range(v.shape[0]):
            merged = list(_Interval.normalized_intervals.fget(got_s[k]))
            ref = _Interval()
            ref.intervals = list(want_s[k])
            assert merged == list(ref.normalized_intervals), (wl, shift, fading, k)


def test_oracle_matches_committed_goldens():
    """Goldens = the oracle's outputs at the time the reference run above agreed with them (scripts/make_vad_goldens.py)."""
    g = np.load(GOLDEN, allow_pickle=False)
    for i, (wl, shift, fading, v) in enumerate(_cases()):
        assert np.array_equal(g[f"{i}/vad"], v)
        assert np.array_equal(O.stft_vad(v, wl, shift, fading), g[f"{i}/frames"])
        iv = O.istft_vad(g[f"{i}/frames"], wl, shift, fading)
        flat = np.array([[k, a, b] for k, row in enumerate(iv) for a, b in row], dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(flat, g[f"{i}/intervals"])


def test_frames_survive_the_round_trip_through_samples():
    """stft_vad(istft_vad(F)) == F for every frame activity F that stft_vad can produce: 'first' / 'last' are the set
    inverses of the sample -> frame mapping."""
    for wl, shift, fading, v in _cases():
        n = v.shape[-1]
        F = O.stft_vad(v, wl, shift, fading)
        iv = O.istft_vad(F, wl, shift, fading, num_samples=n)
        back = np.zeros_like(v)
        for k in range(v.shape[0]):
            for a, b in iv[k]:
                back[k, a:b] = True
        assert np.array_equal(O.stft_vad(back, wl, shift, fading), F), (wl, shift, fading)


@pytest.mark.gpu
def test_product_vad_utils_match_oracle(cuda):
    import torch

    from tssep_b200.util.utils import istft_vad, stft_vad

    g = np.load(GOLDEN, allow_pickle=False)
    for i, (wl, shift, fading, v) in enumerate(_cases()):
        frames = stft_vad(torch.tensor(v, device=cuda), wl, shift, fading)
        assert frames.dtype == torch.bool and np.array_equal(frames.cpu().numpy(), g[f"{i}/frames"])
        assert np.array_equal(stft_vad(v, wl, shift, fading), g[f"{i}/frames"])          # numpy in -> numpy out
        f_float = stft_vad(torch.tensor(v, device=cuda, dtype=torch.float32), wl, shift, fading)
        assert f_float.dtype == torch.float32                                           # utils.py:24-27 returns a float Tensor
        iv = istft_vad(frames, wl, shift, fading)
        assert iv == O.istft_vad(g[f"{i}/frames"], wl, shift, fading)
        clipped = istft_vad(frames, wl, shift, fading, num_samples=v.shape[-1])
        assert clipped == O.istft_vad(g[f"{i}/frames"], wl, shift, fading, num_samples=v.shape[-1])
    assert stft_vad([v[0], v[1]], wl, shift, fading)[1].tolist() == stft_vad(v[1], wl, shift, fading).tolist()
