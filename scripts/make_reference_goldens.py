#!/usr/bin/env python
"""Generates ``tests/golden/reference_net_goldens.npz`` by running the REFERENCE's own code
(``/root/reference/tssep/train/net.py``, ``rnnp.py``, ``feature_extractor_torchaudio.py``, imported behind the
stand-ins of ``tests/ref_stub.py``) on seeded inputs and weights.

Run in the build container only (the reference tree does not exist on the GPU box):

    python scripts/make_reference_goldens.py

The goldens pin the branches the reference's doctests pin by shape only: 'mul' conditioning, the
ts_vad speaker-concat layer, num_averaged_permutations > 1, output_resolution 't', explicit_vad, TorchMFCC (2-D and
batch-coupled 3-D input), InstanceNorm / InstanceNorm_v2 along several axes.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import ref_stub as RS  # noqa: E402


def main():
    ns = RS.load()
    out = {}
    for name, kw in RS.CASES.items():
        torch.manual_seed(0)
        ref = ns.net.MaskEstimator_v2(aux_net=None, **kw).eval()
        for k, v in ref.state_dict().items():
            out[f"{name}/state/{k}"] = v.numpy()
        for batched in (False, True):
            xs, aux = RS.case_inputs(name, batched)
            np.random.seed(3)
            with torch.no_grad():
                o = ref(xs, RS.aux_argument(aux, batched))
            for f in ("mask", "logit", "embedding", "vad_mask", "vad_logit"):
                v = getattr(o, f)
                if v is not None:
                    out[f"{name}/{'batched' if batched else 'single'}/{f}"] = v.numpy()

    # TorchMFCC.stft_to_feature (feature_extractor_torchaudio.py:93-106) on a seeded complex STFT
    rng = np.random.RandomState(7)
    X = (rng.randn(2, 12, 513) + 1j * rng.randn(2, 12, 513)).astype(np.complex64)
    X *= np.exp(rng.uniform(-6, 2, size=(2, 12, 1))).astype(np.float32)  # frame levels spread over > 80 dB
    fe = ns.mfcc.TorchMFCC(size=1024, shift=256, window="hann")
    out["mfcc/X"] = X
    out["mfcc/single"] = fe.stft_to_feature(torch.tensor(X[0])).numpy()
    out["mfcc/batched"] = fe.stft_to_feature(torch.tensor(X)).numpy()  # top_db coupled over the batch

    # InstanceNorm / InstanceNorm_v2 (net.py:250-330)
    t = torch.tensor(rng.randn(3, 11, 7).astype(np.float32) * 3 + 1)
    out["norm/x"] = t.numpy()
    for dim in (-1, -2, 0):
        out[f"norm/v1/dim{dim}"] = ns.net.InstanceNorm(dim=dim)(t).numpy()
        out[f"norm/v1u/dim{dim}"] = ns.net.InstanceNorm(dim=dim, unbiased=True)(t).numpy()
    for md, nd in ((-1, -1), (-2, -2), (-2, -1), ((-2, -1), (-2, -1))):
        key = f"norm/v2/{md}/{nd}".replace(" ", "")
        if isinstance(md, tuple):
            # torch.linalg.norm over two dims is the Frobenius norm; np.sqrt(x.shape[tuple]) is not defined in the
            # reference (x.shape[(-2, -1)] raises), so only single axes are pinned
            continue
        out[key] = ns.net.InstanceNorm_v2(md, nd)(t).numpy()

    path = os.path.join(ROOT, "tests", "golden", "reference_net_goldens.npz")
    np.savez_compressed(path, **out)
    print(path, f"{os.path.getsize(path) / 1024:.0f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
