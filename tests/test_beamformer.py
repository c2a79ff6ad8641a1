"""Mask-based MVDR beamformer (TorchBF 'mvdr_souden', tssep/train/enhancer.py:140-283).

* CPU: the oracle restatement equals the reference's own TorchBF (imported from /root/reference, complex128) exactly;
  committed goldens of that run travel to the GPU box;
* GPU: the product kernels (float64 covariance sums and solves, float32 application) against oracle and goldens.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O
from tests import ref_stub as RS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "reference_torchbf_goldens.npz")


def scene(seed, K, nmask, D, T, F, batch=None):
    """A seeded multi-channel scene: K point sources with random steering vectors plus noise, soft masks from the
    source powers (so that the covariance matrices are well conditioned but not trivial)."""
    rng = np.random.RandomState(seed)
    lead = () if batch is None else (batch,)
    S = rng.randn(*lead, K, T, F) + 1j * rng.randn(*lead, K, T, F)
    S *= (rng.rand(*lead, K, T, 1) > 0.5)                      # speakers come and go
    A = rng.randn(*lead, K, D, 1, F) + 1j * rng.randn(*lead, K, D, 1, F)
    N = 0.3 * (rng.randn(*lead, D, T, F) + 1j * rng.randn(*lead, D, T, F))
    Y = (A * S[..., :, None, :, :]).sum(axis=-4) + N
    p = np.abs(S) ** 2 + 1e-3
    tot = p.sum(axis=-3, keepdims=True) + 0.09
    m = p / tot
    if nmask == 2:
        masks = np.stack([m, 1.0 - m], axis=-3)
    else:
        masks = m[..., None, :, :]
    return torch.tensor(masks.astype(np.float32)), torch.tensor(Y.astype(np.complex128))


CASES = [dict(seed=0, K=2, nmask=2, D=6, T=40, F=17), dict(seed=1, K=3, nmask=1, D=4, T=33, F=9),
         dict(seed=2, K=8, nmask=1, D=7, T=50, F=12, batch=2), dict(seed=3, K=2, nmask=1, D=2, T=20, F=5)]


@pytest.mark.skipif(not RS.available(), reason="reference tree not present on this machine")
@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_torchbf(case):
    enh = RS.load_enhancer()
    masks, Y = scene(**case)
    for kw in (dict(), dict(masking=True, masking_eps=0.1)):
        want = enh.TorchBF("mvdr_souden", **kw)(masks, {"Observation": Y, "reference_channel": 0}, None)
        got = O.torch_bf(masks, Y, 0, **kw)
        assert want.dtype == torch.complex128 and got.shape == want.shape
        assert (want - got).abs().max().item() <= 1e-12 * want.abs().max().item()


def test_oracle_equals_committed_goldens():
    g = np.load(GOLDEN)
    for i, case in enumerate(CASES):
        masks, Y = scene(**case)
        got = O.torch_bf(masks, Y, 0).numpy()
        np.testing.assert_allclose(got, g[f"{i}/enh"], rtol=0, atol=1e-9 * np.abs(g[f"{i}/enh"]).max())
        got = O.torch_bf(masks, Y, 0, masking=True, masking_eps=0.1).numpy()
        np.testing.assert_allclose(got, g[f"{i}/enh_masking"], rtol=0, atol=1e-9 * np.abs(g[f"{i}/enh"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_product_torchbf_matches_oracle(cuda, i):
    from tssep_b200.enhancer import TorchBF

    g = np.load(GOLDEN)
    masks, Y = scene(**CASES[i])
    for kw, key in ((dict(), "enh"), (dict(masking=True, masking_eps=0.1), "enh_masking")):
        got = TorchBF("mvdr_souden", **kw)(masks.to(cuda), {"Observation": Y.to(cuda), "reference_channel": 0}, None)
        assert got.dtype == torch.complex128 and tuple(got.shape) == g[f"{i}/{key}"].shape
        err = np.abs(got.cpu().numpy() - g[f"{i}/{key}"]).max() / np.abs(g[f"{i}/{key}"]).max()
        print(f"case {i} {key}: max rel err {err:.2e}")
        assert err < 2e-5, err  # float32 application of a float64 beamformer to a complex64 copy of Y


@pytest.mark.gpu
def test_product_torchbf_long_meeting(cuda):
    """A 2-minute 7-channel scene (several time chunks of the covariance kernel, double atomics across them)."""
    from tssep_b200.enhancer import TorchBF

    masks, Y = scene(seed=5, K=8, nmask=1, D=7, T=7500, F=33)
    want = O.torch_bf(masks, Y, 0)
    got = TorchBF()(masks.to(cuda), {"Observation": Y.to(torch.complex64).to(cuda), "reference_channel": 0}, None)
    assert got.dtype == torch.complex64
    err = (got.cpu() - want).abs().max().item() / want.abs().max().item()
    print(f"long scene: max rel err {err:.2e}")
    assert err < 5e-5, err
