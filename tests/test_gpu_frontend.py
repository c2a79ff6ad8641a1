"""GPU parity: STFT / features / iSTFT kernels vs the CPU oracle (fp32; tolerances stated inline)."""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O

pytestmark = pytest.mark.gpu


def _sig(n, seed=0, lead=()):
    rng = np.random.RandomState(seed)
    return rng.uniform(-1, 1, size=(*lead, n)).astype(np.float32)


@pytest.mark.parametrize("n,size,shift,window", [(80000, 1024, 256, "hann"), (10000, 1024, 256, "blackman"),
                                                 (160, 64, 32, "hann"), (4097, 512, 128, "hann"), (300, 1024, 256, "hann")])
def test_stft_matches_oracle(cuda, n, size, shift, window):
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    fe = Log1pMaxNormAbsSTFT(size=size, shift=shift, window=window)
    x = _sig(n, lead=(3,))
    want = O.stft(torch.tensor(x), size=size, shift=shift, window=window)
    got = fe.stft(torch.tensor(x, device=cuda))
    assert got.shape == want.shape and got.dtype == torch.complex64
    scale = want.abs().max().item()
    # fp32 radix-2 FFT vs pocketfft: a few ulp of the largest bin
    assert (got.cpu() - want).abs().max().item() <= 2e-6 * scale + 1e-6


def test_stft_feature_golden(cuda):
    """Reference doctest golden tssep/train/feature_extractor.py:197-202 (float64 there; fp32 here)."""
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    fe = Log1pMaxNormAbsSTFT()
    rng = np.random.RandomState(0)
    f = fe(torch.tensor(rng.uniform(0, 1, size=10_000).astype(np.float32), device=cuda)).cpu().numpy().astype(np.float64)
    assert f.shape == (43, 513)
    assert abs(np.mean(f) - 0.03461471931132962) < 1e-6
    assert abs(np.std(f) - 0.051645387514742555) < 1e-6
    assert abs(np.max(f) - 1.0) < 1e-6
    assert abs(np.min(f) - 1.0003006801514706e-06) < 5e-7


@pytest.mark.parametrize("batched", [False, True])
def test_features_match_oracle(cuda, batched):
    from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures

    fe = ConcaternatedSTFTFeatures.new({
        "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC", "size": 1024, "shift": 256,
                "window": "hann"},
        "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
        "size": 1024, "shift": 256, "window": "hann"})
    assert fe.output_size == 553
    exs = [O.dummy_example(s, aux_size=513) for s in range(2 if batched else 1)]
    x = np.stack([e["observation"][0] for e in exs]) if batched else exs[0]["observation"][0]
    if batched:
        x[1] *= 0.01  # different level per item exercises the batch-coupled top_db quirk
    X = O.stft(torch.tensor(x), window="hann")
    want = O.concat_feature(X, O.MFCCTables())
    got = fe.stft_to_feature(X.to(cuda)).cpu()
    assert got.shape == want.shape
    # MFCC part: dB values O(100), fp32 matmul order differs -> 2e-3 abs; log1p part in [0,1] -> 2e-6
    assert (got[..., :40] - want[..., :40]).abs().max().item() < 2e-3
    assert (got[..., 40:] - want[..., 40:]).abs().max().item() < 2e-6


@pytest.mark.parametrize("n,size,shift", [(80000, 1024, 256), (1000, 64, 32), (5000, 512, 128)])
def test_istft_matches_oracle_and_round_trip(cuda, n, size, shift):
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    fe = Log1pMaxNormAbsSTFT(size=size, shift=shift, window="hann")
    x = _sig(n, lead=(2,))
    X = O.stft(torch.tensor(x), size=size, shift=shift, window="hann")
    rng = np.random.RandomState(1)
    Y = X * torch.tensor(rng.uniform(0, 1, size=X.shape).astype(np.float32))
    want = O.istft(Y, size=size, shift=shift, window="hann", num_samples=n)
    got = fe.istft(Y.to(cuda), num_samples=n).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < 5e-6
    back = fe.istft(fe.stft(torch.tensor(x, device=cuda)), num_samples=n).cpu()
    assert (back - torch.tensor(x)).abs().max().item() < 5e-6


def test_masking_and_fused_istft(cuda):
    from tssep_b200.enhancer import Masking
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    fe = Log1pMaxNormAbsSTFT(window="hann")
    x = _sig(20000, lead=(2, 1))
    X = O.stft(torch.tensor(x), window="hann")  # (2,1,T,F)
    rng = np.random.RandomState(2)
    mask = torch.tensor(rng.uniform(0, 1, size=(2, 4, 1, X.shape[-2], 513)).astype(np.float32))
    want_est = O.masking(mask, X, 0)
    want_time = O.istft(want_est, window="hann", num_samples=20000)
    est, time = Masking.apply(mask.to(cuda), X.to(cuda), 0, fe, want_estimate=True, want_time=True, num_samples=20000)
    assert est.dtype == torch.complex64 and est.shape == want_est.shape
    assert (est.cpu() - want_est).abs().max().item() == 0.0  # complex x real product is exact in fp32
    assert (time.cpu() - want_time).abs().max().item() < 5e-6
    # the reference-shaped call: unbatched, through __call__
    ex = {"Observation": X[0].to(cuda), "reference_channel": 0}

    class _M:
        pass

    m = _M()
    m.fe = fe
    got = Masking()(mask[0].to(cuda), ex, m)
    assert (got.cpu() - want_est[0]).abs().max().item() == 0.0


def test_instance_norm(cuda):
    from tssep_b200.net import InstanceNorm

    np.random.seed(0)
    t = torch.tensor(np.array([np.random.randn(50) * 5 - 5, np.random.randn(50) * 0.5 + 100]), dtype=torch.float32)
    got = InstanceNorm(dim=-1)(t.to(cuda)).cpu()
    want = O.instance_norm(t)
    assert (got - want).abs().max().item() < 1e-4


def test_instance_norms_match_reference_goldens(cuda):
    """InstanceNorm along any single axis (biased / unbiased) and InstanceNorm_v2 with equal and different axes against
    outputs of the reference's own classes (tests/golden/reference_net_goldens.npz, tssep/train/net.py:250-330)."""
    import os

    from tssep_b200.net import InstanceNorm, InstanceNorm_v2

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_net_goldens.npz"))
    x = torch.tensor(g["norm/x"]).to(cuda)
    for dim in (-1, -2, 0):
        assert np.abs(InstanceNorm(dim=dim)(x).cpu().numpy() - g[f"norm/v1/dim{dim}"]).max() < 2e-6
        assert np.abs(InstanceNorm(dim=dim, unbiased=True)(x).cpu().numpy() - g[f"norm/v1u/dim{dim}"]).max() < 2e-6
    for md, nd in ((-1, -1), (-2, -2), (-2, -1)):
        assert np.abs(InstanceNorm_v2(md, nd)(x).cpu().numpy() - g[f"norm/v2/{md}/{nd}"]).max() < 2e-6
