"""GPU parity: diarization post-processing vs our written spec (oracle.diarize_reference).
Activity is a float mean (summation order differs): frames whose smoothed activity is within
1e-5 of the threshold are excluded from the bit-exact comparison, as BASELINE.json's north_star allows."""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("width", [1, 5, 11])
def test_diarize_matches_spec(cuda, width):
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT
    from tssep_b200.postprocess import diarize

    fe = Log1pMaxNormAbsSTFT(window="hann")
    rng = np.random.RandomState(0)
    K, T, F = 8, 700, 513
    base = (np.sin(np.arange(T)[None, :] / (9.0 + np.arange(K)[:, None])) > 0.2).astype(np.float32)
    mask = np.clip(base[:, None, :, None] * 0.8 + rng.uniform(0, 0.35, size=(K, 1, T, F)), 0, 1).astype(np.float32)
    n = 16000 * 11
    act, sm, active, segs = O.diarize_reference(mask, threshold=0.5, median_width=width, num_samples=n)
    d = diarize(torch.tensor(mask, device=cuda), fe, num_samples=n, threshold=0.5, median_width=width)
    assert np.abs(d.activity.cpu().numpy() - act).max() < 1e-5
    assert np.abs(d.smooth.cpu().numpy() - sm).max() < 1e-5
    safe = np.abs(sm - 0.5) > 1e-5
    assert (d.active.cpu().numpy().astype(bool) == active)[safe].all()
    if safe.all():
        assert d.to_lists() == segs
        assert d.counts.cpu().numpy().tolist() == [len(s) for s in segs]


def test_segments_edge_cases(cuda):
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT
    from tssep_b200.postprocess import diarize

    fe = Log1pMaxNormAbsSTFT(window="hann")
    T, F = 300, 513
    pat = np.zeros((4, T), dtype=np.float32)
    pat[1, :] = 1          # always active
    pat[2, 0] = pat[2, -1] = 1  # single frames at both ends
    pat[3, ::2] = 1        # 150 runs > max_segments
    mask = np.repeat(pat[:, None, :, None], F, axis=-1)
    _, _, _, segs = O.diarize_reference(mask, threshold=0.5, median_width=1, num_samples=70000)
    d = diarize(torch.tensor(mask, device=cuda), fe, num_samples=70000, threshold=0.5, median_width=1, max_segments=64)
    got = d.to_lists()
    assert got[0] == [] and got[1] == segs[1] and got[2] == segs[2]
    assert d.counts.cpu().numpy().tolist() == [0, 1, 2, 150]
    assert got[3] == segs[3][:64]


@pytest.mark.parametrize("size,shift", [(1024, 256), (256, 64)])
def test_activity_from_the_enhancement_kernel(cuda, size, shift):
    """Masking.apply(activity_out=...) reduces the mask rows it reads anyway to the frame activity: same values as
    the stand-alone reduction (fast 1024/256 kernel and the generic geometry)."""
    from tssep_b200.enhancer import Masking
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT
    from tssep_b200.postprocess import diarize

    fe = Log1pMaxNormAbsSTFT(size=size, shift=shift, window="hann")
    g = torch.Generator().manual_seed(0)
    B, K, T, F = 2, 5, 333, size // 2 + 1
    mask = torch.rand((B, K, 1, T, F), generator=g).to(cuda)
    X = torch.view_as_complex(torch.randn((B, 1, T, F, 2), generator=g)).to(cuda)
    act = torch.empty((B, K, T), dtype=torch.float32, device=cuda)
    Masking.apply(mask, X, 0, fe, want_estimate=True, want_time=True, activity_out=act)
    want = mask[:, :, 0].mean(-1)
    assert (act - want).abs().max().item() < 1e-6
    a = diarize(mask, fe, threshold=0.5, median_width=5)
    b = diarize(mask, fe, threshold=0.5, median_width=5, activity=act)
    assert (a.activity - b.activity).abs().max().item() < 1e-6


def test_pcm16_matches_numpy(cuda):
    """tssep_pcm16: round to nearest even, saturate (the conversion of eval.write_wav, on the device)."""
    from tssep_b200 import ops

    rng = np.random.RandomState(0)
    x = np.concatenate([rng.randn(100_003).astype(np.float32) * 0.4, np.array([1.5, -1.5, 0.5 / 32767, 1.5 / 32767, -2.5 / 32767, 0.0], np.float32)])
    want = np.clip(np.rint(x * np.float32(32767.0)), -32768, 32767).astype(np.int16)
    got = ops.pcm16(torch.as_tensor(x).to(cuda), 32767.0).cpu().numpy()
    assert got.dtype == np.int16 and np.array_equal(got, want)
    assert ops.pcm16(torch.zeros(0, device=cuda)).numel() == 0
