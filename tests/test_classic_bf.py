"""Segment-wise evaluation beamformer ``ClassicBF_np`` + distortion masks (SURVEY.md §8f-1;
tssep/train/enhancer.py:370-590, enhancer_distortion_mask.py:9-55).

Pinned by the reference: the doctest values of ``SumCrossTalker`` / ``OneMinus``.  The beamforming vector itself comes
from the absent ``pb_bss`` package (restated; the oracle header says "parity unpinned"); the CPU tests check the
restatement against first principles on a two-speaker scene, the GPU tests compare the CUDA path with the oracle.
"""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O

M_DOC = np.array([[0, 0.2, 0.8, 1, 0], [0.1, 0, 0.5, 1, 0], [1, 0.1, 1, 0.5, 0]])[None, :, :, None]
SUM_DOC = np.array([[[0., 0.2, 0.8, 1., 0.], [0.1, 0., 0.5, 1., 0.], [1., 0.1, 1., 0.5, 0.]],
                    [[1.1, 0.1, 1.5, 1.5, 0.01], [1., 0.3, 1.8, 1.5, 0.01], [0.1, 0.2, 1.3, 2., 0.01]]])


def test_oracle_distortion_masks_match_reference_doctests():
    np.testing.assert_allclose(np.squeeze(O.sum_cross_talker(M_DOC, eps=0.01)), SUM_DOC, atol=1e-12)
    m = np.array([0, 0.5, 1])[None]
    np.testing.assert_array_equal(O.one_minus(m), [[0., 0.5, 1.], [1., 0.5, 0.]])


def test_product_distortion_masks_have_no_cpu_path():
    from tssep_b200.enhancer_distortion_mask import SumCrossTalker

    with pytest.raises(RuntimeError, match="CUDA"):
        SumCrossTalker()(torch.zeros((1, 3, 4, 2)))


@pytest.mark.gpu
def test_distortion_masks_match_reference_doctests(cuda):
    from tssep_b200.enhancer_distortion_mask import OneMinus, SumCrossTalker

    got = SumCrossTalker(eps=0.01)(M_DOC)                      # numpy in -> numpy out, computed on the device
    assert isinstance(got, np.ndarray)
    np.testing.assert_allclose(np.squeeze(got), SUM_DOC, atol=1e-12)
    m = np.array([0, 0.5, 1])[None]
    np.testing.assert_array_equal(OneMinus()(m), [[0., 0.5, 1.], [1., 0.5, 0.]])
    got = SumCrossTalker(eps=0.01)(torch.tensor(M_DOC).to(cuda))
    assert got.is_cuda
    np.testing.assert_allclose(np.squeeze(got.cpu().numpy()), SUM_DOC, atol=1e-12)
    rng = np.random.RandomState(0)
    big = rng.rand(1, 8, 33, 50)
    np.testing.assert_allclose(SumCrossTalker()(big), O.sum_cross_talker(big), rtol=0, atol=1e-13)   # summation order
    np.testing.assert_array_equal(OneMinus()(big), O.one_minus(big))


def toy_scene(seed=0, F=17, T=79, D=6):
    """Two speakers with different directions, partial overlap (the idea of tssep/data.py:155-240, own random draws)."""
    rng = np.random.RandomState(seed)
    doa = [np.exp(1j * np.zeros(D)), np.exp(1j * np.pi * np.array([0, 1, 0.5, 0.25, 0.75, 0][:D]))]
    dia = [[(0, 55)], [(45, T)]]
    src = []
    for k in range(2):
        s = (rng.randn(T, F) + 1j * rng.randn(T, F)) * np.sqrt(0.5)
        act = np.zeros(T, bool)
        for a, b in dia[k]:
            act[a:b] = True
        s[~act] = 0
        src.append(doa[k][:, None, None] * s[None])                   # (D, T, F)
    noise = 0.05 * (rng.randn(D, T, F) + 1j * rng.randn(D, T, F))
    obs = src[0] + src[1] + noise
    p = np.stack([np.abs(src[0][0]) ** 2, np.abs(src[1][0]) ** 2, np.abs(noise[0]) ** 2])
    mask = p / p.sum(0, keepdims=True)                                # (3, T, F) Wiener-like
    return obs, np.stack(src), mask, dia


def test_oracle_souden_vector_closed_form():
    """Rank-one target: w = Phi_n^-1 a conj(a_ref) / (a^H Phi_n^-1 a), distortionless towards the reference channel."""
    rng = np.random.RandomState(0)
    D, F = 6, 5
    a = rng.randn(F, D) + 1j * rng.randn(F, D)
    B = rng.randn(F, D, D) + 1j * rng.randn(F, D, D)
    phi_n = B @ np.conj(np.swapaxes(B, -1, -2)) + 0.1 * np.eye(D)
    phi_t = 2.5 * a[:, :, None] * np.conj(a[:, None, :])
    w = O.mvdr_souden_vector(phi_t, phi_n, ref_channel=2)
    ia = np.linalg.solve(phi_n, a[..., None])[..., 0]
    want = ia * np.conj(a[:, 2:3]) / np.einsum("fd,fd->f", np.conj(a), ia)[:, None]
    np.testing.assert_allclose(w, want, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.einsum("fd,fd->f", np.conj(w), a), a[:, 2], rtol=1e-9)


def test_oracle_classic_bf_structure():
    obs, src, mask, dia = toy_scene()
    est = O.classic_bf_np(mask[:-1, None], obs, dia)
    assert est.shape == (2, 79, 17) and est.dtype == np.complex128
    for k, (a, b) in enumerate([(0, 55), (45, 79)]):
        active = np.abs(est[k]).sum(-1) != 0
        assert active[a:b].all() and not active[:a].any() and not active[b:].any()      # enhancer.py:404-405
    # the plain-transpose symmetrisation of _get_psd leaves real statistics, so the beamformer is real
    Y = np.transpose(obs, (2, 0, 1))
    m = O.sum_cross_talker(np.transpose(mask[:-1, None], (1, 0, 3, 2))[:1])
    pt, pn = O._get_psd(m[0, 0, :, :55], Y[:, :, :55]), O._get_psd(m[1, 0, :, :55], Y[:, :, :55])
    assert np.abs(pt.imag).max() < 1e-15 and np.abs(pn.imag).max() < 1e-15   # rounding residue only
    w = O.mvdr_souden_vector(pt, pn, 0)
    np.testing.assert_allclose(est[0, :55], np.einsum("fd,fdt->tf", w.conj(), Y[:, :, :55]), atol=1e-12)
    # speaker 0 arrives in phase on every microphone: a real beamformer can pass it and cancel speaker 1, whose
    # contribution to the output must drop well below its level at the reference microphone
    Ys1 = np.transpose(src[1], (2, 0, 1))[:, :, 45:55]
    leak = np.einsum("fd,fdt->tf", w.conj(), Ys1)
    assert np.sum(np.abs(leak) ** 2) < 0.1 * np.sum(np.abs(src[1][0, 45:55]) ** 2)
    # ch0 / ch1 select a microphone
    np.testing.assert_array_equal(O.classic_bf_np(mask[:-1, None], obs, dia, bf="ch1")[0, :55], obs[1, :55])
    # masking multiplies with the (floored) target mask
    est_m = O.classic_bf_np(mask[:-1, None], obs, dia, masking=True, masking_eps=0.2)
    np.testing.assert_allclose(est_m[1, 45:], est[1, 45:] * np.maximum(mask[1, 45:], 0.2), atol=1e-12)


def test_factory_aliases():
    from tssep_b200.configurable import FACTORY_ALIASES

    for name in ("enhancer.ClassicBF_np", "enhancer_distortion_mask.SumCrossTalker", "enhancer_distortion_mask.OneMinus"):
        assert FACTORY_ALIASES["tssep.train." + name] == "tssep_b200." + name


class _AI:  # what the call needs of paderbox's ArrayInterval
    def __init__(self, ivs):
        self.normalized_intervals = tuple(ivs)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [{}, dict(masking=True, masking_eps=0.1), dict(mask_power=2), dict(bf="ch0"), dict(bf="ch1")])
def test_classic_bf_against_oracle(kw, cuda):
    from tssep_b200.enhancer import ClassicBF_np

    obs, src, mask, dia = toy_scene(seed=3, F=33, T=140, D=6)
    dia = [[(0, 55), (90, 140)], [(45, 100)]]
    want = O.classic_bf_np(mask[:-1, None], obs, dia, **kw)
    enh = ClassicBF_np(**kw)
    got = enh(mask[:-1, None], obs, [_AI(d) for d in dia], numpy_out=True)
    assert got.dtype == np.complex128 and got.shape == want.shape
    assert np.abs(got - want).max() < 2e-5 * np.abs(obs).max()
    assert np.all(got[0, 55:90] == 0) and np.all(got[1, :45] == 0)
    # tensors in, dict of segments out (numpy_out=False), boolean activity as diarization
    act = np.zeros((2, 140), bool)
    for k, d in enumerate(dia):
        for a, b in d:
            act[k, a:b] = True
    ret = enh(torch.as_tensor(mask[:-1, None]).to(cuda), torch.as_tensor(obs).to(cuda), [a for a in act])
    assert [sorted(r) for r in ret] == [[(0, 55), (90, 140)], [(45, 100)]]
    for k, r in enumerate(ret):
        for (a, b), v in r.items():
            assert v.shape == (b - a, 33) and v.is_cuda
            assert np.abs(v.cpu().numpy() - want[k, a:b]).max() < 2e-5 * np.abs(obs).max()


@pytest.mark.gpu
def test_classic_bf_edge_cases(cuda):
    """A speaker without any interval stays silent; one-frame and whole-signal intervals work; segment_wpe is applied
    to the interval only."""
    from tssep_b200.enhancer import WPE, ClassicBF_np

    obs, src, mask, _ = toy_scene(seed=5, F=9, T=120, D=6)
    dia = [[], [(0, 1), (10, 120)]]
    want = O.classic_bf_np(mask[:-1, None], obs, dia)
    got = ClassicBF_np()(mask[:-1, None], obs, dia, numpy_out=True)
    assert np.all(got[0] == 0) and np.abs(got - want).max() < 2e-5 * np.abs(obs).max()
    ret = ClassicBF_np()(mask[:-1, None], obs, dia)
    assert ret[0] == {} and sorted(ret[1]) == [(0, 1), (10, 120)] and ret[1][(0, 1)].shape == (1, 9)
    wpe_kw = dict(taps=2, delay=1, iterations=1)
    got = ClassicBF_np(segment_wpe=WPE(**wpe_kw))(mask[:-1, None], obs, [[(20, 100)], []], numpy_out=True)
    seg = O.wpe(obs[:, 20:100].astype(np.complex64), **wpe_kw)
    pad = np.zeros_like(obs)
    pad[:, 20:100] = seg
    want = O.classic_bf_np(mask[:-1, None], pad, [[(20, 100)], []])
    assert np.abs(got - want).max() < 5e-4 * np.abs(obs).max()


@pytest.mark.gpu
def test_classic_bf_whole_signal_and_wpe_hooks(cuda):
    from tssep_b200.enhancer import WPE, ClassicBF_np

    obs, src, mask, dia = toy_scene(seed=4, F=9, T=400, D=6)
    want = O.classic_bf_np(mask[:-1, None], obs, None, segment_bf=False)
    got = ClassicBF_np()(mask[:-1, None], obs, None, segment_bf=False, numpy_out=True)
    assert np.abs(got - want).max() < 2e-5 * np.abs(obs).max()
    # pre_wpe: the beamformer sees the dereverberated observation
    wpe_kw = dict(taps=3, delay=2, iterations=2)
    want = O.classic_bf_np(mask[:-1, None], O.wpe(obs.astype(np.complex64), **wpe_kw), [[(0, 200)], [(100, 400)]])
    got = ClassicBF_np(pre_wpe=WPE(**wpe_kw))(mask[:-1, None], obs, [[(0, 200)], [(100, 400)]], numpy_out=True)
    assert np.abs(got - want).max() < 5e-4 * np.abs(obs).max()
    with pytest.raises(NotImplementedError):
        ClassicBF_np(bf="wmwf")(mask[:-1, None], obs, None, segment_bf=False, numpy_out=True)
    with pytest.raises(AssertionError):
        ClassicBF_np()(mask[:-1, None], obs[:2], None, segment_bf=False, numpy_out=True)   # all channels must be loaded
