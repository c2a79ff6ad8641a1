"""Multi-channel / normalised feature variants: inter-channel phase differences, Log1pAbsIPDSTFT,
Log1pMaxNormAbsIPDSTFT, MVNLog1pAbsSTFT, NoFeatureSTFT.

The oracle restatement is pinned to the values the reference's own doctests hold (tssep/train/feature_extractor.py:40-56,
:85-95, :114-124); the CUDA path is compared with the oracle on seeded inputs, with the same NumPy seed so that both
draw the same partner channels."""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O


def _doctest_signal():
    np.random.seed(0)
    return np.ones([6, 4, 5]) * np.exp(1j * np.random.uniform(0, 2 * np.pi, [6, 1, 1])) * (np.arange(6)[:, None, None] + 1)


C_GOLD = np.array([0.81966208, 0.76070789, 0.93459697, 0.93459697, 0.72366352, 0.90670355])
S_GOLD = np.array([-0.57284734, 0.64909438, 0.35570844, -0.35570844, -0.69015296, -0.42176851])


def test_oracle_ipd_matches_reference_doctest():
    signal = _doctest_signal()   # the seed is consumed further by the channel shuffle, as in the doctest
    c, s = O.interchannel_phase_differences(signal)
    np.testing.assert_allclose(c[0], np.full((4, 5), 0.81966208), atol=5e-9)
    np.testing.assert_allclose(c[:, 0, 0], C_GOLD, atol=5e-9)
    np.testing.assert_allclose(s[:, 0, 0], S_GOLD, atol=5e-9)
    sig = O.interchannel_phase_differences(signal, concatenate=True)   # next shuffle of the same RNG stream
    np.testing.assert_allclose(sig[-1, 0, :], [6.0] * 5 + [0.81966208] * 5 + [0.57284734] * 5, atol=5e-9)
    np.testing.assert_allclose(sig[:, 0, 0], [1, 2, 3, 4, 5, 6])


def test_oracle_log1p_ipd_matches_reference_doctest():
    x = np.array([[1, 5], [3 + 4j, -5]])[:, None, :]
    got = np.squeeze(O.log1p_abs_ipd_feature(x), axis=-2)[0]
    np.testing.assert_allclose(got, [0.69314718, 1.79175947, 0.6, -1.0, -0.8, 0.0], atol=5e-9)


def test_oracle_mvn_matches_reference_doctest():
    got = O.mvn_log1p_abs_feature(np.array([[1, 5], [3 + 4j, -5]]))
    np.testing.assert_allclose(got, [[-0.54930614, 0.0], [0.54930614, 0.0]], atol=5e-9)


def test_output_sizes_and_factory_aliases():
    from tssep_b200 import feature_extractor as F
    from tssep_b200.configurable import FACTORY_ALIASES

    assert F.Log1pAbsIPDSTFT().output_size == 1539
    assert F.Log1pMaxNormAbsIPDSTFT().output_size == 1539
    assert F.MVNLog1pAbsSTFT().output_size == 513
    assert F.NoFeatureSTFT().output_size == 0
    with pytest.raises(AssertionError):
        F.Log1pMaxNormAbsIPDSTFT(output_size=513)
    with pytest.raises(AssertionError):
        F.NoFeatureSTFT(output_size=3)
    for name in ("Log1pAbsIPDSTFT", "Log1pMaxNormAbsIPDSTFT", "MVNLog1pAbsSTFT", "NoFeatureSTFT"):
        assert FACTORY_ALIASES["tssep.train.feature_extractor." + name] == "tssep_b200.feature_extractor." + name


# ---------------------------------------------------------------------------------------------------- CUDA path
@pytest.mark.gpu
def test_ipd_doctest_values_on_gpu():
    from tssep_b200 import feature_extractor as F

    signal = _doctest_signal()
    c, s = F.interchannel_phase_differences(signal)
    np.testing.assert_allclose(c[:, 0, 0], C_GOLD, atol=1e-6)
    np.testing.assert_allclose(s[:, 0, 0], S_GOLD, atol=1e-6)
    sig = F.interchannel_phase_differences(signal, concatenate=True)
    np.testing.assert_allclose(sig[-1, 0, :], [6.0] * 5 + [0.81966208] * 5 + [0.57284734] * 5, atol=1e-6)
    x = np.array([[1, 5], [3 + 4j, -5]])[:, None, :]
    got = np.squeeze(F.Log1pAbsIPDSTFT().stft_to_feature(x), axis=-2)[0]
    np.testing.assert_allclose(got, [0.69314718, 1.79175947, 0.6, -1.0, -0.8, 0.0], atol=1e-6)
    got = F.MVNLog1pAbsSTFT().stft_to_feature(np.array([[1, 5], [3 + 4j, -5]]))
    np.testing.assert_allclose(got, [[-0.54930614, 0.0], [0.54930614, 0.0]], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("lead,D,T,Fq", [((), 2, 7, 5), ((), 6, 43, 513), ((3,), 4, 19, 257), ((2, 2), 3, 5, 33)])
def test_ipd_against_oracle(lead, D, T, Fq, cuda):
    from tssep_b200 import feature_extractor as F

    rng = np.random.RandomState(D * 100 + T)
    X = (rng.randn(*lead, D, T, Fq) + 1j * rng.randn(*lead, D, T, Fq)).astype(np.complex64)
    np.random.seed(5)
    want_c, want_s = O.interchannel_phase_differences(X.astype(np.complex128))
    np.random.seed(5)
    got_c, got_s = F.interchannel_phase_differences(torch.as_tensor(X).to(cuda))
    np.testing.assert_allclose(got_c.cpu().numpy(), want_c, atol=2e-6)
    np.testing.assert_allclose(got_s.cpu().numpy(), want_s, atol=2e-6)
    # explicit partner channels
    sc = (np.arange(D) + 1) % D
    want_c, want_s = O.interchannel_phase_differences(X.astype(np.complex128), second_channel=sc)
    got_c, got_s = F.interchannel_phase_differences(torch.as_tensor(X).to(cuda), second_channel=sc)
    np.testing.assert_allclose(got_c.cpu().numpy(), want_c, atol=2e-6)
    np.testing.assert_allclose(got_s.cpu().numpy(), want_s, atol=2e-6)


@pytest.mark.gpu
def test_feature_classes_against_oracle(cuda):
    from tssep_b200 import feature_extractor as F

    rng = np.random.RandomState(3)
    wav = rng.randn(4, 10_000).astype(np.float32)
    X = O.stft(torch.as_tensor(wav), size=1024, shift=256, window="blackman").numpy().astype(np.complex128)
    for cls, fn in ((F.Log1pAbsIPDSTFT, O.log1p_abs_ipd_feature), (F.Log1pMaxNormAbsIPDSTFT, O.log1p_maxnorm_ipd_feature)):
        np.random.seed(11)
        want = fn(X)
        np.random.seed(11)
        got = cls()(torch.as_tensor(wav).to(cuda))
        assert tuple(got.shape) == (4, 43, 1539)
        # the STFT itself is f32 on the device: phase of near-empty bins is ill-conditioned, compare where |z| is not tiny
        np.testing.assert_allclose(got[..., :513].cpu().numpy(), want[..., :513], atol=2e-5)
        ok = np.minimum(np.abs(X), np.abs(X).min(axis=0, keepdims=True)) > 1e-2
        assert ok.mean() > 0.9
        d = np.abs(got[..., 513:].cpu().numpy() - want[..., 513:])
        assert d[np.concatenate([ok, ok], -1)].max() < 2e-3
    want = O.mvn_log1p_abs_feature(X[0])
    got = F.MVNLog1pAbsSTFT()(torch.as_tensor(wav[0]).to(cuda))
    np.testing.assert_allclose(got.cpu().numpy(), want, atol=2e-5)
    assert F.NoFeatureSTFT()(torch.as_tensor(wav[0]).to(cuda)).shape == (43, 0)
    with pytest.raises(NotImplementedError):
        F.MVNLog1pAbsSTFT(norm_vars=True).stft_to_feature(torch.as_tensor(X[0]).to(cuda))
