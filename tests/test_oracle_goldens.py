"""Pins the CPU oracle against every golden the reference's own doctests hold for the hot path
(SURVEY.md §8c).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import tssep_oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_doctest_goldens.json")))


def test_log1p_maxnorm_small_array():
    g = GOLD["feature_extractor.py:194-196"]
    got = O.log1p_maxnorm_feature(np.array([[1, 5], [3 + 4j, -5]]))
    assert np.allclose(got, g["expected"], atol=1e-8)


def test_stft_plus_feature_statistics_bit_identical():
    g = GOLD["feature_extractor.py:197-202"]
    rng = np.random.RandomState(0)
    f = O.log1p_maxnorm_feature(O.stft(rng.uniform(0, 1, size=10_000)))  # class default: blackman 1024/256
    assert list(f.shape) == g["shape"]
    assert np.mean(f) == g["mean"] and np.min(f) == g["min"] and np.max(f) == g["max"] and np.std(f) == g["std"]


def test_frame_counts():
    for n, t in GOLD["frame_counts"]:
        assert O.num_frames(n, 1024, 256) == t


def _golden_model():
    np.random.seed(0)
    torch.manual_seed(0)
    net = O.OracleMaskEstimator(idim=513, odim=513, units=10, projs=12, aux_net_output_size=100, combination="cat")
    return net


def test_end_to_end_known_answer():
    """tssep/train/model.py:540-575: parameter count, LogMAE of the random-init model, feature norms."""
    g = GOLD["model.py:540-575"]
    net = _golden_model()
    assert sum(p.numel() for p in net.parameters()) == g["parameters"]
    exs = [O.dummy_example(s) for s in (0, 1)]
    obs = torch.tensor(np.stack([e["observation"] for e in exs]))
    aux = torch.tensor(np.stack([e["auxInput"] for e in exs]))
    tgt = torch.tensor(np.stack([e["speaker_reverberation_early_ch0"] for e in exs]))
    out = O.forward_path(obs, aux, net, feature="log1p", window="hann")
    assert list(out.mask.shape) == [2, 8, 1, 316, 513] and out.stft_estimate.dtype == torch.complex64
    loss = O.log_mae(out.time_estimate, tgt)
    assert np.allclose(loss.numpy(), g["validate_LogMAE"], atol=2e-6), loss
    assert abs(loss.sum().item() - g["loss"]) < 1e-4
    assert abs(torch.norm(out.Input).item() - g["input_norm"]) < 1e-3
    assert abs(torch.std(out.Input).item() - g["input_std"]) < 1e-4
    assert abs(torch.amax(out.Input.abs()).item() - g["input_amax"]) < 1e-6


def test_state_dict_key_contract():
    """Names listed at tssep/train/model.py:580-621 (42 tensors)."""
    net = _golden_model()
    keys = ["mask_estimator." + k for k in net.state_dict().keys()]
    assert keys == GOLD["model.py:580-621"]


def test_parameter_counts_ts_vad():
    """tssep/train/net.py:453-483: mul, ts_vad=4, idim=513."""
    g = GOLD["net.py:453-483"]
    net = O.OracleMaskEstimator(idim=513, combination="mul", ts_vad=4, aux_net_output_size=513)
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(net.pre_net.net[0]) == g["pre_net_lstm"] and n(net.pre_net.net[1]) == g["pre_net_linear"]
    assert n(net.post_net.birnn1.net[0]) == g["birnn1_lstm"] and n(net.post_net.birnn2.net[0]) == g["birnn2_lstm"]
    assert n(net.post_net.linear2) == g["linear2"]
    np.random.seed(0)
    obs = torch.tensor(np.random.randn(50, 513).astype(np.float32))
    aux = [torch.tensor(np.random.randn(513).astype(np.float32)) for _ in range(4)]
    out = net(obs, aux)
    assert list(out.mask.shape) == [4, 1, 50, 513]


def test_cat_shapes():
    """tssep/train/net.py:720-729."""
    net = O.OracleMaskEstimator(idim=257, ts_vad=False)
    obs = torch.tensor(np.random.randn(50, 257).astype(np.float32))
    aux = [torch.tensor(np.random.randn(100).astype(np.float32)) for _ in range(3)]
    out = net(obs, aux)
    assert list(out.mask.shape) == [3, 1, 50, 257] and list(out.embedding.shape) == [3, 1, 100]


def test_vad_staircase():
    """tssep/data.py:36-48."""
    vad = O.staircase_vad(71, 8)
    lines = ["".join("_#"[int(c)] for c in line) for line in vad]
    assert lines == GOLD["data.py:36-48"]
    assert vad.sum(axis=1).tolist() == [15] * 8


def test_loss_and_norm_goldens():
    torch.manual_seed(0)
    target = torch.rand((2, 10000))
    estimate = target + 0.5 * torch.rand((2, 10000))
    assert abs(O.log_mae(estimate, target).item() - GOLD["loss.py:223-233"][0]) < 1e-4
    estimate[1, :] = 0
    target[1, :] = 0
    assert abs(O.log_mae(estimate, target).item() - GOLD["loss.py:223-233"][1]) < 1e-4
    np.random.seed(0)
    t = torch.tensor(np.array([np.random.randn(10) * 5 - 5, np.random.randn(10) * 0.5 + 100]))
    want = torch.tensor(GOLD["net.py:293-303"], dtype=torch.float64)
    assert (O.instance_norm(t) - want).abs().max() < 1e-4
    assert (O.instance_norm_v2(t) - want).abs().max() < 1e-4


def test_istft_perfect_reconstruction_and_trial_average_identity():
    x = torch.randn(3, 20000, generator=torch.Generator().manual_seed(0))
    y = O.istft(O.stft(x, window="hann"), window="hann", num_samples=20000)
    assert (x - y).abs().max().item() < 2e-6
    # identical speakers + as many trials as speakers: every speaker's logit is the mean over ALL head
    # blocks of the same hidden sequence, hence all speakers must get the same logits
    torch.manual_seed(1)
    kw = dict(idim=33, odim=33, units=4, projs=5, combination="mul", ts_vad=3, aux_net_output_size=33,
              random_speaker_order=False)
    b = O.OracleMaskEstimator(num_averaged_permutations=3, **kw)
    xs = torch.rand(20, 33)
    e = torch.rand(33)
    with torch.no_grad():
        logit = b(xs, [e, e, e]).logit
    assert (logit - logit[:1]).abs().max().item() < 1e-6


def test_oracle_regression_fixture():
    """Guards the oracle itself against drift: a few mask values of the toy TS-SEP config."""
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_toy_tssep.npz"))
    torch.manual_seed(0)
    net = O.OracleMaskEstimator(idim=553, odim=513, units=40, projs=42, combination="mul", ts_vad=8,
                                aux_net_output_size=513, num_averaged_permutations=2).eval()
    e = O.dummy_example(0, aux_size=513)
    np.random.seed(0)
    out = O.forward_path(torch.tensor(e["observation"]), torch.tensor(e["auxInput"]), net, feature="concat",
                         tables=O.MFCCTables(), window="hann")
    idx = tuple(fx["index"].T)
    assert np.allclose(out.mask.numpy()[idx], fx["mask"], atol=2e-6)
    assert np.allclose(out.time_estimate.numpy()[:, ::4001], fx["time"], atol=2e-6)
    assert np.allclose(out.Input.numpy()[::37, ::29], fx["input"], atol=2e-5)


def test_diarize_reference_spec():
    mask = np.zeros((2, 1, 12, 4), dtype=np.float32)
    mask[0, 0, 2:5] = 1
    mask[0, 0, 7] = 1          # isolated frame: removed by a width-3 median
    mask[1, 0, :] = 1
    act, sm, active, seg = O.diarize_reference(mask, threshold=0.5, median_width=3, num_samples=2500)
    assert active[0].tolist() == [False, False, True, True, True] + [False] * 7
    s0, s1 = O.frame_to_sample_index(2, 1024, 256), O.frame_to_sample_index(5, 1024, 256)
    assert seg[0] == [(int(s0), int(s1))] and seg[1] == [(0, 2500)]
    # index mapping is monotone and consistent between the two directions
    s = np.arange(0, 5000, 7)
    f = O.sample_to_frame_index(s, 1024, 256)
    assert (np.diff(f) >= 0).all()
    assert (O.sample_to_frame_index(O.frame_to_sample_index(f, 1024, 256), 1024, 256) == f).all()
