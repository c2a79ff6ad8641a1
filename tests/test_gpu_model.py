"""GPU parity of the whole path (features -> mask estimator -> masking -> iSTFT) vs the oracle.

Stated tolerances (bf16 GEMM operands, fp32 accumulation / cell state, random init as in
SURVEY.md §8d): max|mask - mask_ref| <= 1e-3, |SDR delta| <= 0.05 dB on the separated signals,
LogMAE of the reference's end-to-end golden within 1e-3."""
import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O
from tests.util import make_pair, sdr_db

pytestmark = pytest.mark.gpu

MASK_TOL = 1e-3


def _ex(seed, aux_size, num_samples=None):
    e = O.dummy_example(seed, aux_size=aux_size, num_samples=num_samples)
    return e


def _me_kwargs(**kw):
    base = dict(idim=553, odim=513, units=40, projs=42, combination="mul", ts_vad=8, aux_net_output_size=513,
                num_averaged_permutations=2, output_resolution="tf")
    base.update(kw)
    return base


def _product_model(me, feature="concat"):
    from tssep_b200.data import DummyReader
    from tssep_b200.enhancer import Masking
    from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures, Log1pMaxNormAbsSTFT
    from tssep_b200.loss import LogMAE
    from tssep_b200.model import Model

    if feature == "concat":
        fe = ConcaternatedSTFTFeatures.new({
            "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC"},
            "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
            "size": 1024, "shift": 256, "window": "hann"})
    else:
        fe = Log1pMaxNormAbsSTFT(size=1024, shift=256, window="hann")
    return Model(fe=fe, reader=DummyReader(), mask_estimator=me, enhancer=Masking(), loss=LogMAE()).eval()


@pytest.mark.parametrize("resolution", ["tf", "t"])
def test_toy_configs_unbatched(cuda, resolution):
    """BASELINE configs[0]/[1]: toy TS-VAD ('t') and TS-SEP ('tf') on DummyReader validate examples."""
    ref, me = make_pair(_me_kwargs(output_resolution=resolution))
    model = _product_model(me)
    tables = O.MFCCTables()
    for seed in range(2):
        e = _ex(seed, 513)
        obs, aux = torch.tensor(e["observation"]), torch.tensor(e["auxInput"])
        np.random.seed(seed)
        want = O.forward_path(obs, aux, ref, feature="concat", tables=tables, window="hann")
        np.random.seed(seed)
        ex = {"observation": obs.to(cuda), "auxInput": aux.to(cuda), "reference_channel": 0}
        got = model(ex, with_time_estimate=True)
        assert got.mask.shape == (8, 1, 316, 513) and got.stft_estimate.dtype == torch.complex64
        assert got.embedding.shape == (8, 1, 513)
        assert (ex["Input"].cpu() - want.Input)[..., 40:].abs().max().item() < 2e-6
        dm = (got.mask.cpu() - want.mask).abs().max().item()
        dl = (got.logit.cpu() - want.logit).abs().max().item()
        print(f"seed {seed} res {resolution}: max|dmask| {dm:.2e} max|dlogit| {dl:.2e}")
        assert dm < MASK_TOL, dm
        assert (got.embedding.cpu() - want.embedding).abs().max().item() == 0.0
        dt = (got.time_estimate.cpu() - want.time_estimate).abs().max().item()
        assert dt < 5e-3, dt
        tgt = e["speaker_reverberation_early_ch0"]
        d_sdr = abs(sdr_db(got.time_estimate.cpu().numpy(), tgt) - sdr_db(want.time_estimate.numpy(), tgt))
        assert d_sdr <= 0.05, d_sdr


def test_reference_end_to_end_golden(cuda):
    """tssep/train/model.py:540-575: cat conditioning, units=10, projs=12, batch of 2 ->
    validate_LogMAE [0.74156505, 0.744494], ||Input|| 58.8257."""
    np.random.seed(0)
    torch.manual_seed(0)
    from tssep_b200.model import Model

    model = Model.new({"mask_estimator": {"units": 10, "projs": 12}}).eval()
    assert sum(p.numel() for p in model.parameters()) == 114038
    model = model.to(cuda)
    exs = [O.dummy_example(s) for s in (0, 1)]
    ex = {
        "observation": torch.tensor(np.stack([e["observation"] for e in exs]), device=cuda),
        "auxInput": torch.tensor(np.stack([e["auxInput"] for e in exs]), device=cuda),
        "reference_channel": 0,
    }
    tgt = torch.tensor(np.stack([e["speaker_reverberation_early_ch0"] for e in exs]), device=cuda)
    out = model(ex, with_time_estimate=True)
    assert out.mask.shape == (2, 8, 1, 316, 513) and out.stft_estimate.shape == (2, 8, 316, 513)
    loss = model.loss(out.time_estimate, tgt).cpu().numpy()
    print("LogMAE", loss)
    assert np.allclose(loss, [0.74156505, 0.744494], atol=1e-3), loss
    assert abs(torch.norm(ex["Input"]).item() - 58.8257) < 1e-2
    assert abs(torch.amax(ex["Input"].abs()).item() - 1.0) < 1e-6


@pytest.mark.parametrize("kw", [
    dict(combination="cat", ts_vad=False, num_averaged_permutations=1, aux_net_output_size=100, idim=513, units=10,
         projs=12),
    dict(combination="cat", ts_vad=8, num_averaged_permutations=3, aux_net_output_size=100, idim=513),
    dict(combination="mul", ts_vad=8, num_averaged_permutations=1, random_speaker_order=False, idim=513),
    dict(combination="mul", ts_vad=False, num_averaged_permutations=1, idim=513, nmask=2),
    dict(combination="mul", ts_vad=8, num_averaged_permutations=2, idim=513, explicit_vad=True),
])
@pytest.mark.parametrize("kernel", ["regs", "ts"])
def test_mask_estimator_variants_batched(cuda, monkeypatch, kw, kernel):
    """MaskEstimator_v2 alone, batched (B, T, F) input, all option combinations the reference exposes,
    through both recurrence kernels (tensor-memory tcgen05 / register-resident mma.sync)."""
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", kernel)
    ref, me = make_pair(_me_kwargs(**kw))
    B, T, K = 2, 120, 8
    A = kw.get("aux_net_output_size", 513)
    g = torch.Generator().manual_seed(5)
    xs = torch.rand((B, T, 513), generator=g)
    aux = torch.rand((B, K, A), generator=g)
    np.random.seed(7)
    with torch.no_grad():
        want = ref(xs, [[a for a in item] for item in aux])
    np.random.seed(7)
    with torch.no_grad():
        got = me(xs.to(cuda), [[a for a in item] for item in aux.to(cuda)])
    assert got.mask.shape == want.mask.shape
    assert (got.mask.cpu() - want.mask).abs().max().item() < MASK_TOL
    if want.logit is None:
        assert got.logit is None
        assert (got.vad_mask.cpu() - want.vad_mask).abs().max().item() < MASK_TOL
        assert (got.vad_logit.cpu() - want.vad_logit).abs().max().item() < 5e-3
    else:
        assert (got.logit.cpu() - want.logit).abs().max().item() < 5e-3
    assert (got.embedding.cpu() - want.embedding).abs().max().item() == 0.0


@pytest.mark.parametrize("kernel,g_dtype", [("regs", "bf16"), ("regs", "f32"), ("ts", "bf16")])
def test_full_size_dims_short_meeting(cuda, monkeypatch, kernel, g_dtype):
    """C3 dims (U=300, P=320, mul, ts_vad=8, R=2) on 20 s of audio; separate() with two meetings; both
    recurrence kernels, both storage types of the input projections."""
    monkeypatch.setenv("TSSEP_LSTM_KERNEL", kernel)
    monkeypatch.setenv("TSSEP_G_DTYPE", g_dtype)
    ref, me = make_pair(_me_kwargs(units=300, projs=320))
    model = _product_model(me)
    tables = O.MFCCTables()
    n = 16000 * 20
    exs = [_ex(s, 513, n) for s in range(2)]
    obs = torch.tensor(np.stack([e["observation"][0] for e in exs]))
    aux = torch.tensor(np.stack([e["auxInput"] for e in exs]))
    np.random.seed(3)
    wants = [O.forward_path(obs[i][None], aux[i], ref, feature="concat", tables=tables, window="hann") for i in range(2)]
    np.random.seed(3)
    got = model.separate(obs.to(cuda), aux.to(cuda), diarize=dict(threshold=0.5, median_width=11))
    for i in range(2):
        dm = (got.mask[i].cpu() - wants[i].mask).abs().max().item()
        print(f"meeting {i}: max|dmask| {dm:.2e}")
        assert dm < MASK_TOL, dm
        tgt = exs[i]["speaker_reverberation_early_ch0"]
        d_sdr = abs(sdr_db(got.time_estimate[i].cpu().numpy(), tgt) - sdr_db(wants[i].time_estimate.numpy(), tgt))
        assert d_sdr <= 0.05, d_sdr
        assert (got.stft_estimate[i].cpu() - wants[i].stft_estimate).abs().max().item() < 5e-2


@pytest.mark.parametrize("kw", [dict(), dict(ts_vad=False, num_averaged_permutations=1, combination="cat",
                                              aux_net_output_size=100)])
def test_separate_waves_equals_one_batch(cuda, monkeypatch, kw):
    """Model.separate_waves shares the row-light layers across waves and runs the rest per wave; every output
    must be identical to the one-batch result (the rows of a recurrence launch are independent)."""
    _, me = make_pair(_me_kwargs(**kw))
    model = _product_model(me)
    n = 16000 * 3
    exs = [_ex(s, kw.get("aux_net_output_size", 513), n) for s in range(5)]
    obs = torch.tensor(np.stack([e["observation"][0] for e in exs])).to(cuda)
    aux = torch.tensor(np.stack([e["auxInput"] for e in exs])).to(cuda)
    diar = dict(threshold=0.5, median_width=5)
    np.random.seed(7)
    want = model.separate(obs, aux, diarize=diar)
    np.random.seed(7)
    seen = []
    for lo, hi, got in model.separate_waves(obs, aux, wave=3, out_wave=2, diarize=diar):
        seen.append((lo, hi))
        assert torch.equal(got.mask, want.mask[lo:hi])
        assert torch.equal(got.logit, want.logit[lo:hi])
        assert torch.equal(torch.view_as_real(got.stft_estimate), torch.view_as_real(want.stft_estimate[lo:hi]))
        assert torch.equal(got.time_estimate, want.time_estimate[lo:hi])
        assert torch.equal(got.segments.segments, want.segments.segments[lo:hi])
        assert torch.equal(got.segments.counts, want.segments.counts[lo:hi])
    assert seen == [(0, 2), (2, 4), (4, 5)]


def test_state_dict_round_trip_and_cache_invalidation(cuda):
    ref, me = make_pair(_me_kwargs(idim=513))
    xs = torch.rand((60, 513), generator=torch.Generator().manual_seed(0))
    aux = [torch.rand(513) for _ in range(8)]
    np.random.seed(0)
    a = me(xs.to(cuda), [x.to(cuda) for x in aux]).mask.clone()
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.01)
    me.load_state_dict(ref.state_dict())  # must invalidate the packed-weight caches
    np.random.seed(0)
    b = me(xs.to(cuda), [x.to(cuda) for x in aux]).mask
    np.random.seed(0)
    with torch.no_grad():
        want = ref(xs, aux).mask
    assert (a - b).abs().max().item() > 1e-4
    assert (b.cpu() - want).abs().max().item() < MASK_TOL


def test_cpu_tensor_is_rejected(lib):
    from tssep_b200.feature_extractor import Log1pMaxNormAbsSTFT

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Log1pMaxNormAbsSTFT().stft(torch.zeros(2000))


def test_runs_on_a_device_that_is_not_current(lib):
    """Every C-ABI call makes the device of its stream current for the launch (tssep_b200/_lib.py::call): a model on
    cuda:1 works while cuda:0 is the current device.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    dev = torch.device("cuda:1")
    ref, me = make_pair(_me_kwargs(), device=dev)
    model = _product_model(me)
    e = _ex(0, 513)
    obs, aux = torch.tensor(e["observation"]), torch.tensor(e["auxInput"])
    np.random.seed(0)
    want = O.forward_path(obs, aux, ref, feature="concat", tables=O.MFCCTables(), window="hann")
    assert torch.cuda.current_device() == 0
    np.random.seed(0)
    got = model({"observation": obs.to(dev), "auxInput": aux.to(dev), "reference_channel": 0}, with_time_estimate=True)
    assert torch.cuda.current_device() == 0
    assert (got.mask.cpu() - want.mask).abs().max().item() < MASK_TOL
    with pytest.raises(RuntimeError, match="different devices|same device"):
        me(torch.zeros((10, 553), device="cuda:0"), [a for a in aux.to(dev)])


def test_steps_in_flight_on_several_streams_match_sequential_runs(cuda):
    """A serving loop keeps several steps in flight on as many streams (bench.py at <= 16 meetings per GPU), with the
    recurrence launches confined to a share of the SMs (rnnp.set_cta_budget): same results as one step at a time
    (bit for bit while the budget leaves the cluster shapes unchanged, as here; the shape with the epilogue-side G add
    rounds differently, hence the tolerance)."""
    from tssep_b200 import rnnp

    ref, me = make_pair(_me_kwargs(units=300, projs=320))
    model = _product_model(me)
    exs = [_ex(s, 513, num_samples=48_000) for s in range(3)]
    obs = [torch.tensor(np.stack([e["observation"]] * 2)).to(cuda) for e in exs]    # 2 meetings per step
    aux = [torch.tensor(np.stack([e["auxInput"]] * 2)).to(cuda) for e in exs]

    def run(i):
        np.random.seed(i)
        return [(o.mask.clone(), o.time_estimate.clone(), o.segments.segments.clone())
                for _, _, o in model.separate_waves(obs[i], aux[i], diarize=dict(threshold=0.5, median_width=5))]

    want = [run(i) for i in range(3)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda) for _ in range(3)]
    rnnp.set_cta_budget(98)
    try:
        got = [None] * 3
        for rep in range(2):      # second round: every stream's allocator pool is warm, kernels really overlap
            for i, st in enumerate(streams):
                st.wait_stream(torch.cuda.current_stream(cuda))
                with torch.cuda.stream(st):
                    got[i] = run(i)
        torch.cuda.synchronize()
    finally:
        rnnp.set_cta_budget(None)
    for w, g in zip(want, got):
        for (wm, wt, ws), (gm, gt, gs) in zip(w, g):
            assert (wm - gm).abs().max().item() <= 2e-5 and (wt - gt).abs().max().item() <= 2e-4
            if torch.equal(wm, gm):
                assert torch.equal(wt, gt) and torch.equal(ws, gs)
