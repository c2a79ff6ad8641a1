"""Evaluation driver (tssep_b200/eval.py): host orchestration around the inference path."""
import os
import wave

import numpy as np
import pytest
import torch

from oracle import tssep_oracle as O


def test_prepare_eval_dataset_batches_equal_lengths_longest_first():
    from tssep_b200.eval import prepare_eval_dataset

    exs = [{"example_id": f"m{i}", "audio_data": {"observation": np.zeros((1, n), np.float32)}, "auxInput": np.zeros((8, 4))}
           for i, n in enumerate([100, 300, 100, 300, 200, 300])]
    batches = prepare_eval_dataset(exs, batch_size=2)
    ids = [[e["example_id"] for e in b] for b in batches]
    assert ids == [["m1", "m3"], ["m5"], ["m4"], ["m0", "m2"]]
    assert all(e["reference_channel"] == 0 for b in batches for e in b)
    # two ranks: every meeting exactly once
    seen = sorted(e["example_id"] for r in range(2) for b in prepare_eval_dataset(exs, rank=r, world_size=2) for e in b)
    assert seen == sorted(e["example_id"] for e in exs)


def test_collate_fn_stacks_and_checks_reference_channel():
    from tssep_b200.eval import collate_fn

    exs = [{"observation": np.ones((1, 5), np.float32) * i, "auxInput": np.zeros((8, 3)), "reference_channel": 0,
            "example_id": str(i)} for i in range(3)]
    ex = collate_fn(exs)
    assert ex["observation"].shape == (3, 1, 5) and ex["auxInput"].shape == (3, 8, 3) and ex["reference_channel"] == 0
    exs[1]["reference_channel"] = 1
    with pytest.raises(AssertionError):
        collate_fn(exs)


def test_rttm_and_wav_writers(tmp_path):
    from tssep_b200.eval import Segment, rttm_lines, write_wav

    segs = [Segment("meetB", 1, 16000, 48000), Segment("meetA", 0, 8000, 12000), Segment("meetA", 3, 0, 4000)]
    assert rttm_lines(segs, 16000) == [
        "SPEAKER meetA 1 0.000 0.250 <NA> <NA> spk3 <NA> <NA>",
        "SPEAKER meetA 1 0.500 0.250 <NA> <NA> spk0 <NA> <NA>",
        "SPEAKER meetB 1 1.000 2.000 <NA> <NA> spk1 <NA> <NA>",
    ]
    path = os.path.join(tmp_path, "x.wav")
    sig = np.sin(np.arange(800) / 10).astype(np.float32) * 0.5
    write_wav(path, sig, 16000)
    with wave.open(path) as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 800)
        pcm = np.frombuffer(w.readframes(800), dtype="<i2")
    assert np.abs(pcm / 32767.0 - sig).max() < 1e-4


@pytest.mark.gpu
def test_eval_driver_end_to_end(cuda, tmp_path):
    """Toy TS-SEP model on four DummyReader meetings of two lengths: segments equal the diarization of the oracle's masks
    (away from the threshold), per-segment audio equals the cut of the oracle's separated signal, files are written."""
    from tests.util import make_pair
    from tssep_b200.data import DummyReader
    from tssep_b200.enhancer import Masking
    from tssep_b200.eval import EvalDriver
    from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures
    from tssep_b200.loss import LogMAE
    from tssep_b200.model import Model

    kw = dict(idim=553, odim=513, units=40, projs=42, combination="mul", ts_vad=8, aux_net_output_size=513,
              num_averaged_permutations=2)
    ref, me = make_pair(kw, device=cuda)
    fe = ConcaternatedSTFTFeatures.new({
        "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC"},
        "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
        "size": 1024, "shift": 256, "window": "hann"})
    model = Model(fe=fe, reader=DummyReader(aux_size=513), mask_estimator=me, enhancer=Masking(), loss=LogMAE()).eval().to(cuda)
    reader = DummyReader(aux_size=513)
    exs = [reader.get_example(s, num_samples=n, with_targets=False) for s, n in ((0, 48000), (1, 32000), (2, 48000), (3, 32000))]
    thr = 0.5002
    drv = EvalDriver(model, threshold=thr, median_width=5, out_dir=os.fspath(tmp_path), min_segment_samples=0)
    np.random.seed(0)
    got = drv.run(exs, device=cuda)
    assert sorted(got) == ["dummy_id_0", "dummy_id_1", "dummy_id_2", "dummy_id_3"]
    # oracle: the driver processes the long meetings first (ids 0, 2) and then the short ones (1, 3), one permutation
    # draw per meeting in that order
    np.random.seed(0)
    tables = O.MFCCTables()
    for sid in (0, 2, 1, 3):
        e = exs[sid]
        obs = torch.tensor(e["audio_data"]["observation"])
        want = O.forward_path(obs, torch.tensor(e["auxInput"]), ref, feature="concat", tables=tables, window="hann")
        _, sm, active, segs = O.diarize_reference(want.mask.numpy(), threshold=thr, median_width=5, num_samples=obs.shape[-1])
        mine = got[f"dummy_id_{sid}"]
        if (np.abs(sm - thr) > 2e-4).all():
            assert sorted((s.speaker, s.start, s.end) for s in mine) == sorted(
                (k, a, b) for k in range(8) for a, b in segs[k] if b > a)
        for s in mine[:5]:
            cut = want.time_estimate[s.speaker, s.start:s.end].numpy()
            assert np.abs(s.audio - cut).max() < 5e-3
            assert os.path.exists(s.path)
    rttm = open(os.path.join(tmp_path, "rank0.rttm")).read().strip().splitlines()
    assert len(rttm) == sum(len(v) for v in got.values()) and all(l.startswith("SPEAKER dummy_id_") for l in rttm)


@pytest.mark.gpu
def test_segment_wise_beamforming(cuda):
    """EvalDriver with a TorchBF segment enhancer on a 3-channel scene: every segment's audio equals the oracle
    beamformer run on the same frame range."""
    from tests.util import make_pair
    from tssep_b200.data import DummyReader
    from tssep_b200.enhancer import Masking, TorchBF
    from tssep_b200.eval import EvalDriver
    from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures
    from tssep_b200.loss import LogMAE
    from tssep_b200.model import Model

    kw = dict(idim=553, odim=513, units=40, projs=42, combination="mul", ts_vad=8, aux_net_output_size=513,
              num_averaged_permutations=1)
    ref, me = make_pair(kw, device=cuda)
    fe = ConcaternatedSTFTFeatures.new({
        "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC"},
        "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
        "size": 1024, "shift": 256, "window": "hann"})
    model = Model(fe=fe, reader=DummyReader(aux_size=513), mask_estimator=me, enhancer=Masking(), loss=LogMAE()).eval().to(cuda)
    e = DummyReader(aux_size=513).get_example(0, num_samples=40000, with_targets=False)
    rng = np.random.RandomState(0)
    obs1 = e["audio_data"]["observation"][0]
    obs = np.stack([obs1, np.roll(obs1, 3) * 0.8 + 0.05 * rng.randn(obs1.size), np.roll(obs1, -2) * 1.1 + 0.05 * rng.randn(obs1.size)])
    e["audio_data"]["observation"] = obs.astype(np.float32)
    drv = EvalDriver(model, threshold=0.5, median_width=5, segment_enhancer=TorchBF(), context=2, min_segment_samples=4000)
    np.random.seed(0)
    got = drv.run([e], device=cuda)["dummy_id_0"]
    assert len(got) >= 1
    np.random.seed(0)
    X = O.stft(torch.tensor(obs.astype(np.float32)), size=1024, shift=256, window="hann")
    inp = O.concat_feature(X[0], O.MFCCTables()).float()
    with torch.no_grad():
        mask = ref(inp, [a for a in torch.tensor(e["auxInput"])]).mask
    pad = 768
    for s in got[:3]:
        f0 = max(0, int(fe.sample_index_to_frame_index(s.start)) - 2)
        f1 = min(X.shape[-2], int(fe.sample_index_to_frame_index(s.end - 1)) + 1 + 2)
        est = O.torch_bf(mask[:, :, f0:f1], X[:, f0:f1], 0)[s.speaker].to(torch.complex64)
        y = O.istft(est, size=1024, shift=256, window="hann", fading=False).numpy()
        start = f0 * 256 - pad
        lo = s.start - start
        want = np.zeros(s.end - s.start, np.float32)
        a, b = max(lo, 0), min(lo + want.size, y.size)
        want[a - lo:b - lo] = y[a:b]
        scale = max(1e-6, np.abs(want).max())
        assert np.abs(s.audio - want).max() / scale < 1e-3
