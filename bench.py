#!/usr/bin/env python
"""Throughput benchmark of the B200 TS-SEP inference hot path.

    python bench.py --gpus N --steps K --warmup W          # product arm (this repo's CUDA path)
    python bench.py --impl reference ...                    # the reference's CPU arithmetic (oracle) on host cores

Metric (BASELINE.json): audio-seconds processed per wall-second (RTF^-1), TS-SEP 8-speaker inference.

Workload = BASELINE config 4: a batch of 64 synthetic LibriCSS-shaped 10-min meetings, sharded data-parallel across
the N GPUs of one box (64/N meetings per GPU, `tssep_b200.dist.assign_meetings`), STRONG scaling.  One "step" = one
pass of the whole path (STFT -> features -> RNNP mask estimator -> mask x STFT -> iSTFT -> diarization) over the
rank's meetings; every output of the reference's ForwardOutput is materialised in HBM.  There is no data-path
collective; one NCCL all-gather of the segment tables per step (`tssep_b200.dist.SegmentGather`).  Rank 0 prints ONE
JSON line; at N = 1 it also carries BASELINE config 3 (ONE 10-min meeting alone: latency, audio-s/s) under
`config3`, a live parity check against the oracle under `parity`, and the CPU baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# GB-sized buffers of different shapes come and go every step: without expandable segments the caching
# allocator strands tens of GB in blocks the next request cannot use (must be set before CUDA initialises).
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
MODEL_KW = dict(idim=553, odim=513, layers=3, units=300, projs=320, combination="mul", ts_vad=8,
                aux_net_output_size=513, num_averaged_permutations=2, output_resolution="tf",
                random_speaker_order=True)
FE_CFG = {
    "fe1": {"factory": "tssep_b200.feature_extractor_torchaudio.TorchMFCC"},
    "fe2": {"factory": "tssep_b200.feature_extractor.Log1pMaxNormAbsSTFT"},
    "size": 1024, "shift": 256, "window": "hann",
}
WORKLOAD = "TS-SEP 8-speaker inference, LibriCSS-shaped synthetic 16 kHz meetings, U=300 P=320 mul ts_vad=8 R=2"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=["c4", "c5"],
                    help="c4 (default): BASELINE config 4, inference of 64 meetings; c5: BASELINE config 5, one training "
                         "forward + backward step on a batch of 60-s segments (extra mode, its own JSON line)")
    ap.add_argument("--segments", type=int, default=8, help="c5: 60-s segments per training step")
    ap.add_argument("--meetings", type=int, default=int(os.environ.get("TSSEP_BENCH_MEETINGS", 64)),
                    help="meetings of the whole job (BASELINE config 4: 64), dealt to the ranks")
    ap.add_argument("--seconds", type=float, default=600.0, help="length of every synthetic meeting")
    ap.add_argument("--wave-meetings", type=int, default=int(os.environ.get("TSSEP_BENCH_WAVE", 0)),
                    help="meetings per pass of the K-rows-per-meeting layers (0 = planned from the recurrence capacity)")
    ap.add_argument("--out-wave-meetings", type=int, default=int(os.environ.get("TSSEP_BENCH_OUT_WAVE", 13)),
                    help="meetings per output wave (head -> mask x STFT -> iSTFT -> diarization)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=60.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="skip the single-meeting (BASELINE config 3) measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the live parity check against the oracle")
    ap.add_argument("--profile-json", default=None, help="also write the line to this file")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
def synth_meeting(seed: int, num_samples: int, aux_size=513, device=None):
    """tssep/data.py:75-139 generator at arbitrary length (product-side restatement).  With ``device`` the
    24 float64 sinusoids are evaluated on the GPU (same formula and RNG draws; the CPU version needs ~4 s per
    10-min meeting), everything else follows tssep_b200.data.DummyReader.get_example."""
    from tssep_b200.data import DummyReader

    reader = DummyReader(sample_rate=SAMPLE_RATE, aux_size=aux_size)
    if device is None:
        ex = reader.get_example(seed, num_samples=num_samples, with_targets=False)
        return ex["audio_data"]["observation"][0].astype(np.float32), ex["auxInput"]
    K = reader.num_speakers
    rng = np.random.RandomState(seed)
    frequency = rng.randint(100, 7000, size=(3, K))
    vad = torch.as_tensor(reader._get_vad(num_samples, K), device=device)
    time_ax = torch.arange(num_samples, device=device, dtype=torch.float64) / SAMPLE_RATE
    obs = torch.zeros(num_samples, dtype=torch.float32, device=device)
    for k in range(K):
        acc = torch.sin(2 * np.pi * float(frequency[0, k]) * time_ax)
        acc += torch.sin(2 * np.pi * float(frequency[1, k]) * time_ax)
        acc += torch.sin(2 * np.pi * float(frequency[2, k]) * time_ax)
        obs += acc.to(torch.float32) * vad[k]
    noise = rng.rand(1, num_samples).astype(np.float32)
    obs = obs.cpu().numpy() + noise[0]
    aux = np.zeros((K, aux_size), dtype=np.float32)
    for spk, fs in enumerate(frequency.T):
        for f in fs:
            f = (f * aux_size) // 7001
            aux[spk, f:f + 2] = 1
    return obs, aux


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:  # one poller per job: rank 0 samples its own GPU
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_mhz_min": min(sm) if sm else None, "sm_mhz_p10": float(np.percentile(sm, 10)) if sm else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measured_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the recurrence, from the committed ncu --set full capture
    (profiles/r2_ncu_traffic.json, written by scripts/ncu_summary.py): {rows -> bytes per frame and row}."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else None


def bind_to_gpu_numa_node(local: int):
    """Pins this process (and the pinned host buffers it allocates afterwards, by first touch) to the CPUs next to its
    GPU: with several ranks on one host the device-to-host copies otherwise cross the socket interconnect."""
    try:
        props = torch.cuda.get_device_properties(local)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        cpus = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
            return {"cpus": cpus, "numa_node": int(node), "n_cpus": len(ids)}
    except Exception as e:  # containers without sysfs PCI topology: leave the affinity alone
        return {"error": f"{type(e).__name__}: {e}"}
    return None


# ----------------------------------------------------------------------------------------------
def oracle_setup(units, projs):
    from oracle import tssep_oracle as O

    torch.manual_seed(0)
    kw = {k: v for k, v in MODEL_KW.items() if k != "layers"}
    kw.update(units=units, projs=projs)
    net = O.OracleMaskEstimator(**kw).eval()
    return O, net, O.MFCCTables()


def time_oracle(sample_seconds: float, reps: int, warmup: int = 1, threads: int = None):
    """The reference's CPU arithmetic (oracle restatement: torch.nn.LSTM/Linear, torchaudio, torch.fft)
    on all host cores (or ``threads``), on a bounded sample of the workload (the path is linear in the audio length)."""
    O, net, tables = oracle_setup(MODEL_KW["units"], MODEL_KW["projs"])
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = threads or avail
    torch.set_num_threads(cores)
    n = int(sample_seconds * SAMPLE_RATE)
    obs, aux = synth_meeting(0, n)
    obs_t, aux_t = torch.tensor(obs)[None], torch.tensor(aux)
    times = []
    for i in range(warmup + reps):
        np.random.seed(0)
        t0 = time.perf_counter()
        O.forward_path(obs_t, aux_t, net, feature="concat", tables=tables, window="hann")
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sample_seconds / float(np.median(times)), cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    value, cores, times = time_oracle(args.cpu_sample_seconds, reps=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    host = os.cpu_count() or cores
    sample = (f"oracle restatement of the reference CPU path (torch.nn.LSTM/Linear, torchaudio, torch.fft; f32) on one "
              f"{args.cpu_sample_seconds:.0f}-s slice of a synthetic 8-speaker meeting per step (cost is linear in audio "
              f"length), torch threads={cores} of {host} host CPUs")
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": 1, "ms_per_step": float(np.median(times)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + f"; BASELINE config 4: {args.meetings} meetings of {args.seconds:.0f} s",
                   "sample_seconds": args.cpu_sample_seconds},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "host_cpus": host, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def build_product_model(device):
    from tssep_b200.data import DummyReader
    from tssep_b200.enhancer import Masking
    from tssep_b200.feature_extractor import ConcaternatedSTFTFeatures
    from tssep_b200.loss import LogMAE
    from tssep_b200.model import Model
    from tssep_b200.net import MaskEstimator_v2

    torch.manual_seed(0)
    me = MaskEstimator_v2.new(dict(MODEL_KW)).eval()
    fe = ConcaternatedSTFTFeatures.new(FE_CFG)
    model = Model(fe=fe, reader=DummyReader(aux_size=513), mask_estimator=me, enhancer=Masking(), loss=LogMAE())
    return model.eval().to(device)


def kernel_breakdown(timeline, steps):
    """Per entry point, and per (entry point, detail) for the calls that carry one (GEMM shapes, recurrence rows)."""
    agg, fine = {}, {}
    for name, s, e, detail in timeline:
        ms = s.elapsed_time(e)
        a = agg.setdefault(name, [0.0, 0])
        a[0] += ms
        a[1] += 1
        if detail is not None:
            f = fine.setdefault(f"{name}[{detail}]", [0.0, 0])
            f[0] += ms
            f[1] += 1
    out = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    detail = {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for k, v in sorted(fine.items(), key=lambda kv: -kv[1][0])}
    return out, detail


def parity_check(model, dev):
    """Live check of the product path against the oracle (the checker, not the thing measured): C3 dims, a 20-s
    meeting, same weights and permutation draws.  The full-length (T = 37 503) comparison is tests/test_gpu_full_size.py."""
    from oracle import tssep_oracle as O

    torch.manual_seed(0)
    kw = {k: v for k, v in MODEL_KW.items() if k != "layers"}
    ref = O.OracleMaskEstimator(**kw).eval()
    ref.load_state_dict({k: v.detach().cpu() for k, v in model.mask_estimator.state_dict().items()}, strict=True)
    n = 20 * SAMPLE_RATE
    e = O.dummy_example(3, aux_size=513, num_samples=n)
    obs, aux = torch.tensor(e["observation"]), torch.tensor(e["auxInput"])
    np.random.seed(11)
    want = O.forward_path(obs, aux, ref, feature="concat", tables=O.MFCCTables(), window="hann")
    np.random.seed(11)
    got = model.separate(obs.to(dev), aux[None].to(dev))
    tgt = e["speaker_reverberation_early_ch0"]

    def sdr(est):
        return float(10 * np.log10((tgt ** 2).sum() / (((est - tgt) ** 2).sum() + 1e-20) + 1e-20))

    return {
        "against": "oracle/tssep_oracle.py (pinned to the reference's doctest goldens and, exactly, to its own net.py / "
                   "rnnp.py / TorchMFCC: tests/test_oracle_vs_reference_net.py)",
        "case": "C3 dims (U=300 P=320 mul ts_vad=8 R=2), one 20-s meeting, random-init weights",
        "max_abs_dmask": float((got.mask[0].cpu() - want.mask).abs().max()),
        "max_abs_dlogit": float((got.logit[0].cpu() - want.logit).abs().max()),
        "max_abs_dtime": float((got.time_estimate[0].cpu() - want.time_estimate).abs().max()),
        "sdr_delta_db": abs(sdr(got.time_estimate[0].cpu().numpy()) - sdr(want.time_estimate.numpy())),
        "tolerance": {"max_abs_dmask": 1e-3, "sdr_delta_db": 0.05},
        "full_length": "tests/test_gpu_full_size.py::test_full_length_meeting_matches_oracle (one 10-min meeting, T = 37503)",
        "diarization": "parity unpinned: thresholding / median smoothing / segment extraction do not exist in the "
                       "reference; checked against this repo's own spec (oracle.diarize_reference)",
    }


def run_b200(args):
    from tssep_b200 import _lib
    from tssep_b200 import dist as tdist
    from tssep_b200 import ops as _ops

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    assert torch.cuda.is_available(), "bench.py needs CUDA (the product has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    _lib.load()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)

    n = int(args.seconds * SAMPLE_RATE)
    model = build_product_model(dev)
    K = 8

    # ---- BASELINE config 4: `meetings` equal meetings dealt to the ranks (strong scaling) ------------------------
    plan = tdist.assign_meetings([n] * args.meetings, world)
    my_ids = plan[rank]
    M = len(my_ids)
    assert M >= 1, f"rank {rank} has no meeting: --meetings {args.meetings} < --gpus {world}"
    # The K=8 speaker rows of every meeting advance together in the tensor-memory recurrence; pre_net (1 row per
    # meeting) and the TS-VAD layer (R rows) run once for all meetings.  Waves of the K-row layers: as few dependent
    # steps as possible given how many rows one wave of clusters holds at each cluster width.
    Up = 304
    cap = {w: _ops.recurrence_ts_capacity(Up, w) for w in (8, 16, 32, 64)}
    free_b, _ = torch.cuda.mem_get_info(dev)
    # memory model per 10-min meeting (calibrated on torch.cuda.memory_stats in round 1): 0.7 GB that lives for the whole
    # step (audio, STFT, pre_net rows, the step's separated audio and the serving loop's buffers), 2.05 GB while its wave
    # is in the K-rows-per-meeting layers (bf16 G + H + layer input), 2.5 GB while its output wave exists
    scale = args.seconds / 600.0
    light, heavy, outs = 0.7e9 * scale, 2.05e9 * scale, 2.5e9 * scale
    out_wave = min(M, max(1, args.out_wave_meetings))
    max_wave = int((0.90 * free_b - M * light - 0 * outs) / heavy)
    max_wave = max(1, min(M, max_wave))
    if args.wave_meetings > 0:
        waves = [min(args.wave_meetings, M - lo) for lo in range(0, M, args.wave_meetings)]
    else:
        waves = tdist.plan_recurrence_waves(M, K, cap, max_items=max_wave)
    while out_wave > 1 and M * light + out_wave * outs > 0.90 * free_b:
        out_wave -= 1

    meetings = [synth_meeting(i, n, device=dev) for i in my_ids]
    obs_host = torch.tensor(np.stack([m[0] for m in meetings])).pin_memory()
    aux_host = torch.tensor(np.stack([m[1] for m in meetings])).pin_memory()
    obs_dev, aux_dev = obs_host.to(dev), aux_host.to(dev)
    diar = dict(threshold=0.5, median_width=11, max_segments=256)
    # Steps in flight.  With few meetings per GPU a step is a chain of latency-bound recurrence launches that occupy
    # 20-80 of the 148 SMs, so a serving loop keeps two or three steps (independent batches) in flight on as many
    # streams (measured per-rank step at 8 meetings: 149 / 101 / 86 / 84 ms with 1 / 2 / 3 / 4 in flight; at 16 meetings
    # 194 / 151 / 145 ms with 1 / 2 / 3 and a collapse with 4 -- profiles/r2_steps_in_flight.txt), each
    # confined to two thirds of the SMs (rnnp.set_cta_budget; measured: 1/2, 2/3 and no bound are within 6 % of each
    # other): the recurrences of both run side by side.  With many meetings
    # per GPU the launches fill the device (and the memory), and steps run back to back.
    in_flight = int(os.environ.get("TSSEP_BENCH_IN_FLIGHT", 0)) or (3 if M <= 8 else (2 if M <= 16 else 1))
    if in_flight >= 2:
        from tssep_b200 import rnnp as _rnnp
        _rnnp.set_cta_budget(int(os.environ.get("TSSEP_BENCH_CTA_BUDGET", 0)) or
                             (2 * torch.cuda.get_device_properties(dev).multi_processor_count) // 3)
    step_streams = [torch.cuda.Stream(device=dev) for _ in range(in_flight)] if in_flight > 1 else [None]
    gathers = [tdist.SegmentGather(plan, rank, K, diar["max_segments"], dev) for _ in range(in_flight)]

    class StepOut:
        pass

    def step(obs, aux, on_wave=None, time_out=None, wave_plan=None, do_gather=True, slot=0):
        """One pass over the rank's meetings (Model.separate_waves).  The big per-wave outputs (mask, logit,
        stft_estimate) are fully written to HBM and released when the next output wave starts; time_estimate and the
        segment tables of all waves stay.  Speaker permutations are drawn on the host in meeting order."""
        np.random.seed(0)
        out = StepOut()
        segs, cnts = [], []
        for lo, hi, o in model.separate_waves(obs, aux, wave=waves if wave_plan is None else wave_plan, out_wave=out_wave,
                                              diarize=diar, time_out=time_out):
            segs.append(o.segments.segments)
            cnts.append(o.segments.counts)
            if on_wave is not None:
                on_wave(lo, hi, o.time_estimate)
            del o
        out.segments = torch.cat(segs) if len(segs) > 1 else segs[0]
        out.counts = torch.cat(cnts) if len(cnts) > 1 else cnts[0]
        if do_gather and world > 1:
            out.all_segments, out.all_counts = gathers[slot].start(out.segments, out.counts).result()
        return out

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    main0 = torch.cuda.current_stream(dev)

    def run_steps(k):
        """k device-resident steps, at most `in_flight` of them overlapping: step i runs on stream i % in_flight (every
        stream runs its steps in order); returns after making the main stream wait for all of them."""
        for i in range(k):
            if in_flight == 1:
                out = step(obs_dev, aux_dev)
            else:
                st = step_streams[i % in_flight]
                if i < in_flight:
                    st.wait_stream(main0)
                with torch.cuda.stream(st):
                    out = step(obs_dev, aux_dev, slot=i % in_flight)
            del out
        for st in step_streams:
            if st is not None:
                main0.wait_stream(st)

    run_steps(max(args.warmup, in_flight))  # also grows every stream's memory pool to its steady state
    barrier()

    # ---- device-resident timed region -----------------------------------------------------
    timeline = []
    _lib.set_timeline(timeline)
    l0 = _lib.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local, enabled=rank == 0) as clocks:
        barrier()
        ev0.record()
        run_steps(args.steps)
        ev1.record()
        barrier()
    _lib.set_timeline(None)
    launches = _lib.launch_count - l0
    ms = ev0.elapsed_time(ev1)
    breakdown, breakdown_detail = kernel_breakdown(timeline, args.steps)

    # ---- end-to-end: pinned host buffers in, separated audio + segments back on the host ----
    pinned = True
    try:
        time_host = torch.empty((M, K, n), dtype=torch.float32).pin_memory()
    except RuntimeError:  # not enough pinnable host memory: pageable destination, slower copies -- recorded in the line
        time_host = torch.empty((M, K, n), dtype=torch.float32)
        pinned = False
    seg_host = torch.empty((M, K, diar["max_segments"], 2), dtype=torch.int32).pin_memory()
    cnt_host = torch.empty((M, K), dtype=torch.int32).pin_memory()

    # The copies run on their own streams, so the D2H read of step i overlaps the compute of step i+1 (what a
    # serving loop does); every step's H2D and D2H lie inside the timed region, which ends when the last result has
    # landed in host memory.  One pinned host buffer suffices: a step's copies are done long before the next
    # step produces output (its output stages come last), and the copy stream is ordered anyway.
    copy_stream = torch.cuda.Stream(device=dev)   # device -> host
    in_stream = torch.cuda.Stream(device=dev)     # host -> device (separate, or it would queue behind the D2H)

    # Persistent device buffers of the serving loop (two of each input, used alternately): the inputs land in them by
    # H2D copy, the separated audio is written into time_buf by the enhancement kernel and read back by D2H copy.
    torch.cuda.empty_cache()
    n_in = max(2, in_flight)
    obs_bufs = [torch.empty_like(obs_dev) for _ in range(n_in)]
    aux_bufs = [torch.empty_like(aux_dev) for _ in range(n_in)]
    # one time buffer per step in flight (with one step in flight it is written only at the end of a step, after the
    # previous step's audio has left it)
    time_bufs = [torch.empty((M, K, n), dtype=torch.float32, device=dev) for _ in range(in_flight)]
    copied = [None] * in_flight  # event: the D2H copies of the previous step out of time_bufs[slot] are done
    in_free = [None] * n_in  # event: the step that read obs_bufs[j] / aux_bufs[j] is done

    pcm = {"on": False}
    pcm_host = None
    pcm_bufs = None

    def e2e_step(i):
        j = i % n_in
        slot = i % in_flight
        main_stream = step_streams[slot] if in_flight > 1 else main0   # the stream this step computes on
        ev_in = torch.cuda.Event()
        with torch.cuda.stream(in_stream):
            if in_free[j] is not None:
                in_stream.wait_event(in_free[j])
            obs_bufs[j].copy_(obs_host, non_blocking=True)
            aux_bufs[j].copy_(aux_host, non_blocking=True)
            ev_in.record(in_stream)
        main_stream.wait_event(ev_in)

        def time_out():  # called right before the first output wave of this step is written
            if copied[slot] is not None:
                main_stream.wait_event(copied[slot])  # the previous step's audio has left this time buffer
            return time_bufs[slot]

        def ship(lo, hi, time_estimate):  # D2H of a wave's separated audio as soon as it exists
            ev = torch.cuda.Event()
            ev.record(main_stream)
            if pcm["on"]:  # 16-bit PCM: converted on the device (tssep_pcm16), half the bytes over PCIe and into host memory
                with torch.cuda.stream(main_stream):
                    _ops.pcm16(time_estimate, 32767.0 / 8.0, out=pcm_bufs[slot][lo:hi])
                ev.record(main_stream)
            copy_stream.wait_event(ev)
            with torch.cuda.stream(copy_stream):
                if pcm["on"]:
                    pcm_host[lo:hi].copy_(pcm_bufs[slot][lo:hi], non_blocking=True)
                else:
                    time_host[lo:hi].copy_(time_estimate, non_blocking=True)

        with torch.cuda.stream(main_stream):
            out = step(obs_bufs[j], aux_bufs[j], on_wave=ship, time_out=time_out, slot=slot)
        ev_done = torch.cuda.Event()
        ev_done.record(main_stream)
        in_free[j] = ev_done  # the inputs of step i + n_in may overwrite obs_bufs[j] only after this step
        copy_stream.wait_event(ev_done)
        with torch.cuda.stream(copy_stream):
            out.segments.record_stream(copy_stream)
            out.counts.record_stream(copy_stream)
            seg_host.copy_(out.segments, non_blocking=True)
            cnt_host.copy_(out.counts, non_blocking=True)
            copied[slot] = torch.cuda.Event()
            copied[slot].record(copy_stream)

    # plain D2H bandwidth of this rank with every rank copying at once (what bounds the end-to-end number: 512 KB of
    # separated audio per audio-second)
    pm = min(M, out_wave)
    probe = torch.empty((pm, K, n), dtype=torch.float32, device=dev)
    barrier()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    time_host[:pm].copy_(probe, non_blocking=True)
    pe1.record()
    torch.cuda.synchronize()
    d2h_gbs = probe.numel() * 4 / (pe0.elapsed_time(pe1) / 1e3) / 1e9
    del probe

    for i in range(max(2, min(args.warmup, 3))):
        e2e_step(i)
    copy_stream.synchronize()
    barrier()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(main0)
    copy_stream.wait_event(ee0)
    in_stream.wait_event(ee0)
    for st in step_streams:
        if st is not None:
            st.wait_event(ee0)
    for i in range(args.steps):
        e2e_step(i)
    main0.wait_stream(copy_stream)
    for st in step_streams:
        if st is not None:
            main0.wait_stream(st)
    ee1.record(main0)
    barrier()
    e2e_ms = ee0.elapsed_time(ee1)

    # The same loop shipping the separated audio as 16-bit PCM (the format the evaluation driver writes to disk) instead
    # of float32: a clearly separate line -- the headline e2e above keeps the reference's float32 output.  Measured where
    # the host side of the device-to-host copies bounds the end-to-end number (several ranks copying at once).
    e2e_pcm_ms = None
    if world > 1:
        try:
            pcm_host = torch.empty((M, K, n), dtype=torch.int16).pin_memory()
        except RuntimeError:
            pcm_host = torch.empty((M, K, n), dtype=torch.int16)
        pcm_bufs = [torch.empty((M, K, n), dtype=torch.int16, device=dev) for _ in range(in_flight)]
        pcm["on"] = True
        for i in range(max(2, in_flight)):
            e2e_step(i)
        copy_stream.synchronize()
        barrier()
        pp0, pp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pp0.record(main0)
        for st in [copy_stream, in_stream] + [st for st in step_streams if st is not None]:
            st.wait_event(pp0)
        for i in range(args.steps):
            e2e_step(i)
        main0.wait_stream(copy_stream)
        for st in step_streams:
            if st is not None:
                main0.wait_stream(st)
        pp1.record(main0)
        barrier()
        e2e_pcm_ms = pp0.elapsed_time(pp1)
        pcm["on"] = False
        del pcm_bufs, pcm_host
    del obs_bufs, aux_bufs, time_bufs, time_host

    d2h_all = d2h_gbs
    if world > 1:
        t = torch.tensor([ms, e2e_ms, e2e_pcm_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, e2e_ms, e2e_pcm_ms = float(t[0]), float(t[1]), float(t[2])
        s = torch.tensor([d2h_gbs, float(pinned)], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM)
        d2h_all, pinned_ranks = float(s[0]), int(round(float(s[1])))
    else:
        pinned_ranks = int(pinned)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ---- BASELINE config 3: ONE meeting alone (latency), rank 0 at N = 1 ---------------------------------------------
    config3 = None
    if world == 1 and not args.no_config3:
        o1, a1 = obs_dev[:1], aux_dev[:1]
        h1 = torch.empty((1, K, n), dtype=torch.float32).pin_memory()
        for _ in range(2):
            step(o1, a1, wave_plan=[1], do_gather=False)
        torch.cuda.synchronize()
        tl3 = []
        _lib.set_timeline(tl3)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(args.steps):
            step(o1, a1, wave_plan=[1], do_gather=False)
        c1.record()
        torch.cuda.synchronize()
        _lib.set_timeline(None)
        ms3 = c0.elapsed_time(c1) / args.steps
        bd3, _ = kernel_breakdown(tl3, args.steps)
        # end to end: host audio in, separated audio back on the host, nothing overlapped (a single request)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            od, ad = obs_host[:1].to(dev, non_blocking=True), aux_host[:1].to(dev, non_blocking=True)
            for lo, hi, o in model.separate_waves(od, ad, wave=[1], out_wave=1, diarize=diar):
                h1.copy_(o.time_estimate, non_blocking=True)
                del o
        g1.record()
        torch.cuda.synchronize()
        e2e3 = g0.elapsed_time(g1) / args.steps
        rec3 = sum(v["ms_per_step"] for k, v in bd3.items() if k.startswith("tssep_blstm_recurrence"))
        T3 = model.fe.num_frames(n)
        config3 = {"workload": f"BASELINE config 3: ONE {args.seconds:.0f}-s meeting alone on one B200",
                   "ms": ms3, "value": args.seconds / (ms3 / 1e3), "unit": "audio-s/s",
                   "e2e_ms": e2e3, "e2e_value": args.seconds / (e2e3 / 1e3),
                   "recurrence_ms": rec3, "us_per_recurrent_step": rec3 * 1e3 / (4 * T3),
                   "kernels": bd3}

    parity = None
    if world == 1 and not args.no_parity:
        parity = parity_check(model, dev)

    audio_s = args.meetings * args.seconds * args.steps
    value = audio_s / (ms / 1e3)
    e2e_value = audio_s / (e2e_ms / 1e3)
    peaks = measured_peaks()
    T = model.fe.num_frames(n)
    # dominant kernel: the BLSTM recurrence (latency bound; expressed against the tensor peak as asked)
    rec_parts = {k: v for k, v in breakdown.items() if k.startswith("tssep_blstm_recurrence")}
    rec = {"ms_per_step": sum(v["ms_per_step"] for v in rec_parts.values()),
           "launches_per_step": sum(v["launches_per_step"] for v in rec_parts.values()) or 1}
    rec_rows = M * (1 + 8 + 8 + 2)  # pre_net, birnn0, birnn1, birnn2 (R=2) batch rows of this rank
    rec_cfg = "tanh.approx gates" if _ops.fast_math_default() else "exp-based gates"
    rec_flops = 2.0 * rec_rows * T * 2 * (4 * 300 * 300)
    rec_tflops = rec_flops / (rec["ms_per_step"] / 1e3) / 1e12 if rec["ms_per_step"] else 0.0
    gemm = breakdown.get("tssep_gemm", {"ms_per_step": 0.0})
    gemm_flops = M * (3.726e12 - 2.0 * 19 * T * 2 * 4 * 300 * 300)
    gemm_tflops = gemm_flops / (gemm["ms_per_step"] / 1e3) / 1e12 if gemm["ms_per_step"] else 0.0
    top = max(breakdown.items(), key=lambda kv: kv[1]["ms_per_step"])[0] if breakdown else None
    # algorithmic HBM bytes of the recurrence launches of one step: G is streamed once (bf16), H written once
    rec_bytes = rec_rows * T * (8 * Up * 2 + 2 * Up * 2)
    rec_launches = max(1.0, rec["launches_per_step"])
    rec_gbs = rec_bytes / (rec["ms_per_step"] / 1e3) / 1e9 if rec["ms_per_step"] else 0.0
    traffic = measured_traffic()
    traffic_per_launch, traffic_note = None, "no ncu --set full capture of the shipped recurrence committed yet"
    if traffic and "bytes_per_row_frame" in traffic:
        traffic_per_launch = traffic["bytes_per_row_frame"] * rec_rows * T / rec_launches
        traffic_note = traffic.get("note")
    roofline = {
        "kernel": "blstm_rec_ts_kernel", "bound": "tensor", "achieved": rec_tflops, "peak": peaks["bf16_tflops_sustained"],
        "unit": "TFLOP/s", "frac": rec_tflops / peaks["bf16_tflops_sustained"],
        "traffic": traffic_per_launch, "traffic_note": traffic_note,
        "algorithmic_flops_per_launch": rec_flops / rec_launches, "algorithmic_bytes_per_launch": rec_bytes / rec_launches,
        "avg_launch_ms": rec["ms_per_step"] / rec_launches,
        "peak_source": peaks["source"] + " (sustained)",
        "note": "the recurrence is bound by the latency of T dependent steps (per step: DSMEM exchange of h with st.async, "
                "tcgen05.mma with W_hh resident in tensor memory, gate math), not by the tensor pipe or HBM; "
                "see us_per_recurrent_step and launches",
        "hbm_GBps": rec_gbs, "hbm_frac": rec_gbs / peaks["hbm_gbs"],
        "us_per_recurrent_step": rec["ms_per_step"] * 1e3 / (rec_launches * T) if rec["ms_per_step"] else None,
        "dependent_steps_per_step": rec_launches * T,
        "launches": {k: dict(v, us_per_dependent_step=v["ms_per_step"] * 1e3 / (v["launches_per_step"] * T))
                     for k, v in breakdown_detail.items() if k.startswith("tssep_blstm_recurrence")},
        "gate_math": rec_cfg,
        "share_of_step": rec["ms_per_step"] / (ms / args.steps),
        "top_kernel_by_time": top,
    }
    roofline_gemm = {"kernel": "gemm_tc_kernel", "bound": "tensor", "achieved": gemm_tflops,
                     "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tflops / peaks["bf16_tflops_sustained"],
                     "share_of_step": gemm["ms_per_step"] / (ms / args.steps)}
    F = 513
    hbm = {}
    for name, nbytes in {
        "tssep_stft": M * (4 * n + 8 * T * F),
        "tssep_feature_stats": M * (8 * T * F + 4 * T * 40),
        "tssep_feature_write": M * (8 * T * F + 4 * T * 40 + 6 * T * 553),
        "tssep_mask_istft": M * (8 * T * F + 4 * 8 * T * F + 8 * 8 * T * F + 4 * 8 * n),
    }.items():
        if name in breakdown and breakdown[name]["ms_per_step"] > 0:
            gbs = nbytes / (breakdown[name]["ms_per_step"] / 1e3) / 1e9
            hbm[name] = {"achieved_GBps": gbs, "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes": nbytes}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, times = time_oracle(args.cpu_sample_seconds, reps=2)
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
               "sample": f"oracle (reference CPU arithmetic) on a {args.cpu_sample_seconds:.0f}-s slice of meeting 0, "
                         f"median of 2 after 1 warm-up, torch threads={cores} (the CPUs this process may run on; the host "
                         f"lists {os.cpu_count()}); cost is linear in audio length"}
        # the reference's documented setting (README.md:51-56, CI): one thread; a shorter slice keeps the run bounded
        s1 = max(5.0, args.cpu_sample_seconds / 2.0)
        v1, _, _ = time_oracle(s1, reps=2, warmup=1, threads=1)
        cpu["single_thread"] = {"value": v1, "unit": "audio-s/s", "cores": 1,
                                "sample": f"same, {s1:.0f}-s slice, median of 2 after 1 warm-up, torch threads=1"}
        torch.set_num_threads(cores)
        if config3 is not None:
            config3["speedup_vs_cpu_baseline"] = config3["value"] / v
            config3["e2e_speedup_vs_cpu_baseline"] = config3["e2e_value"] / v

    d2h_bytes = int(M * K * n * 4 + seg_host.numel() * 4 + cnt_host.numel() * 4)
    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"BASELINE config 4: {args.meetings} LibriCSS-shaped synthetic {args.seconds:.0f}-s 16 kHz "
                               f"meetings per step, sharded over {world} GPU(s) ({M} per GPU), 8 speakers, TS-SEP (U=300, "
                               "P=320, mul, ts_vad=8, 2 averaged permutations), random-init weights; every ForwardOutput "
                               "field + time_estimate + segments written to HBM (mask / logit / stft_estimate buffers are "
                               f"reused from one output wave of {out_wave} meetings to the next)",
                   "precision": "bf16 GEMM / recurrence operands, f32 accumulation, f32 cell state, f32 STFT / iSTFT",
                   "meetings_total": args.meetings, "meetings_per_gpu": M, "recurrence_waves": waves,
                   "steps_in_flight": in_flight,
                   "steps_in_flight_note": ("consecutive steps (independent batches of the same 64 meetings) overlap on "
                                            "as many streams, every recurrence launch confined to 2/3 of the SMs; ms_per_step = "
                                            "timed region / steps; the per-kernel times below overlap and add up to more "
                                            "than that" if in_flight > 1 else "steps run back to back on one stream"),
                   "recurrence_capacity_rows": cap, "meetings_per_output_wave": out_wave,
                   "meeting_seconds": args.seconds, "frames": T,
                   "l2": "inputs and intermediates (GBs per step) far exceed the 126 MB L2; no explicit flush",
                   "parallelism": f"dp{world} over meetings"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(obs_host.numel() * 4 + aux_host.numel() * 4),
                "d2h_bytes_per_step": d2h_bytes,
                "d2h_GBps_rank0": d2h_gbs, "d2h_GBps_all_ranks": d2h_all,
                "host_buffers_pinned_ranks": pinned_ranks, "numa_binding_rank0": numa,
                "note": "per-rank byte counts; bounded by the device-to-host copy of the separated audio (8 speakers x f32 = "
                        "512 KB per audio-second); copies overlap the next step, the last step's copy tail is inside the "
                        "timed region"},
        "e2e_pcm16": (None if e2e_pcm_ms is None else {
            "value": audio_s / (e2e_pcm_ms / 1e3), "unit": "audio-s/s", "ms_per_step": e2e_pcm_ms / args.steps,
            "d2h_bytes_per_step": int(M * K * n * 2),
            "note": "NOT the headline: the same end-to-end loop with the separated audio converted to 16-bit PCM on the device "
                    "(tssep_pcm16; what the evaluation driver writes to disk) before it is copied to the host -- half the "
                    "device-to-host bytes, which is what bounds e2e when several ranks copy into host memory at once"}),
        "gpu_launches": launches,
        "memory": {"max_reserved_GB": torch.cuda.max_memory_reserved(dev) / 1e9,
                   "max_allocated_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
                   "device_total_GB": torch.cuda.get_device_properties(dev).total_memory / 1e9,
                   "note": "rank 0, whole run (device-resident loop, end-to-end loops with their staging buffers, config 3)"},
        "roofline": roofline, "roofline_gemm": roofline_gemm, "hbm_kernels": hbm,
        "cpu_baseline": cpu,
        "config3": config3,
        "parity": parity,
        "kernels": breakdown, "gemm_shapes": {k: v for k, v in breakdown_detail.items() if k.startswith("tssep_gemm")},
    }
    print(json.dumps(line))
    if args.profile_json:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_json)), exist_ok=True)
        json.dump(line, open(args.profile_json, "w"), indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# BASELINE config 5: TS-SEP training forward / backward step, bf16 operands, synthetic 8-speaker 60-s segments, 1 GPU
def c5_batch(n_segments: int, seconds: float):
    from tssep_b200.data import DummyReader

    reader = DummyReader(sample_rate=SAMPLE_RATE, aux_size=513)
    n = int(seconds * SAMPLE_RATE)
    exs = [reader.get_example(s, num_samples=n, with_targets=True) for s in range(n_segments)]
    obs = torch.tensor(np.stack([e["audio_data"]["observation"] for e in exs]).astype(np.float32))       # (B, 1, n)
    tgt = torch.tensor(np.stack([e["audio_data"]["speaker_reverberation_early_ch0"] for e in exs]).astype(np.float32))
    aux = torch.tensor(np.stack([e["auxInput"] for e in exs]).astype(np.float32))
    return obs, aux, tgt


def run_c5(args):
    from tssep_b200 import _lib

    assert args.gpus == 1 and int(os.environ.get("WORLD_SIZE", 1)) == 1, "config 5 is a single-GPU configuration"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _lib.load()
    seconds = 60.0
    model = build_product_model(dev).train()
    obs, aux, tgt = c5_batch(args.segments, seconds)
    obs, aux, tgt = obs.to(dev), aux.to(dev), tgt.to(dev)

    def step():
        np.random.seed(0)
        for p in model.parameters():
            p.grad = None
        ex = {"observation": obs, "auxInput": aux, "reference_channel": 0}
        out = model(ex)
        loss = model.loss(out.time_estimate, tgt).sum()
        loss.backward()
        return loss

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    timeline = []
    _lib.set_timeline(timeline)
    l0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clocks:
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
    _lib.set_timeline(None)
    ms = e0.elapsed_time(e1) / args.steps
    breakdown, detail = kernel_breakdown(timeline, args.steps)
    own_ms = sum(v["ms_per_step"] for v in breakdown.values())
    T = model.fe.num_frames(int(seconds * SAMPLE_RATE))
    grads = sum(int(p.grad is not None) for p in model.parameters())
    line = {
        "metric": "audio_seconds_per_second", "value": args.segments * seconds / (ms / 1e3), "unit": "audio-s/s",
        "segments_per_second": args.segments / (ms / 1e3), "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"BASELINE config 5: TS-SEP training forward + backward step (LogMAE, no optimizer) on a batch "
                               f"of {args.segments} synthetic 8-speaker {seconds:.0f}-s segments, U=300 P=320 mul ts_vad=8 R=2, "
                               "random-init weights, 1 B200",
                   "precision": "bf16 GEMM / recurrence operands, f32 accumulation, f32 cell state, f32 parameter gradients",
                   "segments": args.segments, "frames": T, "parameters_with_gradient": grads,
                   "library_parts": "weight-gradient GEMMs, the head Linear and the elementwise glue run through torch "
                                    "(cuBLAS / ATen); the recurrences (forward and BPTT), the input / projection GEMMs and "
                                    "their data gradients, STFT / iSTFT and its adjoint are this repo's kernels"},
        "clocks": clocks.summary(), "loss": float(loss), "gpu_launches": _lib.launch_count - l0,
        "own_kernel_ms_per_step": own_ms, "kernels": breakdown,
        "recurrence_launches": {k: dict(v, us_per_dependent_step=v["ms_per_step"] * 1e3 / (v["launches_per_step"] * T))
                                for k, v in detail.items() if "recurrence" in k},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = time_oracle_c5(seconds)
    print(json.dumps(line))
    if args.profile_json:
        json.dump(line, open(args.profile_json, "w"), indent=1)


def time_oracle_c5(seconds: float, reps: int = 2):
    """The reference's training arithmetic (oracle restatement, torch autograd, f32) on the host cores: forward + backward
    of ONE 60-s segment per repetition."""
    O, net, tables = oracle_setup(MODEL_KW["units"], MODEL_KW["projs"])
    net.train()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    obs, aux, tgt = c5_batch(1, seconds)
    n = obs.shape[-1]
    times = []
    for i in range(1 + reps):
        np.random.seed(0)
        net.zero_grad()
        t0 = time.perf_counter()
        X = O.stft(obs[0], size=1024, shift=256, window="hann")
        inp = O.concat_feature(X[0], tables).float()
        out = net(inp, [a for a in aux[0]])
        est = O.masking(out.mask, X, 0)
        t_est = O.istft(est, size=1024, shift=256, window="hann", num_samples=n)
        O.log_mae(t_est, tgt[0]).backward()
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
    v = seconds / float(np.median(times))
    return {"value": v, "unit": "audio-s/s", "segments_per_second": v / seconds, "cores": cores, "host_cpus": os.cpu_count(),
            "kind": "port", "sample": f"oracle forward + backward (torch autograd, f32) of one {seconds:.0f}-s segment, median "
                                      f"of {reps} after 1 warm-up, torch threads={cores}"}


def run_reference_c5(args):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cpu = time_oracle_c5(60.0, reps=max(1, args.steps))
    line = {"impl": "reference", "metric": "audio_seconds_per_second", "value": cpu["value"], "unit": "audio-s/s",
            "segments_per_second": cpu["segments_per_second"], "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": 1,
            "ms_per_step": 60.0 / cpu["value"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 5: TS-SEP training forward + backward step, one 60-s segment per step"},
            "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse_args()
    if a.config == "c5":
        run_reference_c5(a) if a.impl == "reference" else run_c5(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
