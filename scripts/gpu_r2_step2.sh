#!/bin/bash
mkdir -p gpurun_out
# (1) where does the step go: MMA-warp counters, fence on/off, 4 columns per epilogue warp
( TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 8 64 --clusters 8 16 --ksplit 0 --frames 8000
  echo "--- NOFENCE (measurement only)"
  TSSEP_DEBUG_KNOBS=1 TSSEP_TS_NOFENCE=1 timeout 300 python scripts/profile_rec.py --rows 8 64 --clusters 8 16 --ksplit 0 --frames 8000
  echo "--- COLS=4 at 8 rows per cluster"
  TSSEP_DEBUG_KNOBS=1 TSSEP_TS_COLS=4 timeout 300 python scripts/profile_rec.py --rows 8 64 104 --clusters 8 --ksplit 0 --frames 8000
  echo "--- COLS=8 at 16 rows per cluster"
  TSSEP_DEBUG_KNOBS=1 TSSEP_TS_COLS=8 timeout 300 python scripts/profile_rec.py --rows 64 208 --clusters 16 --ksplit 0 --frames 8000
  echo "--- 416 rows"
  TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 416 --clusters 32 --ksplit 0 --frames 4000
) > gpurun_out/r2_rec_phases2.txt 2>&1
cat gpurun_out/r2_rec_phases2.txt
# (2) the whole GPU suite
timeout 2400 python -m pytest tests -m gpu -q -s -rf --timeout 1200 > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
grep -E "passed|failed|rc=|FAILED|max\|dmask|stress|full 10-min" gpurun_out/r2_pytest_gpu.log | tail -40
# (3) bench: BASELINE config 4 at N=1 (+ config 3, parity, cpu baseline)
timeout 1500 python bench.py --steps 3 --warmup 2 --profile-json gpurun_out/r2_bench_first.json > gpurun_out/r2_bench_first.out 2> gpurun_out/r2_bench_first.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_first.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_first.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("waves", d["config"]["recurrence_waves"], d["config"]["recurrence_capacity_rows"])
    for k, v in d["kernels"].items(): print("  ", k, round(v["ms_per_step"], 2), v["launches_per_step"])
    for k, v in d["roofline"]["launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
    print("config3", {k: v for k, v in (d["config3"] or {}).items() if k != "kernels"})
    print("parity", {k: v for k, v in (d["parity"] or {}).items() if k.startswith("max") or k.startswith("sdr")})
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
except Exception as e:
    print("no bench json", e)
PY
