#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --profile-json gpurun_out/r2_bench_c5_v2.json > gpurun_out/r2_bench_c5_v2.out 2> gpurun_out/r2_bench_c5_v2.err
echo "c5 rc=$?"; tail -3 gpurun_out/r2_bench_c5_v2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c5_v2.json"))
    print("c5 value", d["value"], "seg/s", d["segments_per_second"], "ms", d["ms_per_step"], "own ms", d["own_kernel_ms_per_step"], "loss", d["loss"])
    for k, v in d["kernels"].items(): print("  ", k, round(v["ms_per_step"], 2), v["launches_per_step"])
    for k, v in d["recurrence_launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no c5 json", e)
PY
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 > gpurun_out/r2_pytest_gpu14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu14.log; tail -5 gpurun_out/r2_pytest_gpu14.log
timeout 1500 python bench.py --profile-json gpurun_out/r2_bench_c4_v2.json > gpurun_out/r2_bench_c4_v2.out 2> gpurun_out/r2_bench_c4_v2.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_c4_v2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c4_v2.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("waves", d["config"]["recurrence_waves"], d["config"]["recurrence_capacity_rows"], "frac", d["roofline"]["frac"])
    for k, v in d["kernels"].items(): print("  ", k, round(v["ms_per_step"], 2), v["launches_per_step"])
    for k, v in d["roofline"]["launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
    print("config3", {k: v for k, v in (d["config3"] or {}).items() if k != "kernels"})
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
except Exception as e:
    print("no bench json", e)
PY
