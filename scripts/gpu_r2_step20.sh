#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 -x > gpurun_out/r2_pytest_gpu20.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu20.log; tail -6 gpurun_out/r2_pytest_gpu20.log
timeout 300 python scripts/profile_istft.py --meetings 4 2>&1 | grep -v Warn | tail -6
timeout 1500 python bench.py --profile-json gpurun_out/r2_bench_c4_v3.json > gpurun_out/r2_bench_c4_v3.out 2> gpurun_out/r2_bench_c4_v3.err
echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_c4_v3.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c4_v3.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print("waves", d["config"]["recurrence_waves"], "frac", d["roofline"]["frac"])
    for k, v in d["kernels"].items(): print("  ", k, round(v["ms_per_step"], 2), v["launches_per_step"])
    for k, v in d["gemm_shapes"].items(): print("  ", k, round(v["ms_per_step"], 2), v["launches_per_step"])
    for k, v in d["roofline"]["launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
    print("config3", {k: v for k, v in (d["config3"] or {}).items() if k != "kernels"})
    print("parity", d.get("parity"))
except Exception as e:
    print("no json", e)
PY
