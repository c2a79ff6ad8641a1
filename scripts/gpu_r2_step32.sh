#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -rf --timeout 1200 -x > gpurun_out/r2_step32_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step32_tests.log; tail -4 gpurun_out/r2_step32_tests.log
timeout 300 python scripts/profile_istft.py --meetings 4 2>&1 | grep -v Warn | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_tmp.json > /dev/null 2> gpurun_out/r2_bench_tmp.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print("ms", round(d["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in d["kernels"].items()})
PY
