#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_tma_store.py 2>&1 | grep -v Warn | tail -8
timeout 1800 python -m pytest tests -m gpu -q -rf --timeout 1200 -x > gpurun_out/r2_step33_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step33_tests.log; tail -4 gpurun_out/r2_step33_tests.log
for i in 1 2; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_tmp.json > /dev/null 2> gpurun_out/r2_bench_tmp.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print("ms", round(d["ms_per_step"], 1), {k: round(v["ms_per_step"], 1) for k, v in list(d["kernels"].items())[:3]}, {k.split('K=')[1][:12]: round(v["ms_per_step"], 1) for k, v in list(d["gemm_shapes"].items())[:5]})
PY
done
