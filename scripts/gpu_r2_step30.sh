#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_rnnp.py tests/test_gpu_model.py tests/test_gpu_train.py -q -rf -m gpu --timeout 900 -x > gpurun_out/r2_step30_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step30_tests.log; tail -4 gpurun_out/r2_step30_tests.log
timeout 300 python scripts/profile_gemm.py --meetings 8 2>&1 | grep -v Warn | tee gpurun_out/r2_gemm_microbench.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_tmp.json > /dev/null 2> gpurun_out/r2_bench_tmp.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print("ms", round(d["ms_per_step"], 1), "gemm", round(d["kernels"]["tssep_gemm"]["ms_per_step"], 1), "rec", round(d["kernels"]["tssep_blstm_recurrence_ts"]["ms_per_step"], 1))
for k, v in d["gemm_shapes"].items(): print("  ", k, round(v["ms_per_step"], 2))
PY
