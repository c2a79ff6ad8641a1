#!/bin/bash
mkdir -p gpurun_out
for c in 0 128 112 96; do
  TSSEP_GEMM_MAX_CTAS=$c timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_tmp.json > /dev/null 2> gpurun_out/r2_bench_tmp.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print("gemm ctas", $c, "ms", round(d["ms_per_step"], 1), "gemm", round(d["kernels"]["tssep_gemm"]["ms_per_step"], 1), "rec", round(d["kernels"]["tssep_blstm_recurrence_ts"]["ms_per_step"], 1), {k.split('[')[1][:9]: round(v["us_per_dependent_step"], 2) for k, v in d["roofline"]["launches"].items()}, d["clocks"]["sm_mhz"], d["clocks"]["sm_mhz_p10"])
PY
done
