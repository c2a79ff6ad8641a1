#!/bin/bash
# two steps in flight at 8 / 16 meetings per GPU (one GPU, the per-rank workload of the 8- and 4-GPU runs)
mkdir -p gpurun_out
for m in 8 16; do
  timeout 900 python bench.py --meetings $m --steps 10 --warmup 3 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_m${m}_if2.json > gpurun_out/r2_bench_m${m}_if2.out 2> gpurun_out/r2_bench_m${m}_if2.err
  echo "m=$m rc=$?"; tail -2 gpurun_out/r2_bench_m${m}_if2.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_m${m}_if2.json"))
print("meetings", $m, "in flight", d["config"]["steps_in_flight"], "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1))
for k, v in d["roofline"]["launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
for k, v in list(d["kernels"].items())[:4]: print("  ", k, round(v["ms_per_step"], 2))
PY
done
