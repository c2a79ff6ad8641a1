#!/bin/bash
# per-rank workloads of the strong-scaling run at 4 and 8 GPUs (16 / 8 meetings per GPU), on one GPU
mkdir -p gpurun_out
for m in 16 8; do
  timeout 900 python bench.py --meetings $m --steps 10 --warmup 3 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_m$m.json > gpurun_out/r2_bench_m$m.out 2> gpurun_out/r2_bench_m$m.err
  echo "m=$m rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_m$m.json"))
print("meetings", $m, "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1), "waves", d["config"]["recurrence_waves"])
for k, v in d["roofline"]["launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
for k, v in list(d["kernels"].items())[:4]: print("  ", k, round(v["ms_per_step"], 2))
PY
done
