#!/bin/bash
mkdir -p gpurun_out
for m in 8 16 64; do
  timeout 900 python bench.py --meetings $m --steps 20 --warmup 5 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_tmp.json > /dev/null 2> gpurun_out/r2_bench_tmp.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print("meetings", $m, "in flight", d["config"]["steps_in_flight"], "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["ms_per_step"], 1), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["memory"].items() if k != "note"})
PY
done
