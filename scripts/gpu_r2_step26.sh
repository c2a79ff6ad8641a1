#!/bin/bash
mkdir -p gpurun_out
for f in 3 4; do
for m in 8 16; do
  TSSEP_BENCH_IN_FLIGHT=$f timeout 900 python bench.py --meetings $m --steps 12 --warmup 4 --no-cpu-baseline --no-config3 --no-parity --profile-json gpurun_out/r2_bench_tmp.json > /dev/null 2> gpurun_out/r2_bench_tmp.err
  echo "rc=$?"; tail -1 gpurun_out/r2_bench_tmp.err | cut -c1-200
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_tmp.json"))
print("in flight", $f, "meetings", $m, "value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1), {k.split('[')[1][:9]: round(v["us_per_dependent_step"], 2) for k, v in d["roofline"]["launches"].items()})
PY
done
done
