#!/bin/bash
# round-end rehearsal: what the driver runs (GPU tests, smoke, default bench with the driver's step counts, reference arm)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rf --timeout 1200 -x > gpurun_out/r2_final_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.log; tail -4 gpurun_out/r2_final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r2_final_ref.json
s0=$(date +%s)
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --profile-json gpurun_out/r2_bench_c4_final.json > gpurun_out/r2_final_bench.out 2> gpurun_out/r2_final_bench.err
echo "bench rc=$? wall=$(( $(date +%s) - s0 ))s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c4_final.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 1), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"].get("traffic"))
print("config3", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in (d["config3"] or {}).items() if k not in ("kernels", "workload")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "launches", d["gpu_launches"], d["clocks"])
PY
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --profile-json gpurun_out/r2_bench_c5_final.json > /dev/null 2> gpurun_out/r2_final_c5.err; echo "c5 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5_final.json"))
print("c5 value", d["value"], "ms", d["ms_per_step"])
PY
