#!/bin/bash
# 8-GPU strong-scaling run of the default bench command (64 meetings in total, 8 per GPU)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_n8.json 2> gpurun_out/r2_bench_c4_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/r2_bench_c4_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_c4_n8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}); print(d["e2e"]); print(d["config"].get("recurrence_waves"), d["config"]["steps_in_flight"], d["clocks"]); print({k: v for k, v in (d.get("e2e_pcm16") or {}).items() if k != "note"})
for k, v in d["roofline"]["launches"].items(): print("  ", k, round(v["ms_per_step"], 2), round(v["us_per_dependent_step"], 3))
PY
