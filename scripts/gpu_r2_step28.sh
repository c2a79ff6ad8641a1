#!/bin/bash
# 2-GPU run: pcm16 test, bench with the e2e_pcm16 leg (32 meetings per GPU, one step in flight)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_postprocess.py -q -rf -m gpu --timeout 600 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_bench_c4_n2.json 2> gpurun_out/r2_bench_c4_n2.err
echo "bench n2 rc=$?"; tail -2 gpurun_out/r2_bench_c4_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_c4_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}); print({k: v for k, v in d["e2e"].items() if k != "note"}); print({k: v for k, v in d["e2e_pcm16"].items() if k != "note"}); print(d["config"].get("recurrence_waves"), d["config"]["steps_in_flight"], d["clocks"])
PY
