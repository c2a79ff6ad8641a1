#!/bin/bash
# Runs the GPU parity suites one file per process, each under its own timeout, so a hung
# kernel cannot wedge the whole call.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  echo "=== $f"
  timeout 600 python -m pytest "$f" -m gpu -q -x --timeout 300 -s > "gpurun_out/$name.log" 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')"
done
