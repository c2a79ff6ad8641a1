#!/bin/bash
# round-2 first GPU visit: new recurrence kernel correctness + step latency
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_rnnp.py -q -x -s > gpurun_out/r2_rnnp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_rnnp.log
tail -5 gpurun_out/r2_rnnp.log
timeout 300 python scripts/profile_rec.py --rows 1 8 64 104 --clusters 8 16 --ksplit 0 1 --frames 8000 > gpurun_out/r2_rec_small.txt 2>&1
timeout 300 python scripts/profile_rec.py --rows 208 256 416 512 --clusters 16 32 --ksplit 0 1 --frames 4000 > gpurun_out/r2_rec_large.txt 2>&1
TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 8 104 --clusters 8 16 --ksplit 0 1 --frames 8000 > gpurun_out/r2_rec_phases.txt 2>&1
TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 416 --clusters 32 --ksplit 0 1 --frames 4000 >> gpurun_out/r2_rec_phases.txt 2>&1
cat gpurun_out/r2_rec_small.txt gpurun_out/r2_rec_large.txt gpurun_out/r2_rec_phases.txt
