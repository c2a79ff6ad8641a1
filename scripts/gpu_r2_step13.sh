#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnp.py tests/test_wpe.py -q -rf -m gpu --timeout 600 > gpurun_out/r2_step13_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step13_tests.log; tail -15 gpurun_out/r2_step13_tests.log
{
echo "## release library: one-tile shapes with and without ping-pong"
timeout 600 python scripts/profile_rec.py --rows 8 56 64 112 128 224 --clusters 8 16 32 --tiles 1 --subs 1 2 --frames 6000 --reps 2
echo "## two-tile reference points"
timeout 600 python scripts/profile_rec.py --rows 64 104 128 208 416 --clusters 8 16 32 --tiles 2 --subs 1 2 --frames 6000 --reps 2
echo "## debug library"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 112 --clusters 16 32 --tiles 1 --subs 2 --frames 4000 --reps 1
} > gpurun_out/r2_step13_microbench.txt 2>&1
grep -v Warn gpurun_out/r2_step13_microbench.txt | tail -70
