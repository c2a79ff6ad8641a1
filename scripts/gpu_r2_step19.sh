#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnp.py -q -rf -m gpu --timeout 600 -k tmem > gpurun_out/r2_step19_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step19_tests.log; tail -4 gpurun_out/r2_step19_tests.log
{
echo "## release library: ping-pong shapes with the proxy fence on a helper warp"
timeout 600 python scripts/profile_rec.py --rows 208 416 832 --clusters 16 32 64 --tiles 2 --subs 2 --frames 6000 --reps 2
echo "## debug library"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 208 --clusters 16 --tiles 2 --subs 2 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 416 --clusters 32 --tiles 2 --subs 2 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 832 --clusters 64 --tiles 2 --subs 2 --frames 4000 --reps 1
} > gpurun_out/r2_step19_microbench.txt 2>&1
grep -v Warn gpurun_out/r2_step19_microbench.txt | tail -20
