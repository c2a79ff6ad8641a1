#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnp.py tests/test_classic_bf.py tests/test_wpe.py -q -rf -m gpu --timeout 600 > gpurun_out/r2_step15_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step15_tests.log; tail -15 gpurun_out/r2_step15_tests.log
{
echo "## release library: ping-pong shapes with G added by the epilogue"
timeout 600 python scripts/profile_rec.py --rows 104 208 416 512 832 --clusters 16 32 64 --tiles 2 --subs 1 2 --frames 6000 --reps 2
echo "## debug library"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 416 --clusters 16 32 64 --subs 2 --frames 4000 --reps 1
} > gpurun_out/r2_step15_microbench.txt 2>&1
grep -v Warn gpurun_out/r2_step15_microbench.txt | tail -40
