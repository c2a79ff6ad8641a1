#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnp.py tests/test_classic_bf.py -q -rf -m gpu --timeout 600 > gpurun_out/r2_step16_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step16_tests.log; tail -8 gpurun_out/r2_step16_tests.log
{
echo "## release library, unrolled MMA issue: auto"
timeout 600 python scripts/profile_rec.py --rows 1 8 56 64 104 128 208 416 512 832 --clusters 0 --tiles 0 --frames 6000 --reps 2
echo "## shapes"
timeout 600 python scripts/profile_rec.py --rows 56 --clusters 8 16 32 --tiles 1 --subs 1 --frames 6000 --reps 2
timeout 600 python scripts/profile_rec.py --rows 104 208 416 832 --clusters 8 16 32 64 --tiles 2 --subs 1 2 --frames 6000 --reps 2
echo "## debug library"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 56 --clusters 8 --tiles 1 --subs 1 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 104 --clusters 8 16 --tiles 2 --subs 1 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 416 --clusters 32 64 --subs 1 2 --frames 4000 --reps 1
} > gpurun_out/r2_step16_microbench.txt 2>&1
grep -v Warn gpurun_out/r2_step16_microbench.txt | tail -60
