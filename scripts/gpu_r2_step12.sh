#!/bin/bash
# interleaved row-tile MMAs: parity + microbench; WPE tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnp.py tests/test_wpe.py tests/test_ipd_features.py -q -rf -m gpu --timeout 600 > gpurun_out/r2_step12_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step12_tests.log; tail -15 gpurun_out/r2_step12_tests.log
{
echo "## release library, auto"
timeout 600 python scripts/profile_rec.py --rows 8 64 104 128 208 416 512 --clusters 0 --tiles 0 --frames 6000 --reps 2
echo "## release library: shapes"
timeout 600 python scripts/profile_rec.py --rows 104 208 416 512 --clusters 8 16 32 64 --subs 1 2 --frames 6000 --reps 2
timeout 600 python scripts/profile_rec.py --rows 8 56 --clusters 8 --tiles 1 --subs 1 --frames 6000 --reps 2
echo "## debug library"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 104 --clusters 8 --subs 1 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 56 --clusters 8 --tiles 1 --subs 1 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 416 --clusters 16 32 64 --subs 1 2 --frames 4000 --reps 1
} > gpurun_out/r2_step12_microbench.txt 2>&1
grep -v Warn gpurun_out/r2_step12_microbench.txt | tail -60
