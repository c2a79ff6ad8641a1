#!/bin/bash
# ping-pong sub-batches at 8 / 16 rows: parity + microbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnp.py -q -rf -m gpu --timeout 600 -k "tmem" > gpurun_out/r2_step11_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step11_tests.log; tail -5 gpurun_out/r2_step11_tests.log
{
echo "## release library: sub-batch shapes"
timeout 600 python scripts/profile_rec.py --rows 32 104 208 256 416 512 --clusters 16 32 64 --subs 1 2 --frames 6000 --reps 2
echo "## debug library"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 208 416 --clusters 16 32 --subs 1 2 --frames 4000 --reps 1
} > gpurun_out/r2_step11_microbench.txt 2>&1
cat gpurun_out/r2_step11_microbench.txt | grep -v Warn | tail -50
