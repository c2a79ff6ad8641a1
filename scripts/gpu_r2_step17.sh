#!/bin/bash
mkdir -p gpurun_out
{
echo "## release library, unrolled MMA issue, P.G on the tensor core for every shape (GEPI = false)"
timeout 600 python scripts/profile_rec.py --rows 104 208 416 832 --clusters 16 32 64 --tiles 2 --subs 2 --frames 6000 --reps 2
} > gpurun_out/r2_step17_microbench.txt 2>&1
grep -v Warn gpurun_out/r2_step17_microbench.txt | tail -20
