#!/bin/bash
mkdir -p gpurun_out
{
echo "## debug library: phase counters of the shipped shapes"
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 8 --clusters 8 --tiles 1 --subs 1 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 104 --clusters 8 --tiles 2 --subs 1 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 208 --clusters 16 --tiles 2 --subs 2 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 416 --clusters 32 --tiles 2 --subs 2 --frames 4000 --reps 1
TSSEP_DEBUG_KNOBS=1 timeout 600 python scripts/profile_rec.py --rows 832 --clusters 64 --tiles 2 --subs 2 --frames 4000 --reps 1
} > gpurun_out/r2_step18_phases.txt 2>&1
grep -v Warn gpurun_out/r2_step18_phases.txt | tail -20
