#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rnnp.py -q -x -s -m gpu > gpurun_out/r2_rnnp5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_rnnp5.log; tail -4 gpurun_out/r2_rnnp5.log
( TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 416 512 832 --clusters 64 --tiles 2 --frames 4000
  timeout 300 python scripts/profile_rec.py --rows 256 416 512 832 --clusters 32 64 --tiles 2 --frames 4000
  timeout 300 python scripts/profile_rec.py --rows 8 64 104 128 208 416 512 --clusters 0 --tiles 0 --frames 4000
) > gpurun_out/r2_rec_pingpong.txt 2>&1
cat gpurun_out/r2_rec_pingpong.txt
