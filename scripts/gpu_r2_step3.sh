#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_rnnp.py tests/test_torch_ops.py -q -x -s -m gpu > gpurun_out/r2_rnnp3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_rnnp3.log
tail -4 gpurun_out/r2_rnnp3.log
( TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 1 8 16 32 --clusters 8 --tiles 1 2 --frames 8000
  TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 32 64 --clusters 16 --tiles 1 2 --frames 8000
  timeout 300 python scripts/profile_rec.py --rows 1 8 16 32 64 96 104 128 208 416 --clusters 0 --tiles 0 --frames 8000
) > gpurun_out/r2_rec_tiles.txt 2>&1
cat gpurun_out/r2_rec_tiles.txt
timeout 600 python scripts/profile_istft.py > gpurun_out/r2_istft.txt 2>&1; tail -5 gpurun_out/r2_istft.txt
