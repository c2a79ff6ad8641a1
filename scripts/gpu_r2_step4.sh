#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_vad_utils.py tests/test_gpu_postprocess.py tests/test_gpu_frontend.py -q -x -s -m gpu > gpurun_out/r2_small4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_small4.log; tail -4 gpurun_out/r2_small4.log
timeout 600 python scripts/profile_istft.py > gpurun_out/r2_istft2.txt 2>&1; tail -5 gpurun_out/r2_istft2.txt
( TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 8 32 --clusters 8 --tiles 1 2 --frames 8000
  TSSEP_DEBUG_KNOBS=1 timeout 300 python scripts/profile_rec.py --rows 32 64 --clusters 16 --tiles 1 2 --frames 8000
) > gpurun_out/r2_rec_tiles_phases.txt 2>&1
cat gpurun_out/r2_rec_tiles_phases.txt
