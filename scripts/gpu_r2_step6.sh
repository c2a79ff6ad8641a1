#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -s -m gpu --timeout 300 > gpurun_out/r2_train6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_train6.log; tail -40 gpurun_out/r2_train6.log
