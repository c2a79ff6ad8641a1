#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -q -rf -m gpu --timeout 600 > gpurun_out/r2_step27_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step27_tests.log; tail -6 gpurun_out/r2_step27_tests.log
