#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_beamformer.py tests/test_eval_driver.py -q -s -rf -m gpu --timeout 600 > gpurun_out/r2_step9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step9.log; tail -25 gpurun_out/r2_step9.log
