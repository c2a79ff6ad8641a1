#!/bin/bash
# 2-GPU run: new tests + the torchrun bench path (NCCL SegmentGather, strong scaling)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ipd_features.py tests/test_eval_driver.py -q -rf -m gpu --timeout 600 > gpurun_out/r2_step10_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_step10_tests.log; tail -8 gpurun_out/r2_step10_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_c4_n2.json 2> gpurun_out/r2_bench_c4_n2.err
echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/r2_bench_c4_n2.json; tail -5 gpurun_out/r2_bench_c4_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err
echo "ref n2 rc=$?"; tail -c 600 gpurun_out/r2_bench_ref_n2.json
timeout 300 python -m pytest tests -q -rf -m gpu -k "two_gpu or multi_gpu or device" --timeout 300 > gpurun_out/r2_step10_2gpu_tests.log 2>&1; tail -3 gpurun_out/r2_step10_2gpu_tests.log
