#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/gpu_debug_train.py 4 160 > gpurun_out/debug_train.log 2>&1
echo "debug train exit $?" > gpurun_out/summary6.txt
timeout 300 python tools/gpu_debug_train.py 16 160 > gpurun_out/debug_train16.log 2>&1
echo "debug train B16 exit $?" >> gpurun_out/summary6.txt
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1
echo "pytest train exit $?" >> gpurun_out/summary6.txt
cat gpurun_out/summary6.txt; tail -50 gpurun_out/debug_train.log; tail -8 gpurun_out/debug_train16.log; tail -30 gpurun_out/pytest_train.log
