#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 300 python tools/gpu_debug_train.py 4 160 > gpurun_out/debug_train.log 2>&1
echo "debug train exit $?" > gpurun_out/summary4.txt
timeout 120 ./tools/micro/tma_bw > gpurun_out/tma_bw.log 2>&1
echo "tma_bw exit $?" >> gpurun_out/summary4.txt
cat gpurun_out/summary4.txt; tail -50 gpurun_out/debug_train.log; cat gpurun_out/tma_bw.log
