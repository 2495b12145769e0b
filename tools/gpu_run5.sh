#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/micro/umma_shift > gpurun_out/umma_shift.log 2>&1
echo "umma_shift exit $?" > gpurun_out/summary5.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_debug_train.py 4 160 > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/summary5.txt
cat gpurun_out/summary5.txt; cat gpurun_out/umma_shift.log; grep -v "^$" gpurun_out/sanitizer.log | head -60
