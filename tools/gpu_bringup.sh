#!/bin/bash
# First GPU bring-up: each case under its own timeout so a hung kernel cannot eat the whole lease.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for c in s1 s2 s3 s4 e2 e3 e4; do
  timeout 150 python tools/gpu_debug_conv.py $c > gpurun_out/conv_$c.log 2>&1
  echo "case $c exit $?" >> gpurun_out/summary.txt
done
timeout 150 python tools/gpu_debug_conv.py s2 bf16 > gpurun_out/conv_s2_bf16.log 2>&1
echo "case s2 bf16 exit $?" >> gpurun_out/summary.txt
timeout 200 python tools/gpu_debug_forward.py > gpurun_out/forward.log 2>&1
echo "forward exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 30 gpurun_out/conv_s1.log gpurun_out/forward.log
