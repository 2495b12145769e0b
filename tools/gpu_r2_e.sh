#!/bin/bash
mkdir -p gpurun_out
for sk in 0 1; do
for cfg in "40 16 128" "20 8 256" "10 4 512"; do
  echo "== trace sk=$sk $cfg" >> gpurun_out/r2e_trace.txt
  DSK_STREAM_K=$sk timeout 120 python tools/micro/trace_halo.py $cfg 3 >> gpurun_out/r2e_trace.txt 2>&1
done
done
cat gpurun_out/r2e_trace.txt
