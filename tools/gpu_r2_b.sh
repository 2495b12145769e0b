#!/bin/bash
# round-2 second GPU pass: all GPU tests (no -x), role traces of the halo kernel on the four stage shapes, new bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
for cfg in "80 32 64" "40 16 128" "20 8 256" "10 4 512"; do
  echo "== trace $cfg" >> gpurun_out/r2b_trace.txt
  timeout 120 python tools/micro/trace_halo.py $cfg 3 >> gpurun_out/r2b_trace.txt 2>&1
done
timeout 120 python tools/micro/time_halo.py >> gpurun_out/r2b_trace.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?" >> gpurun_out/r2b_bench.err
tail -15 gpurun_out/r2b_pytest.log; tail -5 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_bench.json
