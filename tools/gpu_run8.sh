#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/summary8.txt
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
echo "bench train exit $?" >> gpurun_out/summary8.txt
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary8.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary8.txt
cat gpurun_out/summary8.txt; tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err; cat gpurun_out/smoke.log | tail -3
