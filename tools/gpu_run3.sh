#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -s > gpurun_out/pytest_train.log 2>&1
echo "pytest train exit $?" > gpurun_out/summary3.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_train.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rest exit $?" >> gpurun_out/summary3.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary3.txt
cat gpurun_out/summary3.txt; tail -30 gpurun_out/pytest_train.log; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
