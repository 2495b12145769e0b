#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/summary11.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary11.txt
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
echo "bench train exit $?" >> gpurun_out/summary11.txt
cat gpurun_out/summary11.txt; tail -4 gpurun_out/pytest_gpu.log; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print('infer', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['per_launch_ms'])
d=json.load(open('gpurun_out/bench_train.json')); print('train', d['value'], d['ms_per_step'])"
