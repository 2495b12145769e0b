#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_halo_conv.py -m gpu -q -x > gpurun_out/pytest_halo.log 2>&1
echo "pytest halo exit $?" > gpurun_out/summary13.txt
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q > gpurun_out/pytest_fwd.log 2>&1
echo "pytest forward exit $?" >> gpurun_out/summary13.txt
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary13.txt
cat gpurun_out/summary13.txt; grep -v "^E    \|^E  +" gpurun_out/pytest_halo.log | tail -25; tail -5 gpurun_out/pytest_fwd.log; python -c "
import json
d=json.load(open('gpurun_out/bench.json')); print('infer', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['per_launch_ms'])"; tail -3 gpurun_out/bench.err
