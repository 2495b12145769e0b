#!/bin/bash
# Round-end evidence run, part 1: tests, smoke, both bench arms, auxiliary bench workloads.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/summary_final.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary_final.txt
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench.err
echo "bench ref exit $?" >> gpurun_out/summary_final.txt
timeout 600 python bench.py > gpurun_out/bench.json 2>> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary_final.txt
timeout 600 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --lanes 1 --no-cpu-baseline > gpurun_out/bench_lanes1.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2>> gpurun_out/bench.err
echo "bench train exit $?" >> gpurun_out/summary_final.txt
timeout 300 python bench.py --workload allpairs --steps 500 --warmup 20 > gpurun_out/bench_allpairs.json 2>> gpurun_out/bench.err
echo "bench allpairs exit $?" >> gpurun_out/summary_final.txt
cat gpurun_out/summary_final.txt; tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -5 gpurun_out/bench.err
for f in bench bench_bf16 bench_lanes1 bench_ref bench_train bench_allpairs; do python -c "
import json
d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'ms', round(d['ms_per_step'],4), d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))"; done
