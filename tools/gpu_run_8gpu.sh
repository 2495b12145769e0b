#!/bin/bash
# 8-GPU check of both workloads (one process per GPU, torchrun)
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench ${N}gpu exit $?" > gpurun_out/summary_${N}gpu.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload train --steps 20 --warmup 3 > gpurun_out/bench_train_${N}gpu.json 2> gpurun_out/bench_train_${N}gpu.err
echo "bench train ${N}gpu exit $?" >> gpurun_out/summary_${N}gpu.txt
cat gpurun_out/summary_${N}gpu.txt; for f in bench_${N}gpu bench_train_${N}gpu; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d.get('clocks'))"; done; tail -n 3 gpurun_out/bench_${N}gpu.err gpurun_out/bench_train_${N}gpu.err
