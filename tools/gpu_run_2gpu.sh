#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu exit $?" > gpurun_out/summary_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 20 --warmup 3 > gpurun_out/bench_train_2gpu.json 2> gpurun_out/bench_train_2gpu.err
echo "bench train 2gpu exit $?" >> gpurun_out/summary_2gpu.txt
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2>> gpurun_out/bench_2gpu.err
timeout 300 python bench.py --workload train --steps 20 --warmup 3 > gpurun_out/bench_train_1gpu.json 2>> gpurun_out/bench_2gpu.err
cat gpurun_out/summary_2gpu.txt; for f in bench_1gpu bench_2gpu bench_train_1gpu bench_train_2gpu; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'])"; done; tail -3 gpurun_out/bench_2gpu.err gpurun_out/bench_train_2gpu.err
