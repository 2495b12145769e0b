#!/bin/bash
# 2-GPU check of both workloads (one process per GPU, torchrun, NCCL only in the training allreduce / timing barrier)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu exit $?" > gpurun_out/summary_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 20 --warmup 3 > gpurun_out/bench_train_2gpu.json 2> gpurun_out/bench_train_2gpu.err
echo "bench train 2gpu exit $?" >> gpurun_out/summary_2gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_2gpu.json 2> gpurun_out/bench_ref_2gpu.err
echo "bench ref 2gpu exit $?" >> gpurun_out/summary_2gpu.txt
cat gpurun_out/summary_2gpu.txt; for f in bench_2gpu bench_train_2gpu bench_ref_2gpu; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d.get('clocks'))"; done; tail -3 gpurun_out/bench_2gpu.err gpurun_out/bench_train_2gpu.err
