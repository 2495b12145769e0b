#!/bin/bash
# round-2 first GPU pass: state of HEAD on a B200 + the pair-mode micro-benchmarks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
for m in umma_pair umma_pair_pipe halo_pair umma_mnmajor_shift; do
  timeout 60 tools/micro/$m > gpurun_out/r2a_micro_$m.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_micro_$m.txt
done
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.log; for m in umma_pair umma_pair_pipe halo_pair umma_mnmajor_shift; do echo "== $m"; tail -12 gpurun_out/r2a_micro_$m.txt; done; cat gpurun_out/r2a_bench.json
