#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" > gpurun_out/summary2.txt
timeout 600 python bench.py --steps 200 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary2.txt
timeout 300 python bench.py --steps 200 --warmup 10 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
echo "bench ref exit $?" >> gpurun_out/summary2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_umma|conv1_kernel|pool_time|fc_kernel|l2norm" -s 45 -c 30 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 22 -c 11 -o gpurun_out/prof_conv -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
