#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv_umma|conv3x3_halo|conv1_kernel|pool_time|fc_kernel|l2norm" -s 45 -c 30 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
echo "ncu launches exit $?" > gpurun_out/summary12.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo|conv_umma" -s 22 -c 11 -o gpurun_out/prof_v2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary12.txt
cat gpurun_out/summary12.txt
