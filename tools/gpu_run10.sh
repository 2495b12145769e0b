#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_conv.py -m gpu -q > gpurun_out/pytest_halo.log 2>&1
echo "pytest halo exit $?" > gpurun_out/summary10.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo|conv_umma" -s 22 -c 11 -o gpurun_out/prof_halo -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_halo.log 2>&1
echo "ncu exit $?" >> gpurun_out/summary10.txt
cat gpurun_out/summary10.txt; tail -3 gpurun_out/pytest_halo.log
