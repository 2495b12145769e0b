#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward_ops.py tests/test_gpu_train.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1
echo "pytest train+ops exit $?" > gpurun_out/summary7.txt
cat gpurun_out/summary7.txt; grep -v "^E    \|^E  +" gpurun_out/pytest_train.log | tail -60
