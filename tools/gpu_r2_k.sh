#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_parity.py tests/test_gpu_backward_ops.py tests/test_gpu_head.py -m gpu -q --timeout 300 -s > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log; grep -n "loss x\|worst grad\|loss trajectory\|oracle:\|max rel dev\|B=.*T=" gpurun_out/r2k_pytest.log | cut -c1-300
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_train.json 2> gpurun_out/r2k_bench_train.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2k_bench_train.json"))
print("train: %.0f utt/s, %.3f ms/step, e2e %.0f, frac %.3f, loss %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["last_loss"]))
PY
