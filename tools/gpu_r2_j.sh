#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_pipeline.py tests/test_verification.py tests/test_gpu_train.py -m gpu -q --timeout 300 > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -6 gpurun_out/r2j_pytest.log; grep -n "loss x\|with loss scale" gpurun_out/r2j_pytest.log
for lanes in 1 3; do
  timeout 300 python bench.py --workload infer --steps 200 --warmup 10 --lanes $lanes --no-cpu-baseline > gpurun_out/r2j_bench_l${lanes}.json 2> gpurun_out/r2j_bench_l${lanes}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2j_bench_l${lanes}.json"))
    r=d["roofline"]
    print("lanes=$lanes value %.0f ms %.4f e2e %.0f | conv chain %.4f ms frac %.3f sections %s | clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_set_ms"], r["frac"], r["section_ms"], d["clocks"]))
    print("   per launch", r["per_launch_ms_event_bracketed"])
except Exception as e:
    print("lanes=$lanes FAILED", e); print(open("gpurun_out/r2j_bench_l${lanes}.err").read()[-1500:])
PY
done
