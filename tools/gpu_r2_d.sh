#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
for v in 0 1; do
  for lanes in 1 3; do
    DSK_STREAM_K=$v timeout 300 python bench.py --workload infer --steps 200 --warmup 10 --lanes $lanes --no-cpu-baseline > gpurun_out/r2d_bench_sk${v}_l${lanes}.json 2> gpurun_out/r2d_bench_sk${v}_l${lanes}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_bench_sk${v}_l${lanes}.json"))
    r=d["roofline"]
    print("stream_k=$v lanes=$lanes value %.0f ms %.4f e2e %.0f | conv chain %.4f ms frac %.3f | clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_set_ms"], r["frac"], d["clocks"]))
    print("   per launch", r["per_launch_ms_event_bracketed"])
except Exception as e:
    print("stream_k=$v lanes=$lanes FAILED", e); print(open("gpurun_out/r2d_bench_sk${v}_l${lanes}.err").read()[-1500:])
PY
  done
done
