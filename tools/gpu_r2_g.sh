#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -6 gpurun_out/r2g_pytest.log
for lanes in 1 3; do
  timeout 300 python bench.py --workload infer --steps 200 --warmup 10 --lanes $lanes --no-cpu-baseline > gpurun_out/r2g_bench_l${lanes}.json 2> gpurun_out/r2g_bench_l${lanes}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_bench_l${lanes}.json"))
    r=d["roofline"]
    print("lanes=$lanes value %.0f ms %.4f e2e %.0f | conv chain %.4f ms frac %.3f | clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_set_ms"], r["frac"], d["clocks"]))
    print("   per launch", r["per_launch_ms_event_bracketed"])
except Exception as e:
    print("lanes=$lanes FAILED", e); print(open("gpurun_out/r2g_bench_l${lanes}.err").read()[-1500:])
PY
done
for cfg in "80 32 64" "40 16 128" "10 4 512"; do
  echo "== trace $cfg" >> gpurun_out/r2g_trace.txt
  timeout 120 python tools/micro/trace_halo.py $cfg 3 >> gpurun_out/r2g_trace.txt 2>&1
done
grep -v "^single\|^producer\|deltas:" gpurun_out/r2g_trace.txt
