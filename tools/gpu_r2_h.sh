#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_forward.py tests/test_verification.py -m gpu -q --timeout 300 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2h_bench.json"))
r=d["roofline"]; w=d["windows"]; e=d["e2e"]
print("value %.0f ms %.4f (win %d: %.4f..%.4f) host %.4f | e2e %.0f ms %.4f (win %d: %.4f..%.4f) host %.4f" % (d["value"], d["ms_per_step"], w["n"], w["ms_per_step_min"], w["ms_per_step_max"], d["host_enqueue_ms_per_step"], e["value"], e["ms_per_step"], w["e2e_n"], w["e2e_ms_per_step_min"], w["e2e_ms_per_step_max"], e["host_enqueue_ms_per_step"]))
print("conv chain %.4f frac %.3f inprod %.3f clocks %s" % (r["launch_set_ms"], r["frac"], r["in_production"]["frac"], d["clocks"]))
t=d["train"]; print("train %.0f utt/s %.3f ms frac %.3f e2e %.0f" % (t["value"], t["ms_per_step"], t["roofline"]["frac"], t["e2e"]["value"]))
a=d["allpairs"]; print("allpairs %.1f us" % a["value"])
PY
