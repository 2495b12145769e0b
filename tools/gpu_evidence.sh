#!/bin/bash
# Round-end evidence runs (one B200 unless stated).  Usage: bash tools/gpu_evidence.sh <stage> ...
#   tests      pytest -m gpu                                   -> gpurun_out/ev_pytest_gpu.log
#   smoke      __graft_entry__.smoke()                         -> gpurun_out/ev_smoke.log
#   bench      default bench line + reference arm (driver's --steps 20 --warmup 5) + the 2000-step default
#   launches   ncu launch list of the bench command            -> gpurun_out/ev_launches.csv
#   full       ncu --set full of the conv kernels of a forward -> gpurun_out/ev_conv_full.ncu-rep
#   sanitizer  compute-sanitizer memcheck / racecheck / synccheck over the kernel-level GPU tests
#   multi      torchrun bench at the visible GPU count (gpurun --gpus N)
mkdir -p gpurun_out
for stage in "$@"; do
case $stage in
tests)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/ev_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ev_pytest_gpu.log
  tail -4 gpurun_out/ev_pytest_gpu.log ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ev_smoke.log; tail -2 gpurun_out/ev_smoke.log ;;
bench)
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/ev_bench_ref.json 2> gpurun_out/ev_bench_ref.err
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err; echo "bench rc=$?"
  timeout 600 python bench.py --workload infer --no-cpu-baseline > gpurun_out/ev_bench_2000.json 2> gpurun_out/ev_bench_2000.err
  timeout 600 python bench.py --workload infer --no-cpu-baseline --lanes 1 --steps 500 > gpurun_out/ev_bench_lanes1.json 2> gpurun_out/ev_bench_lanes1.err
  timeout 600 python bench.py --workload infer --no-cpu-baseline --dtype bf16 --steps 500 > gpurun_out/ev_bench_bf16.json 2> gpurun_out/ev_bench_bf16.err
  python - <<PY
import json
for f in ("ev_bench", "ev_bench_2000", "ev_bench_lanes1", "ev_bench_bf16", "ev_bench_ref"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        r = d.get("roofline", {})
        print(f, "value %.0f ms %.4f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "| frac", r.get("frac"), "conv ms", r.get("launch_set_ms"), "| clocks", d.get("clocks"))
        if "train" in d: print("   train %.0f utt/s %.3f ms frac %.3f | allpairs %.1f us" % (d["train"]["value"], d["train"]["ms_per_step"], d["train"]["roofline"]["frac"], d["allpairs"]["value"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
  ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/ev_launches.csv python bench.py --workload infer --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ev_launches.log 2>&1
  python tools/ncu_summary.py launches gpurun_out/ev_launches.csv | head -30 ;;
full)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_halo|conv1_umma" -s 24 -c 12 -o gpurun_out/ev_conv_full -f python bench.py --workload infer --steps 6 --warmup 3 --lanes 1 --no-cpu-baseline > gpurun_out/ev_conv_full.log 2>&1
  ls -la gpurun_out/ev_conv_full.ncu-rep; python tools/ncu_summary.py full gpurun_out/ev_conv_full.ncu-rep | tail -20 ;;
sanitizer)
  # kernel-level tests under compute-sanitizer; each tool bounded (sanitized tcgen05 / TMA kernels run 10-100x slower)
  SAN_TESTS="tests/test_gpu_loss.py tests/test_gpu_head.py tests/test_gpu_halo_conv.py tests/test_gpu_backward_ops.py tests/test_gpu_forward.py::test_eval_forward_matches_oracle tests/test_gpu_train.py::test_forward_triplet_is_bit_identical_to_three_sequential_calls"
  for tool in memcheck racecheck synccheck; do
    timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SAN_TESTS -m gpu -q --timeout 3000 -x -k "not (80-32-64 or 40-16-128 or 17-0) and not two_ctas and not stream_k" > gpurun_out/ev_sanitizer_$tool.log 2>&1
    echo "$tool rc=$?" | tee -a gpurun_out/ev_sanitizer_$tool.log; grep -E "ERROR SUMMARY|passed|failed|error" gpurun_out/ev_sanitizer_$tool.log | tail -4
  done ;;
multi)
  N=$(nvidia-smi -L | wc -l)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/ev_bench_${N}gpu.json 2> gpurun_out/ev_bench_${N}gpu.err; echo "bench $N gpus rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/ev_bench_${N}gpu.json"))
print("N=%d value %.0f e2e %.0f ms %.4f | train %.0f utt/s %.3f ms (global batch %d triplets)" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["train"]["value"], d["train"]["ms_per_step"], d["train"]["config"]["global_batch_triplets"]))
PY
  ;;
esac
done
