#!/bin/bash
# round-2 third GPU pass: the two-CTAs-per-SM halo conv: parity tests, then A/B of the bench with DSK_SMALL_CTA=0/1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_halo_conv.py tests/test_gpu_forward.py tests/test_gpu_pipeline.py tests/test_verification.py tests/test_gpu_head.py tests/test_gpu_train_parity.py tests/test_gpu_train.py -m gpu -q --timeout 600 -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
for v in 0 1; do
  for lanes in 1 3; do
    DSK_SMALL_CTA=$v timeout 300 python bench.py --workload infer --steps 200 --warmup 10 --lanes $lanes --no-cpu-baseline > gpurun_out/r2c_bench_small${v}_l${lanes}.json 2> gpurun_out/r2c_bench_small${v}_l${lanes}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_small${v}_l${lanes}.json"))
r=d["roofline"]
print("small=$v lanes=$lanes value %.0f ms %.4f e2e %.0f | conv chain %.4f ms frac %.3f | sections %s | clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_set_ms"], r["frac"], r["section_ms"], d["clocks"]))
print("   per launch", r["per_launch_ms_event_bracketed"])
PY
  done
done

# training step: sequential calls would be --workload train on the previous commit (8.92 ms); now forward_triplet (three streams)
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_train.json 2> gpurun_out/r2c_bench_train.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_train.json"))
print("train: %.0f utt/s, %.3f ms/step, e2e %.0f, frac %.3f, loss %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["last_loss"]))
PY
# launch list of the training step (per-kernel time shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2c_train_launches.csv python bench.py --workload train --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_train_ncu.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2c_train_launches.csv")) if len(r)>5]
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr is None: continue
    d=dict(zip(hdr,r))
    try: v=float(d["Metric Value"].replace(",",""))
    except: continue
    u=d.get("Metric Unit","")
    v = v/1000.0 if u in ("ns","nsecond") else (v*1000.0 if u in ("ms","msecond") else v)
    k=d["Kernel Name"][:60]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print("%-62s %5d %10.1f us %5.1f%%" % (k, v[0], v[1], 100*v[1]/tot))
PY
