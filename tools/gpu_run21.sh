#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_train_r21.csv python bench.py --workload train --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu21.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_train_r21.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
rows=rows[1:]
# last step = second half
n=len(rows); half=rows[n//2:] if n>40 else rows
agg=collections.OrderedDict()
for r in half:
    k=r[ki][:70]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[vi].replace(',',''))/1e3
tot=sum(v[1] for v in agg.values())
print("kernels in window:", len(half), "total us", round(tot,1))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:28]:
    print(f"{v[1]:9.1f} us {v[0]:4d}x  {k}")
PY
