#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 150 -c 45 --csv --log-file gpurun_out/launches_r16.csv python bench.py --steps 6 --warmup 8 --lanes 1 --no-cpu-baseline > gpurun_out/ncu16.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_r16.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:46]:
    print(r[ki][:60].ljust(60), r[vi])
PY
