#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== trace L2 shape"; timeout 120 python tools/micro/trace_halo.py 40 16 128 3 2>&1 | grep -v "^tile\|deltas"
echo "== bench default"; timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_launch_ms'])"
echo "== bench lanes1"; timeout 300 python bench.py --steps 1000 --warmup 20 --lanes 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 150 -c 15 --csv --log-file gpurun_out/launches_r18.csv python bench.py --steps 6 --warmup 8 --lanes 1 --no-cpu-baseline > gpurun_out/ncu18.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_r18.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:16]:
    print(r[ki][:50].ljust(50), r[vi])
PY
