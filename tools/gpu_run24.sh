#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -5
for l in 2 3 4; do
timeout 300 python bench.py --steps 1000 --warmup 20 --lanes $l --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('lanes', d['config']['forwards_in_flight'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'host', round(d['e2e']['host_enqueue_ms_per_step'],3), 'frac', round(r['frac'],4))"
done
