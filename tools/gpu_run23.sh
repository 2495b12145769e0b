#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_halo_conv.py tests/test_gpu_forward.py tests/test_verification.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -3
for l in 1 3; do
timeout 300 python bench.py --steps 1000 --warmup 20 --lanes $l --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('lanes', d['config']['forwards_in_flight'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(r['frac'],4), 'sections', {k: round(v*1e3,1) for k,v in r['section_ms'].items()}, [round(x*1e3,1) for x in r['per_launch_ms_event_bracketed']])"
done
for a in "80 32 64" "40 16 128"; do echo "== $a"; timeout 120 python tools/micro/trace_halo.py $a 3 2>&1 | grep "^tile" | sed -n 3,5p; done
