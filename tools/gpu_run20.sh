#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_verification.py tests/test_gpu_halo_conv.py -m gpu -q -x 2>&1 | tail -3
for g in 1 0; do for l in 1 3; do
DSK_GRAPH=$g timeout 300 python bench.py --steps 600 --warmup 20 --lanes $l --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('graph $g lanes', d['config']['forwards_in_flight'], 'value', round(d['value']), 'host ms/step', round(d['host_enqueue_ms_per_step'],4), '| e2e', round(d['e2e']['value']), 'host', round(d['e2e']['host_enqueue_ms_per_step'],4))"
done; done
