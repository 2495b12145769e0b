#!/bin/bash
DSK_N256=1 DSK_N256_MIN_TILES=1 timeout 600 python -m pytest tests/test_gpu_halo_conv.py tests/test_gpu_forward.py -m gpu -q -x 2>&1 | tail -3
for cfg in "0 80" "1 80" "1 40"; do set -- $cfg
for l in 1 3; do
DSK_N256=$1 DSK_N256_MIN_TILES=$2 timeout 300 python bench.py --steps 600 --warmup 20 --lanes $l --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n256=$1 min=$2 lanes', d['config']['forwards_in_flight'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), [round(x*1e3,1) for x in d['roofline']['per_launch_ms']])"
done; done
