#!/bin/bash
# PDL knobs: conv1 plain launch / late trigger, value vs e2e at 3 lanes
for c in "1 0" "0 0" "1 1" "0 1"; do set -- $c
echo "== DSK_CONV1_PDL=$1 DSK_LATE_TRIGGER=$2"
DSK_CONV1_PDL=$1 DSK_LATE_TRIGGER=$2 timeout 300 python bench.py --steps 600 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
echo "== lanes 2 / 4 (defaults)"
for l in 2 4; do timeout 300 python bench.py --steps 600 --warmup 20 --lanes $l --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', round(d['e2e']['value']))"; done
