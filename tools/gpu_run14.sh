#!/bin/bash
# straight-line MMA issue: halo tests, 5x5 planar trace, default vs planar bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo_conv.py tests/test_gpu_forward.py -m gpu -q 2>&1 | tail -2
for a in "40 16 64 128" "20 8 128 256" "10 4 256 512"; do echo "== planar 5x5 $a"; timeout 120 python tools/micro/trace_s2.py $a 2>&1 | tail -4; done
echo "== bench default"; timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_launch_ms'])"
echo "== bench planar"; DSK_PLANAR_S2=1 timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_launch_ms'])"
echo "== bench lanes1 default"; timeout 300 python bench.py --steps 1000 --warmup 20 --lanes 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
echo "== bench lanes1 planar"; DSK_PLANAR_S2=1 timeout 300 python bench.py --steps 1000 --warmup 20 --lanes 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
