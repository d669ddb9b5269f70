#!/bin/bash
# usage: tools/sweep_scene.sh <scene> "<tune1>" ...   (run on the GPU box)
scene=$1; shift
for t in "$@"; do
  python bench.py --scene $scene --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --tune "$t" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$scene $t', round(d['value'],1), round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()})"
done
