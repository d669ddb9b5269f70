#!/bin/bash
# usage: tools/quick_bench.sh <scene>... (GPU box): resident Msamples/s + per-kernel-class ms, parity tree
for scene in "$@"; do
  python bench.py --scene $scene --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-stats 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$scene', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})"
done
