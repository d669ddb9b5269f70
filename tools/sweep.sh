#!/bin/bash
# usage: tools/sweep.sh "<tune1>" "<tune2>" ...   (run on the GPU box)
for t in "$@"; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tune "$t" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms_per_step']
print('$t', round(d['value'],1), round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items()})"
done
