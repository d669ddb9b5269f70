#!/bin/bash
# GPU box, one GPU: L1 cache-policy hints on the traversal's record fetches (evict_last on nodes, no_allocate on triangle records).
out=gpurun_out; mkdir -p $out
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny grid spheres field bunny_ao; do
  timeout 600 python bench.py --scene $scene $Q > $out/ab14_$scene.json 2> $out/ab14_$scene.err; show $out/ab14_$scene.json "$scene shipped"
  for v in nodeel nodeel_trina trina; do
    GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 600 python bench.py --scene $scene $Q > $out/ab14_${scene}_$v.json 2> $out/ab14_${scene}_$v.err; show $out/ab14_${scene}_$v.json "$scene $v"
  done
done
