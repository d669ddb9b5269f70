#!/bin/bash
# GPU box, one GPU: any-hit walks (shadow, AO) that always visit the left child first (GB_ANY_FIXED_ORDER=1; the answer does not
# depend on the order) against the shipped near-child-first walks.
out=gpurun_out; mkdir -p $out
V=$PWD/goblin_b200/variants/libgoblin_b200_anyfix.so
( GOBLIN_B200_LIB=$V timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny bunny_ao spheres grid; do
  timeout 600 python bench.py --scene $scene $Q > $out/ab21_$scene.json 2> $out/ab21_$scene.err; show $out/ab21_$scene.json "$scene shipped"
  GOBLIN_B200_LIB=$V timeout 600 python bench.py --scene $scene $Q > $out/ab21_${scene}_anyfix.json 2> $out/ab21_${scene}_anyfix.err; show $out/ab21_${scene}_anyfix.json "$scene left-first any-hit"
done
