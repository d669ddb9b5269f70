#!/bin/bash
# GPU box, one GPU: closest-hit record in shared memory -> 64 registers -> 8 CTAs / SM (32 warps) instead of 7.
out=gpurun_out; mkdir -p $out
( GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_hs8.so timeout 600 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -2 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny spheres bunny_ao grid field; do
  timeout 600 python bench.py --scene $scene $Q > $out/ab15_$scene.json 2> $out/ab15_$scene.err; show $out/ab15_$scene.json "$scene shipped"
  for v in hs7 hs8; do
    GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 600 python bench.py --scene $scene $Q > $out/ab15_${scene}_$v.json 2> $out/ab15_${scene}_$v.err; show $out/ab15_${scene}_$v.json "$scene $v"
  done
done
