#!/bin/bash
# GPU box, one GPU: k_shade with queue entries read two trips ahead, next trip's path state asked into L2, one 64-bit
# atomic for both queue reservations; film producer without divisions.  Parity tests, then the bench per scene:
# shipped build, the same without the prefetch (nopf), Lambert / Blinn shade compiled for 5 and 7 CTAs per SM.
out=gpurun_out; mkdir -p $out
( timeout 900 python -m pytest tests/test_scene_variants.py tests/test_api_surface.py tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py tests/test_gpu_vs_reference.py -m gpu -x -q 2>&1 | tail -3 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny spheres grid field; do
  timeout 600 python bench.py --scene $scene $Q > $out/ab17_$scene.json 2> $out/ab17_$scene.err; show $out/ab17_$scene.json "$scene shipped"
  [ $scene = field ] && continue
  for v in nopf sl5 sl7; do
    GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 600 python bench.py --scene $scene $Q > $out/ab17_${scene}_$v.json 2> $out/ab17_${scene}_$v.err; show $out/ab17_${scene}_$v.json "$scene $v"
  done
done
