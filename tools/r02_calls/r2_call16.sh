#!/bin/bash
# GPU box, one GPU: one-warp film kernel (k_film_warp) + shade launches only for material types the scene holds.
# Parity tests that cover the film (every filter variant, both film kernels), the bench on three scenes, and an
# ncu capture of the bounce-0 Lambert shade launch and of the new film kernel.
out=gpurun_out; mkdir -p $out
( timeout 900 python -m pytest tests/test_scene_variants.py tests/test_api_surface.py tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()}, 'launches', d.get('gpu_launches'))
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny spheres grid; do
  timeout 600 python bench.py --scene $scene $Q > $out/ab16_$scene.json 2> $out/ab16_$scene.err; show $out/ab16_$scene.json "$scene"
done
B="python bench.py --scene bunny --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --tune 20,6,4,10,0,1,0"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 1 -o $out/r02j_shade_lambert_b0 -f $B > $out/ncu16a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_film -c 1 -o $out/r02j_film_warp -f $B > $out/ncu16b.log 2>&1
ls -la $out/*.ncu-rep 2>/dev/null | tail -3
