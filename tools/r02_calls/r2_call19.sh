#!/bin/bash
# GPU box, one GPU: any-hit walks that park a reached triangle leaf and keep walking (GB_SPEC_ANY=1): parity (any-hit answers
# are bit-compared with the oracle and the live reference in these files), then the bench per scene against the shipped build,
# with the triangle-stage batch threshold (gb_set_tuning values[1]) at 6 (default), 3 and 10.
out=gpurun_out; mkdir -p $out
V=$PWD/goblin_b200/variants/libgoblin_b200_spec.so
( GOBLIN_B200_LIB=$V timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_gpu_golden.py tests/test_scene_variants.py -m gpu -x -q 2>&1 | tail -3 ) 2>&1
( GOBLIN_B200_LIB=$V timeout 900 python -m pytest tests/test_gpu_vs_reference.py -m gpu -x -q -k "exported" 2>&1 | tail -3 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny bunny_ao spheres grid; do
  timeout 600 python bench.py --scene $scene $Q > $out/ab19_$scene.json 2> $out/ab19_$scene.err; show $out/ab19_$scene.json "$scene shipped"
  GOBLIN_B200_LIB=$V timeout 600 python bench.py --scene $scene $Q > $out/ab19_${scene}_spec.json 2> $out/ab19_${scene}_spec.err; show $out/ab19_${scene}_spec.json "$scene spec"
  for lb in 3 10; do
    GOBLIN_B200_LIB=$V timeout 600 python bench.py --scene $scene $Q --tune 20,$lb,4,10 > $out/ab19_${scene}_spec_lb$lb.json 2> $out/ab19_${scene}_spec_lb$lb.err; show $out/ab19_${scene}_spec_lb$lb.json "$scene spec leafBatch=$lb"
  done
done
