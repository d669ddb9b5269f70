#!/bin/bash
# GPU box, one GPU.  Round 2, fifth call: compile-time variants of the pair walk, shade with selective strata, fault hunt, S4 e2e.
out=gpurun_out; mkdir -p $out
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_vs_oracle.py tests/test_sampler.py -m gpu -x -q 2>&1 | tail -4 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
timeout 300 python bench.py $Q > $out/ab5_bunny.json 2> $out/ab5_bunny.err; show $out/ab5_bunny.json "bunny shipped"
timeout 600 python bench.py --scene grid $Q > $out/ab5_grid.json 2> $out/ab5_grid.err; show $out/ab5_grid.json "grid shipped"
timeout 600 python bench.py --scene spheres $Q > $out/ab5_spheres.json 2> $out/ab5_spheres.err; show $out/ab5_spheres.json "spheres shipped"
for v in steps1 steps3 steps4 poplane poplane3 blocks8; do
  for scene in bunny grid spheres; do
    GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 600 python bench.py --scene $scene $Q > $out/ab5_${scene}_$v.json 2> $out/ab5_${scene}_$v.err; show $out/ab5_${scene}_$v.json "$scene variant $v"
  done
done
echo "--- fault hunt"; bash tools/fault_hunt.sh 12 2>&1 | tail -10 | cut -c1-400
echo "--- S4 end to end"
timeout 1200 python bench.py --scene grid --steps 5 --warmup 3 --no-fast-tree --no-cpu-baseline > $out/bench_grid_r2e.json 2> $out/bench_grid_r2e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_grid_r2e.json').read().strip().splitlines()[-1])
print('grid: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'}, 'frac', d['roofline']['frac'])
PY
