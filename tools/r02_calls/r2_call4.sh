#!/bin/bash
# GPU box, one GPU.  Round 2, fourth call: two wave lanes A/B, scheduling-interval variants, sampler tests.
out=gpurun_out; mkdir -p $out
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_vs_oracle.py tests/test_api_surface.py tests/test_sampler.py tests/test_scene_variants.py tests/test_image_io.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -6 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny grid spheres field bunny_ao; do
  for t in "20,6,4,10,0,2,1" "20,6,4,10,0,1,1" "20,6,4,10,0,2,0"; do
    timeout 600 python bench.py --scene $scene $Q --tune "$t" > $out/ab4_${scene}.json 2> $out/ab4_${scene}.err; show $out/ab4_${scene}.json "$scene tune $t"
  done
done
for v in steps3 steps4 steps1; do
  GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 300 python bench.py $Q > $out/ab4_bunny_$v.json 2> $out/ab4_bunny_$v.err; show $out/ab4_bunny_$v.json "bunny variant $v"
  GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 600 python bench.py --scene grid $Q > $out/ab4_grid_$v.json 2> $out/ab4_grid_$v.err; show $out/ab4_grid_$v.json "grid variant $v"
done
echo "--- strong probe, two lanes"; timeout 300 python tools/strong_probe.py bunny 5 2>&1 | tail -4 | cut -c1-330
echo "--- strong probe, one lane"; GB_WAVE_LANES=1 timeout 300 python tools/strong_probe.py bunny 5 2>&1 | tail -4 | cut -c1-330
echo "--- strong probe grid, two lanes"; timeout 300 python tools/strong_probe.py grid 3 2>&1 | tail -4 | cut -c1-330
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench_r2d.json 2> $out/bench_r2d.err; tail -3 $out/bench_r2d.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2d.json').read().strip().splitlines()[-1])
print('bench: value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'serial', round(d['e2e']['serial_value'],1), 'frac', d['roofline']['frac'], 'fast', (d.get('fast_tree') or {}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'))
print(d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_ms_per_step_overlapped'])
PY
