#!/bin/bash
# GPU box, one GPU.  Round 2, third call: min/max slab test A/B, upload timing, source-level profile of the pair walk.
out=gpurun_out; mkdir -p $out
( timeout 600 python -m pytest tests/test_gpu_vs_oracle.py tests/test_api_surface.py -m gpu -x -q 2>&1 | tail -4 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for mode in wide exact; do
  timeout 300 python bench.py $Q --trace-mode $mode > $out/ab3_bunny_$mode.json 2> $out/ab3_bunny_$mode.err; show $out/ab3_bunny_$mode.json "bunny minmax $mode"
done
for v in ordered w6s2; do
  for mode in wide exact; do
    GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 300 python bench.py $Q --trace-mode $mode > $out/ab3_bunny_${v}_$mode.json 2> $out/ab3_bunny_${v}_$mode.err; show $out/ab3_bunny_${v}_$mode.json "bunny variant $v $mode"
  done
done
for scene in grid spheres field bunny_ao; do
  for mode in wide exact; do
    timeout 600 python bench.py --scene $scene --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats --trace-mode $mode > $out/ab3_${scene}_$mode.json 2> $out/ab3_${scene}_$mode.err; show $out/ab3_${scene}_$mode.json "$scene minmax $mode"
  done
done
GB_UPLOAD_TIMING=1 python - <<'PY' 2>&1 | tail -24
import sys, time
sys.path.insert(0, '.')
import bench
from goblin_b200 import api
sc = api.Scene(bench.scene_path("bunny"))
ctx = api.Context(0)
for i in range(3):
    t0 = time.perf_counter(); ctx.upload_scene(sc); print("upload_scene", round((time.perf_counter() - t0) * 1e3, 3), "ms", flush=True)
for i in range(3):
    t0 = time.perf_counter(); ctx.upload_scene_async(sc); t1 = time.perf_counter(); ctx.synchronize(); print("upload_scene_async host", round((t1 - t0) * 1e3, 3), "ms, +sync", round((time.perf_counter() - t1) * 1e3, 3), flush=True)
PY
timeout 900 python bench.py --steps 20 --warmup 3 --trace-mode exact --no-fast-tree > $out/bench_r2c.json 2> $out/bench_r2c.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2c.json').read().strip().splitlines()[-1])
print('bench exact: value', round(d['value'],1), 'e2e', d['e2e'], 'frac', d['roofline']['frac'])
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --trace-mode exact"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 22 -c 2 -f -o /tmp/prof_extend_r2c $B > $out/ncu_e_r2c.log 2>&1
ncu -i /tmp/prof_extend_r2c.ncu-rep --page raw --csv > $out/r2c_extend_raw.csv 2>/dev/null
ncu -i /tmp/prof_extend_r2c.ncu-rep --page source --csv > $out/r2c_extend_source.csv 2>/dev/null
ls -la /tmp/prof_extend_r2c.ncu-rep $out/r2c_extend_source.csv
tail -n 1 $out/ncu_e_r2c.log | cut -c1-200
