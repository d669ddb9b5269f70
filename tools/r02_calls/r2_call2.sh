#!/bin/bash
# GPU box, one GPU.  Round 2, second call: 256-bit sector loads (A/B against 128-bit), async upload, strong-split probe.
out=gpurun_out; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_vs_oracle.py tests/test_api_surface.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -15 ) > $out/pytest_gpu_r2b.log 2>&1
tail -6 $out/pytest_gpu_r2b.log
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for mode in wide exact; do
  timeout 300 python bench.py $Q --trace-mode $mode > $out/ab2_bunny_$mode.json 2> $out/ab2_bunny_$mode.err; show $out/ab2_bunny_$mode.json "bunny ld256 $mode"
done
for v in ld128 w6s2 w6s1; do
  for mode in wide exact; do
    GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 300 python bench.py $Q --trace-mode $mode > $out/ab2_bunny_${v}_$mode.json 2> $out/ab2_bunny_${v}_$mode.err; show $out/ab2_bunny_${v}_$mode.json "bunny variant $v $mode"
  done
done
for scene in grid spheres field bunny_ao; do
  for mode in wide exact; do
    timeout 600 python bench.py --scene $scene --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats --trace-mode $mode > $out/ab2_${scene}_$mode.json 2> $out/ab2_${scene}_$mode.err; show $out/ab2_${scene}_$mode.json "$scene ld256 $mode"
  done
done
timeout 300 python tools/strong_probe.py bunny 5 2>&1 | tail -4
timeout 300 python tools/strong_probe.py grid 3 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --trace-mode exact > $out/bench_r2b.json 2> $out/bench_r2b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print('bench exact: value', round(d['value'],1), 'e2e', d['e2e'], 'frac', d['roofline']['frac'], 'fast', (d.get('fast_tree') or {}).get('value'))
PY
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --trace-mode exact"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 7 -f -o /tmp/prof_extend_r2b $B > $out/ncu_e_r2b.log 2>&1
ncu -i /tmp/prof_extend_r2b.ncu-rep --page raw --csv > $out/r2b_extend_raw.csv 2>/dev/null
tail -n 1 $out/ncu_e_r2b.log | cut -c1-200
