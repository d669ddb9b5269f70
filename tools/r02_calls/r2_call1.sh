#!/bin/bash
# GPU box, one GPU.  Round 2, first call: the whole -m gpu suite on the new default (4-wide) walk, then A/B numbers.
out=gpurun_out; mkdir -p $out; rm -f $out/parity_vs_reference.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $out/pytest_gpu_r2a.log 2>&1
tail -8 $out/pytest_gpu_r2a.log
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for mode in wide exact; do
  timeout 300 python bench.py $Q --trace-mode $mode > $out/ab_bunny_$mode.json 2> $out/ab_bunny_$mode.err; show $out/ab_bunny_$mode.json "bunny $mode"
done
for v in w5s1 w6s2 w6s1 w4s2; do
  GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_$v.so timeout 300 python bench.py $Q > $out/ab_bunny_$v.json 2> $out/ab_bunny_$v.err; show $out/ab_bunny_$v.json "bunny variant $v"
done
for t in "20,6,4,10,4" "20,6,4,10,6" "24,6,4,10" "16,6,4,10" "20,8,4,10" "20,4,4,10" "20,6,4,14" "20,6,4,6"; do
  timeout 300 python bench.py $Q --tune "$t" > $out/ab_tune.json 2> $out/ab_tune.err; show $out/ab_tune.json "bunny tune $t"
done
for scene in grid spheres field bunny_ao; do
  for mode in wide exact; do
    timeout 600 python bench.py --scene $scene --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats --trace-mode $mode > $out/ab_${scene}_$mode.json 2> $out/ab_${scene}_$mode.err; show $out/ab_${scene}_$mode.json "$scene $mode"
  done
done
# the full default line (stats child, e2e, fast tree, cpu baseline)
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench_r2a.json 2> $out/bench_r2a.err; tail -c 1500 $out/bench_r2a.json
# ncu: launch list + one full capture of the wide k_extend launches of one step
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 40 --csv --log-file $out/launches_r2a.csv $B > $out/ncu_l_r2a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 7 -f -o /tmp/prof_extend_r2a $B > $out/ncu_e_r2a.log 2>&1
ncu -i /tmp/prof_extend_r2a.ncu-rep --page raw --csv > $out/r2a_extend_raw.csv 2>/dev/null
ncu -i /tmp/prof_extend_r2a.ncu-rep --page details > $out/r2a_extend_details.txt 2>/dev/null
sz=$(stat -c %s /tmp/prof_extend_r2a.ncu-rep 2>/dev/null || echo 0); if [ "$sz" -gt 0 ] && [ "$sz" -lt 30000000 ]; then cp /tmp/prof_extend_r2a.ncu-rep $out/; fi
tail -n 2 $out/ncu_l_r2a.log $out/ncu_e_r2a.log
