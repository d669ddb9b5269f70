#!/bin/bash
# GPU box, one GPU.  Round 2, eighth call: film kernel A/B, scheduling-knob sweep on the 3-stage build, fault core dump, AO traffic.
out=gpurun_out; mkdir -p $out
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_vs_oracle.py -m gpu -x -q 2>&1 | tail -3 ) 2>&1
Q="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fast-tree --no-stats"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1]); k=d['roofline']['kernel_ms_per_step']
    print('$2', round(d['value'],1), 'Msamples/s', round(d['ms_per_step'],2), 'ms', {a:round(b,2) for a,b in k.items()})
except Exception as e: print('$2 FAILED', e)
"; }
for scene in bunny spheres; do
  timeout 300 python bench.py --scene $scene $Q > $out/ab8_$scene.json 2> $out/ab8_$scene.err; show $out/ab8_$scene.json "$scene film-lds"
  GOBLIN_B200_LIB=$PWD/goblin_b200/variants/libgoblin_b200_filmshfl.so timeout 300 python bench.py --scene $scene $Q > $out/ab8_${scene}_shfl.json 2> $out/ab8_${scene}_shfl.err; show $out/ab8_${scene}_shfl.json "$scene film-shfl"
done
for t in "16,6,4,10" "24,6,4,10" "28,6,4,10" "20,4,4,10" "20,8,4,10" "20,10,4,10" "20,6,2,10" "20,6,6,10" "20,6,4,6" "20,6,4,14" "24,8,4,10" "24,8,6,14"; do
  timeout 300 python bench.py $Q --tune "$t" > $out/ab8_tune.json 2> $out/ab8_tune.err; show $out/ab8_tune.json "bunny tune $t"
done
for t in "24,6,4,10" "20,8,4,10" "24,8,6,14"; do
  timeout 600 python bench.py --scene grid $Q --tune "$t" > $out/ab8_tune.json 2> $out/ab8_tune.err; show $out/ab8_tune.json "grid tune $t"
done
echo "--- core dump of the dropped variant"; bash tools/fault_core.sh popend 10 2>&1 | tee $out/fault_core_popend.txt | tail -100
echo "--- AO traffic"
B="python bench.py --scene bunny_ao --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --tune 20,6,4,10,0,1,0"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k k_ao -s 3 -c 1 --csv --log-file $out/traffic_bunny_ao.csv $B > $out/traffic_bunny_ao.log 2>&1; grep -v "^==" $out/traffic_bunny_ao.csv | tail -4 | cut -c1-300
