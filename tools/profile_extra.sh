#!/bin/bash
# usage: tools/profile_extra.sh <tag>  (GPU box): ncu --set full of the any-hit kernels (k_shadow on S1, k_ao on S2)
tag=$1; out=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
A="python bench.py --scene bunny_ao --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
$B > $out/plain_x_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:k_shadow -s 21 -c 7 -f -o /tmp/prof_shadow_$tag $B > $out/ncu_sh_$tag.log 2>&1
$A > $out/plain_a_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:k_ao -s 9 -c 3 -f -o /tmp/prof_ao_$tag $A > $out/ncu_ao_$tag.log 2>&1
for n in shadow ao; do
  ncu -i /tmp/prof_${n}_$tag.ncu-rep --page raw --csv > $out/${tag}_${n}_raw.csv 2>/dev/null
done
ls -la /tmp/*.ncu-rep; tail -n 1 $out/ncu_sh_$tag.log $out/ncu_ao_$tag.log
