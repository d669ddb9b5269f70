#!/bin/bash
# usage: tools/profile_round.sh <tag>   (run on the GPU box, one GPU): ncu launch list and ncu --set full captures of
# the dominant traversal kernel on bunny (S1) and on the 10 M-triangle grid (S4), plus the shade kernels.  Only the
# raw / details pages (CSV, text) travel back, and the bunny k_extend report for its source page: gpurun_out is
# capped at 64 MiB.
tag=$1
out=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
G="python bench.py --scene grid --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
$B > $out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 40 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 7 -f -o /tmp/prof_extend_$tag $B > $out/ncu_e_$tag.log 2>&1
ncu --set full --clock-control none -k regex:k_shade -s 42 -c 6 -f -o /tmp/prof_shade_$tag $B > $out/ncu_s_$tag.log 2>&1
$G > $out/plain_grid_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:k_extend -s 21 -c 7 -f -o /tmp/prof_grid_$tag $G > $out/ncu_g_$tag.log 2>&1
for n in extend shade grid; do
  ncu -i /tmp/prof_${n}_$tag.ncu-rep --page raw --csv > $out/${tag}_${n}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_${n}_$tag.ncu-rep --page details > $out/${tag}_${n}_details.txt 2>/dev/null
done
ls -la /tmp/*.ncu-rep
sz=$(stat -c %s /tmp/prof_extend_$tag.ncu-rep)
if [ "$sz" -lt 40000000 ]; then cp /tmp/prof_extend_$tag.ncu-rep $out/; fi
du -sh $out
for f in $out/ncu_?_$tag.log; do tail -n 2 $f; done
