#!/bin/bash
# usage: tools/profile_round.sh <tag>   (run on the GPU box, one GPU): bench line, ncu launch list, and ncu --set full
# captures of the dominant traversal kernel on bunny (S1) and on the 10 M-triangle grid (S4).
tag=$1
out=gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
python bench.py --steps 5 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err || exit 1
$B > $out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 40 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 7 -f -o $out/prof_extend_$tag $B > $out/ncu_e_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 42 -c 6 -f -o $out/prof_shade_$tag $B > $out/ncu_s_$tag.log 2>&1
G="python bench.py --scene grid --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats"
$G > $out/plain_grid_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 7 -f -o $out/prof_grid_$tag $G > $out/ncu_g_$tag.log 2>&1
tail -2 $out/ncu_*_$tag.log
cat $out/bench_$tag.json
