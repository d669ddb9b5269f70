#!/bin/bash
# GPU box, one GPU: ncu --set full of the final build's kernels on S1 (k_extend x7, k_shade, k_film, k_shadow bounce 0) and k_extend on S4.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --tune 20,6,4,10,0,1,0"
G="python bench.py --scene grid --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --tune 20,6,4,10,0,1,0"
$B > $out/plain_r2i.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 7 -f -o /tmp/prof_extend_r2i $B > $out/ncu_e_r2i.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_shade|k_film|k_shadow|k_raygen" -s 84 -c 6 -f -o /tmp/prof_other_r2i $B > $out/ncu_o_r2i.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_extend -s 48 -c 8 -f -o /tmp/prof_grid_r2i $G > $out/ncu_g_r2i.log 2>&1
for n in extend other grid; do
  ncu -i /tmp/prof_${n}_r2i.ncu-rep --page raw --csv > $out/r2i_${n}_raw.csv 2>/dev/null
done
ncu -i /tmp/prof_extend_r2i.ncu-rep --page details > $out/r2i_extend_details.txt 2>/dev/null
ls -la /tmp/*.ncu-rep | head; for f in $out/ncu_?_r2i.log; do tail -n 1 $f | cut -c1-160; done
