#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "--- core dump of the dropped variant"; bash tools/fault_core.sh popend 14 2>&1 | tee $out/fault_core_popend.txt | tail -120
echo "--- shipped build, same hunt (must not fault)"; bash tools/fault_core.sh "" 1 2>&1 | tail -2
echo "--- AO traffic"
B="python bench.py --scene bunny_ao --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-stats --no-fast-tree --tune 20,6,4,10,0,1,0"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:k_ao<" -s 3 -c 1 --csv --log-file $out/traffic_bunny_ao.csv $B > $out/traffic_bunny_ao.log 2>&1; grep -v "^==" $out/traffic_bunny_ao.csv | tail -4 | cut -c1-300
