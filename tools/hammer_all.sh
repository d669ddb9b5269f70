#!/bin/bash
# usage: tools/hammer_all.sh [runs]   (GPU box): repeated random ray batches through both instantiations of the traversal
# kernels (counters off / on) on every named scene, plus repeated counter-on renders; any CUDA fault or closest / any-hit
# disagreement fails.
runs=${1:-3}
fail=0
for scene in bunny spheres grid_small field; do
  for mode in plain stats; do
    for seed in $(seq 1 $runs); do
      python tools/hammer.py $scene $mode $seed > /tmp/hammer.log 2>&1 || { fail=$((fail+1)); echo "FAILED $scene $mode $seed"; tail -2 /tmp/hammer.log; }
    done
  done
done
for scene in bunny spheres; do
  for i in $(seq 1 $runs); do
    python bench.py --scene $scene --stats-only 1000,16,0,16 > /tmp/stats.log 2>&1 || { fail=$((fail+1)); echo "FAILED stats render $scene"; tail -2 /tmp/stats.log; }
  done
done
echo "hammer failures: $fail"
