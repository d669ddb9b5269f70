#!/bin/bash
# GPU box, one GPU: one rank's slice of an N-way strong split (tools/strong_probe.py) with 1 .. 4 wave lanes.
for scene in bunny grid; do
  for lanes in 1 2 3 4; do
    timeout 600 python tools/strong_probe.py $scene 8 $lanes 2>&1 | grep '"split"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['scene'], 'lanes', d['lanes'], 'split', d['split'], d['ms_per_frame'], 'ms', 'eff', d['efficiency_vs_split1'], {k: v for k, v in d['kernel_ms'].items()})
"
  done
done
