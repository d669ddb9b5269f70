#!/bin/bash
# usage: tools/accel_compare.sh <scene>...   (run on the GPU box) -- parity tree vs fast tree, same workload
for scene in "$@"; do
  for accel in equal_count sah; do
    python bench.py --scene $scene --accel $accel --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/accel_${scene}_${accel}.json
    python - "$scene" "$accel" <<'PY'
import json, sys
d = json.load(open("gpurun_out/accel_%s_%s.json" % (sys.argv[1], sys.argv[2])))
k = d["roofline"]["kernel_ms_per_step"]
print(sys.argv[1], sys.argv[2], "Msamples/s", round(d["value"], 1), "Mrays/s", round(d["mrays_per_s"], 1), "ms/step", round(d["ms_per_step"], 2),
      "nodes/ray", round(d["roofline"]["nodes_per_ray"], 1), "load_s", round(d["config"]["scene_load_s"], 2), {a: round(b, 2) for a, b in k.items()})
PY
  done
done
