#!/usr/bin/env python
"""GPU box, one GPU: what one rank of an N-way strong split does, for N = 1, 2, 4, 8 -- the slice [0, spp / N) of the
sample indices of one frame, timed with CUDA events, per kernel class.  Shows where a strong split loses efficiency
before any collective is involved.  usage: tools/strong_probe.py [scene] [frames] [lanes]
(lanes: gb_set_tuning values[5], the wave lanes a render uses; default: the library's choice)"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from goblin_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else None
scene = api.Scene(bench.scene_path(name))
ctx = api.Context(0)
ctx.upload_scene(scene)
if lanes is not None:
    ctx.set_tuning([20, 6, 4, 10, 0, lanes, 1])
spp = scene.spp_squared()
stream = torch.cuda.ExternalStream(ctx.stream())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
base = None
for n in (1, 2, 4, 8):
    end = spp // n
    for i in range(2):
        ctx.film_clear()
        ctx.render(seed=i, spp_total=spp, spp_begin=0, spp_end=end)
    ctx.synchronize()
    ctx.reset_kernel_times()
    ctx.enable_kernel_timing(True)
    ms = 0.0
    for i in range(frames):
        with torch.cuda.stream(stream):
            flush.fill_(i)
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.film_clear()
            ctx.render(seed=10 + i, spp_total=spp, spp_begin=0, spp_end=end)
            z.record()
        ctx.synchronize()
        torch.cuda.synchronize()
        ms += a.elapsed_time(z)
    ctx.enable_kernel_timing(False)
    kt = ctx.kernel_times()
    ms /= frames
    base = base or ms
    print(json.dumps({"scene": name, "lanes": lanes, "split": n, "spp_slice": end, "ms_per_frame": round(ms, 3),
                      "efficiency_vs_split1": round(base / (n * ms), 3),
                      "kernel_ms": {k: round(v[0] / frames, 3) for k, v in kt.items() if v[1]},
                      "launches_per_frame": {k: v[1] // frames for k, v in kt.items() if v[1]}}))
