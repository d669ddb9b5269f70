#!/usr/bin/env python
"""GPU box: upload a named bench scene and render one frame (for ncu captures of a single launch).
usage: tools/one_render.py [bunny_ao|bunny]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from goblin_b200 import api  # noqa: E402

# the bunny scenes of bench.py (scenes/_gen/bunny, made by goblin_b200/bin/scene_gen; no torch import here)
name = sys.argv[1] if len(sys.argv) > 1 else "bunny_ao"
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "scenes", "_gen", "bunny", {"bunny_ao": "bunny_ao.json", "bunny": "bunny_pt.json"}[name])
scene = api.Scene(path)
ctx = api.Context(0)
ctx.upload_scene(scene)
ctx.set_tuning([20, 6, 4, 10, 0, 1, 0])  # one wave lane, shadow kernel in line: whole-frame launches
ctx.film_clear()
ctx.render(seed=3, spp_total=scene.spp_squared())
ctx.synchronize()
print("rendered", name, ctx.counters()["camera_samples"], "camera samples")
