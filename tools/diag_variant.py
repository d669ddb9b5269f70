"""GPU vs oracle Li on a scene variant, broken down by the instance the camera ray hits (debug aid)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from goblin_b200 import api
from tests import oracle_port as op, util
from tests import test_scene_variants as tv
v = sys.argv[1]
d = os.path.dirname(util.TINY_PT)
rng = np.random.default_rng(5)
for w, h in ((40, 24), (32, 16)):
    img = rng.uniform(0.05, 1.0, (h, w, 3)).astype(np.float32)
    img[h // 6:h // 6 + 3, w // 4:w // 4 + 4] *= 40.0
    api.write_rgb(os.path.join(d, f"_env_{w}x{h}.exr"), img)
p = os.path.join(d, "_diag.json")
sc = tv._variant(v)
depth = int(sys.argv[2]) if len(sys.argv) > 2 else sc["render_setting"]["max_ray_depth"]
sc["render_setting"]["max_ray_depth"] = depth
if len(sys.argv) > 3:  # keep bump / normal maps on one material only
    for m in sc["materials"]:
        if m["name"] != sys.argv[3]:
            m.pop("bumpmap", None); m.pop("normalmap", None)
json.dump(sc, open(p, "w"))
scene = api.Scene(p)
ctx = api.Context(0)
ctx.upload_scene(scene)
f = scene.desc.film
n = 20000
rows = rng.uniform(0, 1, (n, 4 + 7 * depth)).astype(np.float32)
rows[:, 0] = rng.uniform(f.sx0, f.sx1, n)
rows[:, 1] = rng.uniform(f.sy0, f.sy1, n)
got, want = ctx.li(rows), op.li(scene, rows)
close = np.isclose(got, want, rtol=2e-3, atol=2e-4).all(axis=1)
hits = op.trace_closest(scene, op.camera_rays(scene, rows[:, :4]))
print(v, "depth", depth, "close", close.mean())
for inst in np.unique(hits["inst"]):
    m = hits["inst"] == inst
    print("  primary inst", inst, "n", m.sum(), "close", round(close[m].mean(), 4))
bad = np.where(~close)[0][:4]
for b in bad:
    print("  ", hits["inst"][b], got[b], want[b])
os.remove(p)
