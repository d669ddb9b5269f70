"""How long does gb_upload_scene / film download take? (e2e overhead breakdown)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from goblin_b200 import api
name = sys.argv[1] if len(sys.argv) > 1 else "bunny"
scene = api.Scene(bench.scene_path(name))
ctx = api.Context(0)
for _ in range(3):
    ctx.upload_scene(scene)
t = time.perf_counter()
n = 20 if name == "bunny" else 3
for _ in range(n):
    ctx.upload_scene(scene)
dt = (time.perf_counter() - t) / n
print(name, "upload ms", dt * 1e3, "bytes", ctx.upload_bytes(), "GB/s", ctx.upload_bytes() / dt * 1e-9)
ctx.film_clear(); ctx.synchronize()
t = time.perf_counter()
for _ in range(n):
    f = ctx.film_download()
print("film download ms", (time.perf_counter() - t) / n * 1e3, f.nbytes)
