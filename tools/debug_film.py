"""Debug helper: GPU film vs oracle-port film with the same Philox seed."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from goblin_b200 import api
from tests import util, oracle_port as op

np.set_printoptions(precision=3, suppress=True, linewidth=250)
for path, spp in [(util.TINY_PT, 64), (util.TINY_PT, 1024), (util.TINY_PT, 4096)]:
    scene = api.Scene(path)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    ctx.film_clear()
    ctx.render(seed=11, spp_total=spp)
    g = ctx.film_download()
    c, cnt, _ = op.render(scene, seed=11, spp_total=spp)
    gc = ctx.counters()
    print(path, spp, "gpu counters", gc["camera_samples"], gc["rays_closest"], gc["rays_any"], gc["kernel_launches"], "port", cnt["camera_samples"], cnt["rays_closest"], cnt["rays_any"])
    d = np.abs(g - c)
    print(" weight max rel diff", (d[..., 3] / c[..., 3]).max(), "rgb rel", (d[..., :3].sum() / c[..., :3].sum()))
    bad = np.argwhere(d.max(axis=2) > 1e-3 * (np.abs(c).max(axis=2) + 1e-3))
    print(" bad pixels", len(bad), "of", g.shape[0] * g.shape[1])
    if len(bad):
        print(" rows", np.unique(bad[:, 0]), "cols", np.unique(bad[:, 1]))
        for y, x in bad[:8]:
            print("  ", y, x, g[y, x], c[y, x])
        print(" weight ratio by row", (g[..., 3].sum(1) / c[..., 3].sum(1)))
    ctx.close()
