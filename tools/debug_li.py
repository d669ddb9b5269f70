"""Debug helper: where do gb_li results leave the golden reference values?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from goblin_b200 import api
from tests import util

def run(path, gold, ao):
    scene = api.Scene(path)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    g = util.golden(gold)
    got = ctx.li(g["samples"])
    ref = g["L"]
    rays = ctx.camera_rays(g["samples"][:, :4])
    hits = ctx.trace_closest(rays)
    models = scene.models()
    insts = scene.instances()
    kind = np.array([models[insts[i].model].kind if i >= 0 else -1 for i in hits["inst"]])
    mat = np.array([models[insts[i].model].material if i >= 0 else -1 for i in hits["inst"]])
    bad = ~np.isclose(got, ref, rtol=2e-3, atol=2e-4).all(axis=1)
    print(gold, "bad", bad.sum(), "of", len(bad), "mean got/ref", got.mean(), ref.mean())
    for k in np.unique(kind):
        m = kind == k
        print("  first-hit kind", k, "n", m.sum(), "bad", bad[m].sum())
    for k in np.unique(hits["inst"]):
        m = hits["inst"] == k
        print("  first-hit inst", k, "n", m.sum(), "bad", bad[m].sum(), "mat", mat[m][0], "kind", kind[m][0])
    calls = g["calls"]
    for c in np.unique(calls[:, 0]):
        m = calls[:, 0] == c
        print("  intersect calls", c, "n", m.sum(), "bad", bad[m].sum())
    idx = np.nonzero(bad)[0][:12]
    for i in idx:
        print("   sample", i, "inst", hits["inst"][i], "got", got[i], "ref", ref[i], "calls", calls[i])
    ctx.close()

run(util.TINY_PT, "tiny_li_pt.npz", False)
run(util.TINY_AO, "tiny_li_ao.npz", True)
