"""The logic of the default (4-wide) device walk, checked on the CPU before a GPU sees it.

tests/native/wide_walk.cpp emulates one lane of persistentTrace<ANY, false, true> with the same
functions the kernels are compiled from (goblin_b200/csrc/wide_node.h, rt_core.cuh).  Here it is
compared with the oracle's walk of the reference tree in the reference's order
(src/GoblinBVH.cpp:189-280): hit instance, primitive, t and epsilon bit for bit, any-hit flags
equal, and the per-thread stack never deeper than the bound the device allocates."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from goblin_b200 import api
from tests import oracle_port as op
from tests import util

LIB = os.path.join(util.ROOT, "tests", "native", "_build", "libwide_walk.so")
_lib = None


def ww():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", os.path.join(util.ROOT, "tests", "native")], check=True, capture_output=True)
        _lib = C.CDLL(LIB)
        P = C.c_void_p
        _lib.ww_trace.argtypes = [P, P, C.c_size_t, P, P, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return _lib


def wide_trace(scene, rays):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    hits = np.zeros(rays.shape[0], dtype=api.HIT_DTYPE)
    occ = np.zeros(rays.shape[0], dtype=np.uint8)
    deepest, bound = C.c_int(), C.c_int()
    assert ww().ww_trace(C.addressof(scene.desc), rays.ctypes.data, rays.shape[0], hits.ctypes.data, occ.ctypes.data,
                         C.byref(deepest), C.byref(bound)) == 0
    return hits, occ, deepest.value, bound.value


def random_rays(scene, n, seed, finite_frac=0.3):
    rng = np.random.default_rng(seed)
    wb = np.array(scene.desc.world_bound[:], np.float32)
    lo, hi = wb[:3], wb[3:]
    span = np.maximum(hi - lo, 1e-3)
    lo, hi = lo - 0.1 * span, hi + 0.1 * span
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    tgt = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    mint = np.full((n, 1), 1e-3, np.float32)
    maxt = np.where(rng.uniform(size=(n, 1)) < finite_frac, rng.uniform(0.05, 1.0, (n, 1)) * np.linalg.norm(span),
                    np.inf).astype(np.float32)
    return np.concatenate([o, d.astype(np.float32), mint, maxt], 1).astype(np.float32)


def _same(scene, rays):
    want = op.trace_closest(scene, rays)
    got, occ, deepest, bound = wide_trace(scene, rays)
    assert np.array_equal(got["inst"], want["inst"])
    assert np.array_equal(got["prim"], want["prim"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    assert np.array_equal(got["eps"].view(np.uint32), want["eps"].view(np.uint32))
    assert np.array_equal(occ, op.trace_any(scene, rays))
    assert deepest <= bound, (deepest, bound)
    return (want["inst"] >= 0).mean(), deepest, bound


@pytest.mark.parametrize("kind,args,json_name,n", [
    (None, (), util.TINY_PT, 200_000),             # every primitive kind, multi-instance top level
    ("bunny", (), "bunny_pt_small.json", 200_000),  # S1 geometry: a 17-level mesh tree
    ("spheres", (), "spheres_pt.json", 200_000),    # S3: 1,028 instances in the top-level tree
    ("grid", (160,), "grid_pt.json", 200_000),      # S4 at 51,200 triangles
    ("field", (5,), "field_pt.json", 200_000),      # S5 at 25 instances: both levels deep
    ("grid", (48,), "grid_pt.json", 100_000),       # the degenerate, very deep tree of the stack-limit test
])
def test_wide_walk_equals_the_reference_order_walk(built, kind, args, json_name, n):
    path = json_name if kind is None else os.path.join(util.gen_scene(kind, *args), json_name)
    scene = api.Scene(path)
    rays = random_rays(scene, n, 5)
    cam = np.random.default_rng(3).uniform(0, 1, (n // 2, 4)).astype(np.float32)
    cam[:, 0] *= scene.desc.film.xres
    cam[:, 1] *= scene.desc.film.yres
    hit_frac, deepest, bound = _same(scene, np.concatenate([rays, op.camera_rays(scene, cam)]))
    assert hit_frac > 0.05
    assert deepest > 0 or scene.desc.n_instances < 2


def test_wide_walk_on_the_sah_and_middle_trees(built):
    """Unbalanced trees (leaves at odd depths, wide nodes with empty slots)."""
    d = util.gen_scene("spheres")
    for accel in ("sah", "middle"):
        scene = api.Scene(os.path.join(d, "spheres_pt.json"), accel=accel)
        _same(scene, random_rays(scene, 100_000, 7))
    scene = api.Scene(os.path.join(util.gen_scene("bunny"), "bunny_pt_small.json"), accel="sah")
    _same(scene, random_rays(scene, 100_000, 8))


def test_axis_aligned_rays_on_box_planes(built):
    """Zero direction components with the origin exactly on box planes: the reference's slab test meets
    0 * inf = NaN there.  The wide walk evaluates the same comparisons on the boxes it tests, but not the
    intermediate node's own box; a NaN there makes the reference prune a subtree the wide walk still enters
    (include/goblin_b200.h, GB_TRACE_WIDE).  So: every hit the reference reports is reported identically or
    replaced by a NEARER one, and the rays that differ are listed and few."""
    d = util.gen_scene("grid", 160)
    scene = api.Scene(os.path.join(d, "grid_pt.json"))
    pos = scene.vert_pos().reshape(-1, 3)
    rng = np.random.default_rng(21)
    n = 60_000
    o = pos[rng.integers(0, len(pos), n)].copy()
    axis = rng.integers(0, 3, n)
    dirs = np.zeros((n, 3), np.float32)
    dirs[np.arange(n), axis] = rng.choice([-1.0, 1.0], n)
    two = rng.uniform(size=n) < 0.5
    other = (axis + 1) % 3
    dirs[np.arange(n)[two], other[two]] = rng.uniform(-1, 1, two.sum()).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o -= dirs * rng.uniform(0.0, 3.0, (n, 1)).astype(np.float32) * (rng.uniform(size=(n, 1)) < 0.7)
    rays = np.concatenate([o, dirs, np.full((n, 1), 1e-3, np.float32), np.full((n, 1), np.inf, np.float32)], 1)
    rays = rays.astype(np.float32)
    want = op.trace_closest(scene, rays)
    got, occ, deepest, bound = wide_trace(scene, rays)
    differ = (got["inst"] != want["inst"]) | (got["prim"] != want["prim"]) | (got["t"].view(np.uint32) != want["t"].view(np.uint32))
    print(f"axis-aligned on-plane rays: {differ.sum()} of {n} differ from the reference-order walk")
    for i in np.nonzero(differ)[0][:10]:
        print("  ray", i, rays[i], "reference", want[i], "wide", got[i])
    # never a lost hit, never a farther one
    assert not ((want["inst"] >= 0) & (got["inst"] < 0)).any()
    both = (want["inst"] >= 0) & (got["inst"] >= 0)
    assert (got["t"][both] <= want["t"][both]).all()
    assert differ.mean() < 0.02
    assert deepest <= bound
