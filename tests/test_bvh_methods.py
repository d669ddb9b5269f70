"""BVH split methods beside the reference's equal_count.

`middle` is the reference's other method (src/GoblinBVH.cpp:124-134; no caller of the reference
selects it, so there is no reference dump to pin it to: its split rule is checked structurally).
`sah` is this library's non-parity fast tree.  Both use the reference's node format, so the
checks are: the tree is well-formed, every traversal returns the same closest-hit distance as the
reference's tree, and the GPU walks them exactly as the oracle does."""
import math

import numpy as np
import pytest

from goblin_b200 import api
from tests import oracle_port as op
from tests import util


def _boxes(n, seed, big_frac=0.02):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-10, 10, (n, 3))
    r = rng.uniform(0.01, 0.3, (n, 3))
    r[rng.uniform(size=n) < big_frac] *= 30  # a few large boxes: where SAH and the median differ
    return np.concatenate([c - r, c + r], 1).astype(np.float32)


def _walk(nodes, order, boxes, method):
    """Checks pre-order structure, boxes and the split rule; returns (leaf slots visited, depth)."""
    n = len(boxes)
    seen = np.zeros(n, bool)
    centers = (0.5 * (boxes[:, :3] + boxes[:, 3:])).astype(np.float32)
    deepest = 0

    def rec(i, depth):
        nonlocal deepest
        deepest = max(deepest, depth)
        nd = nodes[i]
        if nd["nprims"]:
            sl = np.arange(nd["offset"], nd["offset"] + nd["nprims"])
            assert not seen[sl].any()
            seen[sl] = True
            b = boxes[order[sl]]
            assert np.array_equal(nd["bmin"], b[:, :3].min(0)) and np.array_equal(nd["bmax"], b[:, 3:].max(0))
            return sl
        left, right = rec(i + 1, depth + 1), rec(int(nd["offset"]), depth + 1)
        l, r = nodes[i + 1], nodes[int(nd["offset"])]
        assert np.array_equal(nd["bmin"], np.minimum(l["bmin"], r["bmin"]))
        assert np.array_equal(nd["bmax"], np.maximum(l["bmax"], r["bmax"]))
        assert left[-1] + 1 == right[0]  # leaf slots are appended in range order
        ax = int(nd["axis"])
        cl, cr = centers[order[left], ax], centers[order[right], ax]
        assert cl.max() <= cr.min()  # every method partitions along its axis
        if method == "equal_count":
            assert len(left) == (len(left) + len(right)) // 2
        elif method == "middle":
            call = centers[order[np.concatenate([left, right])]]
            mid = np.float32(0.5) * (call[:, ax].min() + call[:, ax].max())
            assert (cl.max() < mid <= cr.min()) or len(left) == (len(left) + len(right)) // 2
        return np.concatenate([left, right])

    import sys
    sys.setrecursionlimit(10000)
    rec(0, 0)
    assert seen.all()
    return deepest


@pytest.mark.parametrize("method", ["equal_count", "middle", "sah"])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 1000, 20000])
def test_tree_is_well_formed(built, method, n):
    boxes = _boxes(n, n)
    nodes, order = api.bvh_build(boxes, method)
    assert len(nodes) == 2 * n - 1
    assert np.array_equal(np.sort(order), np.arange(n))
    depth = _walk(nodes, order, boxes, method)
    if method == "sah":  # the stack in shared memory is sized by the depth: it stays bounded
        assert depth <= math.ceil(math.log2(max(n, 1))) + 3


def test_equal_count_is_the_default(built):
    boxes = _boxes(5000, 3)
    a, ao = api.bvh_build(boxes)
    b, bo = api.bvh_build(boxes, "equal_count")
    assert a.tobytes() == b.tobytes() and np.array_equal(ao, bo)


def test_coincident_centres_make_one_leaf(built):
    boxes = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (5, 1))
    for method in ("middle", "sah"):
        nodes, order = api.bvh_build(boxes, method)
        assert len(nodes) == 1 and nodes[0]["nprims"] == 5


def test_sah_is_cheaper_and_bad_method_is_refused(built):
    boxes = _boxes(20000, 9)

    def sah_cost(nodes):
        ext = nodes["bmax"].astype(np.float64) - nodes["bmin"]
        area = ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 2] * ext[:, 0]
        return area[nodes["nprims"] == 0].sum() / area[0]

    assert sah_cost(api.bvh_build(boxes, "sah")[0]) < 0.8 * sah_cost(api.bvh_build(boxes, "equal_count")[0])
    import ctypes as C
    cnt = C.c_uint32()
    assert api.lib().gb_bvh_build_method(boxes.ctypes.data, 4, 7, None, C.byref(cnt), None) == 1  # GB_ERR_INVALID
    opt = api.LoadOptions(bvh_method=9)
    h = C.c_void_p()
    assert api.lib().gb_scene_load_json_ex(util.TINY_PT.encode(), C.byref(opt), C.byref(h)) == 1  # GB_ERR_INVALID


def _rays(scene, n, seed):
    rng = np.random.default_rng(seed)
    wb = np.array(scene.desc.world_bound[:], np.float32)
    lo, hi = wb[:3], wb[3:]
    o = rng.uniform(lo, hi, (n, 3))
    d = rng.uniform(lo, hi, (n, 3)) - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d, np.full((n, 1), 1e-3), np.full((n, 1), np.inf)], 1).astype(np.float32)


@pytest.mark.parametrize("kind,args,json_name", [("tiny", [], None), ("spheres", [], "spheres_pt.json"),
                                                 ("bunny", [], "bunny_pt.json")])
def test_same_hits_on_every_tree(built, kind, args, json_name):
    """Closest-hit distance, occlusion and (away from exact ties) ids do not depend on the tree."""
    path = util.TINY_PT if json_name is None else util.gen_scene(kind, *args) + "/" + json_name
    ref_scene = api.Scene(path)
    rays = _rays(ref_scene, 20000, 5)
    want = op.trace_closest(ref_scene, rays)
    want_any = op.trace_any(ref_scene, rays)
    for method in ("middle", "sah"):
        sc = api.Scene(path, accel=method)
        assert sc.desc.n_top_nodes == ref_scene.desc.n_top_nodes
        got = op.trace_closest(sc, rays)
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), method
        assert ((got["inst"] == want["inst"]) & (got["prim"] == want["prim"])).mean() > 0.9999, method
        assert np.array_equal(op.trace_any(sc, rays), want_any), method


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["middle", "sah"])
@pytest.mark.parametrize("kind,args,json_name", [("spheres", [], "spheres_pt.json"), ("bunny", [], "bunny_pt.json"),
                                                 ("field", [5], "field_pt.json")])
def test_gpu_walks_the_other_trees_like_the_oracle(built, method, kind, args, json_name):
    """Same kernels, another tree: hits, distances and the traversal counters equal the oracle's
    walk of that tree; distances also equal the parity tree's."""
    path = util.gen_scene(kind, *args) + "/" + json_name
    sc = api.Scene(path, accel=method)
    ctx = api.Context(0)
    ctx.upload_scene(sc)
    rays = _rays(sc, 300_000, 11)
    want, wc = op.trace_closest(sc, rays, counters=True)
    ctx.enable_counters(True)
    ctx.reset_counters()
    got = ctx.trace_closest(rays)
    gc = ctx.counters()
    for k in ("inst", "prim"):
        assert np.array_equal(got[k], want[k])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    for k in ("nodes_visited", "prims_tested", "instances_entered"):
        assert gc[k] == wc[k], k
    ctx.enable_counters(False)
    assert np.array_equal(ctx.trace_any(rays), op.trace_any(sc, rays))
    parity = op.trace_closest(api.Scene(path), rays)
    assert np.array_equal(got["t"].view(np.uint32), parity["t"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_image_does_not_depend_on_the_tree(built):
    """Same seed, same samples: the film rendered on the SAH tree equals the parity tree's up to
    the handful of samples whose path crosses an exact tie."""
    films = []
    for method in ("equal_count", "sah"):
        sc = api.Scene(util.TINY_PT, accel=method)
        ctx = api.Context(0)
        ctx.upload_scene(sc)
        ctx.film_clear()
        ctx.render(seed=3, spp_total=16)
        films.append(util.film_image(ctx.film_download()))
    assert util.rel_mse(films[1], films[0]) < 1e-6
