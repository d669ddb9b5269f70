"""CUDA path vs the oracle port on the same seeded inputs (the oracle is pinned to the reference
by tests/test_oracle_port.py).  Everything on the device goes through the C ABI."""
import numpy as np
import pytest

from goblin_b200 import api
from tests import oracle_port as op
from tests import util

pytestmark = pytest.mark.gpu


def _ctx(path):
    scene = api.Scene(path)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    return ctx, scene


def _random_rays(scene, n, seed, finite_frac=0.3):
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


def _check_traces(ctx, scene, rays, modes=("pair", "wide")):
    """Both walks of the traversal kernels (include/goblin_b200.h GB_TRACE_*): the default pair-node one and the
    4-wide one, bit for bit against the oracle's walk of the reference tree."""
    want = op.trace_closest(scene, rays)
    want_any = op.trace_any(scene, rays)
    for mode in modes:
        ctx.set_trace_mode(mode)
        got = ctx.trace_closest(rays)
        assert np.array_equal(got["inst"], want["inst"]), mode
        assert np.array_equal(got["prim"], want["prim"]), mode
        assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32)), mode
        assert np.array_equal(got["eps"].view(np.uint32), want["eps"].view(np.uint32)), mode
        assert np.array_equal(ctx.trace_any(rays), want_any), mode
    ctx.set_trace_mode("pair")
    return (want["inst"] >= 0).mean()


def test_tiny_random_rays_and_counters(built):
    ctx, scene = _ctx(util.TINY_PT)
    rays = _random_rays(scene, 200_000, 1)
    hit_frac = _check_traces(ctx, scene, rays)
    assert 0.2 < hit_frac < 1.0
    # traversal statistics are properties of the reference tree + order: identical on both sides
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.trace_closest(rays[:50_000])
    ctx.trace_any(rays[:50_000])
    g = ctx.counters()
    ctx.enable_counters(False)
    _, c1 = op.trace_closest(scene, rays[:50_000], counters=True)
    _, c2 = op.trace_any(scene, rays[:50_000], counters=True)
    for k in ("nodes_visited", "prims_tested", "instances_entered"):
        assert g[k] == c1[k] + c2[k], k
        assert g[k + "_any"] == c2[k], k
    ctx.close()


def test_film_same_seed_matches_oracle(built):
    """Same Philox seed -> the device film equals the oracle's film pixel by pixel (float
    reassociation only), and both traced exactly the same number of rays."""
    for path, spp in ((util.TINY_PT, 16), (util.TINY_AO, 4)):
        ctx, scene = _ctx(path)
        ctx.film_clear()
        ctx.render(seed=9, spp_total=spp)
        g = ctx.film_download()
        c, cnt, _ = op.render(scene, seed=9, spp_total=spp)
        gc = ctx.counters()
        assert gc["camera_samples"] == cnt["camera_samples"]
        assert abs(gc["rays_closest"] - cnt["rays_closest"]) <= 1e-5 * cnt["rays_closest"]
        assert abs(gc["rays_any"] - cnt["rays_any"]) <= 1e-5 * cnt["rays_any"]
        assert np.allclose(g[..., 3], c[..., 3], rtol=1e-4, atol=1e-5)
        close = np.isclose(g[..., :3], c[..., :3], rtol=2e-3, atol=1e-4).all(axis=2)
        assert close.mean() > 0.999, path
        assert abs(g[..., :3].sum() - c[..., :3].sum()) < 1e-4 * c[..., :3].sum()
        ctx.close()


def test_multi_wave_equals_single_wave(built):
    """Waves are an execution detail: a render cut into many small waves gives the same film."""
    ctx, scene = _ctx(util.TINY_PT)
    ctx.film_clear()
    ctx.render(seed=4, spp_total=16)
    one = ctx.film_download()
    ctx.set_wave_paths(3000)
    ctx.film_clear()
    ctx.render(seed=4, spp_total=16)
    many = ctx.film_download()
    assert np.allclose(one, many, rtol=1e-4, atol=1e-5)
    ctx.close()


def test_li_fresh_samples(built):
    ctx, scene = _ctx(util.TINY_PT)
    rng = np.random.default_rng(77)
    rows = rng.uniform(0, 1, (20_000, 4 + 7 * scene.desc.setting.max_ray_depth)).astype(np.float32)
    rows[:, 0] *= scene.desc.film.xres
    rows[:, 1] *= scene.desc.film.yres
    want = op.li(scene, rows)
    got = ctx.li(rows)
    close = np.isclose(got, want, rtol=2e-3, atol=2e-4).all(axis=1)
    assert close.mean() >= 0.999
    assert abs(got.sum() - want.sum()) < 2e-3 * want.sum()
    ctx.close()


@pytest.mark.parametrize("kind,args,json_name", [
    ("spheres", (), "spheres_pt.json"),       # S3: 1,028 sphere / disk instances, all materials, area lights
    ("grid", (160,), "grid_pt.json"),         # S4 at 51,200 triangles: one displaced-grid mesh with vn
    ("field", (5,), "field_pt.json"),         # S5 at 25 instances of one 81,920-triangle model
    ("bunny", (), "bunny_pt_small.json"),     # S1 geometry
])
def test_synthetic_scenes(built, kind, args, json_name):
    """The named synthetic scenes at sizes the oracle finishes in seconds: ray
    batches bit-exact, per-sample Li within float tolerance."""
    import os
    d = util.gen_scene(kind, *args)
    ctx, scene = _ctx(os.path.join(d, json_name))
    rays = _random_rays(scene, 100_000, 5)
    cam = np.random.default_rng(3).uniform(0, 1, (50_000, 4)).astype(np.float32)
    cam[:, 0] *= scene.desc.film.xres
    cam[:, 1] *= scene.desc.film.yres
    cam_rays = ctx.camera_rays(cam)
    assert np.array_equal(cam_rays.view(np.uint32), op.camera_rays(scene, cam).view(np.uint32))
    _check_traces(ctx, scene, np.concatenate([rays, cam_rays]))
    rng = np.random.default_rng(11)
    rows = rng.uniform(0, 1, (20_000, 4 + 7 * scene.desc.setting.max_ray_depth)).astype(np.float32)
    rows[:, 0] *= scene.desc.film.xres
    rows[:, 1] *= scene.desc.film.yres
    want = op.li(scene, rows)
    got = ctx.li(rows)
    close = np.isclose(got, want, rtol=2e-3, atol=2e-4).all(axis=1)
    assert close.mean() >= 0.998, f"{(~close).sum()} of {len(close)}"
    assert abs(got.sum() - want.sum()) <= 5e-3 * want.sum()
    ctx.close()


def test_large_ray_batches_bunny(built):
    """R* of SURVEY 8(d) on S1: 2^21 coherent camera rays + 2^21 incoherent rays (+ finite-extent
    shadow-like segments), closest and any-hit, bit-exact against the oracle."""
    import os
    d = util.gen_scene("bunny")
    ctx, scene = _ctx(os.path.join(d, "bunny_pt.json"))
    n = 1 << 21
    cam = np.random.default_rng(8).uniform(0, 1, (n, 4)).astype(np.float32)
    cam[:, 0] *= scene.desc.film.xres
    cam[:, 1] *= scene.desc.film.yres
    rays = np.concatenate([ctx.camera_rays(cam), _random_rays(scene, n, 9)])
    hit_frac = _check_traces(ctx, scene, rays)
    assert hit_frac > 0.3
    ctx.close()


def test_bvh_depth_and_stack_limits(built):
    """Stack sizing follows the trees of the uploaded scene.  A scene whose wide-walk stack column (three
    entries per wide level) would not leave room for four CTAs per SM is walked pair-wise instead, and
    gb_get_trace_mode says so: forced here by shrinking the limit (GB_MAX_WIDE_SMEM, read at upload)."""
    import os
    d = util.gen_scene("grid", 48)
    os.environ["GB_MAX_WIDE_SMEM"] = "4096"
    try:
        ctx, scene = _ctx(os.path.join(d, "grid_pt.json"))
    finally:
        del os.environ["GB_MAX_WIDE_SMEM"]
    ctx.set_trace_mode("wide")
    assert ctx.trace_mode() == "pair"  # asked for, but this scene's wide stack does not fit: walked pair-wise
    _check_traces(ctx, scene, _random_rays(scene, 50_000, 12, finite_frac=0.5), modes=("wide",))
    ctx.close()
    ctx, scene = _ctx(os.path.join(d, "grid_pt.json"))
    assert ctx.trace_mode() == "pair"
    ctx.set_trace_mode("wide")
    assert ctx.trace_mode() == "wide"
    rays = _random_rays(scene, 50_000, 13, finite_frac=0.5)
    _check_traces(ctx, scene, rays)
    ctx.enable_counters(True)
    ctx.reset_counters()
    ctx.trace_closest(rays[:20_000])
    g = ctx.counters()
    ctx.enable_counters(False)
    _, c = op.trace_closest(scene, rays[:20_000], counters=True)
    for k in ("nodes_visited", "prims_tested", "instances_entered"):
        assert g[k] == c[k], k
    ctx.close()


def test_axis_aligned_rays_nan_in_the_slab_test(built):
    """Rays with a zero direction component have an infinite reciprocal; when the origin also lies
    exactly on a box plane the reference's slab test meets 0 * inf = NaN and its ordered
    comparisons decide.  The kernels keep that comparison structure: hits, distances and node
    counts still equal the oracle's.  (A min / max formulation of the slab test with 3-input
    FMNMX3 and this case split off was measured: -6 ALU instructions per box, < 1 % faster.)"""
    import os
    d = util.gen_scene("grid", 160)
    scene = api.Scene(os.path.join(d, "grid_pt.json"))
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    pos = scene.vert_pos().reshape(-1, 3)
    rng = np.random.default_rng(21)
    n = 60_000
    o = pos[rng.integers(0, len(pos), n)].copy()          # coordinates that are box planes of the tree
    axis = rng.integers(0, 3, n)
    dirs = np.zeros((n, 3), np.float32)
    dirs[np.arange(n), axis] = rng.choice([-1.0, 1.0], n)
    two = rng.uniform(size=n) < 0.5                          # half of them: one zero component only
    other = (axis + 1) % 3
    dirs[np.arange(n)[two], other[two]] = rng.uniform(-1, 1, two.sum()).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o -= dirs * rng.uniform(0.0, 3.0, (n, 1)).astype(np.float32) * (rng.uniform(size=(n, 1)) < 0.7)
    # keep the on-plane coordinates exact where the direction component is zero
    rays = np.concatenate([o, dirs, np.full((n, 1), 1e-3, np.float32), np.full((n, 1), np.inf, np.float32)], 1)
    rays = rays.astype(np.float32)
    # a mesh instance with identity-like placement keeps zero components zero in object space too
    want, wc = op.trace_closest(scene, rays, counters=True)
    ctx.enable_counters(True)
    ctx.reset_counters()
    got = ctx.trace_closest(rays)
    gc = ctx.counters()
    assert np.array_equal(got["inst"], want["inst"]) and np.array_equal(got["prim"], want["prim"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    assert gc["nodes_visited"] == wc["nodes_visited"] and gc["prims_tested"] == wc["prims_tested"]
    ctx.enable_counters(False)
    want_any = op.trace_any(scene, rays)
    for mode in ("pair", "wide"):
        # the wide walk skips the intermediate nodes' own box tests (where a NaN would make the reference prune);
        # on this batch the CPU emulation of the wide walk (tests/test_wide_walk.py) agrees with the reference
        # on every ray, so the kernel must as well
        ctx.set_trace_mode(mode)
        got2 = ctx.trace_closest(rays)
        assert np.array_equal(got2["inst"], want["inst"]) and np.array_equal(got2["t"].view(np.uint32), want["t"].view(np.uint32)), mode
        assert np.array_equal(ctx.trace_any(rays), want_any), mode
    assert (want["inst"] >= 0).mean() > 0.2


@pytest.mark.parametrize("kind,json_name,n", [
    ("grid", "grid_pt.json", 1 << 20),    # S4 at full size: one mesh of 9,999,392 triangles, 24-level tree
    ("field", "field_pt.json", 1 << 19),  # S5 at full size: 729 instances of one 81,920-triangle model
])
def test_full_size_scenes_ray_batches(built, kind, json_name, n):
    """The two largest named scenes at BASELINE.json's sizes: coherent camera rays and incoherent
    segments, closest and any-hit, both walks bit-exact against the oracle's walk of the same flattened
    scene, plus a batch-independence property (a ray's answer does not depend on which rays share its warp).
    The reference itself is compared on these scenes in tests/test_gpu_vs_reference.py."""
    import os
    d = util.gen_scene(kind)
    ctx, scene = _ctx(os.path.join(d, json_name))
    cam = np.random.default_rng(18).uniform(0, 1, (n, 4)).astype(np.float32)
    cam[:, 0] *= scene.desc.film.xres
    cam[:, 1] *= scene.desc.film.yres
    rays = np.concatenate([ctx.camera_rays(cam), _random_rays(scene, n, 19)])
    hit_frac = _check_traces(ctx, scene, rays)
    assert hit_frac > 0.3
    # size-independent property: a second, different batch gives the same answers when traced together
    # with the first (no cross-talk between rays sharing a warp through the lane refill)
    half = rays[::2]
    assert np.array_equal(ctx.trace_closest(half)["t"].view(np.uint32), ctx.trace_closest(rays)["t"][::2].view(np.uint32))
    ctx.close()


def test_empty_scene_and_empty_mesh(built):
    """Edge cases of the scene format: no primitives at all (an empty top-level BVH: every ray misses,
    the film only gathers weights) and a mesh whose OBJ file is missing (the reference keeps an empty mesh:
    BVH::intersect returns false for it)."""
    import json
    import os
    d = os.path.dirname(util.TINY_PT)
    sc = json.load(open(util.TINY_PT))
    sc["primitives"] = []
    sc["lights"] = [l for l in sc["lights"] if l["type"] == "point"]
    scene = api.Scene(json_text=json.dumps(sc), scene_dir=d)
    assert scene.desc.n_instances == 0 and scene.desc.n_top_nodes == 0
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    rays = np.array([[0, 0, -5, 0, 0, 1, 1e-3, np.inf]] * 1000, np.float32)
    assert (ctx.trace_closest(rays)["inst"] == -1).all() and not ctx.trace_any(rays).any()
    ctx.film_clear()
    ctx.render(seed=1, spp_total=4)
    film = ctx.film_download()
    assert (film[..., :3] == 0).all() and film[..., 3].min() > 0
    ctx.close()
    sc = json.load(open(util.TINY_PT))
    sc["geometries"][0]["file"] = "models/_missing.obj"
    p = os.path.join(d, "_missing_mesh.json")
    json.dump(sc, open(p, "w"))
    try:
        ctx, scene = _ctx(p)
        _check_traces(ctx, scene, _random_rays(scene, 50_000, 3))
        ctx.close()
    finally:
        os.remove(p)
