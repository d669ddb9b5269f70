"""Pins the oracle (oracle/goblin_oracle.cpp, the CPU restatement) to the reference.

Expected values come from the UNMODIFIED reference run through oracle/_ref/ref_tool
(tests/golden/make_golden.py).  When the built reference binary is present (this container, or
shipped to the GPU box), a fresh cross-check on new random inputs runs as well."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from goblin_b200 import api, gbar
from tests import oracle_port as op
from tests import util


@pytest.fixture(scope="module")
def pt(built):
    return api.Scene(util.TINY_PT)


@pytest.fixture(scope="module")
def ao(built):
    return api.Scene(util.TINY_AO)


def test_closest_hit_bit_exact(pt):
    g = util.golden("tiny_rays.npz")
    h = op.trace_closest(pt, g["rays"])
    hit = g["hit"] != 0
    assert np.array_equal(h["inst"] >= 0, hit)
    assert np.array_equal(h["inst"][hit], g["inst"][hit])
    assert np.array_equal(h["prim"][hit], g["prim"][hit])
    assert np.array_equal(h["t"][hit].view(np.uint32), g["t"][hit].view(np.uint32))
    assert np.array_equal(h["eps"][hit].view(np.uint32), g["eps"][hit].view(np.uint32))


def test_any_hit_bit_exact(pt):
    g = util.golden("tiny_rays.npz")
    assert np.array_equal(op.trace_any(pt, g["rays"]) != 0, g["occluded"] != 0)


def test_fragments_bit_exact(pt):
    """Fragment position and normal after InstancedPrimitive's transform (columns 0-5 of ref_tool's
    14-float fragment record; 8-10 are dpdu)."""
    g = util.golden("tiny_rays.npz")
    hit = g["hit"] != 0
    fr = op.trace_fragments(pt, g["rays"])
    assert np.array_equal(fr[hit, :6].view(np.uint32), g["frag"][hit, :6].view(np.uint32))
    assert np.array_equal(fr[hit, 6:9].view(np.uint32), g["frag"][hit, 8:11].view(np.uint32))


def test_camera_rays_bit_exact(pt):
    g = util.golden("tiny_rays.npz")
    want = g["rays"][g["kind"] == 0]
    got = op.camera_rays(pt, g["cam_samples"])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_li_path_tracer(pt):
    """PathTracer::Li on the reference's own sample values: radiance within 1e-5 relative
    (sin / cos of the sampling maps are the only non-bit-exact inputs) and the exact number of
    Scene::intersect / Scene::occluded calls per sample (pins the path structure)."""
    g = util.golden("tiny_li_pt.npz")
    L, calls = op.li(pt, g["samples"], calls=True)
    assert np.array_equal(calls, g["calls"])
    assert np.allclose(L, g["L"], rtol=1e-5, atol=1e-6)


def test_li_ao(ao):
    g = util.golden("tiny_li_ao.npz")
    L, calls = op.li(ao, g["samples"], calls=True)
    assert np.array_equal(calls, g["calls"])
    assert np.array_equal(L, g["L"])


def test_render_matches_reference_film(pt):
    """go_render (Philox sampler) against the reference's own renders: 16 x 256 spp batches (the batch size of the
    reference renders behind the golden file: with smaller batches the per-pixel distributions are skewed enough --
    a firefly raises a pixel's mean and its variance estimate together -- to pull the mean of the statistic to about
    -0.15 .. -0.2 for ANY unbiased sampler, measured with and without the strata), the per-pixel two-sample statistic
    must look standard normal; reference-equivalent call counts per camera sample agree with what the reference's
    render made (linker --wrap counters)."""
    g = util.golden("tiny_film_pt.npz")
    imgs, total, calls, samples = [], None, [0, 0], 0
    for b in range(16):
        film, counters, ref_calls = op.render(pt, seed=400 + b, spp_total=256)
        imgs.append(util.film_image(film))
        total = film.astype(np.float64) if total is None else total + film
        calls = [calls[0] + ref_calls[0], calls[1] + ref_calls[1]]
        samples += counters["camera_samples"]
    assert samples == 16 * pt.camera_samples(256)
    assert util.rel_mse(util.film_image(total), g["mean"]) < 5e-3
    t = util.film_ttest(np.stack(imgs), g)
    assert abs(t.mean()) < 0.15 and 0.8 < t.std() < 1.2, (t.mean(), t.std())
    per_sample = (calls[0] + calls[1]) / samples
    ref_per_sample = (int(g["intersect_calls"]) + int(g["occluded_calls"])) / int(g["camera_samples"])
    assert abs(per_sample - ref_per_sample) < 0.01 * ref_per_sample


def test_render_threads_and_sharding(pt):
    a, _, _ = op.render(pt, seed=3, spp_total=4, threads=1)
    b, _, _ = op.render(pt, seed=3, spp_total=4, threads=4)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6)
    lo, _, _ = op.render(pt, seed=3, spp_total=4, spp_begin=0, spp_end=2, threads=1)
    both, _, _ = op.render(pt, seed=3, spp_total=4, spp_begin=2, spp_end=4, threads=1, film=lo)
    assert np.allclose(both, a, rtol=1e-5, atol=1e-6)


def test_counters_consistent(pt):
    g = util.golden("tiny_rays.npz")
    _, c = op.trace_closest(pt, g["rays"][:500], threads=1, counters=True)
    assert c["rays_closest"] == 500 and c["nodes_visited"] >= 500 and c["instances_entered"] > 0


@pytest.mark.skipif(not util.have_ref_tool(), reason="oracle/_ref/ref_tool not built")
def test_fresh_cross_check_against_reference_binary(pt):
    """New random rays and samples through the reference binary itself."""
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    n = 2000
    o = rng.uniform(-6, 6, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    rays = np.concatenate([o, d, np.full((n, 1), 1e-3, np.float32), np.full((n, 1), np.inf, np.float32)], 1)
    rows = rng.uniform(0, 1, (500, 4 + 7 * pt.desc.setting.max_ray_depth)).astype(np.float32)
    rows[:, 0] *= pt.desc.film.xres
    rows[:, 1] *= pt.desc.film.yres
    with tempfile.TemporaryDirectory() as td:
        rays.tofile(td + "/rays.f32")
        rows.tofile(td + "/rows.f32")
        subprocess.run([util.REF_TOOL, "trace", util.TINY_PT, td + "/rays.f32", td + "/t.gbar"], check=True,
                       capture_output=True)
        subprocess.run([util.REF_TOOL, "li", util.TINY_PT, td + "/rows.f32", td + "/l.gbar"], check=True,
                       capture_output=True)
        t = {k: v.copy() for k, v in gbar.load(td + "/t.gbar").items()}
        l = {k: v.copy() for k, v in gbar.load(td + "/l.gbar").items()}
    h = op.trace_closest(pt, rays)
    hit = t["hit"] != 0
    assert np.array_equal(h["inst"] >= 0, hit)
    assert np.array_equal(h["inst"][hit], t["inst"][hit]) and np.array_equal(h["prim"][hit], t["prim"][hit])
    assert np.array_equal(h["t"][hit].view(np.uint32), t["t"][hit].view(np.uint32))
    assert np.array_equal(op.trace_any(pt, rays) != 0, t["occluded"] != 0)
    L, calls = op.li(pt, rows, calls=True)
    assert np.array_equal(calls, l["calls"])
    assert np.allclose(L, l["L"], rtol=1e-5, atol=1e-6)
