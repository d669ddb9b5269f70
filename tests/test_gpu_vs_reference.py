"""The CUDA path against the UNMODIFIED reference, live, on BASELINE.json's scenes at their real size.

oracle/_ref/ref_tool (the reference's own objects + a harness) ships to the GPU box, so these tests do
not go through the oracle port: the reference loads the scene once (`ref_tool session`), generates the
camera rays, runs its own path tracer while the harness records every ray it hands to Scene::intersect /
Scene::occluded (linker --wrap), traces all of them, evaluates Li on explicit samples and renders converged
images; the GPU gets the same files through the C ABI.

north_star's gates: flattened BVH bit-exact (tests/test_host_scene.py, run where the reference sources
are), hit ids agree on >= 99.999 % of the exported rays with t within 1e-5 relative, converged 4096-spp
images relMSE < 1e-3 plus a per-pixel variance-normalised test.  What is observed is stronger -- ids, t and
epsilon bit-exact on every ray -- and each run appends what it measured to gpurun_out/parity_vs_reference.jsonl
(copied to profiles/ for the record)."""
import json
import os
import time

import numpy as np
import pytest

from goblin_b200 import api, gbar
from tests import util

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not util.have_ref_tool(), reason="oracle/_ref/ref_tool not built")]

LOG = os.path.join(util.ROOT, "gpurun_out", "parity_vs_reference.jsonl")


def _record(entry):
    print("PARITY", json.dumps(entry))
    try:
        os.makedirs(os.path.dirname(LOG), exist_ok=True)
        with open(LOG, "a") as f:
            f.write(json.dumps(entry) + "\n")
    except OSError:
        pass


def _compare_hits(got, want, what, entry):
    """ids >= 99.999 %, t within 1e-5 relative (north_star); everything that differs is listed."""
    n = len(want["t"])
    hit_r, hit_g = want["inst"] >= 0, got["inst"] >= 0
    same_id = (got["inst"] == want["inst"]) & ((got["prim"] == want["prim"]) | ~hit_r)
    both = hit_r & hit_g
    rel = np.zeros(n)
    rel[both] = np.abs(got["t"][both].astype(np.float64) - want["t"][both]) / np.maximum(np.abs(want["t"][both]), 1e-30)
    t_bits_equal = (got["t"].view(np.uint32) == want["t"].view(np.uint32)) | ~both
    bad = np.nonzero(~same_id)[0]
    entry[what] = {"rays": int(n), "hit_fraction": float(hit_r.mean()), "id_mismatches": int(len(bad)),
                   "id_agreement": float(same_id.mean()), "t_bit_exact": float(t_bits_equal.mean()),
                   "t_max_rel_err": float(rel.max()) if n else 0.0,
                   "mismatch_list": [{"ray": int(i), "reference": [int(want["inst"][i]), int(want["prim"][i]), float(want["t"][i])],
                                      "gpu": [int(got["inst"][i]), int(got["prim"][i]), float(got["t"][i])]} for i in bad[:20]]}
    assert same_id.mean() >= 0.99999, f"{what}: {len(bad)} of {n} hit ids differ from the reference"
    assert rel.max() <= 1e-5, f"{what}: t differs by {rel.max():.3g} relative"
    # a mismatching id must be a tie: the same distance to within the gate
    assert (rel[bad] <= 1e-5).all()


def _ref_hits(g):
    out = np.zeros(len(g["t"]), dtype=api.HIT_DTYPE)
    out["t"], out["eps"], out["inst"], out["prim"] = g["t"], g["eps"], g["inst"], g["prim"]
    return out


@pytest.mark.parametrize("name,kind,json_name,log2n,n_li", [
    ("S1", "bunny", "bunny_pt.json", 22, 1 << 18),     # bunny.json layout, 81,920-triangle stand-in, glass + Lambert
    ("S3", "spheres", "spheres_pt.json", 22, 1 << 18),  # 1,028 sphere / disk instances, all three materials, area lights
    ("S4", "grid", "grid_pt.json", 22, 1 << 17),        # the 9,999,392-triangle mesh of BASELINE.json config 4
    ("S5", "field", "field_pt.json", 21, 1 << 16),      # 729 instances of the 81,920-triangle model (config 5)
])
def test_reference_exported_rays_and_li(built, tmp_path, name, kind, json_name, log2n, n_li):
    """R* of SURVEY 8(d) from the reference itself: 2^22 primary rays (Camera::generateRay), 2^22 rays the
    reference's path tracer handed to Scene::intersect after the first hit (bounce / MIS / attenuation rays) and
    2^22 of its shadow segments (Scene::occluded), each traced by the reference (closest AND any) and by both
    walks of the CUDA kernels.  Then PathTracer::Li (src/GoblinPathtracer.cpp:50-179) on explicit samples."""
    path = os.path.join(util.gen_scene(kind), json_name)
    n = 1 << log2n
    td = str(tmp_path)
    t0 = time.time()
    scene = api.Scene(path)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    load_s = time.time() - t0
    rng = np.random.default_rng(31)
    cam = rng.uniform(0, 1, (n, 4)).astype(np.float32)
    cam[:, 0] *= scene.desc.film.xres
    cam[:, 1] *= scene.desc.film.yres
    depth = scene.desc.setting.max_ray_depth
    rows = rng.uniform(0, 1, (n_li, 4 + 7 * depth)).astype(np.float32)
    rows[:, 0] *= scene.desc.film.xres
    rows[:, 1] *= scene.desc.film.yres
    cam.tofile(td + "/cam.f32")
    rows.tofile(td + "/rows.f32")
    t0 = time.time()
    with util.RefSession(path, cwd=td) as ref:
        ref_load_s = time.time() - t0
        ref.run("camrays", td + "/cam.f32", td + "/camrays.gbar")
        cam_rays = gbar.load(td + "/camrays.gbar")["rays"]
        # the reference's own integrator on its own samples, every ray recorded until 2^n of each kind are held
        ref.run("li", f"@{64 * n}:17:{2 * n}:{n}", td + "/rec.gbar", "--record-mt")
        rec = gbar.load(td + "/rec.gbar")
        rays, rk = rec["rays"], rec["kind"]
        origin = cam_rays[0, :3]
        primary = (rk == 0) & np.all(rays[:, :3] == origin, axis=1)  # pinhole camera: primaries leave one point
        bounce = rays[(rk == 0) & ~primary][:n]
        shadow = rays[rk == 1][:n]
        assert len(bounce) >= n // 2 and len(shadow) >= n // 4, (len(bounce), len(shadow))
        batch = np.ascontiguousarray(np.concatenate([cam_rays, bounce, shadow]), dtype=np.float32)
        batch.tofile(td + "/batch.f32")
        ref.run("trace", td + "/batch.f32", td + "/trace.gbar", "--no-frag")
        want = gbar.load(td + "/trace.gbar")
        ref.run("li", td + "/rows.f32", td + "/li.gbar")
        want_L = gbar.load(td + "/li.gbar")["L"]
    ref_s = time.time() - t0
    entry = {"scene": name, "file": json_name, "triangles": int(scene.desc.n_tris), "instances": int(scene.desc.n_instances),
             "load_s": {"gpu_host_side": load_s, "reference": ref_load_s}, "reference_session_s": ref_s}
    try:
        # ---- camera rays: same arithmetic on both sides up to libm vs CUDA trig in the orientation
        got_cam = ctx.camera_rays(cam)
        entry["camera_rays"] = {"n": int(n), "max_abs_diff": float(np.abs(got_cam[:, :6] - cam_rays[:, :6]).max())}
        assert entry["camera_rays"]["max_abs_diff"] <= 1e-6
        # ---- the exported batches, both walks
        want_hits = _ref_hits(want)
        seg = {"primary": slice(0, n), "bounce": slice(n, n + len(bounce)), "shadow": slice(n + len(bounce), len(batch))}
        for mode in ("pair", "wide"):
            ctx.set_trace_mode(mode)
            got = ctx.trace_closest(batch)
            occ = ctx.trace_any(batch)
            for what, sl in seg.items():
                _compare_hits(got[sl], want_hits[sl], f"{mode}.closest.{what}", entry)
                agree = occ[sl] == (want["occluded"][sl] != 0)
                entry[f"{mode}.any.{what}"] = {"rays": int(agree.size), "agreement": float(agree.mean()),
                                               "occluded_fraction": float(occ[sl].mean())}
                assert agree.mean() >= 0.99999, f"{mode}.any.{what}"
        ctx.set_trace_mode("pair")
        # ---- per-sample radiance
        got_L = ctx.li(rows)
        close = np.isclose(got_L, want_L, rtol=2e-3, atol=2e-4).all(axis=1)
        err = np.abs(got_L - want_L).max(axis=1)
        worst = np.argsort(err)[::-1][:10]
        entry["li"] = {"samples": int(n_li), "tolerance": "rtol 2e-3 + atol 2e-4 per channel", "inside": float(close.mean()),
                       "outside": int((~close).sum()), "mean_gpu": float(got_L.mean()), "mean_reference": float(want_L.mean()),
                       "median_abs_err": float(np.median(err)), "p999_abs_err": float(np.quantile(err, 0.999)),
                       "worst": [{"sample": int(i), "gpu": got_L[i].tolist(), "reference": want_L[i].tolist()} for i in worst]}
        assert close.mean() >= 0.995, f"{(~close).sum()} of {n_li} samples outside tolerance"
        assert abs(got_L.mean() - want_L.mean()) <= 3e-3 * want_L.mean()
    finally:
        _record(entry)  # what was measured is kept even when a gate fails
    ctx.close()


def _lowres(path, xres, yres, tag):
    """The same scene file with a smaller film, written next to the original (mesh paths are relative)."""
    sc = json.load(open(path))
    sc["camera"]["film"]["resolution"] = [xres, yres]
    out = path.replace(".json", f"_{tag}.json")
    with open(out, "w") as f:
        json.dump(sc, f)
    return out


@pytest.mark.parametrize("name,kind,json_name,xres,yres,ref_batches,gpu_batches,gpu_spp", [
    # S3 (glass + mirrors + four area lights) is not converged at 4096 spp: two independent, unbiased 4096-spp
    # estimates differ by relMSE 3.4e-3 from noise alone (measured: the oracle port against the reference, t-test
    # mean -0.06 / std 1.05).  So the reference renders 16,384 spp and the GPU 65,536; the gate stays 1e-3.
    ("S3", "spheres", "spheres_pt.json", 64, 64, 64, 64, 1024),
    ("S4", "grid", "grid_pt.json", 96, 54, 16, 16, 256),  # the full 9,999,392-triangle mesh, reduced film, 4096 spp
])
def test_converged_image_against_live_reference(built, tmp_path, name, kind, json_name, xres, yres, ref_batches,
                                                gpu_batches, gpu_spp):
    """Converged images: GPU (Philox sampler) against the reference rendering now (mt19937 stratified sampler),
    both as independently seeded batches (the reference's of 256 spp): relMSE < 1e-3 and a per-pixel two-sample
    statistic that must look standard normal (Renderer::render + Film, src/GoblinRenderer.cpp:99-126,
    src/GoblinFilm.cpp:61-173)."""
    path = _lowres(os.path.join(util.gen_scene(kind), json_name), xres, yres, "lowres")
    td = str(tmp_path)
    refs, ref_rate = [], []
    t0 = time.time()
    with util.RefSession(path, cwd=td) as ref:
        for b in range(ref_batches):
            r = ref.render(td + "/f.gbar", 700 + b, 256)
            ref_rate.append(r["msamples_per_s"])
            refs.append(util.film_image(gbar.load(td + "/f.gbar")["film"]))
    refs = np.stack(refs)
    gold = {"mean": refs.mean(0), "var_of_mean": refs.var(0, ddof=1) / len(refs)}
    scene = api.Scene(path)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    imgs, total = [], None
    for b in range(gpu_batches):
        ctx.film_clear()
        ctx.render(seed=1300 + b, spp_total=gpu_spp)
        film = ctx.film_download()
        total = film.astype(np.float64) if total is None else total + film
        imgs.append(util.film_image(film))
    ctx.close()
    total = util.film_image(total)
    t = util.film_ttest(np.stack(imgs), gold)
    entry = {"scene": name, "converged": f"{xres}x{yres}, reference {ref_batches} x 256 spp, gpu {gpu_batches} x {gpu_spp} spp",
             "rel_mse": util.rel_mse(total, gold["mean"]),
             "t_mean": float(t.mean()), "t_std": float(t.std()), "t_gt_4.5": float((np.abs(t) > 4.5).mean()),
             "reference_msamples_per_s": float(np.mean(ref_rate)), "seconds": time.time() - t0}
    _record(entry)
    assert entry["rel_mse"] < 1e-3
    assert abs(t.mean()) < 0.1, f"biased: mean t = {t.mean():.3f}"
    assert 0.4 < t.std() < 1.15, f"t spread {t.std():.3f}"
    assert (np.abs(t) > 4.5).mean() < 1e-3
