#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference.

The reference (bachi95/Goblin) ships no tests and no golden data (SURVEY.md
section 4), so every known-answer vector here is produced by running the
reference's own code through oracle/_ref/ref_tool (built by `make -C oracle
ref` from /root/reference/src).  Run from the repo root in the build container:

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
    tiny/                 the scene files the vectors belong to (scene_gen tiny)
    tiny_dump.gbar        flattened BVHs, instance matrices, mesh arrays, ...
    tiny_rays.npz         ray batches + reference hits (closest and any-hit)
    tiny_li_pt.npz        PathTracer::Li on explicit sample values
    tiny_li_ao.npz        AORenderer::Li on explicit sample values
    tiny_film_pt.npz      converged reference film: mean and variance of the mean
    tiny_film_ao.npz
    post.npz              Goblin::bloom / toneMapping / the PPM writer on a small HDR image
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from goblin_b200 import gbar  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
GEN = os.path.join(ROOT, "goblin_b200", "bin", "scene_gen")
OUT = os.path.join(ROOT, "tests", "golden")


def run(*args):
    return subprocess.run(list(args), check=True, capture_output=True, text=True).stdout


def ref(cmd, scene, *rest):
    return run(REF, cmd, scene, *rest)


def ref_arrays(cmd, scene, arr, *extra):
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.f32")
        outp = os.path.join(td, "out.gbar")
        np.ascontiguousarray(arr, dtype=np.float32).tofile(inp)
        ref(cmd, scene, inp, outp, *extra)
        return {k: v.copy() for k, v in gbar.load(outp).items()}


def random_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def make_rays(scene, rng, dump):
    film = dump["film"]
    xres, yres = int(film[0]), int(film[1])
    # A: camera rays
    sm = np.zeros((2048, 4), np.float32)
    sm[:, 0] = rng.uniform(film[6], film[7], 2048)
    sm[:, 1] = rng.uniform(film[8], film[9], 2048)
    sm[:, 2:] = rng.uniform(0, 1, (2048, 2))
    cam = ref_arrays("camrays", scene, sm)["rays"]
    # B: the rays the reference's own path tracer shoots (bounce + shadow segments)
    depth = json.load(open(scene))["render_setting"]["max_ray_depth"]
    rows = np.zeros((192, 4 + 7 * depth), np.float32)
    rows[:, 0] = rng.uniform(0, xres, 192)
    rows[:, 1] = rng.uniform(0, yres, 192)
    rows[:, 2:] = rng.uniform(0, 1, (192, 2 + 7 * depth))
    rec = ref_arrays("li", scene, rows, "--record")
    recorded = rec["rays"][rec["rays"][:, 7] < np.inf][:1024]  # shadow segments (finite maxt)
    bounce = rec["rays"][(rec["rays"][:, 7] == np.inf) & (rec["rays"][:, 6] != np.float32(1e-3))][:2048]
    # C: incoherent random rays through the scene bounds, some with finite extent
    lo = dump["inst.aabb"][:, :3].min(0) - 0.5
    hi = dump["inst.aabb"][:, 3:].max(0) + 0.5
    lo = np.maximum(lo, -8.0)
    hi = np.minimum(hi, 8.0)
    n = 3072
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = random_dirs(rng, n)
    mint = np.where(rng.uniform(size=n) < 0.5, 0.0, 1e-3).astype(np.float32)
    maxt = np.where(rng.uniform(size=n) < 0.3, rng.uniform(0.1, 6.0, n), np.inf).astype(np.float32)
    rnd = np.concatenate([o, d, mint[:, None], maxt[:, None]], 1).astype(np.float32)
    # D: axis-parallel directions (zero components: infinite inverse directions)
    m = 512
    o = rng.uniform(lo, hi, (m, 3)).astype(np.float32)
    d = np.zeros((m, 3), np.float32)
    ax = rng.integers(0, 3, m)
    d[np.arange(m), ax] = rng.choice([-1.0, 1.0], m)
    two = rng.uniform(size=m) < 0.4  # a second non-zero component for some
    d[two, (ax[two] + 1) % 3] = rng.uniform(-1, 1, two.sum())
    axis = np.concatenate([o, d, np.zeros((m, 1), np.float32), np.full((m, 1), np.inf, np.float32)], 1)
    # E: un-normalised directions (the reference never renormalises in object space)
    k = 512
    o = rng.uniform(lo, hi, (k, 3)).astype(np.float32)
    d = random_dirs(rng, k) * rng.uniform(0.05, 20.0, (k, 1)).astype(np.float32)
    unn = np.concatenate([o, d, np.full((k, 1), 1e-3, np.float32), np.full((k, 1), np.inf, np.float32)], 1)
    rays = np.concatenate([cam, bounce, recorded, rnd, axis, unn.astype(np.float32)], 0).astype(np.float32)
    kinds = np.concatenate([np.full(len(cam), 0), np.full(len(bounce), 1), np.full(len(recorded), 2),
                            np.full(len(rnd), 3), np.full(len(axis), 4), np.full(len(unn), 5)]).astype(np.int32)
    res = ref_arrays("trace", scene, rays)
    res["rays"] = rays
    res["kind"] = kinds
    res["cam_samples"] = sm
    return res


def make_li(scene, rng, n, ao=False):
    cfg = json.load(open(scene))["render_setting"]
    film = gbar.load(os.path.join(OUT, "tiny_dump.gbar"))["film"]
    cols = 2 * cfg["ao_sample_num"] if ao else 7 * cfg["max_ray_depth"]
    rows = rng.uniform(0, 1, (n, 4 + cols)).astype(np.float32)
    rows[:, 0] = rng.uniform(film[6], film[7], n)
    rows[:, 1] = rng.uniform(film[8], film[9], n)
    res = ref_arrays("li", scene, rows)
    return {"samples": rows, "L": res["L"], "calls": res["calls"]}


def make_film(scene, batches, spp):
    films = []
    info = None
    with tempfile.TemporaryDirectory() as td:
        for b in range(batches):
            outp = os.path.join(td, "film.gbar")
            txt = ref("render", scene, outp, "--seed", str(1000 + b), "--spp", str(spp))
            info = json.loads(txt.split("REF_RESULT", 1)[1])
            f = gbar.load(outp)["film"]
            films.append(f[..., :3] / f[..., 3:4])
    films = np.stack(films).astype(np.float64)
    mean = films.mean(0)
    var_of_mean = films.var(0, ddof=1) / batches
    return {"mean": mean.astype(np.float32), "var_of_mean": var_of_mean.astype(np.float32),
            "batches": np.int32(batches), "spp_per_batch": np.int32(info["spp"]),
            "intersect_calls": np.int64(info["intersect_calls"]), "occluded_calls": np.int64(info["occluded_calls"]),
            "camera_samples": np.int64(info["camera_samples"])}


def make_post(rng):
    """The reference's image post-processing (ref_tool post) on a 40 x 28 HDR image with a few very
    bright pixels: bloom alone, tone mapping alone, and the PPM text Goblin::writeImage produces
    with and without tone mapping."""
    w, h = 40, 28
    img = rng.uniform(0.0, 1.5, (h, w, 3)).astype(np.float32)
    hot = rng.uniform(size=(h, w)) < 0.03
    img[hot] *= 60.0
    out = {"rgb": img, "bloom_radius": np.float32(0.2), "bloom_weight": np.float32(0.3)}
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.f32")
        img.tofile(inp)

        def post(name, radius, weight, tone):
            o = os.path.join(td, name)
            run(REF, "post", inp, o, str(w), str(h), repr(radius), repr(weight), str(tone))
            return o
        out["bloom"] = np.fromfile(post("a.f32", 0.2, 0.3, 0), np.float32).reshape(h, w, 3)
        out["bloom_wide"] = np.fromfile(post("w.f32", 2.5, 1.0, 0), np.float32).reshape(h, w, 3)
        out["tone"] = np.fromfile(post("b.f32", 0.0, 0.0, 1), np.float32).reshape(h, w, 3)
        out["ppm_plain"] = np.frombuffer(open(post("c.ppm", 0.0, 0.0, 0), "rb").read(), np.uint8)
        out["ppm_tone"] = np.frombuffer(open(post("d.ppm", 0.0, 0.0, 1), "rb").read(), np.uint8)
        out["ppm_bloom_tone"] = np.frombuffer(open(post("e.ppm", 0.2, 0.3, 1), "rb").read(), np.uint8)
    return out


def main():
    assert os.path.exists(REF), "build the reference oracle first: make -C oracle ref"
    tiny = os.path.join(OUT, "tiny")
    run(GEN, "tiny", tiny)
    pt = os.path.join(tiny, "tiny_pt.json")
    ao = os.path.join(tiny, "tiny_ao.json")
    ref("dump", pt, os.path.join(OUT, "tiny_dump.gbar"))
    dump = gbar.load(os.path.join(OUT, "tiny_dump.gbar"))
    rng = np.random.default_rng(20261018)
    if "--post-only" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "post.npz"), **make_post(np.random.default_rng(77)))
        return
    if "--films-only" in sys.argv:
        np.savez_compressed(os.path.join(OUT, "tiny_film_pt.npz"), **make_film(pt, 24, 256))
        np.savez_compressed(os.path.join(OUT, "tiny_film_ao.npz"), **make_film(ao, 16, 64))
        return
    np.savez_compressed(os.path.join(OUT, "tiny_rays.npz"), **make_rays(pt, rng, dump))
    np.savez_compressed(os.path.join(OUT, "tiny_li_pt.npz"), **make_li(pt, rng, 4096))
    np.savez_compressed(os.path.join(OUT, "tiny_li_ao.npz"), **make_li(ao, rng, 2048, ao=True))
    # many short batches rather than few long ones: the per-pixel variance of the mean is itself
    # estimated from the batches, and the two-sample t-test in tests/ needs it to be stable
    np.savez_compressed(os.path.join(OUT, "tiny_film_pt.npz"), **make_film(pt, 24, 256))
    np.savez_compressed(os.path.join(OUT, "tiny_film_ao.npz"), **make_film(ao, 16, 64))
    np.savez_compressed(os.path.join(OUT, "post.npz"), **make_post(np.random.default_rng(77)))
    for f in sorted(os.listdir(OUT)):
        p = os.path.join(OUT, f)
        if os.path.isfile(p):
            print(f, os.path.getsize(p))


if __name__ == "__main__":
    main()
