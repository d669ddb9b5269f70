"""The drop-in CLI: `g_ray scene.json` (src/g_ray.cpp:7-27) with the accelerated integrators on the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

from goblin_b200 import api
from tests import util

G_RAY = os.path.join(util.ROOT, "goblin_b200", "bin", "g_ray")


def _run(*args, cwd=None):
    return subprocess.run([G_RAY, *args], capture_output=True, text=True, cwd=cwd, timeout=600)


def _read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(v) for v in f.readline().split())
        assert float(f.readline()) < 0  # little endian
        return np.frombuffer(f.read(), np.float32).reshape(h, w, 3)[::-1]


def test_usage_and_loud_failures(built, tmp_path):
    r = _run()
    assert r.returncode == 0 and "Usage: g_ray scene.json" in r.stdout
    r = _run(str(tmp_path / "nope.json"))  # ContextLoader::load returns nullptr: nothing rendered
    assert "error reading scene file" in r.stderr and "render complete" not in r.stdout
    r = _run(util.TINY_PT, "--accel", "octree")
    assert r.returncode == 1 and "--accel" in r.stderr
    sc = json.load(open(util.TINY_PT))
    sc["render_setting"]["render_method"] = "sppm"  # what examples/bunny.json ships
    p = os.path.join(os.path.dirname(util.TINY_PT), "_cli_sppm.json")
    json.dump(sc, open(p, "w"))
    try:
        r = _run(p)
        assert r.returncode == 1 and "--method path_tracing" in r.stderr
    finally:
        os.remove(p)


@pytest.mark.gpu
def test_g_ray_renders_what_the_api_renders(built, tmp_path):
    out = str(tmp_path / "tiny.pfm")
    r = _run(util.TINY_PT, "--spp", "16", "--seed", "3", "--out", out, "--stats")
    assert r.returncode == 0, r.stderr
    assert "successfully loaded scene, start rendering..." in r.stdout and "render complete in" in r.stdout
    stats = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    scene = api.Scene(util.TINY_PT)
    x0, x1, y0, y1 = scene.sample_range()
    assert stats["camera_samples"] == (x1 - x0) * (y1 - y0) * 16 and stats["gpus"] == 1
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    ctx.film_clear()
    ctx.render(seed=3, spp_total=16)
    want = util.film_image(ctx.film_download())
    got = _read_pfm(out)
    assert got.shape == want.shape
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5)  # same kernels; float atomics may reorder the film sums


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [["--method", "ao"], ["--accel", "sah"], ["--depth", "3", "--gpus", "1"]])
def test_g_ray_flags(built, tmp_path, flags):
    out = str(tmp_path / "o.exr")
    r = _run(util.TINY_PT, "--spp", "4", "--out", out, *flags)
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(out) > 1000 and open(out, "rb").read(4) == b"\x76\x2f\x31\x01"


@pytest.mark.gpu
def test_g_ray_default_output_is_the_scene_name(built):
    """No --out and no film "file": <scene>.exr next to the scene (src/GoblinContextLoader.cpp:473-484)."""
    sc = json.load(open(util.TINY_PT))
    sc["render_setting"]["sample_per_pixel"] = 1
    sc["render_setting"]["seed"] = 5
    sc["camera"]["film"].pop("file", None)
    d = os.path.dirname(util.TINY_PT)
    p = os.path.join(d, "_cli_default.json")
    json.dump(sc, open(p, "w"))
    try:
        r = _run(p)
        assert r.returncode == 0, r.stderr
        assert os.path.exists(os.path.join(d, "_cli_default.exr"))
    finally:
        for q in (p, os.path.join(d, "_cli_default.exr")):
            if os.path.exists(q):
                os.remove(q)


@pytest.mark.gpu
def test_g_ray_two_gpus_equals_one(built, tmp_path):
    """--gpus 2: one host thread per GPU renders half of the per-pixel sample indices of a scene replica,
    the films are summed on the host (Film::mergeTile) and written through GPU 0: the same sample set,
    so the same image up to float summation order."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    a, b = str(tmp_path / "one.pfm"), str(tmp_path / "two.pfm")
    r1 = _run(util.TINY_PT, "--spp", "16", "--seed", "9", "--out", a)
    r2 = _run(util.TINY_PT, "--spp", "16", "--seed", "9", "--gpus", "2", "--out", b, "--stats")
    assert r1.returncode == 0 and r2.returncode == 0, r1.stderr + r2.stderr
    assert json.loads([l for l in r2.stdout.splitlines() if l.startswith("{")][-1])["gpus"] == 2
    assert np.allclose(_read_pfm(a), _read_pfm(b), rtol=1e-4, atol=1e-5)
