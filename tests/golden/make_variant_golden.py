#!/usr/bin/env python
"""Golden vectors for the widened surface (textures, image based lights, masks, bump maps, orthographic camera):
per-sample radiance of the UNMODIFIED reference (oracle/_ref/ref_tool li) on the scene variants that
tests/test_scene_variants.py builds, written to tests/golden/variants_li.npz.  Run from the repo root in the
build container (needs /root/reference built by `make -C oracle ref`):

    python tests/golden/make_variant_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from goblin_b200 import api, gbar  # noqa: E402
from tests import test_scene_variants as tv  # noqa: E402
from tests import util  # noqa: E402

GOLDEN_VARIANTS = ["tex_filtered", "ibl", "img_ewa", "mask", "mask_ibl", "bump", "ortho"]
N = 600


def main():
    assert util.have_ref_tool(), "build the reference oracle first: make -C oracle ref"
    d = os.path.dirname(util.TINY_PT)
    env = tv.write_env_maps(d)
    out = {}
    try:
        for v in GOLDEN_VARIANTS:
            path = os.path.join(d, f"_golden_{v}.json")
            json.dump(tv._variant(v), open(path, "w"))
            scene = api.Scene(path)
            f = scene.desc.film
            rng = np.random.default_rng(zlib.crc32(("golden_" + v).encode()))
            rows = rng.uniform(0, 1, (N, tv._row_floats(scene))).astype(np.float32)
            rows[:, 0] = rng.uniform(f.sx0, f.sx1, N)
            rows[:, 1] = rng.uniform(f.sy0, f.sy1, N)
            with tempfile.TemporaryDirectory() as td:
                rows.tofile(td + "/rows.f32")
                subprocess.run([util.REF_TOOL, "li", path, td + "/rows.f32", td + "/l.gbar"], check=True, capture_output=True)
                ref = {k: a.copy() for k, a in gbar.load(td + "/l.gbar").items()}
            out[v + ".rows"] = rows
            out[v + ".L"] = ref["L"]
            out[v + ".calls"] = ref["calls"]
            os.remove(path)
    finally:
        for p in env:
            os.remove(p)
    dst = os.path.join(ROOT, "tests", "golden", "variants_li.npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
