"""Shared helpers for the test-suite."""
import os
import subprocess

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TINY_PT = os.path.join(GOLDEN, "tiny", "tiny_pt.json")
TINY_AO = os.path.join(GOLDEN, "tiny", "tiny_ao.json")
REF_TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_tool")
SCENE_GEN = os.path.join(ROOT, "goblin_b200", "bin", "scene_gen")
GEN_DIR = os.path.join(ROOT, "scenes", "_gen")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def have_ref_tool():
    return os.path.exists(REF_TOOL) and os.access(REF_TOOL, os.X_OK)


def gen_scene(kind, *args):
    """Generate (once) one of the deterministic synthetic scenes; returns its directory."""
    d = os.path.join(GEN_DIR, kind + ("_" + "_".join(map(str, args)) if args else ""))
    marker = os.path.join(d, ".done")
    if not os.path.exists(marker):
        os.makedirs(d, exist_ok=True)
        subprocess.run([SCENE_GEN, kind, d, *map(str, args)], check=True)
        open(marker, "w").close()
    return d


def rel_mse(img, ref, eps=1e-2):
    """Relative MSE as used for converged-image comparisons: mean((a-b)^2 / (b^2 + eps))."""
    img = np.asarray(img, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + eps)))
