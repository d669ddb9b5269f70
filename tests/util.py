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


def film_image(film):
    """Film::writeImage's normalisation: colour / weight (src/GoblinFilm.cpp:164-173)."""
    film = np.asarray(film, np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        img = film[..., :3] / film[..., 3:4]
    return np.nan_to_num(img)


def film_ttest(images, gold, rel_floor=2e-3):
    """Per-pixel variance-normalised two-sample test of a stack of independently seeded images
    against the reference's batch statistics in a tiny_film_*.npz.  Returns the t values of the
    pixels the reference lit.  The relative floor keeps pixels with (numerically) zero variance --
    an emitter seen directly has the same radiance in every sample -- from dividing a 1e-6
    rounding difference by ~0."""
    images = np.asarray(images, np.float64)
    b = images.shape[0]
    mean = images.mean(0)
    var_of_mean = images.var(0, ddof=1) / b
    ref = gold["mean"].astype(np.float64)
    ref_var = gold["var_of_mean"].astype(np.float64)
    floor = (rel_floor * np.maximum(ref, mean)) ** 2 + 1e-12
    t = (mean - ref) / np.sqrt(var_of_mean + ref_var + floor)
    return t[ref.sum(axis=2) > 0]
