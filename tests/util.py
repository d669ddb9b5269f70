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


class RefSession:
    """`ref_tool session scene.json`: the UNMODIFIED reference (oracle/_ref) with the scene loaded once,
    then commands from stdin.  Loading the 10 M-triangle scene takes the reference about a minute; the
    tests that compare with it on that scene do everything in one session."""

    def __init__(self, scene_path, cwd=None):
        self.proc = subprocess.Popen([REF_TOOL, "session", scene_path], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                     stderr=subprocess.PIPE, text=True, cwd=cwd)
        self._expect("SESSION_READY")

    def _expect(self, prefix):
        lines = []
        while True:
            line = self.proc.stdout.readline()
            if not line:
                raise RuntimeError("ref_tool session ended: " + "".join(lines[-5:]) + self.proc.stderr.read()[-2000:])
            lines.append(line)
            if line.startswith(prefix):
                return line, lines

    def run(self, *words):
        """One command (e.g. run("trace", rays_path, out_path)); returns the lines it printed."""
        self.proc.stdin.write(" ".join(str(w) for w in words) + "\n")
        self.proc.stdin.flush()
        line, lines = self._expect("SESSION_DONE")
        if int(line.split()[1]) != 0:
            raise RuntimeError(f"ref_tool {words[0]} failed: {line}")
        return lines

    def render(self, out, seed, spp, threads=0):
        """Timed RenderContext::render(); returns the REF_RESULT record."""
        import json
        words = ["render", out, "--seed", seed, "--spp", spp] + (["--threads", threads] if threads else [])
        lines = self.run(*words)
        rec = [l for l in lines if "REF_RESULT" in l]
        return json.loads(rec[-1].split("REF_RESULT", 1)[1])

    def close(self):
        if self.proc.poll() is None:
            try:
                self.proc.stdin.write("quit\n")
                self.proc.stdin.flush()
                self.proc.wait(timeout=60)
            except Exception:
                self.proc.kill()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
