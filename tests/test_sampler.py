"""The counter-based sampler that replaces GoblinSampler (north_star): Philox values placed in the strata the
reference's per-pixel sampler uses (Sampler::requestSamples, src/GoblinSampler.cpp:108-197: every 1-D dimension in
sample_per_pixel jittered strata, every 2-D dimension on a root x root jittered grid, each dimension shuffled
independently across the samples of the pixel).  What must hold: each (pixel, dimension) visits every stratum exactly
once, different pixels / dimensions get unrelated orders, and a sample is worth what a reference sample is worth --
the per-pixel variance of an image equals the reference's."""
import os

import numpy as np
import pytest

from goblin_b200 import api
from tests import oracle_port as op
from tests import util


def test_strata_are_permutations():
    for spp in (4, 16, 100, 256, 1024):
        seen = set()
        for pixel in (0, 1, 77, 12345678901):
            for dim in (0, 1, 4, 5, 39, 0x10000, 0x20003):
                st = op.strata(9, pixel, dim, spp)
                assert np.array_equal(np.sort(st), np.arange(spp)), (spp, pixel, dim)
                seen.add(st.tobytes())
        assert spp < 16 or len(seen) >= 27  # 28 (pixel, dimension) pairs: practically all orders differ
    # another seed: another set of orders
    assert not np.array_equal(op.strata(1, 5, 2, 64), op.strata(2, 5, 2, 64))


def _variance_ratio(render_batch, gold, batches=24):
    imgs = np.stack([util.film_image(render_batch(b)) for b in range(batches)])
    var = imgs.var(0, ddof=1)
    ref_var = gold["var_of_mean"].astype(np.float64) * gold_batches(gold)
    lit = gold["mean"].sum(2) > 0
    per_pixel = var[lit].sum(1) / np.maximum(ref_var[lit].sum(1), 1e-12)
    return var[lit].mean() / ref_var[lit].mean(), float(np.median(per_pixel)), imgs


def gold_batches(gold):
    return 24  # tests/golden/make_golden.py: 24 reference renders of 256 spp


def test_variance_per_sample_matches_the_reference_cpu(built):
    """The oracle port draws the product's numbers (same Philox, same strata): 24 images of 256 spp of the tiny scene
    against the 24 reference renders behind tiny_film_pt.npz.  Without the strata (GO_NO_STRATA, what round 1 drew)
    the same comparison gives 1.25 (mean) / 2.0 (median pixel): a sample was worth half a reference sample."""
    gold = util.golden("tiny_film_pt.npz")
    scene = api.Scene(util.TINY_PT)
    mean_ratio, median_ratio, imgs = _variance_ratio(lambda b: op.render(scene, seed=300 + b, spp_total=256)[0], gold)
    print(f"variance per 256-spp image, product sampler / reference sampler: mean {mean_ratio:.3f}, median pixel {median_ratio:.3f}")
    assert median_ratio <= 1.05 and mean_ratio <= 1.10
    t = util.film_ttest(imgs, gold)
    assert abs(t.mean()) < 0.1 and 0.85 < t.std() < 1.15  # and still unbiased


def _ao_variance_ratio(render_batch):
    gold = util.golden("tiny_film_ao.npz")
    nb, spp = int(gold["batches"]), int(gold["spp_per_batch"])
    imgs = np.stack([util.film_image(render_batch(b, spp)) for b in range(24)])
    var, ref_var = imgs.var(0, ddof=1), gold["var_of_mean"].astype(np.float64) * nb
    lit = gold["mean"].sum(2) > 0
    noisy = ref_var[lit].sum(1) > 1e-8  # unoccluded pixels are exactly 1 on both sides
    return var[lit].mean() / ref_var[lit].mean(), float(np.median(var[lit].sum(1)[noisy] / ref_var[lit].sum(1)[noisy]))


def test_ao_variance_per_sample_matches_the_reference_cpu(built):
    """The AO integrator's 25 direction strata, sub-stratified over the samples of a pixel (stratifiedUniform2D(buffer, n),
    src/GoblinSampler.cpp:276-307): one permutation per camera sample + a hashed cyclic shift per AO ray.  16 x 64 spp
    reference renders behind tiny_film_ao.npz; without the sub-strata (GO_NO_STRATA) the ratio is 2.1."""
    scene = api.Scene(util.TINY_AO)
    mean_ratio, median_ratio = _ao_variance_ratio(lambda b, spp: op.render(scene, seed=300 + b, spp_total=spp)[0])
    print(f"AO variance per image, product sampler / reference sampler: mean {mean_ratio:.3f}, median pixel {median_ratio:.3f}")
    assert median_ratio <= 1.08 and mean_ratio <= 1.08


@pytest.mark.gpu
def test_ao_variance_per_sample_matches_the_reference_gpu(built):
    scene = api.Scene(util.TINY_AO)
    ctx = api.Context(0)
    ctx.upload_scene(scene)

    def batch(b, spp):
        ctx.film_clear()
        ctx.render(seed=300 + b, spp_total=spp)
        return ctx.film_download().copy()

    mean_ratio, median_ratio = _ao_variance_ratio(batch)
    print(f"GPU AO variance per image / reference: mean {mean_ratio:.3f}, median pixel {median_ratio:.3f}")
    assert median_ratio <= 1.08 and mean_ratio <= 1.08
    ctx.close()


@pytest.mark.gpu
def test_variance_per_sample_matches_the_reference_gpu(built):
    gold = util.golden("tiny_film_pt.npz")
    scene = api.Scene(util.TINY_PT)
    ctx = api.Context(0)
    ctx.upload_scene(scene)

    def batch(b):
        ctx.film_clear()
        ctx.render(seed=300 + b, spp_total=256)
        return ctx.film_download().copy()

    mean_ratio, median_ratio, _ = _variance_ratio(batch, gold)
    print(f"GPU variance per 256-spp image / reference: mean {mean_ratio:.3f}, median pixel {median_ratio:.3f}")
    assert median_ratio <= 1.05 and mean_ratio <= 1.10
    ctx.close()


@pytest.mark.gpu
@pytest.mark.skipif(not util.have_ref_tool(), reason="oracle/_ref/ref_tool not built")
def test_variance_per_sample_bunny_against_live_reference(built, tmp_path):
    """S1 geometry at 128 x 96: 16 images of 256 spp from the reference (rendered now) and from the GPU."""
    from goblin_b200 import gbar
    path = os.path.join(util.gen_scene("bunny"), "bunny_pt_small.json")
    refs = []
    with util.RefSession(path, cwd=str(tmp_path)) as ref:
        for b in range(16):
            ref.render(str(tmp_path / "f.gbar"), 900 + b, 256)
            refs.append(util.film_image(gbar.load(str(tmp_path / "f.gbar"))["film"]))
    refs = np.stack(refs)
    scene = api.Scene(path)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    imgs = []
    for b in range(16):
        ctx.film_clear()
        ctx.render(seed=40 + b, spp_total=256)
        imgs.append(util.film_image(ctx.film_download()))
    ctx.close()
    imgs = np.stack(imgs)
    v_gpu, v_ref = imgs.var(0, ddof=1).sum(2), refs.var(0, ddof=1).sum(2)
    # a spot light on smooth surfaces: most pixels have (numerically) no variance; compare where the reference has some
    noisy = v_ref > np.quantile(v_ref, 0.5)
    mean_ratio = v_gpu[noisy].mean() / v_ref[noisy].mean()
    median_ratio = float(np.median(v_gpu[noisy] / v_ref[noisy]))
    print(f"S1 variance per 256-spp image, GPU / reference: mean {mean_ratio:.3f}, median pixel {median_ratio:.3f}")
    assert median_ratio <= 1.10 and mean_ratio <= 1.15  # 16 batches: the estimate itself carries ~5 %
