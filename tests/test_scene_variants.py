"""Scene-format coverage: variants of the tiny scene (thin-lens camera, every reconstruction
filter, a crop window, spp that is not a square, no lights) through the whole stack.

CPU: the host loader against the reference's own structures (needs oracle/_ref/ref_tool, built
in this container and shipped to the GPU box) and the oracle against the reference's Li.
GPU: the CUDA path against the oracle on the same variants."""
import json
import os
import subprocess
import tempfile
import zlib

import numpy as np
import pytest

from goblin_b200 import api, gbar
from tests import oracle_port as op
from tests import util


def _variant(name):
    sc = json.load(open(util.TINY_PT))
    if name == "dof":
        sc["camera"]["lens_radius"] = 0.08
        sc["camera"]["focal_distance"] = 6.5
    elif name == "box":
        sc["camera"]["filter"] = {"type": "box", "width": [0.5, 0.5]}
    elif name == "triangle":
        sc["camera"]["filter"] = {"type": "triangle", "width": [1.5, 1.25]}
    elif name == "mitchell":
        sc["camera"]["filter"] = {"type": "mitchell", "width": [2.0, 2.0], "b": 0.3333333, "c": 0.3333333}
    elif name == "wide_gaussian":
        sc["camera"]["filter"] = {"type": "gaussian", "width": [3.0, 3.0], "falloff": 1.0}
    elif name == "crop":
        sc["camera"]["film"]["crop"] = [0.25, 0.75, 0.3, 0.9]
    elif name == "spp50":
        sc["render_setting"]["sample_per_pixel"] = 50
    elif name == "nolights":
        sc["lights"] = []
    elif name == "depth1":
        sc["render_setting"]["max_ray_depth"] = 1
    elif name == "depth2":
        sc["render_setting"]["max_ray_depth"] = 2
    elif name == "ao10":  # not a square: the reference shoots 16 rays (SampleQuota::requestTwoDQuota)
        sc["render_setting"]["render_method"] = "ao"
        sc["render_setting"]["ao_sample_num"] = 10
    elif name == "delta_only":
        sc["lights"] = [l for l in sc["lights"] if l["type"] != "area"]
    elif name.startswith("ibl"):
        # image based light (src/GoblinLight.cpp:464-629): a 40 x 24 map (resized to 64 x 32 by the
        # MIPMap) beside a point light, or a 32 x 16 one alone
        sky = {"name": "sky", "type": "ibl", "file": "_env_40x24.exr", "filter": [0.8, 0.9, 1.0], "euler": [10.0, 40.0, 0.0]}
        if name == "ibl":
            sc["lights"] = [l for l in sc["lights"] if l["type"] == "point"] + [sky]
        elif name == "ibl_only":
            sky["file"] = "_env_32x16.exr"
            sky["filter"] = [1.0, 1.0, 1.0]
            del sky["euler"]
            sc["lights"] = [sky]
        else:  # every light kind at once, the map missing: the reference falls back to 1 x 1 magenta
            sky["file"] = "_env_missing.exr"
            sc["lights"] = sc["lights"] + [sky]
    elif name.startswith("bump"):
        # bump and normal maps (BumpShaders::evaluate, src/GoblinMaterial.cpp:221-281) on meshes with and without
        # vt / vn, a sphere and a disk; a constant bump map still replaces the normal by the dpdu x dpdv frame's
        tex = sc["textures"]
        tex += [
            {"format": "float", "name": "lo", "type": "constant", "float": 0.0},
            {"format": "float", "name": "hi", "type": "constant", "float": 0.004},
            {"format": "float", "name": "dents", "type": "checkerboard", "texture1": "lo", "texture2": "hi",
             "mapping": "uv", "scale": [24.0, 24.0]},
            {"format": "float", "name": "relief", "type": "image", "file": "_env_32x16.exr", "filter": "nearest",
             "channel": "G", "mapping": "uv", "scale": [2.0, 2.0]},
            {"format": "float", "name": "milli", "type": "constant", "float": 0.001},
            {"format": "float", "name": "relief_small", "type": "scale", "texture": "relief", "scale": "milli"},
            {"format": "color", "name": "nmap", "type": "image", "file": "_env_40x24.exr", "filter": "nearest",
             "mapping": "uv", "scale": [3.0, 3.0]},
            {"format": "color", "name": "bluish", "type": "constant", "color": [0.55, 0.45, 0.95]},
        ]
        by = {m["name"]: m for m in sc["materials"]}
        by["grey"]["bumpmap"] = "dents"           # floor mesh (vt, vn)
        by["green"]["bumpmap"] = "lo"             # box: flat bump map
        by["mirror"]["normalmap"] = "bluish"      # sphere, constant tangent-space normal
        by["gloss"]["bumpmap"] = "relief_small"   # disk
        by["gloss"]["normalmap"] = "nmap"
        by["metal"]["normalmap"] = "nmap"         # flat-shaded mesh without vt
        by["glass"]["bumpmap"] = "relief_small"   # blob meshes + unit sphere
        if name == "bump_ao":
            sc["render_setting"]["render_method"] = "ao"
            sc["render_setting"]["ao_sample_num"] = 9
    elif name == "ortho":  # OrthographicCamera (src/GoblinCamera.cpp:290-329) with a filtered checkerboard floor
        sc["camera"]["type"] = "orthographic"
        sc["camera"]["film_width"] = 9.0
        sc["textures"] += [{"format": "color", "name": "check", "type": "checkerboard", "texture1": "red", "texture2": "white",
                            "mapping": "uv", "scale": [9.0, 9.0], "filter": True}]
        for m in sc["materials"]:
            if m["name"] == "grey":
                m["Kd"] = "check"
    elif name.startswith("mask"):
        # Mask materials (src/GoblinMaterial.cpp:747-811): constant and checkerboard alpha, tinted
        # pass-through, around Lambert / Blinn / glass; they bring back the path tracer's isOpaque filter
        # and evalAttenuation (src/GoblinPathtracer.cpp:5-48) for real
        tex = sc["textures"]
        tex += [
            {"format": "float", "name": "half", "type": "constant", "float": 0.5},
            {"format": "float", "name": "zero", "type": "constant", "float": 0.0},
            {"format": "float", "name": "one", "type": "constant", "float": 1.0},
            {"format": "float", "name": "cutout", "type": "checkerboard", "texture1": "zero", "texture2": "one",
             "mapping": "uv", "scale": [6.0, 6.0]},
            {"format": "color", "name": "amber", "type": "constant", "color": [1.0, 0.8, 0.5]},
        ]
        sc["materials"] += [
            {"name": "veil", "type": "mask", "material": "green", "alpha": "half", "transparent_color": "amber"},
            {"name": "lace", "type": "mask", "material": "gloss", "alpha": "cutout"},
            {"name": "ghost_glass", "type": "mask", "material": "glass", "alpha": "half"},
            {"name": "plain", "type": "mask", "material": "red"},
        ]
        for pr in sc["primitives"]:
            if pr["type"] != "model":
                continue
            pr["material"] = {"box": "veil", "plate": "lace", "blob": "ghost_glass", "ico": "plain"}.get(pr["name"], pr["material"])
        if name == "mask_ibl":
            sc["lights"] = [l for l in sc["lights"] if l["type"] in ("point", "area")][:2] + [
                {"name": "sky", "type": "ibl", "file": "_env_32x16.exr", "filter": [0.6, 0.7, 0.9]}]
    elif name.startswith("img_"):
        # image textures (src/GoblinTexture.cpp:82-288,431-503): every filter and address mode, both
        # mappings, colour and float formats, gamma and channel selection
        tex = sc["textures"]
        filt = {"img_nearest": "nearest", "img_bilinear": "bilinear", "img_trilinear": "trilinear", "img_ewa": "EWA"}[name]
        tex += [
            {"format": "color", "name": "photo", "type": "image", "file": "_env_40x24.exr", "filter": filt,
             "mapping": "uv", "scale": [6.0, 4.0], "offset": [0.3, 0.2]},
            {"format": "color", "name": "photo_clamp", "type": "image", "file": "_env_32x16.exr", "filter": filt,
             "address": "clamp", "gamma": 2.2, "channel": "G", "mapping": "uv", "scale": [1.5, 1.5]},
            {"format": "color", "name": "photo_border", "type": "image", "file": "_env_32x16.exr", "filter": filt,
             "address": "border", "mapping": "spherical", "position": [0.0, 0.85, 1.6]},
            {"format": "float", "name": "gloss_map", "type": "image", "file": "_env_32x16.exr", "filter": filt,
             "channel": "R", "gamma": 0.5, "max_anisotropy": 4.0, "mapping": "uv", "scale": [3.0, 3.0]},
            {"format": "float", "name": "forty", "type": "constant", "float": 40.0},
            {"format": "float", "name": "exp_map", "type": "scale", "texture": "gloss_map", "scale": "forty"},
            {"format": "color", "name": "no_file", "type": "image", "file": "_nope.exr"},
        ]
        by = {m["name"]: m for m in sc["materials"]}
        by["grey"]["Kd"] = "photo"
        by["green"]["Kd"] = "photo_clamp"
        by["mirror"]["Kr"] = "photo_border"
        by["glass"]["Kt"] = "photo"
        by["glass"]["Kr"] = "no_file"
        by["gloss"]["exponent"] = "exp_map"
        by["metal"]["exponent"] = "exp_map"
        if name == "img_bilinear":
            # the reference's bilinear filter indexes one level past the pyramid once the footprint
            # exceeds ~0.7 of the coarsest level it rounds to (MIPMap::lookup clamps to mLevelsNum, not
            # mLevelsNum - 1) and crashes: keep this variant away from the 1 x 1 fallback image and
            # from the spherical mapping's poles
            # and from uv singularities (sphere poles, disk centre, per-face default uvs)
            by["glass"]["Kr"] = "tint"
            by["glass"]["Kt"] = "tint"
            by["mirror"]["Kr"] = "white"
            by["gloss"]["exponent"] = "shiny"
            by["metal"]["exponent"] = "satin"
    elif name.startswith("tex_"):
        # procedural textures (src/GoblinTexture.cpp:292-427) on every material slot that takes one
        tex = sc["textures"]
        filt = name != "tex_point"
        tex += [
            {"format": "color", "name": "check", "type": "checkerboard", "texture1": "red", "texture2": "white",
             "mapping": "uv", "scale": [7.0, 5.0], "offset": [0.25, 0.1], "filter": filt},
            {"format": "float", "name": "half", "type": "constant", "float": 0.5},
            {"format": "float", "name": "one", "type": "constant", "float": 1.0},
            {"format": "float", "name": "fcheck", "type": "checkerboard", "texture1": "half", "texture2": "one",
             "mapping": "uv", "scale": [3.0, 3.0], "filter": filt},
            {"format": "color", "name": "dimmed", "type": "scale", "texture": "check", "scale": "fcheck"},
            {"format": "color", "name": "globe", "type": "checkerboard", "texture1": "green", "texture2": "dimmed",
             "mapping": "spherical", "position": [0.0, 0.85, 1.6], "scale": [1.0, 1.0, 1.0], "filter": filt},
            {"format": "float", "name": "rough", "type": "checkerboard", "texture1": "satin", "texture2": "shiny",
             "mapping": "uv", "scale": [4.0, 4.0], "filter": filt},
            {"format": "color", "name": "undefined_child", "type": "checkerboard", "texture1": "nope", "texture2": "grey",
             "scale": [2.0, 2.0]},
        ]
        by = {m["name"]: m for m in sc["materials"]}
        by["grey"]["Kd"] = "check"          # floor: mesh with vt
        by["green"]["Kd"] = "dimmed"        # box
        by["mirror"]["Kr"] = "globe"        # sphere, spherical mapping in world space
        by["glass"]["Kt"] = "check"         # blob meshes + unit sphere: sphere uv
        by["glass"]["Kr"] = "undefined_child"
        by["gloss"]["Kg"] = "globe"         # disk
        by["gloss"]["exponent"] = "rough"
        by["metal"]["exponent"] = "rough"
        if name == "tex_dof":
            sc["camera"]["lens_radius"] = 0.08
            sc["camera"]["focal_distance"] = 6.5
    else:
        raise KeyError(name)
    return sc


VARIANTS = ["dof", "box", "triangle", "mitchell", "wide_gaussian", "crop", "spp50", "nolights", "delta_only", "depth1",
            "depth2", "ao10", "tex_point", "tex_filtered", "tex_dof", "ibl", "ibl_only", "ibl_missing",
            "img_nearest", "img_bilinear", "img_trilinear", "img_ewa", "mask", "mask_ibl", "ortho", "bump", "bump_ao"]


def write_env_maps(d):
    """The two HDR maps the image-textured / image-lit variants read: a small very bright region on noise,
    written by the product's EXR writer (deterministic: the golden vectors depend on them)."""
    rng = np.random.default_rng(5)
    env = []
    for w, h in ((40, 24), (32, 16)):
        img = rng.uniform(0.05, 1.0, (h, w, 3)).astype(np.float32)
        img[h // 6:h // 6 + 3, w // 4:w // 4 + 4] *= 40.0
        env.append(os.path.join(d, f"_env_{w}x{h}.exr"))
        api.write_rgb(env[-1], img)
    return env


@pytest.fixture(scope="module")
def variant_files(tmp_path_factory, built):
    """JSON files next to the tiny scene's models (mesh paths are relative to the scene file)."""
    out = {}
    d = os.path.dirname(util.TINY_PT)
    env = write_env_maps(d)
    for v in VARIANTS:
        path = os.path.join(d, f"_variant_{v}.json")
        with open(path, "w") as f:
            json.dump(_variant(v), f)
        out[v] = path
    yield out
    for p in env:
        os.remove(p)
    for p in out.values():
        for q in (p, p[:-5] + ".exr"):
            if os.path.exists(q):
                os.remove(q)


def _row_floats(scene):
    st = scene.desc.setting
    return 4 + (2 * st.ao_sample_num if st.method == 1 else 7 * st.max_ray_depth)


def _ref(cmd, scene, *rest):
    subprocess.run([util.REF_TOOL, cmd, scene, *rest], check=True, capture_output=True)


@pytest.mark.skipif(not util.have_ref_tool(), reason="oracle/_ref/ref_tool not built")
@pytest.mark.parametrize("v", VARIANTS)
def test_loader_and_oracle_match_reference(variant_files, v):
    scene = api.Scene(variant_files[v])
    with tempfile.TemporaryDirectory() as td:
        _ref("dump", variant_files[v], td + "/d.gbar")
        dump = {k: a.copy() for k, a in gbar.load(td + "/d.gbar").items()}
        f = scene.desc.film
        assert [f.xres, f.yres, f.xstart, f.xcount, f.ystart, f.ycount, f.sx0, f.sx1, f.sy0, f.sy1] == list(dump["film"])
        assert np.array_equal(np.array(f.filter_table[:], np.float32).view(np.uint32), dump["filter.table"].view(np.uint32))
        assert np.array_equal(np.array(f.filter_width[:], np.float32), dump["filter.width"])
        cam = scene.desc.camera
        mine = np.array(list(cam.position) + list(cam.orientation) + [cam.proj00, cam.proj11, cam.lens_radius,
                                                                      cam.focal_distance], np.float32)
        assert np.array_equal(mine.view(np.uint32), dump["camera"].view(np.uint32))
        nodes = np.frombuffer(np.ascontiguousarray(scene.top_nodes()).tobytes(), np.uint8).reshape(-1, 32)
        assert np.array_equal(nodes, dump["top.nodes"])
        assert np.array_equal(scene.top_order(), dump["top.order"])
        if scene.desc.n_lights:
            assert np.array_equal(scene.light_cdf().view(np.uint32), dump["light.cdf"].view(np.uint32))
        for li in range(scene.desc.n_lights):  # image based lights: map, orientation, sampling tables
            L = scene.desc.lights[li]
            if L.type != 4:
                continue
            pre = f"ibl{li}."
            tex = np.ctypeslib.as_array(scene.desc.image_texels, (scene.desc.n_image_texels * 4,))
            tex = tex[4 * L.image_offset:4 * (L.image_offset + L.image_width * L.image_height)]
            assert [L.image_width, L.image_height] == list(dump[pre + "sizes"][:2])
            assert np.array_equal(tex.view(np.uint32), dump[pre + "level0"].view(np.uint32))
            n_dist = len(dump[pre + "dist"])
            dist = np.ctypeslib.as_array(scene.desc.light_dist, (scene.desc.n_light_dist,))[L.dist_offset:L.dist_offset + n_dist]
            assert [L.dist_width, L.dist_height] == list(dump[pre + "dist_size"])
            assert np.array_equal(dist.view(np.uint32), dump[pre + "dist"].view(np.uint32))
            assert np.array_equal(np.array(L.to_world[:], np.float32).view(np.uint32), dump[pre + "to_world"].view(np.uint32))
            assert np.array_equal(np.array(L.to_object[:], np.float32).view(np.uint32), dump[pre + "to_object"].view(np.uint32))
        # Li and camera rays on fresh samples
        rng = np.random.default_rng(zlib.crc32(v.encode()))
        rows = rng.uniform(0, 1, (1500, _row_floats(scene))).astype(np.float32)
        rows[:, 0] = rng.uniform(f.sx0, f.sx1, 1500)
        rows[:, 1] = rng.uniform(f.sy0, f.sy1, 1500)
        rows.tofile(td + "/rows.f32")
        rows[:, :4].copy().tofile(td + "/cam.f32")
        _ref("li", variant_files[v], td + "/rows.f32", td + "/l.gbar")
        _ref("camrays", variant_files[v], td + "/cam.f32", td + "/c.gbar")
        ref_l = {k: a.copy() for k, a in gbar.load(td + "/l.gbar").items()}
        ref_c = gbar.load(td + "/c.gbar")["rays"].copy()
    got_c = op.camera_rays(scene, rows[:, :4])
    assert np.array_equal(got_c[:, [0, 1, 2, 6, 7]], ref_c[:, [0, 1, 2, 6, 7]]) or "dof" in v
    assert np.allclose(got_c[:, :6], ref_c[:, :6], rtol=0, atol=1e-6)  # lens sampling: sin / cos of libm, 1 ulp at |o| ~ 6.5
    L, calls = op.li(scene, rows, calls=True)
    same = (calls == ref_l["calls"]).all(axis=1)
    assert same.mean() > 0.995  # a 1-ulp lens difference may flip a grazing path
    close = np.isclose(L[same], ref_l["L"][same], rtol=1e-4, atol=1e-5).all(axis=1)
    # image-textured Blinn exponents reach ~1600 on the maps' bright patch: there a last-bit difference
    # in a direction is amplified past the tolerance on a few samples in ten thousand
    # normal-mapped glossy meshes likewise: sampled directions dip below the geometric surface, the path
    # rattles between faces at t ~ 1e-2 and a last-bit difference decides which face comes next
    assert close.all() or (v.startswith("img_") and close.mean() > 0.998) or (v.startswith("bump") and close.mean() > 0.995)


@pytest.mark.gpu
@pytest.mark.parametrize("v", VARIANTS)
def test_gpu_matches_oracle(variant_files, v):
    scene = api.Scene(variant_files[v])
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    f = scene.desc.film
    rng = np.random.default_rng(5)
    rows = rng.uniform(0, 1, (4000, _row_floats(scene))).astype(np.float32)
    rows[:, 0] = rng.uniform(f.sx0, f.sx1, 4000)
    rows[:, 1] = rng.uniform(f.sy0, f.sy1, 4000)
    cam_g, cam_o = ctx.camera_rays(rows[:, :4]), op.camera_rays(scene, rows[:, :4])
    assert np.allclose(cam_g[:, :6], cam_o[:, :6], rtol=0, atol=2e-6)  # thin lens: CUDA sincosf vs libm
    assert np.array_equal(cam_g[:, 6:], cam_o[:, 6:])
    hits_g, hits_o = ctx.trace_closest(cam_o), op.trace_closest(scene, cam_o)
    assert np.array_equal(hits_g["inst"], hits_o["inst"]) and np.array_equal(hits_g["prim"], hits_o["prim"])
    assert np.array_equal(hits_g["t"].view(np.uint32), hits_o["t"].view(np.uint32))
    got, want = ctx.li(rows), op.li(scene, rows)
    close = np.isclose(got, want, rtol=2e-3, atol=2e-4).all(axis=1)
    assert close.mean() >= 0.995
    spp = scene.spp_squared()
    ctx.film_clear()
    ctx.render(seed=12, spp_total=spp)
    g = ctx.film_download()
    c, cnt, _ = op.render(scene, seed=12, spp_total=spp)
    assert ctx.counters()["camera_samples"] >= cnt["camera_samples"]
    assert np.allclose(g[..., 3], c[..., 3], rtol=1e-4, atol=1e-5)
    ok = np.isclose(g[..., :3], c[..., :3], rtol=2e-3, atol=1e-4).all(axis=2)
    # bump: the rattling normal-mapped paths (see above) differ on ~0.1 % of the samples, and a pixel gathers hundreds
    assert ok.mean() > (0.95 if v.startswith("bump") else 0.995), v
    if v == "crop":  # nothing outside the crop window is touched
        mask = np.zeros(g.shape[:2], bool)
        mask[f.ystart:f.ystart + f.ycount, f.xstart:f.xstart + f.xcount] = True
        assert (g[~mask] == 0).all() and (g[mask][:, 3] > 0).all()
    if v == "nolights":
        assert (g[..., :3] == 0).all() and (g[..., 3] > 0).any()
    # the one-warp film kernel (candidate windows of at most 32 pixels: every variant but wide_gaussian) against the
    # general one, gb_set_tuning values[7]: same weights, same order of additions per sample pixel
    g = g.copy()
    ctx.set_tuning([20, 6, 4, 10, 0, 2, 1, 1])
    ctx.film_clear()
    ctx.render(seed=12, spp_total=spp)
    general = ctx.film_download()
    assert np.allclose(general, g, rtol=2e-5, atol=1e-6), v
    assert np.array_equal(general == 0, g == 0), v
    ctx.close()


# ---- the OpenEXR reader behind image based lights against the reference's (tinyexr's LoadEXR)

def _write_exr(path, img, compression, pixel_type, line_order_decreasing=False, channels="BGR"):
    """A small OpenEXR scanline writer for tests: compression 0 none, 1 RLE, 2 ZIPS, 3 ZIP;
    pixel_type 1 HALF, 2 FLOAT."""
    import struct
    h, w, _ = img.shape
    idx = {"R": 0, "G": 1, "B": 2}

    def attr(name, typ, val):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(val)) + val

    names = sorted(channels)
    ch = b"".join(n.encode() + b"\0" + struct.pack("<iBxxxii", pixel_type, 0, 1, 1) for n in names) + b"\0"
    win = struct.pack("<4i", 0, 0, w - 1, h - 1)
    hd = struct.pack("<II", 20000630, 2)
    hd += attr("channels", "chlist", ch) + attr("compression", "compression", bytes([compression]))
    hd += attr("dataWindow", "box2i", win) + attr("displayWindow", "box2i", win)
    hd += attr("lineOrder", "lineOrder", bytes([1 if line_order_decreasing else 0]))
    hd += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    hd += attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) + attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    hd += b"\0"
    lines = 16 if compression == 3 else 1
    dt = np.float16 if pixel_type == 1 else np.float32

    def shuffle(raw):
        b = np.frombuffer(raw, np.uint8)
        t = np.concatenate([b[0::2], b[1::2]]).astype(np.int32)
        d = t.copy()
        d[1:] = (t[1:] - t[:-1] + 128 + 256) % 256
        return d.astype(np.uint8).tobytes()

    def rle(data):
        out = bytearray()
        i = 0
        while i < len(data):
            j = i
            while j + 1 < len(data) and data[j + 1] == data[i] and j - i < 126:
                j += 1
            if j - i >= 2:
                out += struct.pack("b", j - i) + data[i:i + 1]
                i = j + 1
            else:
                k = i
                while k < len(data) and k - i < 127 and not (k + 2 < len(data) and data[k] == data[k + 1] == data[k + 2]):
                    k += 1
                out += struct.pack("b", -(k - i)) + data[i:k]
                i = k
        return bytes(out)

    blocks = []
    for y0 in range(0, h, lines):
        raw = b"".join(img[y, :, idx[n]].astype(dt).tobytes() for y in range(y0, min(h, y0 + lines)) for n in names)
        if compression in (2, 3):
            z = zlib.compress(shuffle(raw))
            data = z if len(z) < len(raw) else raw
        elif compression == 1:
            z = rle(shuffle(raw))
            data = z if len(z) < len(raw) else raw
        else:
            data = raw
        blocks.append((y0, data))
    if line_order_decreasing:
        blocks = blocks[::-1]
    off = len(hd) + 8 * len(blocks)
    table, body = {}, b""
    for y0, data in blocks:
        table[y0] = off + len(body)
        body += struct.pack("<ii", y0, len(data)) + data
    tab = b"".join(struct.pack("<Q", table[y0]) for y0 in sorted(table))
    open(path, "wb").write(hd + tab + body)


@pytest.mark.skipif(not util.have_ref_tool(), reason="oracle/_ref/ref_tool not built")
@pytest.mark.parametrize("compression,pixel_type,decreasing", [(0, 2, False), (1, 1, False), (2, 2, False), (3, 1, False),
                                                                (3, 2, False)])  # increasing-y files only: the pinned case
def test_exr_reader_matches_the_reference_loader(built, tmp_path, compression, pixel_type, decreasing):
    rng = np.random.default_rng(compression * 7 + pixel_type)
    img = rng.uniform(0.0, 3.0, (40, 64, 3)).astype(np.float32)
    img[10:30, 5:40] = 0.25  # runs for RLE / ZIP to chew on
    sc = json.load(open(util.TINY_PT))
    sc["lights"] = [{"name": "sky", "type": "ibl", "file": "env.exr", "filter": [1.0, 1.0, 1.0]}]
    d = tmp_path / "scene"
    os.makedirs(d / "models")
    for f in os.listdir(os.path.join(os.path.dirname(util.TINY_PT), "models")):
        os.symlink(os.path.join(os.path.dirname(util.TINY_PT), "models", f), d / "models" / f)
    json.dump(sc, open(d / "s.json", "w"))
    _write_exr(str(d / "env.exr"), img, compression, pixel_type, decreasing)
    _ref("dump", str(d / "s.json"), str(d / "d.gbar"))
    dump = gbar.load(str(d / "d.gbar"))
    scene = api.Scene(str(d / "s.json"))
    L = scene.desc.lights[0]
    assert (L.image_width, L.image_height) == (64, 64)  # 64 x 40 resized up to powers of two
    tex = np.ctypeslib.as_array(scene.desc.image_texels, (scene.desc.n_image_texels * 4,))
    assert np.array_equal(tex.view(np.uint32), dump["ibl0.level0"].view(np.uint32))
    assert dump["ibl0.level0"].reshape(-1, 4)[:, :3].max() > 1.0  # really the image, not the magenta fallback


# ---- committed golden vectors from the unmodified reference (tests/golden/make_variant_golden.py): they travel
# to the GPU box, where /root/reference does not exist

GOLDEN_VARIANTS = ["tex_filtered", "ibl", "img_ewa", "mask", "mask_ibl", "bump", "ortho"]


def _variant_golden():
    return dict(np.load(os.path.join(util.GOLDEN, "variants_li.npz")))


@pytest.mark.parametrize("v", GOLDEN_VARIANTS)
def test_oracle_matches_reference_goldens(variant_files, v):
    g = _variant_golden()
    scene = api.Scene(variant_files[v])
    L, calls = op.li(scene, g[v + ".rows"], calls=True)
    same = (calls == g[v + ".calls"]).all(axis=1)
    assert same.mean() > 0.995
    close = np.isclose(L[same], g[v + ".L"][same], rtol=1e-4, atol=1e-5).all(axis=1)
    assert close.mean() > 0.995


@pytest.mark.gpu
@pytest.mark.parametrize("v", GOLDEN_VARIANTS)
def test_gpu_matches_reference_goldens(variant_files, v):
    """The CUDA path against radiance values computed by the reference itself."""
    g = _variant_golden()
    scene = api.Scene(variant_files[v])
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    got = ctx.li(g[v + ".rows"])
    close = np.isclose(got, g[v + ".L"], rtol=2e-3, atol=2e-4).all(axis=1)
    assert close.mean() >= 0.99, v
    assert abs(got.sum() - g[v + ".L"].sum()) <= 2e-2 * g[v + ".L"].sum()
    ctx.close()
