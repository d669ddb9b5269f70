"""Film::writeImage (src/GoblinFilm.cpp:164-192): colour / weight, written as half-float BGR
OpenEXR like the reference (src/GoblinImageIO.cpp:35-98), or PFM / PPM by extension."""
import struct

import numpy as np
import pytest

from goblin_b200 import api


def _film(h=7, w=11, seed=3):
    rng = np.random.default_rng(seed)
    rgbw = rng.uniform(0.0, 4.0, (h, w, 4)).astype(np.float32)
    rgbw[..., 3] = rng.uniform(0.5, 2.0, (h, w))
    return rgbw


def _read_exr_half(path):
    """Minimal reader: single-part, uncompressed scanlines, HALF channels."""
    b = open(path, "rb").read()
    assert struct.unpack_from("<I", b, 0)[0] == 20000630
    off = 8
    attrs = {}
    while b[off] != 0:
        e = b.index(b"\0", off)
        name = b[off:e].decode()
        off = e + 1
        e = b.index(b"\0", off)
        typ = b[off:e].decode()
        off = e + 1
        (size,) = struct.unpack_from("<i", b, off)
        off += 4
        attrs[name] = (typ, b[off:off + size])
        off += size
    off += 1
    assert attrs["compression"][1] == b"\0"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans, c = [], attrs["channels"][1]
    p = 0
    while c[p] != 0:
        e = c.index(b"\0", p)
        chans.append(c[p:e].decode())
        assert struct.unpack_from("<i", c, e + 1)[0] == 1  # HALF
        p = e + 1 + 16
    offsets = struct.unpack_from(f"<{h}Q", b, off)
    img = {}
    for y in range(h):
        o = offsets[y]
        yy, n = struct.unpack_from("<ii", b, o)
        assert yy == y0 + y and n == 2 * w * len(chans)
        row = np.frombuffer(b, np.float16, w * len(chans), o + 8).reshape(len(chans), w)
        for k, name in enumerate(chans):
            img.setdefault(name, np.zeros((h, w), np.float16))[y] = row[k]
    return img


def test_exr_half_bgr(built, tmp_path):
    rgbw = _film()
    path = str(tmp_path / "a.exr")
    api.write_image(path, rgbw)
    img = _read_exr_half(path)
    assert sorted(img) == ["B", "G", "R"]
    # Color::operator/(float) multiplies by the reciprocal (src/GoblinColor.h:76-79)
    want = (rgbw[..., :3] * (np.float32(1.0) / rgbw[..., 3:4])).astype(np.float16)  # round-to-nearest-even
    for k, name in enumerate("RGB"):
        assert np.array_equal(img[name].view(np.uint16), want[..., k].view(np.uint16)), name


def test_pfm_and_ppm(built, tmp_path):
    rgbw = _film()
    want = rgbw[..., :3] * (np.float32(1.0) / rgbw[..., 3:4])
    pfm = str(tmp_path / "a.pfm")
    api.write_image(pfm, rgbw)
    b = open(pfm, "rb").read()
    head = b.split(b"\n", 3)
    assert head[0] == b"PF" and head[1].split() == [b"11", b"7"] and float(head[2]) < 0
    data = np.frombuffer(head[3], "<f4").reshape(7, 11, 3)[::-1]  # PFM rows run bottom to top
    assert np.array_equal(data, want)
    ppm = str(tmp_path / "a.ppm")
    api.write_image(ppm, rgbw)
    b = open(ppm, "rb").read()
    # ASCII P3, gamma 2.2, like the reference's writeImagePPM (src/GoblinImageIO.cpp:100-128)
    tok = b.split()
    assert tok[0] == b"P3" and tok[1:4] == [b"11", b"7", b"255"] and len(tok) == 4 + 7 * 11 * 3


def test_zero_weight_pixels_and_bad_path(built, tmp_path):
    rgbw = _film()
    rgbw[0, 0] = 0  # colour / weight with weight 0: the reference divides anyway (NaN); the file still writes
    api.write_image(str(tmp_path / "z.pfm"), rgbw)
    with pytest.raises(api.GoblinError) as e:
        api.write_image(str(tmp_path / "no_such_dir" / "a.exr"), rgbw)
    assert e.value.code == 2


# ---- Film::writeImage post-processing: Goblin::bloom, Goblin::toneMapping (src/GoblinImageIO.cpp:169-237)
# against tests/golden/post.npz, produced by the unmodified reference (ref_tool post).

def _post_golden():
    import os
    from tests import util
    return dict(np.load(os.path.join(util.GOLDEN, "post.npz")))


def test_ppm_and_tone_mapping_match_the_reference_writer(built, tmp_path):
    g = _post_golden()
    p = str(tmp_path / "plain.ppm")
    api.write_rgb(p, g["rgb"], tone_mapping=False)
    assert open(p, "rb").read() == g["ppm_plain"].tobytes()
    p = str(tmp_path / "tone.ppm")
    api.write_rgb(p, g["rgb"], tone_mapping=True)
    assert open(p, "rb").read() == g["ppm_tone"].tobytes()
    # tone mapping is a .ppm-only step in Goblin::writeImage: the float formats ignore the flag
    p = str(tmp_path / "tone.pfm")
    api.write_rgb(p, g["rgb"], tone_mapping=True)
    raw = np.frombuffer(open(p, "rb").read()[-g["rgb"].size * 4:], np.float32).reshape(g["rgb"].shape)[::-1]
    assert np.array_equal(raw, g["rgb"])


def _post_scene(radius, weight, tone):
    import json
    import os
    from tests import util
    g = _post_golden()
    h, w, _ = g["rgb"].shape
    js = json.load(open(util.TINY_PT))
    js["camera"]["film"] = {"resolution": [float(w), float(h)], "tone_mapping": bool(tone),
                            "bloom_radius": radius, "bloom_weight": weight}
    return api.Scene(json_text=json.dumps(js), scene_dir=os.path.dirname(util.TINY_PT)), g


def test_loader_reads_the_film_post_parameters(built):
    scene, _ = _post_scene(0.2, 0.3, True)
    f = scene.desc.film
    assert (f.tone_mapping, f.bloom_radius, f.bloom_weight) == (1, np.float32(0.2), np.float32(0.3))
    f = api.Scene(__import__("tests.util", fromlist=["x"]).TINY_PT).desc.film
    assert (f.tone_mapping, f.bloom_radius, f.bloom_weight) == (0, 0.0, 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("radius,weight,key", [(0.2, 0.3, "bloom"), (2.5, 1.0, "bloom_wide"), (0.0, 0.0, "rgb")])
def test_device_bloom_is_bit_exact(built, radius, weight, key):
    scene, g = _post_scene(radius, weight, False)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    rgbw = np.concatenate([g["rgb"], np.ones(g["rgb"].shape[:2] + (1,), np.float32)], 2)
    ctx.film_upload(rgbw)
    got = ctx.film_resolve()
    assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32))


@pytest.mark.gpu
def test_film_write_bloom_then_tone_mapped_ppm(built, tmp_path):
    scene, g = _post_scene(0.2, 0.3, True)
    ctx = api.Context(0)
    ctx.upload_scene(scene)
    rgbw = np.concatenate([g["rgb"], np.ones(g["rgb"].shape[:2] + (1,), np.float32)], 2)
    ctx.film_upload(rgbw * np.float32(4.0))  # colour / weight: the weight 4 divides back out exactly
    p = str(tmp_path / "out.ppm")
    ctx.film_write(p)
    assert open(p, "rb").read() == g["ppm_bloom_tone"].tobytes()
