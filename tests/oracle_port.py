"""ctypes binding of oracle/_build/libgoblin_oracle.so (TEST INFRASTRUCTURE).

The oracle is the CPU restatement of the reference in oracle/goblin_oracle.cpp.
It is the checker for the CUDA path; the product never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from goblin_b200 import api

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "libgoblin_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "goblin_oracle.cpp")
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
        l = C.CDLL(LIB_PATH)
        P = C.c_void_p
        l.go_trace_closest.argtypes = [P, P, C.c_size_t, P, C.c_int, P]
        l.go_trace_any.argtypes = [P, P, C.c_size_t, P, C.c_int, P]
        l.go_trace_fragments.argtypes = [P, P, C.c_size_t, P]
        l.go_camera_rays.argtypes = [P, P, C.c_size_t, P]
        l.go_li.argtypes = [P, P, C.c_size_t, C.c_size_t, P, P, C.c_int, P]
        l.go_render.argtypes = [P, P, P, C.c_int, C.c_int, P, P]
        l.go_strata.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, P]
        _lib = l
    return _lib


def hardware_threads():
    return lib().go_hardware_threads()


def strata(seed, pixel, dim, spp):
    """Stratum index of every sample of a pixel in one dimension (the product's sampler, restated)."""
    out = np.zeros(spp, dtype=np.uint32)
    assert lib().go_strata(seed, pixel, dim, spp, out.ctypes.data) == 0
    return out


def _counters(c):
    return {k: getattr(c, k) for k, _ in api.Counters._fields_}


def trace_closest(scene, rays, threads=0, counters=False):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    hits = np.zeros(rays.shape[0], dtype=api.HIT_DTYPE)
    c = api.Counters()
    rc = lib().go_trace_closest(C.addressof(scene.desc), rays.ctypes.data, rays.shape[0], hits.ctypes.data, threads,
                                C.addressof(c))
    assert rc == 0
    return (hits, _counters(c)) if counters else hits


def trace_any(scene, rays, threads=0, counters=False):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    occ = np.zeros(rays.shape[0], dtype=np.uint8)
    c = api.Counters()
    rc = lib().go_trace_any(C.addressof(scene.desc), rays.ctypes.data, rays.shape[0], occ.ctypes.data, threads,
                            C.addressof(c))
    assert rc == 0
    return (occ, _counters(c)) if counters else occ


def trace_fragments(scene, rays):
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    out = np.zeros((rays.shape[0], 9), dtype=np.float32)
    assert lib().go_trace_fragments(C.addressof(scene.desc), rays.ctypes.data, rays.shape[0], out.ctypes.data) == 0
    return out


def camera_rays(scene, samples):
    samples = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
    rays = np.zeros((samples.shape[0], 8), dtype=np.float32)
    assert lib().go_camera_rays(C.addressof(scene.desc), samples.ctypes.data, samples.shape[0], rays.ctypes.data) == 0
    return rays


def li(scene, samples, threads=0, calls=False, counters=False):
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    n, row = samples.shape
    out = np.zeros((n, 3), dtype=np.float32)
    ref_calls = np.zeros((n, 2), dtype=np.uint32)
    c = api.Counters()
    rc = lib().go_li(C.addressof(scene.desc), samples.ctypes.data, n, row, out.ctypes.data, ref_calls.ctypes.data,
                     threads, C.addressof(c))
    assert rc == 0, "sample rows too short for this integrator"
    res = (out,)
    if calls:
        res += (ref_calls,)
    if counters:
        res += (_counters(c),)
    return res if len(res) > 1 else out


def render(scene, seed=1, spp_total=None, spp_begin=0, spp_end=None, max_ray_depth=0, method=-1, ao_sample_num=0,
           threads=0, pixel_stride=1, film=None):
    """Returns (film yres x xres x 4, counters dict, (ref intersect calls, ref occluded calls))."""
    if spp_total is None:
        spp_total = scene.spp_squared()
    if spp_end is None:
        spp_end = spp_total
    f = scene.desc.film
    if film is None:
        film = np.zeros((f.yres, f.xres, 4), dtype=np.float32)
    p = api.RenderParams(seed, spp_total, spp_begin, spp_end, max_ray_depth, method, ao_sample_num)
    c = api.Counters()
    ref_calls = (C.c_uint64 * 2)(0, 0)
    rc = lib().go_render(C.addressof(scene.desc), C.addressof(p), film.ctypes.data, threads, pixel_stride,
                         C.addressof(c), C.addressof(ref_calls))
    assert rc == 0, "bad render parameters"
    return film, _counters(c), (ref_calls[0], ref_calls[1])
