"""ctypes binding of include/goblin_b200.h (the C ABI of libgoblin_b200.so).

Thin by design: every function here forwards to one extern "C" entry point and
raises GoblinError on a non-zero status.  There is no Python or CPU fallback;
if the shared library is missing the import fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GOBLIN_B200_LIB: another build of the same library (tools/build_variants.sh, compile-time knob sweeps)
LIB_PATH = os.environ.get("GOBLIN_B200_LIB") or os.path.join(_HERE, "libgoblin_b200.so")


class GoblinError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"goblin_b200 error {code}: {msg}")
        self.code = code


class BvhNode(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("offset", C.c_uint32),
                ("nprims", C.c_uint8), ("axis", C.c_uint8), ("pad", C.c_uint8 * 2)]


class Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("radius", C.c_float), ("material", C.c_int32),
                ("area_light", C.c_int32), ("node_offset", C.c_uint32), ("node_count", C.c_uint32),
                ("tri_offset", C.c_uint32), ("tri_count", C.c_uint32), ("vert_offset", C.c_uint32),
                ("vert_count", C.c_uint32), ("has_normal", C.c_int32), ("has_uv", C.c_int32),
                ("is_camera_lens", C.c_int32), ("bound", C.c_float * 6)]


class Instance(C.Structure):
    _fields_ = [("to_world", C.c_float * 12), ("to_object", C.c_float * 12), ("aabb", C.c_float * 6),
                ("model", C.c_int32), ("pad", C.c_int32)]


class Material(C.Structure):
    _fields_ = [("type", C.c_int32), ("kd", C.c_float * 3), ("kt", C.c_float * 3), ("eta", C.c_float),
                ("k", C.c_float), ("exponent", C.c_float), ("fresnel", C.c_int32),
                ("kd_tex", C.c_int32), ("kt_tex", C.c_int32), ("exponent_tex", C.c_int32),
                ("mask", C.c_int32), ("alpha", C.c_float), ("transparent_color", C.c_float * 3),
                ("alpha_tex", C.c_int32), ("transparent_tex", C.c_int32), ("bump_tex", C.c_int32),
                ("normal_tex", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("type", C.c_int32), ("is_float", C.c_int32), ("value", C.c_float * 3),
                ("child", C.c_int32 * 2), ("filter", C.c_int32), ("mapping", C.c_int32),
                ("map_scale", C.c_float * 2), ("map_offset", C.c_float * 2), ("to_tex", C.c_float * 12),
                ("image_filter", C.c_int32), ("address_mode", C.c_int32), ("max_anisotropy", C.c_float),
                ("first_level", C.c_int32), ("n_levels", C.c_int32)]


class ImageLevel(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("texel_offset", C.c_uint64)]


class Light(C.Structure):
    _fields_ = [("type", C.c_int32), ("color", C.c_float * 3), ("position", C.c_float * 3),
                ("direction", C.c_float * 3), ("cos_theta_max", C.c_float),
                ("cos_falloff_start", C.c_float), ("geom_kind", C.c_int32), ("radius", C.c_float),
                ("area", C.c_float), ("to_world", C.c_float * 12), ("to_object", C.c_float * 12),
                ("instance", C.c_int32), ("model", C.c_int32), ("area_offset", C.c_uint32),
                ("cdf_offset", C.c_uint32), ("image_width", C.c_int32), ("image_height", C.c_int32),
                ("image_offset", C.c_uint64), ("dist_width", C.c_int32), ("dist_height", C.c_int32),
                ("dist_offset", C.c_uint64)]


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("orientation", C.c_float * 4), ("proj00", C.c_float),
                ("proj11", C.c_float), ("lens_radius", C.c_float), ("focal_distance", C.c_float),
                ("orthographic", C.c_int32), ("film_width", C.c_float), ("film_height", C.c_float)]


class FilmDesc(C.Structure):
    _fields_ = [("xres", C.c_int32), ("yres", C.c_int32), ("xstart", C.c_int32), ("xcount", C.c_int32),
                ("ystart", C.c_int32), ("ycount", C.c_int32), ("sx0", C.c_int32), ("sx1", C.c_int32),
                ("sy0", C.c_int32), ("sy1", C.c_int32), ("filter_width", C.c_float * 2),
                ("filter_table", C.c_float * 256), ("tone_mapping", C.c_int32), ("bloom_radius", C.c_float),
                ("bloom_weight", C.c_float)]


class RenderSetting(C.Structure):
    _fields_ = [("method", C.c_int32), ("spp", C.c_int32), ("max_ray_depth", C.c_int32),
                ("ao_sample_num", C.c_int32), ("gpu_num", C.c_int32), ("seed", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("top_nodes", C.POINTER(BvhNode)), ("n_top_nodes", C.c_uint32),
                ("top_order", C.POINTER(C.c_uint32)), ("instances", C.POINTER(Instance)),
                ("n_instances", C.c_uint32), ("models", C.POINTER(Model)), ("n_models", C.c_uint32),
                ("model_nodes", C.POINTER(BvhNode)), ("n_model_nodes", C.c_uint64),
                ("model_order", C.POINTER(C.c_uint32)), ("tri_index", C.POINTER(C.c_uint32)),
                ("n_tris", C.c_uint64), ("vert_pos", C.POINTER(C.c_float)),
                ("vert_nrm", C.POINTER(C.c_float)), ("vert_uv", C.POINTER(C.c_float)),
                ("n_verts", C.c_uint64), ("materials", C.POINTER(Material)), ("n_materials", C.c_uint32),
                ("lights", C.POINTER(Light)), ("n_lights", C.c_uint32),
                ("light_power", C.POINTER(C.c_float)), ("light_cdf", C.POINTER(C.c_float)),
                ("light_tri_area", C.POINTER(C.c_float)), ("light_tri_cdf", C.POINTER(C.c_float)),
                ("n_light_tri_area", C.c_uint32), ("n_light_tri_cdf", C.c_uint32),
                ("world_bound", C.c_float * 6), ("camera", Camera), ("film", FilmDesc),
                ("setting", RenderSetting), ("textures", C.POINTER(Texture)), ("n_textures", C.c_uint32),
                ("image_levels", C.POINTER(ImageLevel)), ("n_image_levels", C.c_uint32),
                ("image_texels", C.POINTER(C.c_float)), ("n_image_texels", C.c_uint64),
                ("light_dist", C.POINTER(C.c_float)), ("n_light_dist", C.c_uint64)]


class LoadOptions(C.Structure):
    _fields_ = [("bvh_method", C.c_int32), ("reserved", C.c_int32 * 7)]


# GB_BVH_*: equal_count = the reference's tree (parity), middle = its unused other method,
# sah = the non-parity fast tree
BVH_METHODS = {"equal_count": 0, "middle": 1, "sah": 2}


class RenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("spp_total", C.c_int32), ("spp_begin", C.c_int32),
                ("spp_end", C.c_int32), ("max_ray_depth", C.c_int32), ("method", C.c_int32),
                ("ao_sample_num", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("camera_samples", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_any", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("prims_tested", C.c_uint64),
                ("instances_entered", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("nodes_visited_any", C.c_uint64), ("prims_tested_any", C.c_uint64),
                ("instances_entered_any", C.c_uint64)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("mint", np.float32), ("maxt", np.float32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("eps", np.float32), ("inst", np.int32), ("prim", np.int32)])
NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("bmax", np.float32, 3), ("offset", np.uint32),
                       ("nprims", np.uint8), ("axis", np.uint8), ("pad", np.uint8, 2)])

# every symbol include/goblin_b200.h declares
EXPORTS = [
    "gb_scene_load_json", "gb_scene_load_json_string", "gb_scene_load_json_ex",
    "gb_scene_load_json_string_ex", "gb_scene_destroy", "gb_scene_get_desc",
    "gb_scene_output_path", "gb_bvh_build", "gb_bvh_build_method", "gb_device_count", "gb_create", "gb_destroy",
    "gb_upload_scene", "gb_trace_closest", "gb_trace_any", "gb_trace_closest_device",
    "gb_trace_any_device", "gb_camera_rays", "gb_li", "gb_render", "gb_film_clear",
    "gb_film_download", "gb_film_upload", "gb_film_device_ptr", "gb_film_write", "gb_write_image", "gb_film_resolve", "gb_write_rgb",
    "gb_synchronize", "gb_stream", "gb_enable_counters", "gb_get_counters", "gb_reset_counters",
    "gb_last_kernel_ms", "gb_last_error", "gb_version", "gb_enable_kernel_timing", "gb_get_kernel_times",
    "gb_reset_kernel_times", "gb_set_wave_paths", "gb_set_tuning", "gb_upload_bytes",
    "gb_set_trace_mode", "gb_get_trace_mode", "gb_upload_scene_async",
    "gb_comm_init_all", "gb_comm_unique_id", "gb_comm_init_rank", "gb_comm_attach", "gb_comm_destroy", "gb_comm_size",
    "gb_film_allreduce", "gb_film_allreduce_all", "gb_nccl_version", "gb_debug_stack_violation",
]
COMM_ID_BYTES = 128

# GB_TRACE_*: how the traversal kernels walk the reference's tree
TRACE_MODES = {"wide": 0, "pair": 1}

KERNEL_CLASSES = ["raygen", "extend", "shade", "shadow", "ao", "film", "trace", "other"]


class KernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 8), ("launches", C.c_uint64 * 8)]

_lib = None


def lib():
    """Load libgoblin_b200.so (once).  Raises OSError if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C goblin_b200/csrc`")
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        l.gb_last_error.restype = C.c_char_p
        l.gb_version.restype = C.c_char_p
        l.gb_scene_output_path.restype = C.c_char_p
        l.gb_scene_output_path.argtypes = [C.c_void_p]
        l.gb_scene_destroy.restype = None
        l.gb_scene_destroy.argtypes = [C.c_void_p]
        l.gb_scene_load_json.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        l.gb_scene_load_json_string.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
        l.gb_scene_load_json_ex.argtypes = [C.c_char_p, C.POINTER(LoadOptions), C.POINTER(C.c_void_p)]
        l.gb_scene_load_json_string_ex.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(LoadOptions),
                                                   C.POINTER(C.c_void_p)]
        l.gb_bvh_build_method.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint32),
                                          C.c_void_p]
        l.gb_scene_get_desc.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        l.gb_bvh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p]
        l.gb_device_count.argtypes = [C.POINTER(C.c_int)]
        l.gb_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        l.gb_destroy.argtypes = [C.c_void_p]
        l.gb_upload_scene.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        l.gb_upload_scene_async.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        l.gb_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.gb_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.gb_trace_closest_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.gb_trace_any_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.gb_camera_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.gb_li.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
        l.gb_render.argtypes = [C.c_void_p, C.POINTER(RenderParams)]
        l.gb_film_clear.argtypes = [C.c_void_p]
        l.gb_film_download.argtypes = [C.c_void_p, C.c_void_p]
        l.gb_film_upload.argtypes = [C.c_void_p, C.c_void_p]
        l.gb_film_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.gb_film_write.argtypes = [C.c_void_p, C.c_char_p]
        l.gb_write_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        l.gb_film_resolve.argtypes = [C.c_void_p, C.c_void_p]
        l.gb_write_rgb.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        l.gb_synchronize.argtypes = [C.c_void_p]
        l.gb_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        l.gb_enable_counters.argtypes = [C.c_void_p, C.c_int]
        l.gb_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
        l.gb_reset_counters.argtypes = [C.c_void_p]
        l.gb_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        l.gb_enable_kernel_timing.argtypes = [C.c_void_p, C.c_int]
        l.gb_get_kernel_times.argtypes = [C.c_void_p, C.POINTER(KernelTimes)]
        l.gb_reset_kernel_times.argtypes = [C.c_void_p]
        l.gb_set_wave_paths.argtypes = [C.c_void_p, C.c_size_t]
        l.gb_upload_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        l.gb_set_tuning.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        l.gb_comm_init_all.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        l.gb_comm_unique_id.argtypes = [C.c_void_p, C.c_size_t]
        l.gb_comm_init_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        l.gb_comm_attach.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        l.gb_comm_destroy.argtypes = [C.c_void_p]
        l.gb_comm_size.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        l.gb_film_allreduce.argtypes = [C.c_void_p]
        l.gb_film_allreduce_all.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        l.gb_nccl_version.argtypes = [C.POINTER(C.c_int)]
        l.gb_debug_stack_violation.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        l.gb_set_trace_mode.argtypes = [C.c_void_p, C.c_int]
        l.gb_get_trace_mode.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise GoblinError(rc, lib().gb_last_error().decode(errors="replace"))


def _np(ptr, n, dtype):
    """View n items behind a ctypes pointer as a numpy array (no copy)."""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    addr = C.cast(ptr, C.c_void_p).value
    buf = (C.c_uint8 * (int(n) * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype)


class Scene:
    """Host-side flattened scene: ContextLoader::load of the reference
    (src/GoblinContextLoader.cpp:447-503)."""

    def __init__(self, path=None, json_text=None, scene_dir=None, accel="equal_count"):
        self._h = C.c_void_p()
        self.accel = accel
        opt = LoadOptions(bvh_method=BVH_METHODS[accel])
        if path is not None:
            check(lib().gb_scene_load_json_ex(os.fsencode(path), C.byref(opt), C.byref(self._h)))
        else:
            check(lib().gb_scene_load_json_string_ex(json_text.encode(), os.fsencode(scene_dir or "."),
                                                     C.byref(opt), C.byref(self._h)))
        self.desc = SceneDesc()
        check(lib().gb_scene_get_desc(self._h, C.byref(self.desc)))

    def close(self):
        if self._h:
            lib().gb_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def output_path(self):
        return lib().gb_scene_output_path(self._h).decode()

    # numpy views of the flattened arrays (valid while the Scene is alive)
    def top_nodes(self):
        return _np(self.desc.top_nodes, self.desc.n_top_nodes, NODE_DTYPE)

    def top_order(self):
        return _np(self.desc.top_order, self.desc.n_instances, np.uint32)

    def model_nodes(self):
        return _np(self.desc.model_nodes, self.desc.n_model_nodes, NODE_DTYPE)

    def model_order(self):
        return _np(self.desc.model_order, self.desc.n_tris, np.uint32)

    def tri_index(self):
        return _np(self.desc.tri_index, self.desc.n_tris * 3, np.uint32).reshape(-1, 3)

    def vert_pos(self):
        return _np(self.desc.vert_pos, self.desc.n_verts * 3, np.float32).reshape(-1, 3)

    def vert_nrm(self):
        return _np(self.desc.vert_nrm, self.desc.n_verts * 3, np.float32).reshape(-1, 3)

    def vert_uv(self):
        return _np(self.desc.vert_uv, self.desc.n_verts * 2, np.float32).reshape(-1, 2)

    def instances(self):
        return [self.desc.instances[i] for i in range(self.desc.n_instances)]

    def models(self):
        return [self.desc.models[i] for i in range(self.desc.n_models)]

    def lights(self):
        return [self.desc.lights[i] for i in range(self.desc.n_lights)]

    def light_cdf(self):
        return _np(self.desc.light_cdf, self.desc.n_lights + 1, np.float32)

    def light_power(self):
        return _np(self.desc.light_power, self.desc.n_lights, np.float32)

    def sample_range(self):
        f = self.desc.film
        return f.sx0, f.sx1, f.sy0, f.sy1

    def spp_squared(self, spp=None):
        """spp rounded up to a perfect square (src/GoblinSampler.cpp:72)."""
        spp = self.desc.setting.spp if spp is None else spp
        root = int(np.ceil(np.sqrt(np.float32(spp))))
        return root * root

    def camera_samples(self, spp=None):
        sx0, sx1, sy0, sy1 = self.sample_range()
        return (sx1 - sx0) * (sy1 - sy0) * self.spp_squared(spp)


def bvh_build(aabbs, method="equal_count"):
    """BVH::BVH on raw boxes (n x 6 float32) -> (nodes, order)."""
    aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
    n = aabbs.shape[0]
    nodes = np.zeros(max(2 * n, 1), dtype=NODE_DTYPE)
    order = np.zeros(max(n, 1), dtype=np.uint32)
    cnt = C.c_uint32()
    check(lib().gb_bvh_build_method(aabbs.ctypes.data, n, BVH_METHODS[method], nodes.ctypes.data, C.byref(cnt),
                                    order.ctypes.data))
    return nodes[:cnt.value].copy(), order[:n].copy()


class Context:
    """One GPU.  Every method is a direct call through the C ABI."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        check(lib().gb_create(device, C.byref(self._h)))
        self.scene = None
        self.device = device

    def close(self):
        if self._h:
            lib().gb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, scene):
        check(lib().gb_upload_scene(self._h, C.byref(scene.desc)))
        self.scene = scene

    def upload_scene_async(self, scene):
        """gb_upload_scene_async: stage + copy on the copy stream while queued kernels still read the current
        scene; what is queued afterwards uses the new one.  Does not clear the film."""
        check(lib().gb_upload_scene_async(self._h, C.byref(scene.desc)))
        self.scene = scene

    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        check(lib().gb_trace_closest(self._h, rays.ctypes.data, rays.shape[0], hits.ctypes.data))
        return hits

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        check(lib().gb_trace_any(self._h, rays.ctypes.data, rays.shape[0], occ.ctypes.data))
        return occ

    def trace_closest_device(self, d_rays_ptr, n, d_hits_ptr):
        check(lib().gb_trace_closest_device(self._h, d_rays_ptr, n, d_hits_ptr))

    def trace_any_device(self, d_rays_ptr, n, d_occ_ptr):
        check(lib().gb_trace_any_device(self._h, d_rays_ptr, n, d_occ_ptr))

    def camera_rays(self, samples):
        samples = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1, 4)
        rays = np.zeros((samples.shape[0], 8), dtype=np.float32)
        check(lib().gb_camera_rays(self._h, samples.ctypes.data, samples.shape[0], rays.ctypes.data))
        return rays

    def li(self, samples):
        samples = np.ascontiguousarray(samples, dtype=np.float32)
        n, row = samples.shape
        out = np.zeros((n, 3), dtype=np.float32)
        check(lib().gb_li(self._h, samples.ctypes.data, n, row, out.ctypes.data))
        return out

    def render(self, seed=1, spp_total=None, spp_begin=0, spp_end=None, max_ray_depth=0, method=-1,
               ao_sample_num=0):
        if spp_total is None:
            spp_total = self.scene.spp_squared()
        if spp_end is None:
            spp_end = spp_total
        p = RenderParams(seed, spp_total, spp_begin, spp_end, max_ray_depth, method, ao_sample_num)
        check(lib().gb_render(self._h, C.byref(p)))

    def film_clear(self):
        check(lib().gb_film_clear(self._h))

    def film_download(self, out=None):
        """Film::mPixels as yres x xres x (r, g, b, weight); `out` reuses a caller-owned buffer."""
        f = self.scene.desc.film
        if out is None:
            out = np.empty((f.yres, f.xres, 4), dtype=np.float32)
        assert out.shape == (f.yres, f.xres, 4) and out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]
        check(lib().gb_film_download(self._h, out.ctypes.data))
        return out

    def film_upload(self, rgbw):
        f = self.scene.desc.film
        rgbw = np.ascontiguousarray(rgbw, dtype=np.float32)
        if rgbw.size != f.yres * f.xres * 4:  # gb_film_upload copies exactly the film's size from this buffer
            raise ValueError(f"film_upload: expected {f.yres} x {f.xres} x 4 floats, got {rgbw.size}")
        check(lib().gb_film_upload(self._h, rgbw.ctypes.data))

    def film_device_ptr(self):
        p = C.c_void_p()
        n = C.c_size_t()
        check(lib().gb_film_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def film_write(self, path):
        check(lib().gb_film_write(self._h, os.fsencode(path)))

    def film_resolve(self):
        """colour / weight + the film's bloom, on the device: what Film::writeImage hands to
        Goblin::writeImage (yres x xres x 3)."""
        f = self.scene.desc.film
        out = np.zeros((f.yres, f.xres, 3), dtype=np.float32)
        check(lib().gb_film_resolve(self._h, out.ctypes.data))
        return out

    def synchronize(self):
        check(lib().gb_synchronize(self._h))

    def stream(self):
        s = C.c_void_p()
        check(lib().gb_stream(self._h, C.byref(s)))
        return s.value or 0

    def enable_counters(self, on=True):
        check(lib().gb_enable_counters(self._h, 1 if on else 0))

    def counters(self):
        c = Counters()
        check(lib().gb_get_counters(self._h, C.byref(c)))
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    def reset_counters(self):
        check(lib().gb_reset_counters(self._h))

    def enable_kernel_timing(self, on=True):
        check(lib().gb_enable_kernel_timing(self._h, 1 if on else 0))

    def kernel_times(self):
        """{class: (ms, launches)} since the last reset_kernel_times()."""
        t = KernelTimes()
        check(lib().gb_get_kernel_times(self._h, C.byref(t)))
        return {name: (t.ms[k], t.launches[k]) for k, name in enumerate(KERNEL_CLASSES)}

    def reset_kernel_times(self):
        check(lib().gb_reset_kernel_times(self._h))

    def set_wave_paths(self, max_paths):
        check(lib().gb_set_wave_paths(self._h, max_paths))

    def upload_bytes(self):
        n = C.c_size_t()
        check(lib().gb_upload_bytes(self._h, C.byref(n)))
        return n.value

    def set_tuning(self, values):
        arr = (C.c_int * len(values))(*values)
        check(lib().gb_set_tuning(self._h, arr, len(values)))

    # ---- Film::mergeTile across GPUs: NCCL all-reduce of the device film, issued by the library
    def comm_init_rank(self, comm_id, nranks, rank):
        """One process per GPU: join the film communicator (comm_id: the 128 bytes of comm_unique_id()
        made on rank 0 and carried here by the launcher's own channel)."""
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(comm_id))
        check(lib().gb_comm_init_rank(self._h, buf, COMM_ID_BYTES, nranks, rank))

    def comm_size(self):
        n = C.c_int()
        check(lib().gb_comm_size(self._h, C.byref(n)))
        return n.value

    def comm_destroy(self):
        check(lib().gb_comm_destroy(self._h))

    def film_allreduce(self):
        """Sum the device film over all ranks, in place, asynchronously on the context's stream."""
        check(lib().gb_film_allreduce(self._h))

    def debug_stack_violation(self):
        out = (C.c_int * 4)()
        check(lib().gb_debug_stack_violation(self._h, out))
        return list(out)

    def set_trace_mode(self, mode):
        """"pair" (default: pair nodes, every box test of the reference; "exact" is its older name) or
        "wide" (4-wide nodes)."""
        check(lib().gb_set_trace_mode(self._h, TRACE_MODES["pair" if mode == "exact" else mode]))

    def trace_mode(self):
        m = C.c_int()
        check(lib().gb_get_trace_mode(self._h, C.byref(m)))
        return {v: k for k, v in TRACE_MODES.items()}[m.value]

    def last_kernel_ms(self):
        ms = C.c_float()
        check(lib().gb_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value


def comm_unique_id():
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    check(lib().gb_comm_unique_id(buf, COMM_ID_BYTES))
    return bytes(buf)


def comm_init_all(contexts):
    """One process, one Context per GPU: a communicator over all of them (ncclCommInitAll)."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    check(lib().gb_comm_init_all(arr, len(contexts)))


def film_allreduce_all(contexts):
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    check(lib().gb_film_allreduce_all(arr, len(contexts)))


def nccl_version():
    v = C.c_int()
    check(lib().gb_nccl_version(C.byref(v)))
    return v.value


def device_count():
    n = C.c_int()
    check(lib().gb_device_count(C.byref(n)))
    return n.value


def write_rgb(path, rgb, tone_mapping=False):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    check(lib().gb_write_rgb(os.fsencode(path), rgb.ctypes.data, rgb.shape[1], rgb.shape[0], int(tone_mapping)))


def write_image(path, rgbw):
    rgbw = np.ascontiguousarray(rgbw, dtype=np.float32)
    check(lib().gb_write_image(os.fsencode(path), rgbw.ctypes.data, rgbw.shape[1], rgbw.shape[0]))
