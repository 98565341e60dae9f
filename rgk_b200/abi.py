"""ctypes mirror of include/rgk_b200.h (struct layouts + the product library loader).

The product library is rgk_b200/librgk_b200.so (built by __graft_entry__.build()).
There is no CPU fallback: load_library() raises if the CUDA extension is missing.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RGK_B200_LIB") or os.path.join(_HERE, "librgk_b200.so")   # RGK_B200_LIB: A/B builds of the same library

RGK_NO_TRIANGLE = 0xFFFFFFFF
(BXDF_DIFFUSE, BXDF_MIX, BXDF_DIELECTRIC, BXDF_MIRROR, BXDF_TRANSPARENT, BXDF_LTC_BECKMANN,
 BXDF_LTC_GGX, BXDF_LTC_BECKMANN_DIFFUSE, BXDF_LTC_GGX_DIFFUSE) = range(9)
SAMPLER_MT19937, SAMPLER_TABLES, SAMPLER_FAST = 0, 1, 2
PROBE_BXDF_SAMPLE, PROBE_BXDF_VALUE, PROBE_TEXTURE, PROBE_RANDOM_LIGHT, PROBE_SKY, PROBE_FRAME = range(6)
PROBE_WIDTHS = {0: (7, 7), 1: (8, 3), 2: (2, 5), 3: (5, 12), 4: (3, 3), 5: (6, 6)}

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("color", C.c_float * 3), ("texels", f32p)]


class Material(C.Structure):
    _fields_ = [("bxdf", C.c_uint32), ("no_russian", C.c_uint32), ("emission", C.c_float * 3),
                ("roughness", C.c_float), ("ior", C.c_float), ("amount", C.c_float),
                ("mix_a", C.c_int32), ("mix_b", C.c_int32),
                ("tex_diffuse", C.c_int32), ("tex_color", C.c_int32), ("tex_bump", C.c_int32),
                ("_pad", C.c_uint32 * 3)]


class Mesh(C.Structure):
    _fields_ = [("first_triangle", C.c_uint32), ("n_triangles", C.c_uint32), ("material", C.c_uint32),
                ("_pad", C.c_uint32)]


class PointLight(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("color", C.c_float * 3), ("intensity", C.c_float), ("size", C.c_float)]


class Sky(C.Structure):
    _fields_ = [("mode", C.c_uint32), ("color", C.c_float * 3), ("intensity", C.c_float), ("rotate", C.c_float),
                ("envmap", C.c_int32)]


class LtcTable(C.Structure):
    _fields_ = [("M", f32p), ("amplitude", f32p)]


class SceneDesc(C.Structure):
    _fields_ = [("n_vertices", C.c_uint32), ("positions", f32p), ("normals", f32p), ("tangents", f32p),
                ("texcoords", f32p),
                ("n_triangles", C.c_uint32), ("indices", u32p),
                ("n_meshes", C.c_uint32), ("meshes", C.POINTER(Mesh)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(Material)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(Texture)),
                ("n_point_lights", C.c_uint32), ("point_lights", C.POINTER(PointLight)),
                ("sky", Sky), ("ltc_ggx", LtcTable), ("ltc_beckmann", LtcTable), ("thinglass", C.c_uint32)]


class KdTree(C.Structure):
    _fields_ = [("n_nodes", C.c_uint32), ("nodes", u32p), ("n_refs", C.c_uint32), ("refs", u32p)]


class SceneInfo(C.Structure):
    _fields_ = [("epsilon", C.c_float), ("bbox", C.c_float * 6), ("n_nodes", C.c_uint32), ("n_refs", C.c_uint32),
                ("n_triangles", C.c_uint32), ("n_areal_lights", C.c_uint32), ("max_depth", C.c_uint32),
                ("total_point_power", C.c_float), ("total_areal_power", C.c_float)]


class Ray(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("direction", C.c_float * 3), ("tnear", C.c_float), ("tfar", C.c_float)]


class Hit(C.Structure):
    _fields_ = [("triangle", C.c_uint32), ("t", C.c_float), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float)]


class TravStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("inner", C.c_uint64), ("leaf", C.c_uint64), ("refs", C.c_uint64),
                ("tests", C.c_uint64), ("exact", C.c_uint64), ("prefiltered", C.c_uint64), ("prefilter_wrong", C.c_uint64)]

    def as_dict(self):
        """The algorithmic counters (identical on the CPU checkers and the device)."""
        return {k: int(getattr(self, k)) for k in ("rays", "inner", "leaf", "refs", "tests")}

    def device_dict(self):
        return {k: int(getattr(self, k)) for k in ("exact", "prefiltered", "prefilter_wrong")}

    def bytes_per_ray(self, out_bytes=20):
        """36 in + out + 8*inner + 8*leaf + 4*refs + 48*tests, averaged over the batch."""
        n = max(1, int(self.rays))
        return 36 + out_bytes + (8 * self.inner + 8 * self.leaf + 4 * self.refs + 48 * self.tests) / n


class Camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lookat", C.c_float * 3), ("direction", C.c_float * 3),
                ("cameraup", C.c_float * 3), ("cameraleft", C.c_float * 3), ("viewscreen", C.c_float * 3),
                ("viewscreen_x", C.c_float * 3), ("viewscreen_y", C.c_float * 3), ("lens_size", C.c_float),
                ("xsize", C.c_int32), ("ysize", C.c_int32)]


class RenderParams(C.Structure):
    _fields_ = [("xres", C.c_uint32), ("yres", C.c_uint32), ("multisample", C.c_uint32), ("depth", C.c_uint32),
                ("clamp", C.c_float), ("russian", C.c_float), ("bumpmap_scale", C.c_float),
                ("force_fresnell", C.c_uint32), ("reverse", C.c_uint32), ("sampler_mode", C.c_uint32)]


class Task(C.Structure):
    _fields_ = [("x1", C.c_uint32), ("x2", C.c_uint32), ("y1", C.c_uint32), ("y2", C.c_uint32)]


class RoundStats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("samples", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("gpu_ms", C.c_float), ("trace_ms", C.c_float),
                ("closest_ms", C.c_float), ("shadow_ms", C.c_float), ("sampler_ms", C.c_float), ("shade_ms", C.c_float),
                ("closest_launches", C.c_uint32), ("shadow_launches", C.c_uint32), ("shadow_rays_skipped", C.c_uint64)]

    def as_dict(self):
        return {k: (int(getattr(self, k)) if "ms" not in k else float(getattr(self, k))) for k, _ in self._fields_}


TRAVERSAL_BVH, TRAVERSAL_KD = 0, 1


class DeviceCfg(C.Structure):
    """rgk_device_cfg: how the library runs the path on the device (the library reads no environment variable)."""
    _fields_ = [("struct_size", C.c_uint32), ("traversal", C.c_uint32),
                ("build_threads", C.c_uint32), ("bvh_bins", C.c_uint32), ("bvh_all_axes", C.c_uint32), ("bvh_leaf_max", C.c_uint32),
                ("bvh_greedy_collapse", C.c_uint32), ("bvh_c_prim", C.c_float), ("bvh_reinsert_iters", C.c_uint32),
                ("bvh_reinsert_frac", C.c_float),
                ("chunk_paths", C.c_uint64), ("table_bytes", C.c_uint64), ("reverse_bytes", C.c_uint64),
                ("refill_batch", C.c_uint32), ("refill_coherent", C.c_uint32), ("refill_incoherent", C.c_uint32),
                ("refill_shadow", C.c_uint32), ("binning", C.c_uint32), ("bin_shadow_first", C.c_uint32), ("bin_items", C.c_uint32),
                ("bin_min_frac", C.c_float), ("shade_path_order", C.c_uint32), ("skip_null_shadow", C.c_uint32),
                ("const_light", C.c_uint32), ("arb_grid", C.c_uint32), ("bvh_shadow_nosort", C.c_uint32),
                ("bvh_closest_nearest", C.c_uint32), ("sampler_smem", C.c_uint32), ("trace_threads", C.c_uint32),
                ("kd_variant", C.c_uint32), ("sampler_ctas_per_sm", C.c_uint32), ("sampler_kernel", C.c_uint32),
                ("sampler_slots", C.c_uint32), ("_reserved", C.c_uint32 * 5)]


def device_cfg(lib=None, **fields):
    """The library's defaults (rgk_device_cfg_init) with `fields` changed; traversal may be 'bvh' / 'kd'."""
    lib = lib or load_library()
    cfg = DeviceCfg()
    lib.rgk_device_cfg_init(C.byref(cfg))
    for k, v in fields.items():
        if k == "traversal" and isinstance(v, str):
            v = {"bvh": TRAVERSAL_BVH, "kd": TRAVERSAL_KD}[v]
        if not hasattr(cfg, k) or k.startswith("_") or k == "struct_size":
            raise AttributeError("rgk_device_cfg has no field " + k)
        setattr(cfg, k, v)
    return cfg


assert C.sizeof(Material) == 64 and C.sizeof(Ray) == 32 and C.sizeof(Hit) == 20

# Every symbol include/rgk_b200.h declares (tests check that the library exports all of them).
ABI_VERSION = 4      # RGK_ABI_VERSION of include/rgk_b200.h these ctypes structures mirror

EXPORTS = [
    "rgk_abi_version", "rgk_status_string", "rgk_context_create", "rgk_context_destroy", "rgk_last_error",
    "rgk_scene_commit", "rgk_scene_get_info", "rgk_scene_get_kdtree", "rgk_trace_closest", "rgk_trace_shadow",
    "rgk_trace_closest_device", "rgk_trace_shadow_device", "rgk_camera_init", "rgk_camera_rays",
    "rgk_generate_tasks", "rgk_sampler_set_size", "rgk_sampler_tables", "rgk_render_round",
    "rgk_render_round_device", "rgk_render_frame", "rgk_render_set_tables", "rgk_synchronize",
    "rgk_render_set_counting", "rgk_render_get_trav_stats", "rgk_probe", "rgk_render_set_shard",
    "rgk_host_scene_create", "rgk_host_scene_destroy", "rgk_host_last_error", "rgk_host_scene_get_info",
    "rgk_host_scene_get_kdtree", "rgk_host_scene_get_records", "rgk_host_scene_get_bounds",
    "rgk_host_scene_get_bvh_size", "rgk_host_scene_get_bvh", "rgk_bvh_stats",
    "rgk_device_cfg_init", "rgk_context_configure", "rgk_context_get_cfg", "rgk_accumulate_device", "rgk_render_get_shade_stats",
]


_LIBS = {}


def load_library(path=None):
    """dlopen the CUDA library. Fails loudly when it has not been built (no CPU fallback exists)."""
    path = path or LIB_PATH
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "rgk_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.rgk_abi_version.restype = C.c_uint32
    if lib.rgk_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.rgk_abi_version()}, this binding expects {ABI_VERSION} (rebuild: __graft_entry__.build())")
    lib.rgk_status_string.restype = C.c_char_p
    lib.rgk_status_string.argtypes = [C.c_int]
    lib.rgk_context_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.rgk_context_destroy.argtypes = [vp]
    lib.rgk_context_destroy.restype = None
    lib.rgk_last_error.restype = C.c_char_p
    lib.rgk_last_error.argtypes = [vp]
    lib.rgk_scene_commit.argtypes = [vp, C.POINTER(SceneDesc), C.POINTER(KdTree)]
    lib.rgk_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    lib.rgk_scene_get_kdtree.argtypes = [vp, vp, vp]
    lib.rgk_trace_closest.argtypes = [vp, vp, vp, C.c_uint64, vp, C.POINTER(TravStats)]
    lib.rgk_trace_shadow.argtypes = [vp, vp, vp, C.c_uint64, vp, C.POINTER(TravStats)]
    lib.rgk_trace_closest_device.argtypes = [vp, vp, vp, C.c_uint64, vp, vp]
    lib.rgk_trace_shadow_device.argtypes = [vp, vp, vp, C.c_uint64, vp, vp]
    lib.rgk_camera_init.argtypes = [C.POINTER(Camera), f32p, f32p, f32p, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                    C.c_float, C.c_float]
    lib.rgk_camera_init.restype = None
    lib.rgk_camera_rays.argtypes = [vp, C.POINTER(Camera), C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint64, vp]
    lib.rgk_generate_tasks.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Task), C.c_uint32]
    lib.rgk_generate_tasks.restype = C.c_uint32
    lib.rgk_sampler_set_size.argtypes = [C.c_uint32]
    lib.rgk_sampler_set_size.restype = C.c_uint32
    lib.rgk_sampler_tables.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    rr = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.POINTER(Task), C.c_uint32, C.c_uint32, C.c_uint32,
          vp, vp, C.POINTER(RoundStats)]
    lib.rgk_render_round.argtypes = rr
    lib.rgk_render_round_device.argtypes = rr
    lib.rgk_render_frame.argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.c_uint32, vp, vp,
                                     C.POINTER(RoundStats)]
    lib.rgk_render_set_tables.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint64]
    lib.rgk_synchronize.argtypes = [vp]
    lib.rgk_probe.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp]
    lib.rgk_host_scene_create.argtypes = [C.POINTER(SceneDesc), C.POINTER(KdTree), C.POINTER(DeviceCfg), C.POINTER(vp)]
    lib.rgk_device_cfg_init.argtypes = [C.POINTER(DeviceCfg)]
    lib.rgk_device_cfg_init.restype = None
    lib.rgk_context_configure.argtypes = [vp, C.POINTER(DeviceCfg)]
    lib.rgk_context_get_cfg.argtypes = [vp, C.POINTER(DeviceCfg)]
    lib.rgk_render_get_shade_stats.argtypes = [vp, C.POINTER(C.c_uint64 * 8)]
    lib.rgk_accumulate_device.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, vp]
    lib.rgk_host_scene_destroy.argtypes = [vp]
    lib.rgk_host_scene_destroy.restype = None
    lib.rgk_host_last_error.restype = C.c_char_p
    lib.rgk_host_scene_get_info.argtypes = [vp, C.POINTER(SceneInfo)]
    lib.rgk_host_scene_get_kdtree.argtypes = [vp, vp, vp]
    lib.rgk_host_scene_get_records.argtypes = [vp, vp, vp]
    lib.rgk_host_scene_get_bounds.argtypes = [vp, vp]
    lib.rgk_host_scene_get_bvh_size.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.rgk_host_scene_get_bvh.argtypes = [vp, vp, vp]
    lib.rgk_bvh_stats.argtypes = [vp, C.POINTER(C.c_uint64 * 10)]
    lib.rgk_render_set_counting.argtypes = [vp, C.c_int]
    lib.rgk_render_set_shard.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.rgk_render_get_trav_stats.argtypes = [vp, C.POINTER(TravStats), C.POINTER(TravStats)]
    _LIBS[path] = lib
    return lib
