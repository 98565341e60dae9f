"""Thin ctypes wrapper over librgk_b200.so for tests, bench.py and tools.

Everything here goes through the C ABI (include/rgk_b200.h); numpy arrays are host buffers,
integer addresses (e.g. torch tensor .data_ptr()) are device buffers.  There is no CPU path:
creating a Context without a CUDA device raises.
"""
import ctypes as C

import numpy as np

from . import abi

RAY_DT = np.dtype([("origin", np.float32, 3), ("direction", np.float32, 3), ("tnear", np.float32), ("tfar", np.float32)])
HIT_DT = np.dtype([("triangle", np.uint32), ("t", np.float32), ("a", np.float32), ("b", np.float32), ("c", np.float32)])


class RgkError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"rgk status {status}: {text}")
        self.status = status


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Context:
    def __init__(self, device=0, stream=None, lib=None, cfg=None, **cfg_fields):
        """cfg: an abi.DeviceCfg, or keyword fields of rgk_device_cfg changed from the library's defaults
        (e.g. traversal="kd", chunk_paths=1 << 20).  Applied before any scene is committed."""
        self.lib = lib or abi.load_library()
        h = C.c_void_p()
        # stream: None -> the library creates its own stream; an integer cudaStream_t handle -> run on that stream
        # (torch's default stream has handle 0, which CUDA spells cudaStreamLegacy = 0x1)
        if stream is not None and int(stream) == 0:
            stream = 1
        st = self.lib.rgk_context_create(int(device), C.c_void_p(stream) if stream is not None else None, C.byref(h))
        if st != 0:
            raise RgkError(st, self.lib.rgk_last_error(None).decode())
        self.h = h
        self._desc = None
        if cfg is not None or cfg_fields:
            self.configure(cfg, **cfg_fields)

    def configure(self, cfg=None, **fields):
        """rgk_context_configure: host-build / traversal fields take effect at the next commit, the others at the next call."""
        if cfg is None:
            cfg = self.cfg()
        for k, v in fields.items():
            if k == "traversal" and isinstance(v, str):
                v = {"bvh": abi.TRAVERSAL_BVH, "kd": abi.TRAVERSAL_KD}[v]
            if not hasattr(cfg, k) or k.startswith("_") or k == "struct_size":
                raise AttributeError("rgk_device_cfg has no field " + k)
            setattr(cfg, k, v)
        self._check(self.lib.rgk_context_configure(self.h, C.byref(cfg)))

    def cfg(self):
        out = abi.DeviceCfg()
        self._check(self.lib.rgk_context_get_cfg(self.h, C.byref(out)))
        return out

    def close(self):
        if self.h:
            self.lib.rgk_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != 0:
            raise RgkError(st, self.lib.rgk_last_error(self.h).decode())

    # ---- scene
    def commit(self, desc, tree=None):
        self._desc = desc
        self._check(self.lib.rgk_scene_commit(self.h, C.byref(desc), C.byref(tree) if tree is not None else None))

    def scene_info(self):
        info = abi.SceneInfo()
        self._check(self.lib.rgk_scene_get_info(self.h, C.byref(info)))
        return info

    def scene_kdtree(self):
        info = self.scene_info()
        nodes = np.zeros(2 * info.n_nodes, np.uint32)
        refs = np.zeros(max(1, info.n_refs), np.uint32)
        self._check(self.lib.rgk_scene_get_kdtree(self.h, _p(nodes), _p(refs)))
        return nodes, refs[:info.n_refs]

    # ---- traversal (host buffers)
    def trace_closest(self, rays, ignore=None, want_stats=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DT)
        hits = np.zeros(len(rays), dtype=HIT_DT)
        st = abi.TravStats()
        if ignore is not None:
            ignore = np.ascontiguousarray(ignore, np.uint32)
        self._check(self.lib.rgk_trace_closest(self.h, _p(rays), _p(ignore), C.c_uint64(len(rays)), _p(hits),
                                               C.byref(st) if want_stats else None))
        return (hits, st) if want_stats else hits

    def trace_shadow(self, a, b, want_stats=False):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        vis = np.zeros(len(a), np.uint8)
        st = abi.TravStats()
        self._check(self.lib.rgk_trace_shadow(self.h, _p(a), _p(b), C.c_uint64(len(a)), _p(vis),
                                              C.byref(st) if want_stats else None))
        return (vis, st) if want_stats else vis

    # ---- traversal (device buffers: integer addresses)
    def trace_closest_device(self, d_rays, d_ignore, n, d_hits, d_stats=None):
        self._check(self.lib.rgk_trace_closest_device(self.h, _p(d_rays), _p(d_ignore), C.c_uint64(n), _p(d_hits), _p(d_stats)))

    def trace_shadow_device(self, d_a, d_b, n, d_visible, d_stats=None):
        self._check(self.lib.rgk_trace_shadow_device(self.h, _p(d_a), _p(d_b), C.c_uint64(n), _p(d_visible), _p(d_stats)))

    # ---- camera / tasks / sampler
    def camera(self, pos, lookat, up, yview, xview, xres, yres, focus_plane=1.0, lens_size=0.0):
        cam = abi.Camera()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        self.lib.rgk_camera_init(C.byref(cam), f3(pos), f3(lookat), f3(up), yview, xview, xres, yres, focus_plane, lens_size)
        return cam

    def camera_rays(self, cam, xres, yres, xy, offsets, lens=None):
        xy = np.ascontiguousarray(xy, np.int32)
        offsets = np.ascontiguousarray(offsets, np.float32)
        if lens is not None:
            lens = np.ascontiguousarray(lens, np.float32)
        rays = np.zeros(len(xy), dtype=RAY_DT)
        self._check(self.lib.rgk_camera_rays(self.h, C.byref(cam), xres, yres, _p(xy), _p(offsets), _p(lens),
                                             C.c_uint64(len(xy)), _p(rays)))
        return rays

    def generate_tasks(self, tile, xres, yres):
        n = self.lib.rgk_generate_tasks(tile, xres, yres, None, 0)
        out = (abi.Task * n)()
        self.lib.rgk_generate_tasks(tile, xres, yres, out, n)
        return out

    def sampler_set_size(self, ms):
        return int(self.lib.rgk_sampler_set_size(ms))

    def sampler_tables(self, seeds, ms, n1d, n2d):
        seeds = np.ascontiguousarray(seeds, np.uint32)
        ss = self.sampler_set_size(ms)
        t1 = np.zeros((len(seeds), n1d, ss), np.float32)
        t2 = np.zeros((len(seeds), n2d, ss, 2), np.float32)
        self._check(self.lib.rgk_sampler_tables(self.h, _p(seeds), len(seeds), ms, n1d, n2d, _p(t1), _p(t2)))
        return t1, t2

    # ---- rendering
    def render_round(self, cam, params, tasks, seedstart=42, seedcount_base=0, fb=None):
        if fb is None:
            fb = (np.zeros((params.yres, params.xres, 3), np.float32), np.zeros((params.yres, params.xres), np.uint32))
        st = abi.RoundStats()
        self._check(self.lib.rgk_render_round(self.h, C.byref(cam), C.byref(params), tasks, len(tasks), seedstart,
                                              seedcount_base, _p(fb[0]), _p(fb[1]), C.byref(st)))
        return fb[0], fb[1], st

    def render_round_device(self, cam, params, tasks, d_rgb, d_count, seedstart=42, seedcount_base=0):
        st = abi.RoundStats()
        self._check(self.lib.rgk_render_round_device(self.h, C.byref(cam), C.byref(params), tasks, len(tasks), seedstart,
                                                     seedcount_base, _p(d_rgb), _p(d_count), C.byref(st)))
        return st

    def render_frame(self, cam, params, rounds=1):
        fb = np.zeros((params.yres, params.xres, 3), np.float32)
        cnt = np.zeros((params.yres, params.xres), np.uint32)
        st = abi.RoundStats()
        self._check(self.lib.rgk_render_frame(self.h, C.byref(cam), C.byref(params), rounds, _p(fb), _p(cnt), C.byref(st)))
        return fb, cnt, st

    def probe(self, kind, index, rows):
        rows = np.ascontiguousarray(rows, np.float32)
        win, wout = abi.PROBE_WIDTHS[kind]
        assert rows.ndim == 2 and rows.shape[1] == win
        out = np.zeros((len(rows), wout), np.float32)
        self._check(self.lib.rgk_probe(self.h, kind, index, _p(rows), C.c_uint64(len(rows)), _p(out)))
        return out

    def set_tables(self, multisample, t1, t2):
        """t1[pixel][dim][set], t2[pixel][dim][set][2] for the pixels of the next render calls (RGK_SAMPLER_TABLES)."""
        t1 = np.ascontiguousarray(t1, np.float32)
        t2 = np.ascontiguousarray(t2, np.float32)
        self._check(self.lib.rgk_render_set_tables(self.h, multisample, t1.shape[1], t2.shape[1], _p(t1), _p(t2), C.c_uint64(t1.shape[0])))

    def set_shard(self, first, stride):
        self._check(self.lib.rgk_render_set_shard(self.h, first, stride))

    def set_counting(self, enabled):
        self._check(self.lib.rgk_render_set_counting(self.h, int(bool(enabled))))

    def render_trav_stats(self):
        a, b = abi.TravStats(), abi.TravStats()
        self._check(self.lib.rgk_render_get_trav_stats(self.h, C.byref(a), C.byref(b)))
        return a, b

    def accumulate_device(self, d_dst, d_src, n_floats, d_count=None, d_other_count=None, stream=None):
        """EXRTexture::Accumulate on device buffers (integer addresses); stream: a cudaStream_t handle, None = the context stream."""
        if stream is not None and int(stream) == 0:
            stream = 1
        self._check(self.lib.rgk_accumulate_device(self.h, _p(d_dst), _p(d_src), C.c_uint64(n_floats), _p(d_count), _p(d_other_count),
                                                   C.c_void_p(stream) if stream is not None else None))

    def shade_stats(self):
        """k_shade work counters of the last counting round (rgk_render_get_shade_stats)."""
        out = (C.c_uint64 * 8)()
        self._check(self.lib.rgk_render_get_shade_stats(self.h, C.byref(out)))
        return dict(zip(("vertices", "sky_vertices", "texels", "ltc_evals", "light_evals", "continuations"), (int(x) for x in out)))

    def synchronize(self):
        self._check(self.lib.rgk_synchronize(self.h))

    def bvh_stats(self):
        """Counters of the wide-BVH launches since the previous call (all 0 on the kd-only traversal)."""
        out = (C.c_uint64 * 10)()
        self._check(self.lib.rgk_bvh_stats(self.h, C.byref(out)))
        names = ("rays", "ambiguous", "nodes", "tests", "slots")
        d = {n: out[k] + out[5 + k] for k, n in enumerate(names)}           # totals; per query kind below
        d["closest"] = {n: out[k] for k, n in enumerate(names)}
        d["shadow"] = {n: out[5 + k] for k, n in enumerate(names)}
        return d


class HostScene:
    """Host-only scene commit (rgk_host_scene_*): planes, areal lights, epsilon, bbox, kd-tree build + flatten -- what
    rgk_scene_commit computes on the CPU before uploading.  Needs no GPU."""

    def __init__(self, desc, tree=None, lib=None, cfg=None, **cfg_fields):
        self.lib = lib or abi.load_library()
        h = C.c_void_p()
        if cfg is None and cfg_fields:
            cfg = abi.device_cfg(self.lib, **cfg_fields)
        st = self.lib.rgk_host_scene_create(C.byref(desc), C.byref(tree) if tree is not None else None,
                                            C.byref(cfg) if cfg is not None else None, C.byref(h))
        if st != 0:
            raise RgkError(st, self.lib.rgk_host_last_error().decode())
        self.h = h

    def info(self):
        info = abi.SceneInfo()
        self.lib.rgk_host_scene_get_info(self.h, C.byref(info))
        return info

    def kdtree(self):
        info = self.info()
        nodes = np.zeros(2 * info.n_nodes, np.uint32)
        refs = np.zeros(max(1, info.n_refs), np.uint32)
        self.lib.rgk_host_scene_get_kdtree(self.h, _p(nodes), _p(refs))
        return nodes, refs[:info.n_refs]

    def records(self):
        info = self.info()
        planes = np.zeros((info.n_triangles, 4), np.float32)
        rec = np.zeros((info.n_triangles, 12), np.float32)
        self.lib.rgk_host_scene_get_records(self.h, _p(planes), _p(rec))
        return planes, rec

    def bounds(self):
        b = np.zeros((self.info().n_triangles, 4), np.float32)
        self.lib.rgk_host_scene_get_bounds(self.h, _p(b))
        return b

    def bvh(self):
        """(nodes [n, 32] float32, order [n_triangles] uint32, depth) of the wide BVH; n == 0 when the scene was committed
        with traversal="kd" (or has NaN-prone triangles, which keep it on the kd-tree)."""
        n, slots, depth = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.lib.rgk_host_scene_get_bvh_size(self.h, C.byref(n), C.byref(slots), C.byref(depth))
        nodes = np.zeros((max(1, n.value), 32), np.float32)
        order = np.zeros(max(1, slots.value), np.uint32)
        if n.value:
            self.lib.rgk_host_scene_get_bvh(self.h, _p(nodes), _p(order))
        return nodes[:n.value], order[:slots.value], depth.value

    def close(self):
        if self.h:
            self.lib.rgk_host_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
