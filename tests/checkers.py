"""Loaders for the two CPU checkers (test infrastructure): the oracle restatement
(oracle/_build/librgk_oracle.so, prefix rgko_) and the reference build
(oracle/_ref/librgk_ref.so, prefix rgkref_).  Product code never imports this."""
import contextlib
import ctypes as C
import os

import numpy as np

from rgk_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "librgk_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "librgk_ref.so")
vp = C.c_void_p


@contextlib.contextmanager
def scoped_env(**kv):
    """Set environment variables for the duration of a block (None = unset) and restore what was there before."""
    old = {k: os.environ.get(k) for k in kv}
    try:
        for k, v in kv.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


class Checker:
    """Uniform wrapper over the oracle ('rgko') and the reference build ('rgkref')."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.is_ref = prefix == "rgkref"
        f = self._f
        f("scene_create").restype = vp
        f("scene_destroy").argtypes = [vp]
        f("scene_destroy").restype = None
        f("describe").restype = C.c_char_p
        f("sampler_set_size").restype = C.c_uint32
        f("generate_tasks").restype = C.c_uint32

    def _f(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def describe(self):
        return self._f("describe")().decode()

    # ---- scene
    def scene_create(self, desc, tree=None):
        if self.is_ref:
            assert tree is None
            h = self._f("scene_create")(C.byref(desc))
        else:
            h = self._f("scene_create")(C.byref(desc), C.byref(tree) if tree is not None else None)
        if not h:
            raise RuntimeError("scene_create failed")
        return vp(h)

    def scene_destroy(self, h):
        self._f("scene_destroy")(h)

    def scene_info(self, h):
        info = abi.SceneInfo()
        self._f("scene_get_info")(h, C.byref(info))
        return info

    def scene_kdtree(self, h):
        info = self.scene_info(h)
        nodes = np.zeros(2 * info.n_nodes, np.uint32)
        refs = np.zeros(max(1, info.n_refs), np.uint32)
        self._f("scene_get_kdtree")(h, _ptr(nodes), _ptr(refs))
        return nodes, refs[:info.n_refs]

    def scene_planes(self, h):
        info = self.scene_info(h)
        p = np.zeros((info.n_triangles, 4), np.float32)
        self._f("scene_get_planes")(h, _ptr(p))
        return p

    # ---- traversal
    def trace_closest(self, h, rays, ignore=None, nthreads=8, want_stats=False):
        n = len(rays)
        hits = np.zeros(n, dtype=HIT_DT)
        if self.is_ref:
            self._f("trace_closest")(h, _ptr(rays), _ptr(ignore), C.c_uint64(n), _ptr(hits), nthreads)
            return hits
        st = abi.TravStats()
        self._f("trace_closest")(h, _ptr(rays), _ptr(ignore), C.c_uint64(n), _ptr(hits),
                                 C.byref(st) if want_stats else None, nthreads)
        return (hits, st) if want_stats else hits

    def trace_shadow(self, h, a, b, nthreads=8, want_stats=False):
        n = len(a)
        vis = np.zeros(n, np.uint8)
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        if self.is_ref:
            self._f("trace_shadow")(h, _ptr(a), _ptr(b), C.c_uint64(n), _ptr(vis), nthreads)
            return vis
        st = abi.TravStats()
        self._f("trace_shadow")(h, _ptr(a), _ptr(b), C.c_uint64(n), _ptr(vis), C.byref(st) if want_stats else None, nthreads)
        return (vis, st) if want_stats else vis

    # ---- camera / tasks / sampler
    def camera_init(self, pos, lookat, up, yview, xview, xres, yres, focus_plane=1.0, lens_size=0.0):
        cam = abi.Camera()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        fn = self._f("camera_init")
        fn.argtypes = [C.POINTER(abi.Camera), abi.f32p, abi.f32p, abi.f32p, C.c_float, C.c_float, C.c_int32, C.c_int32,
                       C.c_float, C.c_float]
        fn.restype = None
        fn(C.byref(cam), f3(pos), f3(lookat), f3(up), yview, xview, xres, yres, focus_plane, lens_size)
        return cam

    def camera_rays(self, cam, xres, yres, xy, offsets, lens=None):
        n = len(xy)
        rays = np.zeros(n, dtype=RAY_DT)
        xy = np.ascontiguousarray(xy, np.int32); offsets = np.ascontiguousarray(offsets, np.float32)
        if lens is not None:
            lens = np.ascontiguousarray(lens, np.float32)
        self._f("camera_rays")(C.byref(cam), C.c_uint32(xres), C.c_uint32(yres), _ptr(xy), _ptr(offsets), _ptr(lens),
                               C.c_uint64(n), _ptr(rays))
        return rays

    def generate_tasks(self, tile, xres, yres):
        n = self._f("generate_tasks")(tile, xres, yres, None, 0)
        out = (abi.Task * n)()
        self._f("generate_tasks")(tile, xres, yres, out, n)
        return out

    def sampler_set_size(self, ms):
        return int(self._f("sampler_set_size")(C.c_uint32(ms)))

    def sampler_tables(self, seeds, ms, n1d, n2d):
        seeds = np.ascontiguousarray(seeds, np.uint32)
        ss = self.sampler_set_size(ms)
        t1 = np.zeros((len(seeds), n1d, ss), np.float32)
        t2 = np.zeros((len(seeds), n2d, ss, 2), np.float32)
        self._f("sampler_tables")(_ptr(seeds), C.c_uint32(len(seeds)), C.c_uint32(ms), C.c_uint32(n1d), C.c_uint32(n2d),
                                  _ptr(t1), _ptr(t2))
        return t1, t2

    # ---- rendering
    def render_round(self, h, cam, params, tasks, seedstart=42, seedcount_base=0, fb=None, nthreads=8):
        npx = params.xres * params.yres
        if fb is None:
            fb = (np.zeros((params.yres, params.xres, 3), np.float32), np.zeros((params.yres, params.xres), np.uint32))
        st = abi.RoundStats()
        rc = self._f("render_round")(h, C.byref(cam), C.byref(params), tasks, C.c_uint32(len(tasks)), C.c_uint32(seedstart),
                                     C.c_uint32(seedcount_base), _ptr(fb[0]), _ptr(fb[1]), C.byref(st), nthreads)
        assert rc == 0, rc
        return fb[0], fb[1], st

    # ---- probes
    def bxdf_sample(self, h, material, Vi, uv, sample):
        n = len(Vi); out = np.zeros((n, 7), np.float32)
        Vi, uv, sample = (np.ascontiguousarray(x, np.float32) for x in (Vi, uv, sample))
        self._f("bxdf_sample")(h, C.c_uint32(material), _ptr(Vi), _ptr(uv), _ptr(sample), C.c_uint64(n), _ptr(out))
        return out

    def bxdf_value(self, h, material, Vi, Vr, uv):
        n = len(Vi); out = np.zeros((n, 3), np.float32)
        Vi, Vr, uv = (np.ascontiguousarray(x, np.float32) for x in (Vi, Vr, uv))
        self._f("bxdf_value")(h, C.c_uint32(material), _ptr(Vi), _ptr(Vr), _ptr(uv), C.c_uint64(n), _ptr(out))
        return out

    def texture_fetch(self, h, tex, uv):
        n = len(uv); out = np.zeros((n, 5), np.float32)
        uv = np.ascontiguousarray(uv, np.float32)
        self._f("texture_fetch")(h, C.c_uint32(tex), _ptr(uv), C.c_uint64(n), _ptr(out))
        return out

    def random_light(self, h, samples5):
        n = len(samples5); out = np.zeros((n, 12), np.float32)
        s = np.ascontiguousarray(samples5, np.float32)
        self._f("random_light")(h, _ptr(s), C.c_uint64(n), _ptr(out))
        return out

    def sky(self, h, dirs):
        n = len(dirs); out = np.zeros((n, 3), np.float32)
        d = np.ascontiguousarray(dirs, np.float32)
        self._f("sky")(h, _ptr(d), C.c_uint64(n), _ptr(out))
        return out


RAY_DT = np.dtype([("origin", np.float32, 3), ("direction", np.float32, 3), ("tnear", np.float32), ("tfar", np.float32)])
HIT_DT = np.dtype([("triangle", np.uint32), ("t", np.float32), ("a", np.float32), ("b", np.float32), ("c", np.float32)])
assert RAY_DT.itemsize == 32 and HIT_DT.itemsize == 20


def oracle():
    if not os.path.exists(ORACLE_SO):      # normally built by __graft_entry__.build(); one g++ call if it did not travel
        import subprocess
        r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(ORACLE_SO):
            raise RuntimeError(f"{ORACLE_SO} missing and `make -C oracle oracle` failed: {r.stderr[-400:]}")
    return Checker(ORACLE_SO, "rgko")


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    if not have_ref():
        raise RuntimeError(f"{REF_SO} missing (built only where /root/reference exists): run `make -C oracle ref`")
    return Checker(REF_SO, "rgkref")
