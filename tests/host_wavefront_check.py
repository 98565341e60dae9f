"""Larger one-off runs of the host-compiled wavefront (device sources + host loop of render.cu, wide-BVH path, device sampler)
against the oracle on the three stand-ins at 240x136x16 spp: ray counts and framebuffer bits.  ~1 minute of CPU.
  python tests/host_wavefront_check.py > profiles/r1_host_wavefront_vs_oracle.txt"""
import os, sys, time, numpy as np, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers, test_device_on_host as T
from rgk_b200 import standin
lib = C.CDLL(T.SO)
vp = C.c_void_p
lib.doh_shade_scene_create.restype = vp; lib.doh_shade_scene_create.argtypes = [vp, vp]; lib.doh_shade_scene_destroy.argtypes = [vp]
lib.doh_render_round.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
O = checkers.oracle()
for name in ("sponza", "sibenik", "conference"):
    pack, cfg = standin.BUILDERS[name](width=240, height=136, multisample=16)
    t = time.time()
    (rgb, cnt, st, bvh), (fo, co, so) = T._host_round(lib, O, pack, cfg, seedcount_base=11, wide_bvh=True, device_sampler=True)
    same = rgb.view(np.uint32) == fo.view(np.uint32)
    print(name, "depth", cfg.recursion_level, "paths", 240*136*16, "closest", int(st.closest_rays), int(so.closest_rays), "shadow", int(st.shadow_rays)+int(st.shadow_rays_skipped), int(so.shadow_rays),
          "bvh rays/deferred", bvh, "framebuffer words differing", int((~same).sum()), "of", same.size, "counts equal", bool(np.array_equal(cnt, co)), "%.1f s" % (time.time() - t), flush=True)
