#!/usr/bin/env python
"""Driver of tools/tsan_lanes_on_host.sh: whole rounds of the wavefront through the 32-lane host build (threads as lanes) named by
$DOH_MT_SO -- the ThreadSanitizer-instrumented libdevice_on_host_mt_tsan.so, run with libtsan preloaded -- compared with the oracle."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers
from rgk_b200 import scenes, standin
import test_wavefront_lanes_on_host as W
from test_device_on_host import _host_round, vp
lib = C.CDLL(os.environ.get("DOH_MT_SO", W.SO))
lib.doh_shade_scene_create.restype = vp; lib.doh_shade_scene_create.argtypes = [vp, vp]; lib.doh_shade_scene_destroy.argtypes = [vp]
lib.doh_render_round.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
O = checkers.oracle()
ok = True
for name, (pack, cfg), bvh in (("material zoo, wide BVH + kd arbiter", scenes.material_zoo(width=40, height=24, multisample=4, recursion_max=3, lens=0.04), True),
                               ("Cornell box, recursion-max 8, kd-tree", scenes.load_builtin("cornell-box", width=24, height=24, multisample=4, recursion_max=8), False),
                               ("atrium stand-in, wide BVH", standin.sponza(width=48, height=27, multisample=4), True)):
    (rgb, cnt, st, _), (fo, co, so) = _host_round(lib, O, pack, cfg, seedcount_base=3, wide_bvh=bvh, device_sampler=True, **W.DEFAULTS)
    same = np.array_equal(rgb.view(np.uint32), fo.view(np.uint32)) and np.array_equal(cnt, co)
    print("round:", name, "-- framebuffer identical to the oracle's:", same, flush=True)
    ok &= same
sys.exit(0 if ok else 1)
