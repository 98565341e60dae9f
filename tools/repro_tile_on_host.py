#!/usr/bin/env python
"""CPU reproduction of one tile of a round: the host build of the device sources (tests/host_cpp, one-lane warps) on the wide-BVH
traversal and on the kd-only one, and the oracle, for the tile tools/find_bvh_mismatch.py reported.
   python tools/repro_tile_on_host.py gpurun_out/mismatch_<workload>.json <workload> [tile number in the file]
Following one path: DOH_DUMP_PIXEL=x,y prints the pixel's per-sample sums on both traversals (diff them for the sample's slot), then
DOH_SLOT=<slot> with a -DRGK_DOH_DEBUG build of tests/host_cpp/device_on_host.cpp (DOH_SO=<that .so>; same g++ line as
__graft_entry__.build()'s libdevice_on_host.so) prints every closest-hit and shadow query of that slot."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, checkers
from rgk_b200 import abi
vp = C.c_void_p
info = json.load(open(sys.argv[1])); workload = sys.argv[2]; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rec = info["tiles"][which]
pack, cfg, label = bench.build_workload(workload, None, None)
desc = pack.desc()
doh = C.CDLL(os.environ.get("DOH_SO", os.path.join(ROOT, "build", "host", "libdevice_on_host.so")))
doh.doh_shade_scene_create.restype = vp; doh.doh_shade_scene_create.argtypes = [vp, vp]; doh.doh_shade_scene_destroy.argtypes = [vp]
doh.doh_render_round.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
O = checkers.oracle()
ca = cfg.camera_args()
cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
p = cfg.params(abi.SAMPLER_MT19937)
x1, x2, y1, y2 = rec["tile"]
one = (abi.Task * 1)(); one[0].x1, one[0].x2, one[0].y1, one[0].y2 = x1, x2, y1, y2
for f in ("xm", "ym"):
    if hasattr(one[0], f): setattr(one[0], f, 0.0)
base = rec["seedcount_base_for_single_tile"]
imgs = {}
for trav in ("bvh", "kd"):
    dcfg = abi.device_cfg(traversal=trav, binning=0, refill_coherent=1, refill_incoherent=1, refill_shadow=1, sampler_kernel=2)
    t = time.time()
    h = vp(doh.doh_shade_scene_create(C.byref(desc), C.byref(dcfg)))
    rgb = np.zeros((p.yres, p.xres, 3), np.float32); cnt = np.zeros((p.yres, p.xres), np.uint32); st = abi.RoundStats(); bv = np.zeros(2, np.uint64)
    rc = doh.doh_render_round(h, C.byref(dcfg), C.byref(cam), C.byref(p), one, 1, 42, base, None, None, 0, 0, 0, rgb.ctypes.data, cnt.ctypes.data, C.byref(st), bv.ctypes.data)
    doh.doh_shade_scene_destroy(h)
    imgs[trav] = rgb[y1:y2, x1:x2].copy()
    print(trav, "rc", rc, "closest", int(st.closest_rays), "shadow", int(st.shadow_rays), "deferred", int(bv[1]), "sec", round(time.time() - t, 1), flush=True)
d = np.argwhere((imgs["bvh"].view(np.uint32) != imgs["kd"].view(np.uint32)).any(axis=2))
print("host bvh vs host kd: differing pixels (x, y):", [(int(x1 + c), int(y1 + r)) for r, c in d], "| GPU reported:", rec["pixels_differing"])
for r, c in d:
    print("  ", (x1 + c, y1 + r), "bvh", imgs["bvh"][r, c], "kd", imgs["kd"][r, c])
