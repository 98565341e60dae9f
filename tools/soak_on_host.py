#!/usr/bin/env python
"""CPU twin of tools/soak_bvh_vs_kd.py: rounds of a (small) workload on the host build of the device sources (tests/host_cpp, one-lane
warps), wide-BVH traversal against kd-only, framebuffers and ray counts compared bit for bit.  A differing round is written in the
format tools/repro_tile_on_host.py reads (tile, seedcount_base of the single-tile render).
   python tools/soak_on_host.py <workload> <first round> <rounds> [out.json]"""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench, checkers
from rgk_b200 import abi
vp = C.c_void_p
workload, first, rounds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
pack, cfg, label = bench.build_workload(workload, None, None)
desc = pack.desc()
doh = C.CDLL(os.environ.get("DOH_SO", os.path.join(ROOT, "build", "host", "libdevice_on_host.so")))
doh.doh_shade_scene_create.restype = vp; doh.doh_shade_scene_create.argtypes = [vp, vp]; doh.doh_shade_scene_destroy.argtypes = [vp]
doh.doh_render_round.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
O = checkers.oracle()
ca = cfg.camera_args()
cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
p = cfg.params(abi.SAMPLER_MT19937)
lib = abi.load_library()
nt = lib.rgk_generate_tasks(32, p.xres, p.yres, None, 0)
tasks = (abi.Task * nt)()
lib.rgk_generate_tasks(32, p.xres, p.yres, tasks, nt)                      # RenderDriver's task order
tiles = [(t.x1, t.x2, t.y1, t.y2) for t in tasks]
H = {}
for trav in ("bvh", "kd"):
    dcfg = abi.device_cfg(traversal=trav, binning=0, refill_coherent=1, refill_incoherent=1, refill_shadow=1, sampler_kernel=2)
    H[trav] = (dcfg, vp(doh.doh_shade_scene_create(C.byref(desc), C.byref(dcfg))))
out = {"workload": label, "tiles": []}
for r in range(first, first + rounds):
    base = r * len(tiles)
    img, rays = {}, {}
    for trav, (dcfg, h) in H.items():
        rgb = np.zeros((p.yres, p.xres, 3), np.float32); cnt = np.zeros((p.yres, p.xres), np.uint32); st = abi.RoundStats(); bv = np.zeros(2, np.uint64)
        rc = doh.doh_render_round(h, C.byref(dcfg), C.byref(cam), C.byref(p), tasks, len(tiles), 42, base, None, None, 0, 0, 0, rgb.ctypes.data, cnt.ctypes.data, C.byref(st), bv.ctypes.data)
        assert rc == 0
        img[trav] = rgb; rays[trav] = [int(st.closest_rays), int(st.shadow_rays)]
    d = np.argwhere((img["bvh"].view(np.uint32) != img["kd"].view(np.uint32)).any(axis=2))
    print("round", r, "differing pixels", [(int(c), int(q)) for q, c in d], rays, flush=True)
    for q, c in d:
        ti = next(i for i, (x1, x2, y1, y2) in enumerate(tiles) if x1 <= c < x2 and y1 <= q < y2)
        out["tiles"].append({"round": r, "tile_index": ti, "tile": list(tiles[ti]), "seedcount_base_for_single_tile": base + ti, "pixels_differing": [[int(c), int(q)]]})
    if rays["bvh"] != rays["kd"] and len(d) == 0:
        out["tiles"].append({"round": r, "rays_only": rays})
if len(sys.argv) > 4:
    json.dump(out, open(sys.argv[4], "w"), indent=1)
