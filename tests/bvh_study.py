#!/usr/bin/env python
"""Design study for DESIGN.md "Next": work per ray of a wide BVH (binned SAH, collapsed to 4 / 8 children) against the
reference kd-tree on the stand-in scene, and how often its closest hit lands on a different triangle.  CPU only (the
oracle library); nothing here is a product path or a test gate.   python tests/bvh_study.py [--scene sponza] [--rays 200000]"""
import argparse, ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers, raybatches
from rgk_b200 import standin

ap = argparse.ArgumentParser(); ap.add_argument("--scene", default="sponza"); ap.add_argument("--rays", type=int, default=200000)
args = ap.parse_args()
pack, cfg = standin.BUILDERS[args.scene](width=640, height=360, multisample=1)
O = checkers.oracle()
h = O.scene_create(pack.desc())
ca = cfg.camera_args()
cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
rays = raybatches.primary(O, cam, 640, 360, jitter_seed=3)
hits, st = O.trace_closest(h, rays, want_stats=True)
brays, ign = raybatches.bounce(rays, hits, O.scene_planes(h)[:, :3], O.scene_info(h).epsilon)
_, sb = O.trace_closest(h, brays, ign, want_stats=True)
O.lib.rgko_bvh_study.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(1)
for name, r, g, kd in (("primary", rays, None, st), ("bounce", brays, ign, sb)):
    sel = rng.choice(len(r), min(args.rays, len(r)), replace=False)
    rr = np.ascontiguousarray(r[sel]); gg = None if g is None else np.ascontiguousarray(g[sel])
    print(json.dumps({"batch": name, "kd_tree": {"inner_per_ray": kd.inner / kd.rays, "leaves_per_ray": kd.leaf / kd.rays, "tests_per_ray": kd.tests / kd.rays}}))
    for width, leaf in ((4, 4), (8, 4), (8, 2)):
        out = np.zeros(10)
        O.lib.rgko_bvh_study(h, rr.ctypes.data_as(C.c_void_p), None if gg is None else gg.ctypes.data_as(C.c_void_p), C.c_uint64(len(rr)), width, leaf, out.ctypes.data_as(C.c_void_p))
        n = out[0]
        print(json.dumps({"batch": name, "bvh_width": width, "leaf_size": leaf, "nodes_per_ray": out[1] / n, "boxes_per_ray": out[2] / n, "tests_per_ray": out[3] / n,
                          "hit_id_differs": out[4] / n, "of_which_within_2eps": out[5] / max(out[4], 1), "ambiguous": out[8] / n, "differs_unflagged": int(out[9]), "wide_nodes": int(out[6]), "node_MB": out[7] / 1e6}))
