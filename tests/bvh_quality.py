#!/usr/bin/env python
"""Work per ray of the PRODUCT-built wide BVH (rgk_b200/csrc/host_bvh.cpp), counted by the oracle's CPU mirror of the device
traversal: wide nodes visited and exact tests per primary / bounce / shadow ray on the stand-ins.  The builder is tuned
against these counters (the kernel does not change with the tree).   python tests/bvh_quality.py [--scenes sponza,conference]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers, raybatches
from rgk_b200 import device, standin
import test_bvh_host as T

ap = argparse.ArgumentParser(); ap.add_argument("--scenes", default="sponza,conference"); ap.add_argument("--res", default="320x180")
args = ap.parse_args()
w, hgt = (int(x) for x in args.res.split("x"))
tot = {"nodes": 0.0, "tests": 0.0, "rays": 0}
for scene in args.scenes.split(","):
    pack, cfg = standin.BUILDERS[scene](width=w, height=hgt, multisample=1)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    O = checkers.oracle(); h = O.scene_create(pack.desc())
    ca = cfg.camera_args()
    cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    rays = raybatches.primary(O, cam, w, hgt, jitter_seed=11)
    want = O.trace_closest(h, rays)
    nodes, order, depth = hs.bvh()
    closest, shadow = T._mirror(O, h, nodes, order)
    eps = O.scene_info(h).epsilon
    brays, ign = raybatches.bounce(rays, want, O.scene_planes(h)[:, :3], eps, seed=100)
    wantb = O.trace_closest(h, brays, ign)
    light = np.asarray(pack.point_lights[0][0], np.float32) if pack.point_lights else np.array([0, 5, 0], np.float32)
    a, b = raybatches.shadow_segments(brays, wantb, light)
    row = {"scene": scene, "wide_nodes": len(nodes), "depth": depth}
    for name, fn in (("primary", lambda: closest(rays)), ("bounce", lambda: closest(brays, ign)), ("shadow", lambda: shadow(a, b))):
        got, df, cnt = fn()
        n = len(df)
        row[name] = {"nodes_per_ray": round(float(cnt[0]) / n, 3), "tests_per_ray": round(float(cnt[1]) / n, 3), "deferred": round(float(df.mean()), 5)}
        tot["nodes"] += float(cnt[0]); tot["tests"] += float(cnt[1]); tot["rays"] += n
    print(json.dumps(row))
print(json.dumps({"all": {"nodes_per_ray": round(tot["nodes"] / tot["rays"], 3), "tests_per_ray": round(tot["tests"] / tot["rays"], 3),
                          "cost (nodes + 0.4 tests)": round((tot["nodes"] + 0.4 * tot["tests"]) / tot["rays"], 3)}}))
