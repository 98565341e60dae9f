#!/usr/bin/env python
"""Where does a round on the default traversal (wide BVH + kd arbiter) differ from the kd-only round?  Renders one round of a
workload on both, lists the differing pixels and their 32x32 tiles, re-renders each such tile alone on both traversals (a tile's
seed depends only on its index, so the single-tile render repeats its part of the frame) and saves what a CPU reproduction needs.
   python tools/find_bvh_mismatch.py <workload> [--spp N] [--round R] -> gpurun_out/mismatch_<workload>.json"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from rgk_b200 import abi, device, multigpu

ap = argparse.ArgumentParser()
ap.add_argument("workload"); ap.add_argument("--spp", type=int, default=None); ap.add_argument("--round", type=int, default=100)
ap.add_argument("--max-tiles", type=int, default=6)
a = ap.parse_args()
pack, cfg, label = bench.build_workload(a.workload, a.spp, None)
desc = pack.desc()
stream = torch.cuda.current_stream().cuda_stream
ctxs = {t: device.Context(0, stream=stream, traversal=t) for t in ("bvh", "kd")}
for c in ctxs.values():
    c.commit(desc)
cam = ctxs["bvh"].camera(**cfg.camera_args())
p = cfg.params(abi.SAMPLER_MT19937)
tasks = ctxs["bvh"].generate_tasks(32, p.xres, p.yres)
nt = len(tasks)
base = multigpu.seedcount_base(a.round, nt)
img, cnt, st = {}, {}, {}
for t, c in ctxs.items():
    img[t] = torch.zeros((p.yres, p.xres, 3), dtype=torch.float32, device="cuda"); cnt[t] = torch.zeros((p.yres, p.xres), dtype=torch.int32, device="cuda")
    st[t] = c.render_round_device(cam, p, tasks, img[t].data_ptr(), cnt[t].data_ptr(), 42, base)
torch.cuda.synchronize()
diff = (img["bvh"].view(torch.int32) != img["kd"].view(torch.int32)).any(dim=2)
ys, xs = torch.nonzero(diff, as_tuple=True)
pix = [(int(x), int(y)) for x, y in zip(xs.tolist(), ys.tolist())]
out = {"workload": label, "round": a.round, "seedcount_base": base, "pixels_differing": pix,
       "rays": {t: [int(st[t].closest_rays), int(st[t].shadow_rays), int(st[t].shadow_rays_skipped)] for t in ctxs}, "tiles": []}
print(label, "differing pixels:", pix, out["rays"], flush=True)
seen = set()
for (x, y) in pix:
    ti = next(i for i in range(nt) if tasks[i].x1 <= x < tasks[i].x2 and tasks[i].y1 <= y < tasks[i].y2)
    if ti in seen or len(seen) >= a.max_tiles:
        continue
    seen.add(ti)
    one = (abi.Task * 1)(tasks[ti])
    t_img = {}
    for t, c in ctxs.items():
        f = torch.zeros((p.yres, p.xres, 3), dtype=torch.float32, device="cuda"); k = torch.zeros((p.yres, p.xres), dtype=torch.int32, device="cuda")
        s = c.render_round_device(cam, p, one, f.data_ptr(), k.data_ptr(), 42, base + ti)
        torch.cuda.synchronize()
        t_img[t] = (f, int(s.closest_rays), int(s.shadow_rays))
    same_as_frame = bool((t_img["bvh"][0][tasks[ti].y1:tasks[ti].y2, tasks[ti].x1:tasks[ti].x2] == img["bvh"][tasks[ti].y1:tasks[ti].y2, tasks[ti].x1:tasks[ti].x2]).all())
    d = (t_img["bvh"][0].view(torch.int32) != t_img["kd"][0].view(torch.int32)).any(dim=2)
    dy, dx = torch.nonzero(d, as_tuple=True)
    rec = {"tile_index": ti, "tile": [tasks[ti].x1, tasks[ti].x2, tasks[ti].y1, tasks[ti].y2], "seedcount_base_for_single_tile": base + ti,
           "single_tile_repeats_the_frame": same_as_frame, "pixels_differing": [(int(u), int(v)) for u, v in zip(dx.tolist(), dy.tolist())],
           "rays_bvh": t_img["bvh"][1:], "rays_kd": t_img["kd"][1:],
           "values": {t: [t_img[t][0][v, u].tolist() for u, v in zip(dx.tolist(), dy.tolist())] for t in ctxs}}
    print(json.dumps(rec), flush=True)
    out["tiles"].append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/mismatch_%s.json" % a.workload, "w"), indent=1)
