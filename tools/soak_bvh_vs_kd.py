#!/usr/bin/env python
"""Soak of the default traversal (wide BVH + kd arbiter) against the kd-only one: renders rounds R, R+1, ... of a workload on both
(different seeds every round) and compares framebuffers and ray counts bit for bit.  A differing round is reported with its pixels;
tools/find_bvh_mismatch.py <workload> --round <that round> then isolates its tiles for tools/repro_tile_on_host.py.
   python tools/soak_bvh_vs_kd.py <workload> [--spp N] [--first R] [--rounds N] [--seconds S]  -> one JSON line"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rgk_b200 import abi, device, multigpu

ap = argparse.ArgumentParser()
ap.add_argument("workload"); ap.add_argument("--spp", type=int, default=None); ap.add_argument("--first", type=int, default=200)
ap.add_argument("--rounds", type=int, default=10); ap.add_argument("--seconds", type=float, default=1e9)
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
img = {t: torch.zeros((p.yres, p.xres, 3), dtype=torch.float32, device="cuda") for t in ctxs}
cnt = {t: torch.zeros((p.yres, p.xres), dtype=torch.int32, device="cuda") for t in ctxs}
t0 = time.time(); rays = 0; deferred = 0; bad = []; done = 0
for r in range(a.first, a.first + a.rounds):
    if time.time() - t0 > a.seconds:
        break
    st = {}
    for t, c in ctxs.items():
        img[t].zero_(); cnt[t].zero_()
        st[t] = c.render_round_device(cam, p, tasks, img[t].data_ptr(), cnt[t].data_ptr(), 42, multigpu.seedcount_base(r, len(tasks)))
    torch.cuda.synchronize()
    diff = (img["bvh"].view(torch.int32) != img["kd"].view(torch.int32)).any(dim=2) | (cnt["bvh"] != cnt["kd"])
    counts = {t: [int(st[t].closest_rays), int(st[t].shadow_rays), int(st[t].shadow_rays_skipped)] for t in ctxs}
    rays += counts["bvh"][0] + counts["bvh"][1]; done += 1
    if bool(diff.any()) or counts["bvh"] != counts["kd"]:
        ys, xs = torch.nonzero(diff, as_tuple=True)
        bad.append({"round": r, "pixels": [(int(x), int(y)) for x, y in zip(xs.tolist(), ys.tolist())][:32], "rays": counts})
print(json.dumps({"workload": label, "spp_override": a.spp, "rounds": [a.first, a.first + done - 1], "rays_per_traversal": rays, "rounds_differing": bad, "seconds": round(time.time() - t0, 1)}), flush=True)
