"""Traversal-only benchmark (SURVEY 8d "ray batches"): primary, bounce and shadow batches on a stand-in scene,
device-resident, timed with CUDA events on the launch stream.  Prints one JSON line per batch with Mrays/s,
the oracle-defined algorithmic bytes per ray and the achieved GB/s.

  python tools/trace_bench.py [--scene sponza] [--res 1920x1080] [--reps 5] [--check N]

--check N compares the first N rays of every batch against the CPU oracle (bit-exact).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="sponza")
    ap.add_argument("--res", default="1920x1080")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--copies", type=int, default=2, help="jittered copies of the primary batch (>= 2^22 rays)")
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--traversal", default="kd", choices=["kd", "bvh"])
    ap.add_argument("--kd-variant", dest="kd_variant", type=int, default=6, choices=[2, 6])
    ap.add_argument("--sort", type=int, default=0, help="sort secondary batches by (direction octant, origin cell on an N^3 grid)")
    args = ap.parse_args()
    import torch
    from rgk_b200 import device, standin, abi
    w, h = (int(x) for x in args.res.split("x"))
    pack, cfg = standin.BUILDERS[args.scene](width=w, height=h, multisample=1)
    desc = pack.desc()
    stream = torch.cuda.current_stream().cuda_stream
    ctx = device.Context(0, stream=stream, traversal=args.traversal, kd_variant=args.kd_variant)
    ctx.commit(desc)
    info = ctx.scene_info()
    cam = ctx.camera(**cfg.camera_args())
    ys, xs = np.mgrid[0:h, 0:w]
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    rng = np.random.default_rng(1)
    rays = np.concatenate([ctx.camera_rays(cam, w, h, xy, rng.random((len(xy), 2), dtype=np.float32)) for _ in range(args.copies)])
    hits = ctx.trace_closest(rays)
    import raybatches
    planes = np.frombuffer(ctx.host_planes(), np.float32).reshape(-1, 4) if hasattr(ctx, "host_planes") else None
    if planes is None:
        # geometric normals from the scene pack
        a = pack.arrays()
        P, I = a["positions"], a["indices"]
        n = np.cross(P[I[:, 2]] - P[I[:, 0]], P[I[:, 1]] - P[I[:, 0]])
        planes = n / np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    brays, ign = raybatches.bounce(rays, hits, planes[:, :3], info.epsilon)
    light = pack.point_lights[0][0] if pack.point_lights else (0.0, 10.0, 0.0)
    sa, sb = raybatches.shadow_segments(rays, hits, light)
    if args.sort:
        bb = np.array(list(info.bbox), np.float32).reshape(3, 2)
        def key(o, d):
            g = np.clip(((o - bb[:, 0]) / (bb[:, 1] - bb[:, 0]) * args.sort).astype(np.int64), 0, args.sort - 1)
            octant = (d[:, 0] > 0).astype(np.int64) | ((d[:, 1] > 0).astype(np.int64) << 1) | ((d[:, 2] > 0).astype(np.int64) << 2)
            return (octant * args.sort ** 3) + (g[:, 0] * args.sort + g[:, 1]) * args.sort + g[:, 2]
        o = np.argsort(key(brays["origin"], brays["direction"]), kind="stable")
        brays, ign = brays[o], ign[o]
    bh = ctx.trace_closest(brays, ign)
    sa2, sb2 = raybatches.shadow_segments(brays, bh, light)
    if args.sort:
        d = sb2 - sa2
        o = np.argsort(key(sb2, -d), kind="stable")
        sa2, sb2 = sa2[o], sb2[o]

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()

    out = []
    batches = [("primary", rays, None), ("bounce", brays, ign)]
    for name, r, ig in batches:
        d_r, d_i = dev(r), (dev(ig) if ig is not None else None)
        d_h = torch.empty(len(r) * 20, dtype=torch.uint8, device="cuda")
        _, st = ctx.trace_closest(r[: min(len(r), 1 << 20)], ig[: min(len(r), 1 << 20)] if ig is not None else None, want_stats=True)
        for _ in range(3):
            ctx.trace_closest_device(d_r.data_ptr(), d_i.data_ptr() if d_i is not None else None, len(r), d_h.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(args.reps):
            ctx.trace_closest_device(d_r.data_ptr(), d_i.data_ptr() if d_i is not None else None, len(r), d_h.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        bpr = st.bytes_per_ray(20)
        rec = {"batch": name, "rays": len(r), "ms": ms, "Mrays_s": len(r) / ms / 1e3, "bytes_per_ray": bpr,
               "GB_s": len(r) * bpr / ms / 1e6, "per_ray": {k: v / st.rays for k, v in st.as_dict().items() if k != "rays"}}
        if args.check:
            import checkers
            O = checkers.oracle(); ho = O.scene_create(desc)
            got = np.frombuffer(d_h.cpu().numpy().tobytes(), dtype=device.HIT_DT)[: args.check]
            ref = O.trace_closest(ho, r[: args.check], ig[: args.check] if ig is not None else None)
            rec["bit_exact_vs_oracle"] = bool(got.tobytes() == ref.tobytes())
        out.append(rec)
    for name, a, b in (("shadow", sa, sb), ("shadow2", sa2, sb2)):
        d_a, d_b = dev(a), dev(b)
        d_v = torch.empty(len(a), dtype=torch.uint8, device="cuda")
        _, st = ctx.trace_shadow(a[: 1 << 20], b[: 1 << 20], want_stats=True)
        for _ in range(3):
            ctx.trace_shadow_device(d_a.data_ptr(), d_b.data_ptr(), len(a), d_v.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(args.reps):
            ctx.trace_shadow_device(d_a.data_ptr(), d_b.data_ptr(), len(a), d_v.data_ptr())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        bpr = st.bytes_per_ray(1)
        rec = {"batch": name, "rays": len(a), "ms": ms, "Mrays_s": len(a) / ms / 1e3, "bytes_per_ray": bpr, "GB_s": len(a) * bpr / ms / 1e6,
               "visible_frac": float(d_v.float().mean()), "per_ray": {k: v / st.rays for k, v in st.as_dict().items() if k != "rays"}}
        if args.check:
            import checkers
            O = checkers.oracle(); ho = O.scene_create(desc)
            rec["bit_exact_vs_oracle"] = bool(np.array_equal(d_v.cpu().numpy()[: args.check], O.trace_shadow(ho, a[: args.check], b[: args.check])))
        out.append(rec)
    tag = {"scene": args.scene, "triangles": info.n_triangles, "variant": args.kd_variant}
    for r in out:
        r.update(tag)
        print(json.dumps(r))


if __name__ == "__main__":
    main()
