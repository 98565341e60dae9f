#!/usr/bin/env python
"""A/B of the wide-BVH traversal (RGK_TRAVERSAL_BVH, the default) against the kd-tree kernels on device-resident ray batches of a
stand-in scene: Mrays/s of both (CUDA events on the launch stream), the deferred fraction, and a FULL-SIZE bit-exact
comparison of every hit record / visibility flag between the two (the kd kernels are the ones the parity suite pins to
the oracle).  torch-free (ctypes on libcudart) so that it starts in seconds on a fresh box.

  python tools/bvh_ab.py [--scene sponza] [--res 1920x1080] [--reps 5]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


class Cuda:
    def __init__(self):
        self.rt = C.CDLL("libcudart.so")
        self.rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    def ck(self, e):
        if e != 0: raise RuntimeError(f"cuda error {e}")
    def stream(self):
        s = C.c_void_p(); self.ck(self.rt.cudaStreamCreate(C.byref(s))); return s
    def event(self):
        e = C.c_void_p(); self.ck(self.rt.cudaEventCreate(C.byref(e))); return e
    def to_device(self, a):
        a = np.ascontiguousarray(a); p = C.c_void_p()
        self.ck(self.rt.cudaMalloc(C.byref(p), C.c_size_t(max(a.nbytes, 256))))
        self.ck(self.rt.cudaMemcpy(p, C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes), 1))
        return p.value
    def empty(self, nbytes):
        p = C.c_void_p(); self.ck(self.rt.cudaMalloc(C.byref(p), C.c_size_t(max(nbytes, 256)))); return p.value
    def to_host(self, p, dtype, n):
        a = np.zeros(n, dtype); self.ck(self.rt.cudaMemcpy(C.c_void_p(a.ctypes.data), C.c_void_p(p), C.c_size_t(a.nbytes), 2)); return a
    def free(self, p): self.rt.cudaFree(C.c_void_p(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="sponza"); ap.add_argument("--res", default="1920x1080"); ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--render", type=int, default=2, help="also time N full render rounds of the bench workload (default scene config) on both contexts")
    ap.add_argument("--spp", type=int, default=None, help="override the samples per pixel of the render rounds")
    ap.add_argument("--render-only", action="store_true", help="skip the ray-batch A/B (for ncu runs)")
    ap.add_argument("--sweep", action="store_true", help="render-only: time the BVH context under a list of rgk_device_cfg settings")
    args = ap.parse_args()
    from rgk_b200 import device, standin, abi
    import raybatches
    cu = Cuda()
    w, h = (int(x) for x in args.res.split("x"))
    pack, cfg = standin.BUILDERS[args.scene](width=w, height=h, multisample=1)
    desc = pack.desc()
    st = cu.stream()
    kd = device.Context(0, stream=st.value, traversal="kd"); kd.commit(desc)
    t0 = time.time(); bv = device.Context(0, stream=st.value, traversal="bvh"); bv.commit(desc); t_commit = time.time() - t0
    if args.sweep:
        return sweep(cu, bv, args.scene, args.render, args.spp)
    if args.render_only:
        render_ab(cu, st, kd, bv, args.scene, args.render, args.spp)
        return
    info = kd.scene_info()
    cam = kd.camera(**cfg.camera_args())
    ys, xs = np.mgrid[0:h, 0:w]
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    rays = kd.camera_rays(cam, w, h, xy, np.random.default_rng(1).random((len(xy), 2), dtype=np.float32))
    hits = kd.trace_closest(rays)
    a = pack.arrays(); P, I = a["positions"], a["indices"]
    nrm = np.cross(P[I[:, 2]] - P[I[:, 0]], P[I[:, 1]] - P[I[:, 0]]); nrm = nrm / np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    brays, ign = raybatches.bounce(rays, hits, nrm, info.epsilon)
    bh = kd.trace_closest(brays, ign)
    light = pack.point_lights[0][0] if pack.point_lights else (0.0, 10.0, 0.0)
    sa, sb = raybatches.shadow_segments(brays, bh, light)
    e0, e1 = cu.event(), cu.event()

    def timed(fn):
        for _ in range(2): fn()
        cu.ck(cu.rt.cudaEventRecord(e0, st))
        for _ in range(args.reps): fn()
        cu.ck(cu.rt.cudaEventRecord(e1, st)); cu.ck(cu.rt.cudaEventSynchronize(e1))
        ms = C.c_float(); cu.ck(cu.rt.cudaEventElapsedTime(C.byref(ms), e0, e1))
        return ms.value / args.reps

    for name, r, ig in (("primary", rays, None), ("bounce", brays, ign)):
        d_r, d_i = cu.to_device(r), (cu.to_device(ig) if ig is not None else None)
        d_h1, d_h2 = cu.empty(len(r) * 20), cu.empty(len(r) * 20)
        ms_kd = timed(lambda: kd.trace_closest_device(d_r, d_i, len(r), d_h1))
        bv.bvh_stats()
        ms_bv = timed(lambda: bv.trace_closest_device(d_r, d_i, len(r), d_h2))
        s = bv.bvh_stats()
        h1, h2 = cu.to_host(d_h1, np.uint32, len(r) * 5), cu.to_host(d_h2, np.uint32, len(r) * 5)
        print(json.dumps({"batch": name, "rays": len(r), "kd_ms": ms_kd, "bvh_ms": ms_bv, "kd_mrays_s": len(r) / ms_kd / 1e3, "bvh_mrays_s": len(r) / ms_bv / 1e3,
                          "speedup": ms_kd / ms_bv, "deferred_frac": s["ambiguous"] / max(1, s["rays"]), "bvh_rays_counted": s["rays"],
                          "mismatching_records": int((h1.reshape(-1, 5) != h2.reshape(-1, 5)).any(1).sum())}), flush=True)
        for p in (d_r, d_i, d_h1, d_h2):
            if p is not None: cu.free(p)
    d_a, d_b = cu.to_device(sa), cu.to_device(sb)
    d_v1, d_v2 = cu.empty(len(sa)), cu.empty(len(sa))
    ms_kd = timed(lambda: kd.trace_shadow_device(d_a, d_b, len(sa), d_v1))
    bv.bvh_stats()
    ms_bv = timed(lambda: bv.trace_shadow_device(d_a, d_b, len(sa), d_v2))
    s = bv.bvh_stats()
    v1, v2 = cu.to_host(d_v1, np.uint8, len(sa)), cu.to_host(d_v2, np.uint8, len(sa))
    print(json.dumps({"batch": "shadow", "rays": len(sa), "kd_ms": ms_kd, "bvh_ms": ms_bv, "kd_mrays_s": len(sa) / ms_kd / 1e3, "bvh_mrays_s": len(sa) / ms_bv / 1e3,
                      "speedup": ms_kd / ms_bv, "deferred_frac": s["ambiguous"] / max(1, s["rays"]), "visible_frac": float(v1.mean()),
                      "mismatching_records": int((v1 != v2).sum()), "bvh_commit_s": t_commit}), flush=True)
    for p in (d_a, d_b, d_v1, d_v2): cu.free(p)
    if args.render > 0:
        render_ab(cu, st, kd, bv, args.scene, args.render, args.spp)


SWEEP = [{}, {"arb_grid": 1}, {"arb_grid": 4}, {"refill_incoherent": 16}, {"refill_incoherent": 28}, {"refill_incoherent": 32},
         {"refill_coherent": 32}, {"refill_shadow": 6}, {"refill_shadow": 20}, {"binning": 0}, {"bin_shadow_first": 0}, {"shade_path_order": 0},
         {"bvh_shadow_nosort": 1}, {"bvh_closest_nearest": 1}, {}]


def sweep(cu, bv, scene, rounds, spp=None):
    """ms per round of the BVH context under each rgk_device_cfg setting (scheduling fields apply to the next call)."""
    from rgk_b200 import standin
    pack, cfg = standin.BUILDERS[scene](**({"multisample": spp} if spp else {}))
    params = cfg.params()
    cam = bv.camera(**cfg.camera_args())
    tasks = bv.generate_tasks(64, params.xres, params.yres)
    npx = params.xres * params.yres
    d_rgb, d_cnt = cu.empty(npx * 12), cu.empty(npx * 4)
    ref = None
    from rgk_b200 import abi
    base = bv.cfg()
    for knobs in SWEEP:
        bv.configure(abi.DeviceCfg.from_buffer_copy(bytes(base)), **knobs)
        ms = []
        for r in range(rounds + 1):
            cu.ck(cu.rt.cudaMemset(C.c_void_p(d_rgb), 0, C.c_size_t(npx * 12))); cu.ck(cu.rt.cudaMemset(C.c_void_p(d_cnt), 0, C.c_size_t(npx * 4)))
            stats = bv.render_round_device(cam, params, tasks, d_rgb, d_cnt)
            bv.synchronize()
            ms.append((stats.gpu_ms, stats.closest_ms, stats.shadow_ms, stats.shade_ms, stats.sampler_ms))
        fb = cu.to_host(d_rgb, np.uint32, npx * 3)
        if ref is None: ref = fb
        best = min(ms[1:])
        print(json.dumps({"knobs": knobs, "ms": best[0], "closest_ms": best[1], "shadow_ms": best[2], "shade_ms": best[3], "sampler_ms": best[4],
                          "same_framebuffer": bool((fb == ref).all())}), flush=True)


def render_ab(cu, st, kd, bv, scene, rounds, spp=None):
    """ms per RenderDriver round of the bench workload (the stand-in's own resolution / spp / depth) through the kd-only and
    the BVH context, the framebuffers compared bit for bit."""
    from rgk_b200 import standin
    pack, cfg = standin.BUILDERS[scene](**({"multisample": spp} if spp else {}))
    params = cfg.params()
    out = {}
    for name, ctx in (("kd", kd), ("bvh", bv)):
        cam = ctx.camera(**cfg.camera_args())
        tasks = ctx.generate_tasks(64, params.xres, params.yres)
        npx = params.xres * params.yres
        d_rgb, d_cnt = cu.empty(npx * 12), cu.empty(npx * 4)
        ms = []
        for r in range(rounds + 1):
            cu.ck(cu.rt.cudaMemset(C.c_void_p(d_rgb), 0, C.c_size_t(npx * 12))); cu.ck(cu.rt.cudaMemset(C.c_void_p(d_cnt), 0, C.c_size_t(npx * 4)))
            ctx.bvh_stats()
            stats = ctx.render_round_device(cam, params, tasks, d_rgb, d_cnt)
            ctx.synchronize()
            ms.append(stats.gpu_ms)
        b = ctx.bvh_stats()
        out[name] = {"ms": ms[1:], "closest_rays": stats.closest_rays, "shadow_rays": stats.shadow_rays, "bvh_rays": b["rays"], "bvh_deferred": b["ambiguous"],
                     "fb": cu.to_host(d_rgb, np.uint32, npx * 3), "cnt": cu.to_host(d_cnt, np.uint32, npx)}
        cu.free(d_rgb); cu.free(d_cnt)
    k, b = out["kd"], out["bvh"]
    print(json.dumps({"render": scene, "xres": params.xres, "yres": params.yres, "spp": params.multisample, "depth": params.depth,
                      "kd_ms_per_round": k["ms"], "bvh_ms_per_round": b["ms"], "speedup": min(k["ms"]) / min(b["ms"]),
                      "rays_equal": k["closest_rays"] == b["closest_rays"] and k["shadow_rays"] == b["shadow_rays"],
                      "bvh_rays": b["bvh_rays"], "bvh_deferred_frac": b["bvh_deferred"] / max(1, b["bvh_rays"]),
                      "framebuffer_words_differing": int((k["fb"] != b["fb"]).sum()), "counts_differing": int((k["cnt"] != b["cnt"]).sum())}), flush=True)


if __name__ == "__main__":
    main()
