"""Randomized differential campaign: the host-compiled device source (kd traversal, variant 6, and the wide-BVH pass, variant 4)
against the oracle on random triangle soups (tiny / huge / sliver / far / axis-aligned / duplicated triangles, scenes moved far
from the origin, rays aimed at vertices and edges from inside and from ~2000 units away).  Found the two rules added at the end
of round 1 (scaled boundary width, NaN-prone triangles); every fourth soup has sheared twins (round 2).   python tests/bvh_campaign.py <seconds> [first seed]"""
import os, sys, time, numpy as np, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers, test_device_on_host as T
from test_prefilter_bounds import _scene, _triangles
lib = C.CDLL(T.SO); vp = C.c_void_p
lib.doh_scene_create.restype = vp
lib.doh_scene_create.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_float, vp]
lib.doh_scene_destroy.argtypes = [vp]
lib.doh_closest.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, vp, vp, vp]
lib.doh_shadow.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, vp, vp]
O = checkers.oracle()
t_end = time.time() + float(sys.argv[1])
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
seed = seed0; bad = 0; rays_total = 0
while time.time() < t_end:
    seed += 1
    rng = np.random.default_rng(seed)
    n_t = int(rng.choice([50, 400, 3000]))
    tris = _triangles(rng, n_t)
    if seed % 2: tris = np.concatenate([tris, tris[: n_t // 10]])
    if seed % 3 == 0: tris = (tris + np.float32(rng.choice([0, 100, -5000]))).astype(np.float32)     # far from the origin
    if seed % 4 == 0:                                                   # sheared twins: v1 within ~eps of v0 in one coordinate (|q1.x| < eps branch)
        diam = np.linalg.norm(tris.reshape(-1, 3).max(0) - tris.reshape(-1, 3).min(0))
        sel = rng.random(len(tris)) < 0.4
        ax = rng.integers(0, 3, len(tris))
        snap = tris[np.arange(len(tris)), 0, ax] + (rng.uniform(-2, 2, len(tris)) * 1e-5 * diam).astype(np.float32)
        tris[np.arange(len(tris))[sel], 1, ax[sel]] = snap[sel]
    pack = _scene(tris)
    h = O.scene_create(pack.desc()); S = T.Scene(lib, pack)
    has_bvh = len(S.keep[5]) > 0
    nobvh = globals().get('nobvh', 0) + (0 if has_bvh else 1); globals()['nobvh'] = nobvh
    n = 20000
    pick = rng.integers(0, len(tris), n)
    w = rng.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)
    kind = rng.integers(0, 4, n)
    w[kind == 0] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (kind == 0).sum())]
    e = kind == 1; w[e, 2] = 0; w[e, :2] /= w[e, :2].sum(1, keepdims=True)
    target = np.einsum("nk,nkd->nd", w, tris[pick]).astype(np.float32)
    c = tris.reshape(-1, 3).mean(0)
    origin = (c + np.where(rng.random((n, 1)) < 0.5, rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(-2000, 2000, (n, 3)))).astype(np.float32)
    d = target - origin; keep = np.linalg.norm(d, axis=1) > 1e-6
    rays = np.zeros(int(keep.sum()), checkers.RAY_DT)
    rays["origin"] = origin[keep]; rays["direction"] = (d[keep] / np.linalg.norm(d[keep], axis=1, keepdims=True)).astype(np.float32); rays["tfar"] = 10000.0
    ign = np.where(rng.random(len(rays)) < 0.3, pick[keep], 0xFFFFFFFF).astype(np.uint32)
    want = O.trace_closest(h, rays, ign)
    for variant in ((6, 4) if has_bvh else (6,)):
        got, df, _ = S.closest(variant, rays, ign)
        if not T._same(got[~df], want[~df]):
            bad += 1; print("MISMATCH closest seed", seed, "variant", variant, flush=True)
    a, b = rays["origin"], target[keep]; far = np.linalg.norm(a - b, axis=1) > 0.1
    wv = O.trace_shadow(h, a[far], b[far])
    for variant in ((6, 4) if has_bvh else (6,)):
        vis, df = S.shadow(variant, a[far], b[far])
        if not (vis[~df] == wv[~df]).all():
            bad += 1; print("MISMATCH shadow seed", seed, "variant", variant, int(((vis != wv) & ~df).sum()), flush=True)
    rays_total += 2 * len(rays)
    S.close(); O.scene_destroy(h)
print("scenes without bvh", globals().get("nobvh", 0)); print("scenes", seed - seed0, "rays", rays_total, "mismatching batches", bad, flush=True)
