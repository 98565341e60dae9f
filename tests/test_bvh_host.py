"""The wide BVH (RGK_TRAVERSAL_BVH, the default; rgk_b200/csrc/host_bvh.cpp + bvh_device.cuh), checked without a GPU: structure of
the product-built nodes, and -- through the oracle's CPU mirror of the device traversal logic (rgko_bvh4_*) -- that every
ray the BVH pass would commit has exactly the kd-tree's answer (Scene::FindIntersectKdOtherThan / Scene::Visibility,
src/scene_intersect.cpp:211-327, src/scene.cpp:670-673) while only a tiny fraction is deferred to the kd pass."""
import ctypes as C
import os

import numpy as np
import pytest

import checkers
import raybatches
from rgk_b200 import device, standin, scenes


@pytest.fixture(scope="module")
def setup():
    pack, cfg = standin.sponza(width=320, height=180, multisample=1)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    ca = cfg.camera_args()
    cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    rays = raybatches.primary(O, cam, 320, 180, jitter_seed=5)
    hits = O.trace_closest(h, rays)
    yield pack, cfg, hs, O, h, rays, hits
    hs.close()


def test_off_on_the_kd_traversal():
    pack = scenes.load_builtin("cornell-box")[0]
    hs = device.HostScene(pack.desc(), traversal="kd")
    nodes, order, depth = hs.bvh()
    assert len(nodes) == 0 and len(order) == 0 and depth == 0
    hs.close()


def test_structure(setup):
    pack, cfg, hs, O, h, rays, hits = setup
    _check_structure(pack, hs)


@pytest.mark.parametrize("knobs", [{"bvh_reinsert_iters": 1}, {"bvh_greedy_collapse": 1, "bvh_all_axes": 0, "bvh_bins": 16}])
def test_structure_under_the_builder_knobs(knobs):
    """The study knobs of host_bvh.cpp (insertion-based optimisation, the first version's greedy collapse) still produce a
    valid tree whose committed rays match the kd-tree."""
    pack, cfg = standin.sponza(width=160, height=90, multisample=1)
    hs = device.HostScene(pack.desc(), traversal="bvh", **knobs)
    _check_structure(pack, hs)
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    ca = cfg.camera_args()
    cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    rays = raybatches.primary(O, cam, 160, 90, jitter_seed=4)
    nodes, order, _ = hs.bvh()
    closest, _ = _mirror(O, h, nodes, order)
    got, deferred, _ = closest(rays)
    assert _same(got[~deferred], O.trace_closest(h, rays)[~deferred])
    hs.close()


def _check_structure(pack, hs):
    nodes, order, depth = hs.bvh()
    nt = hs.info().n_triangles
    assert len(nodes) > 0 and 3 * depth + 1 <= 64
    assert sorted(order.tolist()) == list(range(nt))                       # every triangle in exactly one leaf slot
    codes = nodes[:, 24:28].view(np.uint32)
    assert not nodes[:, 28:].any()
    arr = pack.arrays()
    pos = np.asarray(arr["positions"], np.float32).reshape(-1, 3)
    tri = np.asarray(arr["indices"], np.uint32).reshape(-1, 3)
    tlo, thi = pos[tri].min(1), pos[tri].max(1)
    # what TestIntersection accepts is the projected triangle (v1 moved to v0's first projected coordinate in the |q1.x| < eps
    # branch, src/primitives.cpp:141-147) lifted onto the stored fp32 plane: the boxes hold those three corners too (2 ulps outwards)
    planes, rec = hs.records()
    planes = np.asarray(planes, np.float32).reshape(-1, 4).astype(np.float64); rec = np.asarray(rec, np.float32).reshape(-1, 12)
    flags = rec[:, 11].view(np.uint32)
    i1 = np.where((flags & 3) == 0, 1, 0); i2 = np.where((flags & 3) == 2, 1, 2); k = 3 - i1 - i2
    rows = np.arange(nt)
    ok = (planes[rows, k] != 0) & np.isfinite(planes).all(1)
    v = pos[tri].astype(np.float64)                                         # [nt, 3 corners, 3]
    for c in range(3):
        a = np.where((c == 1) & ((flags & 4) != 0), v[rows, 0, i1], v[rows, c, i1]); b = v[rows, c, i2]
        with np.errstate(all="ignore"):
            lifted = -(planes[:, 3] + planes[rows, i1] * a + planes[rows, i2] * b) / planes[rows, k]
        w = np.zeros((nt, 3)); w[rows, i1] = a; w[rows, i2] = b; w[rows, k] = lifted
        use = ok & np.isfinite(lifted)
        lo = hi = w.astype(np.float32)
        for _ in range(2):
            lo = np.nextafter(lo, np.float32(-np.inf)); hi = np.nextafter(hi, np.float32(np.inf))
        tlo[use] = np.minimum(tlo[use], lo[use]); thi[use] = np.maximum(thi[use], hi[use])
    seen_inner, covered = set(), np.zeros(nt, bool)
    # subtree bounds bottom-up: children always have larger indices than their parent (preorder emission)
    sub_lo, sub_hi = np.zeros((len(nodes), 3), np.float32), np.zeros((len(nodes), 3), np.float32)
    for i in range(len(nodes) - 1, -1, -1):
        lo_i, hi_i = np.full(3, np.inf, np.float32), np.full(3, -np.inf, np.float32)
        for c in range(4):
            code = int(codes[i, c])
            lo = nodes[i, [0 + c, 8 + c, 16 + c]]; hi = nodes[i, [4 + c, 12 + c, 20 + c]]
            if code == 0x7FFFFFFF:
                assert np.isposinf(lo).all() and np.isposinf(hi).all()
                continue
            if code & 0x80000000:
                first, cnt = code & 0x1FFFFFFF, ((code >> 29) & 3) + 1
                t = order[first:first + cnt]
                assert not covered[t].any()
                covered[t] = True
                assert (tlo[t].min(0) == lo).all() and (thi[t].max(0) == hi).all()     # exact bounds of its triangles' accept regions
            else:
                assert i < code < len(nodes) and code not in seen_inner
                seen_inner.add(code)
                assert (sub_lo[code] == lo).all() and (sub_hi[code] == hi).all()
            lo_i, hi_i = np.minimum(lo_i, lo), np.maximum(hi_i, hi)
        sub_lo[i], sub_hi[i] = lo_i, hi_i
    assert covered.all() and len(seen_inner) == len(nodes) - 1


def _mirror(O, h, nodes, order):
    lib = O.lib
    vp = C.c_void_p
    lib.rgko_bvh4_closest.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, vp, vp, vp]
    lib.rgko_bvh4_shadow.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, vp, vp, vp]
    nodes = np.ascontiguousarray(nodes); order = np.ascontiguousarray(order)

    def closest(rays, ignore=None):
        hits = np.zeros(len(rays), checkers.HIT_DT); status = np.zeros(len(rays), np.uint8); cnt = np.zeros(2, np.uint64)
        rays = np.ascontiguousarray(rays)
        lib.rgko_bvh4_closest(h, nodes.ctypes.data, order.ctypes.data, rays.ctypes.data, None if ignore is None else ignore.ctypes.data,
                              len(rays), hits.ctypes.data, status.ctypes.data, cnt.ctypes.data)
        return hits, status.astype(bool), cnt

    def shadow(a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        vis = np.zeros(len(a), np.uint8); status = np.zeros(len(a), np.uint8); cnt = np.zeros(2, np.uint64)
        lib.rgko_bvh4_shadow(h, nodes.ctypes.data, order.ctypes.data, a.ctypes.data, b.ctypes.data, len(a), vis.ctypes.data, status.ctypes.data, cnt.ctypes.data)
        return vis, status.astype(bool), cnt
    return closest, shadow


def _same(a, b):
    return all((a[f].view(np.uint32) == b[f].view(np.uint32)).all() for f in ("triangle", "t", "a", "b", "c"))


def test_committed_rays_match_the_kdtree(setup):
    pack, cfg, hs, O, h, rays, hits = setup
    nodes, order, _ = hs.bvh()
    closest, shadow = _mirror(O, h, nodes, order)
    # primary rays
    got, deferred, cnt = closest(rays)
    assert deferred.mean() < 3e-3
    assert _same(got[~deferred], hits[~deferred])
    assert cnt[0] / len(rays) < 25 and cnt[1] / len(rays) < 12
    # incoherent bounce rays with the hit triangle ignored
    eps = O.scene_info(h).epsilon
    brays, ign = raybatches.bounce(rays, hits, O.scene_planes(h)[:, :3], eps)
    want = O.trace_closest(h, brays, ign)
    got, deferred, cnt = closest(brays, ign)
    assert deferred.mean() < 3e-3
    assert _same(got[~deferred], want[~deferred])
    # shadow segments from the light to the first hits, and between random pairs of hit points (mostly blocked)
    light = np.asarray(pack.point_lights[0][0], np.float32)
    a, b = raybatches.shadow_segments(rays, hits, light)
    b2 = raybatches.shadow_segments(brays, want, light)[1]          # points on every wall, not just the visible ones
    # The second batch has segments lying exactly in the floor plane (a zero direction component on a kd split plane:
    # the reference's NaN-interval case, where its leaves accept hits anywhere along the line) -- all of them must defer --
    # and the third has segments that run within an ulp of a wall and graze the edges of the triangles touching it, which
    # the kd-tree sees or not depending on the side of the cell face the ray is on (the edge-hit deferral).
    for aa, bb, max_deferred in ((a, b, 1e-3), (b[np.random.default_rng(2).permutation(len(b))], b, 0.2),
                                 (b2[np.random.default_rng(3).permutation(len(b2))], b2, 0.2)):
        keep = np.linalg.norm(aa - bb, axis=1) > 1e-3
        aa, bb = aa[keep], bb[keep]
        want = O.trace_shadow(h, aa, bb)
        vis, deferred, _ = shadow(aa, bb)
        assert deferred.mean() < max_deferred
        assert (vis[~deferred] == want[~deferred]).all()
        assert 0 < want.mean() < 1
        in_plane = ((bb - aa) == 0).any(1)
        assert deferred[in_plane].all()


@pytest.mark.parametrize("name", ["cornell-box", "material-zoo"])
def test_small_scenes_through_the_mirror(name):
    """Few, large, axis-aligned triangles (Cornell box: pixel-centre rays with zero direction components, hits on shared
    edges) and the material zoo: committed rays equal the kd-tree's answer over primary rays and three bounces."""
    pack, cfg = scenes.load_builtin("cornell-box", width=128, height=128, multisample=1) if name == "cornell-box" else scenes.material_zoo(width=128, height=96)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    ca = cfg.camera_args()
    cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    nodes, order, _ = hs.bvh()
    closest, shadow = _mirror(O, h, nodes, order)
    eps = O.scene_info(h).epsilon
    light = np.asarray(pack.point_lights[0][0], np.float32) if pack.point_lights else np.array([0, 0.9, 0], np.float32)
    for jitter in (None, 3):
        cur = raybatches.primary(O, cam, cfg.xres, cfg.yres, jitter_seed=jitter)
        ign = None
        for bounce in range(4):
            want = O.trace_closest(h, cur, ign)
            got, deferred, _ = closest(cur, ign)
            assert deferred.mean() < 0.05
            assert _same(got[~deferred], want[~deferred])
            a, b = raybatches.shadow_segments(cur, want, light)
            if len(a):
                vis, dfs, _ = shadow(a, b)
                assert (vis[~dfs] == O.trace_shadow(h, a, b)[~dfs]).all()
            cur, ign = raybatches.bounce(cur, want, O.scene_planes(h)[:, :3], eps, seed=7 + bounce)
            if len(cur) == 0:
                break
    hs.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_triangle_soup_stress(seed):
    """Goes looking for trouble: a soup of well-shaped, tiny (smaller than epsilon), huge, sliver, far-from-origin and
    axis-aligned triangles plus exact coplanar duplicates, with rays aimed at vertices, edges and interiors from inside and
    far outside the scene.  Whatever the BVH pass would commit must be the kd-tree's answer; everything else must defer."""
    from test_prefilter_bounds import _scene, _triangles
    rng = np.random.default_rng(seed)
    tris = _triangles(rng, 3000)
    tris = np.concatenate([tris, tris[:200], tris[:100] + np.float32(1e-6)])          # exact and near-exact duplicates
    pack = _scene(tris)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    nodes, order, depth = hs.bvh()
    assert len(nodes) > 0
    closest, shadow = _mirror(O, h, nodes, order)
    n = 60000
    pick = rng.integers(0, len(tris), n)
    w = rng.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)                          # clusters near vertices and edges
    kind = rng.integers(0, 4, n)
    w[kind == 0] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (kind == 0).sum())]     # exactly a vertex
    e = kind == 1                                                                       # exactly on an edge
    w[e, 2] = 0; w[e, :2] /= w[e, :2].sum(1, keepdims=True)
    target = np.einsum("nk,nkd->nd", w, tris[pick]).astype(np.float32)
    origin = np.where(rng.random((n, 1)) < 0.5, rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(-2000, 2000, (n, 3))).astype(np.float32)
    d = target - origin
    keep = np.linalg.norm(d, axis=1) > 1e-6
    rays = np.zeros(int(keep.sum()), checkers.RAY_DT)
    rays["origin"] = origin[keep]
    rays["direction"] = (d[keep] / np.linalg.norm(d[keep], axis=1, keepdims=True)).astype(np.float32)
    rays["tnear"] = 0.0
    rays["tfar"] = 10000.0
    ax = rng.random(len(rays)) < 0.05                                                   # some axis-aligned directions
    rays["direction"][ax] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, ax.sum())] * np.where(rng.random((ax.sum(), 1)) < 0.5, 1, -1).astype(np.float32)
    ign = np.where(rng.random(len(rays)) < 0.3, pick[keep], 0xFFFFFFFF).astype(np.uint32)
    want = O.trace_closest(h, rays, ign)
    got, deferred, _ = closest(rays, ign)
    assert deferred.mean() < 0.6 and (want["triangle"] != 0xFFFFFFFF).mean() > 0.3
    assert _same(got[~deferred], want[~deferred])
    a, b = rays["origin"], target[keep]
    far_enough = np.linalg.norm(a - b, axis=1) > 0.1
    vis, dfs, _ = shadow(a[far_enough], b[far_enough])
    assert (vis[~dfs] == O.trace_shadow(h, a[far_enough], b[far_enough])[~dfs]).all()
    hs.close()


@pytest.mark.parametrize("n", [1, 2, 4, 5, 9])
def test_tiny_scenes(n):
    """One wide node with a single leaf, a full leaf, the first split: structure and committed rays on scenes of 1-9 triangles."""
    from test_prefilter_bounds import _scene
    rng = np.random.default_rng(n)
    base = rng.uniform(-1, 1, (n, 1, 3)).astype(np.float32)
    tris = (base + rng.normal(scale=0.5, size=(n, 3, 3))).astype(np.float32)
    pack = _scene(tris)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    _check_structure(pack, hs)
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    nodes, order, depth = hs.bvh()
    closest, shadow = _mirror(O, h, nodes, order)
    m = 4000
    rays = np.zeros(m, checkers.RAY_DT)
    rays["origin"] = rng.uniform(-3, 3, (m, 3)).astype(np.float32)
    target = tris[rng.integers(0, n, m)].mean(1) + rng.normal(scale=0.05, size=(m, 3)).astype(np.float32)
    d = target - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tfar"] = 10000.0
    want = O.trace_closest(h, rays)
    got, deferred, _ = closest(rays)
    assert (want["triangle"] != 0xFFFFFFFF).mean() > 0.05 and deferred.mean() < 0.2
    assert _same(got[~deferred], want[~deferred])
    hs.close()


@pytest.mark.parametrize("case", ["identical", "zero-area", "collinear-centroids"])
def test_degenerate_inputs(case):
    """200 copies of one triangle (every hit is an exact tie: all must defer), zero-area triangles among ordinary ones, and a
    row of triangles with collinear centroids: valid tree, committed rays equal to the kd-tree's answer."""
    from test_prefilter_bounds import _scene
    rng = np.random.default_rng(0)
    if case == "identical":
        tris = np.repeat(rng.normal(size=(1, 3, 3)).astype(np.float32), 200, 0)
    elif case == "zero-area":
        tris = np.concatenate([rng.normal(size=(50, 3, 3)).astype(np.float32), np.repeat(rng.normal(size=(20, 1, 3)).astype(np.float32), 3, 1)])
    else:
        tris = np.stack([np.stack([[i, 0, 0], [i + 0.5, 1, 0], [i + 0.5, 0, 1]]) for i in range(300)]).astype(np.float32)
    pack = _scene(tris)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    _check_structure(pack, hs)
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    nodes, order, _ = hs.bvh()
    closest, _ = _mirror(O, h, nodes, order)
    m = 5000
    rays = np.zeros(m, checkers.RAY_DT)
    rays["origin"] = rng.uniform(-3, 3, (m, 3)).astype(np.float32)
    d = tris[rng.integers(0, len(tris), m)].mean(1) + rng.normal(scale=0.05, size=(m, 3)).astype(np.float32) - rays["origin"]
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tfar"] = 10000.0
    want = O.trace_closest(h, rays)
    got, deferred, _ = closest(rays)
    hit = want["triangle"] != 0xFFFFFFFF
    assert hit.mean() > 0.5 and _same(got[~deferred], want[~deferred])
    if case == "identical":
        assert deferred[hit].all()
    else:
        assert deferred.mean() < 0.02
    hs.close()


@pytest.mark.parametrize("seed", [5019, 5038, 5541, 5550])
def test_campaign_regressions(seed):
    """Seeds a randomized differential campaign (22 M rays) tripped over: rays from ~2000 units away aimed at vertices of small
    triangles (the boundary width of the edge rule has to scale with the ray's position uncertainty over the triangle's
    height), and soups moved to -5000 where tiny triangles collapse onto collinear float32 vertices (the reference accepts
    their NaN-barycentric hits anywhere on the plane inside a referencing kd leaf: such scenes keep the kd-tree for every ray)."""
    from test_prefilter_bounds import _scene, _triangles
    rng = np.random.default_rng(seed)
    n_t = int(rng.choice([50, 400, 3000]))
    tris = _triangles(rng, n_t)
    if seed % 2:
        tris = np.concatenate([tris, tris[: n_t // 10]])
    if seed % 3 == 0:
        tris = (tris + np.float32(rng.choice([0, 100, -5000]))).astype(np.float32)
    pack = _scene(tris)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    nodes, order, _ = hs.bvh()
    if seed in (5541, 5550):
        assert len(nodes) == 0                       # NaN-prone triangles: no wide BVH for this scene
        hs.close()
        return
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    closest, _ = _mirror(O, h, nodes, order)
    n = 20000
    pick = rng.integers(0, len(tris), n)
    w = rng.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)
    kind = rng.integers(0, 4, n)
    w[kind == 0] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (kind == 0).sum())]
    e = kind == 1
    w[e, 2] = 0; w[e, :2] /= w[e, :2].sum(1, keepdims=True)
    target = np.einsum("nk,nkd->nd", w, tris[pick]).astype(np.float32)
    c = tris.reshape(-1, 3).mean(0)
    origin = (c + np.where(rng.random((n, 1)) < 0.5, rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(-2000, 2000, (n, 3)))).astype(np.float32)
    d = target - origin
    keep = np.linalg.norm(d, axis=1) > 1e-6
    rays = np.zeros(int(keep.sum()), checkers.RAY_DT)
    rays["origin"] = origin[keep]
    rays["direction"] = (d[keep] / np.linalg.norm(d[keep], axis=1, keepdims=True)).astype(np.float32)
    rays["tfar"] = 10000.0
    ign = np.where(rng.random(len(rays)) < 0.3, pick[keep], 0xFFFFFFFF).astype(np.uint32)
    got, deferred, _ = closest(rays, ign)
    assert _same(got[~deferred], O.trace_closest(h, rays, ign)[~deferred])
    hs.close()


def sheared_fan(seed=7, n=40000):
    """A fan of triangles facing +z whose v1 is q1x (|q1x| < eps, != 0) off v0 in the first projected coordinate, inside a
    20 x 200 x 20 frame (eps = 1e-5 * diameter ~ 2e-3, the rays' box margin ~ 1e-4).  v1 is the lowest vertex of a plane sloped
    in x, so for q1x > 0 the sheared corner lies ~1e-3 below the triangle's extents; the neighbour across v0-v1 is attached for
    q1x < 0 and moved 1 below otherwise (a free sheared corner).  Rays aimed across the true and the sheared edge, and segments
    through the same points.  -> pack, triangles, rays, (a, b)"""
    from test_prefilter_bounds import _scene
    rng = np.random.default_rng(seed)
    box = np.array([[[-10, -100, -10], [10, -100, -10], [-10, 100, -10]], [[10, 100, 10], [-10, 100, 10], [10, -100, 10]]], np.float32)
    tris = [box[0], box[1]]
    q1xs = [1.8e-3, -1.8e-3, 6e-4, -1.2e-3, 2e-3, 1e-4]
    for j, q1x in enumerate(q1xs):
        v0 = np.array([j - 3.0, -0.5, 0.1 * j], np.float32)
        tris.append(np.stack([v0, v0 + np.array([q1x, 0.02, -1e-4], np.float32), v0 + np.array([0.03, 0.0, 0.015], np.float32)]))
        nb = np.stack([v0, v0 + np.array([-0.03, 0.02, -0.014], np.float32), v0 + np.array([q1x, 0.02, -1e-4], np.float32)])
        tris.append(nb if j % 2 == 1 else nb + np.array([0, 0, -1], np.float32))       # every other neighbour 1 below: a free sheared corner
    tris = np.stack(tris).astype(np.float32)
    pick = 2 + 2 * rng.integers(0, len(q1xs), n)
    v0, v1 = tris[pick, 0], tris[pick, 1]
    s = rng.random((n, 1)).astype(np.float32)
    target = v0 + s * (v1 - v0)
    target[:, 0] += rng.uniform(-4e-3, 4e-3, n).astype(np.float32)
    origin = target + np.stack([rng.normal(scale=0.3, size=n), rng.normal(scale=0.3, size=n), rng.uniform(0.5, 3.0, n)], 1).astype(np.float32)
    d = (target - origin).astype(np.float32)
    rays = np.zeros(n, checkers.RAY_DT)
    rays["origin"] = origin
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tfar"] = 10000.0
    return _scene(tris), tris, rays, (origin, (target + 0.5 * d).astype(np.float32))


def test_sheared_accept_region():
    """Triangle::TestIntersection's |q1.x| < eps branch (src/primitives.cpp:141-147) drops q1.x: a triangle whose second vertex
    is within eps of the first in the first projected coordinate is accepted as a sheared twin, up to eps outside its true
    extents.  Found by the 4K x 1024 spp conference frame (one pixel of 8.3 M: a chair-leg triangle hit 4e-5 inside the sheared
    edge and 1e-4 outside its box; tools/find_bvh_mismatch.py, tools/repro_tile_on_host.py).  The boxes hold the sheared region,
    and hits near its boundary defer."""
    pack, tris, rays, (a, b) = sheared_fan()
    hs = device.HostScene(pack.desc(), traversal="bvh")
    _check_structure(pack, hs)
    rec = np.asarray(hs.records()[1], np.float32).reshape(-1, 12)
    sheared = ((rec[:, 11].view(np.uint32) & 4) != 0) & (rec[:, 6] != 0)
    assert sheared[2::2].all()
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    nodes, order, _ = hs.bvh()
    closest, shadow = _mirror(O, h, nodes, order)
    want = O.trace_closest(h, rays)
    got, deferred, _ = closest(rays)
    hit = want["triangle"] != 0xFFFFFFFF
    on_sheared = np.isin(want["triangle"], np.nonzero(sheared)[0])
    # some of the reference's hits lie outside the hit triangle's true extents: the case the exact boxes missed
    p = rays["origin"] + rays["direction"] * want["t"][:, None]
    tw = tris[np.where(hit, want["triangle"], 0)]
    outside = hit & ((p < tw.min(1) - 3e-4) | (p > tw.max(1) + 3e-4)).any(1)
    assert on_sheared.mean() > 0.2 and outside.sum() > 100 and 0.0 < deferred.mean() < 0.9
    assert _same(got[~deferred], want[~deferred])
    vis, dfs, _ = shadow(a, b)
    assert (vis[~dfs] == O.trace_shadow(h, a, b)[~dfs]).all()
    hs.close()


@pytest.mark.parametrize("seed", [11, 12])
def test_slivers_off_their_plane(seed):
    """Long slivers (5 units by 1e-4 .. 1e-2): the cross product behind Triangle::CalculatePlane (src/primitives.cpp:24-36) loses
    its direction, so the stored plane misses the triangle's own far vertices and the accepted region (projected triangle
    lifted onto that plane) leaves the true extents.  The boxes hold the lifted corners; triangles whose plane is off by more
    than eps / 4 carry flag 8 and always defer.  Committed rays must be the kd-tree's answer."""
    from test_prefilter_bounds import _scene
    rng = np.random.default_rng(seed)
    n_t = 400
    a = rng.uniform(-6, 6, (n_t, 3))
    u = rng.normal(size=(n_t, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = np.cross(u, rng.normal(size=(n_t, 3))); w /= np.linalg.norm(w, axis=1, keepdims=True)
    height = 10.0 ** rng.uniform(-4, -2, (n_t, 1))
    tris = np.stack([a, a + 5.0 * u, a + 2.5 * u + height * w], 1).astype(np.float32)
    walls = rng.uniform(-8, 8, (40, 1, 3)) + rng.normal(scale=2.0, size=(40, 3, 3))
    tris = np.concatenate([tris, walls.astype(np.float32)])
    pack = _scene(tris)
    hs = device.HostScene(pack.desc(), traversal="bvh")
    nodes, order, _ = hs.bvh()
    if len(nodes) == 0:
        pytest.skip("a sliver collapsed onto collinear fp32 vertices: the scene keeps the kd-tree")
    _check_structure(pack, hs)
    rec = np.asarray(hs.records()[1], np.float32).reshape(-1, 12)
    off_plane = (rec[:, 11].view(np.uint32) & 8) != 0
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    closest, shadow = _mirror(O, h, nodes, order)
    n = 40000
    pick = rng.integers(0, n_t, n)
    s = rng.random((n, 1)); t = rng.random((n, 1)) * 1.2 - 0.1
    target = (tris[pick, 0] * (1 - s) + tris[pick, 1] * s) * (1 - t) + tris[pick, 2] * t
    origin = target + rng.normal(scale=2.0, size=(n, 3))
    d = (target - origin).astype(np.float32)
    rays = np.zeros(n, checkers.RAY_DT)
    rays["origin"] = origin.astype(np.float32)
    rays["direction"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tfar"] = 10000.0
    want = O.trace_closest(h, rays)
    got, deferred, _ = closest(rays)
    hit = want["triangle"] != 0xFFFFFFFF
    print("off-plane triangles", int(off_plane.sum()), "hits", hit.mean(), "hits on slivers", (want["triangle"][hit] < n_t).mean(), "deferred", deferred.mean())
    assert (want["triangle"][hit] < n_t).mean() > 0.05
    assert _same(got[~deferred], want[~deferred])
    assert off_plane[:n_t].sum() > 20 and not off_plane[n_t:].any()
    assert deferred[hit & off_plane[np.where(hit, want["triangle"], 0)]].all()         # every ray whose answer is such a triangle went to the kd pass
    b = (target + 0.3 * d).astype(np.float32)
    vis, dfs, _ = shadow(rays["origin"], b)
    assert (vis[~dfs] == O.trace_shadow(h, rays["origin"], b)[~dfs]).all()
    hs.close()


def grazing_rays(n=200000, seed=3):
    """The Cornell-box ray of test_grazing_hits (first entry; its ignore triangle is 1) and n rays skimming the box's walls:
    origins 1e-5 .. 1e-2 off a wall, drifting towards it at 1e-5 .. 3e-2 of the direction.  -> rays, ignore"""
    rng = np.random.default_rng(seed)
    lo, hi = np.array([-1, 0, -1], np.float32), np.array([1, 2, 1], np.float32)
    origin = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    ax = rng.integers(0, 3, n); side = rng.integers(0, 2, n)
    wall = np.where(side == 1, hi[ax], lo[ax]); inward = np.where(side == 1, -1.0, 1.0)
    origin[np.arange(n), ax] = wall + inward * 10.0 ** rng.uniform(-5, -2, n)
    d = rng.normal(size=(n, 3))
    d[np.arange(n), ax] = -inward * 10.0 ** rng.uniform(-5, -1.5, n)
    rays = np.zeros(n + 1, checkers.RAY_DT)
    rays["origin"][1:] = origin
    rays["direction"][1:] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["origin"][0] = (0.403505802, 1.99871016, -0.999653339); rays["direction"][0] = (-0.23824501, 0.000627510715, 0.971204877)
    rays["tfar"] = 10000.0
    ign = np.full(n + 1, 0xFFFFFFFF, np.uint32); ign[0] = 1
    return rays, ign


def test_grazing_hits():
    """A ray 0.04 degrees off the Cornell box's ceiling (whose fp32 plane passes a quarter ulp above the triangle's highest
    vertex): the hit is interior to the triangle but lies beyond its extents, in a kd cell that does not reference it, 1.9e-4 of
    ray parameter (5 eps) past the leaf that does -- the kd-tree reports a miss (1 path in 36 rounds of the 256 x 256 x 16 spp
    Cornell render, tools/soak_bvh_vs_kd.py).  Grazing hits defer.  The ray itself, and 200 k rays skimming the walls."""
    pack = scenes.load_builtin("cornell-box")[0]
    hs = device.HostScene(pack.desc(), traversal="bvh")
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    nodes, order, _ = hs.bvh()
    closest, shadow = _mirror(O, h, nodes, order)
    rays, ign = grazing_rays()
    want = O.trace_closest(h, rays, ign)
    got, deferred, _ = closest(rays, ign)
    assert want["triangle"][0] == 0xFFFFFFFF and deferred[0]
    assert deferred.mean() < 0.8 and (want["triangle"] != 0xFFFFFFFF).mean() > 0.5
    assert _same(got[~deferred], want[~deferred])
    hs.close()
