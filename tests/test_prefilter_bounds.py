"""The device's second conservative filter (Traverser::outside_bounds, rgk_b200/csrc/trace_device.cuh) against a float32
emulation of Triangle::TestIntersection (src/primitives.cpp:75-166), without a GPU: whenever the exact test accepts a ray,
the approximate hit point must lie inside the triangle's widened 2-D bounds (tri_prefilter_bounds, host_scene.cpp) plus
the run-time margin -- for well-shaped, sliver, tiny, huge and far-from-origin triangles, rays aimed at edges and
vertices, and the worst-case error of the approximate t.  (The GPU suite checks the same on real batches through the
`prefilter_wrong` counter; this test goes looking for trouble.)"""
import numpy as np
import pytest

from rgk_b200 import abi, device
from rgk_b200.scene import ScenePack

F = np.float32


def _scene(tris):
    """A pack holding the given triangles [n,3,3] (one mesh, one diffuse material)."""
    pack = ScenePack()
    pack.add_material("m", abi.BXDF_DIFFUSE, tex_diffuse=pack.add_solid_texture((0.5, 0.5, 0.5)))
    P = tris.reshape(-1, 3).astype(F)
    n = len(P)
    pack.add_mesh(P, np.tile(np.array([[0, 1, 0]], F), (n, 1)), np.zeros((n, 2), F), np.zeros((n, 3), F),
                  np.arange(n, dtype=np.uint32).reshape(-1, 3), "m")
    return pack


def _triangles(rng, n):
    kinds = rng.integers(0, 6, n)
    base = rng.uniform(-1, 1, (n, 3))
    e1, e2 = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
    scale = np.ones(n)
    scale[kinds == 1] = 1e-3                       # tiny
    scale[kinds == 2] = 50.0                       # huge
    sl = kinds == 3                                # slivers: second edge almost parallel to the first
    e2[sl] = e1[sl] * rng.uniform(0.3, 2.0, (sl.sum(), 1)) + rng.normal(size=(sl.sum(), 3)) * 10.0 ** rng.uniform(-5, -2, (sl.sum(), 1))
    far = kinds == 4                               # far from the origin: large coordinates, ordinary size
    base[far] *= 500.0
    ax = kinds == 5                                # axis-aligned edges (q1.x == 0 cases of the 2-D projection)
    e1[ax] = np.eye(3)[rng.integers(0, 3, ax.sum())] * rng.uniform(0.1, 2, (ax.sum(), 1))
    v0 = base
    v1 = v0 + e1 * scale[:, None]
    v2 = v0 + e2 * scale[:, None]
    return np.stack([v0, v1, v2], 1).astype(F)


def _exact_accepts(o, d, plane, rec, eps):
    """float32 emulation of TestIntersection on the precomputed record; returns (accept, t)."""
    dot = F(F(F(d[0] * plane[0]) + F(d[1] * plane[1])) + F(d[2] * plane[2]))
    if np.isnan(dot) or abs(dot) < eps:
        return False, F(0)
    dot2 = F(F(F(o[0] * plane[0]) + F(o[1] * plane[1])) + F(o[2] * plane[2]))
    t = F(-(np.float64(plane[3]) + np.float64(dot2)) / np.float64(dot))
    flags = rec[11].view(np.uint32)
    code = int(flags & 3)
    i1, i2 = ((1, 2), (0, 2), (0, 1))[code]
    q0x = F(F(o[i1] + F(d[i1] * t)) - rec[4])
    q0y = F(F(o[i2] + F(d[i2] * t)) - rec[5])
    q1x, q1y, q2x, q2y, den = rec[6], rec[7], rec[8], rec[9], rec[10]
    with np.errstate(all="ignore"):
        if flags & 4:
            beta = F(q0x / q2x)
            if beta < 0 or beta > 1:
                return False, t
            alpha = F(F(q0y - F(beta * q2y)) / q1y)
        else:
            beta = F(F(F(q0y * q1x) - F(q0x * q1y)) / den)
            if beta < 0 or beta > 1:
                return False, t
            alpha = F(F(q0x - F(beta * q2x)) / q1x)
    if np.isnan(alpha) or np.isnan(beta) or alpha < 0 or F(alpha + beta) > 1:
        return False, t
    return True, t


def _outside(o, d, t32, b):
    code = int(b[0].view(np.uint32) & 3)
    i1, i2 = ((1, 2), (0, 2), (0, 1))[code]
    m = F(F(F(F(abs(t32) + abs(o[0])) + abs(o[1])) + abs(o[2])) * F(7.62939453125e-6))
    p1 = F(np.float64(d[i1]) * np.float64(t32) + np.float64(o[i1]))     # fma: one rounding
    p2 = F(np.float64(d[i2]) * np.float64(t32) + np.float64(o[i2]))
    return bool(F(p1 + m) < b[0] or F(p1 - m) > b[1] or F(p2 + m) < b[2] or F(p2 - m) > b[3])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_bounds_never_reject_what_the_exact_test_accepts(seed):
    rng = np.random.default_rng(seed)
    tris = _triangles(rng, 400)
    pack = _scene(tris)
    hs = device.HostScene(pack.desc())
    planes, rec = hs.records()
    bounds = hs.bounds()
    eps = F(hs.info().epsilon)
    hs.close()
    accepted = rejected_ok = 0
    for ti in range(len(tris)):
        v0, v1, v2 = tris[ti].astype(np.float64)
        # targets: vertices, edge points, interior, and points just outside, each nudged by a few ulps of the triangle size
        bary = [(0, 0), (1, 0), (0, 1), (0.5, 0), (0, 0.5), (0.5, 0.5), (1 / 3, 1 / 3), (1e-7, 1e-7), (1 - 1e-7, 0), (0.999999, 1e-6), (-1e-7, 0.3), (0.3, -1e-7)]
        for a, b in bary:
            target = v0 + a * (v1 - v0) + b * (v2 - v0)
            for _ in range(2):
                dirn = rng.normal(size=3); dirn /= np.linalg.norm(dirn)
                dist = 10.0 ** rng.uniform(-2, 3)
                o = (target - dirn * dist).astype(F)
                d = dirn.astype(F)
                d = (d / F(np.sqrt(F(F(F(d[0] * d[0]) + F(d[1] * d[1])) + F(d[2] * d[2]))))).astype(F)
                ok, t = _exact_accepts(o, d, planes[ti], rec[ti], eps)
                if not ok:
                    continue
                accepted += 1
                # the device's approximate t differs from the exact one by < 2^-21 |t|: try both extremes and the value itself
                for t32 in (t, F(t * F(1 + 2.0 ** -21)), F(t * F(1 - 2.0 ** -21))):
                    assert not _outside(o, d, t32, bounds[ti]), (ti, a, b, float(t), tris[ti].tolist())
    assert accepted > 2000
    # and the filter is not vacuous: points far off the triangle in its plane are outside the bounds of ordinary triangles
    for ti in range(len(tris)):
        v0, v1, v2 = tris[ti].astype(np.float64)
        L = max(np.abs(v1 - v0).max(), np.abs(v2 - v0).max())
        if not np.isfinite(bounds[ti]).all() or abs(bounds[ti][1]) > 1e30:
            continue
        target = v0 + 4.0 * (v1 - v0) + 4.0 * (v2 - v0)
        n = np.cross(v1 - v0, v2 - v0)
        if np.linalg.norm(n) < 1e-12 * L * L:
            continue
        dirn = -n / np.linalg.norm(n)
        o = (target - dirn * L).astype(F)
        d = dirn.astype(F)
        t32 = F(L)
        rejected_ok += _outside(o, d, t32, bounds[ti])
    assert rejected_ok > len(tris) // 3
