"""Ray batches for traversal parity: primary rays of a camera, cosine-distributed bounce
rays leaving the first hits (origin offset faceN*10*eps, ignore = hit triangle, like
src/path_tracer.cpp:291-295) and light->surface shadow segments (src/scene.cpp:670-673)."""
import numpy as np

from checkers import RAY_DT


def primary(checker, cam, w, h, jitter_seed=None):
    ys, xs = np.mgrid[0:h, 0:w]
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    if jitter_seed is None:
        off = np.full((len(xy), 2), 0.5, np.float32)
    else:
        off = np.random.default_rng(jitter_seed).random((len(xy), 2), dtype=np.float32)
    return checker.camera_rays(cam, w, h, xy, off)


def bounce(rays, hits, normals_of_tri, eps, seed=0x52474B31):
    """Secondary rays from the hit points with cosine-distributed directions about the geometric normal."""
    rng = np.random.default_rng(seed)
    ok = hits["triangle"] != 0xFFFFFFFF
    r, h = rays[ok], hits[ok]
    n = normals_of_tri[h["triangle"]].astype(np.float32)
    d = r["direction"]
    facing = (np.einsum("ij,ij->i", n, d) < 0)[:, None]
    n = np.where(facing, n, -n).astype(np.float32)
    pos = (r["origin"] + h["t"][:, None] * d).astype(np.float32)
    u = rng.random((len(r), 2), dtype=np.float32)
    rr, ph = np.sqrt(u[:, 0]), 2 * np.pi * u[:, 1]
    lx, ly, lz = rr * np.cos(ph), rr * np.sin(ph), np.sqrt(np.maximum(0, 1 - u[:, 0]))
    a = np.where(np.abs(n[:, :1]) > 0.9, np.array([[0, 1, 0]], np.float32), np.array([[1, 0, 0]], np.float32))
    t = np.cross(n, a); t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    dirs = (lx[:, None] * t + ly[:, None] * b + lz[:, None] * n).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True).astype(np.float32)
    out = np.zeros(len(r), dtype=RAY_DT)
    out["origin"] = pos + n * np.float32(eps * 10.0)
    out["direction"] = dirs.astype(np.float32)
    out["tnear"] = 0.0
    out["tfar"] = 10000.0
    return out, h["triangle"].astype(np.uint32).copy()


def shadow_segments(rays, hits, light_pos):
    ok = hits["triangle"] != 0xFFFFFFFF
    r, h = rays[ok], hits[ok]
    b = (r["origin"] + h["t"][:, None] * r["direction"]).astype(np.float32)
    a = np.broadcast_to(np.asarray(light_pos, np.float32), b.shape).copy()
    return a, b
