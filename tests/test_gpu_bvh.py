"""GPU parity of the wide-BVH traversal (RGK_TRAVERSAL_BVH, the library default; bvh_device.cuh + the kd arbiter pass) against the
oracle, through the C ABI: hit records and visibility flags bit-exact, like the kd path (test_gpu_trace.py), with the
BVH counters proving that the BVH kernels -- not the kd ones -- produced them."""
import os

import numpy as np
import pytest

import checkers
import raybatches
from rgk_b200 import device, standin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    pack, cfg = standin.sponza(width=480, height=270, multisample=1)
    ctx = device.Context(0, traversal="bvh")
    ctx.commit(pack.desc())
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    cam = ctx.camera(**cfg.camera_args())
    ys, xs = np.mgrid[0:270, 0:480]
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    rays = ctx.camera_rays(cam, 480, 270, xy, np.random.default_rng(9).random((len(xy), 2), dtype=np.float32))
    yield pack, ctx, O, h, rays
    ctx.close()


def _same(a, b):
    return all((a[f].view(np.uint32) == b[f].view(np.uint32)).all() for f in ("triangle", "t", "a", "b", "c"))


def test_closest_and_shadow_match_the_oracle(setup):
    pack, ctx, O, h, rays = setup
    ctx.bvh_stats()
    want = O.trace_closest(h, rays)
    got = ctx.trace_closest(rays)
    s = ctx.bvh_stats()
    assert s["rays"] == len(rays) and s["ambiguous"] < 3e-3 * len(rays)
    assert _same(got, want)
    eps = O.scene_info(h).epsilon
    brays, ign = raybatches.bounce(rays, want, O.scene_planes(h)[:, :3], eps)
    wantb = O.trace_closest(h, brays, ign)
    gotb = ctx.trace_closest(brays, ign)
    s = ctx.bvh_stats()
    assert s["rays"] == len(brays) and s["ambiguous"] < 3e-3 * len(brays)
    assert _same(gotb, wantb)
    light = np.asarray(pack.point_lights[0][0], np.float32)
    a, b = raybatches.shadow_segments(brays, wantb, light)
    # in-plane segments between surface points exercise the degenerate-ray deferral (kd NaN-interval case)
    a2 = b[np.random.default_rng(3).permutation(len(b))]
    keep = np.linalg.norm(a2 - b, axis=1) > 1e-3
    for aa, bb in ((a, b), (a2[keep], b[keep])):
        wantv = O.trace_shadow(h, aa, bb)
        gotv = ctx.trace_shadow(aa, bb)
        s = ctx.bvh_stats()
        assert s["rays"] == len(aa)
        assert (gotv == wantv).all()
        assert 0 < wantv.mean() < 1


def test_counting_mode_and_empty_batch(setup):
    pack, ctx, O, h, rays = setup
    ctx.set_counting(True)
    try:
        ctx.bvh_stats()
        got = ctx.trace_closest(rays[:50000])
        s = ctx.bvh_stats()
        assert s["rays"] == 50000 and 5 * s["rays"] < s["nodes"] < 40 * s["rays"] and s["tests"] < 20 * s["rays"]
        assert _same(got, O.trace_closest(h, rays[:50000]))
    finally:
        ctx.set_counting(False)
    assert len(ctx.trace_closest(rays[:0])) == 0


def test_render_round_is_bit_identical_to_the_kd_path():
    """The whole wavefront (closest + shadow launches through the BVH, ambiguous rays through the kd arbiter) must produce
    the same framebuffer, bit for bit, and the same ray counts as the kd-only context."""
    pack, cfg = standin.sponza(width=256, height=144, multisample=4)
    desc = pack.desc()
    kd = device.Context(0, traversal="kd"); kd.commit(desc)
    bv = device.Context(0, traversal="bvh"); bv.commit(desc)
    out = []
    for ctx in (kd, bv):
        cam = ctx.camera(**cfg.camera_args())
        tasks = ctx.generate_tasks(64, 256, 144)
        params = cfg.params()
        params.depth = 3
        ctx.bvh_stats()
        rgb, cnt, stats = ctx.render_round(cam, params, tasks)
        out.append(((rgb, cnt), stats, ctx.bvh_stats()))
    (fb0, st0, b0), (fb1, st1, b1) = out
    assert b0["rays"] == 0 and b1["rays"] == st1.closest_rays + st1.shadow_rays
    assert b1["ambiguous"] < 3e-3 * b1["rays"]
    assert st0.closest_rays == st1.closest_rays and st0.shadow_rays == st1.shadow_rays
    assert (fb0[0].view(np.uint32) == fb1[0].view(np.uint32)).all() and (fb0[1] == fb1[1]).all()
    kd.close(); bv.close()


def test_sheared_accept_region():
    """Triangles of TestIntersection's |q1.x| < eps branch (src/primitives.cpp:141-147) are accepted up to eps outside their true
    extents (tests/test_bvh_host.py::test_sheared_accept_region): BVH pass + kd arbiter against the oracle on rays across the
    true and the sheared edges."""
    from test_bvh_host import sheared_fan
    pack, tris, rays, (a, b) = sheared_fan()
    ctx = device.Context(0, traversal="bvh")
    ctx.commit(pack.desc())
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    ctx.bvh_stats()
    got = ctx.trace_closest(rays)
    s = ctx.bvh_stats()
    assert s["rays"] == len(rays) and 0 < s["ambiguous"] < len(rays)
    assert _same(got, O.trace_closest(h, rays))
    assert (ctx.trace_shadow(a, b) == O.trace_shadow(h, a, b)).all()
    ctx.close()


def test_grazing_hits():
    """Rays skimming the Cornell box's walls (tests/test_bvh_host.py::test_grazing_hits; the first one is the ray of the soak that the
    kd-tree misses and the BVH alone would hit): BVH pass + kd arbiter against the oracle."""
    from test_bvh_host import grazing_rays
    from rgk_b200 import scenes
    pack = scenes.load_builtin("cornell-box")[0]
    ctx = device.Context(0, traversal="bvh")
    ctx.commit(pack.desc())
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    rays, ign = grazing_rays()
    ctx.bvh_stats()
    got = ctx.trace_closest(rays, ign)
    s = ctx.bvh_stats()
    want = O.trace_closest(h, rays, ign)
    assert s["rays"] == len(rays) and 0 < s["ambiguous"] < len(rays)
    assert want["triangle"][0] == 0xFFFFFFFF and _same(got, want)
    ctx.close()
