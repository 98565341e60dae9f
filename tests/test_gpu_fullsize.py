"""BASELINE-size checks (configs[1]: 1920x1080 on the ~71 k-triangle stand-in): the oracle cannot trace 2 M+ rays in
seconds, so the full batch is checked through size-independent properties and a random subset bit-exactly."""
import numpy as np
import pytest

import raybatches
from rgk_b200 import abi, standin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sponza(gpu_ctx, oracle):
    pack, cfg = standin.sponza(multisample=1)
    desc = pack.desc()
    gpu_ctx.commit(desc)
    return pack, cfg, desc, oracle.scene_create(desc)


def test_full_hd_primary_and_shadow_properties(gpu_ctx, oracle, sponza):
    pack, cfg, desc, ho = sponza
    # the host-built tree of a 71 k-triangle scene equals the reference-procedure tree word for word
    no, ro = oracle.scene_kdtree(ho)
    ng, rg = gpu_ctx.scene_kdtree()
    assert np.array_equal(no, ng) and np.array_equal(ro, rg)
    cam = gpu_ctx.camera(**cfg.camera_args())
    ys, xs = np.mgrid[0:1080, 0:1920]
    xy = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
    rays = gpu_ctx.camera_rays(cam, 1920, 1080, xy, np.random.default_rng(0).random((len(xy), 2), dtype=np.float32))
    hits, st = gpu_ctx.trace_closest(rays, want_stats=True)
    ok = hits["triangle"] != 0xFFFFFFFF
    assert 0.5 < ok.mean() < 1.0 and int(st.rays) == len(rays)
    assert st.prefilter_wrong == 0 and st.prefiltered > st.exact     # the pre-filter settles most candidates, never wrongly
    a = pack.arrays()
    P, I = a["positions"].astype(np.float64), a["indices"]
    h, r = hits[ok], rays[ok]
    assert h["triangle"].max() < len(I)
    # barycentrics form a partition of unity inside the triangle, t is positive and inside the ray's range
    assert np.all(h["b"] >= 0) and np.all(h["c"] >= 0) and np.all(h["b"] + h["c"] <= 1.0)
    assert np.allclose(h["a"] + h["b"] + h["c"], 1.0, atol=2e-6)
    assert np.all(h["t"] > 0) and np.all(h["t"] <= 10000.0 + 1.0)
    # the hit point lies on the triangle: interpolated vertices == origin + t * direction (to fp32 accuracy)
    tri = I[h["triangle"]]
    interp = h["a"][:, None] * P[tri[:, 0]] + h["b"][:, None] * P[tri[:, 1]] + h["c"][:, None] * P[tri[:, 2]]
    point = r["origin"].astype(np.float64) + h["t"][:, None].astype(np.float64) * r["direction"]
    assert np.abs(interp - point).max() < 2e-3          # scene diameter 33: ~1e-5 relative (+ the 2-D projected test)
    # a random subset, bit-exact against the oracle (closest hits and light->surface visibility)
    sel = np.random.default_rng(1).choice(len(rays), 40000, replace=False)
    assert hits[sel].tobytes() == oracle.trace_closest(ho, rays[sel]).tobytes()
    sa, sb = raybatches.shadow_segments(rays, hits, pack.point_lights[0][0])
    vis = gpu_ctx.trace_shadow(sa, sb)
    sel2 = np.random.default_rng(2).choice(len(sa), 40000, replace=False)
    assert np.array_equal(vis[sel2], oracle.trace_shadow(ho, sa[sel2], sb[sel2]))
    # idempotence: the same batch again gives the same bytes; reversing the batch order reverses the result
    assert gpu_ctx.trace_closest(rays).tobytes() == hits.tobytes()
    assert gpu_ctx.trace_closest(rays[::-1].copy())[::-1].tobytes() == hits.tobytes()


def test_full_hd_round_linearity_and_counts(gpu_ctx, sponza):
    """One 1920x1080 round at 4 spp: counts are analytic, two rounds accumulate linearly, a second call with the same
    seeds reproduces the image bit for bit (deterministic per-pixel ownership, no float atomics)."""
    pack, cfg, desc, ho = sponza
    cfg.multisample = 4
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params(abi.SAMPLER_MT19937)
    tasks = gpu_ctx.generate_tasks(32, 1920, 1080)
    f1, c1, s1 = gpu_ctx.render_round(cam, p, tasks, seedcount_base=0)
    assert int(c1.min()) == 4 and int(c1.max()) == 4 and int(s1.samples) == 1920 * 1080 * 4
    assert np.isfinite(f1).all() and f1.min() >= 0.0
    assert int(s1.closest_rays) >= int(s1.samples) and int(s1.shadow_rays) <= int(s1.closest_rays)
    f1b, _, _ = gpu_ctx.render_round(cam, p, tasks, seedcount_base=0)
    assert np.array_equal(f1.view(np.uint32), f1b.view(np.uint32))
    f2, c2, _ = gpu_ctx.render_round(cam, p, tasks, seedcount_base=len(tasks))
    acc = (f1.copy(), c1.copy())
    gpu_ctx.render_round(cam, p, tasks, seedcount_base=len(tasks), fb=acc)
    assert np.array_equal(acc[0], f1 + f2) and int(acc[1].min()) == 8
    assert not np.array_equal(f1, f2)                       # different seeds -> different noise
    assert abs(float(f1.mean()) - float(f2.mean())) < 0.05 * float(f1.mean())


def test_queue_order_never_changes_a_result(gpu_ctx, sponza, monkeypatch):
    """Direction-binned queues (k_bin) vs atomic compaction, different reordering group sizes and refill thresholds: the
    order in which a launch processes its rays must not change one bit of the image."""
    pack, cfg, desc, ho = sponza
    cfg.multisample = 4
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params(abi.SAMPLER_MT19937)
    tasks = gpu_ctx.generate_tasks(32, 1920, 1080)
    ref_img = None
    base = gpu_ctx.cfg()
    for env in ({"binning": 0}, {"binning": 1}, {"binning": 1, "bin_items": 8192, "refill_incoherent": 4, "refill_shadow": 32},
                {"binning": 1, "bin_shadow_first": 0, "chunk_paths": 3000000}):
        gpu_ctx.configure(abi.DeviceCfg.from_buffer_copy(bytes(base)), **env)      # scheduling fields apply to the next call
        f, c, st = gpu_ctx.render_round(cam, p, tasks, seedcount_base=7)
        if ref_img is None:
            ref_img, ref_rays = f, (int(st.closest_rays), int(st.shadow_rays))
        else:
            assert np.array_equal(f.view(np.uint32), ref_img.view(np.uint32)), env
            assert (int(st.closest_rays), int(st.shadow_rays)) == ref_rays
    gpu_ctx.configure(base)
