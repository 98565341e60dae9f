"""GPU parity of what bench.py and the BASELINE configs actually run ("make the benched path the tested
path"): the large-set sampler instantiation k_sampler_mt<false> (spp 121 / 256 / 512 -> 529 / 1024: BASELINE C3, C4, C5), a
scene with the thinglass set non-empty (src/scene_intersect.cpp:330-455, selected at src/path_tracer.cpp:128-133,431-432), the
bidirectional mode on the library-default wide-BVH traversal, the randomized differential campaign of tests/bvh_campaign.py
on the device, and the headline workload at full size (1080p x 64 spp) kd against BVH, bit for bit.  All through the C ABI."""
import os

import numpy as np
import pytest

import checkers
from rgk_b200 import abi, device, scenes, standin

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("ms", [121, 256, 512, 1024])
def test_large_set_sampler_tables_bit_exact(gpu_ctx, oracle, ms):
    """Set sizes above 100 leave the shared-memory instantiation (MT_LANES * ss * 8 bytes > 100 KB) and are shuffled in
    place in global memory (src/sampler.cpp:77-116): the instantiation every BASELINE config above 64 spp takes."""
    seeds = np.array([42 + 0x42424242, 7, 0, 0xFFFFFFFF, 123456789, 0x42424242 * 3 + 99], np.uint32)
    n1d, n2d = (64, 64) if ms <= 256 else (7, 9)      # 1024 spp x 64 dims x 6 seeds on the scalar oracle takes a while: fewer dims
    t1g, t2g = gpu_ctx.sampler_tables(seeds, ms, n1d, n2d)
    t1o, t2o = oracle.sampler_tables(seeds, ms, n1d, n2d)
    ss = gpu_ctx.sampler_set_size(ms)
    assert ss == oracle.sampler_set_size(ms) and ss > 100 and t1g.shape == (len(seeds), n1d, ss)
    assert np.array_equal(t1g.view(np.uint32), t1o.view(np.uint32))
    assert np.array_equal(t2g.view(np.uint32), t2o.view(np.uint32))


def test_large_set_sampler_against_the_reference_fixture(gpu_ctx):
    """The same instantiation against tables the REFERENCE build produced (tools/make_golden.py, sampler_large.npz)."""
    g = np.load(os.path.join(G, "sampler_large.npz"))
    for ms in (256, 512):
        t1, t2 = gpu_ctx.sampler_tables(g["seeds"], ms, g[f"t1_{ms}"].shape[1], g[f"t2_{ms}"].shape[1])
        assert np.array_equal(t1.view(np.uint32), g[f"t1_{ms}"].view(np.uint32))
        assert np.array_equal(t2.view(np.uint32), g[f"t2_{ms}"].view(np.uint32))


def test_large_set_sampler_inside_a_round(gpu_ctx, oracle):
    """A whole round at 121 spp (set size 121 -> k_sampler_mt<false>) with the device sampler equals the same round fed with
    the oracle's tables (RGK_SAMPLER_TABLES), bit for bit: the in-place global tables are indexed the way k_shade reads them."""
    from test_gpu_render import _pixel_seeds
    pack, cfg = scenes.material_zoo(width=32, height=32, multisample=121, recursion_max=3, lens=0.04)
    gpu_ctx.commit(pack.desc())
    cam = gpu_ctx.camera(**cfg.camera_args())
    tasks = gpu_ctx.generate_tasks(32, 32, 32)
    p = cfg.params(abi.SAMPLER_MT19937)
    a, ca, sa = gpu_ctx.render_round(cam, p, tasks, seedcount_base=3)
    seeds = _pixel_seeds(tasks, 42, 3)
    t1, t2 = oracle.sampler_tables(seeds, 121, 1 + p.depth, 5 + p.depth)
    gpu_ctx.set_tables(121, t1, t2)
    b, cb, sb = gpu_ctx.render_round(cam, cfg.params(abi.SAMPLER_TABLES), tasks, seedcount_base=3)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)
    assert int(sa.closest_rays) == int(sb.closest_rays) and int(sa.shadow_rays) == int(sb.shadow_rays)


def test_thinglass_scene(gpu_ctx, oracle):
    """rgk_scene_desc::thinglass != 0: the reference traces such a scene through FindIntersectKdOtherThanWithThinglass and
    VisibilityWithThinglass.  Material::is_thinglass is never set upstream, so both are result-identical to the plain
    variants -- the fixture is the REFERENCE build's own round with a non-empty thinglass set, the oracle takes its
    thinglass branch, and the GPU must give the same hits bit for bit, the same ray counts and the image within the
    stated tolerance (rel-mean <= 1e-3, RMSE <= 2 % of the mean)."""
    g = np.load(os.path.join(G, "cornell_thinglass.npz"))
    pack, cfg = scenes.load_builtin("cornell-box", width=64, height=64, multisample=4)
    pack.thinglass = 1
    desc = pack.desc()
    assert desc.thinglass == 1
    gpu_ctx.commit(desc)
    ho = oracle.scene_create(desc)
    hits = gpu_ctx.trace_closest(g["brays"], g["ign"])
    assert hits.tobytes() == g["bhits"].tobytes()                                     # the reference's FindIntersectKdOtherThanWithThinglass
    assert (gpu_ctx.trace_shadow(g["sa"], g["sb"]) == g["vis"]).all()                 # VisibilityWithThinglass
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params()
    p.depth = 40
    tasks = gpu_ctx.generate_tasks(32, 64, 64)
    fg, cg, sg = gpu_ctx.render_round(cam, p, tasks)
    fo, co, so = oracle.render_round(ho, cam, p, tasks)
    assert np.array_equal(fo.view(np.uint32), g["fb"].view(np.uint32)) and int(so.closest_rays) == int(g["closest_rays"])
    mean = float(g["fb"].mean())
    assert np.array_equal(cg, g["cnt"])
    assert abs(float(fg.mean()) - mean) <= 1e-3 * mean and float(np.sqrt(np.mean((fg - g["fb"]) ** 2))) <= 0.02 * mean
    assert abs(int(sg.closest_rays) - int(g["closest_rays"])) <= 0.001 * int(g["closest_rays"])


@pytest.mark.parametrize("scene,reverse", [("cornell", 2), ("zoo", 3)])
def test_bidirectional_mode_on_the_default_traversal(oracle, scene, reverse):
    """reverse > 0 with the wide BVH on (its closest-hit launches go through k_closest_bvh + the arbiter, its shadow and
    connection segments through the kd kernels): ray counts equal the oracle's, the image equals the kd-only context's bit
    for bit (same kernels apart from the traversal structure, same splat order) and the oracle's within tolerance."""
    if scene == "cornell":
        pack, cfg = scenes.load_builtin("cornell-box", width=64, height=64, multisample=4, recursion_max=6)
    else:
        pack, cfg = scenes.material_zoo(width=48, height=32, multisample=4, recursion_max=4, lens=0.05)
    desc = pack.desc()
    out = []
    for trav in ("bvh", "kd"):
        ctx = device.Context(0, traversal=trav)
        ctx.commit(desc)
        cam = ctx.camera(**cfg.camera_args())
        p = cfg.params()
        p.reverse = reverse
        tasks = ctx.generate_tasks(32, p.xres, p.yres)
        ctx.bvh_stats()
        f, c, st = ctx.render_round(cam, p, tasks)
        out.append((f, c, st, ctx.bvh_stats()))
        ctx.close()
    (fb, cb, sb, bb), (fk, ck, sk, bk) = out
    assert bb["rays"] > 0 and bk["rays"] == 0
    assert int(sb.closest_rays) == int(sk.closest_rays) and int(sb.shadow_rays) == int(sk.shadow_rays)
    # light-path splats are atomic adds: their order is not defined, so kd vs BVH is compared like GPU vs oracle
    ho = oracle.scene_create(desc)
    fo, co, so = oracle.render_round(ho, cam, p, tasks, nthreads=1)
    assert int(sb.closest_rays) == int(so.closest_rays) and int(sb.shadow_rays) + int(sb.shadow_rays_skipped) == int(so.shadow_rays)
    mean = float(fo.mean())
    for f in (fb, fk):
        assert abs(float(f.mean()) - mean) <= 2e-3 * mean
        assert float(np.sqrt(np.mean((f - fo) ** 2))) <= 0.05 * mean
    assert float(np.abs(fb - fk).max()) <= 1e-3 * max(1.0, float(np.abs(fk).max()))


@pytest.mark.parametrize("seed", range(9100, 9124))
def test_bvh_campaign_on_the_device(seed):
    """tests/bvh_campaign.py's random soups (sliver / tiny / huge / far-away triangles, exact duplicates, rays aimed at
    vertices and edges, far origins, a third with an ignored triangle) through the BVH + arbiter KERNELS: every hit record
    and every visibility flag equals the kd oracle's.  On the host build this campaign found the two deferral rules added
    late in round 1; here it runs with real 32-lane warps, ballots and the arbiter launch."""
    from test_prefilter_bounds import _scene, _triangles
    rng = np.random.default_rng(seed)
    n_t = int(rng.choice([50, 400, 3000]))
    tris = _triangles(rng, n_t)
    if seed % 2:
        tris = np.concatenate([tris, tris[: n_t // 10]])
    if seed % 3 == 0:
        tris = (tris + np.float32(rng.choice([0, 100, -5000]))).astype(np.float32)
    pack = _scene(tris)
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    ctx = device.Context(0, traversal="bvh")
    ctx.commit(pack.desc())
    n = 20000
    pick = rng.integers(0, len(tris), n)
    w = rng.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)
    kind = rng.integers(0, 4, n)
    w[kind == 0] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (kind == 0).sum())]
    e = kind == 1
    w[e, 2] = 0
    w[e, :2] /= w[e, :2].sum(1, keepdims=True)
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
    ctx.bvh_stats()
    got = ctx.trace_closest(rays, ign)
    s = ctx.bvh_stats()
    want = O.trace_closest(h, rays, ign)
    has_bvh = s["rays"] > 0            # scenes with NaN-prone triangles stay on the kd-tree (host_scene.cpp)
    assert got.tobytes() == want.tobytes(), (seed, has_bvh)
    a, b = rays["origin"], target[keep]
    far = np.linalg.norm(a - b, axis=1) > 0.1
    assert (ctx.trace_shadow(a[far], b[far]) == O.trace_shadow(h, a[far], b[far])).all(), (seed, has_bvh)
    ctx.close()
    O.scene_destroy(h)


def test_headline_round_kd_equals_bvh_at_full_size():
    """The bench workload itself -- sponza stand-in, 1920x1080, 64 spp, one round -- through the kd-only context and the
    default one: identical ray counts, 0 framebuffer words differing."""
    pack, cfg = standin.sponza()
    desc = pack.desc()
    out = []
    for trav in ("kd", "bvh"):
        ctx = device.Context(0, traversal=trav)
        ctx.commit(desc)
        cam = ctx.camera(**cfg.camera_args())
        p = cfg.params()
        tasks = ctx.generate_tasks(32, p.xres, p.yres)
        ctx.bvh_stats()
        f, c, st = ctx.render_round(cam, p, tasks, seedcount_base=5 * len(tasks))
        out.append((f, c, (int(st.closest_rays), int(st.shadow_rays), int(st.shadow_rays_skipped)), ctx.bvh_stats()))
        ctx.close()
    (fk, ck, rk, bk), (fb, cb, rb, bb) = out
    assert bk["rays"] == 0 and bb["rays"] == rb[0] + rb[1] and bb["ambiguous"] < 5e-3 * bb["rays"]
    assert rk == rb
    assert int((fk.view(np.uint32) != fb.view(np.uint32)).sum()) == 0 and np.array_equal(ck, cb)
    assert int(ck.min()) == cfg.multisample
