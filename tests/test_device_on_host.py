"""The DEVICE SOURCE of the traversal (rgk_b200/csrc/trace_device.cuh, bvh_device.cuh) compiled for the host
(tests/host_cpp/device_shim.h, device_on_host.cpp) and run against the oracle -- not a restatement of it: the kd-tree
traversal in both control structures with its two conservative pre-filters (bit-exact hit records, oracle-identical
counters, prefilter_wrong == 0) and the wide-BVH candidate pass with its deferral rules, through the same persistent-warp
drivers the kernels use (one-lane warps).  Scene arrays come from the product's host commit (rgk_host_scene_*).
Further down: the shading functions, and whole rounds of the wavefront (render.cu's kernels and host loop, the device sampler
included) against the oracle's framebuffer.  Covers what nvcc compiles except the two inline-PTX helpers (host branches),
k_bin (queue order only) and real warp divergence."""
import ctypes as C
import os

import numpy as np
import pytest

import checkers
import raybatches
from rgk_b200 import device, scenes, standin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.environ.get("RGK_DEVICE_ON_HOST_SO") or os.path.join(ROOT, "build", "host", "libdevice_on_host.so")    # tools/asan_device_on_host.sh
vp = C.c_void_p


@pytest.fixture(scope="module")
def doh():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(SO)
    lib.doh_scene_create.restype = vp
    lib.doh_scene_create.argtypes = [vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_float, vp]
    lib.doh_scene_destroy.argtypes = [vp]
    lib.doh_closest.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, vp, vp, vp]
    lib.doh_shadow.argtypes = [vp, C.c_int, vp, vp, C.c_uint64, vp, vp]
    lib.doh_shade_scene_create.restype = vp
    lib.doh_shade_scene_create.argtypes = [vp, vp]
    lib.doh_shade_scene_destroy.argtypes = [vp]
    lib.doh_probe.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp]
    lib.doh_render_round.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
    return lib


class Scene:
    def __init__(self, lib, pack):
        hs = device.HostScene(pack.desc(), traversal="bvh")
        info = hs.info()
        self.keep = [np.ascontiguousarray(a) for a in (*hs.kdtree(), *hs.records(), hs.bounds(), *hs.bvh()[:2], np.array(list(info.bbox), np.float32))]
        nodes, refs, planes, rec, bounds, bnodes, border, bbox = self.keep
        p = lambda a: a.ctypes.data
        self.lib = lib
        self.h = vp(lib.doh_scene_create(p(nodes), len(nodes) // 2, p(refs), len(refs), p(planes), p(rec), p(bounds), info.n_triangles,
                                         p(bnodes), len(bnodes), p(border), info.epsilon, p(bbox)))
        hs.close()

    def closest(self, variant, rays, ignore=None):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), checkers.HIT_DT); status = np.zeros(len(rays), np.uint8); cnt = np.zeros(8, np.uint64)
        rc = self.lib.doh_closest(self.h, variant, rays.ctypes.data, None if ignore is None else ignore.ctypes.data, len(rays), hits.ctypes.data,
                                  status.ctypes.data, cnt.ctypes.data)
        assert rc == 0
        return hits, status.astype(bool), cnt

    def shadow(self, variant, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        vis = np.zeros(len(a), np.uint8); status = np.zeros(len(a), np.uint8)
        assert self.lib.doh_shadow(self.h, variant, a.ctypes.data, b.ctypes.data, len(a), vis.ctypes.data, status.ctypes.data) == 0
        return vis, status.astype(bool)

    def close(self):
        self.lib.doh_scene_destroy(self.h)


def _same(a, b):
    return all((a[f].view(np.uint32) == b[f].view(np.uint32)).all() for f in ("triangle", "t", "a", "b", "c"))


def _batches(O, h, pack, cfg, w, hgt):
    ca = cfg.camera_args()
    cam = O.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    rays = raybatches.primary(O, cam, w, hgt, jitter_seed=21)
    hits, st = O.trace_closest(h, rays, want_stats=True)
    eps = O.scene_info(h).epsilon
    brays, ign = raybatches.bounce(rays, hits, O.scene_planes(h)[:, :3], eps, seed=5)
    bhits, bst = O.trace_closest(h, brays, ign, want_stats=True)
    light = np.asarray(pack.point_lights[0][0], np.float32) if pack.point_lights else np.array([0, 0.9, 0], np.float32)
    a, b = raybatches.shadow_segments(brays, bhits, light)
    return (rays, None, hits, st), (brays, ign, bhits, bst), (a, b, O.trace_shadow(h, a, b))


@pytest.mark.parametrize("name", ["sponza", "cornell-box", "material-zoo"])
def test_device_source_on_the_host(doh, name):
    if name == "sponza":
        pack, cfg = standin.sponza(width=240, height=135, multisample=1)
    elif name == "cornell-box":
        pack, cfg = scenes.load_builtin("cornell-box", width=128, height=128, multisample=1)
    else:
        pack, cfg = scenes.material_zoo(width=128, height=96)
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    S = Scene(doh, pack)
    prim, bounce, (a, b, want_vis) = _batches(O, h, pack, cfg, cfg.xres, cfg.yres)
    for rays, ign, want, st in (prim, bounce):
        for variant in (2, 6):                                   # trace_persistent / trace_phased
            got, _, cnt = S.closest(variant, rays, ign)
            assert _same(got, want)
            # counters of the counting instantiation: the oracle's, and the second pre-filter never wrong
            assert (int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3])) == (st.inner, st.leaf, st.refs, st.tests)
            assert int(cnt[6]) == 0 and int(cnt[4]) + int(cnt[5]) > 0
        got, deferred, cnt = S.closest(4, rays, ign)             # wide-BVH pass
        assert deferred.mean() < 0.05
        assert _same(got[~deferred], want[~deferred])
        assert 0 < cnt[0] < 40 * len(rays)
    for variant in (2, 6):
        vis, _ = S.shadow(variant, a, b)
        assert (vis == want_vis).all()
    vis, deferred = S.shadow(4, a, b)
    assert deferred.mean() < 0.05 and (vis[~deferred] == want_vis[~deferred]).all()
    S.close()


def test_device_source_on_a_triangle_soup(doh):
    """The stress scene of test_bvh_host.py (tiny, huge, sliver, far, axis-aligned and duplicated triangles; rays aimed at
    vertices and edges from inside and far outside) through the device source: kd traversal bit-exact with
    prefilter_wrong == 0, wide-BVH pass equal on everything it would commit."""
    from test_prefilter_bounds import _scene, _triangles
    rng = np.random.default_rng(11)
    tris = _triangles(rng, 2500)
    tris = np.concatenate([tris, tris[:150], tris[:80] + np.float32(1e-6)])
    pack = _scene(tris)
    O = checkers.oracle()
    h = O.scene_create(pack.desc())
    S = Scene(doh, pack)
    n = 40000
    pick = rng.integers(0, len(tris), n)
    w = rng.dirichlet([0.3, 0.3, 0.3], n).astype(np.float32)
    kind = rng.integers(0, 4, n)
    w[kind == 0] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, (kind == 0).sum())]
    e = kind == 1
    w[e, 2] = 0; w[e, :2] /= w[e, :2].sum(1, keepdims=True)
    target = np.einsum("nk,nkd->nd", w, tris[pick]).astype(np.float32)
    origin = np.where(rng.random((n, 1)) < 0.5, rng.uniform(-1.5, 1.5, (n, 3)), rng.uniform(-2000, 2000, (n, 3))).astype(np.float32)
    d = target - origin
    keep = np.linalg.norm(d, axis=1) > 1e-6
    rays = np.zeros(int(keep.sum()), checkers.RAY_DT)
    rays["origin"] = origin[keep]
    rays["direction"] = (d[keep] / np.linalg.norm(d[keep], axis=1, keepdims=True)).astype(np.float32)
    rays["tfar"] = 10000.0
    ax = rng.random(len(rays)) < 0.05
    rays["direction"][ax] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, ax.sum())]
    # a third of the rays with a bounded [tnear, tfar] that often ends on or just before the target
    bounded = rng.random(len(rays)) < 0.33
    dist = np.linalg.norm(d[keep], axis=1).astype(np.float32)
    rays["tnear"][bounded] = (dist[bounded] * rng.uniform(0.0, 0.6, bounded.sum())).astype(np.float32)
    rays["tfar"][bounded] = (dist[bounded] * rng.choice([0.8, 1.0, 1.0, 1.2], bounded.sum())).astype(np.float32)
    ign = np.where(rng.random(len(rays)) < 0.3, pick[keep], 0xFFFFFFFF).astype(np.uint32)
    want, st = O.trace_closest(h, rays, ign, want_stats=True)
    for variant in (2, 6):
        got, _, cnt = S.closest(variant, rays, ign)
        assert _same(got, want)
        assert (int(cnt[0]), int(cnt[1]), int(cnt[2]), int(cnt[3])) == (st.inner, st.leaf, st.refs, st.tests) and int(cnt[6]) == 0
    got, deferred, _ = S.closest(4, rays, ign)
    assert deferred.mean() < 0.6 and _same(got[~deferred], want[~deferred])
    a, b = rays["origin"], target[keep]
    far = np.linalg.norm(a - b, axis=1) > 0.1
    want_vis = O.trace_shadow(h, a[far], b[far])
    for variant in (2, 6):
        assert (S.shadow(variant, a[far], b[far])[0] == want_vis).all()
    vis, dfs = S.shadow(4, a[far], b[far])
    assert (vis[~dfs] == want_vis[~dfs]).all()
    S.close()


class HostProber:
    """Stands in for device.Context in the shading unit tests: .probe() runs probe_one (probe_device.cuh) on the host."""

    def __init__(self, lib, desc):
        self.lib, self.desc = lib, desc
        self.h = vp(lib.doh_shade_scene_create(C.byref(desc), None))
        assert self.h.value

    def probe(self, kind, index, rows):
        from rgk_b200 import abi
        rows = np.ascontiguousarray(rows, np.float32)
        win, wout = abi.PROBE_WIDTHS[kind]
        assert rows.ndim == 2 and rows.shape[1] == win
        out = np.zeros((len(rows), wout), np.float32)
        assert self.lib.doh_probe(self.h, kind, index, rows.ctypes.data, len(rows), out.ctypes.data) == 0
        return out

    def close(self):
        self.lib.doh_shade_scene_destroy(self.h)


def test_shading_device_source_on_the_host(doh, oracle):
    """The unit-level shading parity tests of test_gpu_shading.py (every live BxDF incl. the LTC lobes, textures and bump
    slopes, light picking, envmap sky, local frames) with shade_device.cuh compiled for the host instead of the GPU: same
    inputs, same golden fixture, same tolerances."""
    import test_gpu_shading as T
    pack, cfg = scenes.material_zoo(width=48, height=32, multisample=4, lens=0.05)
    desc = pack.desc()
    ctx = HostProber(doh, desc)
    zoo = (pack, cfg, desc, oracle.scene_create(desc), np.load(os.path.join(T.G, "zoo.npz")))
    T.test_bxdf_value_and_sample_all_kinds(ctx, oracle, zoo)
    T.test_textures_lights_sky_frames(ctx, oracle, zoo)
    ctx.close()


def _host_round(doh, oracle, pack, cfg, depth=None, seedcount_base=0, wide_bvh=False, reverse=None, device_sampler=False, **cfg_fields):
    """One round of render_round_impl compiled for the host (kernels + host loop of render.cu), fed with the oracle's
    StratifiedSampler tables, next to the oracle's own round."""
    from rgk_b200 import abi
    from test_gpu_render import _pixel_seeds
    desc = pack.desc()
    # one-lane warps: refill after every ray; no k_bin (it cooperates through shared memory)
    if device_sampler:
        cfg_fields.setdefault("sampler_kernel", 2)       # the warp-per-pixel builder (the GPU default) unless a test asks otherwise
    fields = dict(binning=0, refill_coherent=1, refill_incoherent=1, refill_shadow=1)
    fields.update(cfg_fields)       # (the 32-lane build of test_wavefront_lanes_on_host.py passes the library's own scheduling defaults)
    dcfg = abi.device_cfg(traversal="bvh" if wide_bvh else "kd", **fields)
    if True:
        h = vp(doh.doh_shade_scene_create(C.byref(desc), C.byref(dcfg)))
        assert h.value
        ho = oracle.scene_create(desc)
        ca = cfg.camera_args()
        cam = oracle.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
        p = cfg.params(abi.SAMPLER_MT19937 if device_sampler else abi.SAMPLER_TABLES)
        if depth is not None:
            p.depth = depth
        if reverse is not None:
            p.reverse = reverse
        tasks = oracle.generate_tasks(32, p.xres, p.yres)
        lens = 1 if ca["lens_size"] != 0.0 else 0
        n1d, n2d = 1 + p.depth, 4 + lens + p.depth + p.reverse
        seeds = _pixel_seeds(tasks, 42, seedcount_base)
        t1, t2 = oracle.sampler_tables(seeds, p.multisample, n1d, n2d)
        t1 = np.ascontiguousarray(t1, np.float32); t2 = np.ascontiguousarray(t2, np.float32)
        rgb = np.zeros((p.yres, p.xres, 3), np.float32); cnt = np.zeros((p.yres, p.xres), np.uint32)
        st = abi.RoundStats(); bvh = np.zeros(2, np.uint64)
        rc = doh.doh_render_round(h, C.byref(dcfg), C.byref(cam), C.byref(p), tasks, len(tasks), 42, seedcount_base, None if device_sampler else t1.ctypes.data,
                                  None if device_sampler else t2.ctypes.data, n1d, n2d,
                                  len(seeds), rgb.ctypes.data, cnt.ctypes.data, C.byref(st), bvh.ctypes.data)
        assert rc == 0
        po = cfg.params(abi.SAMPLER_MT19937); po.depth = p.depth; po.reverse = p.reverse
        fo, co, so = oracle.render_round(ho, cam, po, tasks, seedcount_base=seedcount_base)
        doh.doh_shade_scene_destroy(h)
    return (rgb, cnt, st, bvh), (fo, co, so)


def _zoo():
    return scenes.material_zoo(width=40, height=24, multisample=4, recursion_max=3, lens=0.04)


def _cornell():
    return scenes.load_builtin("cornell-box", width=48, height=48, multisample=4)          # recursion-max 40, Russian roulette


def _sponza():
    return standin.sponza(width=64, height=36, multisample=4)


@pytest.mark.parametrize("scene,wide_bvh,reverse", [(_zoo, False, 0), (_zoo, True, 0), (_cornell, False, 0), (_cornell, True, 0),
                                                    (_sponza, True, 0), (_cornell, False, 2), (_zoo, False, 1), (_cornell, True, 2), (_zoo, True, 1)])
def test_wavefront_device_source_on_the_host(doh, oracle, scene, wide_bvh, reverse):
    """The whole wavefront -- render.cu's kernels AND its host loop -- on the CPU, with the oracle's sampler tables: pixel
    setup, camera rays, closest-hit (kd, or wide BVH + arbiter), k_shade (BxDFs, textures, NEE, roulette), shadow resolve,
    accumulation; with reverse > 0 the bidirectional mode's light paths, splats and connections.  The libm is the oracle's
    here, so the framebuffer must be the oracle's BIT FOR BIT: whatever the GPU image differs by is CUDA's sinf/cosf."""
    pack, cfg = scene()
    (rgb, cnt, st, bvh), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=7, wide_bvh=wide_bvh, reverse=reverse)
    assert np.array_equal(cnt, co)
    assert int(st.closest_rays) == int(so.closest_rays)
    assert int(st.shadow_rays) + int(st.shadow_rays_skipped) == int(so.shadow_rays)
    assert (bvh[0] > 0) == wide_bvh
    if reverse:
        # splats and per-vertex sums are accumulated in another order than the reference's sequential loop (atomics, one
        # thread per connection): float addition does not associate, a few ulps of the pixel sum remain
        assert (np.abs(rgb - fo) <= 2e-5 * np.maximum(np.abs(fo), float(fo.mean()))).all()
        return
    same = rgb.view(np.uint32) == fo.view(np.uint32)
    assert same.all(), f"{(~same).sum()} of {same.size} framebuffer words differ; max abs diff {np.abs(rgb - fo).max()}"


def test_reference_scene_files_through_the_wavefront_on_the_host(doh, oracle):
    """Every scene file of the reference that loads here (its own JSON configs: OBJ meshes, MTL materials, image and bump
    textures, point / sphere / areal lights, envmap and colour skies, thin-glass flag, unidirectional and bidirectional) at
    32x24x2: the device source on the host against the oracle -- bit for bit where reverse == 0, 2e-5 otherwise.
    Needs /root/reference (this container only)."""
    import glob
    import warnings
    from rgk_b200 import scene
    done = []
    for path in sorted(glob.glob("/root/reference/scenes/*.json")):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pack, cfg = scene.load_json_config(path, overrides={"output-width": 32, "output-height": 24, "multisample": 2})
        except (scene.ConfigFileException, FileNotFoundError):
            continue
        if cfg.recursion_level + cfg.reverse > 59:
            continue
        for wide_bvh in (False, True):
            (rgb, cnt, st, bvh), (fo, co, so) = _host_round(doh, oracle, pack, cfg, wide_bvh=wide_bvh)
            name = (os.path.basename(path), wide_bvh)
            assert np.array_equal(cnt, co), name
            assert int(st.closest_rays) == int(so.closest_rays), name
            if cfg.reverse:
                assert (np.abs(rgb - fo) <= 2e-5 * np.maximum(np.abs(fo), float(fo.mean()) + 1e-6)).all(), name
            else:
                assert np.array_equal(rgb.view(np.uint32), fo.view(np.uint32)), name
        done.append((os.path.basename(path), cfg.reverse))
    if not done:
        pytest.skip("no reference scenes here")
    assert len(done) >= 15


def test_host_loop_chunking_on_the_host(doh, oracle):
    """render_round_impl's chunk loop (rgk_device_cfg::chunk_paths far below the paths of the call: many chunks, tiles split across
    them) gives the same bits as one chunk -- the host loop of render.cu itself, run on the CPU."""
    pack, cfg = _zoo()
    (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=3, wide_bvh=True, chunk_paths=700)
    assert np.array_equal(cnt, co) and np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))
    assert int(st.closest_launches) > cfg.recursion_level                          # more than one chunk was traced


def test_more_chunks_than_pending_counter_blocks_on_the_host(doh, oracle):
    """A call of 21 chunks: the per-chunk counter blocks wait in pinned memory for the end of the round, 16 of them at most
    (H_CHUNKS) -- the 17th chunk makes the host drain them first.  Ray counts (summed from those blocks) and image as the oracle's."""
    pack, cfg = scenes.material_zoo(width=224, height=96, multisample=1, recursion_max=3, lens=0.0)
    (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=1, wide_bvh=True, chunk_paths=64)
    assert np.array_equal(cnt, co) and np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))
    assert int(st.closest_rays) == int(so.closest_rays) and int(st.shadow_rays) + int(st.shadow_rays_skipped) == int(so.shadow_rays)
    assert int(st.closest_launches) >= 21 * 1


@pytest.mark.parametrize("scene", [_zoo, _cornell, _sponza])
def test_ab_knob_kernels_on_the_host(doh, oracle, scene):
    """The A/B candidates that are built but off by default give the same bits: rgk_device_cfg::bvh_shadow_nosort (any-hit BVH
    traversal entering the children in slot order) and bvh_closest_nearest (closest-hit BVH traversal entering the nearest
    child first without sorting the others)."""
    pack, cfg = scene()
    (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=5, wide_bvh=True, bvh_shadow_nosort=1, bvh_closest_nearest=1)
    assert np.array_equal(cnt, co) and np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))
    assert int(st.closest_rays) == int(so.closest_rays)


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("multisample", [4, 9, 121])
def test_device_sampler_in_the_wavefront_on_the_host(doh, oracle, multisample, kernel):
    """RGK_SAMPLER_MT19937: k_sampler_mt (the device replica of StratifiedSampler over libstdc++'s mt19937 / shuffle) inside
    the host-compiled wavefront instead of caller-supplied tables -- still the oracle's framebuffer bit for bit.  Set size
    121 exceeds the shared-memory budget of the thread-per-pixel kernel (1) and takes its in-place global-memory instantiation;
    kernel 2 is the warp-per-pixel builder k_sampler_warp (one-lane warps here: batches of one draw, one table slot), which
    only builds the tables the round reads."""
    pack, cfg = scenes.material_zoo(width=16, height=12, multisample=multisample, recursion_max=3, lens=0.04)
    (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=2, wide_bvh=True, device_sampler=True, sampler_kernel=kernel)
    assert np.array_equal(cnt, co) and np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))


@pytest.mark.parametrize("seed", range(6))
def test_random_material_parameters_on_the_host(doh, oracle, seed):
    """The material zoo with random roughness / ior / mix amounts (including the 0 and 1 ends), with and without lens and
    envmap sky, recursion-max 5, device sampler: the wavefront's device source against the oracle, bit for bit."""
    rng = np.random.default_rng(100 + seed)
    pack, cfg = scenes.material_zoo(width=32, height=20, multisample=4, recursion_max=5, lens=0.04 if seed % 2 else 0.0, envmap_sky=bool(seed % 3))
    for m in pack.materials:
        m["roughness"] = float(np.float32(rng.choice([0.0, 0.01, 0.05, 0.3, 0.7, 1.0]) if rng.random() < 0.5 else rng.uniform(0.0, 1.0)))
        m["ior"] = float(np.float32(rng.uniform(1.0, 2.6)))
        m["amount"] = float(np.float32(rng.uniform(0.0, 1.0)))
    (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=seed, wide_bvh=bool(seed % 2), device_sampler=True)
    assert np.array_equal(cnt, co) and int(st.closest_rays) == int(so.closest_rays)
    assert np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))


@pytest.mark.parametrize("seed", range(6))
def test_random_cameras_on_the_host(doh, oracle, seed):
    """Cameras dropped at random inside the scene's bounding box (often inside or behind geometry, looking anywhere), Cornell
    box / material zoo / cathedral stand-in, BVH path, device sampler, odd seeds also in the bidirectional mode."""
    rng = np.random.default_rng(1000 + seed)
    if seed % 3 == 0:
        pack, cfg = scenes.load_builtin("cornell-box", width=32, height=32, multisample=4, recursion_max=8)
    elif seed % 3 == 1:
        pack, cfg = scenes.material_zoo(width=32, height=20, multisample=4, recursion_max=4)
    else:
        pack, cfg = standin.sibenik(width=32, height=20, multisample=2)
    a = pack.arrays()
    lo, hi = a["positions"].min(0), a["positions"].max(0)
    cfg.camera["position"] = [float(x) for x in (lo + (hi - lo) * rng.uniform(0.05, 0.95, 3)).astype(np.float32)]
    cfg.camera["lookat"] = [float(x) for x in (lo + (hi - lo) * rng.uniform(0.0, 1.0, 3)).astype(np.float32)]
    for rev in ((0, 1) if seed % 2 else (0,)):
        (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, seedcount_base=seed, wide_bvh=True, device_sampler=True, reverse=rev)
        assert np.array_equal(cnt, co) and int(st.closest_rays) == int(so.closest_rays)
        if rev:
            assert (np.abs(rgb - fo) <= 2e-5 * np.maximum(np.abs(fo), float(fo.mean()) + 1e-6)).all()
        else:
            assert np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))


@pytest.mark.parametrize("reverse", [0, 2])
def test_obj_fixture_scene_on_the_host(doh, oracle, reverse):
    """tests/golden/objscene/room.json (OBJ mesh with quads and n-gons, MTL materials, PNG texture, JPEG bump map) through the
    importers and then the host-compiled wavefront -- travels with the repo, unlike the reference's own scene files."""
    import warnings
    from rgk_b200 import scene
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        pack, cfg = scene.load_json_config(os.path.join(ROOT, "tests", "golden", "objscene", "room.json"),
                                           overrides={"output-width": 40, "output-height": 30, "multisample": 4})
    (rgb, cnt, st, _), (fo, co, so) = _host_round(doh, oracle, pack, cfg, wide_bvh=True, device_sampler=True, reverse=reverse)
    assert np.array_equal(cnt, co) and int(st.closest_rays) == int(so.closest_rays)
    if reverse:
        assert (np.abs(rgb - fo) <= 2e-5 * np.maximum(np.abs(fo), float(fo.mean()) + 1e-6)).all()
    else:
        assert np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))
