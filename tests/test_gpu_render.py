"""Parity of the CUDA bounce loop against the CPU oracle, through the C ABI (rgk_render_round)."""
import numpy as np
import pytest

from rgk_b200 import abi, scenes

pytestmark = pytest.mark.gpu


def test_sampler_tables_bit_exact(gpu_ctx, oracle):
    """The device mt19937 / shuffle replica reproduces StratifiedSampler's tables bit for bit."""
    seeds = np.array([42 + 0x42424242, 7, 0, 0xFFFFFFFF, 123456789], np.uint32)
    for ms in (1, 4, 16, 40, 64):                     # 40 -> set_size 49 (odd), others even
        t1g, t2g = gpu_ctx.sampler_tables(seeds, ms, 64, 64)
        t1o, t2o = oracle.sampler_tables(seeds, ms, 64, 64)
        assert gpu_ctx.sampler_set_size(ms) == oracle.sampler_set_size(ms)
        assert np.array_equal(t1g.view(np.uint32), t1o.view(np.uint32)), ms
        assert np.array_equal(t2g.view(np.uint32), t2o.view(np.uint32)), ms


def test_cornell_image_same_sample_sequence(gpu_ctx, oracle):
    """BASELINE configs[0] (Cornell box, 256x256, 16 spp, recursion-max 40) with the SAME sample sequence as the
    CPU path.  Trajectories differ only where CUDA's sinf/cosf differ from glibc in the last ulp, so the images
    agree far below Monte-Carlo noise: rel-mean <= 1e-3, RMSE <= 2% of the mean (stated tolerance)."""
    pack, cfg = scenes.load_builtin("cornell-box")
    desc = pack.desc()
    gpu_ctx.commit(desc)
    ho = oracle.scene_create(desc)
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params(abi.SAMPLER_MT19937)
    tasks = gpu_ctx.generate_tasks(32, p.xres, p.yres)
    fg, cg, sg = gpu_ctx.render_round(cam, p, tasks)
    fo, co, so = oracle.render_round(ho, cam, p, tasks)
    assert np.array_equal(cg, co) and int(cg.min()) == 16
    ig, io = fg / 16.0, fo / 16.0
    mean = float(io.mean())
    rel_mean = abs(float(ig.mean()) - mean) / mean
    rmse = float(np.sqrt(np.mean((ig - io) ** 2)))
    frac_equal = float(np.mean(ig == io))
    print(f"cornell 256x256x16: rel_mean={rel_mean:.3e} rmse/mean={rmse / mean:.3e} pixels bit-equal={frac_equal:.4f} "
          f"rays gpu={sg.closest_rays}/{sg.shadow_rays} cpu={so.closest_rays}/{so.shadow_rays}")
    assert rel_mean <= 1e-3
    assert rmse <= 0.02 * mean
    assert abs(int(sg.closest_rays) - int(so.closest_rays)) <= 0.001 * so.closest_rays
    assert frac_equal > 0.5


def test_cornell_fast_sampler_statistics(gpu_ctx, oracle):
    """RGK_SAMPLER_FAST draws a different sequence with the same distribution: compare at equal spp against the
    oracle's own seed-to-seed difference (rel-mean <= 0.5%, RMSE <= 1.5x the oracle's seed-to-seed RMSE)."""
    pack, cfg = scenes.load_builtin("cornell-box", width=128, height=128, multisample=16)
    desc = pack.desc()
    gpu_ctx.commit(desc)
    ho = oracle.scene_create(desc)
    cam = gpu_ctx.camera(**cfg.camera_args())
    tasks = gpu_ctx.generate_tasks(32, 128, 128)
    fo1, _, _ = oracle.render_round(ho, cam, cfg.params(), tasks, seedstart=42)
    fo2, _, _ = oracle.render_round(ho, cam, cfg.params(), tasks, seedstart=4242)
    fg, _, _ = gpu_ctx.render_round(cam, cfg.params(abi.SAMPLER_FAST), tasks)
    mean = float(fo1.mean())
    self_rmse = float(np.sqrt(np.mean((fo1 - fo2) ** 2)))
    rmse = float(np.sqrt(np.mean((fg - fo1) ** 2)))
    rel_mean = abs(float(fg.mean()) - mean) / mean
    print(f"fast sampler: rel_mean={rel_mean:.3e} rmse={rmse:.4f} oracle seed-to-seed rmse={self_rmse:.4f}")
    assert rel_mean <= 5e-3
    assert rmse <= 1.5 * self_rmse


def test_round_accumulates_and_seeds_advance(gpu_ctx, oracle):
    pack, cfg = scenes.load_builtin("cornell-box", width=64, height=48, multisample=4, recursion_max=5)
    desc = pack.desc()
    gpu_ctx.commit(desc)
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params()
    tasks = gpu_ctx.generate_tasks(32, 64, 48)     # ragged tiles: 64x48 -> 2x2 tiles, bottom row 16 high
    f1, c1, _ = gpu_ctx.render_round(cam, p, tasks, seedcount_base=0)
    f2, c2, _ = gpu_ctx.render_round(cam, p, tasks, seedcount_base=len(tasks), fb=(f1.copy(), c1.copy()))
    assert int(c2.min()) == 8 and int(c2.max()) == 8
    ff, cf, _ = gpu_ctx.render_frame(cam, p, rounds=2)   # RenderFrame's loop == two rounds with running seedcount
    assert np.array_equal(cf, c2) and np.array_equal(ff, f2)
    ho = oracle.scene_create(desc)
    fo, co, _ = oracle.render_round(ho, cam, p, tasks)
    fo, co, _ = oracle.render_round(ho, cam, p, tasks, seedcount_base=len(tasks), fb=(fo, co))
    assert abs(float(ff.mean()) - float(fo.mean())) <= 2e-3 * float(fo.mean())


def test_error_paths(gpu_ctx):
    from rgk_b200.device import RgkError
    pack, cfg = scenes.load_builtin("cornell-box", width=32, height=32, multisample=1)
    gpu_ctx.commit(pack.desc())
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params()
    p.depth = 61                                   # leaves the 64 tabulated sampler dimensions
    with pytest.raises(RgkError) as e:
        gpu_ctx.render_round(cam, p, gpu_ctx.generate_tasks(32, 32, 32))
    assert e.value.status == 6
    p = cfg.params()
    p.reverse = 17
    with pytest.raises(RgkError) as e:
        gpu_ctx.render_round(cam, p, gpu_ctx.generate_tasks(32, 32, 32))
    assert e.value.status == 6


@pytest.mark.parametrize("scene_name,reverse,depth", [("cornell", 3, 6), ("cornell", 1, 40), ("zoo", 2, 5)])
def test_bidirectional_round_matches_oracle(gpu_ctx, oracle, scene_name, reverse, depth):
    """reverse > 0 (src/path_tracer.cpp:336-398,462-480): light path, camera splats (count 0), vertex connections.
    Same sample sequence as the CPU; splats are summed with atomics, so the comparison is by tolerance."""
    if scene_name == "cornell":
        pack, cfg = scenes.load_builtin("cornell-box", width=64, height=64, multisample=4)
    else:
        pack, cfg = scenes.material_zoo(width=64, height=40, multisample=4, lens=0.03)
    desc = pack.desc()
    gpu_ctx.commit(desc)
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params()
    p.depth, p.reverse = depth, reverse
    tasks = gpu_ctx.generate_tasks(32, p.xres, p.yres)
    fb, cnt, st = gpu_ctx.render_round(cam, p, tasks)
    ho = oracle.scene_create(desc)
    fo, co, so = oracle.render_round(ho, cam, p, tasks, nthreads=1)
    assert np.array_equal(cnt, co) and np.all(cnt == 4)                 # splats add radiance, never samples
    assert int(st.closest_rays) == int(so.closest_rays)                 # camera + light path segments
    assert int(st.shadow_rays) + int(st.shadow_rays_skipped) == int(so.shadow_rays)   # NEE + both kinds of connection rays
    mean = float(fo.mean())
    assert abs(float(fb.mean()) - mean) / mean < 1e-3
    assert float(np.sqrt(np.mean((fb - fo) ** 2))) / mean < 0.05
    # and the unidirectional image differs from it (the mode is doing something)
    p.reverse = 0
    f0, _, _ = gpu_ctx.render_round(cam, p, tasks)
    assert float(np.abs(f0 - fb).mean()) > 1e-3 * mean
    oracle.scene_destroy(ho)


def _pixel_seeds(tasks, seedstart=42, seedcount_base=0):
    """Seeds of the pixels of a call in task order, y-major/x-minor (src/render_driver.cpp:160,173, src/path_tracer.cpp:47)."""
    out = []
    for i, t in enumerate(tasks):
        n = (t.x2 - t.x1) * (t.y2 - t.y1)
        base = (seedstart + seedcount_base + i) & 0xFFFFFFFF
        out.append((base + (np.arange(1, n + 1, dtype=np.uint64) * 0x42424242)) & 0xFFFFFFFF)
    return np.concatenate(out).astype(np.uint32)


def test_caller_supplied_tables_equal_device_sampler(gpu_ctx, oracle):
    """RGK_SAMPLER_TABLES fed with the oracle's StratifiedSampler tables gives the image of RGK_SAMPLER_MT19937 bit for
    bit (lens camera -> one more 2-D dim; multisample 9 -> odd set size; 40x24 -> ragged tiles)."""
    pack, cfg = scenes.material_zoo(width=40, height=24, multisample=9, recursion_max=3, lens=0.04)
    desc = pack.desc()
    gpu_ctx.commit(desc)
    cam = gpu_ctx.camera(**cfg.camera_args())
    tasks = gpu_ctx.generate_tasks(32, 40, 24)
    f_mt, c_mt, _ = gpu_ctx.render_round(cam, cfg.params(abi.SAMPLER_MT19937), tasks, seedcount_base=7)
    n1d, n2d = 1 + 3, 5 + 3
    t1, t2 = oracle.sampler_tables(_pixel_seeds(tasks, 42, 7), 9, n1d, n2d)
    gpu_ctx.set_tables(9, t1, t2)
    f_tb, c_tb, _ = gpu_ctx.render_round(cam, cfg.params(abi.SAMPLER_TABLES), tasks, seedcount_base=7)
    assert np.array_equal(f_mt.view(np.uint32), f_tb.view(np.uint32)) and np.array_equal(c_mt, c_tb)
    ho = oracle.scene_create(desc)
    fo, _, _ = oracle.render_round(ho, cam, cfg.params(), tasks, seedcount_base=7)
    assert abs(float(f_mt.mean()) - float(fo.mean())) <= 2e-3 * float(fo.mean())


def test_tile_sharding_is_independent_of_n(gpu_ctx):
    """rgk_render_set_shard: the union of N tile-sharded renders equals the unsharded round pixel for pixel."""
    pack, cfg = scenes.load_builtin("cornell-box", width=96, height=80, multisample=4, recursion_max=5)
    gpu_ctx.commit(pack.desc())
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params()
    tasks = gpu_ctx.generate_tasks(32, 96, 80)
    full, cfull, _ = gpu_ctx.render_round(cam, p, tasks, seedcount_base=3)
    try:
        for n in (2, 3):
            fb = (np.zeros_like(full), np.zeros_like(cfull))
            for r in range(n):
                gpu_ctx.set_shard(r, n)
                gpu_ctx.render_round(cam, p, tasks, seedcount_base=3, fb=fb)
            assert np.array_equal(fb[0].view(np.uint32), full.view(np.uint32)) and np.array_equal(fb[1], cfull)
    finally:
        gpu_ctx.set_shard(0, 1)


def test_chunking_does_not_change_the_image(gpu_ctx, monkeypatch):
    """Tiny chunks (rgk_device_cfg::chunk_paths) split the tile list into many passes; the result must be identical."""
    pack, cfg = scenes.load_builtin("cornell-box", width=96, height=64, multisample=4, recursion_max=4)
    gpu_ctx.commit(pack.desc())
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params()
    tasks = gpu_ctx.generate_tasks(32, 96, 64)
    a, ca, sa = gpu_ctx.render_round(cam, p, tasks)
    base = gpu_ctx.cfg()
    gpu_ctx.configure(chunk_paths=5000)
    b, cb, sb = gpu_ctx.render_round(cam, p, tasks)
    gpu_ctx.configure(base)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)
    assert int(sb.kernel_launches) > int(sa.kernel_launches)
    # more chunks than the 16 counter blocks that can wait in pinned memory for the end of the round: the host drains in between
    pack2, cfg2 = scenes.load_builtin("cornell-box", width=224, height=96, multisample=1, recursion_max=4)
    cam2 = gpu_ctx.camera(**cfg2.camera_args())
    tasks2 = gpu_ctx.generate_tasks(32, 224, 96)
    c, cc, sc = gpu_ctx.render_round(cam2, cfg2.params(), tasks2)
    gpu_ctx.configure(chunk_paths=64)
    d, cd, sd = gpu_ctx.render_round(cam2, cfg2.params(), tasks2)
    gpu_ctx.configure(base)
    assert np.array_equal(c.view(np.uint32), d.view(np.uint32)) and np.array_equal(cc, cd)
    assert (int(sc.closest_rays), int(sc.shadow_rays), int(sc.shadow_rays_skipped)) == (int(sd.closest_rays), int(sd.shadow_rays), int(sd.shadow_rays_skipped))
    assert int(sd.closest_launches) >= 21
    empty = (abi.Task * 0)()
    z, cz, sz = gpu_ctx.render_round(cam, p, empty)
    assert not z.any() and int(sz.samples) == 0


def test_two_contexts_interleaved(gpu_ctx):
    """Two contexts in one process (what a multi-GPU host would hold, here on one device) with different scenes,
    called alternately: no state leaks between them."""
    from rgk_b200 import device
    pa, ca = scenes.load_builtin("cornell-box", width=64, height=48, multisample=4)
    pb, cb = scenes.material_zoo(width=48, height=32, multisample=4)
    other = device.Context(0)
    try:
        gpu_ctx.commit(pa.desc()); other.commit(pb.desc())
        cam_a, cam_b = gpu_ctx.camera(**ca.camera_args()), other.camera(**cb.camera_args())
        ta, tb = gpu_ctx.generate_tasks(32, 64, 48), other.generate_tasks(32, 48, 32)
        a1, _, _ = gpu_ctx.render_round(cam_a, ca.params(), ta)
        b1, _, _ = other.render_round(cam_b, cb.params(), tb)
        a2, _, _ = gpu_ctx.render_round(cam_a, ca.params(), ta)
        b2, _, _ = other.render_round(cam_b, cb.params(), tb)
        assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32)) and np.array_equal(b1.view(np.uint32), b2.view(np.uint32))
        assert a1.shape != b1.shape and other.scene_info().n_triangles != gpu_ctx.scene_info().n_triangles
    finally:
        other.close()
