"""GPU parity of the two table builders of the mt19937 sampler replica (src/sampler.cpp:85-116): the warp-per-pixel kernel
k_sampler_warp (rgk_device_cfg::sampler_kernel 0 / 2: generator state in shared memory, 32 draws per step, only the tables
a round reads) and the thread-per-pixel kernel k_sampler_mt (1).  Tables against the oracle bit for bit at every set size
class (1, even, odd, larger than a warp, larger than a generation of 624 draws, with 1 .. 7 table slots per warp, pixel
counts that leave the last CTA / warp partly empty), and whole rounds under one kernel against the other."""
import numpy as np
import pytest

from rgk_b200 import abi, device, scenes

pytestmark = pytest.mark.gpu


def _seeds(n):
    rng = np.random.default_rng(20261018)
    s = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    s[:4] = (42 + 0x42424242, 0, 0xFFFFFFFF, 7)
    return s


@pytest.mark.parametrize("kernel,slots", [(1, 0), (2, 0), (2, 1), (2, 3), (2, 7)])
@pytest.mark.parametrize("ms", [1, 2, 4, 16, 40, 64, 121, 256, 400])
def test_tables_bit_exact_under_each_kernel(oracle, kernel, slots, ms):
    """40 -> 49 and 121 are odd set sizes (no single first swap), 400 draws 1600 words per dimension (several generations
    inside one table), 300 seeds leave the second CTA of the warp kernel with 44 pixels (5 and 6 per warp)."""
    n1d, n2d = (64, 64) if ms <= 64 else (5, 6)
    seeds = _seeds(300 if ms <= 64 else 41)
    ctx = device.Context(0, sampler_kernel=kernel, sampler_slots=slots)
    try:
        t1g, t2g = ctx.sampler_tables(seeds, ms, n1d, n2d)
    finally:
        ctx.close()
    t1o, t2o = oracle.sampler_tables(seeds, ms, n1d, n2d)
    assert np.array_equal(t1g.view(np.uint32), t1o.view(np.uint32))
    assert np.array_equal(t2g.view(np.uint32), t2o.view(np.uint32))


def test_lemire_rejections_are_replayed(oracle):
    """At set size 1024 a shuffle draw is rejected with probability up to 2.4e-4 (range / 2^32): over 40 pixels x 16 shuffles of
    512 draws about 30 rejections, each of which shifts the rest of the pixel's stream by one draw."""
    seeds = _seeds(40)
    ctx = device.Context(0, sampler_kernel=2)
    try:
        t1g, t2g = ctx.sampler_tables(seeds, 1024, 8, 8)
    finally:
        ctx.close()
    t1o, t2o = oracle.sampler_tables(seeds, 1024, 8, 8)
    assert np.array_equal(t1g.view(np.uint32), t1o.view(np.uint32))
    assert np.array_equal(t2g.view(np.uint32), t2o.view(np.uint32))


@pytest.mark.parametrize("depth,lens,ms", [(1, 0.0, 4), (2, 0.0, 64), (5, 0.04, 16), (3, 0.04, 121)])
def test_round_is_independent_of_the_table_builder(depth, lens, ms):
    """A round reads only some tables (the keep masks of render_round_impl: pixel jitter, lens, light samples unless the light
    is one fixed point light, one direction per vertex but the last, the Russian-roulette cursor); the warp kernel builds
    only those, the thread-per-pixel kernel all of them: same framebuffer, same ray counts."""
    pack, cfg = scenes.material_zoo(width=72, height=40, multisample=ms, recursion_max=depth, lens=lens)
    out = []
    for kernel in (1, 2):
        ctx = device.Context(0, sampler_kernel=kernel)
        try:
            ctx.commit(pack.desc())
            cam = ctx.camera(**cfg.camera_args())
            tasks = ctx.generate_tasks(32, cfg.xres, cfg.yres)
            out.append(ctx.render_round(cam, cfg.params(abi.SAMPLER_MT19937), tasks, seedcount_base=5))
        finally:
            ctx.close()
    (a, ca, sa), (b, cb, sb) = out
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)
    assert int(sa.closest_rays) == int(sb.closest_rays) and int(sa.shadow_rays) == int(sb.shadow_rays)


def test_round_with_one_fixed_point_light():
    """The headline's light set-up (one point light of size 0: const_light, no light tables kept), Cornell-sized."""
    from rgk_b200 import standin
    pack, cfg = standin.sponza(width=96, height=64, multisample=16)
    out = []
    for kernel in (1, 2):
        ctx = device.Context(0, sampler_kernel=kernel)
        try:
            ctx.commit(pack.desc())
            cam = ctx.camera(**cfg.camera_args())
            tasks = ctx.generate_tasks(32, cfg.xres, cfg.yres)
            out.append(ctx.render_round(cam, cfg.params(abi.SAMPLER_MT19937), tasks))
        finally:
            ctx.close()
    (a, ca, sa), (b, cb, sb) = out
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(ca, cb)
    assert int(sa.closest_rays) == int(sb.closest_rays) and int(sa.shadow_rays) == int(sb.shadow_rays)
