"""Whole rounds of the wavefront -- render.cu's kernels and host loop -- on the CPU with 32-LANE warps (tests/host_cpp/
device_shim_mt.h: CUDA threads are host threads, __syncwarp / __syncthreads are pthread barriers, ballots / shuffles go through a
per-warp exchange, atomics are atomics), at the library's own scheduling defaults: k_bin's shared-memory counting sort and the
direction-binned queues, the ballots of the queue compaction, the persistent traversal refilling at 24 / 12 idle lanes, the
warp-per-pixel sampler, the device-side queue lengths and binning decision of the host-free chunk.  The framebuffer must be the
oracle's bit for bit (the order of a queue never changes a result).  The one-lane build of test_device_on_host.py cannot run any
of that cooperative code; compute-sanitizer is closed on the GPU pool."""
import ctypes as C
import os

import numpy as np
import pytest

from rgk_b200 import scenes, standin
from test_device_on_host import _host_round, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "host", "libdevice_on_host_mt.so")
DEFAULTS = dict(binning=1, refill_coherent=0, refill_incoherent=24, refill_shadow=12)      # rgk_device_cfg_init's values


@pytest.fixture(scope="module")
def doh32():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(SO)
    lib.doh_shade_scene_create.restype = vp
    lib.doh_shade_scene_create.argtypes = [vp, vp]
    lib.doh_shade_scene_destroy.argtypes = [vp]
    lib.doh_render_round.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, C.c_uint32, C.c_uint64, vp, vp, vp, vp]
    return lib


@pytest.mark.parametrize("name,wide_bvh", [("zoo", True), ("zoo", False), ("cornell", True), ("sponza", True)])
def test_rounds_with_32_lane_warps(doh32, oracle, name, wide_bvh):
    if name == "zoo":
        pack, cfg = scenes.material_zoo(width=40, height=24, multisample=4, recursion_max=3, lens=0.04)
    elif name == "cornell":
        pack, cfg = scenes.load_builtin("cornell-box", width=32, height=32, multisample=4, recursion_max=12)      # Russian roulette, deep paths: early read-backs
    else:
        pack, cfg = standin.sponza(width=64, height=36, multisample=4)
    (rgb, cnt, st, bvh), (fo, co, so) = _host_round(doh32, oracle, pack, cfg, seedcount_base=3, wide_bvh=wide_bvh, device_sampler=True, **DEFAULTS)
    assert np.array_equal(cnt, co) and int(st.closest_rays) == int(so.closest_rays)
    assert np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))
    if wide_bvh:
        assert int(bvh[0]) > 0


def test_binned_and_compacted_bounces_with_32_lane_warps(doh32, oracle):
    """bin_min_frac 0.6: the first bounces are binned (k_bin), later ones fall below the threshold and are compacted by k_shade's
    atomics -- the decision is taken on the device from the queue length; 36 spp takes the warp-per-pixel sampler at a set size that
    is not a multiple of the warp."""
    pack, cfg = scenes.material_zoo(width=24, height=16, multisample=36, recursion_max=4, lens=0.0)
    fields = dict(DEFAULTS); fields.update(bin_min_frac=0.6, bin_items=512)
    (rgb, cnt, st, bvh), (fo, co, so) = _host_round(doh32, oracle, pack, cfg, seedcount_base=1, wide_bvh=True, device_sampler=True, **fields)
    assert np.array_equal(cnt, co) and int(st.closest_rays) == int(so.closest_rays) and int(st.shadow_rays) + int(st.shadow_rays_skipped) == int(so.shadow_rays)
    assert np.array_equal(rgb.view(np.uint32), fo.view(np.uint32))


def test_bidirectional_round_with_32_lane_warps(doh32, oracle):
    """reverse = 2 on the Cornell box: the light paths, the camera connections (atomic splats: another order of summation than
    the oracle's, hence the tolerance of test_device_on_host's bidirectional test) and the vertex connections with real warps."""
    pack, cfg = scenes.load_builtin("cornell-box", width=24, height=24, multisample=4, recursion_max=3)
    (rgb, cnt, st, bvh), (fo, co, so) = _host_round(doh32, oracle, pack, cfg, seedcount_base=2, wide_bvh=True, reverse=2, **DEFAULTS)
    assert int(st.closest_rays) == int(so.closest_rays) and int(st.shadow_rays) + int(st.shadow_rays_skipped) == int(so.shadow_rays)
    assert np.array_equal(cnt, co)
    assert np.allclose(rgb, fo, rtol=2e-5, atol=2e-5 * float(np.abs(fo).max()))
