"""k_sampler_warp -- the warp-per-pixel builder of the mt19937 sampler tables (src/sampler.cpp:85-116) -- on the CPU with
32-LANE warps: tests/host_cpp/device_shim_mt.h runs every CUDA thread as a host thread, __syncwarp / __syncthreads as
pthread barriers and the warp collectives (ballot, shuffle, match, reduce) through a per-warp exchange, so the code paths the
one-lane emulation of test_device_on_host.py cannot reach are compared with the oracle here, without a GPU: 128-word twist
steps across the lanes, batches of 32 draws, a Lemire rejection ending a batch at a lane, shuffles above 64 entries applied
in waves of independent exchanges, table slots shared by several pixels, the CTA-wide copy-out, CTAs whose last pixel row is
partly empty.  tools/tsan_lanes_on_host.sh runs the same build under ThreadSanitizer (a missing barrier between two lanes'
shared-memory accesses is a reported race; checked with a barrier removed) and AddressSanitizer + UBSan."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "host", "libsampler_mt.so")


@pytest.fixture(scope="module")
def lanes32():
    if not os.path.exists(SO):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(SO)
    lib.doh32_sampler_tables.restype = C.c_int
    return lib


def _tables(lib, oracle, seeds, ms, n1d, n2d, kernel=2, slots=0):
    ss, n = oracle.sampler_set_size(ms), len(seeds)
    o1 = np.zeros((n1d, ss, n), np.float32)
    o2 = np.zeros((n2d, ss, n, 2), np.float32)
    err = C.c_char_p()
    st = lib.doh32_sampler_tables(seeds.ctypes.data_as(C.c_void_p), n, ms, n1d, n2d, o1.ctypes.data_as(C.c_void_p), o2.ctypes.data_as(C.c_void_p),
                                  kernel, slots, C.byref(err))
    assert st == 0, err.value
    return np.ascontiguousarray(o1.transpose(2, 0, 1)), np.ascontiguousarray(o2.transpose(2, 0, 1, 3))


def _seeds(n, salt=0):
    rng = np.random.default_rng(777 + salt)
    s = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    s[:3] = (42 + 0x42424242, 0, 0xFFFFFFFF)
    return s


@pytest.mark.parametrize("ms,n,n1d,n2d,slots", [
    (64, 43, 3, 6, 0), (64, 16, 2, 3, 1), (64, 21, 4, 4, 5),      # the headline's set size: slot counts, partly empty last rows
    (1, 40, 6, 6, 0), (4, 37, 5, 6, 0), (36, 19, 3, 4, 3),        # batches shorter than a warp
    (40, 12, 3, 3, 0), (121, 11, 2, 3, 0),                        # odd set sizes (49, 121): no single first exchange
    (256, 10, 2, 2, 0), (400, 9, 1, 2, 2),                        # waves of independent exchanges; several generations per table
])
def test_warp_builder_tables_with_32_lane_warps(lanes32, oracle, ms, n, n1d, n2d, slots):
    seeds = _seeds(n, ms)
    t1, t2 = _tables(lanes32, oracle, seeds, ms, n1d, n2d, kernel=2, slots=slots)
    a1, a2 = oracle.sampler_tables(seeds, ms, n1d, n2d)
    assert np.array_equal(t1.view(np.uint32), a1.view(np.uint32))
    assert np.array_equal(t2.view(np.uint32), a2.view(np.uint32))


def test_lemire_rejection_inside_a_batch(lanes32, oracle):
    """Set size 1024: a shuffle draw is rejected with probability up to 2.4e-4; 6 pixels x 6 shuffles of 512 draws see a few,
    each ending its batch at the rejecting lane and shifting the rest of the pixel's stream."""
    seeds = _seeds(6, 5)
    t1, t2 = _tables(lanes32, oracle, seeds, 1024, 3, 3)
    a1, a2 = oracle.sampler_tables(seeds, 1024, 3, 3)
    assert np.array_equal(t1.view(np.uint32), a1.view(np.uint32)) and np.array_equal(t2.view(np.uint32), a2.view(np.uint32))


def test_thread_builder_with_32_lane_warps(lanes32, oracle):
    """k_sampler_mt (thread per pixel, the builder of set sizes <= 16) under the same emulation, with a partly filled last warp:
    its lanes past the end of the chunk go through the warp barriers of the shared-memory tables with a pixel they drop."""
    seeds = _seeds(150, 9)
    t1, t2 = _tables(lanes32, oracle, seeds, 16, 3, 4, kernel=1)
    a1, a2 = oracle.sampler_tables(seeds, 16, 3, 4)
    assert np.array_equal(t1.view(np.uint32), a1.view(np.uint32)) and np.array_equal(t2.view(np.uint32), a2.view(np.uint32))
