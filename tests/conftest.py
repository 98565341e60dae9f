import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # RGK_TEST_TRAVERSAL=bvh runs the whole GPU suite with every scene committed under RGK_WIDE_BVH=1 (the wide-BVH
    # candidate pass + kd-tree arbiter): the gate for making that path the library default.  The tests that set or clear
    # RGK_WIDE_BVH themselves (test_gpu_bvh.py, test_bvh_host.py) are unaffected.
    if os.environ.get("RGK_TEST_TRAVERSAL") == "bvh":
        os.environ["RGK_WIDE_BVH"] = "1"


@pytest.fixture(scope="session")
def oracle():
    import checkers
    return checkers.oracle()


@pytest.fixture(scope="session")
def ref():
    import checkers
    if not checkers.have_ref():
        pytest.skip("oracle/_ref/librgk_ref.so not built (needs /root/reference)")
    return checkers.ref()


@pytest.fixture(scope="session")
def cornell():
    from rgk_b200 import scenes
    pack, cfg = scenes.load_builtin("cornell-box")
    return pack, cfg, pack.desc()


@pytest.fixture(scope="session")
def gpu_ctx():
    from rgk_b200 import device
    ctx = device.Context(0)
    yield ctx
    ctx.close()


def pixel_grid(w, h):
    ys, xs = np.mgrid[0:h, 0:w]
    return np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)
