"""The C-ABI library loads, exports every symbol include/rgk_b200.h declares, and fails loudly (never falls back
to a CPU path) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from rgk_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    header = open(os.path.join(ROOT, "include", "rgk_b200.h")).read()
    declared = set(re.findall(r"\b(rgk_[a-z0-9_]+)\s*\(", header))
    declared -= {"rgk_status"}
    assert declared == set(abi.EXPORTS), declared ^ set(abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rgk_abi_version() == abi.ABI_VERSION
    assert lib.rgk_status_string(4).decode().startswith("no CUDA device")


def test_struct_layouts_match_header():
    assert C.sizeof(abi.Material) == 64 and C.sizeof(abi.Ray) == 32 and C.sizeof(abi.Hit) == 20
    assert C.sizeof(abi.Mesh) == 16 and C.sizeof(abi.PointLight) == 32 and C.sizeof(abi.Task) == 16
    assert C.sizeof(abi.Camera) == 8 * 12 + 12 and C.sizeof(abi.TravStats) == 64


def test_host_only_entry_points_work_without_gpu():
    lib = abi.load_library()
    assert lib.rgk_sampler_set_size(16) == 16 and lib.rgk_sampler_set_size(512) == 529 and lib.rgk_sampler_set_size(40) == 49
    n = lib.rgk_generate_tasks(32, 1920, 1080, None, 0)
    assert n == 60 * 34


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rgk_b200 import device
    with pytest.raises(device.RgkError) as e:
        device.Context(0)
    assert e.value.status == 4 and "no CPU fallback" in str(e.value)


def test_missing_library_is_an_error(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load_library(str(tmp_path / "nope.so"))
