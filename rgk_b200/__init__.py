"""rgk_b200 -- B200 (sm_100a) implementation of RGKrt's path-tracing hot path.

The compute lives in librgk_b200.so (CUDA, C ABI: include/rgk_b200.h).  This
package is the thin host-side binding used by tests, bench.py and tools: scene
packing (scene.py, scenes.py) and a ctypes wrapper (device.py).  No CPU fallback.
"""
from . import abi  # noqa: F401
