"""Extract the LTC fit tables (GGX, Beckmann) from the reference build into
rgk_b200/data/ltc_tables.npz.

The tables are numeric data the hot path needs at run time (src/LTC/ltc_ggx.cpp,
src/LTC/ltc_beckmann.cpp: tabM as double[9] per entry, tabAmplitude as float).  They
are read through oracle/_ref (the reference compiled here) with the double->float
cast of src/LTC/ltc.hpp:6-9, so the file holds exactly what the reference's
mat33::operator glm::mat3 produces.  Run in the build container only.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librgk_ref.so"))
out = {}
for which, name in ((0, "ggx"), (1, "beckmann")):
    M = np.zeros((4096, 9), np.float32)
    amp = np.zeros(4096, np.float32)
    size = lib.rgkref_ltc_tables(which, M.ctypes.data_as(C.c_void_p), amp.ctypes.data_as(C.c_void_p))
    assert size == 64
    out[name + "_M"], out[name + "_amp"] = M, amp
path = os.path.join(ROOT, "rgk_b200", "data", "ltc_tables.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: (v.shape, float(v.sum())) for k, v in out.items()})
