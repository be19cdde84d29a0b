"""TEST INFRASTRUCTURE ONLY: ctypes front-end of the serial host executor of the
aggregation algorithm (tests/hostsim/agg_hostsim.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libagg_hostsim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "agg_hostsim.cpp")
        hdr = os.path.join(_HERE, "..", "..", "astrolink_b200", "csrc", "aggregate_impl.h")
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _LIB, src])
        _lib = ctypes.CDLL(_LIB)
        for s in ("f32", "f64"):
            getattr(_lib, "hostsim_aggregate_" + s).restype = ctypes.c_int64
    return _lib


def aggregate(logRho, ordered_pairs):
    logRho = np.ascontiguousarray(logRho)
    ordered_pairs = np.ascontiguousarray(ordered_pairs, dtype=np.uint32)
    n, E = logRho.size, ordered_pairs.shape[0]
    ordering = np.empty(n, dtype=np.uint32)
    groups = np.empty((n // 2 + 2, 3), dtype=np.uint32)
    prom = np.empty((n // 2 + 2, 2), dtype=logRho.dtype)
    suf = "f32" if logRho.dtype == np.float32 else "f64"
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    G = getattr(lib(), "hostsim_aggregate_" + suf)(p(logRho), p(ordered_pairs), ctypes.c_int64(n), ctypes.c_int64(E),
                                                   p(ordering), p(groups), p(prom))
    if G < 0:
        raise RuntimeError(f"hostsim aggregate failed: {G}")
    return ordering, groups[:G].copy(), prom[:G].copy()
