"""ctypes front-end to the CPU oracle (oracle/astrolink_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs.  The product package
(astrolink_b200) never imports this module.

Each function names the reference step it restates
(/root/reference/astrolink/astrolink.py:<line>).
"""
import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile liboracle.so with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "astrolink_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], env=env,
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_num_threads.restype = ctypes.c_int
        for suf in ("f32", "f64"):
            getattr(_lib, "oracle_aggregate_" + suf).restype = ctypes.c_int64
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _suf(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n=None):
    """Use n host threads (default: all cores) regardless of OMP_NUM_THREADS (torchrun sets it to 1)."""
    n = (os.cpu_count() or 1) if n is None else int(n)
    lib().oracle_set_num_threads(ctypes.c_int(n))
    return num_threads()


def rescale(P):
    """transform_data / _rescale_njit (astrolink.py:218-243)."""
    P = np.ascontiguousarray(P)
    out = np.empty_like(P)
    std = np.empty(P.shape[1], dtype=np.float64)
    getattr(lib(), "oracle_rescale_" + _suf(P.dtype))(
        _p(P), ctypes.c_int64(P.shape[0]), ctypes.c_int(P.shape[1]), _p(out), _p(std))
    return out, std


def knn(P, k, q_begin=0, q_end=None, brute=False):
    """KDTree(P).query(P[q_begin:q_end], k, sqr_dists=True) (astrolink.py:270,279)
    with the defined fma distance chain and (d2, index) ordering."""
    P = np.ascontiguousarray(P)
    n, d = P.shape
    q_end = n if q_end is None else q_end
    nq = q_end - q_begin
    idx = np.empty((nq, k), dtype=np.uint32)
    d2 = np.empty((nq, k), dtype=P.dtype)
    name = ("oracle_knn_brute_" if brute else "oracle_knn_") + _suf(P.dtype)
    rc = getattr(lib(), name)(_p(P), ctypes.c_int64(n), ctypes.c_int(d), ctypes.c_int64(q_begin),
                              ctypes.c_int64(q_end), ctypes.c_int(k), _p(idx), _p(d2))
    if rc != 0:
        raise ValueError("oracle knn: invalid k")
    return d2, idx


def logrho(d2, k_den, d_intrinsic, idx=None, weights=None):
    """_compute_logRho_njit / _compute_weighted_logRho_njit (astrolink.py:293-303);
    result stored in the dtype of d2 as at astrolink.py:282."""
    d2 = np.ascontiguousarray(d2)
    assert d2.shape[1] == k_den
    out = np.empty(d2.shape[0], dtype=d2.dtype)
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        wp, ip = _p(weights), _p(idx)
    else:
        wp, ip = None, None
    getattr(lib(), "oracle_logrho_" + _suf(d2.dtype))(
        _p(d2), ip, wp, ctypes.c_int64(d2.shape[0]), ctypes.c_int(k_den), ctypes.c_int(d_intrinsic), _p(out))
    return out


def normalise(x):
    """_normalise_njit (astrolink.py:305-313)."""
    x = np.array(x, copy=True)
    getattr(lib(), "oracle_normalise_" + _suf(x.dtype))(_p(x), ctypes.c_int64(x.size))
    return x


def make_graph(logRho, kNN, k_link=None):
    """_make_graph_njit (astrolink.py:355-382) -> (pairs [E,2] u32, edges [E])."""
    logRho = np.ascontiguousarray(logRho)
    kNN = np.ascontiguousarray(kNN, dtype=np.uint32)
    n, stride = kNN.shape
    L = stride if k_link is None else k_link
    E = n * (L - 1)
    edges = np.empty(E, dtype=logRho.dtype)
    pairs = np.empty((E, 2), dtype=np.uint32)
    getattr(lib(), "oracle_make_graph_" + _suf(logRho.dtype))(
        _p(logRho), _p(kNN), ctypes.c_int64(n), ctypes.c_int(L), ctypes.c_int(stride), _p(edges), _p(pairs))
    return pairs, edges


def sort_edges(edges):
    """Canonical replacement of ``edges.argsort()[::-1]`` (astrolink.py:341):
    weight descending, slot ascending == np.argsort(-edges, kind='stable')."""
    edges = np.ascontiguousarray(edges)
    order = np.empty(edges.size, dtype=np.int64)
    getattr(lib(), "oracle_sort_edges_" + _suf(edges.dtype))(_p(edges), ctypes.c_int64(edges.size), _p(order))
    return order


def aggregate(logRho, ordered_pairs):
    """_aggregate_njit_float32_uint32 / _float64_uint32 (astrolink.py:384-604),
    variant chosen by logRho.dtype as at astrolink.py:345-350."""
    logRho = np.ascontiguousarray(logRho)
    ordered_pairs = np.ascontiguousarray(ordered_pairs, dtype=np.uint32)
    n = logRho.size
    E = ordered_pairs.shape[0]
    ordering = np.empty(n, dtype=np.uint32)
    groups = np.empty((n // 2 + 2, 3), dtype=np.uint32)
    prom = np.empty((n // 2 + 2, 2), dtype=logRho.dtype)
    G = getattr(lib(), "oracle_aggregate_" + _suf(logRho.dtype))(
        _p(logRho), _p(ordered_pairs), ctypes.c_int64(n), ctypes.c_int64(E), _p(ordering), _p(groups), _p(prom))
    if G < 0:
        raise RuntimeError("oracle aggregate failed (unassigned points)")
    return ordering, groups[:G].copy(), prom[:G].copy()


def hot_path(P, k_den=20, k_link=7, adaptive=1, d_intrinsic=None, weights=None, P_transform=None):
    """The reference's transform_data -> estimate_density_and_kNN -> aggregate
    (astrolink.py:218-353) on the CPU, with per-stage wall times."""
    t = {}
    d_intrinsic = P.shape[1] if d_intrinsic is None else d_intrinsic
    t0 = time.perf_counter()
    if P_transform is None:
        P_transform = rescale(P)[0] if adaptive else np.ascontiguousarray(P)
    t["transform"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    d2, idx = knn(P_transform, k_den)
    t["knn"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    lr_raw = logrho(d2, k_den, d_intrinsic, idx=idx, weights=weights)
    lr = normalise(lr_raw)
    t["density"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    kNN = np.ascontiguousarray(idx[:, :k_link])
    pairs, edges = make_graph(lr, kNN)
    t["graph"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    order = sort_edges(edges)
    ordered_pairs = pairs[order]
    t["sort"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ordering, groups, prominences = aggregate(lr, ordered_pairs)
    t["aggregate"] = time.perf_counter() - t0
    return dict(P_transform=P_transform, d2=d2, idx=idx, logRho_raw=lr_raw, logRho=lr, kNN=kNN,
                pairs=pairs, edges=edges, order=order, ordered_pairs=ordered_pairs,
                ordering=ordering, groups=groups, prominences=prominences, times=t)
