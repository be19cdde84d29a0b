"""ctypes binding of libastrolink_b200.so (the C ABI in include/astrolink_b200.h)
plus thin torch-tensor wrappers.  PyTorch only owns device memory and streams.

There is no CPU fallback: if the library is missing or no CUDA device is
present, every entry point raises.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("AL_LIB") or os.path.join(_HERE, "libastrolink_b200.so")   # AL_LIB: development A/B builds
_lib = None

AL_F32, AL_F64 = 0, 1
LEAF = 32

_c = ctypes
_vp, _i64, _int, _sz = _c.c_void_p, _c.c_int64, _c.c_int, _c.c_size_t

_SIGS = {
    "al_version": (_int, []),
    "al_last_error": (_c.c_char_p, []),
    "al_rescale_workspace_bytes": (_sz, [_i64, _int]),
    "al_rescale": (_int, [_vp, _int, _i64, _int, _i64, _vp, _vp, _vp, _sz, _vp]),
    "al_index_bytes": (_sz, [_int, _i64, _int]),
    "al_index_build_workspace_bytes": (_sz, [_int, _i64, _int]),
    "al_index_build": (_int, [_vp, _int, _i64, _int, _vp, _sz, _vp, _sz, _vp]),
    "al_index_positions": (_i64, [_i64]),
    "al_index_perm": (_vp, [_vp]),
    "al_index_release": (None, [_vp]),
    "al_knn_workspace_bytes": (_sz, [_int, _i64, _int, _int]),
    "al_knn": (_int, [_vp, _i64, _i64, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "al_logrho": (_int, [_vp, _vp, _vp, _int, _i64, _int, _int, _vp, _vp]),
    "al_normalise_workspace_bytes": (_sz, [_i64]),
    "al_minmax": (_int, [_vp, _int, _i64, _vp, _vp, _sz, _vp]),
    "al_normalise": (_int, [_vp, _int, _i64, _vp, _vp]),
    "al_make_edges": (_int, [_vp, _int, _vp, _i64, _i64, _int, _vp, _vp, _vp]),
    "al_sort_edges_workspace_bytes": (_sz, [_int, _i64]),
    "al_sort_edges": (_int, [_vp, _int, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "al_aggregate_workspace_bytes": (_sz, [_int, _i64, _i64]),
    "al_aggregate": (_int, [_vp, _int, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "al_scatter_rows": (_int, [_vp, _vp, _i64, _int, _int, _int, _int, _vp, _vp]),
    "al_fp32_peak_tflops": (_int, [_vp, _vp, _sz, _vp]),
    "al_launch_count": (_i64, []),
    "al_peer_alloc": (_int, [_sz, _vp, _vp]),
    "al_peer_open": (_int, [_vp, _vp]),
    "al_peer_release": (_int, [_vp, _int]),
    "al_peer_copy": (_int, [_vp, _vp, _sz, _vp]),
    "al_knn_link": (_int, [_vp, _i64, _i64, _int, _int, _vp, _vp, _vp, _int, _vp, _vp, _sz, _vp]),
    "al_knn_dealt": (_int, [_vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _int, _vp, _vp, _sz, _vp]),
    "al_knn_dealt_rows": (_i64, [_i64, _i64, _i64, _i64]),
    "al_knn_density": (_int, [_vp, _i64, _i64, _i64, _i64, _int, _int, _vp, _vp, _vp, _int, _vp, _int, _vp, _vp, _sz, _vp]),
    "al_logrho_dealt": (_int, [_vp, _vp, _vp, _int, _i64, _int, _int, _i64, _i64, _i64, _i64, _vp, _vp]),
    "al_logrho_rows": (_int, [_vp, _vp, _vp, _int, _i64, _int, _int, _vp, _vp, _vp]),
    "al_enable_peer_access": (_int, [_int]),
    "al_w1_workspace_bytes": (_sz, [_i64]),
    "al_w1_prepare": (_int, [_vp, _int, _i64, _int, _int, _vp, _sz, _vp, _vp]),
    "al_w1_eval": (_int, [_vp, _sz, _i64, _c.c_double, _c.c_double, _vp, _vp]),
    "al_w1_batch_chunks": (_int, [_i64]),
    "al_w1_eval_batch": (_int, [_vp, _sz, _i64, _vp, _int, _vp, _vp]),
    "al_significance": (_int, [_vp, _int, _i64, _c.c_double, _c.c_double, _vp, _vp]),
}


class NativeError(RuntimeError):
    pass


def lib_path():
    return _LIB_PATH


def lib():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise NativeError(
                f"{_LIB_PATH} is missing: build it with `python -m astrolink_b200.build` "
                "(there is no CPU fallback)")
        L = ctypes.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def _check(rc, what):
    if rc != 0:
        msg = lib().al_last_error().decode(errors="replace")
        raise NativeError(f"{what} failed (code {rc}): {msg}")


def _dt(t):
    if t.dtype == torch.float32:
        return AL_F32
    if t.dtype == torch.float64:
        return AL_F64
    raise TypeError(f"float32/float64 expected, got {t.dtype}")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def require_cuda():
    if not torch.cuda.is_available():
        raise NativeError("astrolink_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


# --------------------------------------------------------------------------- stages
def rescale(P):
    """a1: P [n,d] cuda tensor -> (P_transform, std[d] float64)."""
    L = lib()
    n, d = P.shape
    assert P.stride(1) == 1
    out = torch.empty((n, d), dtype=P.dtype, device=P.device)
    std = torch.empty(d, dtype=torch.float64, device=P.device)
    ws = _ws(L.al_rescale_workspace_bytes(n, d), P.device)
    _check(L.al_rescale(_ptr(P), _dt(P), n, d, P.stride(0), _ptr(out), _ptr(std), _ptr(ws), ws.numel(), _stream()), "al_rescale")
    return out, std


class Index:
    """a2: spatial index over P_transform (replaces KDTree(P_transform))."""

    def __init__(self, Pt):
        L = lib()
        assert Pt.is_contiguous()
        self.n, self.d = Pt.shape
        self.dtype = Pt.dtype
        self.device = Pt.device
        nbytes = L.al_index_bytes(_dt(Pt), self.n, self.d)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=Pt.device)
        ws = _ws(L.al_index_build_workspace_bytes(_dt(Pt), self.n, self.d), Pt.device)
        _check(L.al_index_build(_ptr(Pt), _dt(Pt), self.n, self.d, _ptr(self.buf), self.buf.numel(), _ptr(ws), ws.numel(),
                                _stream()), "al_index_build")
        self.positions = int(L.al_index_positions(self.n))

    def __del__(self):
        # the library caches the index header per device pointer: drop the entry before the caching allocator can hand the
        # same address to another buffer
        try:
            if _lib is not None and getattr(self, "buf", None) is not None:
                _lib.al_index_release(_ptr(self.buf))
        except Exception:
            pass

    def perm(self):
        """index position -> original point (uint32 viewed as int32 tensor; padding = -1)."""
        return self.buf[256:256 + 4 * self.positions].view(torch.int32)

    def knn(self, k, pos_begin=0, pos_end=None, row_order=0, count_pairs=False, link_out=None, want_idx=True,
            chunk=0, stride=0, logrho_out=None, d_intrinsic=None, want_d2=True):
        """a3: exact kNN of the points at index positions [pos_begin, pos_end).  link_out: optional
        [n, L] int32 tensor (possibly peer-mapped) that receives the L nearest ids at the original row;
        with want_idx=False (link_out required) the full [rows, k] index array is not written (idx is None).
        chunk/stride (positions): search the dealt chunks [pos_begin + c*stride, +chunk) below pos_end
        (al_knn_dealt; row_order must be 1, rows in launch order).
        logrho_out (+ d_intrinsic): a4 fused into the search (al_knn_density) -- the unweighted raw log-density goes to
        logrho_out[original index] (index position for dealt launches); with want_d2=False the distance rows are not
        written at all (d2 is None)."""
        L = lib()
        pos_end = self.positions if pos_end is None else pos_end
        rows = self.n if row_order == 0 else int(L.al_knn_dealt_rows(pos_begin, pos_end, chunk, stride))
        assert want_idx or link_out is not None
        assert want_d2 or logrho_out is not None
        idx = torch.empty((rows, k), dtype=torch.int32, device=self.device) if want_idx else None
        d2 = torch.empty((rows, k), dtype=self.dtype, device=self.device) if want_d2 else None
        pairs = torch.zeros(8, dtype=torch.int64, device=self.device) if count_pairs else None
        ws = _ws(L.al_knn_workspace_bytes(AL_F32 if self.dtype == torch.float32 else AL_F64, self.n, self.d, k), self.device)
        link_cols = 0 if link_out is None else int(link_out.shape[1])
        if rows > 0:
            _check(L.al_knn_density(_ptr(self.buf), pos_begin, pos_end, chunk, stride, k, row_order, _ptr(idx), _ptr(d2), _ptr(link_out),
                                    link_cols, _ptr(logrho_out), int(d_intrinsic or 0), _ptr(pairs), _ptr(ws), ws.numel(), _stream()),
                   "al_knn_density")
        if count_pairs:
            return d2, idx, pairs
        return d2, idx


def logrho(d2, idx, k_den, d_intrinsic, weights=None, row_ids=None, out=None):
    """a4/a5: raw log-density from squared kNN distances.  With row_ids/out, row i goes to out[row_ids[i]]
    (out may be a peer-mapped tensor)."""
    L = lib()
    nq, k = d2.shape
    assert k == k_den and d2.is_contiguous()
    if out is None:
        out = torch.empty(nq, dtype=d2.dtype, device=d2.device)
    if weights is not None:
        assert weights.dtype == torch.float64 and idx is not None and idx.is_contiguous()
    _check(L.al_logrho_rows(_ptr(d2), _ptr(idx), _ptr(weights), _dt(d2), nq, k, int(d_intrinsic), _ptr(row_ids), _ptr(out), _stream()),
           "al_logrho_rows")
    return out


def logrho_dealt(d2, idx, k_den, d_intrinsic, begin, end, chunk, stride, out, weights=None):
    """a4/a5 for the rows of a dealt kNN launch, written at their index positions of `out` (possibly peer-mapped)."""
    L = lib()
    nq, k = d2.shape
    assert k == k_den and d2.is_contiguous()
    if nq == 0:
        return out
    if weights is not None:
        assert weights.dtype == torch.float64 and idx is not None and idx.is_contiguous()
    _check(L.al_logrho_dealt(_ptr(d2), _ptr(idx), _ptr(weights), _dt(d2), nq, k, int(d_intrinsic), begin, end, chunk, stride,
                             _ptr(out), _stream()), "al_logrho_dealt")
    return out


class PeerArray:
    """A device array that lives on the owner rank's GPU.  On the owner it also carries `.tensor`, a torch
    view of the same memory; on the other ranks it is only an address their kernels may store to."""

    _TYPESTR = {torch.int32: "<i4", torch.float32: "<f4", torch.float64: "<f8"}

    def __init__(self, shape, dtype, handle=None):
        L = lib()
        self.shape = tuple(int(s) for s in shape)
        self.dtype = dtype
        nbytes = int(np.prod(self.shape)) * torch.empty(0, dtype=dtype).element_size()
        p = ctypes.c_void_p(0)
        if handle is None:                       # owner: allocate and export
            buf = (ctypes.c_ubyte * 64)()
            _check(L.al_peer_alloc(nbytes, ctypes.byref(p), buf), "al_peer_alloc")
            self.handle = bytes(buf)
            self.owner = True
            self.ptr = int(p.value)
            self.__cuda_array_interface__ = {"shape": self.shape, "typestr": self._TYPESTR[dtype], "data": (self.ptr, False), "version": 2}
            self.tensor = torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device()))
        else:                                    # peer: map with the current device as the accessing device
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
            _check(L.al_peer_open(buf, ctypes.byref(p)), "al_peer_open")
            self.handle = handle
            self.owner = False
            self.ptr = int(p.value)
            self.tensor = None

    def data_ptr(self):
        return self.ptr

    def element_size(self):
        return torch.empty(0, dtype=self.dtype).element_size()

    def stride(self, i):
        st = 1
        for s in self.shape[i + 1:]:
            st *= s
        return st


def index_positions(n):
    """Number of index positions of an n-point index (n rounded up to whole leaves)."""
    return int(lib().al_index_positions(int(n)))


def peer_copy(dst, src, nbytes):
    """Asynchronous device-to-device copy on the current stream; dst / src: tensors or PeerArrays."""
    _check(lib().al_peer_copy(_ptr(dst), _ptr(src), int(nbytes), _stream()), "al_peer_copy")


def enable_peer_access(peer_device):
    _check(lib().al_enable_peer_access(int(peer_device)), "al_enable_peer_access")


def minmax(x):
    L = lib()
    mm = torch.empty(2, dtype=x.dtype, device=x.device)
    ws = _ws(L.al_normalise_workspace_bytes(x.numel()), x.device)
    _check(L.al_minmax(_ptr(x), _dt(x), x.numel(), _ptr(mm), _ptr(ws), ws.numel(), _stream()), "al_minmax")
    return mm


def normalise_(x, mm=None):
    """a7: in-place min-max normalisation."""
    L = lib()
    if mm is None:
        mm = minmax(x)
    _check(L.al_normalise(_ptr(x), _dt(x), x.numel(), _ptr(mm), _stream()), "al_normalise")
    return x


def make_edges(logRho, kNN, k_link):
    """a8: -> (weights [E], pairs [E,2] int32)."""
    L = lib()
    n = logRho.numel()
    assert kNN.shape[0] == n and kNN.stride(1) == 1
    E = n * (k_link - 1)
    w = torch.empty(E, dtype=logRho.dtype, device=logRho.device)
    pairs = torch.empty((E, 2), dtype=torch.int32, device=logRho.device)
    _check(L.al_make_edges(_ptr(logRho), _dt(logRho), _ptr(kNN), kNN.stride(0), n, int(k_link), _ptr(w), _ptr(pairs), _stream()),
           "al_make_edges")
    return w, pairs


def sort_edges(w, pairs, want_order=False):
    """a9: canonical order (weight desc, slot asc) -> sorted pairs [E,2] (and slot order)."""
    L = lib()
    E = w.numel()
    out = torch.empty_like(pairs)
    order = torch.empty(E, dtype=torch.int32, device=w.device) if want_order else None
    ws = _ws(L.al_sort_edges_workspace_bytes(_dt(w), E), w.device)
    _check(L.al_sort_edges(_ptr(w), _dt(w), _ptr(pairs), E, _ptr(out), _ptr(order), _ptr(ws), ws.numel(), _stream()), "al_sort_edges")
    return (out, order) if want_order else out


def aggregate(logRho, sorted_pairs):
    """a10/a11: -> (ordering [n] int32, groups [G,3] int32, prominences [G,2])."""
    L = lib()
    n = logRho.numel()
    E = sorted_pairs.shape[0]
    cap = n // 2 + 2
    ordering = torch.empty(n, dtype=torch.int32, device=logRho.device)
    groups = torch.empty((cap, 3), dtype=torch.int32, device=logRho.device)
    prom = torch.empty((cap, 2), dtype=logRho.dtype, device=logRho.device)
    G = ctypes.c_int64(0)
    ws = _ws(L.al_aggregate_workspace_bytes(_dt(logRho), n, E), logRho.device)
    _check(L.al_aggregate(_ptr(logRho), _dt(logRho), _ptr(sorted_pairs), n, E, _ptr(ordering), _ptr(groups), _ptr(prom),
                          ctypes.byref(G), _ptr(ws), ws.numel(), _stream()), "al_aggregate")
    g = int(G.value)
    return ordering, groups[:g], prom[:g]


def scatter_rows(src, perm, dst):
    """dst[perm[p]] = src[p] for perm[p] >= 0."""
    L = lib()
    rows, cols = src.shape
    assert src.element_size() == dst.element_size() and src.stride(1) == 1 and dst.stride(1) == 1
    _check(L.al_scatter_rows(_ptr(src), _ptr(perm), rows, cols, src.stride(0), dst.stride(0), src.element_size(), _ptr(dst), _stream()),
           "al_scatter_rows")
    return dst


def fp32_peak_tflops():
    L = lib()
    v = ctypes.c_double(0.0)
    scratch = torch.empty(64, dtype=torch.uint8, device="cuda")
    _check(L.al_fp32_peak_tflops(ctypes.byref(v), _ptr(scratch), scratch.numel(), _stream()), "al_fp32_peak_tflops")
    return float(v.value)


def significance(prom, a, b):
    """f1: norm.isf(beta.sf(prom, a, b)) element-wise -> float64 tensor of prom's shape."""
    L = lib()
    assert prom.is_contiguous()
    out = torch.empty(prom.shape, dtype=torch.float64, device=prom.device)
    _check(L.al_significance(_ptr(prom), _dt(prom), prom.numel(), float(a), float(b), _ptr(out), _stream()), "al_significance")
    return out


class W1Objective:
    """f1: GPU evaluation of the fit objective; `sorted` holds the sorted prominences (host, float64)."""

    def __init__(self, prom, col=1):
        L = lib()
        assert prom.dim() == 2 and prom.is_contiguous()
        self.G = prom.shape[0]
        self.ws = _ws(L.al_w1_workspace_bytes(self.G), prom.device)
        self.sorted = np.empty(self.G, dtype=np.float64)
        _check(L.al_w1_prepare(_ptr(prom), _dt(prom), self.G, prom.shape[1], col, _ptr(self.ws), self.ws.numel(),
                               self.sorted.ctypes.data_as(ctypes.c_void_p), _stream()), "al_w1_prepare")
        self._val = ctypes.c_double(0.0)
        self._cache = {}
        self.calls = self.hits = 0

    def __call__(self, p):
        """The objective at p.  SciPy's L-BFGS-B with jac='3-point' calls this at x and then at the finite-difference points
        around x, one by one; the points around x are predicted here (same step rule as scipy.optimize's approx_derivative:
        h = eps^(1/3) * max(1, |x|), central unless a bound is in the way) and evaluated in the same GPU call as x itself.
        A value is only ever returned for exactly the point asked for, so a wrong prediction costs a cache miss, nothing else."""
        key = (float(p[0]), float(p[1]))
        v = self._cache.get(key)
        if v is not None:
            self.hits += 1
            return v
        pts = [key] + [q for q in self._around(key) if q != key]
        ab = np.asarray(pts, dtype=np.float64)
        out = np.empty(len(pts), dtype=np.float64)
        _check(lib().al_w1_eval_batch(_ptr(self.ws), self.ws.numel(), self.G, ab.ctypes.data_as(ctypes.c_void_p), len(pts),
                                      out.ctypes.data_as(ctypes.c_void_p), _stream()), "al_w1_eval_batch")
        self._cache = {q: float(o) for q, o in zip(pts, out)}
        self.calls += 1
        return self._cache[key]

    _H = np.finfo(np.float64).eps ** (1.0 / 3.0)

    def _around(self, x, lb=1.0):
        """The points scipy's 3-point scheme evaluates around x for the bounds ((1, inf), (1, inf)) of the fit."""
        pts = []
        for i in (0, 1):
            h = self._H * max(1.0, abs(x[i]))
            if x[i] - h >= lb:                       # central
                steps = (-h, h)
            else:                                    # one-sided, away from the lower bound
                steps = (h, 2.0 * h)
            for st in steps:
                q = list(x)
                q[i] = x[i] + st
                pts.append((q[0], q[1]))
        return pts


def launch_count():
    return int(lib().al_launch_count())


def u32_numpy(t):
    """int32 cuda tensor holding uint32 bit patterns -> numpy uint32."""
    return t.detach().cpu().numpy().view(np.uint32)
