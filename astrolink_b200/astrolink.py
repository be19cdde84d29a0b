"""`AstroLink(P, ...).run()` -- the reference's public interface
(/root/reference/astrolink/astrolink.py:24-216) with its data-parallel hot path
(transform_data -> estimate_density_and_kNN -> aggregate, :218-353) executed by
hand-written sm_100a CUDA kernels behind the C ABI of include/astrolink_b200.h.

Same constructor arguments, defaults, assertion messages, stage methods, result
attributes and per-stage timers as the reference.  Differences, all additive:
  * `device` / `dist` keyword arguments (defaults preserve single-GPU behaviour);
  * `workers=-1` is accepted as documented (the reference crashes on it);
  * the unstable edge sort at :341 is replaced by a defined total order
    (weight descending, edge slot ascending) -- see DESIGN.md.
There is no CPU fallback: without a CUDA device or the built library the
constructor raises.
"""
import os
import time

import numpy as np
import torch

from . import _native as nat
from . import stats


class _Lazy:
    """Attribute that lives on the GPU and is copied to a numpy array the first
    time host code asks for it (P_transform, kNN are scratch in the reference
    and usually never read by the user; results are fetched once)."""

    def __init__(self, name, as_uint32=False):
        self.name = name
        self.dev = "_d_" + name
        self.host = "_h_" + name
        self.as_uint32 = as_uint32

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        h = obj.__dict__.get(self.host)
        if h is not None:
            return h
        d = obj.__dict__.get(self.dev)
        if d is None:
            raise AttributeError(f"'AstroLink' object has no attribute '{self.name}'")
        h = d.detach().cpu().numpy()
        if self.as_uint32:
            h = h.view(np.uint32)
        obj.__dict__[self.host] = h
        return h

    def __set__(self, obj, value):
        obj.__dict__.pop(self.dev, None)
        obj.__dict__[self.host] = value

    def __delete__(self, obj):
        found = obj.__dict__.pop(self.dev, None) is not None
        found = (obj.__dict__.pop(self.host, None) is not None) or found
        if not found:
            raise AttributeError(self.name)


class AstroLink:
    """Drop-in for the reference class (constructor: astrolink.py:132-176)."""

    P_transform = _Lazy("P_transform")
    kNN = _Lazy("kNN", as_uint32=True)
    logRho = _Lazy("logRho")
    ordering = _Lazy("ordering", as_uint32=True)
    groups = _Lazy("groups", as_uint32=True)
    prominences = _Lazy("prominences")

    def __init__(self, P, d_intrinsic=None, weights=None, k_den=20, adaptive=1, S='auto', k_link='auto', h_style=1,
                 workers=8, verbose=0, device=None, dist=None):
        # Input Data
        check_P = isinstance(P, np.ndarray) and len(P.shape) == 2
        assert check_P, "Input data 'P' needs to be a 2D numpy array!"
        self.P = P
        self.n_samples, self.d_features = self.P.shape

        check_d_intrinsic = d_intrinsic is None or (isinstance(d_intrinsic, int) and 1 <= d_intrinsic <= self.d_features)
        assert check_d_intrinsic, "Parameter 'd_intrinsic' must be None or an integer between 1 and the number of features in 'P'!"
        self.d_intrinsic = d_intrinsic if d_intrinsic is not None else self.d_features

        check_weights = weights is None or (isinstance(weights, np.ndarray) and len(weights.shape) == 1 and weights.size == self.n_samples)
        assert check_weights, "Parameter 'weights' must be None or a 1D numpy array with the same length as the number of samples in 'P'!"
        self.weights = weights

        check_k_den = int(k_den) == k_den and k_den <= self.n_samples
        assert check_k_den, "Parameter 'k_den' must be an integer less than P.shape[0]!"
        self.k_den = int(k_den)

        check_adaptive = adaptive in [0, 1]
        assert check_adaptive, "Parameter 'adaptive' must be set as either '0' or '1'!"
        self.adaptive = adaptive

        check_S = S == 'auto' or S <= 38.44939448087599
        assert check_S, "Parameter 'S' must be set to 'auto' or be less than or equal to 38.44939448087599 due to the limiting precision of the (inverse) survival functions from scipy.stats!"
        self.S = S

        check_k_link = k_link == 'auto' or 1 < k_link <= self.k_den
        assert check_k_link, "Parameter 'k_link' needs to be an integer satisfying 1 < k_link <= k_den or 'auto'!"
        if k_link == 'auto':
            # empirical fit of the reference (astrolink.py:164)
            self.k_link = max(int(np.ceil(11.97 * self.d_intrinsic ** (-2.23) - 22.97 * self.k_den ** (-0.57) + 10.03)), 7)
        else:
            self.k_link = k_link

        check_h_style = h_style in [0, 1]
        assert check_h_style, "Parameter 'h_style' must be set as either '0' or '1'!"
        self.h_style = h_style

        check_workers = 1 <= workers or workers == -1
        assert check_workers, f"Parameter 'workers' must be set as either '-1' or needs to be an integer that is >= 1 (values > N_cpu will be set to N_cpu)"
        # host threads only matter for the SciPy fit; the hot path runs on the GPU
        self.workers = workers
        self.verbose = verbose

        # ---- additions
        nat.require_cuda()
        nat.lib()
        assert self.n_samples < 2 ** 31 - 1, "astrolink_b200 supports up to 2**31 - 2 points (uint32 indices)"
        assert self.n_samples * (self.k_link - 1) < 2 ** 32 - 1, \
            "astrolink_b200 supports up to 2**32 - 2 graph edges (n_samples * (k_link - 1)): edge ranks are uint32"
        assert self.P.dtype in (np.float32, np.float64), "astrolink_b200 supports float32 / float64 input"
        assert self.d_features <= 16, "astrolink_b200 supports up to 16 features in this release"
        assert self.k_den <= 128, "astrolink_b200 supports k_den <= 128 in this release"
        self._device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._dist = dist
        self._events = {}          # stage -> (start, end) CUDA events on the launching stream
        self._knn_pairs = None     # device counter of (query, candidate) distance evaluations

    def _use_device_input(self, dP):
        """Benchmark hook: P is already resident in HBM (same values as self.P)."""
        assert dP.shape == self.P.shape and dP.is_cuda
        self._d_P_in = dP

    def _stage(self, name):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._events[name] = ev
        return ev

    def stage_ms(self):
        """Device time of each stage of the last run (CUDA events), in ms."""
        torch.cuda.synchronize(self._device)
        return {k: a.elapsed_time(b) for k, (a, b) in self._events.items()}

    # ------------------------------------------------------------------ helpers
    def _printFunction(self, message, returnLine=True, urgent=False):
        if self.verbose or urgent:
            if returnLine:
                print(f"AstroLink: {message}\r", end='')
            else:
                print(f"AstroLink: {message}")

    def _to_device(self, a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self._device, non_blocking=True)

    def _dev(self, name):
        """Device tensor behind a lazy attribute (uploading a user-injected numpy array)."""
        d = self.__dict__.get("_d_" + name)
        if d is None:
            h = self.__dict__.get("_h_" + name)
            if h is None:
                raise AttributeError(f"'{name}' has not been created yet")
            if h.dtype == np.uint32:
                h = h.view(np.int32)
            d = self._to_device(h)
            self.__dict__["_d_" + name] = d
        return d

    def _set_dev(self, name, tensor):
        self.__dict__.pop("_h_" + name, None)
        self.__dict__["_d_" + name] = tensor

    def __getstate__(self):
        # picklable like the reference object: device tensors become numpy arrays
        state = {}
        for k, v in self.__dict__.items():
            if k.startswith("_d_") or k in ("_dist", "_device", "_events", "_knn_pairs", "_pending"):
                continue
            state[k] = v
        for name in ("P_transform", "kNN", "logRho", "ordering", "groups", "prominences"):
            if ("_d_" + name) in self.__dict__ and ("_h_" + name) not in self.__dict__:
                state["_h_" + name] = getattr(self, name)
        state["_device"] = str(self._device)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._device = torch.device(state.get("_device", "cuda"))
        self._dist = None
        self._events = {}
        self._knn_pairs = None

    # ------------------------------------------------------------------ run
    def run(self):
        """transform_data, estimate_density_and_kNN, aggregate, compute_significances,
        extract_clusters (astrolink.py:183-216)."""
        self._printFunction(f"Started             | {time.strftime('%Y-%m-%d %H:%M:%S')}", returnLine=False)
        begin = time.perf_counter()
        self.transform_data()
        self.estimate_density_and_kNN()
        self.aggregate()
        if self._is_result_rank():
            self.compute_significances()
            self.extract_clusters()
        else:
            self._regrTime = self._rejTime = 0.0
        self._totalTime = time.perf_counter() - begin
        if self.verbose > 1:
            self._printFunction(f"Transformation time | {100*self._transformTime/self._totalTime:.2f}%    ", returnLine=False)
            self._printFunction(f"kNN query time      | {100*self._logRhoTime/self._totalTime:.2f}%       ", returnLine=False)
            self._printFunction(f"Aggregation time    | {100*self._aggregateTime/self._totalTime:.2f}%    ", returnLine=False)
            self._printFunction(f"Regression time     | {100*self._regrTime/self._totalTime:.2f}%         ", returnLine=False)
            self._printFunction(f"Rejection time      | {100*self._rejTime/self._totalTime:.2f}%          ", returnLine=False)
        self._printFunction(f"Completed           | {time.strftime('%Y-%m-%d %H:%M:%S')}       ", returnLine=False)

    def _world(self):
        d = self._dist
        if d is None:
            return 0, 1
        return d.get_rank(), d.get_world_size()

    def _search_width(self):
        """This rank's share of the search in dealt chunks per round (peer path; see dist.tail_aware_widths); None when the
        run is not sharded that way."""
        d = self._dist
        if d is None or d.get_world_size() == 1 or getattr(d, "peer", None) is None:
            return None
        return d.widths[d.get_rank()]

    def _is_result_rank(self):
        return self._world()[0] == 0

    # ------------------------------------------------------------------ a1
    def transform_data(self):
        """adaptive=1: every feature divided by its standard deviation; adaptive=0:
        untouched (astrolink.py:218-234).  Creates `P_transform`."""
        start = time.perf_counter()
        if self._search_width() == 0:
            # multi-GPU, dedicated tail rank: it neither searches nor builds an index, so P never goes to its GPU
            self._transformTime = time.perf_counter() - start
            return
        with torch.cuda.device(self._device):
            dP = self.__dict__.get("_d_P_in")
            if dP is None:
                dP = self._to_device(self.P)
            a, b = self._stage("rescale")
            a.record()
            if self.adaptive:
                if self.verbose > 1:
                    self._printFunction('Transforming data...      ')
                Pt, _ = nat.rescale(dP)
            else:
                Pt = dP
            b.record()
            self._set_dev("P_transform", Pt)
        self._transformTime = time.perf_counter() - start

    # ------------------------------------------------------------------ a2-a7
    def estimate_density_and_kNN(self):
        """k_den nearest neighbours of every point, Epanechnikov/balloon log-density,
        min-max normalisation; keeps the k_link nearest (astrolink.py:245-291).
        Requires `P_transform`; creates `logRho`, `kNN`; deletes `P_transform`."""
        if self.verbose > 1:
            self._printFunction('Computing densities...    ')
        start = time.perf_counter()
        with torch.cuda.device(self._device):
            n, K, L = self.n_samples, self.k_den, self.k_link
            w = None
            if self.weights is not None:
                w = self._to_device(np.asarray(self.weights, dtype=np.float64))
            rank, world = self._world()
            if self._search_width() is not None:
                raw, kNN = self._sharded_knn_peer(w, rank)
            else:
                Pt = self._dev("P_transform")
                if not Pt.is_contiguous():
                    Pt = Pt.contiguous()
                a, b = self._stage("index_build")
                a.record()
                index = nat.Index(Pt)
                b.record()
                self._index_positions = index.positions
                if world == 1:
                    a, b = self._stage("knn")
                    a.record()
                    # the k_link nearest ids go straight to their own [n, k_link] array (astrolink.py:286) from the
                    # search kernel's epilogue; the full index rows are only written when the weights need them
                    kNN = torch.empty((n, L), dtype=torch.int32, device=self._device)
                    if w is None:
                        # a4 fused into the search epilogue (al_knn_density): the [n, k_den] distance array is never written
                        raw = torch.empty((n,), dtype=Pt.dtype, device=self._device)
                        _, _, self._knn_pairs = index.knn(K, row_order=0, count_pairs=True, link_out=kNN, want_idx=False,
                                                          logrho_out=raw, d_intrinsic=self.d_intrinsic, want_d2=False)
                        b.record()
                    else:
                        d2, idx, self._knn_pairs = index.knn(K, row_order=0, count_pairs=True, link_out=kNN, want_idx=True)
                        b.record()
                        a, b = self._stage("density")
                        a.record()
                        raw = nat.logrho(d2, idx, K, self.d_intrinsic, weights=w)
                        b.record()
                        del d2, idx
                else:
                    raw, kNN = self._sharded_knn(index, w, rank, world)
            # with the peer exchange only rank 0 holds the densities (results live on rank 0)
            a, b = self._stage("normalise")
            a.record()
            lr = nat.normalise_(raw) if (world == 1 or rank == 0 or getattr(self._dist, "peer", None) is None) else raw
            b.record()
            self._set_dev("logRho", lr)
            self._set_dev("kNN", kNN)
            if "_d_P_transform" in self.__dict__ or "_h_P_transform" in self.__dict__:
                del self.P_transform
        self._logRhoTime = time.perf_counter() - start

    def _sharded_knn(self, index, w, rank, world):
        """Query-sharded kNN: every rank searches a contiguous block of index
        positions against the replicated point set, then the k_link neighbour
        lists and raw densities are all-gathered over NCCL."""
        from . import dist as adist
        n, K, L = self.n_samples, self.k_den, self.k_link
        lo, hi, per = adist.shard_positions(index.positions, rank, world, nat.LEAF)
        a, b = self._stage("knn")
        a.record()
        d2, idx, self._knn_pairs = index.knn(K, pos_begin=lo, pos_end=hi, row_order=1, count_pairs=True)
        b.record()
        raw_s = nat.logrho(d2, idx, K, self.d_intrinsic, weights=w)
        del d2
        knn_s = torch.full((per, L), -1, dtype=torch.int32, device=self._device)
        knn_s[:hi - lo] = idx[:, :L]
        raw_p = torch.full((per,), float("inf"), dtype=raw_s.dtype, device=self._device)
        raw_p[:hi - lo] = raw_s
        knn_all = torch.empty((per * world, L), dtype=torch.int32, device=self._device)
        raw_all = torch.empty((per * world,), dtype=raw_s.dtype, device=self._device)
        a, b = self._stage("all_gather")
        a.record()
        self._dist.all_gather_into_tensor(knn_all, knn_s)
        self._dist.all_gather_into_tensor(raw_all, raw_p)
        b.record()
        # back to original point order
        sel = adist.gathered_rows(index.positions, world, nat.LEAF, self._device)
        perm = index.perm()
        kNN = torch.empty((n, L), dtype=torch.int32, device=self._device)
        raw = torch.empty((n,), dtype=raw_s.dtype, device=self._device)
        nat.scatter_rows(knn_all[sel].contiguous(), perm, kNN)
        nat.scatter_rows(raw_all[sel].contiguous().view(-1, 1), perm, raw.view(-1, 1))
        return raw, kNN

    def _sharded_knn_peer(self, w, rank):
        """Exchange fused into the kernels: every searching rank's kNN epilogue stores the k_link nearest ids and the raw
        log-density of its queries directly into rank 0's arrays (peer-mapped over NVLink), in index-position order.
        The leaves are dealt to the ranks in chunks (dist.deal_chunks); rank 0, which also runs everything after the
        exchange, takes a smaller share or none (dist.tail_aware_widths) -- with none it builds no index either, and the
        first searching rank sends it the index permutation it needs to put the rows back into point order."""
        from . import dist as adist
        n, K, L = self.n_samples, self.k_den, self.k_link
        dt = torch.float32 if self.P.dtype == np.float32 else torch.float64
        peer = self._dist.peer
        world = self._world()[1]
        widths = self._dist.widths
        P_ = nat.index_positions(n)
        self._index_positions = P_
        half = peer.next_epoch()
        knn0 = peer.get(f"kNN_by_position_{half}", (P_, L), torch.int32, self._device)
        raw0 = peer.get(f"logRho_raw_by_position_{half}", (P_,), dt, self._device)
        perm0 = peer.get(f"perm_by_position_{half}", (P_,), torch.int32, self._device) if widths[0] == 0 else None
        begin, chunk, stride, rows = adist.deal_chunks(P_, rank, world, nat.LEAF, widths=widths)
        index = None
        if widths[rank] > 0:
            Pt = self._dev("P_transform")
            if not Pt.is_contiguous():
                Pt = Pt.contiguous()
            a, b = self._stage("index_build")
            a.record()
            index = nat.Index(Pt)
            b.record()
            a, b = self._stage("knn")
            a.record()
            if w is None:
                # density fused into the search epilogue: both results go straight to rank 0 (position order)
                _, _, self._knn_pairs = index.knn(K, pos_begin=begin, pos_end=P_, row_order=1, count_pairs=True, link_out=knn0,
                                                  want_idx=False, chunk=chunk, stride=stride, logrho_out=raw0,
                                                  d_intrinsic=self.d_intrinsic, want_d2=False)
                b.record()
            else:
                d2, idx, self._knn_pairs = index.knn(K, pos_begin=begin, pos_end=P_, row_order=1, count_pairs=True,
                                                     link_out=knn0, want_idx=True, chunk=chunk, stride=stride)
                b.record()
                nat.logrho_dealt(d2, idx, K, self.d_intrinsic, begin, P_, chunk, stride, raw0, weights=w)
                del d2, idx
            if perm0 is not None and rank == min(r for r in range(world) if widths[r] > 0):
                nat.peer_copy(perm0, index.perm(), 4 * P_)
        a, b = self._stage("all_gather")
        a.record()
        torch.cuda.synchronize(self._device)      # this rank's peer stores are complete ...
        self._dist.barrier()                      # ... and so are everybody else's
        b.record()
        if rank != 0:
            return torch.zeros((1,), dtype=dt, device=self._device), torch.zeros((1, L), dtype=torch.int32, device=self._device)
        # back to original point order (rank 0, local memory)
        a, b = self._stage("unpermute")
        a.record()
        perm = index.perm() if index is not None else perm0.tensor
        # rows padded to 32 bytes when they fit: every scattered row is then one full-sector store (0.5 instead of 1.2 ms at C4)
        kNN = torch.empty((n, 8), dtype=torch.int32, device=self._device)[:, :L] if L <= 8 else \
            torch.empty((n, L), dtype=torch.int32, device=self._device)
        raw = torch.empty((n,), dtype=dt, device=self._device)
        nat.scatter_rows(knn0.tensor, perm, kNN)
        nat.scatter_rows(raw0.tensor.view(-1, 1), perm, raw.view(-1, 1))
        b.record()
        return raw, kNN

    # ------------------------------------------------------------------ a8-a11
    def aggregate(self):
        """Edges between every point and its k_link neighbours weighted by the
        smaller logRho, sorted descending, aggregated into the ordered list with
        groups and prominences (astrolink.py:315-353).  Requires `logRho`, `kNN`;
        creates `ordering`, `groups`, `prominences`; deletes `kNN`.  Runs on rank 0
        when sharded."""
        if self.verbose > 1:
            self._printFunction('Aggregating points...        ')
        start = time.perf_counter()
        if not self._is_result_rank():
            del self.kNN
            self._aggregateTime = time.perf_counter() - start
            return
        with torch.cuda.device(self._device):
            lr = self._dev("logRho")
            kNN = self._dev("kNN")
            a, b = self._stage("edges")
            a.record()
            w, pairs = nat.make_edges(lr, kNN, self.k_link)
            b.record()
            del self.kNN
            a, b = self._stage("sort")
            a.record()
            sorted_pairs = nat.sort_edges(w, pairs)
            b.record()
            del w, pairs
            a, b = self._stage("aggregate")
            a.record()
            ordering, groups, prom = nat.aggregate(lr, sorted_pairs)
            b.record()
            del sorted_pairs
            self._set_dev("ordering", ordering)
            self._set_dev("groups", groups)
            self._set_dev("prominences", prom)
            # results the host needs: small ones now, the big ones while the fit runs
            self._fetch_results()
        self._aggregateTime = time.perf_counter() - start

    def _fetch_results(self):
        side = torch.cuda.Stream(device=self._device)
        side.wait_stream(torch.cuda.current_stream())
        bufs, done = {}, {}
        with torch.cuda.stream(side):
            # small results first: the fit only waits for them, the large copies run underneath it
            for name in ("groups", "prominences", "ordering", "logRho"):
                d = self.__dict__["_d_" + name]
                d.record_stream(side)          # the copy reads d on the side stream: keep its memory until then
                h = torch.empty(d.shape, dtype=d.dtype, pin_memory=True)
                h.copy_(d, non_blocking=True)
                bufs[name] = h
                done[name] = torch.cuda.Event()
                done[name].record(side)
        self._pending = (side, bufs, done)

    def _finish_fetch(self, names=None):
        """Waits for the device-to-host copies of `names` (None: all that are still pending)."""
        pend = self.__dict__.get("_pending")
        if pend is None:
            return
        side, bufs, done = pend
        for name in list(bufs):
            if names is not None and name not in names:
                continue
            done.pop(name).synchronize()
            h = bufs.pop(name).numpy()
            if name in ("groups", "ordering"):
                h = h.view(np.uint32)
            # (an attribute the user has assigned or deleted in the meantime keeps the user's value)
            if ("_d_" + name) in self.__dict__ and ("_h_" + name) not in self.__dict__:
                self.__dict__["_h_" + name] = h
        if not bufs:
            self._pending = None

    # ------------------------------------------------------------------ host statistics
    def compute_significances(self):
        """Beta model of the prominences fitted by minimising the Wasserstein-1
        distance; significances through the standard normal (astrolink.py:828-860).
        Requires `prominences`; creates `groups_sigs`, `pFit`."""
        if self.verbose > 1:
            self._printFunction('Finding significances...  ')
        start = time.perf_counter()
        self._finish_fetch(("groups", "prominences"))
        prom = self.prominences
        if prom.shape[0] >= 512:
            # large group counts: the O(G) objective is evaluated on the GPU, SciPy still drives the fit
            with torch.cuda.device(self._device):
                obj = nat.W1Objective(self._dev("prominences"))
                self.pFit, ok = stats.fit_prominence_model(None, objective=obj, p_sorted=obj.sorted)
        else:
            self.pFit, ok = stats.fit_prominence_model(prom[:, 1])
        if not ok:
            self._printFunction('[Warning] Prominence model fitting failed to converge!', returnLine=False, urgent=True)
        with torch.cuda.device(self._device):
            self.groups_sigs = nat.significance(self._dev("prominences"), self.pFit[0], self.pFit[1]).cpu().numpy()
        self._regrTime = time.perf_counter() - start

    def extract_clusters(self):
        """Groups at least S-sigma significant become clusters; hierarchy per h_style
        (astrolink.py:909-980).  Requires `groups`, `groups_sigs`; creates `clusters`,
        `ids`, `significances`.  Like the reference, S='auto' is replaced by its value."""
        if self.verbose > 1:
            self._printFunction('Finding clusters...       ')
        start = time.perf_counter()
        self._finish_fetch()
        if self.S == 'auto':
            self.S = stats.auto_threshold(self.prominences.shape[0])
        self.clusters, self.ids, self.significances = stats.extract(self.groups, self.groups_sigs, self.S, self.h_style,
                                                                    self.n_samples)
        self._rejTime = time.perf_counter() - start
