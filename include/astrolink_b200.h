/*
 * astrolink_b200.h -- C ABI of the B200-native AstroLink hot path.
 *
 * The reference (william-h-oliver/astrolink) has no FFI layer of its own: its
 * hot path is Python methods calling numba kernels and pykdtree.  Each entry
 * point below replaces one of those calls; the citation is the reference
 * file:line a maintainer would redirect (see INTEGRATION.md for the ctypes
 * binding).  All citations are into astrolink/astrolink.py.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ or torch types.
 *   - every data pointer is a DEVICE pointer owned by the caller, unless its
 *     name ends in _host.  Functions never allocate device memory: scratch is
 *     passed in as (ws, ws_bytes), sized by the matching *_workspace_bytes().
 *   - `dtype` selects the floating type of the data: AL_F32 or AL_F64 (the
 *     dtype of P in the reference; float64 input is computed in float64).
 *   - `stream` is a cudaStream_t passed as void*.
 *   - functions are asynchronous on `stream` unless stated otherwise.
 *   - return 0 on success, a negative AL_E* code otherwise; al_last_error()
 *     returns a message for the calling thread's last failure.
 *   - indices are uint32 (n < 2^32-1, as the reference's reachable dispatch,
 *     astrolink.py:268,345-350).
 */
#ifndef ASTROLINK_B200_H
#define ASTROLINK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AL_F32 0
#define AL_F64 1

#define AL_OK 0
#define AL_EINVAL (-1)    /* bad argument */
#define AL_ENOMEM (-2)    /* workspace / index buffer too small */
#define AL_ECUDA (-3)     /* CUDA runtime error, see al_last_error() */
#define AL_EINTERNAL (-4) /* internal consistency check failed */

int al_version(void);
const char* al_last_error(void);

/* ---- a1: transform_data / _rescale_njit (astrolink.py:218-243) ----------
 * Pt[i][j] = P[i][j] / std_j (population std, accumulated in float64).  A
 * zero-variance column is copied through.  P is row-major [n][d] with a row
 * stride of `row_stride` elements.  std_out (device, d doubles) may be NULL. */
size_t al_rescale_workspace_bytes(int64_t n, int d);
int al_rescale(const void* P, int dtype, int64_t n, int d, int64_t row_stride, void* Pt_out,
               double* std_out, void* ws, size_t ws_bytes, void* stream);

/* ---- a2: KDTree(P_transform) (astrolink.py:270) -------------------------
 * Builds the spatial index used by al_knn: a balanced median-split tree held
 * as a permutation of the points into 32-point leaf tiles plus bounding boxes.
 * `index` is an opaque device buffer of al_index_bytes(); Pt is row-major
 * [n][d] contiguous and is not referenced after the call returns (the index
 * holds its own tiled copy).  Synchronous (it sizes its passes on the host). */
size_t al_index_bytes(int dtype, int64_t n, int d);
size_t al_index_build_workspace_bytes(int dtype, int64_t n, int d);
int al_index_build(const void* Pt, int dtype, int64_t n, int d, void* index, size_t index_bytes,
                   void* ws, size_t ws_bytes, void* stream);
/* number of index positions (n rounded up to whole leaves) and the position ->
 * original-index permutation (device, uint32[al_index_positions], padding = 0xffffffff) */
int64_t al_index_positions(int64_t n);
const uint32_t* al_index_perm(const void* index);
/* Forget the host-side copy of an index header (the library caches it per device pointer so that al_knn need not read
 * it back from the device); call before the index buffer is freed or reused for anything but a rebuild (al_index_build
 * replaces the entry of its buffer).  Unknown pointers are ignored. */
void al_index_release(const void* index);

/* ---- a3: KDTree.query(P_t, k=k_den, sqr_dists=True) (astrolink.py:279) ---
 * Exact k nearest neighbours (self included) of the points at index positions
 * [pos_begin, pos_end) (multiples of 32, or the end) against all n points.
 * Distance: acc = fma(q_j - p_j, q_j - p_j, acc), j = 0..d-1, in `dtype`.
 * Order: (distance, original index) ascending -- ties broken by index.
 * row_order 0: row of original point i written at row i of idx_out/d2_out
 *              ([n][k] arrays);
 * row_order 1: row of index position p written at row p - pos_begin
 *              ([pos_end - pos_begin][k] arrays; padding rows are filled with
 *              0xffffffff / +inf).
 * pairs_evaluated (device uint64[8], may be NULL) accumulates kernel statistics
 * for roofline accounting: [0] (query, candidate) distance evaluations,
 * and, in diagnostic builds (-DKNN_STATS=1) only, [1] tiles fetched, [2] tiles evaluated, [3] lanes needing a
 * tile, [4] queue pushes, [5] top-k insertions, [6] merge iterations, [7] lanes needing a tile before the refresh. */
size_t al_knn_workspace_bytes(int dtype, int64_t n, int d, int k);
int al_knn(const void* index, int64_t pos_begin, int64_t pos_end, int k, int row_order,
           uint32_t* idx_out, void* d2_out, unsigned long long* pairs_evaluated,
           void* ws, size_t ws_bytes, void* stream);

/* a3 with the multi-GPU exchange fused in: additionally writes the first `link_cols` neighbour ids of every
 * query to link_out[original index][link_cols].  link_out may be a peer-mapped pointer into another GPU's
 * memory (NVLink): the rows then land directly in the consumer's kNN buffer and no all-gather is needed.
 * idx_out may be NULL when link_out is given (callers that only keep the k_link columns, astrolink.py:286). */
int al_knn_link(const void* index, int64_t pos_begin, int64_t pos_end, int k, int row_order, uint32_t* idx_out, void* d2_out,
                uint32_t* link_out, int link_cols, unsigned long long* pairs_evaluated, void* ws, size_t ws_bytes, void* stream);

/* a3 for one rank of a query-sharded run: the leaves of the spatial order are dealt to the
 * ranks in chunks so that every rank gets a sample of all regions (per-query cost varies with the local
 * structure of the cloud).  This launch searches the index positions
 *   [pos_begin + c * stride_positions, pos_begin + c * stride_positions + chunk_positions)  for c = 0, 1, ...
 * below pos_end; rank r of R passes pos_begin = r * chunk_positions, stride_positions = R * chunk_positions.
 * link_out rows are written in index-position order ([positions][link_cols], full 128-byte warp stores: the
 * consumer un-permutes them with al_scatter_rows), not by original index as in al_knn_link.
 * Rows are written in launch order (row_order must be 1): row i of the output is position
 * pos_begin + (i / chunk_positions) * stride_positions + i % chunk_positions; rows past pos_end are not written.
 * al_knn_dealt_rows gives the number of output rows.  chunk_positions == 0: identical to al_knn_link. */
int64_t al_knn_dealt_rows(int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions);
int al_knn_dealt(const void* index, int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions,
                 int k, int row_order, uint32_t* idx_out, void* d2_out, uint32_t* link_out, int link_cols,
                 unsigned long long* pairs_evaluated, void* ws, size_t ws_bytes, void* stream);

/* a3 + a4 in one launch (astrolink.py:279 and :293-297): as al_knn_dealt, and additionally the unweighted raw
 * log-density of every query -- the value al_logrho computes from the same distance row, bit for bit -- is written
 * from the search epilogue while the sorted distances are still in shared memory: logrho_out[original index] (or
 * [index position] for dealt launches, like link_out; may be peer-mapped).  With logrho_out given, d2_out may be
 * NULL: the [rows][k] distance array (the largest output of the search) is then never written.  logrho_out == NULL:
 * identical to al_knn_dealt.  The weighted density (a5) needs the neighbour ids and stays in al_logrho. */
int al_knn_density(const void* index, int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions,
                   int k, int row_order, uint32_t* idx_out, void* d2_out, uint32_t* link_out, int link_cols,
                   void* logrho_out, int d_intrinsic, unsigned long long* pairs_evaluated, void* ws, size_t ws_bytes, void* stream);

/* ---- a4/a5: _compute_logRho_njit / _compute_weighted_logRho_njit
 *      (astrolink.py:293-303, gather of weights[indices] at :283) ----------
 * out[i] = log((sum_j w_j - sum_j w_j d2[i][j] / d2[i][k-1]) / d2[i][k-1]^(d_intrinsic/2)),
 * w_j = weights[idx[i][j]] (double) or 1 when weights == NULL.  Evaluated in
 * float64, stored in `dtype` (as the reference's store at :282). */
int al_logrho(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq,
              int k, int d_intrinsic, void* logrho_out, void* stream);
/* a4/a5 for the rows of an al_knn_dealt launch: row i is written at its index position
 * pos_begin + (i / chunk_positions) * stride_positions + i % chunk_positions of logrho_by_position_out (rows past
 * pos_end are skipped).  Consecutive threads store consecutive elements, so the output may be a peer-mapped buffer. */
int al_logrho_dealt(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k, int d_intrinsic,
                    int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions,
                    void* logrho_by_position_out, void* stream);

/* as al_logrho, but row i is written to logrho_out[row_ids[i]] (row_ids == 0xffffffff: skipped); logrho_out may be
 * a peer-mapped pointer (multi-GPU: every rank scatters its densities into the consumer's array). */
int al_logrho_rows(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k, int d_intrinsic,
                   const uint32_t* row_ids, void* logrho_out, void* stream);

/* ---- a7: _normalise_njit (astrolink.py:305-313) -------------------------
 * x <- (x - min) / (max - min) over all n.  minmax_out (device, 2 values of
 * `dtype`, may be NULL) receives {min, max}. */
size_t al_normalise_workspace_bytes(int64_t n);
int al_minmax(const void* x, int dtype, int64_t n, void* minmax_out, void* ws, size_t ws_bytes, void* stream);
int al_normalise(void* x, int dtype, int64_t n, const void* minmax, void* stream);

/* ---- a8: _make_graph_njit (astrolink.py:355-382) ------------------------
 * knn is [n][knn_stride] uint32, the first L columns are used.  Row i owns
 * edge slots [i(L-1), (i+1)(L-1)): the entry equal to i is skipped (if absent
 * the last entry is dropped); pair = (denser, sparser), ties -> (i, j);
 * weight = logrho of the sparser end.  weights_out: `dtype`[E];
 * pairs_out: uint32[E][2]. */
int al_make_edges(const void* logrho, int dtype, const uint32_t* knn, int64_t knn_stride, int64_t n,
                  int L, void* weights_out, uint32_t* pairs_out, void* stream);

/* ---- a9: pairs[edges.argsort()[::-1]] (astrolink.py:341) ----------------
 * Canonical total order: weight descending, slot ascending (the reference's
 * argsort is unstable).  Stable descending radix sort (CUB).
 * order_out (uint32[E], may be NULL) receives the sorted slot indices; without
 * it the pairs travel through the sort as 64-bit values (no slot numbers, no
 * gather).  pairs and sorted_pairs_out must be 8-byte aligned. */
size_t al_sort_edges_workspace_bytes(int dtype, int64_t E);
int al_sort_edges(const void* weights, int dtype, const uint32_t* pairs, int64_t E,
                  uint32_t* sorted_pairs_out, uint32_t* order_out, void* ws, size_t ws_bytes, void* stream);

/* ---- a10/a11: _aggregate_njit_float{32,64}_uint32 (astrolink.py:384-604)
 * sorted_pairs: uint32[E][2] in the order above.  The float32 variant applies
 * the final noise correction to prominence column 0 (:484), the float64
 * variant to column 1 (:595); the variant follows `dtype` as the reference's
 * dispatch does (:345-350).
 * Outputs: ordering uint32[n]; groups uint32[G][3] = [geq_start, leq_start,
 * leq_end]; prominences `dtype`[G][2]; capacity n/2 rows each.
 * *n_groups_host (HOST pointer) receives G.  Synchronous. */
size_t al_aggregate_workspace_bytes(int dtype, int64_t n, int64_t E);
int al_aggregate(const void* logrho, int dtype, const uint32_t* sorted_pairs, int64_t n, int64_t E,
                 uint32_t* ordering_out, uint32_t* groups_out, void* prominences_out,
                 int64_t* n_groups_host, void* ws, size_t ws_bytes, void* stream);

/* ---- f1: groups_sigs = norm.isf(beta.sf(prominences, a, b)) (astrolink.py:858) ----
 * Element-wise over `count` prominence values (`dtype`), result float64.  (a, b)
 * are the Beta parameters fitted on the host; the fit itself stays in Python. */
int al_significance(const void* prominences, int dtype, int64_t count, double a, double b, double* sig_out, void* stream);

/* ---- f1: objective of the Beta prominence-model fit (astrolink.py:880-885, 896-907) ----
 * The optimiser stays on the host (SciPy); these evaluate its O(G) objective.
 * al_w1_prepare: sorts column `col` of the [G][stride] prominence array (`dtype`), builds
 * the midpoint tables in `ws`, copies the sorted values (float64) to sorted_host (HOST, may
 * be NULL).  al_w1_eval: the Wasserstein-1 distance between the midpoint-rule cdf of
 * Beta(a, b) and the empirical cdf, written to *value_host (HOST).  Both synchronous. */
size_t al_w1_workspace_bytes(int64_t G);
int al_w1_prepare(const void* prominences, int dtype, int64_t G, int stride, int col, void* ws, size_t ws_bytes,
                  double* sorted_host, void* stream);
int al_w1_eval(void* ws, size_t ws_bytes, int64_t G, double a, double b, double* value_host, void* stream);
/* The same objective at B <= 16 parameter pairs (ab_host: B x {a, b}) in one call: two launches and one host read for
 * all of them -- the optimiser's finite-difference points are evaluated together with the point itself.  Needs
 * G >= 16 * al_w1_batch_chunks(G).  Sums run in a fixed order: a given (a, b) always yields the same value, whichever
 * batch it is part of. */
int al_w1_batch_chunks(int64_t G);
int al_w1_eval_batch(void* ws, size_t ws_bytes, int64_t G, const double* ab_host, int B, double* values_host, void* stream);

/* ---- helpers ------------------------------------------------------------ */
/* scatter rows: dst[perm[p]][0..cols) = src[p][0..cols) for p in [0, rows), skipping
 * perm == 0xffffffff; elem_bytes is 4 or 8. */
int al_scatter_rows(const void* src, const uint32_t* perm, int64_t rows, int cols, int src_stride,
                    int dst_stride, int elem_bytes, void* dst, void* stream);
/* Peer buffers of the multi-GPU path.  al_peer_alloc: device memory on the current device plus its 64-byte CUDA
 * IPC handle (both returned through HOST pointers); al_peer_open: maps such a buffer into this process with the
 * CURRENT device as the accessing device (its kernels may then store into it over NVLink); al_peer_release frees
 * (owner != 0) or unmaps (owner == 0).  Set-up calls, not on the hot path. */
int al_peer_alloc(size_t bytes, void** ptr_host, unsigned char* handle64_host);
int al_peer_open(const unsigned char* handle64_host, void** ptr_host);
int al_peer_release(void* ptr, int owner);
/* Device-to-device copy on `stream`; dst or src may be a buffer mapped with al_peer_open (multi-GPU path: the index
 * permutation travels to the rank that un-permutes the neighbour rows). */
int al_peer_copy(void* dst, const void* src, size_t bytes, void* stream);
/* enables peer access from the current device to `peer_device` (idempotent). */
int al_enable_peer_access(int peer_device);
/* kernels launched by this library so far in this process (its own kernels; CUB's
 * internal launches are not counted). */
int64_t al_launch_count(void);
/* measured FP32 FMA issue rate of this GPU in TFLOP/s (2 flop per FMA); used as
 * the roofline denominator of the kNN kernel.  Synchronous.  `scratch`: >= 4 bytes of
 * caller-owned device memory (like every entry point, this one allocates nothing). */
int al_fp32_peak_tflops(double* tflops_host, void* scratch, size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
