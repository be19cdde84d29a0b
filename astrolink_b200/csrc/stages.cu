// The streaming stages of the hot path around the kNN search:
//   al_rescale      <- transform_data / _rescale_njit      (astrolink.py:218-243)
//   al_logrho       <- _compute_logRho_njit / weighted      (astrolink.py:293-303, gather :283)
//   al_minmax / al_normalise <- _normalise_njit             (astrolink.py:305-313)
//   al_make_edges   <- _make_graph_njit                     (astrolink.py:355-382)
//   al_sort_edges   <- pairs[edges.argsort()[::-1]]         (astrolink.py:341), canonical stable order
// All are HBM-bound passes; see DESIGN.md for the algorithmic bytes of each.
#include <cub/cub.cuh>
#include <stdarg.h>

#include "common.cuh"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void al_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* al_last_error(void) { return g_err; }
extern "C" int al_version(void) { return 1; }

#include <atomic>
static std::atomic<long> g_launches{0};
void al_count_launches(long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t al_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

constexpr int RS_MAXD = 16;
constexpr int RS_BLOCKS = 592;   // 4 x 148 SMs
constexpr int RS_THREADS = 256;

// ---------------------------------------------------------------------------
// a1 rescale: deterministic two-level column reductions in float64
// ---------------------------------------------------------------------------
template <typename T, bool CENTRED>
__global__ void __launch_bounds__(RS_THREADS) column_sums_kernel(const T* __restrict__ P, i64 n, int d, i64 row_stride,
                                                                 const double* __restrict__ mean, double* __restrict__ partial) {
    double acc[RS_MAXD];
#pragma unroll
    for (int j = 0; j < RS_MAXD; j++) acc[j] = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const T* row = P + i * row_stride;
#pragma unroll
        for (int j = 0; j < RS_MAXD; j++)
            if (j < d) {
                double v = (double)row[j];
                if (CENTRED) { v -= mean[j]; v *= v; }
                acc[j] += v;
            }
    }
    __shared__ double sm[RS_THREADS / 32][RS_MAXD];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < RS_MAXD; j++)
        if (j < d) {
            double v = acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            if (lane == 0) sm[w][j] = v;
        }
    __syncthreads();
    if (threadIdx.x < d) {
        double v = 0.0;
        for (int k = 0; k < RS_THREADS / 32; k++) v += sm[k][threadIdx.x];
        partial[(i64)blockIdx.x * d + threadIdx.x] = v;
    }
}

// one thread per column: fixed-order sum of the block partials
template <bool FINAL_STD>
__global__ void column_finish_kernel(const double* __restrict__ partial, int blocks, int d, i64 n, double* out) {
    int j = threadIdx.x;
    if (j >= d) return;
    double v = 0.0;
    for (int b = 0; b < blocks; b++) v += partial[(i64)b * d + j];
    out[j] = FINAL_STD ? sqrt(v / (double)n) : v / (double)n;
}

template <typename T>
__global__ void __launch_bounds__(256) rescale_apply_kernel(const T* __restrict__ P, i64 n, int d, i64 row_stride,
                                                            const double* __restrict__ sd, T* __restrict__ out) {
    i64 total = n * d;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64)gridDim.x * blockDim.x) {
        i64 i = t / d;
        int j = (int)(t - i * d);
        T v = P[i * row_stride + j];
        double s = sd[j];
        out[t] = s > 0.0 ? (T)(v / (T)s) : v;
    }
}

extern "C" size_t al_rescale_workspace_bytes(int64_t n, int d) {
    (void)n;
    ArenaSizer s;
    s.add<double>((size_t)RS_BLOCKS * d);
    s.add<double>(d);
    s.add<double>(d);
    return s.off + 256;
}

template <typename T>
static int rescale_impl(const T* P, i64 n, int d, i64 row_stride, T* out, double* std_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    Arena ar(ws, ws_bytes);
    double* partial = ar.get<double>((size_t)RS_BLOCKS * d);
    double* mean = ar.get<double>(d);
    double* sd = ar.get<double>(d);
    AL_REQUIRE(ar.ok, AL_ENOMEM, "al_rescale: workspace too small");
    int blocks = al_grid(n, RS_THREADS, RS_BLOCKS);
    column_sums_kernel<T, false><<<blocks, RS_THREADS, 0, st>>>(P, n, d, row_stride, nullptr, partial);
    AL_CHECK_LAUNCH();
    column_finish_kernel<false><<<1, 32, 0, st>>>(partial, blocks, d, n, mean);
    AL_CHECK_LAUNCH();
    column_sums_kernel<T, true><<<blocks, RS_THREADS, 0, st>>>(P, n, d, row_stride, mean, partial);
    AL_CHECK_LAUNCH();
    column_finish_kernel<true><<<1, 32, 0, st>>>(partial, blocks, d, n, sd);
    AL_CHECK_LAUNCH();
    rescale_apply_kernel<T><<<al_grid(n * d, 256), 256, 0, st>>>(P, n, d, row_stride, sd, out);
    AL_CHECK_LAUNCH();
    if (std_out) AL_CUDA(cudaMemcpyAsync(std_out, sd, sizeof(double) * d, cudaMemcpyDeviceToDevice, st));
    return AL_OK;
}

extern "C" int al_rescale(const void* P, int dtype, int64_t n, int d, int64_t row_stride, void* Pt_out, double* std_out,
                          void* ws, size_t ws_bytes, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_rescale: bad dtype");
    AL_REQUIRE(n >= 1 && d >= 1 && d <= RS_MAXD && row_stride >= d, AL_EINVAL, "al_rescale: bad shape (d must be <= %d)", RS_MAXD);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AL_F32) return rescale_impl<float>((const float*)P, n, d, row_stride, (float*)Pt_out, std_out, ws, ws_bytes, st);
    return rescale_impl<double>((const double*)P, n, d, row_stride, (double*)Pt_out, std_out, ws, ws_bytes, st);
}

// ---------------------------------------------------------------------------
// a4/a5 density: one thread per query row
// ---------------------------------------------------------------------------
// One thread per query row.  A warp's 32 rows are contiguous (32*k elements), so DRAM traffic is ideal;
// rows are read with 128-bit loads when their pitch allows it (float: k % 4 == 0), which quarters the
// number of L1 requests.  The per-row arithmetic is the reference's: T-precision sequential sum, ratio
// in T, the rest in float64.
struct DealMap { i64 begin, chunk, stride, end; };      // chunk == 0: rows are written in place

template <typename T, bool VEC4>
__global__ void __launch_bounds__(256) logrho_kernel(const T* __restrict__ d2, const u32* __restrict__ idx,
                                                     const double* __restrict__ weights, i64 nq, int k, double half_d,
                                                     const u32* __restrict__ row_ids, DealMap dm, T* __restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    i64 orow = i;
    if (row_ids != nullptr) {
        const u32 rid = row_ids[i];
        if (rid == 0xffffffffu) return;      // padding row of the last leaf
        orow = rid;
    } else if (dm.chunk > 0) {
        orow = dm.begin + (i / dm.chunk) * dm.stride + i % dm.chunk;      // index position of row i of a dealt launch
        if (orow >= dm.end) return;
    }
    const T* r = d2 + i * k;
    const T coreT = r[k - 1];
    const double core = (double)coreT;
    double v;
    if (weights == nullptr) {
        T s = (T)0;
        if (VEC4) {
            const float4* r4 = reinterpret_cast<const float4*>(r);
            for (int j = 0; j < k / 4; j++) {
                const float4 f = r4[j];
                s += f.x; s += f.y; s += f.z; s += f.w;      // same left-to-right order as the scalar loop
            }
        } else {
            for (int j = 0; j < k; j++) s += r[j];
        }
        out[orow] = al_raw_logrho<T>(s, coreT, k, half_d);
        return;
    } else {
        const u32* ir = idx + i * k;
        if (ir[0] == 0xffffffffu) return;      // padding row of the last leaf (position-ordered launches): no neighbours
        double sw = 0.0, swd = 0.0;
        for (int j = 0; j < k; j++) {
            const double w = weights[ir[j]];
            sw += w;
            swd += w * (double)r[j];
        }
        v = (sw - swd / core) / pow(core, half_d);
    }
    out[orow] = (T)log(v);
}

extern "C" int al_logrho(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k,
                         int d_intrinsic, void* logrho_out, void* stream) {
    return al_logrho_rows(d2, idx, weights, dtype, nq, k, d_intrinsic, nullptr, logrho_out, stream);
}

static int logrho_launch(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k, int d_intrinsic,
                         const uint32_t* row_ids, DealMap dm, void* logrho_out, void* stream);

extern "C" int al_logrho_rows(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k,
                              int d_intrinsic, const uint32_t* row_ids, void* logrho_out, void* stream) {
    return logrho_launch(d2, idx, weights, dtype, nq, k, d_intrinsic, row_ids, DealMap{0, 0, 0, 0}, logrho_out, stream);
}

extern "C" int al_logrho_dealt(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k,
                               int d_intrinsic, int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions,
                               void* logrho_by_position_out, void* stream) {
    AL_REQUIRE(chunk_positions > 0 && stride_positions >= chunk_positions, AL_EINVAL, "al_logrho_dealt: bad chunk/stride");
    return logrho_launch(d2, idx, weights, dtype, nq, k, d_intrinsic, nullptr, DealMap{pos_begin, chunk_positions, stride_positions, pos_end},
                         logrho_by_position_out, stream);
}

static int logrho_launch(const void* d2, const uint32_t* idx, const double* weights, int dtype, int64_t nq, int k, int d_intrinsic,
                         const uint32_t* row_ids, DealMap dm, void* logrho_out, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_logrho: bad dtype");
    AL_REQUIRE(nq >= 0 && k >= 1 && d_intrinsic >= 1, AL_EINVAL, "al_logrho: bad shape");
    AL_REQUIRE(weights == nullptr || idx != nullptr, AL_EINVAL, "al_logrho: weights need idx");
    if (nq == 0) return AL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)((nq + 255) / 256);
    if (dtype == AL_F32) {
        if (k % 4 == 0 && ((uintptr_t)d2 & 15) == 0)
            logrho_kernel<float, true><<<grid, 256, 0, st>>>((const float*)d2, idx, weights, nq, k, d_intrinsic / 2.0, row_ids, dm, (float*)logrho_out);
        else
            logrho_kernel<float, false><<<grid, 256, 0, st>>>((const float*)d2, idx, weights, nq, k, d_intrinsic / 2.0, row_ids, dm, (float*)logrho_out);
    } else {
        logrho_kernel<double, false><<<grid, 256, 0, st>>>((const double*)d2, idx, weights, nq, k, d_intrinsic / 2.0, row_ids, dm, (double*)logrho_out);
    }
    AL_CHECK_LAUNCH();
    return AL_OK;
}

// ---------------------------------------------------------------------------
// a7 min/max + normalise
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) minmax_partial_kernel(const T* __restrict__ x, i64 n, T* __restrict__ partial) {
    T mn = (T)INFINITY, mx = -(T)INFINITY;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        T v = x[i];
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    __shared__ T smn[8], smx[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { mn = smn[w] < mn ? smn[w] : mn; mx = smx[w] > mx ? smx[w] : mx; }
        partial[2 * blockIdx.x] = mn;
        partial[2 * blockIdx.x + 1] = mx;
    }
}
template <typename T>
__global__ void minmax_final_kernel(const T* __restrict__ partial, int blocks, T* out) {
    T mn = (T)INFINITY, mx = -(T)INFINITY;
    for (int b = threadIdx.x; b < blocks; b += 32) {
        T a = partial[2 * b], c = partial[2 * b + 1];
        mn = a < mn ? a : mn;
        mx = c > mx ? c : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; }
}
template <typename T>
__global__ void __launch_bounds__(256) normalise_kernel(T* x, i64 n, const T* __restrict__ mm) {
    const T mn = mm[0], range = mm[1] - mm[0];
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] = (x[i] - mn) / range;
}

extern "C" size_t al_normalise_workspace_bytes(int64_t n) {
    (void)n;
    return al_align((size_t)RS_BLOCKS * 2 * 8) + al_align(16) + 256;
}

extern "C" int al_minmax(const void* x, int dtype, int64_t n, void* minmax_out, void* ws, size_t ws_bytes, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_minmax: bad dtype");
    AL_REQUIRE(n >= 1 && minmax_out != nullptr, AL_EINVAL, "al_minmax: bad arguments");
    Arena ar(ws, ws_bytes);
    double* partial = ar.get<double>((size_t)RS_BLOCKS * 2);
    AL_REQUIRE(ar.ok, AL_ENOMEM, "al_minmax: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    int blocks = al_grid(n, 256, RS_BLOCKS);
    if (dtype == AL_F32) {
        minmax_partial_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, n, (float*)partial);
        minmax_final_kernel<float><<<1, 32, 0, st>>>((const float*)partial, blocks, (float*)minmax_out);
        al_count_launches(1);
    } else {
        minmax_partial_kernel<double><<<blocks, 256, 0, st>>>((const double*)x, n, partial);
        minmax_final_kernel<double><<<1, 32, 0, st>>>(partial, blocks, (double*)minmax_out);
        al_count_launches(1);
    }
    AL_CHECK_LAUNCH();
    return AL_OK;
}

extern "C" int al_normalise(void* x, int dtype, int64_t n, const void* minmax, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_normalise: bad dtype");
    AL_REQUIRE(n >= 1 && minmax != nullptr, AL_EINVAL, "al_normalise: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AL_F32) normalise_kernel<float><<<al_grid(n, 256), 256, 0, st>>>((float*)x, n, (const float*)minmax);
    else normalise_kernel<double><<<al_grid(n, 256), 256, 0, st>>>((double*)x, n, (const double*)minmax);
    AL_CHECK_LAUNCH();
    return AL_OK;
}

// ---------------------------------------------------------------------------
// a8 edges: one thread per point, L-1 edge slots each
// ---------------------------------------------------------------------------
constexpr int ME_CHUNK = 8;
template <typename T>
__global__ void __launch_bounds__(256) make_edges_kernel(const T* __restrict__ logrho, const u32* __restrict__ knn, i64 knn_stride,
                                                         i64 n, int L, T* __restrict__ w_out, uint2* __restrict__ pairs_out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T li = logrho[i];
    i64 pos = i * (L - 1);
    const i64 end = pos + (L - 1);
    const u32* row = knn + i * knn_stride;
    for (int r0 = 0; r0 < L; r0 += ME_CHUNK) {
        u32 j[ME_CHUNK];
        T lj[ME_CHUNK];
#pragma unroll
        for (int e = 0; e < ME_CHUNK; e++) j[e] = r0 + e < L ? row[r0 + e] : (u32)i;          // neighbour ids first ...
#pragma unroll
        for (int e = 0; e < ME_CHUNK; e++) lj[e] = logrho[j[e]];                                // ... then all the gathers in flight
#pragma unroll
        for (int e = 0; e < ME_CHUNK; e++) {
            if ((i64)j[e] == i || pos >= end) continue;      // the point itself (and chunk padding) is skipped
            if (li >= lj[e]) { w_out[pos] = lj[e]; pairs_out[pos] = make_uint2((u32)i, j[e]); }
            else { w_out[pos] = li; pairs_out[pos] = make_uint2(j[e], (u32)i); }
            pos++;
        }
    }
}

// Same edges, memory traffic reshaped for HBM: a block of 256 points reads its 256 x L neighbour ids as one
// contiguous run (coalesced), every thread gathers the L densities of its row (the 4 n bytes of logRho stay in
// L2), builds its L-1 edges in shared memory, and the block's 256 (L-1) edges -- contiguous in the output --
// leave with fully coalesced stores.  Every point emits exactly L-1 edges (its row holds itself once, or --
// rows of coincident duplicates, App. B Q7 -- the last entry is dropped).
constexpr int ME_BLOCK = 256;
template <typename T>
static size_t make_edges_smem_bytes(int L) { return (size_t)ME_BLOCK * ((size_t)L * 4 + (size_t)(L - 1) * (sizeof(T) + 8)); }

template <typename T>
__global__ void __launch_bounds__(ME_BLOCK) make_edges_staged_kernel(const T* __restrict__ logrho, const u32* __restrict__ knn,
                                                                     i64 knn_stride, i64 n, int L, T* __restrict__ w_out,
                                                                     uint2* __restrict__ pairs_out) {
    extern __shared__ __align__(16) unsigned char me_smem[];
    const int E1 = L - 1;
    uint2* s_pairs = (uint2*)me_smem;                                  // [ME_BLOCK * (L-1)]
    T* s_w = (T*)(s_pairs + (size_t)ME_BLOCK * E1);                    // [ME_BLOCK * (L-1)]
    u32* s_knn = (u32*)(s_w + (size_t)ME_BLOCK * E1);                  // [ME_BLOCK * L]
    const i64 b0 = (i64)blockIdx.x * ME_BLOCK;
    const int rows = (int)((n - b0) < ME_BLOCK ? (n - b0) : ME_BLOCK);
    const int t = threadIdx.x;
    if (knn_stride == L) {
        const u32* src = knn + b0 * L;
        for (int e = t; e < rows * L; e += ME_BLOCK) s_knn[e] = src[e];
    } else if (knn_stride == 8 && L <= 8) {
        // rows padded to 32 bytes (the un-permute pass of the sharded path writes them so): one thread per row, two 128-bit loads
        if (t < rows) {
            const uint4* src = reinterpret_cast<const uint4*>(knn + (b0 + t) * 8);
            const uint4 a = src[0], b = src[1];
            const u32 v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int c = 0; c < 8; c++) if (c < L) s_knn[t * L + c] = v[c];
        }
    } else {
        for (int e = t; e < rows * L; e += ME_BLOCK) s_knn[e] = knn[(b0 + e / L) * knn_stride + e % L];
    }
    __syncthreads();
    if (t < rows) {
        const i64 i = b0 + t;
        const T li = logrho[i];
        const u32* row = s_knn + t * L;
        int o = t * E1;
        const int oend = o + E1;
        for (int r0 = 0; r0 < L; r0 += ME_CHUNK) {
            u32 j[ME_CHUNK];
            T lj[ME_CHUNK];
#pragma unroll
            for (int e = 0; e < ME_CHUNK; e++) j[e] = r0 + e < L ? row[r0 + e] : (u32)i;
#pragma unroll
            for (int e = 0; e < ME_CHUNK; e++) lj[e] = logrho[j[e]];                            // all the gathers in flight
#pragma unroll
            for (int e = 0; e < ME_CHUNK; e++) {
                if ((i64)j[e] == i || o >= oend) continue;
                if (li >= lj[e]) { s_w[o] = lj[e]; s_pairs[o] = make_uint2((u32)i, j[e]); }
                else { s_w[o] = li; s_pairs[o] = make_uint2(j[e], (u32)i); }
                o++;
            }
        }
    }
    __syncthreads();
    const i64 pos0 = b0 * E1;
    const int cnt = rows * E1;
    for (int e = t; e < cnt; e += ME_BLOCK) w_out[pos0 + e] = s_w[e];
    for (int e = t; e < cnt; e += ME_BLOCK) pairs_out[pos0 + e] = s_pairs[e];
}

template <typename T>
static int make_edges_launch(const T* logrho, const u32* knn, i64 knn_stride, i64 n, int L, T* w_out, uint2* pairs_out, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + ME_BLOCK - 1) / ME_BLOCK);
    const size_t smem = make_edges_smem_bytes<T>(L);
    if (smem <= 160 * 1024) {
        if (smem > 48 * 1024)
            AL_CUDA(cudaFuncSetAttribute(make_edges_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        make_edges_staged_kernel<T><<<grid, ME_BLOCK, smem, st>>>(logrho, knn, knn_stride, n, L, w_out, pairs_out);
    } else {
        make_edges_kernel<T><<<grid, 256, 0, st>>>(logrho, knn, knn_stride, n, L, w_out, pairs_out);      // very long rows
    }
    AL_CHECK_LAUNCH();
    return AL_OK;
}

extern "C" int al_make_edges(const void* logrho, int dtype, const uint32_t* knn, int64_t knn_stride, int64_t n, int L,
                             void* weights_out, uint32_t* pairs_out, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_make_edges: bad dtype");
    AL_REQUIRE(n >= 1 && L >= 2 && knn_stride >= L, AL_EINVAL, "al_make_edges: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AL_F32)
        return make_edges_launch<float>((const float*)logrho, knn, knn_stride, n, L, (float*)weights_out, (uint2*)pairs_out, st);
    return make_edges_launch<double>((const double*)logrho, knn, knn_stride, n, L, (double*)weights_out, (uint2*)pairs_out, st);
}

// ---------------------------------------------------------------------------
// a9 canonical edge order: stable descending radix sort of (weight, slot)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) iota_kernel(u32* v, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) v[i] = (u32)i;
}
__global__ void __launch_bounds__(256) gather_pairs_kernel(const uint2* __restrict__ pairs, const u32* __restrict__ order, i64 E,
                                                           uint2* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (i64)gridDim.x * blockDim.x) out[i] = pairs[order[i]];
}

template <typename T>
static size_t sort_tmp_bytes(i64 E) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, b, (const T*)nullptr, (T*)nullptr, (const u32*)nullptr, (u32*)nullptr, E, 0,
                                              (int)sizeof(T) * 8, (cudaStream_t)0);
    return b;
}

template <typename T>
static size_t sort_tmp_bytes_direct(i64 E) {
    size_t b = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, b, (const T*)nullptr, (T*)nullptr, (const unsigned long long*)nullptr,
                                              (unsigned long long*)nullptr, E, 0, (int)sizeof(T) * 8, (cudaStream_t)0);
    return b;
}
extern "C" size_t al_sort_edges_workspace_bytes(int dtype, int64_t E) {
    // the larger of the two layouts of sort_edges_impl (with and without the slot order)
    const size_t kb = dtype == AL_F64 ? 8 : 4;
    ArenaSizer d;
    d.add<char>((size_t)E * kb);
    d.add<char>(dtype == AL_F64 ? sort_tmp_bytes_direct<double>(E) : sort_tmp_bytes_direct<float>(E));
    ArenaSizer s;
    s.add<char>((size_t)E * kb);
    s.add<u32>(E);
    s.add<u32>(E);
    s.add<char>(dtype == AL_F64 ? sort_tmp_bytes<double>(E) : sort_tmp_bytes<float>(E));
    return (d.off > s.off ? d.off : s.off) + 256;
}

template <typename T>
static int sort_edges_impl(const T* w, const uint2* pairs, i64 E, uint2* out, u32* order_out, void* ws, size_t ws_bytes, cudaStream_t st) {
    Arena ar(ws, ws_bytes);
    T* keys_out = (T*)ar.get<char>((size_t)E * sizeof(T));
    if (order_out == nullptr) {
        // the 64-bit pairs ride through the sort as values: no slot numbers, no gather (2.58 -> 2.08 ms at C4; the same stable
        // order, so the same bytes)
        size_t tbd = sort_tmp_bytes_direct<T>(E);
        void* tmpd = ar.get<char>(tbd);
        AL_REQUIRE(ar.ok, AL_ENOMEM, "al_sort_edges: workspace too small");
        AL_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmpd, tbd, w, keys_out, (const unsigned long long*)pairs, (unsigned long long*)out, E, 0,
                                                          (int)sizeof(T) * 8, st));
        return AL_OK;
    }
    u32* vals_in = ar.get<u32>(E);
    u32* vals_out = ar.get<u32>(E);
    size_t tb = sort_tmp_bytes<T>(E);
    void* tmp = ar.get<char>(tb);
    AL_REQUIRE(ar.ok, AL_ENOMEM, "al_sort_edges: workspace too small");
    iota_kernel<<<al_grid(E, 256), 256, 0, st>>>(vals_in, E);
    AL_CHECK_LAUNCH();
    u32* dst = order_out ? order_out : vals_out;
    AL_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, tb, w, keys_out, vals_in, dst, E, 0, (int)sizeof(T) * 8, st));
    gather_pairs_kernel<<<al_grid(E, 256), 256, 0, st>>>(pairs, dst, E, out);
    AL_CHECK_LAUNCH();
    return AL_OK;
}

extern "C" int al_sort_edges(const void* weights, int dtype, const uint32_t* pairs, int64_t E, uint32_t* sorted_pairs_out,
                             uint32_t* order_out, void* ws, size_t ws_bytes, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_sort_edges: bad dtype");
    AL_REQUIRE(E >= 1 && E < 0xffffffffll, AL_EINVAL, "al_sort_edges: E out of range");
    AL_REQUIRE((((uintptr_t)pairs | (uintptr_t)sorted_pairs_out) & 7u) == 0, AL_EINVAL, "al_sort_edges: pairs must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AL_F32)
        return sort_edges_impl<float>((const float*)weights, (const uint2*)pairs, E, (uint2*)sorted_pairs_out, order_out, ws, ws_bytes, st);
    return sort_edges_impl<double>((const double*)weights, (const uint2*)pairs, E, (uint2*)sorted_pairs_out, order_out, ws, ws_bytes, st);
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
template <typename V>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const V* __restrict__ src, const u32* __restrict__ perm, i64 rows, int cols,
                                                           int src_stride, int dst_stride, V* __restrict__ dst) {
    // one thread per row: the permutation entry is read once and the (short) row is copied with a column loop
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (i64)gridDim.x * blockDim.x) {
        const u32 to = perm[r];
        if (to == 0xffffffffu) continue;
        const V* s = src + r * src_stride;
        V* d = dst + (i64)to * dst_stride;
        for (int c = 0; c < cols; c++) d[c] = s[c];
    }
}

// rows of up to 8 four-byte words into 32-byte destination rows: the whole sector is written (padding words zero), so the
// scattered stores are full-sector writes instead of partial ones that make the memory system read the sector first
__global__ void __launch_bounds__(256) scatter_rows8_kernel(const u32* __restrict__ src, const u32* __restrict__ perm, i64 rows, int cols,
                                                            int src_stride, uint4* __restrict__ dst) {
    for (i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (i64)gridDim.x * blockDim.x) {
        const u32 to = perm[r];
        if (to == 0xffffffffu) continue;
        const u32* s = src + r * src_stride;
        u32 v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) v[c] = c < cols ? s[c] : 0u;
        dst[2 * (i64)to] = make_uint4(v[0], v[1], v[2], v[3]);
        dst[2 * (i64)to + 1] = make_uint4(v[4], v[5], v[6], v[7]);
    }
}

extern "C" int al_scatter_rows(const void* src, const uint32_t* perm, int64_t rows, int cols, int src_stride, int dst_stride,
                               int elem_bytes, void* dst, void* stream) {
    AL_REQUIRE(elem_bytes == 4 || elem_bytes == 8, AL_EINVAL, "al_scatter_rows: elem_bytes must be 4 or 8");
    if (rows <= 0) return AL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (elem_bytes == 4 && dst_stride == 8 && cols <= 8 && ((uintptr_t)dst & 31) == 0)
        scatter_rows8_kernel<<<al_grid(rows, 256), 256, 0, st>>>((const u32*)src, perm, rows, cols, src_stride, (uint4*)dst);
    else if (elem_bytes == 4)
        scatter_rows_kernel<u32><<<al_grid(rows, 256), 256, 0, st>>>((const u32*)src, perm, rows, cols, src_stride, dst_stride, (u32*)dst);
    else
        scatter_rows_kernel<u64><<<al_grid(rows, 256), 256, 0, st>>>((const u64*)src, perm, rows, cols, src_stride, dst_stride, (u64*)dst);
    AL_CHECK_LAUNCH();
    return AL_OK;
}

// FP32 FMA issue-rate probe (roofline denominator of the kNN kernel)
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, float a, float b, int iters) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678f) out[0] = s;
}

extern "C" int al_fp32_peak_tflops(double* tflops_host, void* scratch, size_t scratch_bytes, void* stream) {
    AL_REQUIRE(tflops_host != nullptr && scratch != nullptr && scratch_bytes >= 4, AL_EINVAL, "al_fp32_peak_tflops: needs 4 bytes of device scratch");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    AL_CUDA(cudaGetDevice(&dev));
    AL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* out = (float*)scratch;
    cudaEvent_t e0, e1;
    AL_CUDA(cudaEventCreate(&e0));
    AL_CUDA(cudaEventCreate(&e1));
    const int iters = 8192, blocks = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        AL_CUDA(cudaEventRecord(e0, st));
        fp32_probe_kernel<<<blocks, 256, 0, st>>>(out, 1.0001f, 0.5f, iters);
        AL_CUDA(cudaEventRecord(e1, st));
        AL_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        AL_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double tf = (double)blocks * 256 * iters * 8 * 2 / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_host = best;
    return AL_OK;
}

// ---------------------------------------------------------------------------
// f1: groups_sigs = norm.isf(beta.sf(prominences, a, b))  (astrolink.py:858)
// Element-wise, float64: the survival function of Beta(a, b) by the continued
// fraction of the regularised incomplete beta function (modified Lentz), taken
// on the side that keeps the small tail accurate, then the inverse normal
// survival function.  The Beta-model FIT stays on the host (north-star item 5).
// ---------------------------------------------------------------------------
__device__ static double beta_cf(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 1000; m++) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

__device__ static double beta_sf(double x, double a, double b, double lbeta) {
    if (!(x > 0.0)) return x != x ? x : 1.0;
    if (x >= 1.0) return 0.0;
    const double front = exp(a * log(x) + b * log1p(-x) - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0)) return 1.0 - front * beta_cf(a, b, x) / a;   // cdf is the small side
    return front * beta_cf(b, a, 1.0 - x) / b;                                       // sf is the small side
}

template <typename T>
__global__ void __launch_bounds__(256) significance_kernel(const T* __restrict__ prom, i64 count, double a, double b, double lbeta,
                                                           double* __restrict__ out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double sf = beta_sf((double)prom[i], a, b, lbeta);
    out[i] = -normcdfinv(sf);    // norm.isf(p) = -ndtri(p)
}

extern "C" int al_significance(const void* prominences, int dtype, int64_t count, double a, double b, double* sig_out, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_significance: bad dtype");
    AL_REQUIRE(count >= 0 && a > 0.0 && b > 0.0, AL_EINVAL, "al_significance: bad arguments");
    if (count == 0) return AL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    unsigned grid = (unsigned)((count + 255) / 256);
    if (dtype == AL_F32) significance_kernel<float><<<grid, 256, 0, st>>>((const float*)prominences, count, a, b, lbeta, sig_out);
    else significance_kernel<double><<<grid, 256, 0, st>>>((const double*)prominences, count, a, b, lbeta, sig_out);
    AL_CHECK_LAUNCH();
    return AL_OK;
}

// ---------------------------------------------------------------------------
// f1: objective of the prominence-model fit (astrolink.py:880-885, 896-907).
// The optimiser (SciPy L-BFGS-B, jac='3-point') stays on the host; only the O(G)
// evaluation of the Wasserstein-1 distance between the numerical cdf of Beta(a, b)
// and the empirical cdf runs here.  al_w1_prepare sorts the prominences and builds
// the per-gap tables once; al_w1_eval is called once per objective evaluation.
// ---------------------------------------------------------------------------
struct W1Ws {
    double *p_sorted, *log_mid, *log_omm, *widths, *mass, *partial, *keys_in;
    void* cub_tmp;
    size_t cub_bytes;
};
constexpr int W1_BLOCKS = 592;

static size_t w1_cub_bytes(i64 G) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (const double*)nullptr, (double*)nullptr, G, 0, 64, (cudaStream_t)0);
    cub::DeviceScan::InclusiveSum(nullptr, b, (const double*)nullptr, (double*)nullptr, G, (cudaStream_t)0);
    return a > b ? a : b;
}
static bool w1_layout(void* ws, size_t ws_bytes, i64 G, W1Ws* w) {
    Arena ar(ws, ws_bytes);
    w->p_sorted = ar.get<double>(G);
    w->log_mid = ar.get<double>(G);
    w->log_omm = ar.get<double>(G);
    w->widths = ar.get<double>(G);
    w->mass = ar.get<double>(G);
    w->keys_in = ar.get<double>(G);
    w->partial = ar.get<double>(W1_BLOCKS);
    w->cub_bytes = w1_cub_bytes(G);
    w->cub_tmp = ar.get<char>(w->cub_bytes);
    return ar.ok;
}
extern "C" size_t al_w1_workspace_bytes(int64_t G) {
    ArenaSizer s;
    for (int i = 0; i < 6; i++) s.add<double>(G);
    s.add<double>(W1_BLOCKS);
    s.add<char>(w1_cub_bytes(G));
    return s.off + 256;
}

template <typename T>
__global__ void __launch_bounds__(256) w1_column_kernel(const T* __restrict__ prom, i64 G, int stride, int col, double* __restrict__ out) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (i64)gridDim.x * blockDim.x) out[i] = (double)prom[i * stride + col];
}
__global__ void __launch_bounds__(256) w1_tables_kernel(const double* __restrict__ p, i64 G, double* __restrict__ log_mid,
                                                        double* __restrict__ log_omm, double* __restrict__ widths) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (i64)gridDim.x * blockDim.x) {
        const double lo = i == 0 ? 0.0 : p[i - 1], hi = p[i];
        const double w = hi - lo, mid = 0.5 * (hi + lo);
        widths[i] = w;
        // a midpoint of exactly 0 (or 1) only occurs in a zero-width gap, whose mass is 0 anyway
        log_mid[i] = w > 0.0 ? log(mid) : 0.0;
        log_omm[i] = w > 0.0 ? log1p(-mid) : 0.0;
    }
}
__global__ void __launch_bounds__(256) w1_mass_kernel(const double* __restrict__ log_mid, const double* __restrict__ log_omm,
                                                      const double* __restrict__ widths, i64 G, double am1, double bm1, double lbeta,
                                                      double* __restrict__ mass) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (i64)gridDim.x * blockDim.x)
        mass[i] = exp(am1 * log_mid[i] + bm1 * log_omm[i] - lbeta) * widths[i];
}
__global__ void __launch_bounds__(256) w1_distance_kernel(const double* __restrict__ cdf, const double* __restrict__ widths, i64 G,
                                                          double q0, double qstep, double* __restrict__ partial) {
    double acc = 0.0;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < G; i += (i64)gridDim.x * blockDim.x)
        acc += fabs(cdf[i] - (q0 + (double)i * qstep)) * widths[i];
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += sm[w];
        partial[blockIdx.x] = v;
    }
}
__global__ void w1_finish_kernel(const double* __restrict__ partial, int blocks, double* out) {
    double v = 0.0;
    for (int b = 0; b < blocks; b++) v += partial[b];
    out[0] = v;
}

extern "C" int al_w1_prepare(const void* prominences, int dtype, int64_t G, int stride, int col, void* ws, size_t ws_bytes,
                             double* sorted_host, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_w1_prepare: bad dtype");
    AL_REQUIRE(G >= 1 && stride >= 1 && col >= 0 && col < stride, AL_EINVAL, "al_w1_prepare: bad shape");
    W1Ws w;
    AL_REQUIRE(w1_layout(ws, ws_bytes, G, &w), AL_ENOMEM, "al_w1_prepare: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = al_grid(G, 256, W1_BLOCKS);
    if (dtype == AL_F32) w1_column_kernel<float><<<grid, 256, 0, st>>>((const float*)prominences, G, stride, col, w.keys_in);
    else w1_column_kernel<double><<<grid, 256, 0, st>>>((const double*)prominences, G, stride, col, w.keys_in);
    AL_CHECK_LAUNCH();
    size_t tb = w.cub_bytes;
    AL_CUDA(cub::DeviceRadixSort::SortKeys(w.cub_tmp, tb, w.keys_in, w.p_sorted, G, 0, 64, st));
    w1_tables_kernel<<<grid, 256, 0, st>>>(w.p_sorted, G, w.log_mid, w.log_omm, w.widths);
    AL_CHECK_LAUNCH();
    if (sorted_host) AL_CUDA(cudaMemcpyAsync(sorted_host, w.p_sorted, sizeof(double) * G, cudaMemcpyDeviceToHost, st));
    AL_CUDA(cudaStreamSynchronize(st));
    return AL_OK;
}

extern "C" int al_w1_eval(void* ws, size_t ws_bytes, int64_t G, double a, double b, double* value_host, void* stream) {
    AL_REQUIRE(G >= 1 && a > 0.0 && b > 0.0 && value_host != nullptr, AL_EINVAL, "al_w1_eval: bad arguments");
    W1Ws w;
    AL_REQUIRE(w1_layout(ws, ws_bytes, G, &w), AL_ENOMEM, "al_w1_eval: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = al_grid(G, 256, W1_BLOCKS);
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    w1_mass_kernel<<<grid, 256, 0, st>>>(w.log_mid, w.log_omm, w.widths, G, a - 1.0, b - 1.0, lbeta, w.mass);
    AL_CHECK_LAUNCH();
    size_t tb = w.cub_bytes;
    AL_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, tb, w.mass, w.mass, G, st));
    const double half = 1.0 / (2.0 * (double)G);
    const double qstep = G > 1 ? (1.0 - 2.0 * half) / (double)(G - 1) : 0.0;
    w1_distance_kernel<<<grid, 256, 0, st>>>(w.mass, w.widths, G, half, qstep, w.partial);
    AL_CHECK_LAUNCH();
    w1_finish_kernel<<<1, 1, 0, st>>>(w.partial, grid, w.keys_in);
    AL_CHECK_LAUNCH();
    AL_CUDA(cudaMemcpyAsync(value_host, w.keys_in, sizeof(double), cudaMemcpyDeviceToHost, st));
    AL_CUDA(cudaStreamSynchronize(st));
    return AL_OK;
}

// ---- batched evaluation: B parameter pairs per call, two launches and one host read for all of them.
// SciPy's L-BFGS-B with jac='3-point' evaluates the objective at x and at 2 d points around it, one call at a time; the
// host wrapper predicts those points and asks for all of them at once (a wrong prediction only costs a cache miss).
// The range of gaps is cut into chunks; pass 1 sums the probability mass of every chunk, pass 2 re-derives the masses of
// its chunk on top of the preceding chunks' total (no global scan) and integrates |cdf - empirical cdf|.  All sums run in
// a fixed order, so a given (a, b) always yields the same bits.
constexpr int W1B_THREADS = 256;
constexpr int W1B_MAX = 16;          // parameter pairs per call

struct W1Batch { double am1[W1B_MAX], bm1[W1B_MAX], lbeta[W1B_MAX]; };

template <bool DIST>
__global__ void __launch_bounds__(W1B_THREADS) w1_batch_kernel(const double* __restrict__ log_mid, const double* __restrict__ log_omm,
                                                               const double* __restrict__ widths, i64 G, i64 chunk, int n_chunks,
                                                               W1Batch prm, double q0, double qstep,
                                                               const double* __restrict__ chunk_mass, double* __restrict__ out) {
    const int c = blockIdx.x, b = blockIdx.y;
    const i64 lo = (i64)c * chunk, hi = lo + chunk < G ? lo + chunk : G;
    const i64 per = (chunk + W1B_THREADS - 1) / W1B_THREADS;
    const i64 t0 = lo + (i64)threadIdx.x * per, t1 = t0 + per < hi ? t0 + per : hi;
    const double am1 = prm.am1[b], bm1 = prm.bm1[b], lbeta = prm.lbeta[b];
    double local = 0.0;
    for (i64 i = t0; i < t1; i++) local += exp(am1 * log_mid[i] + bm1 * log_omm[i] - lbeta) * widths[i];
    __shared__ double sm[W1B_THREADS];
    sm[threadIdx.x] = local;
    __syncthreads();
    if (!DIST) {
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int k = 0; k < W1B_THREADS; k++) v += sm[k];
            out[(i64)b * n_chunks + c] = v;
        }
        return;
    }
    // mass before this thread's items: the preceding chunks (fixed order), then the preceding threads of this chunk
    double before = 0.0;
    for (int k = 0; k < c; k++) before += chunk_mass[(i64)b * n_chunks + k];
    for (int k = 0; k < (int)threadIdx.x; k++) before += sm[k];
    double cdf = before, acc = 0.0;
    for (i64 i = t0; i < t1; i++) {
        const double w = widths[i];
        cdf += exp(am1 * log_mid[i] + bm1 * log_omm[i] - lbeta) * w;
        acc += fabs(cdf - (q0 + (double)i * qstep)) * w;
    }
    __syncthreads();
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int k = 0; k < W1B_THREADS; k++) v += sm[k];
        out[(i64)b * n_chunks + c] = v;
    }
}

extern "C" int al_w1_batch_chunks(int64_t G) {
    i64 c = (G + 4095) / 4096;
    if (c < 1) c = 1;
    if (c > 128) c = 128;
    return (int)c;
}

extern "C" int al_w1_eval_batch(void* ws, size_t ws_bytes, int64_t G, const double* ab_host, int B, double* values_host, void* stream) {
    AL_REQUIRE(G >= 1 && ab_host != nullptr && values_host != nullptr && B >= 1 && B <= W1B_MAX, AL_EINVAL, "al_w1_eval_batch: bad arguments");
    W1Ws w;
    AL_REQUIRE(w1_layout(ws, ws_bytes, G, &w), AL_ENOMEM, "al_w1_eval_batch: workspace too small");
    const int n_chunks = al_w1_batch_chunks(G);
    // scratch: the mass / keys_in arrays of the workspace (G doubles each) are free between evaluations
    AL_REQUIRE((i64)B * n_chunks <= G, AL_EINVAL, "al_w1_eval_batch: G too small for the batched path (use al_w1_eval)");
    double* chunk_mass = w.mass;
    double* chunk_dist = w.keys_in;
    W1Batch prm;
    for (int b = 0; b < B; b++) {
        const double a = ab_host[2 * b], bb = ab_host[2 * b + 1];
        AL_REQUIRE(a > 0.0 && bb > 0.0, AL_EINVAL, "al_w1_eval_batch: parameters must be positive");
        prm.am1[b] = a - 1.0;
        prm.bm1[b] = bb - 1.0;
        prm.lbeta[b] = lgamma(a) + lgamma(bb) - lgamma(a + bb);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const i64 chunk = (G + n_chunks - 1) / n_chunks;
    const double half = 1.0 / (2.0 * (double)G);
    const double qstep = G > 1 ? (1.0 - 2.0 * half) / (double)(G - 1) : 0.0;
    dim3 grid((unsigned)n_chunks, (unsigned)B);
    w1_batch_kernel<false><<<grid, W1B_THREADS, 0, st>>>(w.log_mid, w.log_omm, w.widths, G, chunk, n_chunks, prm, half, qstep, nullptr, chunk_mass);
    AL_CHECK_LAUNCH();
    w1_batch_kernel<true><<<grid, W1B_THREADS, 0, st>>>(w.log_mid, w.log_omm, w.widths, G, chunk, n_chunks, prm, half, qstep, chunk_mass, chunk_dist);
    AL_CHECK_LAUNCH();
    double part[W1B_MAX * 128];
    AL_CUDA(cudaMemcpyAsync(part, chunk_dist, sizeof(double) * (size_t)B * n_chunks, cudaMemcpyDeviceToHost, st));
    AL_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < B; b++) {
        double v = 0.0;
        for (int c = 0; c < n_chunks; c++) v += part[(size_t)b * n_chunks + c];
        values_host[b] = v;
    }
    return AL_OK;
}

// peer access from the current device to `peer_device` (multi-GPU path: kernels on rank r store into rank 0's buffers)
extern "C" int al_enable_peer_access(int peer_device) {
    int dev = 0;
    AL_CUDA(cudaGetDevice(&dev));
    if (dev == peer_device) return AL_OK;
    int can = 0;
    AL_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    AL_REQUIRE(can == 1, AL_EINVAL, "al_enable_peer_access: device %d cannot access device %d", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return AL_OK; }
    AL_CUDA(e);
    return AL_OK;
}

// ---------------------------------------------------------------------------
// Peer buffers of the multi-GPU path: allocated on the consumer GPU (rank 0) and mapped
// into the other ranks' processes through CUDA IPC, opened with the ACCESSING device
// current so that its kernels can store into them over NVLink.
// ---------------------------------------------------------------------------
extern "C" int al_peer_alloc(size_t bytes, void** ptr_host, unsigned char* handle64_host) {
    AL_REQUIRE(bytes > 0 && ptr_host != nullptr && handle64_host != nullptr, AL_EINVAL, "al_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    AL_CUDA(cudaMalloc(&p, bytes));          // set-up path, not the hot path
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        al_set_error("al_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
        return AL_ECUDA;
    }
    memcpy(handle64_host, &h, 64);
    *ptr_host = p;
    return AL_OK;
}
extern "C" int al_peer_open(const unsigned char* handle64_host, void** ptr_host) {
    AL_REQUIRE(handle64_host != nullptr && ptr_host != nullptr, AL_EINVAL, "al_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void* p = nullptr;
    AL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_host = p;
    return AL_OK;
}
// device-to-device copy on `stream`; either pointer may be a peer-mapped buffer (al_peer_open)
extern "C" int al_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
    AL_REQUIRE(dst != nullptr && src != nullptr, AL_EINVAL, "al_peer_copy: bad arguments");
    if (bytes == 0) return AL_OK;
    AL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return AL_OK;
}
extern "C" int al_peer_release(void* ptr, int owner) {
    if (ptr == nullptr) return AL_OK;
    if (owner) AL_CUDA(cudaFree(ptr));
    else AL_CUDA(cudaIpcCloseMemHandle(ptr));
    return AL_OK;
}
