// Spatial index build: replaces pykdtree's KDTree(P_transform)
// (reference astrolink/astrolink.py:270).
//
// A balanced axis-aligned split tree is built level-synchronously: every round
// sorts all points by (node, coordinate along the node's widest dimension) with
// one CUB radix sort, then every node of more than one leaf is cut at the
// multiple of the largest power of two inside its leaf range.  That rule makes
// every aligned block of 32^h leaves a subtree, so the 32-ary box hierarchy the
// search walks has tight boxes.  The tree only decides the ORDER of the points;
// the search result does not depend on it (boxes are exact bounds).
#include <cub/cub.cuh>

#include <mutex>
#include <unordered_map>

#include "index.cuh"
#include <stdlib.h>

constexpr int MAXD = 16;

// ---------------------------------------------------------------------------
// header registry
// ---------------------------------------------------------------------------
static std::mutex g_reg_mu;
static std::unordered_map<const void*, IndexHeader> g_reg;

bool index_lookup(const void* index, IndexHeader* out) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_reg.find(index);
    if (it == g_reg.end()) return false;
    *out = it->second;
    return true;
}
void index_register(const void* index, const IndexHeader& h) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_reg[index] = h;
}
extern "C" void al_index_release(const void* index) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    g_reg.erase(index);
}

// ---------------------------------------------------------------------------
__host__ __device__ static inline u32 split_mid(u32 a, u32 b) {
    u32 x = a ^ (b - 1);
#ifdef __CUDA_ARCH__
    int h = 31 - __clz(x);
#else
    int h = 31 - __builtin_clz(x);
#endif
    return ((b - 1) >> h) << h;
}

static int tree_depth(i64 n_leaves) {
    if (n_leaves <= 1) return 0;
    int p = 63 - __builtin_clzll((unsigned long long)(n_leaves - 1));
    int r = tree_depth(n_leaves - ((i64)1 << p));
    return 1 + (p > r ? p : r);
}

__global__ void tree_init_kernel(u32* perm, u32* nodeA, u32* nodeB, i64 n, i64 n_leaves) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 p = i; p < n; p += stride) perm[p] = (u32)p;
    for (i64 l = i; l < n_leaves; l += stride) { nodeA[l] = 0; nodeB[l] = (u32)n_leaves; }
}

constexpr int BOUNDS_CHUNK = 16;   // leaves per warp

template <typename T>
__global__ void __launch_bounds__(256) tree_node_bounds_kernel(const T* __restrict__ P, const u32* __restrict__ perm,
                                                               const u32* __restrict__ nodeA, const u32* __restrict__ nodeB,
                                                               u32* bbox_lo, u32* bbox_hi, i64 n, i64 n_leaves, int d) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 block_leaf0 = (i64)blockIdx.x * (256 / 32) * BOUNDS_CHUNK;
    if (block_leaf0 >= n_leaves) return;
    const i64 block_leaf1 = min(block_leaf0 + (256 / 32) * BOUNDS_CHUNK, n_leaves) - 1;
    // early rounds: the whole block lies inside one (large) node -> reduce in the block, one atomic set per block
    const u32 a0 = nodeA[block_leaf0];
    const bool uniform = a0 == nodeA[block_leaf1];
    const i64 leaf0 = block_leaf0 + (i64)wib * BOUNDS_CHUNK;
    float mn[MAXD], mx[MAXD];
#pragma unroll
    for (int j = 0; j < MAXD; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    auto accumulate = [&](i64 leaf) {
        const i64 p = leaf * LEAF + lane;
        if (p < n) {
            const T* row = P + (i64)perm[p] * d;
#pragma unroll
            for (int j = 0; j < MAXD; j++)
                if (j < d) {
                    const float v = (float)row[j];
                    mn[j] = fminf(mn[j], v);
                    mx[j] = fmaxf(mx[j], v);
                }
        }
    };
    auto warp_reduce = [&](int j, float& a, float& b) {
        a = mn[j];
        b = mx[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
    };
    if (uniform) {
        if (nodeB[block_leaf0] - a0 <= 1) return;      // a finished node: nothing to do
        __shared__ float smn[256 / 32][MAXD], smx[256 / 32][MAXD];
        for (int c = 0; c < BOUNDS_CHUNK; c++) {
            const i64 leaf = leaf0 + c;
            if (leaf < n_leaves) accumulate(leaf);
        }
#pragma unroll
        for (int j = 0; j < MAXD; j++)
            if (j < d) {
                float a, b;
                warp_reduce(j, a, b);
                if (lane == 0) { smn[wib][j] = a; smx[wib][j] = b; }
            }
        __syncthreads();
        if (threadIdx.x < d) {
            float a = smn[0][threadIdx.x], b = smx[0][threadIdx.x];
            for (int w = 1; w < 256 / 32; w++) { a = fminf(a, smn[w][threadIdx.x]); b = fmaxf(b, smx[w][threadIdx.x]); }
            atomicMin(&bbox_lo[(i64)a0 * d + threadIdx.x], f32_to_ordered(a));
            atomicMax(&bbox_hi[(i64)a0 * d + threadIdx.x], f32_to_ordered(b));
        }
        return;
    }
    // general case: every warp flushes per run of leaves of the same node
    if (leaf0 >= n_leaves) return;
    i64 cur = -1;
    auto flush = [&]() {
        if (cur < 0) return;
#pragma unroll
        for (int j = 0; j < MAXD; j++) {
            if (j < d) {
                float a, b;
                warp_reduce(j, a, b);
                if (lane == 0) {
                    atomicMin(&bbox_lo[cur * d + j], f32_to_ordered(a));
                    atomicMax(&bbox_hi[cur * d + j], f32_to_ordered(b));
                }
            }
        }
    };
    for (int c = 0; c < BOUNDS_CHUNK; c++) {
        const i64 leaf = leaf0 + c;
        if (leaf >= n_leaves) break;
        const u32 a = nodeA[leaf];
        if (nodeB[leaf] - a <= 1) continue;
        if ((i64)a != cur) {
            flush();
            cur = a;
#pragma unroll
            for (int j = 0; j < MAXD; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
        }
        accumulate(leaf);
    }
    flush();
}

// After a round's sort: the point rows and ids follow the new order into the other buffer -- so every later pass of the
// build reads them sequentially, and this gather is local to the node that was sorted -- and, in the same pass, the tight
// boxes of the nodes of the NEXT round are accumulated (the rows are in registers).  Work decomposition as in
// tree_node_bounds_kernel: a warp per BOUNDS_CHUNK leaves, lane = point of a leaf.
// V8: rows are a whole number of 8-byte units; the gather takes 64-bit loads (it is bound by sector requests, one per load
// and lane), NF leaves are in flight per step (the pass is latency bound: slot -> row is a dependent pair of loads), and
// the leaves' rows -- contiguous in the destination -- leave through shared memory as full-sector stores.  MD bounds d.
template <typename T, bool V8, int MD, int NF>
__global__ void __launch_bounds__(256) tree_permute_bounds_kernel(const T* __restrict__ Psrc, const u32* __restrict__ ids_src,
                                                                  const u32* __restrict__ slots, T* __restrict__ Pdst,
                                                                  u32* __restrict__ ids_dst, const u32* __restrict__ nodeA,
                                                                  const u32* __restrict__ nodeB, u32* bbox_lo, u32* bbox_hi, i64 n,
                                                                  i64 n_leaves, int d, int want_bounds) {
    constexpr int EPU = 8 / (int)sizeof(T);         // elements per 8-byte unit
    constexpr int MU = MD / EPU;                     // units per row, at most
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    __shared__ uint2 stage[V8 ? 256 / 32 : 1][V8 ? NF * LEAF * MU : 1];
    const i64 leaf0 = ((i64)blockIdx.x * (256 / 32) + wib) * BOUNDS_CHUNK;
    if (leaf0 >= n_leaves) return;
    float mn[MD], mx[MD];
#pragma unroll
    for (int j = 0; j < MD; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    i64 cur = -1;
    auto flush = [&]() {
        if (cur < 0) return;
#pragma unroll
        for (int j = 0; j < MD; j++) {
            if (j < d) {
                float a = mn[j], b = mx[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
                    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
                }
                if (lane == 0) {
                    atomicMin(&bbox_lo[cur * d + j], f32_to_ordered(a));
                    atomicMax(&bbox_hi[cur * d + j], f32_to_ordered(b));
                }
            }
        }
    };
    // enters a leaf's node: true if its box is being accumulated
    auto enter = [&](i64 leaf) -> bool {
        const u32 a = nodeA[leaf];
        const bool active = want_bounds && nodeB[leaf] - a > 1;
        if (active && (i64)a != cur) {
            flush();
            cur = a;
#pragma unroll
            for (int j = 0; j < MD; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
        }
        return active;
    };
    if (V8) {
        const int U = d / EPU;                       // units per row
        uint2* st = stage[wib];
        for (int c = 0; c < BOUNDS_CHUNK; c += NF) {
            const i64 leaf = leaf0 + c;
            if (leaf >= n_leaves) break;
            bool valid[NF];
            u32 sl[NF];
            uint2 r[NF][MU];
#pragma unroll
            for (int f = 0; f < NF; f++) {
                const i64 pf = (leaf + f) * LEAF + lane;
                valid[f] = leaf + f < n_leaves && pf < n;
                sl[f] = valid[f] ? slots[pf] : 0u;
            }
#pragma unroll
            for (int f = 0; f < NF; f++) {
                const uint2* row = reinterpret_cast<const uint2*>(Psrc + (i64)sl[f] * d);
#pragma unroll
                for (int u = 0; u < MU; u++)
                    if (u < U) r[f][u] = valid[f] ? __ldg(row + u) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int f = 0; f < NF; f++)
                if (valid[f]) ids_dst[(leaf + f) * LEAF + lane] = ids_src[sl[f]];
#pragma unroll
            for (int f = 0; f < NF; f++) {
                if (leaf + f < n_leaves && enter(leaf + f)) {
#pragma unroll
                    for (int u = 0; u < MU; u++)
                        if (u < U && valid[f]) {
                            if (sizeof(T) == 4) {
                                const float v0 = __uint_as_float(r[f][u].x), v1 = __uint_as_float(r[f][u].y);
                                mn[2 * u] = fminf(mn[2 * u], v0);
                                mx[2 * u] = fmaxf(mx[2 * u], v0);
                                mn[(2 * u + 1) % MD] = fminf(mn[(2 * u + 1) % MD], v1);
                                mx[(2 * u + 1) % MD] = fmaxf(mx[(2 * u + 1) % MD], v1);
                            } else {
                                const float v0 = (float)__longlong_as_double((long long)(((unsigned long long)r[f][u].y << 32) | r[f][u].x));
                                mn[u] = fminf(mn[u], v0);
                                mx[u] = fmaxf(mx[u], v0);
                            }
                        }
                }
#pragma unroll
                for (int u = 0; u < MU; u++)
                    if (u < U) st[(f * LEAF + lane) * U + u] = r[f][u];
            }
            __syncwarp();
            i64 rows_here = n - leaf * LEAF;
            i64 cap = (n_leaves - leaf) * LEAF;
            cap = cap < (i64)NF * LEAF ? cap : (i64)NF * LEAF;
            rows_here = rows_here < cap ? rows_here : cap;
            uint2* out = reinterpret_cast<uint2*>(Pdst + leaf * LEAF * d);
            for (int k = lane; k < (int)rows_here * U; k += 32) out[k] = st[k];
            __syncwarp();
        }
    } else {
        for (int c = 0; c < BOUNDS_CHUNK; c++) {
            const i64 leaf = leaf0 + c;
            if (leaf >= n_leaves) break;
            const bool active = enter(leaf);
            const i64 p = leaf * LEAF + lane;
            if (p < n) {
                const u32 s = slots[p];
                ids_dst[p] = ids_src[s];
                const T* row = Psrc + (i64)s * d;
                T* out = Pdst + p * d;
#pragma unroll
                for (int j = 0; j < MD; j++)
                    if (j < d) {
                        const T v = row[j];
                        out[j] = v;
                        if (active) {
                            mn[j] = fminf(mn[j], (float)v);
                            mx[j] = fmaxf(mx[j], (float)v);
                        }
                    }
            }
        }
    }
    flush();
}

// per node being split: widest axis, its lower bound and the quantisation scale of the sort key
__global__ void __launch_bounds__(256) tree_node_split_info_kernel(const u32* __restrict__ nodeA, const u32* __restrict__ nodeB,
                                                                   const u32* __restrict__ bbox_lo, const u32* __restrict__ bbox_hi,
                                                                   int* __restrict__ sdim, float* __restrict__ slo,
                                                                   float* __restrict__ sscale, i64 n_leaves, int d, int cbits) {
    i64 leaf = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n_leaves) return;
    const u32 a = nodeA[leaf];
    if (a != (u32)leaf || nodeB[leaf] - a <= 1) return;
    auto dec = [](u32 o) { u32 b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; return __uint_as_float(b); };
    int dim = 0;
    float best = -1.f, blo = 0.f, bext = 0.f;
    for (int j = 0; j < d; j++) {
        const float lo = dec(bbox_lo[(i64)a * d + j]), hi = dec(bbox_hi[(i64)a * d + j]);
        float ext = hi - lo;
        ext = ext > 0.f ? ext : 0.f;
        if (ext > best) { best = ext; dim = j; blo = lo; bext = ext; }
    }
    sdim[leaf] = dim;
    slo[leaf] = blo;
    sscale[leaf] = bext > 0.f ? (float)((1u << cbits) - 1u) / bext : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(256) tree_make_keys_kernel(const T* __restrict__ P, const u32* __restrict__ perm,
                                                             const u32* __restrict__ nodeA, const u32* __restrict__ nodeB,
                                                             const int* __restrict__ sdim, const float* __restrict__ slo,
                                                             const float* __restrict__ sscale, const u32* __restrict__ noderank,
                                                             u32* keys, i64 n, int d, int cbits) {
    i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const i64 leaf = p / LEAF;
    const u32 a = nodeA[leaf];
    u32 low;
    if (nodeB[leaf] - a > 1) {
        // the coordinate along the node's widest axis, quantised to cbits within the node's box: the tree only
        // needs a balanced partition of reasonable quality, so ties inside one quantum may fall either side
        const float x = (float)P[(perm ? (i64)perm[p] : p) * d + sdim[a]];      // perm == nullptr: the rows are in position order
        const float top = (float)((1u << cbits) - 1u);
        float qf = (x - slo[a]) * sscale[a];
        qf = qf < 0.f ? 0.f : (qf > top ? top : qf);
        low = (u32)qf;
    } else {
        low = (u32)(p - leaf * LEAF);
    }
    // nodes are numbered in order (rank of their first leaf among the node starts): a round that has made 2^L nodes so far
    // needs only L node bits in the key, so the early rounds sort fewer digits
    keys[p] = (noderank[a] << cbits) | low;
}

__global__ void __launch_bounds__(256) tree_node_start_kernel(const u32* __restrict__ nodeA, u32* __restrict__ flag, i64 n_leaves) {
    i64 l = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n_leaves) flag[l] = nodeA[l] == (u32)l ? 1u : 0u;
}

// After a round's sort every node is ordered along its widest axis, so the same order serves `levels`
// successive cuts: the node is divided into up to 2^levels quantile slabs at the aligned split positions.
__global__ void tree_split_kernel(u32* nodeA, u32* nodeB, i64 n_leaves, int levels) {
    i64 l = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_leaves) return;
    u32 a = nodeA[l], b = nodeB[l];
    for (int s = 0; s < levels && b - a > 1; s++) {
        u32 mid = split_mid(a, b);
        if ((u32)l < mid) b = mid; else a = mid;
    }
    nodeA[l] = a;
    nodeB[l] = b;
}


// ---------------------------------------------------------------------------
// Bottom of the tree in shared memory: one CTA finishes an aligned block of
// BOT_LEAVES leaves (a subtree, by the aligned split rule) -- all remaining
// split levels, one level per pass, tight boxes and widest-axis choice per node
// exactly as the global rounds do -- without touching global memory in between.
// Points stay in place in shared memory; only an order array is sorted
// (cub::BlockRadixSort over (node, quantised coordinate) keys).
// ---------------------------------------------------------------------------
constexpr int BOT_LEAVES = 128;
constexpr int BOT_POINTS = BOT_LEAVES * LEAF;     // 4096
#ifndef AL_BOT_THREADS
#define AL_BOT_THREADS 512
#endif
constexpr int BOT_THREADS = AL_BOT_THREADS;
constexpr int BOT_ITEMS = BOT_POINTS / BOT_THREADS;
#ifndef AL_BOT_RADIX_BITS
#define AL_BOT_RADIX_BITS 4
#endif
typedef cub::BlockRadixSort<u32, BOT_THREADS, BOT_ITEMS, u32, AL_BOT_RADIX_BITS> BotSort;

static size_t tree_bottom_smem(int d) {
    return (size_t)BOT_POINTS * d * 4 + (size_t)BOT_POINTS * 4 * 2 + (size_t)BOT_LEAVES * (2 + 2 * d + 4) * 4 + sizeof(BotSort::TempStorage) + 64;
}

template <typename T>
__global__ void __launch_bounds__(BOT_THREADS) tree_bottom_kernel(const T* __restrict__ Pord, T* __restrict__ Pout, u32* perm, i64 n, i64 n_leaves,
                                                                  int d, int cb) {
    extern __shared__ __align__(16) unsigned char bsm[];
    float* coords = (float*)bsm;                               // [d][BOT_POINTS]
    u32* ids = (u32*)(coords + (size_t)d * BOT_POINTS);        // [BOT_POINTS] original point ids
    u32* order = ids + BOT_POINTS;                             // [BOT_POINTS] slot -> local point
    u32* nA = order + BOT_POINTS;                              // [BOT_LEAVES] node range of every leaf (block-relative)
    u32* nB = nA + BOT_LEAVES;
    u32* blo = nB + BOT_LEAVES;                                // [BOT_LEAVES][d] boxes (ordered uints), by node start
    u32* bhi = blo + BOT_LEAVES * d;
    int* sdim = (int*)(bhi + BOT_LEAVES * d);                  // [BOT_LEAVES]
    float* slo = (float*)(sdim + BOT_LEAVES);
    float* ssc = slo + BOT_LEAVES;
    u32* nrank = (u32*)(ssc + BOT_LEAVES);                     // [BOT_LEAVES] rank of the node starting at a leaf
    BotSort::TempStorage* tmp = (BotSort::TempStorage*)((((uintptr_t)(nrank + BOT_LEAVES)) + 15) & ~(uintptr_t)15);
    __shared__ u32 wtot[4];

    const int tid = threadIdx.x;
    const i64 leaf0 = (i64)blockIdx.x * BOT_LEAVES;
    const i64 p0 = leaf0 * LEAF;
    const int nl = (int)min((i64)BOT_LEAVES, n_leaves - leaf0);        // leaves of this block
    const int np = (int)min((i64)BOT_POINTS, n - p0);                  // real points of this block
    for (int s = tid; s < BOT_POINTS; s += BOT_THREADS) {
        u32 id = 0xffffffffu;
        if (s < np) {
            id = perm[p0 + s];
            const T* row = Pord + (p0 + s) * d;        // the rows are in position order (tree_permute_bounds_kernel)
            for (int j = 0; j < d; j++) coords[(size_t)j * BOT_POINTS + s] = (float)row[j];
        }
        ids[s] = id;
        order[s] = (u32)s;
    }
    for (int l = tid; l < BOT_LEAVES; l += BOT_THREADS) { nA[l] = 0; nB[l] = (u32)nl; }
    __syncthreads();
    auto dec = [](u32 o) { u32 b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; return __uint_as_float(b); };
    for (int level = 0; level < 8; level++) {
        // any node left with more than one leaf?
        int active = 0;
        for (int l = tid; l < nl; l += BOT_THREADS) active |= (nB[l] - nA[l] > 1) ? 1 : 0;
        if (!__syncthreads_or(active)) break;
        for (int e = tid; e < BOT_LEAVES * d; e += BOT_THREADS) { blo[e] = 0xffffffffu; bhi[e] = 0u; }
        __syncthreads();
        // tight boxes of the nodes being split: a warp covers one leaf per step, reduces, then one atomic per axis
        for (int s = tid; s < BOT_POINTS; s += BOT_THREADS) {
            const int leaf = s >> 5;
            if (leaf >= nl) break;
            const u32 a = nA[leaf];
            if (nB[leaf] - a > 1) {
                const bool valid = s < np;
                const u32 pt = order[s];
                for (int j = 0; j < d; j++) {
                    const u32 o = valid ? f32_to_ordered(coords[(size_t)j * BOT_POINTS + pt]) : 0u;
                    const u32 mn = __reduce_min_sync(0xffffffffu, valid ? o : 0xffffffffu);
                    const u32 mx = __reduce_max_sync(0xffffffffu, o);
                    if ((tid & 31) == 0) {
                        atomicMin(&blo[a * d + j], mn);
                        atomicMax(&bhi[a * d + j], mx);
                    }
                }
            }
        }
        // nodes are numbered in order so the sort key needs only log2(#nodes) node bits
        if (tid < BOT_LEAVES) {
            const bool head = tid < nl && nA[tid] == (u32)tid;
            const u32 bal = __ballot_sync(0xffffffffu, head);
            nrank[tid] = __popc(bal & ((1u << (tid & 31)) - 1u));
            if ((tid & 31) == 0) wtot[tid >> 5] = __popc(bal);
        }
        __syncthreads();
        u32 n_nodes = 0;
        for (int w = 0; w < BOT_LEAVES / 32; w++) {
            if (tid < BOT_LEAVES && w < (tid >> 5)) nrank[tid] += wtot[w];
            n_nodes += wtot[w];
        }
        const int nbits = n_nodes > 1 ? 32 - __clz(n_nodes - 1) : 0;
        for (int l = tid; l < nl; l += BOT_THREADS) {
            if (nA[l] == (u32)l && nB[l] - nA[l] > 1) {
                int dim = 0;
                float best = -1.f, lo_ = 0.f, ext_ = 0.f;
                for (int j = 0; j < d; j++) {
                    const float lo = dec(blo[l * d + j]), hi = dec(bhi[l * d + j]);
                    float ext = hi - lo;
                    ext = ext > 0.f ? ext : 0.f;
                    if (ext > best) { best = ext; dim = j; lo_ = lo; ext_ = ext; }
                }
                sdim[l] = dim;
                slo[l] = lo_;
                ssc[l] = ext_ > 0.f ? (float)((1u << cb) - 1u) / ext_ : 0.f;
            }
        }
        __syncthreads();
        // keys (blocked arrangement: thread t owns slots t*ITEMS .. t*ITEMS+ITEMS-1), sort, new order
        u32 keys[BOT_ITEMS], vals[BOT_ITEMS];
#pragma unroll
        for (int k = 0; k < BOT_ITEMS; k++) {
            const int s = tid * BOT_ITEMS + k;
            u32 key = 0xffffffffu;
            u32 pt = order[s];
            if (s < np) {
                const int leaf = s >> 5;
                const u32 a = nA[leaf];
                u32 low;
                if (nB[leaf] - a > 1) {
                    const float top = (float)((1u << cb) - 1u);
                    float qf = (coords[(size_t)sdim[a] * BOT_POINTS + pt] - slo[a]) * ssc[a];
                    qf = qf < 0.f ? 0.f : (qf > top ? top : qf);
                    low = (u32)qf;
                } else {
                    low = (u32)(s & 31);
                }
                key = (nrank[a] << cb) | low;
            }
            keys[k] = key;                        // padding slots (all ones) stay last: the sort is stable
            vals[k] = pt;
        }
        __syncthreads();
        BotSort(*tmp).Sort(keys, vals, 0, cb + nbits);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BOT_ITEMS; k++) order[tid * BOT_ITEMS + k] = vals[k];
        for (int l = tid; l < nl; l += BOT_THREADS) {
            u32 a = nA[l], b = nB[l];
            if (b - a > 1) {
                const u32 mid = split_mid(a + (u32)leaf0, b + (u32)leaf0) - (u32)leaf0;   // aligned in GLOBAL leaf indices
                if ((u32)l < mid) b = mid; else a = mid;
                nA[l] = a;
                nB[l] = b;
            }
        }
        __syncthreads();
    }
    // the ids and -- for the tile builder, which then reads sequentially -- the rows, in final order
    for (int s = tid; s < np; s += BOT_THREADS) perm[p0 + s] = ids[order[s]];
    if (sizeof(T) == 4) {
        for (int e = tid; e < np * d; e += BOT_THREADS) {
            const int s = e / d, j = e - s * d;
            Pout[p0 * d + e] = (T)coords[(size_t)j * BOT_POINTS + order[s]];
        }
    } else {
        // float64: shared memory holds rounded copies, so the exact rows are permuted through global memory
        for (int e = tid; e < np * d; e += BOT_THREADS) {
            const int s = e / d, j = e - s * d;
            Pout[p0 * d + e] = Pord[(p0 + order[s]) * d + j];
        }
    }
}

// one warp per leaf: the leaf's points (rows already in position order) into its tile
template <typename T>
__global__ void __launch_bounds__(256) tree_build_tiles_kernel(const T* __restrict__ P, u32* perm, char* tiles,
                                                               T* box0, IndexHeader h) {
    int lane = threadIdx.x & 31;
    i64 leaf = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (leaf >= h.n_leaves) return;
    const int d = h.d;
    i64 p = leaf * LEAF + lane;
    bool valid = p < h.n;
    u32 id = valid ? perm[p] : 0xffffffffu;
    if (!valid) perm[p] = 0xffffffffu;
    char* tile = tiles + leaf * (i64)h.tile_stride;
    T* coords = (T*)(tile + h.tile_hdr_bytes);
    u32* ids = (u32*)(tile + h.tile_ids_off);
    ids[lane] = id;
    const T inf = (T)INFINITY;
    int pb = lane >> 1, half = lane & 1;
    for (int j = 0; j < d; j++) {
        T v = valid ? P[p * d + j] : inf;          // rows in position order
        coords[pb * h.pair_block + 2 * j + half] = v;
        T a = valid ? v : inf, b = valid ? v : -inf;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            T a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
            a = a2 < a ? a2 : a;
            b = b2 > b ? b2 : b;
        }
        if (lane == 0) {
            box0[leaf * h.box_stride + 2 * j] = a;
            box0[leaf * h.box_stride + 2 * j + 1] = b;
        }
    }
    // padding elements of the box row are defined too
    if (lane == 0)
        for (int e = 2 * d; e < h.box_stride; e++) box0[leaf * h.box_stride + e] = (T)0;
    // zero the padding elements of the pair block so the tile is fully defined
    for (int e = 2 * d + half; e < h.pair_block; e += 2) coords[pb * h.pair_block + e] = (T)0;
}

template <typename T>
__global__ void tree_build_level_kernel(const T* __restrict__ cbox, i64 n_child, T* pbox, i64 n_parent, int d, int bs) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_parent * bs) return;
    const i64 node = t / bs;
    const int e = (int)(t - node * bs);      // element of the box row: 2j = lo_j, 2j + 1 = hi_j, then padding
    if (e >= 2 * d) { pbox[t] = (T)0; return; }
    const bool is_hi = e & 1;
    i64 c0 = node * FANOUT, c1 = c0 + FANOUT < n_child ? c0 + FANOUT : n_child;
    T r = is_hi ? -(T)INFINITY : (T)INFINITY;
    for (i64 c = c0; c < c1; c++) {
        const T x = cbox[c * bs + e];
        r = is_hi ? (x > r ? x : r) : (x < r ? x : r);
    }
    pbox[t] = r;
}

// ---------------------------------------------------------------------------
struct BuildWs {
    u32 *keys0, *keys1;
    u32 *vals1;
    u32 *nodeA, *nodeB;
    u32 *bbox_lo, *bbox_hi;
    int* sdim;
    float *slo, *sscale;
    u32 *nflag, *nrank;
    u32 *iota, *slots;
    void *rows0, *rows1;     // the point rows in the current order (ping-pong)
    void* cub_tmp;
    size_t cub_bytes;
};

static size_t sort_temp_bytes(i64 n) {
    size_t bytes = 0, scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const u32*)nullptr, (u32*)nullptr, (const u32*)nullptr, (u32*)nullptr,
                                    (i64)n, 0, 32, (cudaStream_t)0);
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (const u32*)nullptr, (u32*)nullptr, (i64)((n + LEAF - 1) / LEAF), (cudaStream_t)0);
    return bytes > scan ? bytes : scan;       // the same buffer serves the sorts and the node-rank scans
}

extern "C" size_t al_index_bytes(int dtype, int64_t n, int d) {
    IndexHeader h;
    index_layout(dtype, n, d, &h);
    return h.total_bytes;
}

extern "C" int64_t al_index_positions(int64_t n) { return (n + LEAF - 1) / LEAF * LEAF; }

extern "C" const uint32_t* al_index_perm(const void* index) { return (const uint32_t*)((const char*)index + 256); }

extern "C" size_t al_index_build_workspace_bytes(int dtype, int64_t n, int d) {
    i64 n_leaves = (n + LEAF - 1) / LEAF;
    ArenaSizer s;
    s.add<u32>(n);
    s.add<u32>(n);
    s.add<u32>(n);
    s.add<u32>(n_leaves);
    s.add<u32>(n_leaves);
    s.add<u32>(n_leaves * d);
    s.add<u32>(n_leaves * d);
    s.add<int>(n_leaves);
    s.add<float>(n_leaves);
    s.add<float>(n_leaves);
    s.add<u32>(n_leaves);
    s.add<u32>(n_leaves);
    s.add<u32>(n);
    s.add<u32>(n);
    s.add<char>((size_t)n * d * (dtype == AL_F64 ? 8 : 4));
    s.add<char>((size_t)n * d * (dtype == AL_F64 ? 8 : 4));
    s.add<char>(sort_temp_bytes(n));
    return s.off + 256;
}

template <typename T>
static int index_build_impl(const T* P, const IndexHeader& h, void* index, void* ws, size_t ws_bytes, cudaStream_t st) {
    const i64 n = h.n, n_leaves = h.n_leaves;
    const int d = h.d;
    Arena ar(ws, ws_bytes);
    BuildWs w;
    w.keys0 = ar.get<u32>(n);
    w.keys1 = ar.get<u32>(n);
    w.vals1 = ar.get<u32>(n);
    w.nodeA = ar.get<u32>(n_leaves);
    w.nodeB = ar.get<u32>(n_leaves);
    w.bbox_lo = ar.get<u32>(n_leaves * d);
    w.bbox_hi = ar.get<u32>(n_leaves * d);
    w.sdim = ar.get<int>(n_leaves);
    w.slo = ar.get<float>(n_leaves);
    w.sscale = ar.get<float>(n_leaves);
    w.nflag = ar.get<u32>(n_leaves);
    w.nrank = ar.get<u32>(n_leaves);
    w.iota = ar.get<u32>(n);
    w.slots = ar.get<u32>(n);
    w.rows0 = ar.get<char>((size_t)n * d * sizeof(T));
    w.rows1 = ar.get<char>((size_t)n * d * sizeof(T));
    w.cub_bytes = sort_temp_bytes(n);
    w.cub_tmp = ar.get<char>(w.cub_bytes);
    AL_REQUIRE(ar.ok, AL_ENOMEM, "al_index_build: workspace too small (%zu bytes given)", ws_bytes);

    char* base = (char*)index;
    u32* perm = (u32*)(base + h.perm_off);
    u32* ids[2] = {perm, w.vals1};       // point ids by position (ping-pong)
    T* rows[2] = {(T*)w.rows0, (T*)w.rows1};
    const T* Pcur = P;                   // point rows by position: the input itself while the order is the identity
    int cur = 0, rc = 0;

    tree_init_kernel<<<al_grid(n, 256), 256, 0, st>>>(perm, w.nodeA, w.nodeB, n, n_leaves);
    AL_CHECK_LAUNCH();
    tree_init_kernel<<<al_grid(n, 256), 256, 0, st>>>(w.iota, w.nodeA, w.nodeB, n, n_leaves);
    AL_CHECK_LAUNCH();

    const int depth = tree_depth(n_leaves);
    // Split levels served by one sort.  After a sort a node is ordered along its widest axis, so the same
    // order can serve several cuts (quantile slabs).  Measured on B200: in <= 3-D three levels per sort are
    // best throughout; in higher dimensions slabs cost search time near the leaves, so the upper levels take
    // two per sort and the last 7 levels (128-leaf subtrees) one each.  AL_TREE_LEVELS / AL_TREE_TOP override.
    // The last 7 levels (subtrees of 128 leaves) are finished in shared memory by tree_bottom_kernel when a
    // block's coordinates fit (d <= 9); the global rounds then only go down to 128-leaf nodes.
    const bool use_bottom = tree_bottom_smem(d) <= 220 * 1024 && getenv("AL_TREE_NO_BOTTOM") == nullptr;
    int top_levels = d <= 3 ? 3 : 2;
    int top_rounds = d <= 3 ? 64 : (depth > 7 ? (depth - 7) / 2 : 0);
    const int global_depth = use_bottom ? (depth > 7 ? depth - 7 : 0) : depth;
    if (const char* e = getenv("AL_TREE_LEVELS")) { int v = atoi(e); if (v >= 1 && v <= 8) top_levels = v; }
    if (const char* e = getenv("AL_TREE_TOP")) { int v = atoi(e); if (v >= 0 && v <= 64) top_rounds = v; }
    int sched[64];
    int rounds = 0;
    for (int done = 0; done < global_depth && rounds < 64; rounds++) {
        sched[rounds] = rounds < top_rounds ? top_levels : 1;
        done += sched[rounds];
    }
    int node_bits = 1;
    while (((i64)1 << node_bits) < n_leaves) node_bits++;
    // quantised-coordinate bits of the 32-bit sort key: what the node bits of the round leave, up to a cap
    int cbits_cap = 14;      // measured on C4 (build + search): 13: 70.44, 14: 70.25, 16: 70.34, 18: 70.48 ms
    if (const char* e = getenv("AL_TREE_CBITS")) { int v = atoi(e); if (v >= 3 && v <= 24) cbits_cap = v; }   // development knob
    AL_REQUIRE(32 - node_bits >= 5, AL_EINVAL, "al_index_build: too many leaves for 32-bit sort keys");
    int levels_done = 0;
    const i64 leaves_per_block = (256 / 32) * BOUNDS_CHUNK;
    const unsigned bounds_grid = (unsigned)((n_leaves + leaves_per_block - 1) / leaves_per_block);
    for (int r = 0; r < rounds; r++) {
        const int rank_bits = levels_done < node_bits ? levels_done : node_bits;      // at most 2^levels_done nodes exist
        const int cbits = 32 - rank_bits < cbits_cap ? 32 - rank_bits : cbits_cap;
        if (r == 0) {
            // (later rounds: the boxes were accumulated by the previous round's permutation pass)
            AL_CUDA(cudaMemsetAsync(w.bbox_lo, 0xff, sizeof(u32) * n_leaves * d, st));
            AL_CUDA(cudaMemsetAsync(w.bbox_hi, 0x00, sizeof(u32) * n_leaves * d, st));
            tree_node_bounds_kernel<T><<<bounds_grid, 256, 0, st>>>(P, ids[cur], w.nodeA, w.nodeB, w.bbox_lo, w.bbox_hi, n, n_leaves, d);
            AL_CHECK_LAUNCH();
        }
        tree_node_split_info_kernel<<<(unsigned)((n_leaves + 255) / 256), 256, 0, st>>>(w.nodeA, w.nodeB, w.bbox_lo, w.bbox_hi, w.sdim,
                                                                                       w.slo, w.sscale, n_leaves, d, cbits);
        AL_CHECK_LAUNCH();
        // node ranks: an exclusive sum over the node-start flags of the leaves
        tree_node_start_kernel<<<(unsigned)((n_leaves + 255) / 256), 256, 0, st>>>(w.nodeA, w.nflag, n_leaves);
        AL_CHECK_LAUNCH();
        size_t sb = w.cub_bytes;
        AL_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, sb, w.nflag, w.nrank, n_leaves, st));
        tree_make_keys_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
            Pcur, nullptr, w.nodeA, w.nodeB, w.sdim, w.slo, w.sscale, w.nrank, w.keys0, n, d, cbits);
        AL_CHECK_LAUNCH();
        size_t tb = w.cub_bytes;
        // the sort carries positions (slots), not ids: the rows and ids follow in the permutation pass below
        AL_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys0, w.keys1, w.iota, w.slots, n, 0, cbits + rank_bits, st));
        levels_done += sched[r];
        tree_split_kernel<<<(unsigned)((n_leaves + 255) / 256), 256, 0, st>>>(w.nodeA, w.nodeB, n_leaves, sched[r]);
        AL_CHECK_LAUNCH();
        const int want_bounds = r + 1 < rounds ? 1 : 0;
        if (want_bounds) {
            AL_CUDA(cudaMemsetAsync(w.bbox_lo, 0xff, sizeof(u32) * n_leaves * d, st));
            AL_CUDA(cudaMemsetAsync(w.bbox_hi, 0x00, sizeof(u32) * n_leaves * d, st));
        }
        // rows of a whole number of 8-byte units (and an 8-byte aligned input) take the vector path
        const bool v8 = ((size_t)d * sizeof(T)) % 8 == 0 && ((uintptr_t)Pcur & 7u) == 0;
#define AL_PERMUTE(V, M, F)                                                                                                               \
    tree_permute_bounds_kernel<T, V, M, F><<<bounds_grid, 256, 0, st>>>(Pcur, ids[cur], w.slots, rows[rc], ids[cur ^ 1], w.nodeA, w.nodeB, \
                                                                        w.bbox_lo, w.bbox_hi, n, n_leaves, d, want_bounds)
        // (two leaves in flight; one or four measured the same: 5.92 / 5.88 / 5.89 ms for the whole build at C4)
        bool launched = false;
        if (v8 && d <= 8) { AL_PERMUTE(true, 8, 2); launched = true; }
        if constexpr (sizeof(T) == 4) {      // (float64 rows of more than 8 coordinates would not fit the staging buffer)
            if (!launched && v8) { AL_PERMUTE(true, MAXD, 2); launched = true; }
        }
        if (!launched) {
            if (d <= 8) AL_PERMUTE(false, 8, 1);
            else AL_PERMUTE(false, MAXD, 1);
        }
#undef AL_PERMUTE
        AL_CHECK_LAUNCH();
        Pcur = rows[rc];
        rc ^= 1;
        cur ^= 1;
    }
    if (ids[cur] != perm) AL_CUDA(cudaMemcpyAsync(perm, ids[cur], sizeof(u32) * n, cudaMemcpyDeviceToDevice, st));
    if (use_bottom && depth > 0) {
        const size_t smem = tree_bottom_smem(d);
        AL_CUDA(cudaFuncSetAttribute(tree_bottom_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int cb = 10;      // coordinate bits of the in-block sort key: children overlap by at most 1/1024 of the node extent
        if (const char* e = getenv("AL_TREE_BOT_CBITS")) { int v = atoi(e); if (v >= 5 && v <= 16) cb = v; }   // development knob
        tree_bottom_kernel<T><<<(unsigned)((n_leaves + BOT_LEAVES - 1) / BOT_LEAVES), BOT_THREADS, smem, st>>>(Pcur, rows[rc], perm, n, n_leaves, d, cb);
        AL_CHECK_LAUNCH();
        Pcur = rows[rc];
    }

    T* box0 = (T*)(base + h.box_off[0]);
    tree_build_tiles_kernel<T><<<(unsigned)((n_leaves * 32 + 255) / 256), 256, 0, st>>>(Pcur, perm, base + h.tiles_off, box0, h);
    AL_CHECK_LAUNCH();
    for (int l = 1; l <= h.n_levels; l++) {
        const T* cbox = (const T*)(base + h.box_off[l - 1]);
        T* pbox = (T*)(base + h.box_off[l]);
        const i64 work = h.count[l] * h.box_stride;
        tree_build_level_kernel<T><<<(unsigned)((work + 127) / 128), 128, 0, st>>>(cbox, h.count[l - 1], pbox, h.count[l], d, h.box_stride);
        AL_CHECK_LAUNCH();
    }
    AL_CUDA(cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    AL_CUDA(cudaStreamSynchronize(st));
    index_register(index, h);
    return AL_OK;
}

extern "C" int al_index_build(const void* Pt, int dtype, int64_t n, int d, void* index, size_t index_bytes, void* ws,
                              size_t ws_bytes, void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_index_build: dtype must be AL_F32 or AL_F64");
    AL_REQUIRE(n >= 1 && n < 0xffffffffll, AL_EINVAL, "al_index_build: n out of range");
    AL_REQUIRE(d >= 1 && d <= MAXD, AL_EINVAL, "al_index_build: d must be in [1, %d]", MAXD);
    static_assert(sizeof(IndexHeader) <= 256, "IndexHeader must fit its slot");
    IndexHeader h;
    index_layout(dtype, n, d, &h);
    AL_REQUIRE(index_bytes >= h.total_bytes, AL_ENOMEM, "al_index_build: index buffer too small");
    AL_REQUIRE(((uintptr_t)index & 255) == 0, AL_EINVAL, "al_index_build: index must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == AL_F32) return index_build_impl<float>((const float*)Pt, h, index, ws, ws_bytes, st);
    return index_build_impl<double>((const double*)Pt, h, index, ws, ws_bytes, st);
}
