// Exact k-nearest-neighbour search kernel: replaces pykdtree's
// KDTree.query(P_transform, k=k_den, sqr_dists=True)
// (reference astrolink/astrolink.py:279).
//
// One warp owns one leaf of 32 queries (one query per lane).  It walks the
// 32-ary box hierarchy of the index bottom-up from its own leaf.  Opening a wide
// node, lane j takes child j: a box-to-box bound against the warp's bound, then,
// for every query that still needs the parent, the distance from that query to
// the child's box against the query's own k-th distance -- a 32-bit need mask
// per child.  Needed leaves are pulled into shared memory with TMA bulk copies
// (cp.async.bulk + mbarrier, double buffered: while tile t is evaluated the
// traversal has already found tile t+1 and its copy is in flight).  A tile that
// many queries need is evaluated densely (every lane all 32 candidates, packed
// f32x2 SUB/FMA, two candidates per instruction); otherwise lane l holds
// candidate l and the needy queries are evaluated two per pass.  Candidates
// that beat a lane's current k-th distance go to a per-lane queue in shared
// memory; when a queue is close to full the warp merges the queues into
// per-lane max-heaps in shared memory ([slot][lane], conflict-free), heap-sorted
// in place at the end.  k is a runtime value (<= 128).
//
// Exactness: the distance is acc = fma(q_j - p_j, q_j - p_j, acc), j ascending,
// in the data dtype -- the same chain as the oracle -- and lists are ordered by
// (distance, original index).  Box lower bounds use the same chain on per-axis
// gaps; rounding is monotone, so a bound never exceeds the distance of any
// point in the box and pruning (strict >) never drops a neighbour or a tie.
#pragma once
#include "index.cuh"

#ifndef KNN_SPARSE_MAX
#define KNN_SPARSE_MAX 12
#endif
#ifndef KNN_ORDER
#define KNN_ORDER 1            // 1: visit the leaves of a wide node nearest box first
#endif
#ifndef KNN_BRUTE
#define KNN_BRUTE 0            // 1 (microbenchmark only): no pruning, every tile evaluated densely by every warp
#endif
#ifndef KNN_STATS
#define KNN_STATS 0            // 1: maintain the diagnostic counters [3..7] of the stats array
#endif
#ifndef KNN_QCAP
#define KNN_QCAP 8
#endif
// Branch-trimming switches, each measured on C4 (tools/knn_tournament.py): 92.9 ms with all three off,
// 92.0 / 91.6 / 90.1 ms with one on, 88.7 ms with all three.  0 selects the earlier code for A/B runs.
#ifndef KNN_FLUSH_ONE_PREDICATE
#define KNN_FLUSH_ONE_PREDICATE 1               // one predicate per flush step
#endif
#ifndef KNN_SIFT_CLAMPED_CHILD
#define KNN_SIFT_CLAMPED_CHILD 1               // sift-down with a clamped right child (no branch on its existence)
#endif
#ifndef KNN_SPARSE_ONE_VOTE
#define KNN_SPARSE_ONE_VOTE 1               // one combined vote per sparse pass, the per-query ballots only when something hit
#endif
#ifndef KNN_MINBLOCKS
#define KNN_MINBLOCKS 5
#endif
#ifndef KNN_WARPS
#define KNN_WARPS 4
#endif
constexpr int WARPS = KNN_WARPS;     // warps (query leaves) per CTA
constexpr int QCAP = KNN_QCAP;       // queue entries per lane
#ifndef KNN_CHUNK_PB
#define KNN_CHUNK_PB 2
#endif
constexpr int CHUNK_PB = KNN_CHUNK_PB;   // pair blocks (2 candidates each) between queue checks
constexpr unsigned FULL = 0xffffffffu;
constexpr int KNN_NSTATS = 8;        // counters written to the stats array (see al_knn)
constexpr int SPARSE_MAX = KNN_SPARSE_MAX;       // tiles needed by at most this many queries are evaluated one query at a time
static_assert(QCAP >= 4 * CHUNK_PB, "queue must absorb two chunks");

struct KnnParams {
    const char* tiles;
    const void* box[MAX_WIDE_LEVELS];
    i64 count[MAX_WIDE_LEVELS];
    int n_levels;
    int tile_stride, tile_hdr_bytes, tile_ids_off, pair_block, box_stride;
    i64 leaf_begin, leaf_end;
    // interleaved sharding (multi-GPU): CTAs are dealt to this launch in chunks of `chunk_blocks` consecutive
    // CTAs, consecutive chunks of one launch lying `chunk_stride_blocks` CTAs apart in the leaf order
    // (0: one contiguous range)
    i64 chunk_blocks, chunk_stride_blocks;
    i64 n;
    int k;
    int row_order;
    u32* idx_out;
    void* d2_out;
    u32* link_out;      // optional [n][link_cols] at the original point index; may live on a peer GPU
    int link_cols;
    int link_by_position;   // 0: link_out row = original point index; 1: row = index position (warp-coalesced block stores)
    void* logrho_out;   // optional: raw log-density of every query (unweighted), fused into the epilogue; row as link_out's
    double half_d;      // d_intrinsic / 2 (exponent of the core distance in the density)
    unsigned long long* stats;   // [KNN_NSTATS] or nullptr
    int warp_smem;
    int use_generic;    // float32 only: 1 = search with this file's kernel instead of knn_fast.cuh's
};

// per-warp shared memory: 2 tiles | queue | heap | traversal stack | 2 mbarriers
// query-table row: D coordinates + the query's current k-th distance, padded to 16 bytes
template <typename T>
static inline int knn_qrow_elems(int d) {
    int per16 = 16 / (int)sizeof(T);
    return (d + 1 + per16 - 1) / per16 * per16;
}
template <typename T>
static inline int knn_warp_smem_bytes(int tile_stride, int k, int d, int n_levels) {
    size_t b = 2 * (size_t)tile_stride;
    b += (size_t)QCAP * 32 * (sizeof(T) + 4);
    b += (size_t)k * 32 * (sizeof(T) + 4);
    b += (size_t)32 * knn_qrow_elems<T>(d) * sizeof(T);
    b += (size_t)(n_levels + 1) * 32 * 4;
    b += (size_t)MAX_WIDE_LEVELS * (8 + 4 + 4);
    b += 16;
    return (int)al_align(b, 128);
}

#ifdef __CUDACC__
// ---- mbarrier / TMA bulk copy ---------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u64* bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- packed f32x2 arithmetic ------------------------------------------------
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return __fma_rn(a, b, c); }

template <typename T> __device__ __forceinline__ T inf_t() { return (T)INFINITY; }
// per-axis gap between an interval [lo, hi] and a point/interval: max(lo - x_hi, x_lo - hi, 0)
__device__ __forceinline__ float gap_t(float a, float b) { return fmaxf(fmaxf(a, b), 0.f); }
__device__ __forceinline__ double gap_t(double a, double b) { return fmax(fmax(a, b), 0.0); }

template <typename T> __device__ __forceinline__ T warp_max_t(T v);
template <> __device__ __forceinline__ float warp_max_t<float>(float v) {
    // values are >= 0, +inf, or -1 (padding lane): int order == float order
    return __int_as_float(__reduce_max_sync(FULL, __float_as_int(v)));
}
template <> __device__ __forceinline__ double warp_max_t<double>(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(FULL, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// (d2, id) lexicographic: is a strictly worse (later) than b?
template <typename T>
__device__ __forceinline__ bool worse(T da, u32 ia, T db, u32 ib) {
    return da > db || (da == db && ia > ib);
}

// explicit shared-space loads (32-bit shared addresses): tile reads must be LDS, not generic LD
__device__ __forceinline__ float4 lds_f4(u32 addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_d2(u32 addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

// one row of the query table (RS elements, 16-byte aligned) with 128-bit loads
template <int RS>
__device__ __forceinline__ void load_row(const float* row, float* out) {
#pragma unroll
    for (int v = 0; v < RS / 4; v++) {
        const float4 f = reinterpret_cast<const float4*>(row)[v];
        out[4 * v] = f.x; out[4 * v + 1] = f.y; out[4 * v + 2] = f.z; out[4 * v + 3] = f.w;
    }
}
template <int RS>
__device__ __forceinline__ void load_row(const double* row, double* out) {
#pragma unroll
    for (int v = 0; v < RS / 2; v++) {
        const double2 f = reinterpret_cast<const double2*>(row)[v];
        out[2 * v] = f.x; out[2 * v + 1] = f.y;
    }
}

// ---------------------------------------------------------------------------
// (distance, id) keys.  float: one 64-bit word, distance bits in the high half
// (distances are >= 0, so unsigned order == (distance, id) lexicographic order);
// double: a (double, u32) pair held in two arrays.
template <typename T> struct KeyStore;
template <> struct KeyStore<float> {
    typedef u64 Key;
    u64* k;
    static constexpr int BYTES = 8;
    __device__ __forceinline__ static Key make(float d, u32 id) { return ((u64)__float_as_uint(d) << 32) | id; }
    __device__ __forceinline__ static bool worse(Key a, Key b) { return a > b; }
    __device__ __forceinline__ static float dist(Key a) { return __uint_as_float((u32)(a >> 32)); }
    __device__ __forceinline__ static u32 id(Key a) { return (u32)a; }
    __device__ __forceinline__ Key ld(int i) const { return k[i]; }
    __device__ __forceinline__ void st(int i, Key v) const { k[i] = v; }
    __device__ __forceinline__ static KeyStore at(unsigned char* base, int entries) { (void)entries; KeyStore s; s.k = (u64*)base; return s; }
};
struct KeyD { double d; u32 id; };
template <> struct KeyStore<double> {
    typedef KeyD Key;
    double* dd;
    u32* ii;
    static constexpr int BYTES = 12;
    __device__ __forceinline__ static Key make(double d, u32 id) { KeyD k; k.d = d; k.id = id; return k; }
    __device__ __forceinline__ static bool worse(Key a, Key b) { return a.d > b.d || (a.d == b.d && a.id > b.id); }
    __device__ __forceinline__ static double dist(Key a) { return a.d; }
    __device__ __forceinline__ static u32 id(Key a) { return a.id; }
    __device__ __forceinline__ Key ld(int i) const { KeyD k; k.d = dd[i]; k.id = ii[i]; return k; }
    __device__ __forceinline__ void st(int i, Key v) const { dd[i] = v.d; ii[i] = v.id; }
    __device__ __forceinline__ static KeyStore at(unsigned char* base, int entries) {
        KeyStore s; s.dd = (double*)base; s.ii = (u32*)(base + (size_t)entries * 8); return s;
    }
};

// Per-lane top-k: a max-heap of keys in shared memory, laid out [slot][lane].
// Root = current k-th (worst).  Replaces the root by `v` and restores the heap.
template <typename T>
__device__ __forceinline__ void heap_replace_root(const KeyStore<T>& h, int K, int lane, typename KeyStore<T>::Key v) {
    typedef KeyStore<T> KS;
    int pos = 0;
    for (;;) {
        const int l = 2 * pos + 1;
        if (l >= K) break;
        int c = l;
        typename KS::Key ck = h.ld(l * 32 + lane);
#if KNN_SIFT_CLAMPED_CHILD
        {
            const int r = l + 1 < K ? l + 1 : l;        // a missing right child reads the left one again: never "worse"
            const typename KS::Key rk = h.ld(r * 32 + lane);
            if (KS::worse(rk, ck)) { c = r; ck = rk; }
        }
#else
        if (l + 1 < K) {
            const typename KS::Key rk = h.ld((l + 1) * 32 + lane);
            if (KS::worse(rk, ck)) { c = l + 1; ck = rk; }
        }
#endif
        if (!KS::worse(ck, v)) break;
        h.st(pos * 32 + lane, ck);
        pos = c;
    }
    h.st(pos * 32 + lane, v);
}

template <typename T, int D>
struct WarpState {
    T q[D];
    T worst;        // distance of the heap root for real queries, -1 for padding lanes
    T bound;        // warp max of worst
    typename KeyStore<T>::Key root;   // cached heap root
    int cnt;        // queue fill
    bool active;
    u32 n_push, n_insert;   // statistics
};

template <typename T, int D>
__device__ __forceinline__ void flush_queues(WarpState<T, D>& s, const KeyStore<T>& qk, const KeyStore<T>& hk, T* qworst, int K,
                                             int lane, u32& n_flush, u32& n_iter) {
    typedef KeyStore<T> KS;
    const int maxc = __reduce_max_sync(FULL, s.cnt);
    if (KNN_STATS) { n_flush++; n_iter += maxc; }
#pragma unroll 1
    for (int t = 0; t < maxc; t++) {
#if KNN_FLUSH_ONE_PREDICATE
        const typename KS::Key v = qk.ld(t * 32 + lane);          // slots past cnt hold stale keys: masked by the test below
        if (t < s.cnt && KS::worse(s.root, v)) {
            heap_replace_root<T>(hk, K, lane, v);
            s.root = hk.ld(lane);
            if (KNN_STATS) s.n_insert++;
        }
#else
        if (t < s.cnt) {
            const typename KS::Key v = qk.ld(t * 32 + lane);
            if (KS::worse(s.root, v)) {
                heap_replace_root<T>(hk, K, lane, v);
                s.root = hk.ld(lane);
                if (KNN_STATS) s.n_insert++;
            }
        }
#endif
    }
    s.cnt = 0;
    if (s.active) s.worst = KS::dist(s.root);
    qworst[0] = s.worst;            // this lane's row of the query table
    __syncwarp();
    s.bound = warp_max_t<T>(s.worst);
}

// distances of this lane's query to the 2*CHUNK_PB candidates of one chunk; `coords` is a shared address
template <int D>
__device__ __forceinline__ void chunk_dist(const float* q, u32 coords, int pair_block, int pb0, float* out) {
    constexpr int PB4 = (2 * D + 3) / 4;   // float4 per pair block
    (void)pair_block;
    u64 qq[D];
#pragma unroll
    for (int j = 0; j < D; j++) qq[j] = pack2(q[j], q[j]);
#pragma unroll
    for (int c = 0; c < CHUNK_PB; c++) {
        const int pb = pb0 + c;
        u64 acc = 0ull;   // (+0, +0)
#pragma unroll
        for (int v4 = 0; v4 < PB4; v4++) {
            const float4 f = lds_f4(coords + (u32)(pb * PB4 + v4) * 16u);
            if (2 * v4 < D) {
                u64 df = sub2(qq[2 * v4], pack2(f.x, f.y));
                acc = fma2(df, df, acc);
            }
            if (2 * v4 + 1 < D) {
                u64 df = sub2(qq[2 * v4 + 1], pack2(f.z, f.w));
                acc = fma2(df, df, acc);
            }
        }
        unpack2(acc, out[2 * c], out[2 * c + 1]);
    }
}
template <int D>
__device__ __forceinline__ void chunk_dist(const double* q, u32 coords, int pair_block, int pb0, double* out) {
#pragma unroll
    for (int c = 0; c < CHUNK_PB; c++) {
        const int pb = pb0 + c;
        double va = 0.0, vb = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) {
            const double2 cc = lds_d2(coords + (u32)(pb * pair_block + 2 * j) * 8u);
            const double da = q[j] - cc.x, db = q[j] - cc.y;
            va = __fma_rn(da, da, va);
            vb = __fma_rn(db, db, vb);
        }
        out[2 * c] = va;
        out[2 * c + 1] = vb;
    }
}

// distance from query rows ra / rb (query table) to this lane's candidate cc: two queries per pass
template <int D>
__device__ __forceinline__ void pair_dist(const float* ra, const float* rb, const float* cc, float& va, float& vb) {
    // two independent scalar chains: the FP32 pipe is not the limit here, issue slots are, and packing
    // two query rows into f32x2 operands costs more register moves than the packed math saves
    va = 0.f;
    vb = 0.f;
#pragma unroll
    for (int j = 0; j < D; j++) {
        const float da = ra[j] - cc[j], db = rb[j] - cc[j];
        va = __fmaf_rn(da, da, va);
        vb = __fmaf_rn(db, db, vb);
    }
}
template <int D>
__device__ __forceinline__ void pair_dist(const double* ra, const double* rb, const double* cc, double& va, double& vb) {
    va = 0.0;
    vb = 0.0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        const double da = ra[j] - cc[j], db = rb[j] - cc[j];
        va = __fma_rn(da, da, va);
        vb = __fma_rn(db, db, vb);
    }
}

template <typename T, int D>
__global__ void __launch_bounds__(WARPS * 32, KNN_MINBLOCKS) knn_kernel(const KnnParams p) {
    typedef KeyStore<T> KS;
    typedef typename KS::Key Key;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // leaf and node indices fit 32 bits (n < 2^31, 32-point leaves): the traversal uses int arithmetic
    i64 gblock = blockIdx.x;
    if (p.chunk_blocks > 0) gblock = (gblock / p.chunk_blocks) * p.chunk_stride_blocks + gblock % p.chunk_blocks;
    const i64 ql64 = p.leaf_begin + gblock * WARPS + wib;
    if (ql64 >= p.leaf_end) return;
    const int ql = (int)ql64;
    const int K = p.k;

    // ---- per-warp shared memory carve-up (integer offsets from the shared base)
    constexpr int PER16 = 16 / (int)sizeof(T);
    constexpr int RS = (D + 1 + PER16 - 1) / PER16 * PER16;   // query-table row stride (elements)
    const u32 wo = (u32)wib * (u32)p.warp_smem;
    unsigned char* wbase = smem + wo;
    unsigned char* o = wbase + 2u * (u32)p.tile_stride;
    const KS qk = KS::at(o, QCAP * 32);                     // [QCAP][32] queue of keys
    o += (size_t)QCAP * 32 * KS::BYTES;
    const KS hk = KS::at(o, K * 32);                        // [K][32] heap of keys
    o += (size_t)K * 32 * KS::BYTES;
    T* qtab = (T*)o;                                        // [32][RS]: query coordinates, k-th distance
    u32* sneed = (u32*)(qtab + 32 * RS);                    // [n_levels + 1][32] per-child query masks
    int* snode = (int*)(sneed + (p.n_levels + 1) * 32);     // [MAX_WIDE_LEVELS] (8 bytes reserved per entry)
    u32* smask = (u32*)(snode + 2 * MAX_WIDE_LEVELS);       // [MAX_WIDE_LEVELS]
    int* slevel = (int*)(smask + MAX_WIDE_LEVELS);          // [MAX_WIDE_LEVELS]
    u64* bars = (u64*)(slevel + MAX_WIDE_LEVELS);           // [2]
    const u32 tile_sa = smem_u32(wbase);                    // shared address of tile buffer 0
    T* qworst = qtab + lane * RS + D;

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const Key kinf = KS::make(inf_t<T>(), 0xffffffffu);
#pragma unroll 1
    for (int m = 0; m < K; m++) hk.st(m * 32 + lane, kinf);
#pragma unroll
    for (int j = 0; j < RS; j++) qtab[lane * RS + j] = j == D ? (T)-1 : (T)0;
    __syncwarp();

    WarpState<T, D> s;
    s.cnt = 0;
    s.active = false;
    s.worst = (T)-1;
    s.bound = inf_t<T>();
    s.root = kinf;
    s.n_push = 0;
    s.n_insert = 0;
    u32 n_flush = 0, n_iter = 0, n_loaded = 0, n_eval = 0, n_needy = 0, n_slow = 0;
    u32 n_qt = 0;       // (query, tile) evaluations: 32 pairs each
    // the query box: tight bounds of the own leaf (level-0 box arrays)
    T qlo[D], qhi[D];
    {
        const T* b0 = (const T*)p.box[0] + (size_t)ql * p.box_stride;
#pragma unroll
        for (int j = 0; j < D; j++) {
            qlo[j] = b0[2 * j];
            qhi[j] = b0[2 * j + 1];
            s.q[j] = (T)0;
        }
    }
    u32 my_id = 0xffffffffu;

    // Open a wide node: lane j takes child j (level `hc`), tests box-to-box against the warp bound, then
    // for every query in `qmask` the distance from the query to the child's box against the query's own
    // k-th distance.  Returns the children some query needs; per-child query masks go to need_out[lane].
    u32 top_dist = 0xffffffffu;     // this lane's child-box distance (float bits) of the most recent open
    auto open_node = [&](int hc, int node, int excl, u32 qmask, u32* need_out) -> u32 {
        const int cnt = (int)p.count[hc];
        const int c = node * FANOUT + lane;
        const bool ok = c < cnt && c != excl;
        const T* brow = (const T*)p.box[hc] + (size_t)(ok ? c : 0) * p.box_stride;
        T blo[D], bhi[D];
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < D; j++) {
            blo[j] = ok ? brow[2 * j] : (T)0;
            bhi[j] = ok ? brow[2 * j + 1] : (T)0;
            T a = blo[j] - qhi[j], b = qlo[j] - bhi[j];
            const T g = gap_t(a, b);
            acc = fma_t<T>(g, g, acc);
        }
        const bool pass = ok && (KNN_BRUTE || acc <= s.bound);
        u32 mask = 0;
        if (KNN_BRUTE) {
            mask = pass ? qmask : 0u;
        } else if (__any_sync(FULL, pass)) {
            u32 m = qmask;
#pragma unroll 1
            while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                T row[RS];
                load_row<RS>(qtab + q * RS, row);
                T pd = (T)0;
#pragma unroll
                for (int j = 0; j < D; j++) {
                    T a = blo[j] - row[j], b = row[j] - bhi[j];
                    const T g = gap_t(a, b);
                    pd = fma_t<T>(g, g, pd);
                }
                if (pass && pd <= row[D]) mask |= 1u << q;
            }
        }
        need_out[lane] = mask;
        top_dist = mask != 0 ? __float_as_uint((float)acc) : 0xffffffffu;   // order only; float precision is enough
        return __ballot_sync(FULL, mask != 0);
    };

    // ---- traversal state: ancestors of the own leaf, bottom-up; a small stack for descents.
    // The top entry (children mask, node, level) lives in registers; lower entries in shared memory.
    int anc = 1;    // next ancestor level of the own leaf to open
    int sp = 0;     // entries on the stack, including the register-resident top
    u32 t_mask = 0;
    int t_node = 0;
    int t_h = 0;
    u32 cand_need = FULL;
    auto next_candidate = [&]() -> int {
        for (;;) {
            if (sp == 0) {
                if (anc > p.n_levels) return -1;
                // open the children of the level-`anc` ancestor, except the branch already done
                if (__any_sync(FULL, s.cnt > 0)) flush_queues<T, D>(s, qk, hk, qworst, K, lane, n_flush, n_iter);
                const int node = ql >> (5 * anc), excl = ql >> (5 * (anc - 1));
                t_mask = open_node(anc - 1, node, excl, FULL, sneed);
                t_node = node;
                t_h = anc;
                __syncwarp();
                sp = 1;
                anc++;
                continue;
            }
            if (t_mask == 0) {
                sp--;
                if (sp > 0) { t_mask = smask[sp - 1]; t_node = snode[sp - 1]; t_h = slevel[sp - 1]; }
                continue;
            }
            int bit;
            if (KNN_ORDER && t_h == 1) {
                // leaves: nearest box first, so the k-th distances tighten as early as possible
                // the lane number rides in the low 5 bits of the key (the order is a heuristic: 27 bits of distance are plenty),
                // so one warp reduction yields both the minimum and who holds it
                const u32 key = ((t_mask >> lane) & 1u) ? ((top_dist & ~31u) | (u32)lane) : 0xffffffffu;
                bit = (int)(__reduce_min_sync(FULL, key) & 31u);
            } else {
                bit = __ffs(t_mask) - 1;
            }
            t_mask &= ~(1u << bit);
            const u32 need = sneed[(sp - 1) * 32 + bit];
            const int c = t_node * FANOUT + bit;
            if (t_h == 1) { cand_need = need; return c; }
            // descend: save the top entry, open the children of node c (a level t_h-1 node) for the queries that need it
            if (lane == 0) { smask[sp - 1] = t_mask; snode[sp - 1] = t_node; slevel[sp - 1] = t_h; }
            t_mask = open_node(t_h - 2, c, -1, need, sneed + sp * 32);
            t_node = c;
            t_h = t_h - 1;
            __syncwarp();
            sp++;
        }
    };

    // hits of one needy query (sparse path): owner lane `q` queues every candidate flagged in `hm`
    auto emit_hits = [&](u32 hm, T v, u32 cid, int q) {
#pragma unroll 1
        while (hm) {
            const int b = __ffs(hm) - 1;
            hm &= hm - 1;
            const T vv = __shfl_sync(FULL, v, b);
            const u32 ii = __shfl_sync(FULL, cid, b);
            if (__shfl_sync(FULL, s.cnt, q) >= QCAP) flush_queues<T, D>(s, qk, hk, qworst, K, lane, n_flush, n_iter);
            if (lane == q) {
                qk.st(s.cnt * 32 + lane, KS::make(vv, ii));
                s.cnt++;
                if (KNN_STATS) s.n_push++;
            }
        }
    };

    // Tile t of this warp lands in buffer t & 1 and completes phase (t >> 1) & 1 of that buffer's barrier.
    auto issue_tile = [&](int leaf, u32 buf) {
        if (lane == 0) {
            mbar_expect_tx(&bars[buf], (u32)p.tile_stride);
            tma_load_1d(wbase + buf * (u32)p.tile_stride, p.tiles + (size_t)(u32)leaf * (u32)p.tile_stride, (u32)p.tile_stride,
                        &bars[buf]);
        }
        if (KNN_STATS) n_loaded++;
    };
    auto wait_tile = [&](u32 t) {
        while (!mbar_try_wait(&bars[t & 1u], (t >> 1) & 1u)) {}
    };

    // distances from the queries in `need` to the 32 candidates of the tile in buffer `bw`
    auto eval_tile = [&](u32 bw, u32 need) {
        const unsigned char* tile = wbase + bw * (u32)p.tile_stride;
        const T* tcoord = (const T*)(tile + p.tile_hdr_bytes);
        const u32* ids = (const u32*)(tile + p.tile_ids_off);
        const int n_need = __popc(need);
        if (KNN_BRUTE || n_need > SPARSE_MAX) {
            // ---- dense: every lane evaluates all 32 candidates (packed f32x2)
            if (KNN_STATS) n_eval++;
            n_qt += 32;
            const u32 coords = tile_sa + bw * (u32)p.tile_stride + (u32)p.tile_hdr_bytes;
            // each chunk may push 2*CHUNK_PB entries per lane; the sparse path can leave queues fuller than that allows
            if (__any_sync(FULL, s.cnt > QCAP - 2 * CHUNK_PB)) flush_queues<T, D>(s, qk, hk, qworst, K, lane, n_flush, n_iter);
#pragma unroll 1
            for (int pb0 = 0; pb0 < LEAF / 2; pb0 += CHUNK_PB) {
                T v[2 * CHUNK_PB];
                chunk_dist<D>(s.q, coords, p.pair_block, pb0, v);
                T mn = v[0];
#pragma unroll
                for (int e = 1; e < 2 * CHUNK_PB; e++) mn = v[e] < mn ? v[e] : mn;
                if (__any_sync(FULL, mn <= s.worst)) {
#pragma unroll
                    for (int e = 0; e < 2 * CHUNK_PB; e++) {
                        if (v[e] <= s.worst) {
                            qk.st(s.cnt * 32 + lane, KS::make(v[e], ids[2 * pb0 + e]));
                            s.cnt++;
                            if (KNN_STATS) s.n_push++;
                        }
                    }
                    if (__any_sync(FULL, s.cnt > QCAP - 2 * CHUNK_PB))
                        flush_queues<T, D>(s, qk, hk, qworst, K, lane, n_flush, n_iter);
                }
            }
        } else if (n_need > 0) {
            // ---- sparse: two needy queries per pass, lane l evaluates candidate l for both
            if (KNN_STATS) n_eval++;
            n_qt += n_need;
            T cc[D];
#pragma unroll
            for (int j = 0; j < D; j++) cc[j] = tcoord[(lane >> 1) * p.pair_block + 2 * j + (lane & 1)];
            const u32 cid = ids[lane];
            u32 m = need;
#pragma unroll 1
            while (m) {
                // two needy queries per pass (two independent scalar chains, see pair_dist)
                const int qa = __ffs(m) - 1;
                m &= m - 1;
                const bool two = m != 0;
                const int qb = two ? __ffs(m) - 1 : qa;
                m &= m - 1;
                T ra[RS], rb[RS];
                load_row<RS>(qtab + qa * RS, ra);
                load_row<RS>(qtab + qb * RS, rb);
                T va, vb;
                pair_dist<D>(ra, rb, cc, va, vb);
#if KNN_SPARSE_ONE_VOTE
                const bool hit_a = va <= ra[D], hit_b = two && vb <= rb[D];
                if (__any_sync(FULL, hit_a || hit_b)) {      // the common pass has no hit at all: one vote, one branch
                    const u32 ha = __ballot_sync(FULL, hit_a);
                    const u32 hb = __ballot_sync(FULL, hit_b);
                    if (ha) emit_hits(ha, va, cid, qa);
                    if (hb) emit_hits(hb, vb, cid, qb);
                }
#else
                const u32 ha = __ballot_sync(FULL, va <= ra[D]);
                const u32 hb = two ? __ballot_sync(FULL, vb <= rb[D]) : 0u;
                if (ha | hb) {      // the common pass has no hit at all: one branch
                    if (ha) emit_hits(ha, va, cid, qa);
                    if (hb) emit_hits(hb, vb, cid, qb);
                }
#endif
            }
        }
    };

    // ---- own leaf (tile 0, no look-ahead: the box tests need the queries and their first bounds)
    issue_tile(ql, 0u);
    wait_tile(0u);
    {
        const unsigned char* tile = wbase;
        const T* tcoord = (const T*)(tile + p.tile_hdr_bytes);
        const u32* ids = (const u32*)(tile + p.tile_ids_off);
        // pick up this lane's query point and publish it in the query table
#pragma unroll
        for (int j = 0; j < D; j++) {
            s.q[j] = tcoord[(lane >> 1) * p.pair_block + 2 * j + (lane & 1)];
            qtab[lane * RS + j] = s.q[j];
        }
        my_id = ids[lane];
        s.active = my_id != 0xffffffffu;
        s.worst = s.active ? inf_t<T>() : (T)-1;
        qworst[0] = s.worst;
        __syncwarp();
        eval_tile(0u, FULL);
    }
    // ---- the other leaves: while tile t is evaluated, the traversal has already found tile t + 1 and its copy is in flight
    anc = 1;
    int cur = next_candidate();
    if (cur >= 0) {
        u32 cur_need = cand_need;
        u32 t = 1;
        issue_tile(cur, 1u);
#pragma unroll 1
        for (;;) {
            const int nxt = next_candidate();
            const u32 nxt_need = cand_need;
            if (nxt >= 0) {
                __syncwarp();       // every lane is done reading the other buffer
                issue_tile(nxt, (t & 1u) ^ 1u);
            }
            wait_tile(t);
            if (KNN_STATS) { n_needy += __popc(cur_need); n_slow += __popc(cur_need); }
            eval_tile(t & 1u, cur_need);
            if (nxt < 0) break;
            cur_need = nxt_need;
            t++;
        }
    }
    if (__any_sync(FULL, s.cnt > 0)) flush_queues<T, D>(s, qk, hk, qworst, K, lane, n_flush, n_iter);

    // ---- heap -> ascending order in place (heapsort), then write the rows
#pragma unroll 1
    for (int m = K - 1; m >= 1; m--) {
        const Key t = hk.ld(m * 32 + lane);
        hk.st(m * 32 + lane, hk.ld(lane));
        heap_replace_root<T>(hk, m, lane, t);
    }
    const i64 pos_row = ((i64)blockIdx.x * WARPS + wib) * LEAF + lane;      // row of this launch (row_order 1)
    if (s.active) {
        const i64 row = p.row_order == 0 ? (i64)my_id : pos_row;
        if (p.d2_out != nullptr) {       // optional: the fused density below needs no distance rows in memory
            T* dd = (T*)p.d2_out + row * K;
#pragma unroll 1
            for (int m = 0; m < K; m++) dd[m] = KS::dist(hk.ld(m * 32 + lane));
        }
        if (p.idx_out != nullptr) {      // optional: callers that only need the link columns pass NULL
            u32* io = p.idx_out + row * K;
#pragma unroll 1
            for (int m = 0; m < K; m++) io[m] = KS::id(hk.ld(m * 32 + lane));
        }
        if (p.logrho_out != nullptr) {
            // a4 (astrolink.py:293-297) fused: the sorted distances are still in shared memory.  Same arithmetic
            // and summation order as logrho_kernel (stages.cu), so both give the same bits.
            T sum = (T)0;
#pragma unroll 1
            for (int m = 0; m < K; m++) sum += KS::dist(hk.ld(m * 32 + lane));
            const i64 orow = p.link_by_position ? (i64)ql64 * LEAF + lane : (i64)my_id;
            ((T*)p.logrho_out)[orow] = al_raw_logrho<T>(sum, KS::dist(hk.ld((K - 1) * 32 + lane)), K, p.half_d);
        }
    } else if (p.row_order == 1) {
#pragma unroll 1
        for (int m = 0; m < K; m++) {
            if (p.idx_out != nullptr) p.idx_out[pos_row * K + m] = 0xffffffffu;
            if (p.d2_out != nullptr) ((T*)p.d2_out)[pos_row * K + m] = inf_t<T>();
        }
    }
    if (p.link_out != nullptr && p.link_by_position) {
        // Position-ordered rows: the 32 rows of this leaf are one contiguous block of 32 * link_cols ids, written with
        // full 128-byte warp stores -- the shape NVLink wants when link_out is a peer-mapped buffer (4-byte stores
        // scattered by original index are one packet each and made the remote ranks packet-rate bound).  Padding
        // lanes hold 0xffffffff ids; the consumer's un-permute pass skips them.
        __syncwarp();
        u32* blk = p.link_out + (ql64 * LEAF) * p.link_cols;
        const int total = LEAF * p.link_cols;
#pragma unroll 1
        for (int e = lane; e < total; e += 32) {
            const int r = e / p.link_cols, c = e - r * p.link_cols;
            blk[e] = KS::id(hk.ld(c * 32 + r));
        }
    } else if (p.link_out != nullptr && s.active) {
        // the exchange step of the multi-GPU path, fused into the epilogue: the k_link nearest ids go
        // straight to the consumer's buffer (a peer-mapped pointer over NVLink on ranks other than the owner)
        u32* lo_ = p.link_out + (i64)my_id * p.link_cols;
#pragma unroll 1
        for (int m = 0; m < p.link_cols; m++) lo_[m] = KS::id(hk.ld(m * 32 + lane));
    }
    if (p.stats != nullptr) {
        const u32 pushes = __reduce_add_sync(FULL, s.n_push), inserts = __reduce_add_sync(FULL, s.n_insert);
        if (lane == 0) {
            atomicAdd(p.stats + 0, (unsigned long long)n_qt * LEAF);                                         // (query, candidate) pairs evaluated
            atomicAdd(p.stats + 1, (unsigned long long)n_loaded);                       // tiles fetched by TMA
            atomicAdd(p.stats + 2, (unsigned long long)n_eval);                         // tiles evaluated
            atomicAdd(p.stats + 3, (unsigned long long)n_needy);                        // lanes that needed a fetched tile
            atomicAdd(p.stats + 4, (unsigned long long)pushes);                         // queue pushes
            atomicAdd(p.stats + 5, (unsigned long long)inserts);                        // heap insertions
            atomicAdd(p.stats + 6, (unsigned long long)n_iter);                         // merge iterations
            atomicAdd(p.stats + 7, (unsigned long long)n_slow);                         // lanes needing a tile before the refresh
        }
        (void)n_flush;
    }
}

template <typename T, int D>
static int launch_knn(const KnnParams& p, cudaStream_t st) {
    i64 n_q = p.leaf_end - p.leaf_begin;
    if (p.chunk_blocks > 0) {
        // leaves of this launch: every chunk that starts before leaf_end, whole chunks (the kernel skips leaves past the end)
        const i64 stride_leaves = p.chunk_stride_blocks * WARPS, chunk_leaves = p.chunk_blocks * WARPS;
        const i64 n_chunks = n_q <= 0 ? 0 : (n_q + stride_leaves - 1) / stride_leaves;
        n_q = n_chunks * chunk_leaves;
    }
    if (n_q <= 0) return AL_OK;
    size_t smem = (size_t)p.warp_smem * WARPS;
    AL_REQUIRE(smem <= 227 * 1024, AL_EINVAL, "al_knn: k too large for shared memory");
    if (smem > 48 * 1024)
        AL_CUDA(cudaFuncSetAttribute(knn_kernel<T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<T, D><<<(unsigned)((n_q + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p);
    AL_CHECK_LAUNCH();
    return AL_OK;
}
#endif  // __CUDACC__
