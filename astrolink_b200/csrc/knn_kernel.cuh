// Exact k-nearest-neighbour search kernel: replaces pykdtree's
// KDTree.query(P_transform, k=k_den, sqr_dists=True)
// (reference astrolink/astrolink.py:279).
//
// One warp owns one leaf of 32 queries (one query per lane).  It walks the
// 32-ary box hierarchy of the index bottom-up from its own leaf (one lane tests
// one child box), and for every surviving candidate leaf pulls that leaf's
// 32-point tile into shared memory with a TMA bulk copy (cp.async.bulk +
// mbarrier, double buffered: the next tile is in flight while the current one
// is evaluated) and evaluates the 32 x 32 distances on the FP32 pipe with
// packed f32x2 SUB/FMA (two candidates per instruction).  Candidates that beat
// a lane's current k-th distance go to a per-lane queue in shared memory; when
// a queue is close to full the whole warp merges its queues into per-lane
// sorted top-k lists held in registers (k <= 32) or local memory (k <= 128).
//
// Exactness: the distance is acc = fma(q_j - p_j, q_j - p_j, acc), j ascending,
// in the data dtype -- the same chain as the oracle -- and lists are ordered by
// (distance, original index).  Box lower bounds use the same chain on per-axis
// gaps; rounding is monotone, so a bound never exceeds the distance of any
// point in the box and pruning (strict >) never drops a neighbour or a tie.
#pragma once
#include "index.cuh"

constexpr int WARPS = 4;             // warps (query leaves) per CTA
constexpr int QCAP = 16;             // queue entries per lane
constexpr int CHUNK_PB = 4;          // pair blocks (2 candidates each) between queue checks
constexpr unsigned FULL = 0xffffffffu;
constexpr int KNN_NSTATS = 8;        // counters written to the stats array (see al_knn)
constexpr int SPARSE_MAX = 12;       // tiles needed by at most this many queries are evaluated one query at a time
static_assert(QCAP >= 4 * CHUNK_PB, "queue must absorb two chunks");

struct KnnParams {
    const char* tiles;
    const void* box[MAX_WIDE_LEVELS];
    i64 count[MAX_WIDE_LEVELS];
    int n_levels;
    int tile_stride, tile_hdr_bytes, tile_ids_off, pair_block;
    i64 leaf_begin, leaf_end;
    i64 n;
    int k;
    int row_order;
    u32* idx_out;
    void* d2_out;
    unsigned long long* stats;   // [KNN_NSTATS] or nullptr
    int warp_smem;
};

// per-warp shared memory: 2 tiles | queue | heap | traversal stack | 2 mbarriers
// query-table row: D coordinates + the query's current k-th distance, padded to 16 bytes
template <typename T>
static inline int knn_qrow_elems(int d) {
    int per16 = 16 / (int)sizeof(T);
    return (d + 1 + per16 - 1) / per16 * per16;
}
template <typename T>
static inline int knn_warp_smem_bytes(int tile_stride, int k, int d) {
    size_t b = 2 * (size_t)tile_stride;
    b += (size_t)QCAP * 32 * (sizeof(T) + 4);
    b += (size_t)k * 32 * (sizeof(T) + 4);
    b += (size_t)32 * knn_qrow_elems<T>(d) * sizeof(T);
    b += (size_t)MAX_WIDE_LEVELS * 32 * 4;
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

// ---------------------------------------------------------------------------
// Per-lane top-k: a max-heap by (d2, id) in shared memory, laid out [slot][lane]
// (bank = lane, conflict-free for any per-lane slot).  Root = current k-th (worst).
template <typename T>
__device__ __forceinline__ void heap_replace_root(T* hd, u32* hi, int K, int lane, T v, u32 id) {
    int pos = 0;
    for (;;) {
        const int l = 2 * pos + 1;
        if (l >= K) break;
        int c = l;
        T cd = hd[l * 32 + lane];
        u32 ci = hi[l * 32 + lane];
        if (l + 1 < K) {
            const T rd = hd[(l + 1) * 32 + lane];
            const u32 ri = hi[(l + 1) * 32 + lane];
            if (worse<T>(rd, ri, cd, ci)) { c = l + 1; cd = rd; ci = ri; }
        }
        if (!worse<T>(cd, ci, v, id)) break;
        hd[pos * 32 + lane] = cd;
        hi[pos * 32 + lane] = ci;
        pos = c;
    }
    hd[pos * 32 + lane] = v;
    hi[pos * 32 + lane] = id;
}

template <typename T, int D>
struct WarpState {
    T q[D];
    T worst;        // root of the heap for real queries, -1 for padding lanes
    T bound;        // warp max of worst
    int cnt;        // queue fill
    bool active;
    u32 n_push, n_insert;   // statistics
};

template <typename T, int D>
__device__ __forceinline__ void flush_queues(WarpState<T, D>& s, const T* qd, const u32* qi, T* hd, u32* hi, T* qworst, int K,
                                             int lane, u32& n_flush, u32& n_iter) {
    const int maxc = __reduce_max_sync(FULL, s.cnt);
    n_flush++;
    n_iter += maxc;
#pragma unroll 1
    for (int t = 0; t < maxc; t++) {
        if (t < s.cnt) {
            const T v = qd[t * 32 + lane];
            const u32 id = qi[t * 32 + lane];
            if (worse<T>(hd[lane], hi[lane], v, id)) {
                heap_replace_root<T>(hd, hi, K, lane, v, id);
                s.n_insert++;
            }
        }
    }
    s.cnt = 0;
    if (s.active) s.worst = hd[lane];
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

template <typename T, int D>
__global__ void __launch_bounds__(WARPS * 32, 4) knn_kernel(const KnnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 ql = p.leaf_begin + (i64)blockIdx.x * WARPS + wib;
    if (ql >= p.leaf_end) return;
    const int K = p.k;

    // ---- per-warp shared memory carve-up (integer offsets from the shared base)
    constexpr int PER16 = 16 / (int)sizeof(T);
    constexpr int RS = (D + 1 + PER16 - 1) / PER16 * PER16;   // query-table row stride (elements)
    const u32 wo = (u32)wib * (u32)p.warp_smem;
    const u32 o_q = 2u * (u32)p.tile_stride;
    T* qd = (T*)(smem + wo + o_q);                          // [QCAP][32]
    u32* qi = (u32*)(qd + QCAP * 32);                       // [QCAP][32]
    T* hd = (T*)(qi + QCAP * 32);                           // [K][32] heap distances
    u32* hi = (u32*)(hd + K * 32);                          // [K][32] heap ids
    T* qtab = (T*)(hi + K * 32);                            // [32][RS]: query coordinates, k-th distance
    u32* sneed = (u32*)(qtab + 32 * RS);                    // [MAX_WIDE_LEVELS][32] per-child query masks
    i64* snode = (i64*)(sneed + MAX_WIDE_LEVELS * 32);      // [MAX_WIDE_LEVELS]
    u32* smask = (u32*)(snode + MAX_WIDE_LEVELS);           // [MAX_WIDE_LEVELS]
    int* slevel = (int*)(smask + MAX_WIDE_LEVELS);          // [MAX_WIDE_LEVELS]
    u64* bars = (u64*)(slevel + MAX_WIDE_LEVELS);           // [2]
    const u32 tile_sa = smem_u32(smem + wo);                // shared address of tile buffer 0
    T* qworst = qtab + lane * RS + D;

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll 1
    for (int m = 0; m < K; m++) { hd[m * 32 + lane] = inf_t<T>(); hi[m * 32 + lane] = 0xffffffffu; }
#pragma unroll
    for (int j = 0; j < RS; j++) qtab[lane * RS + j] = j == D ? (T)-1 : (T)0;
    __syncwarp();

    WarpState<T, D> s;
    s.cnt = 0;
    s.active = false;
    s.worst = (T)-1;
    s.bound = inf_t<T>();
    s.n_push = 0;
    s.n_insert = 0;
    u32 n_flush = 0, n_iter = 0, n_loaded = 0, n_eval = 0, n_needy = 0, n_slow = 0;
    unsigned long long n_pairs = 0;
    // the query box: tight bounds of the own leaf (level-0 box arrays)
    T qlo[D], qhi[D];
    {
        const T* lo0 = (const T*)p.box[0];
        const T* hi0 = lo0 + (i64)D * p.count[0];
#pragma unroll
        for (int j = 0; j < D; j++) {
            qlo[j] = lo0[(i64)j * p.count[0] + ql];
            qhi[j] = hi0[(i64)j * p.count[0] + ql];
            s.q[j] = (T)0;
        }
    }
    u32 my_id = 0xffffffffu;

    // Open a wide node: lane j takes child j (level `hc`), tests box-to-box against the warp bound, then
    // for every query in `qmask` the distance from the query to the child's box against the query's own
    // k-th distance.  Returns the children some query needs; per-child query masks go to need_out[lane].
    auto open_node = [&](int hc, i64 node, i64 excl, u32 qmask, u32* need_out) -> u32 {
        const i64 cnt = p.count[hc];
        const i64 c = node * FANOUT + lane;
        const bool ok = c < cnt && c != excl;
        const T* lo = (const T*)p.box[hc];
        const T* hi_ = lo + (i64)D * cnt;
        T blo[D], bhi[D];
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < D; j++) {
            blo[j] = ok ? lo[(i64)j * cnt + c] : (T)0;
            bhi[j] = ok ? hi_[(i64)j * cnt + c] : (T)0;
            T a = blo[j] - qhi[j], b = qlo[j] - bhi[j];
            T g = a > b ? a : b;
            g = g > (T)0 ? g : (T)0;
            acc = fma_t<T>(g, g, acc);
        }
        const bool pass = ok && acc <= s.bound;
        u32 mask = 0;
        if (__any_sync(FULL, pass)) {
            u32 m = qmask;
#pragma unroll 1
            while (m) {
                const int q = __ffs(m) - 1;
                m &= m - 1;
                const T* row = qtab + q * RS;
                T pd = (T)0;
#pragma unroll
                for (int j = 0; j < D; j++) {
                    const T x = row[j];
                    T a = blo[j] - x, b = x - bhi[j];
                    T g = a > b ? a : b;
                    g = g > (T)0 ? g : (T)0;
                    pd = fma_t<T>(g, g, pd);
                }
                if (pass && pd <= row[D]) mask |= 1u << q;
            }
        }
        need_out[lane] = mask;
        return __ballot_sync(FULL, mask != 0);
    };

    // ---- traversal state: ancestors of the own leaf, bottom-up; a small stack for descents
    int anc = 0;    // 0: own leaf not yet emitted; g >= 1: next ancestor level to open
    int sp = 0;
    u32 cand_need = FULL;
    auto next_candidate = [&]() -> i64 {
        if (anc == 0) { anc = 1; cand_need = FULL; return ql; }
        for (;;) {
            if (sp == 0) {
                if (anc > p.n_levels) return -1;
                // open the children of the level-`anc` ancestor, except the branch already done
                if (__any_sync(FULL, s.cnt > 0)) flush_queues<T, D>(s, qd, qi, hd, hi, qworst, K, lane, n_flush, n_iter);
                const i64 node = ql >> (5 * anc), excl = ql >> (5 * (anc - 1));
                const u32 m = open_node(anc - 1, node, excl, FULL, sneed);
                if (lane == 0) { smask[0] = m; snode[0] = node; slevel[0] = anc; }
                __syncwarp();
                sp = 1;
                anc++;
                continue;
            }
            const int top = sp - 1;
            u32 m = smask[top];
            if (m == 0) { sp--; continue; }
            const int h = slevel[top];
            const i64 node = snode[top];
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const u32 need = sneed[top * 32 + bit];
            __syncwarp();
            if (lane == 0) smask[top] = m;
            __syncwarp();
            const i64 c = node * FANOUT + bit;
            if (h == 1) { cand_need = need; return c; }
            // descend: open the children of node c (a level h-1 node) for the queries that need it
            {
                const u32 mm = open_node(h - 2, c, -1, need, sneed + sp * 32);
                if (lane == 0) { smask[sp] = mm; snode[sp] = c; slevel[sp] = h - 1; }
                __syncwarp();
                sp++;
            }
        }
    };

    u32 phase = 0;
    i64 pending = -1;
    u32 pending_need = 0;
    int bi = 0;                 // buffer the next issue goes to
    for (;;) {
        // no look-ahead while the own leaf is pending: box tests need the queries and their first bounds
        const bool lookahead = pending != ql;
        const i64 cand = lookahead ? next_candidate() : -1;
        const u32 cneed = cand_need;
        if (cand >= 0) {
            __syncwarp();       // every lane is done reading buffer bi
            if (lane == 0) {
                mbar_expect_tx(&bars[bi], (u32)p.tile_stride);
                tma_load_1d(smem + wo + (u32)bi * (u32)p.tile_stride, p.tiles + cand * (i64)p.tile_stride, (u32)p.tile_stride,
                            &bars[bi]);
            }
            n_loaded++;
        }
        if (pending >= 0) {
            const int bw = bi ^ 1;
            while (!mbar_try_wait(&bars[bw], (phase >> bw) & 1u)) {}
            phase ^= (1u << bw);
            const unsigned char* tile = smem + wo + (u32)bw * (u32)p.tile_stride;
            const T* hdr = (const T*)tile;
            const T* tcoord = (const T*)(tile + p.tile_hdr_bytes);
            const u32* ids = (const u32*)(tile + p.tile_ids_off);
            u32 need;
            if (pending == ql) {
                // own leaf: pick up this lane's query point and publish it in the query table
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
                need = FULL;
            } else {
                // refresh the need mask with the current k-th distances
                T acc = (T)0;
#pragma unroll
                for (int j = 0; j < D; j++) {
                    T a = hdr[j] - s.q[j], b = s.q[j] - hdr[D + j];
                    T g = a > b ? a : b;
                    g = g > (T)0 ? g : (T)0;
                    acc = fma_t<T>(g, g, acc);
                }
                need = __ballot_sync(FULL, acc <= s.worst) & pending_need;
                n_needy += __popc(need);
            }
            const int n_need = __popc(need);
            if (n_need > SPARSE_MAX) {
                // ---- dense: every lane evaluates all 32 candidates (packed f32x2)
                n_eval++;
                n_pairs += (unsigned long long)LEAF * 32;
                const u32 coords = tile_sa + (u32)bw * (u32)p.tile_stride + (u32)p.tile_hdr_bytes;
                // each chunk may push 2*CHUNK_PB entries per lane; the sparse path can leave queues fuller than that allows
                if (__any_sync(FULL, s.cnt > QCAP - 2 * CHUNK_PB))
                    flush_queues<T, D>(s, qd, qi, hd, hi, qworst, K, lane, n_flush, n_iter);
#pragma unroll 1
                for (int pb0 = 0; pb0 < LEAF / 2; pb0 += CHUNK_PB) {
                    T v[2 * CHUNK_PB];
                    chunk_dist<D>(s.q, coords, p.pair_block, pb0, v);
                    T mn = v[0];
#pragma unroll
                    for (int e = 1; e < 2 * CHUNK_PB; e++) mn = v[e] < mn ? v[e] : mn;
                    if (__any_sync(FULL, mn <= s.worst)) {
                        n_slow++;
#pragma unroll
                        for (int e = 0; e < 2 * CHUNK_PB; e++) {
                            if (v[e] <= s.worst) {
                                qd[s.cnt * 32 + lane] = v[e];
                                qi[s.cnt * 32 + lane] = ids[2 * pb0 + e];
                                s.cnt++;
                                s.n_push++;
                            }
                        }
                        if (__any_sync(FULL, s.cnt > QCAP - 2 * CHUNK_PB))
                            flush_queues<T, D>(s, qd, qi, hd, hi, qworst, K, lane, n_flush, n_iter);
                    }
                }
            } else if (n_need > 0) {
                // ---- sparse: one needy query at a time, lane l evaluates candidate l
                n_eval++;
                n_pairs += (unsigned long long)LEAF * n_need;
                T cc[D];
#pragma unroll
                for (int j = 0; j < D; j++) cc[j] = tcoord[(lane >> 1) * p.pair_block + 2 * j + (lane & 1)];
                const u32 cid = ids[lane];
                u32 m = need;
#pragma unroll 1
                while (m) {
                    const int q = __ffs(m) - 1;
                    m &= m - 1;
                    const T* row = qtab + q * RS;
                    T v = (T)0;
#pragma unroll
                    for (int j = 0; j < D; j++) {
                        const T df = row[j] - cc[j];
                        v = fma_t<T>(df, df, v);
                    }
                    u32 hm = __ballot_sync(FULL, v <= row[D]);
#pragma unroll 1
                    while (hm) {
                        const int b = __ffs(hm) - 1;
                        hm &= hm - 1;
                        const T vv = __shfl_sync(FULL, v, b);
                        const u32 ii = __shfl_sync(FULL, cid, b);
                        if (__shfl_sync(FULL, s.cnt, q) >= QCAP)
                            flush_queues<T, D>(s, qd, qi, hd, hi, qworst, K, lane, n_flush, n_iter);
                        if (lane == q) {
                            qd[s.cnt * 32 + lane] = vv;
                            qi[s.cnt * 32 + lane] = ii;
                            s.cnt++;
                            s.n_push++;
                        }
                    }
                }
            }
        }
        if (!lookahead) { pending = -1; continue; }
        if (cand < 0) break;
        pending = cand;
        pending_need = cneed;
        bi ^= 1;
    }
    if (__any_sync(FULL, s.cnt > 0)) flush_queues<T, D>(s, qd, qi, hd, hi, qworst, K, lane, n_flush, n_iter);

    // ---- heap -> ascending order in place (heapsort), then write the rows
#pragma unroll 1
    for (int m = K - 1; m >= 1; m--) {
        const T td = hd[m * 32 + lane];
        const u32 ti = hi[m * 32 + lane];
        hd[m * 32 + lane] = hd[lane];
        hi[m * 32 + lane] = hi[lane];
        heap_replace_root<T>(hd, hi, m, lane, td, ti);
    }
    const i64 pos_row = (ql - p.leaf_begin) * LEAF + lane;
    if (s.active) {
        const i64 row = p.row_order == 0 ? (i64)my_id : pos_row;
        u32* io = p.idx_out + row * K;
        T* dd = (T*)p.d2_out + row * K;
#pragma unroll 1
        for (int m = 0; m < K; m++) { io[m] = hi[m * 32 + lane]; dd[m] = hd[m * 32 + lane]; }
    } else if (p.row_order == 1) {
#pragma unroll 1
        for (int m = 0; m < K; m++) { p.idx_out[pos_row * K + m] = 0xffffffffu; ((T*)p.d2_out)[pos_row * K + m] = inf_t<T>(); }
    }
    if (p.stats != nullptr) {
        const u32 pushes = __reduce_add_sync(FULL, s.n_push), inserts = __reduce_add_sync(FULL, s.n_insert);
        if (lane == 0) {
            atomicAdd(p.stats + 0, n_pairs);                                         // (query, candidate) pairs evaluated
            atomicAdd(p.stats + 1, (unsigned long long)n_loaded);                       // tiles fetched by TMA
            atomicAdd(p.stats + 2, (unsigned long long)n_eval);                         // tiles evaluated
            atomicAdd(p.stats + 3, (unsigned long long)n_needy);                        // lanes that needed an evaluated/skipped tile
            atomicAdd(p.stats + 4, (unsigned long long)pushes);                         // queue pushes
            atomicAdd(p.stats + 5, (unsigned long long)inserts);                        // heap insertions
            atomicAdd(p.stats + 6, (unsigned long long)n_iter);                         // flush iterations (n_flush in high half)
            atomicAdd(p.stats + 7, (unsigned long long)n_slow);                         // chunks that took the push path
        }
        (void)n_flush;
    }
}

template <typename T, int D>
static int launch_knn(const KnnParams& p, cudaStream_t st) {
    i64 n_q = p.leaf_end - p.leaf_begin;
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
