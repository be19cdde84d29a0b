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
    unsigned long long* pairs;
    int warp_smem;
};

template <typename T>
static inline int knn_warp_smem_bytes(int tile_stride) {
    size_t b = 2 * (size_t)tile_stride;
    b += (size_t)QCAP * 32 * (sizeof(T) + 4);
    b += (size_t)MAX_WIDE_LEVELS * 32 * sizeof(T);
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

// ---------------------------------------------------------------------------
// KCAP <= 32: list in registers, fully unrolled; KCAP == 128: list in local
// memory, rolled loops.
template <typename T, int D, int KCAP>
struct WarpState {
    T dk[KCAP];     // sorted ascending in the TOP k slots (worst at KCAP-1)
    u32 ik[KCAP];
    T q[D];
    T worst;        // dk[KCAP-1] for real queries, -1 for padding lanes
    T bound;        // warp max of worst
    int cnt;        // queue fill
    int lo;         // KCAP - k
    bool active;
};

template <typename T, int D, int KCAP>
__device__ __forceinline__ void insert_sorted(WarpState<T, D, KCAP>& s, T v, u32 id) {
    if constexpr (KCAP <= 32) {
        bool prev = true;
#pragma unroll
        for (int m = KCAP - 1; m >= 1; m--) {
            if (m > s.lo) {
                bool sh = worse<T>(s.dk[m - 1], s.ik[m - 1], v, id);
                T nd = sh ? s.dk[m - 1] : v;
                u32 ni = sh ? s.ik[m - 1] : id;
                if (prev) { s.dk[m] = nd; s.ik[m] = ni; }
                prev = sh;
            }
        }
#pragma unroll
        for (int m = KCAP - 1; m >= 0; m--)
            if (m == s.lo && prev) { s.dk[m] = v; s.ik[m] = id; }
    } else {
        int m = KCAP - 1;
#pragma unroll 1
        while (m > s.lo && worse<T>(s.dk[m - 1], s.ik[m - 1], v, id)) {
            s.dk[m] = s.dk[m - 1];
            s.ik[m] = s.ik[m - 1];
            m--;
        }
        s.dk[m] = v;
        s.ik[m] = id;
    }
}

template <typename T, int D, int KCAP>
__device__ __forceinline__ void flush_queues(WarpState<T, D, KCAP>& s, const T* qd, const u32* qi, int lane) {
    int maxc = __reduce_max_sync(FULL, s.cnt);
#pragma unroll 1
    for (int t = 0; t < maxc; t++) {
        if (t < s.cnt) {
            T v = qd[t * 32 + lane];
            u32 id = qi[t * 32 + lane];
            if (worse<T>(s.dk[KCAP - 1], s.ik[KCAP - 1], v, id)) insert_sorted<T, D, KCAP>(s, v, id);
        }
    }
    s.cnt = 0;
    if (s.active) s.worst = s.dk[KCAP - 1];
    s.bound = warp_max_t<T>(s.worst);
}

// distances of this lane's query to the 2*CHUNK_PB candidates of one chunk
template <int D>
__device__ __forceinline__ void chunk_dist(const float* q, const unsigned char* coords, int pair_block, int pb0, float* out) {
    const float4* blk = (const float4*)coords;
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
            float4 f = blk[pb * PB4 + v4];
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
__device__ __forceinline__ void chunk_dist(const double* q, const unsigned char* coords, int pair_block, int pb0, double* out) {
    const double* blk = (const double*)coords;
#pragma unroll
    for (int c = 0; c < CHUNK_PB; c++) {
        const int pb = pb0 + c;
        double va = 0.0, vb = 0.0;
#pragma unroll
        for (int j = 0; j < D; j++) {
            double2 cc = *(const double2*)(blk + pb * pair_block + 2 * j);
            double da = q[j] - cc.x, db = q[j] - cc.y;
            va = __fma_rn(da, da, va);
            vb = __fma_rn(db, db, vb);
        }
        out[2 * c] = va;
        out[2 * c + 1] = vb;
    }
}

template <typename T, int D, int KCAP>
__global__ void __launch_bounds__(WARPS * 32) knn_kernel(const KnnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 ql = p.leaf_begin + (i64)blockIdx.x * WARPS + wib;
    if (ql >= p.leaf_end) return;

    // ---- per-warp shared memory carve-up
    unsigned char* wb = smem + (size_t)wib * p.warp_smem;
    unsigned char* tbuf[2] = {wb, wb + p.tile_stride};
    T* qd = (T*)(wb + 2 * p.tile_stride);                   // [QCAP][32]
    u32* qi = (u32*)(qd + QCAP * 32);                       // [QCAP][32]
    T* sdist = (T*)(qi + QCAP * 32);                        // [MAX_WIDE_LEVELS][32]
    i64* snode = (i64*)(sdist + MAX_WIDE_LEVELS * 32);      // [MAX_WIDE_LEVELS]
    u32* smask = (u32*)(snode + MAX_WIDE_LEVELS);           // [MAX_WIDE_LEVELS]
    int* slevel = (int*)(smask + MAX_WIDE_LEVELS);          // [MAX_WIDE_LEVELS]
    u64* bars = (u64*)(slevel + MAX_WIDE_LEVELS);           // [2]

    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    WarpState<T, D, KCAP> s;
    s.cnt = 0;
    s.lo = KCAP - p.k;
    s.active = false;
    s.worst = (T)-1;
    s.bound = inf_t<T>();
    if constexpr (KCAP <= 32) {
#pragma unroll
        for (int m = 0; m < KCAP; m++) { s.dk[m] = inf_t<T>(); s.ik[m] = 0xffffffffu; }
    } else {
#pragma unroll 1
        for (int m = 0; m < KCAP; m++) { s.dk[m] = inf_t<T>(); s.ik[m] = 0xffffffffu; }
    }
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

    auto boxdist = [&](int h, i64 c) -> T {
        const T* lo = (const T*)p.box[h];
        const i64 cnt = p.count[h];
        const T* hi = lo + (i64)D * cnt;
        T acc = (T)0;
#pragma unroll
        for (int j = 0; j < D; j++) {
            T a = lo[(i64)j * cnt + c] - qhi[j], b = qlo[j] - hi[(i64)j * cnt + c];
            T g = a > b ? a : b;
            g = g > (T)0 ? g : (T)0;
            acc = fma_t<T>(g, g, acc);
        }
        return acc;
    };

    // ---- traversal state: ancestors of the own leaf, bottom-up; a small stack for descents
    int anc = 0;    // 0: own leaf not yet emitted; g >= 1: next ancestor level to open
    int sp = 0;
    auto next_candidate = [&]() -> i64 {
        if (anc == 0) { anc = 1; return ql; }
        for (;;) {
            if (sp == 0) {
                if (anc > p.n_levels) return -1;
                // open the children of the level-`anc` ancestor, except the branch already done
                if (__any_sync(FULL, s.cnt > 0)) flush_queues<T, D, KCAP>(s, qd, qi, lane);
                const i64 node = ql >> (5 * anc), excl = ql >> (5 * (anc - 1));
                const i64 c = node * FANOUT + lane;
                const bool ok = c < p.count[anc - 1] && c != excl;
                const T dist = ok ? boxdist(anc - 1, c) : inf_t<T>();
                const u32 m = __ballot_sync(FULL, ok && dist <= s.bound);
                sdist[lane] = dist;
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
            const T dist = sdist[top * 32 + bit];
            __syncwarp();
            if (lane == 0) smask[top] = m;
            __syncwarp();
            if (!(dist <= s.bound)) continue;
            const i64 c = node * FANOUT + bit;
            if (h == 1) return c;
            // descend: open the children of node c (level h-1 node, children at level h-2)
            {
                const i64 cc = c * FANOUT + lane;
                const bool ok = cc < p.count[h - 2];
                const T dd = ok ? boxdist(h - 2, cc) : inf_t<T>();
                const u32 mm = __ballot_sync(FULL, ok && dd <= s.bound);
                sdist[sp * 32 + lane] = dd;
                if (lane == 0) { smask[sp] = mm; snode[sp] = c; slevel[sp] = h - 1; }
                __syncwarp();
                sp++;
            }
        }
    };

    u32 phase = 0;
    i64 pending = -1;
    int bi = 0;                 // buffer the next issue goes to
    unsigned long long tiles_done = 0;
    for (;;) {
        const i64 cand = next_candidate();
        if (cand >= 0) {
            __syncwarp();       // every lane is done reading buffer bi
            if (lane == 0) {
                mbar_expect_tx(&bars[bi], (u32)p.tile_stride);
                tma_load_1d(tbuf[bi], p.tiles + cand * (i64)p.tile_stride, (u32)p.tile_stride, &bars[bi]);
            }
        }
        if (pending >= 0) {
            const int bw = bi ^ 1;
            while (!mbar_try_wait(&bars[bw], (phase >> bw) & 1u)) {}
            phase ^= (1u << bw);
            const unsigned char* tile = tbuf[bw];
            const T* hdr = (const T*)tile;
            bool go;
            if (pending == ql) {
                // own leaf: pick up this lane's query point
                const T* c = (const T*)(tile + p.tile_hdr_bytes);
#pragma unroll
                for (int j = 0; j < D; j++) s.q[j] = c[(lane >> 1) * p.pair_block + 2 * j + (lane & 1)];
                my_id = ((const u32*)(tile + p.tile_ids_off))[lane];
                s.active = my_id != 0xffffffffu;
                s.worst = s.active ? inf_t<T>() : (T)-1;
                go = true;
            } else {
                // per-lane refinement: distance from my query to the candidate leaf's box
                T acc = (T)0;
#pragma unroll
                for (int j = 0; j < D; j++) {
                    T a = hdr[j] - s.q[j], b = s.q[j] - hdr[D + j];
                    T g = a > b ? a : b;
                    g = g > (T)0 ? g : (T)0;
                    acc = fma_t<T>(g, g, acc);
                }
                go = __any_sync(FULL, acc <= s.worst);
            }
            if (go) {
                tiles_done++;
                const unsigned char* coords = tile + p.tile_hdr_bytes;
                const u32* ids = (const u32*)(tile + p.tile_ids_off);
#pragma unroll 1
                for (int pb0 = 0; pb0 < LEAF / 2; pb0 += CHUNK_PB) {
                    T v[2 * CHUNK_PB];
                    chunk_dist<D>(s.q, coords, p.pair_block, pb0, v);
#pragma unroll
                    for (int e = 0; e < 2 * CHUNK_PB; e++) {
                        if (v[e] <= s.worst) {
                            qd[s.cnt * 32 + lane] = v[e];
                            qi[s.cnt * 32 + lane] = ids[2 * pb0 + e];
                            s.cnt++;
                        }
                    }
                    if (__any_sync(FULL, s.cnt > QCAP - 2 * CHUNK_PB)) flush_queues<T, D, KCAP>(s, qd, qi, lane);
                }
            }
        }
        if (cand < 0) break;
        pending = cand;
        bi ^= 1;
    }
    if (__any_sync(FULL, s.cnt > 0)) flush_queues<T, D, KCAP>(s, qd, qi, lane);

    // ---- write the rows
    const i64 pos_row = (ql - p.leaf_begin) * LEAF + lane;
    if (s.active) {
        const i64 row = p.row_order == 0 ? (i64)my_id : pos_row;
        u32* io = p.idx_out + row * p.k;
        T* dd = (T*)p.d2_out + row * p.k;
        if constexpr (KCAP <= 32) {
#pragma unroll
            for (int m = 0; m < KCAP; m++)
                if (m >= s.lo) { io[m - s.lo] = s.ik[m]; dd[m - s.lo] = s.dk[m]; }
        } else {
#pragma unroll 1
            for (int m = s.lo; m < KCAP; m++) { io[m - s.lo] = s.ik[m]; dd[m - s.lo] = s.dk[m]; }
        }
    } else if (p.row_order == 1) {
#pragma unroll 1
        for (int m = 0; m < p.k; m++) { p.idx_out[pos_row * p.k + m] = 0xffffffffu; ((T*)p.d2_out)[pos_row * p.k + m] = inf_t<T>(); }
    }
    const u32 act = __ballot_sync(FULL, s.active);
    if (p.pairs != nullptr && lane == 0) atomicAdd(p.pairs, tiles_done * (unsigned long long)LEAF * __popc(act));
}

template <typename T, int D, int KCAP>
static int launch_knn(const KnnParams& p, cudaStream_t st) {
    i64 n_q = p.leaf_end - p.leaf_begin;
    if (n_q <= 0) return AL_OK;
    size_t smem = (size_t)p.warp_smem * WARPS;
    if (smem > 48 * 1024)
        AL_CUDA(cudaFuncSetAttribute(knn_kernel<T, D, KCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<T, D, KCAP><<<(unsigned)((n_q + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p);
    AL_CHECK_LAUNCH();
    return AL_OK;
}

template <typename T, int D>
static int knn_dispatch_k(const KnnParams& p, cudaStream_t st) {
    if (p.k <= 8) return launch_knn<T, D, 8>(p, st);
    if (p.k <= 16) return launch_knn<T, D, 16>(p, st);
    if (p.k <= 20) return launch_knn<T, D, 20>(p, st);
    if (p.k <= 32) return launch_knn<T, D, 32>(p, st);
    return launch_knn<T, D, 128>(p, st);
}
#endif  // __CUDACC__
