// float32 specialisation of the exact kNN search kernel (the generic one is in
// knn_kernel.cuh and still serves float64): replaces pykdtree's
// KDTree.query(P_transform, k=k_den, sqr_dists=True) (reference
// astrolink/astrolink.py:279), with the density of astrolink.py:293-297 fused
// into its epilogue.
//
// Same algorithm, same arithmetic, same results as the generic kernel -- one warp
// per 32-query leaf, 32-ary box hierarchy with per-child need masks, TMA tile
// copies, per-lane queues merged into per-lane max-heaps in shared memory --
// rebuilt around the instruction and stall counts of the ncu captures of this
// round (profiles/r02_knn_*):
//   * every shared-memory access goes through a 32-bit shared address computed once;
//     tile geometry is a compile-time function of D;
//   * sparse tiles: a half-warp holds the whole tile, two candidates per lane as
//     packed f32x2 operands straight from the pair-interleaved tile, so one pass
//     evaluates two needy queries (one per half-warp) with D FADD2 + D FFMA2;
//   * box tests: the index stores a node's box as interleaved (lo_j, hi_j) pairs, one
//     128-bit-aligned row per node: a lane reads its child's box with 128-bit loads,
//     a per-axis gap is one FADD2 (q - lo, q - hi), one 3-input max, one FFMA, and two
//     needy queries are tested per iteration (two independent chains);
//   * the children of a node -- leaves and wide nodes alike -- are visited in the order
//     of a per-lane key register, one warp reduction per pick: the child most queries
//     need first, then the nearest box (the k-th distances tighten early; 16 bits of
//     the keys of an interrupted wide node are spilled with its need masks);
//   * the tile copy is issued by an elected lane and the warp's shared base is a
//     uniform register the compiler knows about (a `lane == 0` test around the
//     uniform-operand copy costs a divergence-proof broadcast loop per tile); the
//     lane number is opaque so that it is not rematerialised from %tid in the loops;
//   * the sparse passes locate the next pass's query rows under the loads of the
//     current one (the bit-scan chain and the FMA chain overlap);
//   * a 4-ary heap with the four children of a node adjacent in shared memory (two
//     128-bit loads per level, two levels for k = 20); the root lives in a register;
//   * the kernel is latency bound at the occupancy its shared memory allows, so the
//     per-warp footprint is cut to fit 6 CTAs (24 warps) per SM at d = 6, k = 20:
//     ONE tile buffer (24 resident warps hide the copy of the next tile better than
//     a second buffer at 20 warps did), need masks of the list being walked in
//     registers (lane = child; spilled only on a descent), tiles without a header,
//     a 7-entry queue.
#pragma once
#include <stdlib.h>

#include "knn_kernel.cuh"

#ifndef KF_SPARSE_MAX
#define KF_SPARSE_MAX 12       // tiles needed by at most this many queries take the sparse path
#endif
#ifndef KF_QCAP
#define KF_QCAP 7
#endif
#ifndef KF_OPEN2
#define KF_OPEN2 1             // box tests: two needy queries per iteration (two independent chains)
#endif
#ifndef KF_MINBLOCKS
#define KF_MINBLOCKS 6         // CTAs per SM the register allocation is held to for d <= 8 (80 registers)
#endif

namespace kf {

constexpr int QC = KF_QCAP;
static_assert(QC >= 5, "the queue must hold a dense chunk of four candidates on top of one pending entry");

template <int D>
struct Geo {
    static constexpr int PB = (2 * D + 3) / 4 * 4;                  // floats per pair block (two candidates, interleaved)
    static constexpr int PB4 = PB / 4;                              // 128-bit words per pair block
    static constexpr int IDS = (LEAF / 2) * PB * 4;                 // byte offset of the ids inside a tile
    static constexpr int TS = (IDS + LEAF * 4 + 127) / 128 * 128;   // tile stride (bytes)
    static constexpr int RS = (D + 1 + 3) / 4 * 4;                  // floats per query-table row (coordinates, k-th distance)
    static constexpr int RS4 = RS / 4;
    static constexpr int BS = PB;                                   // floats per box row of the index
    static constexpr int MINB = D <= 8 ? KF_MINBLOCKS : (D <= 12 ? 4 : 3);
};

// per-warp shared memory (bytes from the warp's base).  Everything except the heap's SIZE is a compile-time function of D, and
// the heap comes last but one, so every address the hot loops use is the warp's base plus a constant (plus a lane term):
//   tile | queue [QC][32] u64 | query table [32][RS] f32 | own leaf box [D] (lo, hi) | stack masks [8] | mbarrier |
//   heap quads [ceil((K-1)/4)][2][32][2] u64 (node h's four children in quad h; the root in a register) | spilled need masks [n_levels - 1][32] u32
__host__ __device__ static inline int spill_rows(int n_levels) { return n_levels > 1 ? n_levels - 1 : 1; }
// bytes of the per-warp heap below the root: ceil((k - 1) / 4) quads of 1 KB
__host__ __device__ static inline unsigned heap_bytes(int k) { return (unsigned)((k + 2) >> 2) * 1024u; }
template <int D>
struct Smem {
    typedef Geo<D> G;
    static constexpr int QUEUE = G::TS;
    static constexpr int QT = QUEUE + QC * 256;
    static constexpr int QBOX = QT + 32 * G::RS * 4;
    static constexpr int STK = QBOX + G::BS * 4;
    static constexpr int BAR = STK + MAX_WIDE_LEVELS * 4;
    static constexpr int KEYS = BAR + 16;                      // 16-bit visiting-order keys of the spilled level-2 and level-3 entries, [2][32]
    static constexpr int HEAP = KEYS + 128;
    static_assert(HEAP % 16 == 0, "heap quads are read with 128-bit loads");
};
template <int D>
static inline int warp_smem_bytes(int k, int n_levels) {
    size_t b = (size_t)Smem<D>::HEAP;
    b += heap_bytes(k);
    b += (size_t)spill_rows(n_levels) * 128;
    return (int)al_align(b, 16);      // d=6, k=20, 4 levels: 9 312 B -> 6 CTAs of 4 warps per SM
}

#ifdef __CUDACC__
// ---- shared memory through 32-bit shared addresses -------------------------------------------------------------------
__device__ __forceinline__ u32 lds32(u32 a) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float ldsf(u32 a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ u64 lds64(u32 a) {
    u64 v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds64x2(u32 a, u64& x, u64& y) { asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(a)); }
__device__ __forceinline__ void lds32x2(u32 a, u32& x, u32& y) { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
__device__ __forceinline__ void ldsf2(u32 a, float& x, float& y) { asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); }
__device__ __forceinline__ u32 lds16(u32 a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return (u32)v;
}
__device__ __forceinline__ void sts16(u32 a, u32 v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void stsf(u32 a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts64(u32 a, u64 v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts64x2(u32 a, u64 x, u64 y) { asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(a), "l"(x), "l"(y) : "memory"); }

__device__ __forceinline__ void mbar_init_sa(u32 bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_sa(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d_sa(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok));
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_sa(u32 bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ u64 make_key(float d, u32 id) { return ((u64)__float_as_uint(d) << 32) | id; }
__device__ __forceinline__ float key_dist(u64 k) { return __uint_as_float((u32)(k >> 32)); }
__device__ __forceinline__ u32 key_id(u64 k) { return (u32)k; }

// a row of the query table (uniform or per-half-warp address: a broadcast load)
template <int D>
__device__ __forceinline__ void load_qrow(u32 row_sa, float* r) {
#pragma unroll
    for (int v = 0; v < Geo<D>::RS4; v++) {
        const float4 f = lds_f4(row_sa + 16u * v);
        r[4 * v] = f.x; r[4 * v + 1] = f.y; r[4 * v + 2] = f.z; r[4 * v + 3] = f.w;
    }
}

// Replaces the root of a 4-ary max-heap of `size` nodes by v and restores the heap.  Node 0 is `root` (a register).
// Node h's children 4h + 1 .. 4h + 4 are quad h: two 16-byte halves (children 0-1, children 2-3), each
// laid out [lane][2] (conflict-free 128-bit loads), a quad every 1024 bytes from quad0_sa (this lane's quad 0, first half).
// For k = 20 the heap is two levels deep: a replacement costs at most two quads.  Missing children hold key 0.
__device__ __forceinline__ u32 quad_slot_sa(u32 quad0_sa, u32 quad_off, u32 r) { return quad0_sa + quad_off + (r >> 1) * 512u + (r & 1u) * 8u; }
template <bool SHRUNK>
__device__ __forceinline__ bool heap4_level(u32 quad0_sa, u32 off, u32 q_end, int size, u64 v, u64& ck, u32& r) {
    u64 c0, c1, c2, c3;
    lds64x2(quad0_sa + off, c0, c1);
    lds64x2(quad0_sa + off + 512u, c2, c3);
    if (SHRUNK && off + 1024u == q_end) {       // the last quad of a heap that is being shrunk: slots past `size` hold extracted keys
        const int valid = size - 1 - (int)(off >> 8);
        if (valid < 2) c1 = 0ull;
        if (valid < 3) c2 = 0ull;
        if (valid < 4) c3 = 0ull;
    }
    const bool r01 = c1 > c0, r23 = c3 > c2;
    const u64 m01 = r01 ? c1 : c0, m23 = r23 ? c3 : c2;
    const bool hi = m23 > m01;
    ck = hi ? m23 : m01;
    r = hi ? (r23 ? 3u : 2u) : (r01 ? 1u : 0u);
    return ck > v;
}
template <bool SHRUNK>
__device__ __forceinline__ void heap4_replace(u64& root, u32 quad0_sa, int size, u64 v) {
    const u32 q_end = (u32)((size + 2) >> 2) * 1024u;       // quads with a node below `size`
    if (q_end == 0u) { root = v; return; }
    u64 ck;
    u32 r;
    if (!heap4_level<SHRUNK>(quad0_sa, 0u, q_end, size, v, ck, r)) { root = v; return; }
    root = ck;
    u32 hole = quad_slot_sa(quad0_sa, 0u, r);
    u32 off = 1024u + r * 1024u;                            // node h's child r is node 4h + 1 + r: offset -> 4 offset + 1024 (1 + r)
#pragma unroll 1
    while (off < q_end) {
        if (!heap4_level<SHRUNK>(quad0_sa, off, q_end, size, v, ck, r)) break;
        sts64(hole, ck);
        hole = quad_slot_sa(quad0_sa, off, r);
        off = 4u * off + 1024u + r * 1024u;
    }
    sts64(hole, v);
}

template <int D>
__global__ void __launch_bounds__(WARPS * 32, Geo<D>::MINB) knn_fast_kernel(const KnnParams p) {
    typedef Geo<D> G;
    extern __shared__ __align__(128) unsigned char smem[];
    int lane;
    asm volatile("and.b32 %0, %1, 31;" : "=r"(lane) : "r"(threadIdx.x));
    const int wib = threadIdx.x >> 5;
    i64 gblock = blockIdx.x;
    if (p.chunk_blocks > 0) gblock = (gblock / p.chunk_blocks) * p.chunk_stride_blocks + gblock % p.chunk_blocks;
    const i64 ql64 = p.leaf_begin + gblock * WARPS + wib;
    if (ql64 >= p.leaf_end) return;
    const int ql = (int)ql64;
    const int K = p.k;
    const int n_levels = p.n_levels;

    // ---- shared addresses: the warp's base plus compile-time offsets (see Smem)
    typedef Smem<D> S;
    u32 tile_sa;                                                           // the tile buffer = the warp's base
    // (warp-uniform by construction -- a shuffle from lane 0 -- so it lives in a uniform register: addresses are base + constant
    // (+ lane term), and the bulk copy gets its shared operands without a broadcast)
    tile_sa = smem_u32(smem) + (u32)__shfl_sync(FULL, wib, 0) * (u32)p.warp_smem;      // warp-uniform, and the compiler knows
    const u32 queue_sa = tile_sa + S::QUEUE + (u32)lane * 8u;              // this lane's queue slot 0 (slot stride 256)
    const u32 qt_sa = tile_sa + S::QT;                                     // query table
    const u32 qbox_sa = tile_sa + S::QBOX;                                 // own leaf's box, (lo_j, hi_j) pairs
    const u32 stk_sa = tile_sa + S::STK;                                   // children masks of the entries below the top
    const u32 keys_sa = tile_sa + S::KEYS + (u32)lane * 2u;                // this lane's spilled order keys (row stride 64)
    const u32 bar_sa = tile_sa + S::BAR;                                   // the tile buffer's mbarrier
    const u32 quad_sa = tile_sa + S::HEAP + (u32)lane * 16u;               // this lane's quad 0, first half (quad stride 1024)
    const u32 need_sa = tile_sa + S::HEAP + heap_bytes(K) + (u32)lane * 4u;    // spilled need masks, this lane's column (row stride 128)
    const u32 myrow_sa = qt_sa + (u32)lane * (G::RS * 4u);

    const u64 kinf = make_key(INFINITY, 0xffffffffu);
    if (lane == 0) {
        mbar_init_sa(bar_sa, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll 1
    for (int h = 0; 4 * h + 1 < K; h++) {       // a missing child is never "worse"
        sts64x2(quad_sa + (u32)h * 1024u, kinf, (4 * h + 2 < K) ? kinf : 0ull);
        sts64x2(quad_sa + (u32)h * 1024u + 512u, (4 * h + 3 < K) ? kinf : 0ull, (4 * h + 4 < K) ? kinf : 0ull);
    }
#pragma unroll
    for (int j = 0; j < G::RS; j++) stsf(myrow_sa + 4u * j, j == D ? -1.f : 0.f);
    if (lane < G::BS) stsf(qbox_sa + 4u * lane, ((const float*)p.box[0])[(size_t)ql * G::BS + lane]);
    __syncwarp();

    float q[D];
    float worst = -1.f;         // k-th distance so far (-1: padding lane)
    float bound = INFINITY;     // warp max of worst
    u64 root = kinf;            // heap node 0
    int cnt = 0;                // queue fill
    u32 my_id = 0xffffffffu;
    u32 n_qt = 0;               // (query, tile) evaluations
    u32 st_n[3] = {0, 0, 0};            // (KNN_STATS == 2) evaluated tiles needed by 1, 2, 3-4 queries
    u32 st_lvl[4] = {0, 0, 0, 0};       // (KNN_STATS == 2) needy queries of the evaluated tiles, by ancestor level
    u32 st_loaded = 0, st_dense = 0, st_push = 0, st_insert = 0, st_rounds = 0, st_open_iter = 0, st_descents = 0;

    auto flush = [&]() {
        const int maxc = __reduce_max_sync(FULL, cnt);
#pragma unroll 1
        for (int t = 0; t < maxc; t++) {
            const u64 v = lds64(queue_sa + (u32)t * 256u);      // slots past cnt hold stale keys: masked by the test below
            if (t < cnt && root > v) {
                heap4_replace<false>(root, quad_sa, K, v);
                if (KNN_STATS) st_insert++;
            }
            if (KNN_STATS) st_rounds++;
        }
        cnt = 0;
        if (my_id != 0xffffffffu) worst = key_dist(root);
        stsf(myrow_sa + 4u * D, worst);
        __syncwarp();
        bound = warp_max_t<float>(worst);
    };


    // Open a wide node: lane j takes child j (a level-`hc` node), box-to-box test against the warp bound, then for every
    // query in `qmask` the distance from the query to the child's box against the query's own k-th distance.  The child's
    // need mask stays in this lane's `cmask`; `okey` becomes its visiting-order key (box distance | lane).
    u32 okey = 0xffffffffu, cmask = 0;
    auto open_node = [&](int hc, int node, int excl, u32 qmask) -> u32 {
        const int count = (int)p.count[hc];
        const int c = node * FANOUT + lane;
        const bool ok = c < count && c != excl;
        const float4* brow = reinterpret_cast<const float4*>((const float*)p.box[hc] + (size_t)(ok ? c : 0) * G::BS);
        u64 bx[D + 1];
#pragma unroll
        for (int v = 0; v < G::PB4; v++) {
            const float4 f = __ldg(brow + v);
            if (2 * v < D) bx[2 * v] = pack2(f.x, f.y);
            if (2 * v + 1 < D) bx[2 * v + 1] = pack2(f.z, f.w);
        }
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < D; j++) {
            float lo, hi, qlo, qhi;
            unpack2(bx[j], lo, hi);
            ldsf2(qbox_sa + 8u * j, qlo, qhi);
            const float g = fmaxf(fmaxf(lo - qhi, qlo - hi), 0.f);
            acc = __fmaf_rn(g, g, acc);
        }
        const bool pass = ok && acc <= bound;
        u32 mask = 0;
        if (__any_sync(FULL, pass)) {
            u32 m = qmask;
            // distance from query row `qi` to this lane's child box
            auto box_dist = [&](int qi, float& kth) -> float {
                float row[G::RS];
                load_qrow<D>(qt_sa + (u32)(qi * (G::RS * 4)), row);
                float pd = 0.f;
#pragma unroll
                for (int j = 0; j < D; j++) {
                    float a, b;
                    unpack2(sub2(pack2(row[j], row[j]), bx[j]), a, b);      // (q - lo, q - hi)
                    const float g = fmaxf(fmaxf(-a, b), 0.f);
                    pd = __fmaf_rn(g, g, pd);
                }
                kth = row[D];
                return pd;
            };
#if KF_OPEN2
#pragma unroll 1
            while (m) {
                // two needy queries per iteration (independent chains); an odd count tests the last one twice
                const u32 ba = m & (0u - m);
                m ^= ba;
                const u32 bb = m & (0u - m);
                m ^= bb;
                float ka, kb;
                const float pa = box_dist(31 - __clz(ba), ka);
                const float pb = box_dist(31 - __clz(bb ? bb : ba), kb);
                if (pa <= ka) mask |= ba;
                if (pb <= kb) mask |= bb;
                if (KNN_STATS) st_open_iter++;
            }
#else
#pragma unroll 1
            while (m) {
                const u32 lb = m & (0u - m);        // lowest needy query, as a bit
                m ^= lb;
                float kth;
                const float pd = box_dist(31 - __clz(lb), kth);
                if (pd <= kth) mask |= lb;
                if (KNN_STATS) st_open_iter++;
            }
#endif
            mask = pass ? mask : 0u;
        }
        cmask = mask;
        // order only: a few bits of distance are plenty
        const u32 ord = ((u32)(32 - __popc(mask)) << 27) | ((__float_as_uint(acc) >> 5) & 0x07ffffe0u);
        okey = mask != 0 ? (ord | (u32)lane) : 0xffffffffu;
        return __ballot_sync(FULL, mask != 0);
    };

    // ---- traversal state: ancestors of the own leaf bottom-up; a small stack for descents.  The top entry (children still to
    // visit, node, level; per-child need masks in cmask, lane = child) lives in registers.  Entries below it keep their
    // children mask and need masks in shared memory; their node and level follow from the child's (node >> 5, level + 1).
    int anc = 1, sp = 0;
    u32 t_mask = 0;
    int t_node = 0, t_h = 0;
    u32 cand_need = FULL;
    auto next_candidate = [&]() -> int {
        for (;;) {
            if (sp == 0) {
                if (anc > n_levels) return -1;
                if (__any_sync(FULL, cnt > 0)) flush();
                const int node = ql >> (5 * anc), excl = ql >> (5 * (anc - 1));
                t_mask = open_node(anc - 1, node, excl, FULL);
                t_node = node;
                t_h = anc;
                sp = 1;
                anc++;
                continue;
            }
            if (t_h == 1) {
                // leaves: nearest box first (the k-th distances tighten as early as possible); the lane number rides in
                // the low bits of the key, so one reduction yields the minimum and who holds it
                const u32 kmin = __reduce_min_sync(FULL, okey);
                if (kmin != 0xffffffffu) {
                    const u32 bit = kmin & 31u;
                    if ((u32)lane == bit) okey = 0xffffffffu;
                    cand_need = __shfl_sync(FULL, cmask, (int)bit);
                    return t_node * FANOUT + (int)bit;
                }
                t_mask = 0;
            }
            if (t_mask == 0) {
                sp--;
                if (sp > 0) {
                    t_mask = lds32(stk_sa + (u32)(sp - 1) * 4u);
                    cmask = lds32(need_sa + (u32)(sp - 1) * 128u);
                    t_node >>= 5;
                    t_h++;
                    // the entry's order keys: 16 bits of box distance for levels 2 and 3 (index order above)
                    const u32 k16 = t_h <= 3 ? lds16(keys_sa + (u32)(t_h - 2) * 64u) : 0u;
                    okey = ((t_mask >> lane) & 1u) ? ((k16 << 16) | (u32)lane) : 0xffffffffu;
                }
                continue;
            }
            const int bit = (int)(__reduce_min_sync(FULL, okey) & 31u);
            t_mask &= ~(1u << bit);
            if (lane == bit) okey = 0xffffffffu;
            if (t_h <= 3) sts16(keys_sa + (u32)(t_h - 2) * 64u, okey >> 16);
            const u32 need = __shfl_sync(FULL, cmask, bit);
            const int c = t_node * FANOUT + bit;
            // descend: spill the top entry, open the children of node c (a level t_h-1 node) for the queries that need it
            sts32(need_sa + (u32)(sp - 1) * 128u, cmask);
            if (lane == 0) sts32(stk_sa + (u32)(sp - 1) * 4u, t_mask);
            if (KNN_STATS) st_descents++;
            t_mask = open_node(t_h - 2, c, -1, need);
            t_node = c;
            t_h = t_h - 1;
            __syncwarp();
            sp++;
        }
    };

    // Tile number t of this warp completes phase t & 1 of the buffer's barrier.
    auto issue_tile = [&](int leaf) {
        // (an elected lane instead of `lane == 0`: the copy takes uniform-register operands, and a lane test costs a
        // divergence-proof broadcast loop around it)
        const u32 uleaf = (u32)leaf;
        if (elect_one()) {
            mbar_expect_tx_sa(bar_sa, (u32)G::TS);
            tma_load_1d_sa(tile_sa, p.tiles + (size_t)uleaf * (u32)G::TS, (u32)G::TS, bar_sa);
        }
        if (KNN_STATS) st_loaded++;
    };
    auto wait_tile = [&](u32 t) {
        while (!mbar_try_wait_sa(bar_sa, t & 1u)) {}
    };

    // hits of one query of a sparse pass: owner lane `qi` queues every candidate flagged in `hm` (held by lane bit + laneoff)
    auto emit = [&](u32 hm, float val, u32 id, int qi, int laneoff) {
#pragma unroll 1
        while (hm) {
            const int b = __ffs(hm) - 1 + laneoff;
            hm &= hm - 1;
            const float vv = __shfl_sync(FULL, val, b);
            const u32 ii = __shfl_sync(FULL, id, b);
            if (__shfl_sync(FULL, cnt, qi) >= QC) flush();
            if (lane == qi) {
                sts64(queue_sa + (u32)cnt * 256u, make_key(vv, ii));
                cnt++;
                if (KNN_STATS) st_push++;
            }
        }
    };

    // ---- dense: every lane evaluates all 32 candidates of the tile, two per instruction (packed f32x2, broadcast tile reads)
    auto eval_dense = [&]() {
        const u32 ids_sa = tile_sa + (u32)G::IDS;
        if (KNN_STATS) st_dense++;
        n_qt += 32;
        if (__any_sync(FULL, cnt > QC - 4)) flush();        // the sparse path can leave queues fuller than a chunk allows
#pragma unroll 1
        for (int pb0 = 0; pb0 < LEAF / 2; pb0 += 2) {
            float v[4];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                u64 acc = 0ull;
#pragma unroll
                for (int v4 = 0; v4 < G::PB4; v4++) {
                    const float4 f = lds_f4(tile_sa + (u32)((pb0 + c) * G::PB4 + v4) * 16u);
                    if (2 * v4 < D) {
                        const u64 df = sub2(pack2(q[2 * v4], q[2 * v4]), pack2(f.x, f.y));
                        acc = fma2(df, df, acc);
                    }
                    if (2 * v4 + 1 < D) {
                        const u64 df = sub2(pack2(q[2 * v4 + 1], q[2 * v4 + 1]), pack2(f.z, f.w));
                        acc = fma2(df, df, acc);
                    }
                }
                unpack2(acc, v[2 * c], v[2 * c + 1]);
            }
            const float mn = fminf(fminf(v[0], v[1]), fminf(v[2], v[3]));
            if (__any_sync(FULL, mn <= worst)) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (v[e] <= worst) {
                        sts64(queue_sa + (u32)cnt * 256u, make_key(v[e], lds32(ids_sa + (u32)(2 * pb0 + e) * 4u)));
                        cnt++;
                        if (KNN_STATS) st_push++;
                    }
                }
                if (__any_sync(FULL, cnt > QC - 4)) flush();
            }
        }
    };

    // ---- sparse: each half-warp holds the whole tile (lanes l and l + 16: candidates 2 (l & 15) and 2 (l & 15) + 1 as packed
    // pairs -- exactly the tile's pair block -- in `cand`, their ids in id0/id1); per pass the lower half evaluates one needy
    // query, the upper half another (an odd count: the same one again, its hits ignored).
    const u32 hl = (u32)lane & 15u;
    auto eval_sparse = [&](const u64* cand, u32 id0, u32 id1, u32 need) {
        n_qt += __popc(need);
        u32 m = need;
        u32 ba = m & (0u - m);
        m ^= ba;
        u32 bb = m & (0u - m);
        m ^= bb;
        u32 row_sa = qt_sa + (u32)((31 - __clz((lane < 16 || bb == 0u) ? ba : bb)) * (G::RS * 4));
#pragma unroll 1
        for (;; ) {
            float row[G::RS];
            load_qrow<D>(row_sa, row);
            // the next pass's pair of queries and their rows: an integer chain that runs under the loads
            const u32 nba = m & (0u - m);
            m ^= nba;
            const u32 nbb = m & (0u - m);
            m ^= nbb;
            const u32 nrow_sa = qt_sa + (u32)((31 - __clz((lane < 16 || nbb == 0u) ? nba : nbb)) * (G::RS * 4));
            u64 acc = 0ull;
#pragma unroll
            for (int j = 0; j < D; j++) {
                const u64 df = sub2(pack2(row[j], row[j]), cand[j]);
                acc = fma2(df, df, acc);
            }
            float d0, d1;
            unpack2(acc, d0, d1);
            const bool h0 = d0 <= row[D], h1 = d1 <= row[D];
            if (__any_sync(FULL, h0 || h1)) {       // the common pass has no hit at all: one vote, one branch
                u32 b0 = __ballot_sync(FULL, h0), b1 = __ballot_sync(FULL, h1);
                if (bb == 0u) { b0 &= 0xffffu; b1 &= 0xffffu; }
                const int qa = 31 - __clz(ba), qb = 31 - __clz(bb | 1u);
                emit(b0 & 0xffffu, d0, id0, qa, 0);
                emit(b1 & 0xffffu, d1, id1, qa, 0);
                emit(b0 >> 16, d0, id0, qb, 16);
                emit(b1 >> 16, d1, id1, qb, 16);
            }
            if (nba == 0u) break;
            ba = nba;
            bb = nbb;
            row_sa = nrow_sa;
        }
    };

    // ---- own leaf (tile 0)
    issue_tile(ql);
    wait_tile(0u);
    {
        const u32 mine = tile_sa + (u32)(lane >> 1) * (G::PB * 4u) + (u32)(lane & 1) * 4u;
#pragma unroll
        for (int j = 0; j < D; j++) {
            q[j] = ldsf(mine + 8u * j);
            stsf(myrow_sa + 4u * j, q[j]);
        }
        my_id = lds32(tile_sa + (u32)G::IDS + 4u * lane);
        worst = my_id != 0xffffffffu ? INFINITY : -1.f;
        stsf(myrow_sa + 4u * D, worst);
        __syncwarp();
        eval_dense();
    }
    // ---- the other leaves.  One buffer: the traversal finds tile t + 1 while the copy of tile t is in flight, and the copy of
    // tile t + 1 is issued when tile t has been evaluated.  (Issuing it as soon as a sparse tile has been moved to registers,
    // under its passes, was measured slower -- 74.0 against 71.8 ms on C4 -- and needs the copy ordered after the RETURN of
    // those loads, not their issue: with 128-byte pair blocks, d >= 15, their bank conflicts outlast a copy served from L2.)
    const int first = next_candidate();
    if (first >= 0) {
        u32 cur_need = cand_need;
        u32 t = 1;
        int cur_lvl = anc - 2;
        __syncwarp();       // every lane is done reading the own leaf
        issue_tile(first);
#pragma unroll 1
        for (;;) {
            const int nxt = next_candidate();
            const u32 nxt_need = cand_need;
            const int nxt_lvl = anc - 2;        // (diagnostics) the ancestor the candidate was found under: 0 = the parent
            wait_tile(t);
            if (KNN_STATS == 2) {
                st_lvl[cur_lvl < 4 ? cur_lvl : 3] += __popc(cur_need);
                const int nn = __popc(cur_need);
                if (nn == 1) st_n[0]++;
                else if (nn == 2) st_n[1]++;
                else if (nn <= 4) st_n[2]++;
            }
            if (KNN_BRUTE || __popc(cur_need) > KF_SPARSE_MAX) {
                eval_dense();
            } else {
                u64 cand[D + 1];
#pragma unroll
                for (int v4 = 0; v4 < G::PB4; v4++) {
                    const float4 f = lds_f4(tile_sa + hl * (G::PB * 4u) + 16u * v4);
                    if (2 * v4 < D) cand[2 * v4] = pack2(f.x, f.y);
                    if (2 * v4 + 1 < D) cand[2 * v4 + 1] = pack2(f.z, f.w);
                }
                u32 id0, id1;
                lds32x2(tile_sa + (u32)G::IDS + hl * 8u, id0, id1);
                eval_sparse(cand, id0, id1, cur_need);
            }
            if (nxt < 0) break;
            __syncwarp();       // every lane is done with the tile (its loads have been consumed)
            issue_tile(nxt);
            cur_need = nxt_need;
            if (KNN_STATS == 2) cur_lvl = nxt_lvl;
            t++;
        }
    }
    if (__any_sync(FULL, cnt > 0)) flush();

    // ---- heap -> ascending order in place (heapsort over the paired layout; node 0 ends up in queue slot 0), then the outputs
    auto node_sa = [&](int m, int r) -> u32 {       // heap node m of lane r
        if (m == 0) return queue_sa + (u32)(r - lane) * 8u;
        return quad_slot_sa(quad_sa + (u32)(r - lane) * 16u, (u32)((m - 1) >> 2) * 1024u, (u32)((m - 1) & 3));
    };
#pragma unroll 1
    for (int m = K - 1; m >= 1; m--) {
        const u64 v = lds64(node_sa(m, lane));
        sts64(node_sa(m, lane), root);      // the largest of the m + 1 remaining keys
        heap4_replace<true>(root, quad_sa, m, v);
    }
    sts64(queue_sa, root);
    const bool active = my_id != 0xffffffffu;
    const i64 pos_row = ((i64)blockIdx.x * WARPS + wib) * LEAF + lane;      // row of this launch (row_order 1)
    if (active) {
        const i64 row = p.row_order == 0 ? (i64)my_id : pos_row;
        if (p.d2_out != nullptr) {
            float* dd = (float*)p.d2_out + row * K;
#pragma unroll 1
            for (int m = 0; m < K; m++) dd[m] = key_dist(lds64(node_sa(m, lane)));
        }
        if (p.idx_out != nullptr) {
            u32* io = p.idx_out + row * K;
#pragma unroll 1
            for (int m = 0; m < K; m++) io[m] = key_id(lds64(node_sa(m, lane)));
        }
        if (p.logrho_out != nullptr) {
            // a4 (astrolink.py:293-297) fused: same arithmetic and summation order as logrho_kernel (stages.cu)
            float sum = 0.f;
#pragma unroll 1
            for (int m = 0; m < K; m++) sum += key_dist(lds64(node_sa(m, lane)));
            const i64 orow = p.link_by_position ? ql64 * LEAF + lane : (i64)my_id;
            ((float*)p.logrho_out)[orow] = al_raw_logrho<float>(sum, key_dist(lds64(node_sa(K - 1, lane))), K, p.half_d);
        }
    } else if (p.row_order == 1) {
#pragma unroll 1
        for (int m = 0; m < K; m++) {
            if (p.idx_out != nullptr) p.idx_out[pos_row * K + m] = 0xffffffffu;
            if (p.d2_out != nullptr) ((float*)p.d2_out)[pos_row * K + m] = INFINITY;
        }
    }
    if (p.link_out != nullptr && p.link_by_position) {
        // Position-ordered rows: the 32 rows of this leaf are one contiguous block of 32 * link_cols ids, written with
        // full 128-byte warp stores -- the shape NVLink wants when link_out is a peer-mapped buffer.  Padding lanes
        // hold 0xffffffff ids; the consumer's un-permute pass skips them.
        __syncwarp();
        u32* blk = p.link_out + (ql64 * LEAF) * p.link_cols;
        const int total = LEAF * p.link_cols;
#pragma unroll 1
        for (int e = lane; e < total; e += 32) {
            const int r = e / p.link_cols, c = e - r * p.link_cols;
            blk[e] = key_id(lds64(node_sa(c, r)));
        }
    } else if (p.link_out != nullptr && active) {
        u32* lo_ = p.link_out + (i64)my_id * p.link_cols;
#pragma unroll 1
        for (int m = 0; m < p.link_cols; m++) lo_[m] = key_id(lds64(node_sa(m, lane)));
    }
    if (p.stats != nullptr) {
        // [0] (query, candidate) pairs evaluated; the rest are diagnostics (only maintained in KNN_STATS builds)
        if (lane == 0) {
            atomicAdd(p.stats + 0, (unsigned long long)n_qt * LEAF);
            if (KNN_STATS == 2) {
                for (int l = 0; l < 4; l++) atomicAdd(p.stats + 1 + l, (unsigned long long)st_lvl[l]);
                for (int l = 0; l < 3; l++) atomicAdd(p.stats + 5 + l, (unsigned long long)st_n[l]);
            } else if (KNN_STATS) {
                atomicAdd(p.stats + 1, (unsigned long long)st_loaded);          // tiles fetched
                atomicAdd(p.stats + 2, (unsigned long long)st_dense);           // of which evaluated densely
                atomicAdd(p.stats + 3, (unsigned long long)st_descents);        // wide nodes opened below the ancestors
                atomicAdd(p.stats + 6, (unsigned long long)st_rounds);          // flush rounds
                atomicAdd(p.stats + 7, (unsigned long long)st_open_iter);       // iterations of the box-test loop
            }
        }
        if (KNN_STATS == 1) {
            const u32 pushes = __reduce_add_sync(FULL, st_push), inserts = __reduce_add_sync(FULL, st_insert);
            if (lane == 0) {
                atomicAdd(p.stats + 4, (unsigned long long)pushes);
                atomicAdd(p.stats + 5, (unsigned long long)inserts);
            }
        }
    }
}

template <int D>
static int launch(const KnnParams& p0, cudaStream_t st) {
    typedef Geo<D> G;
    KnnParams p = p0;
    AL_REQUIRE(p.tile_stride == G::TS && p.tile_hdr_bytes == 0 && p.tile_ids_off == G::IDS && p.pair_block == G::PB &&
               p.box_stride == G::BS, AL_EINTERNAL, "al_knn: index geometry does not match the float kernel's (d=%d)", D);
    i64 n_q = p.leaf_end - p.leaf_begin;
    if (p.chunk_blocks > 0) {
        const i64 stride_leaves = p.chunk_stride_blocks * WARPS, chunk_leaves = p.chunk_blocks * WARPS;
        const i64 n_chunks = n_q <= 0 ? 0 : (n_q + stride_leaves - 1) / stride_leaves;
        n_q = n_chunks * chunk_leaves;
    }
    if (n_q <= 0) return AL_OK;
    p.warp_smem = warp_smem_bytes<D>(p.k, p.n_levels);
    const size_t smem = (size_t)p.warp_smem * WARPS;
    AL_REQUIRE(smem <= 227 * 1024, AL_EINVAL, "al_knn: k too large for shared memory");
    if (smem > 48 * 1024) AL_CUDA(cudaFuncSetAttribute(knn_fast_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // all of the SM's shared memory for CTAs: occupancy is set by the per-warp footprint above
    AL_CUDA(cudaFuncSetAttribute(knn_fast_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    if (getenv("AL_KNN_DEBUG") != nullptr) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, knn_fast_kernel<D>, WARPS * 32, smem);
        fprintf(stderr, "[al_knn] d=%d k=%d levels=%d: %zu B shared memory per CTA, %d CTAs (%d warps) per SM\n", D, p.k, p.n_levels, smem, nb,
                nb * WARPS);
    }
    knn_fast_kernel<D><<<(unsigned)((n_q + WARPS - 1) / WARPS), WARPS * 32, smem, st>>>(p);
    AL_CHECK_LAUNCH();
    return AL_OK;
}
#endif  // __CUDACC__

}  // namespace kf
