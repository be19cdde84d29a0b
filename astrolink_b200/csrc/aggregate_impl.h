// Parallel aggregation: replaces _aggregate_njit_float32_uint32 /
// _aggregate_njit_float64_uint32 (reference astrolink/astrolink.py:384-604).
//
// The reference walks the sorted edge list once with a union-find
// (Kruskal), concatenating member lists "larger group first" and recording a
// row per merged group.  Everything it outputs is a function of the
// single-linkage dendrogram of the accepted edges under the total order
// (rank = position in the sorted list):
//   ordering    = leaves in depth-first order, larger child first, ties ->
//                 the child holding pair[0] (the denser end of the edge);
//   a group row = a dendrogram node's smaller child when that child has >= 2
//                 points: [start(node), start(node)+|larger|, +|smaller|];
//   prominences = subtree maxima of logRho minus the edge weight.
// This file builds that dendrogram with data-parallel passes only:
//   (A-down) Boruvka rounds on edge ranks: every component picks its
//            lowest-rank outgoing edge; picked edges are exactly the accepted
//            edges, and the rounds give a contraction hierarchy;
//   (A-up)   the hierarchy is expanded level by level: within one component
//            of picked edges ranks grow away from the mutually-picked seed
//            edge, so its edges form chains in the dendrogram; each edge finds
//            the skeleton link it hangs on by a binary-lifting ancestor search
//            and chains are formed by one sort per level;
//   (B)      an Euler tour of the dendrogram ranked by pointer jumping gives
//            subtree sizes and start offsets; rows, prominences and the noise
//            correction follow with sorts, scans and a range-max tree.
// The algorithm is written against an executor X (parallel-for, scan, sort):
// aggregate.cu supplies the CUDA executor used by the product; the serial host
// executor in tests/hostsim exists only so the logic can be tested without a GPU.
#pragma once
#include <math.h>

#include "agg_common.h"

struct uint2_agg { u32 x, y; };
// element of a segmented sum: running sum, running count, segment-head flag
struct SegItem { double s; u32 c; u32 f; };
struct SegOp {
    AGG_HD SegItem operator()(const SegItem& a, const SegItem& b) const {
        if (b.f) return b;
        SegItem r;
        r.s = a.s + b.s;
        r.c = a.c + b.c;
        r.f = a.f;
        return r;
    }
};

namespace agg {

struct TagInit {}; struct TagPickFirst {}; struct TagHook {}; struct TagJump {}; struct TagRoots {}; struct TagLabel {};
struct TagRecord {}; struct TagNodeFirst {}; struct TagKeep {}; struct TagCompact {}; struct TagLift0 {}; struct TagLift {};
struct TagSearch {}; struct TagLink1 {}; struct TagLink2 {}; struct TagLeafChild {}; struct TagTourSucc {}; struct TagRank {};
struct TagSizes {}; struct TagWeights {}; struct TagStarts {}; struct TagOrdering {}; struct TagRows {}; struct TagSeg {};
struct TagNoise {}; struct TagOwn {}; struct TagOut {}; struct TagMisc {};

template <typename T> AGG_HD inline T agg_max(T a, T b) { return a > b ? a : b; }

// leaves under a node reference (leaf: ref < n; internal: n + storage index)
struct NodeSize {
    const u32* size_int;
    u32 n;
    AGG_HD u32 operator()(u32 ref) const { return ref >= n ? size_int[ref - n] : 1u; }
};

// iterative range-max query on a bottom-up max tree of n2 leaves
template <typename T>
struct RangeMax {
    const T* seg;
    i64 n2;
    AGG_HD T operator()(u32 lo, u32 len) const {
        T r = (T)(-INFINITY);
        i64 a = (i64)lo + n2, b = (i64)lo + len + n2;
        while (a < b) {
            if (a & 1) { r = agg_max<T>(r, seg[a]); a++; }
            if (b & 1) { b--; r = agg_max<T>(r, seg[b]); }
            a >>= 1;
            b >>= 1;
        }
        return r;
    }
};

constexpr int MAX_LEVELS = 48;

// Workspace upper bound (bytes) for n points and E edges.
inline size_t workspace_bound(size_t n, size_t E, size_t tsize) {
    size_t a = 256;
    size_t b = 0;
    b += 2 * (E * 8 + a) + 2 * (E * 4 + a);          // ping-pong edge lists
    b += 7 * (n * 4 + a);                            // pk_* arrays, par, child0, child1
    b += 2 * n * 4 + MAX_LEVELS * a + a;             // nodefirst of all levels
    // per-level temporaries: bounded by level 0 (n nodes, E edges) and by the lifting tables
    size_t down = 6 * (n * 4 + a) + 3 * (E * 4 + 4 + a) + (E * 8 + a);
    size_t lift = (n / 2 + 1) * 4 * 34 + 40 * a + 2 * (n * 8 + a) + 2 * (n * 4 + a);
    b += (down > lift ? down : lift);
    // phase B
    b += 4 * (2 * n * 8 + a) + 8 * (n * 4 + a) + 2 * (2 * n * 4 + a);
    b += 4 * n * tsize + 4 * a;                      // range-max tree, ordered logRho
    b += 10 * (n * 4 + a) + 2 * (n * 8 + a) + 4 * (n * tsize + a);   // rows
    b += 64 * a + (64u << 20);                       // slack incl. sort/scan temporaries
    return b;
}

// ---------------------------------------------------------------------------
template <typename X, typename T>
int run(X& x, AggArena& ar, const T* logrho, const u32* sorted_pairs, i64 n, i64 E, bool quirk_col1, u32* ordering_out,
        u32* groups_out, T* prom_out, i64* n_groups_out) {
    const u32 NONE = AGG_NONE;
    const uint2_agg* ep0 = (const uint2_agg*)sorted_pairs;

    // ---------------- persistent arrays
    u32* pk_rank = ar.get<u32>(n);      // rank (sorted position) of each accepted edge, by storage index
    u32* pk_K = ar.get<u32>(n);         // next-level component that contains it
    u32* pk_p0node = ar.get<u32>(n);    // level node holding pair[0]
    u32* pk_picker = ar.get<u32>(n);    // level node that picked it (NONE: picked by both ends = seed)
    u32* par = ar.get<u32>(n);          // dendrogram parent (storage index) or NONE
    u32* child0 = ar.get<u32>(n);       // child on pair[0]'s side: leaf v, or n + storage index
    u32* child1 = ar.get<u32>(n);
    u32* nodefirst_all = ar.get<u32>(2 * n + MAX_LEVELS * 64 + 64);
    uint2_agg* epA = ar.get<uint2_agg>(E);
    uint2_agg* epB = ar.get<uint2_agg>(E);
    u32* erA = ar.get<u32>(E);
    u32* erB = ar.get<u32>(E);
    // device-side totals of the prefix sums: the host reads them in batches (one read per Boruvka round, three in
    // phase B) instead of once per scan
    u32* tot = ar.get<u32>(16);
    if (!ar.ok) return -2;
    x.fill_u32(tot, 0, 16);
    u32* err = tot + 8;                  // consistency checks evaluated on the device, read with the last batch
    const i64 nodefirst_cap = 2 * n + MAX_LEVELS * 64 + 64;

    i64 lvl_off[MAX_LEVELS + 1], lvl_nodes[MAX_LEVELS + 1], lvl_node_off[MAX_LEVELS + 1];

    x.phase("A-down");
    // ================= phase A-down: Boruvka rounds on ranks =================
    i64 N = n, m = E, M = 0, node_off = 0;
    int L = 0;
    const uint2_agg* ep = ep0;
    const u32* er = nullptr;      // nullptr: rank == list index (level 0)
    uint2_agg* ep_next = epA;
    u32* er_next = erA;
    while (m > 0) {
        if (L >= MAX_LEVELS) return -3;
        x.phase_level("A-down", L, m, N);
        const size_t mk = ar.mark();
        u32* first = ar.get<u32>(N);
        u32* hook = ar.get<u32>(N);
        u32* flag = ar.get<u32>(N + 1);
        u32* newid = ar.get<u32>(N + 1);
        u32* pickbits = ar.get<u32>(m / 32 + 2);     // one bit per edge of the list: picked by a component this round
        u32* changed = ar.get<u32>(8);
        if (!ar.ok) return -2;
        if (node_off + N > nodefirst_cap) return -2;
        u32* nodefirst = nodefirst_all + node_off;
        x.fill_u32(first, NONE, N);
        x.fill_u32(pickbits, 0, m / 32 + 2);
        {
            const uint2_agg* e_ = ep;
            x.template launch<TagPickFirst>(m, AGG_LAMBDA(i64 i) {
                // the list is in rank order, so after the first few edges of a component almost every
                // later edge loses: a plain load (possibly stale = larger, hence safe) filters the atomics
                uint2_agg e = e_[i];
                if ((u32)i < first[e.x]) agg_atomic_min(first + e.x, (u32)i);
                if ((u32)i < first[e.y]) agg_atomic_min(first + e.y, (u32)i);
            });
            x.template launch<TagHook>(N, AGG_LAMBDA(i64 s) {
                u32 i = first[s];
                if (i == NONE) { hook[s] = (u32)s; return; }
                uint2_agg e = e_[i];
                u32 other = e.x == (u32)s ? e.y : e.x;
                bool mutual = first[other] == i;
                hook[s] = (mutual && (u32)s < other) ? (u32)s : other;
                agg_atomic_or(pickbits + (i >> 5), 1u << (i & 31u));
            });
        }
        // Pointer jumping to the component roots.  Every vertex chases its pointer up to 24 hops and stores where it got
        // to; concurrent stores only ever replace a pointer by an ancestor, so stale reads are safe.  One pass divides
        // every depth by at least 24, so 7 passes reach the roots from any depth below 24^7 > 2^32: a fixed number of
        // launches, and no host round trip to find out when to stop -- a pass returns at once when the one before it
        // found every vertex at its root.
        x.fill_u32(changed, 0, 8);
        for (int it = 0; it < 7; it++) {
            const u32* prev = it > 0 ? changed + (it - 1) : nullptr;
            u32* cur = changed + it;
            x.template launch_if<TagJump>(prev, N, AGG_LAMBDA(i64 s) {
                u32 h = agg_load_relaxed(hook + s);
                const u32 h0 = h;
                bool at_root = false;
                for (int hop = 0; hop < 24; hop++) {
                    const u32 hh = agg_load_relaxed(hook + h);
                    if (hh == h) { at_root = true; break; }
                    h = hh;
                }
                if (h != h0) hook[s] = h;
                if (!at_root) *cur = 1u;
            });
        }
        x.template launch<TagRoots>(N, AGG_LAMBDA(i64 s) {
            flag[s] = hook[s] == (u32)s ? 1u : 0u;
            if (s == 0 && changed[6] != 0u) err[0] = 1u;        // the jumping did not converge (cannot happen below 24^7 vertices)
        });
        x.exclusive_sum_u32(flag, newid, N, ar, tot + 0);
        // label[s] kept in flag[]
        x.template launch<TagLabel>(N, AGG_LAMBDA(i64 s) { flag[s] = newid[hook[s]]; });
        const u32* label = flag;
        // every picked edge merges two components, so M + cnt = n - Nn <= n: the records below stay inside pk_*[n]
        {
            const uint2_agg* e_ = ep;
            const u32* r_ = er;
            const i64 M_ = M;
            // Two stable selections in ONE pass over the edge list (no flag arrays, no prefix sums of them):
            // (1) the picked edges, in list (= rank) order, become this level's records; the vertices that picked an edge
            //     learn where it went (nodefirst);
            // (2) the survivors -- not picked, endpoints in different new components -- relabelled, form the next level's list.
            x.fill_u32(nodefirst, NONE, N);
            uint2_agg* eo = ep_next;
            u32* ro = er_next;
            x.template select2<TagRecord>(m, AGG_LAMBDA(i64 i) { return ((pickbits[i >> 5] >> (i & 31)) & 1u) != 0u; }, AGG_LAMBDA(i64 i, i64 pos) {
                const uint2_agg e = e_[i];
                const i64 c = M_ + pos;
                pk_rank[c] = r_ ? r_[i] : (u32)i;
                pk_K[c] = label[e.x];
                pk_p0node[c] = e.x;
                const bool fx = first[e.x] == (u32)i, fy = first[e.y] == (u32)i;
                pk_picker[c] = (fx && fy) ? NONE : (fx ? e.x : e.y);
                if (fx) nodefirst[e.x] = (u32)c;
                if (fy) nodefirst[e.y] = (u32)c;
            }, AGG_LAMBDA(i64 i) {
                const uint2_agg e = e_[i];
                return label[e.x] != label[e.y];
            }, AGG_LAMBDA(i64 i, i64 pos) {
                const uint2_agg e = e_[i];
                uint2_agg ne;
                ne.x = label[e.x];
                ne.y = label[e.y];
                eo[pos] = ne;
                ro[pos] = r_ ? r_[i] : (u32)i;
            }, tot + 1, ar);
            // the round's only host read: components, picked edges and surviving edges
            u32 h3[3];
            x.read_u32s(tot, h3, 3);
            const i64 Nn = h3[0], cnt = h3[1], m_next = h3[2];
            if (M + cnt > n || Nn + cnt != N) return -4;
            lvl_off[L] = M;
            lvl_nodes[L] = N;
            lvl_node_off[L] = node_off;
            M += cnt;
            node_off += N;
            L++;
            ep = ep_next;
            er = er_next;
            if (ep_next == epA) { ep_next = epB; er_next = erB; } else { ep_next = epA; er_next = erA; }
            m = m_next;
            N = Nn;
        }
        x.sync();
        ar.reset(mk);
    }
    lvl_off[L] = M;
    lvl_nodes[L] = N;          // number of connected components
    lvl_node_off[L] = node_off;
    const i64 n_comp = N;
    if (M != n - n_comp) return -5;
    if (M == 0) return -6;

    x.phase("A-up");
    // ================= phase A-up: expand the hierarchy into the dendrogram =================
    x.fill_u32(par, NONE, M);
    x.fill_u32(child0, NONE, M);
    x.fill_u32(child1, NONE, M);
    for (int l = L - 1; l >= 0; l--) {
        const size_t mk = ar.mark();
        const i64 lo = lvl_off[l], hi = lvl_off[l + 1], cnt = hi - lo;
        const i64 suffix = M - hi;
        const i64 Nup = lvl_nodes[l + 1];
        const u32* nf_up = (l + 1 < L) ? nodefirst_all + lvl_node_off[l + 1] : nullptr;   // first edge of level-(l+1) nodes
        int J = 0;
        u32* up = nullptr;
        if (suffix > 0) {
            J = agg_log2_ceil((u64)suffix) + 1;
            up = ar.get<u32>((size_t)J * suffix);
            if (!ar.ok) return -2;
            x.template launch<TagLift0>(suffix, AGG_LAMBDA(i64 i) { up[i] = par[hi + i]; });
            for (int j = 1; j < J; j++) {
                u32* uj = up + (size_t)j * suffix;
                const u32* ujm = up + (size_t)(j - 1) * suffix;
                x.template launch<TagLift>(suffix, AGG_LAMBDA(i64 i) {
                    u32 a = ujm[i];
                    uj[i] = a == NONE ? NONE : ujm[a - hi];
                });
            }
        }
        // accepted edges of one level are stored in rank order, so a STABLE sort by link alone orders every
        // chain by rank: 32-bit keys over the bits the link ids need
        u32* key_in = ar.get<u32>(cnt);
        u32* key_out = ar.get<u32>(cnt);
        u32* val_in = ar.get<u32>(cnt);
        u32* val_out = ar.get<u32>(cnt);
        if (!ar.ok) return -2;
        x.template launch<TagSearch>(cnt, AGG_LAMBDA(i64 i) {
            const i64 c = lo + i;
            const u32 K = pk_K[c], r = pk_rank[c];
            u32 link = K;
            if (pk_picker[c] != NONE && nf_up != nullptr) {
                u32 a0 = nf_up[K];
                if (a0 != NONE && pk_rank[a0] < r) {
                    u32 cur = a0;
                    for (int j = J - 1; j >= 0; j--) {
                        u32 nx = up[(size_t)j * suffix + (cur - hi)];
                        if (nx != NONE && pk_rank[nx] < r) cur = nx;
                    }
                    link = (u32)(Nup + (cur - hi));
                }
            }
            key_in[i] = link;
            val_in[i] = (u32)c;
        });
        x.sort_pairs_u32_u32(key_in, key_out, val_in, val_out, cnt, agg_log2_ceil((u64)(Nup + suffix + 1)), ar);
        // pass 1: chain links, and the top of every chain takes over the link's old parent
        x.template launch<TagLink1>(cnt, AGG_LAMBDA(i64 i) {
            const u32 c = val_out[i];
            const u32 link = key_out[i];
            const bool nextsame = (i + 1 < cnt) && key_out[i + 1] == link;
            if (nextsame) {
                const u32 P = val_out[i + 1];
                par[c] = P;
                if (pk_picker[P] != pk_p0node[P]) child0[P] = (u32)n + c; else child1[P] = (u32)n + c;
            } else if ((i64)link < Nup) {
                const u32 P = nf_up ? nf_up[link] : NONE;
                par[c] = P;
                if (P != NONE) {
                    if (link == pk_p0node[P]) child0[P] = (u32)n + c; else child1[P] = (u32)n + c;
                }
            } else {
                const u32 f = (u32)(hi + ((i64)link - Nup));
                const u32 P = par[f];
                par[c] = P;
                if (P != NONE) {
                    if (child0[P] == (u32)n + f) child0[P] = (u32)n + c; else child1[P] = (u32)n + c;
                }
            }
        });
        // pass 2: the bottom of a chain on an edge link adopts that edge
        x.template launch<TagLink2>(cnt, AGG_LAMBDA(i64 i) {
            const u32 link = key_out[i];
            if ((i64)link < Nup) return;
            if (i > 0 && key_out[i - 1] == link) return;
            const u32 c = val_out[i];
            const u32 f = (u32)(hi + ((i64)link - Nup));
            par[f] = c;
            if (pk_picker[c] != pk_p0node[c]) child0[c] = (u32)n + f; else child1[c] = (u32)n + f;
        });
        x.sync();
        ar.reset(mk);
    }
    const u32* nf0 = nodefirst_all + lvl_node_off[0];    // first (parent) edge of every point
    x.template launch<TagLeafChild>(n, AGG_LAMBDA(i64 v) {
        const u32 P = nf0[v];
        if (P == NONE) return;
        if ((u32)v == pk_p0node[P]) child0[P] = (u32)v; else child1[P] = (u32)v;
    });

    x.phase("B-tour");
    // ================= phase B: sizes, starts, ordering, rows =================
    // ---- Euler tour over internal nodes: event 2c = enter(c), 2c+1 = exit(c)
    const i64 TL = 2 * M;
    u32* roots = ar.get<u32>(n_comp + 1);
    u32* rflag = ar.get<u32>(M + 1);
    u32* rpos = ar.get<u32>(M + 1);
    uint2_agg* la = ar.get<uint2_agg>(TL);   // (next, rank) ping
    uint2_agg* lb = ar.get<uint2_agg>(TL);   // pong
    if (!ar.ok) return -2;
    x.template launch<TagMisc>(M, AGG_LAMBDA(i64 c) { rflag[c] = par[c] == NONE ? 1u : 0u; });
    x.exclusive_sum_u32(rflag, rpos, M, ar, tot + 3);      // number of roots: checked against n_comp with the next read
    const i64 nroots = n_comp;
    x.template launch<TagMisc>(M, AGG_LAMBDA(i64 c) { if (rflag[c]) roots[rpos[c]] = (u32)c; });
    x.template launch<TagTourSucc>(M, AGG_LAMBDA(i64 c) {
        const u32 k0 = child0[c], k1 = child1[c];
        // enter(c)
        u32 nx;
        if (k0 >= (u32)n) nx = 2 * (k0 - (u32)n);
        else if (k1 >= (u32)n) nx = 2 * (k1 - (u32)n);
        else nx = 2 * (u32)c + 1;
        uint2_agg a; a.x = nx; a.y = 1;
        la[2 * c] = a;
        // exit(c)
        const u32 P = par[c];
        u32 ny;
        if (P == NONE) {
            const u32 ri = rpos[c];
            ny = (ri + 1 < (u32)nroots) ? 2 * roots[ri + 1] : NONE;
        } else if (child0[P] == (u32)n + (u32)c && child1[P] >= (u32)n) ny = 2 * (child1[P] - (u32)n);
        else ny = 2 * P + 1;
        uint2_agg b; b.x = ny; b.y = ny == NONE ? 0u : 1u;
        la[2 * c + 1] = b;
    });
    {
        // List ranking (hops to the end of the tour) in O(TL) work: a pseudo-random 1/32 of the elements
        // (by a hash of the index: exits have odd indices, chains have regular strides, so a plain modulus
        // would leave long ruler-free runs) and the head are rulers; each ruler walks to the next ruler
        // labelling the elements it passes (owner, offset); only the short list of rulers is ranked by
        // pointer jumping; the ranks of all elements follow.
        u32* headp = ar.get<u32>(1);
        u32* rflag2 = ar.get<u32>(TL + 1);
        u32* rid = ar.get<u32>(TL + 1);
        if (!ar.ok) return -2;
        x.template launch<TagMisc>(1, AGG_LAMBDA(i64) { headp[0] = 2u * roots[0]; });
        x.template launch<TagRank>(TL, AGG_LAMBDA(i64 e) {
            u32 h = (u32)e * 2654435761u;
            h ^= h >> 15;
            rflag2[e] = ((h & 31u) == 0u || (u32)e == headp[0]) ? 1u : 0u;
        });
        x.exclusive_sum_u32(rflag2, rid, TL, ar, tot + 4);
        u32 h2[2];
        x.read_u32s(tot + 3, h2, 2);
        if ((i64)h2[0] != n_comp) return -7;
        const i64 nR = h2[1];
        u32* relem = ar.get<u32>(nR + 1);
        uint2_agg* ra = ar.get<uint2_agg>(nR + 1);     // (next ruler id, elements up to the next ruler) ping
        uint2_agg* rb = ar.get<uint2_agg>(nR + 1);     // pong
        if (!ar.ok) return -2;
        x.template launch<TagRank>(TL, AGG_LAMBDA(i64 e) { if (rflag2[e]) relem[rid[e]] = (u32)e; });
        x.template launch<TagRank>(nR, AGG_LAMBDA(i64 r) {
            u32 e = relem[r];
            u32 off = 0;
            for (;;) {
                uint2_agg o; o.x = (u32)r; o.y = off;
                lb[e] = o;
                const u32 nx = la[e].x;
                off++;
                if (nx == NONE) { uint2_agg z; z.x = NONE; z.y = off; ra[r] = z; return; }
                // (the ruler test is recomputed from the index instead of read from rflag2: one random sector less per
                // element; the head, the only ruler the hash does not decide, is nobody's successor)
                u32 hn = nx * 2654435761u;
                hn ^= hn >> 15;
                if ((hn & 31u) == 0u) { uint2_agg z; z.x = rid[nx]; z.y = off; ra[r] = z; return; }
                e = nx;
            }
        });
        // suffix sums of the segment lengths along the ruler list
        {
            const int rounds = agg_log2_ceil((u64)nR) + 1;
            uint2_agg* src = ra;
            uint2_agg* dst = rb;
            for (int r = 0; r < rounds; r++) {
                const uint2_agg* s_ = src;
                uint2_agg* d_ = dst;
                x.template launch<TagRank>(nR, AGG_LAMBDA(i64 i) {
                    uint2_agg a = s_[i];
                    if (a.x != NONE) {
                        const uint2_agg b2 = s_[a.x];
                        a.y += b2.y;
                        a.x = b2.x;
                    }
                    d_[i] = a;
                });
                uint2_agg* t = src; src = dst; dst = t;
            }
            ra = src;      // ra[rid].y = elements from ruler rid (inclusive) to the end of the tour
        }
        const uint2_agg* rs = ra;
        x.template launch<TagRank>(TL, AGG_LAMBDA(i64 e) {
            const uint2_agg o = lb[e];
            la[e].y = rs[o.x].y - 1u - o.y;    // hops from e to the last element
        });
    }
    x.phase("B-starts");
    // tour position of event e: TL-1 - rank(e); subtree(c) holds (exit-enter+1)/2 internal nodes
    u32* size_int = ar.get<u32>(M);       // leaves under internal node c
    u32* wts = ar.get<u32>(TL);           // +w at enter, -w at exit, in tour order
    u32* wscan = ar.get<u32>(TL);
    u32* start_int = ar.get<u32>(M);
    if (!ar.ok) return -2;
    const uint2_agg* rk = la;
    x.template launch<TagSizes>(M, AGG_LAMBDA(i64 c) {
        const u32 pe = (u32)(TL - 1) - rk[2 * c].y, px = (u32)(TL - 1) - rk[2 * c + 1].y;
        size_int[c] = (px - pe + 1) / 2 + 1;
    });
    // component offsets (only when the link graph is disconnected; astrolink.py:442-461)
    u32* comp_off = ar.get<u32>(n_comp + 1);
    u32* comp_order = ar.get<u32>(n_comp + 1);       // roots in final order (descending (size, seed rank))
    if (!ar.ok) return -2;
    x.fill_u32(comp_off, 0, n_comp + 1);
    const NodeSize node_size{size_int, (u32)n};
    u32* seedrank = ar.get<u32>(n_comp);
    u32* renter = ar.get<u32>(n_comp);
    u64* ck_in = ar.get<u64>(n_comp);
    u64* ck_out = ar.get<u64>(n_comp);
    u32* cv_in = ar.get<u32>(n_comp);
    u32* csz = ar.get<u32>(n_comp + 1);
    u32* cex = ar.get<u32>(n_comp + 1);
    if (!ar.ok) return -2;
    // pass 0: starts with zero component offsets; pass 1 (disconnected graphs only): with the final offsets
    for (int pass = 0; pass < (n_comp > 1 ? 2 : 1); pass++) {
        if (pass == 1) {
            // seed rank of every component: the seed edge (two leaf children) whose local start is 0
            x.template launch<TagMisc>(n_comp, AGG_LAMBDA(i64 g) { renter[g] = (u32)(TL - 1) - rk[2 * roots[g]].y; });
            x.template launch<TagMisc>(M, AGG_LAMBDA(i64 c) {
                if (child0[c] >= (u32)n || child1[c] >= (u32)n || start_int[c] != 0) return;
                const u32 pe = (u32)(TL - 1) - rk[2 * c].y;
                i64 a = 0, b = n_comp;     // last root with renter <= pe
                while (b - a > 1) { i64 mid = (a + b) / 2; if (renter[mid] <= pe) a = mid; else b = mid; }
                seedrank[a] = pk_rank[c];
            });
            // components by (size, seed rank) descending == ascending complemented key
            x.template launch<TagMisc>(n_comp, AGG_LAMBDA(i64 g) {
                const u64 key = ((u64)size_int[roots[g]] << 32) | seedrank[g];
                ck_in[g] = ~key;
                cv_in[g] = (u32)g;
            });
            x.sort_pairs_u64_u32(ck_in, ck_out, cv_in, comp_order, n_comp, 64, ar);
            x.template launch<TagMisc>(n_comp, AGG_LAMBDA(i64 j) { csz[j] = size_int[roots[comp_order[j]]]; });
            x.exclusive_sum_u32(csz, cex, n_comp, ar, tot + 7);
            x.template launch<TagMisc>(n_comp, AGG_LAMBDA(i64 j) { comp_off[comp_order[j]] = cex[j]; });
        }
        x.template launch<TagWeights>(M, AGG_LAMBDA(i64 c) {
            const u32 P = par[c];
            u32 w;
            if (P == NONE) w = comp_off[rpos[c]];
            else {
                const u32 k0 = child0[P], k1 = child1[P];
                const u32 s0 = node_size(k0), s1 = node_size(k1);
                const u32 firstc = s0 >= s1 ? k0 : k1;
                w = (firstc == (u32)n + (u32)c) ? 0u : (s0 >= s1 ? s0 : s1);
            }
            wts[(u32)(TL - 1) - rk[2 * c].y] = w;
            wts[(u32)(TL - 1) - rk[2 * c + 1].y] = 0u - w;
        });
        x.inclusive_sum_u32(wts, wscan, TL, ar);
        x.template launch<TagStarts>(M, AGG_LAMBDA(i64 c) { start_int[c] = wscan[(u32)(TL - 1) - rk[2 * c].y]; });
    }
    // A leaf's place in the ordering follows from its parent node: the larger child (tie: child 0) starts where the node
    // starts, the other one right after it.  One pass over the internal nodes (sequential reads of their children and
    // starts) instead of one over the leaves (six random gathers each).
    x.template launch<TagOrdering>(M, AGG_LAMBDA(i64 c) {
        const u32 k0 = child0[c], k1 = child1[c];
        if (k0 >= (u32)n && k1 >= (u32)n) return;
        const u32 s0 = node_size(k0), s1 = node_size(k1);
        const u32 st = start_int[c];
        if (k0 < (u32)n) ordering_out[st + (s0 >= s1 ? 0u : s1)] = k0;
        if (k1 < (u32)n) ordering_out[st + (s0 >= s1 ? s0 : 0u)] = k1;
    });

    x.phase("B-rows");
    // ---- range-max tree over logRho in final order
    i64 n2 = 1;
    while (n2 < n) n2 *= 2;
    T* seg = ar.get<T>(2 * n2);
    if (!ar.ok) return -2;
    const T NEG = (T)(-INFINITY);
    x.template launch<TagMisc>(n2, AGG_LAMBDA(i64 i) { seg[n2 + i] = i < n ? logrho[ordering_out[i]] : NEG; });
    for (i64 w = n2 / 2; w >= 1; w /= 2)
        x.template launch<TagSeg>(w, AGG_LAMBDA(i64 i) { seg[w + i] = agg_max<T>(seg[2 * (w + i)], seg[2 * (w + i) + 1]); });
    const RangeMax<T> range_max{seg, n2};

    // ---- group rows: one per dendrogram node whose smaller child has >= 2 points,
    //      plus one per extra connected component (fix-up rows)
    u32* rowflag = ar.get<u32>(M + 1);
    u32* rowpos = ar.get<u32>(M + 1);
    if (!ar.ok) return -2;
    x.template launch<TagMisc>(M, AGG_LAMBDA(i64 c) {
        const u32 s0 = node_size(child0[c]), s1 = node_size(child1[c]);
        rowflag[c] = (s0 < s1 ? s0 : s1) >= 2 ? 1u : 0u;
    });
    x.exclusive_sum_u32(rowflag, rowpos, M, ar, tot + 5);
    u32 hR0 = 0;
    x.read_u32s(tot + 5, &hR0, 1);
    const i64 R0 = hR0;
    const i64 R = R0 + (n_comp - 1);
    u32* r_geq = ar.get<u32>(R + 1);
    u32* r_leq = ar.get<u32>(R + 1);
    u32* r_size = ar.get<u32>(R + 1);
    T* r_pg = ar.get<T>(R + 1);
    T* r_pl = ar.get<T>(R + 1);
    u64* nk_in = ar.get<u64>(R + 1);
    u64* nk_out = ar.get<u64>(R + 1);
    u32* nv_in = ar.get<u32>(R + 1);
    u32* nv_out = ar.get<u32>(R + 1);
    u32* row_of_start = ar.get<u32>(n);
    SegItem* items = ar.get<SegItem>(R + 1);
    double* seg_noise = ar.get<double>(R + 1);
    u32* seg_m = ar.get<u32>(R + 1);
    if (!ar.ok) return -2;
    x.fill_u32(row_of_start, NONE, n);
    x.template launch<TagRows>(M, AGG_LAMBDA(i64 c) {
        if (!rowflag[c]) return;
        const u32 r = rowpos[c];
        const u32 k0 = child0[c], k1 = child1[c];
        const u32 s0 = node_size(k0), s1 = node_size(k1);
        const bool first0 = s0 >= s1;
        const u32 sf = first0 ? s0 : s1, ss = first0 ? s1 : s0;
        const u32 st = start_int[c];
        // edge weight = logRho of the sparser end = pair[1] of the accepted edge
        const T sad = logrho[sorted_pairs[2 * (i64)pk_rank[c] + 1]];
        r_geq[r] = st;
        r_leq[r] = st + sf;
        r_size[r] = ss;
        r_pg[r] = range_max(st, sf) - sad;
        r_pl[r] = range_max(st + sf, ss) - sad;
        nk_in[r] = ((u64)st << 32) | pk_rank[c];
        nv_in[r] = r;
        row_of_start[st + sf] = r;
    });
    if (n_comp > 1) {
        // fix-up rows (astrolink.py:450-460): components after the first, in final order
        x.template launch<TagRows>(n_comp - 1, AGG_LAMBDA(i64 j) {
            const u32 r = (u32)(R0 + j);
            const u32 g = comp_order[j + 1];
            const u32 off = comp_off[g], sz = size_int[roots[g]];
            const u32 sz0 = size_int[roots[comp_order[0]]];
            r_geq[r] = 0;
            r_leq[r] = off;
            r_size[r] = sz;
            r_pg[r] = range_max(0, sz0);
            r_pl[r] = range_max(off, sz);
            nk_in[r] = ((u64)0 << 32) | (u32)(E + j);
            nv_in[r] = r;
            row_of_start[off] = r;
        });
    }
    if (R > 0) {
        // ---- noise correction (astrolink.py:471-485): children of a group in merge order
        x.sort_pairs_u64_u32(nk_in, nk_out, nv_in, nv_out, R, 64, ar);
        // exclusive per-group prefix (noise = sum of pl^2 of the earlier children, t = their count):
        // a segmented scan over the rows in (group, merge time) order
        x.template launch<TagMisc>(R, AGG_LAMBDA(i64 i) {
            const bool head = i == 0 || (u32)(nk_out[i - 1] >> 32) != (u32)(nk_out[i] >> 32);
            SegItem it;
            if (head) { it.s = 0.0; it.c = 0; it.f = 1; }
            else {
                const T pl = r_pl[nv_out[i - 1]];
                it.s = (double)(T)(pl * pl);
                it.c = 1;
                it.f = 0;
            }
            items[i] = it;
        });
        x.segmented_inclusive_scan(items, R, ar);
        x.template launch<TagNoise>(R, AGG_LAMBDA(i64 i) {
            const u32 r = nv_out[i];
            const SegItem it = items[i];
            if (it.c > 0) r_pg[r] = (T)((double)r_pg[r] - sqrt(it.s / (double)it.c));
            const bool last = i + 1 == R || (u32)(nk_out[i + 1] >> 32) != (u32)(nk_out[i] >> 32);
            if (last) {
                const T pl = r_pl[r];
                seg_noise[i] = it.s + (double)(T)(pl * pl);
                seg_m[i] = it.c + 1;
            }
        });
        x.template launch<TagOwn>(R, AGG_LAMBDA(i64 i) {
            const u32 key = (u32)(nk_out[i] >> 32);
            if (i + 1 < R && (u32)(nk_out[i + 1] >> 32) == key) return;    // totals live at the last row of a group
            const u32 own = row_of_start[key];      // the row in which this group is the smaller side
            if (own == NONE) return;                // the root group has no row
            const double corr = sqrt(seg_noise[i] / (double)seg_m[i]);
            if (quirk_col1) r_pl[own] = (T)((double)r_pl[own] - corr);
            else r_pg[own] = (T)((double)r_pg[own] - corr);
        });
    }
    // ---- output: rows with size > 2, ascending leq_start
    u32* oflag = ar.get<u32>(R + 1);
    u32* opos = ar.get<u32>(R + 1);
    u64* ok_in = ar.get<u64>(R + 1);
    u64* ok_out = ar.get<u64>(R + 1);
    u32* ov_in = ar.get<u32>(R + 1);
    u32* ov_out = ar.get<u32>(R + 1);
    if (!ar.ok) return -2;
    i64 G = 0;
    if (R > 0) {
        x.template launch<TagMisc>(R, AGG_LAMBDA(i64 r) { oflag[r] = r_size[r] > 2 ? 1u : 0u; });
        x.exclusive_sum_u32(oflag, opos, R, ar, tot + 6);
        u32 hG[3];
        x.read_u32s(tot + 6, hG, 3);       // [6] groups, [7] (scratch of the component offsets), [8] device-side checks
        if (hG[2] != 0u) return -8;
        G = hG[0];
        if (G > 0) {
            x.template launch<TagMisc>(R, AGG_LAMBDA(i64 r) {
                if (!oflag[r]) return;
                ok_in[opos[r]] = (u64)r_leq[r];
                ov_in[opos[r]] = (u32)r;
            });
            x.sort_pairs_u64_u32(ok_in, ok_out, ov_in, ov_out, G, 32, ar);
            x.template launch<TagOut>(G, AGG_LAMBDA(i64 i) {
                const u32 r = ov_out[i];
                groups_out[3 * i] = r_geq[r];
                groups_out[3 * i + 1] = r_leq[r];
                groups_out[3 * i + 2] = r_leq[r] + r_size[r];
                prom_out[2 * i] = r_pg[r];
                prom_out[2 * i + 1] = r_pl[r];
            });
        }
    }
    x.phase("end");
    x.sync();
    *n_groups_out = G;
    return 0;
}

}  // namespace agg
