// Layout of the spatial index shared by tree.cu (build) and knn.cu (search).
//
// Replaces pykdtree's KDTree(P_transform) (reference astrolink.py:270).  The
// index is a permutation of the points into LEAF-point leaves (one warp's worth
// of queries / one TMA tile of candidates) produced by a balanced axis-aligned
// split tree, plus tight bounding boxes of every aligned block of 32^h leaves
// (a 32-ary box hierarchy a warp tests with one lane per child).
//
// Device memory layout of the opaque `index` buffer:
//   [IndexHeader, 256 B]
//   [perm: u32[NP]]                     index position -> original point (0xffffffff = padding)
//   [tiles: NLf x tile_stride bytes]    per leaf: 16 pair blocks of PB T | ids u32[32]
//   [boxes level 0..G: count_h rows of box_stride T: lo_0 hi_0 lo_1 hi_1 ... (one row = one node, 16-byte aligned,
//    so the lane that tests child c reads its whole box with 128-bit loads and (lo_j, hi_j) is one packed f32x2 operand)]
#pragma once
#include "common.cuh"

constexpr int LEAF = 32;
constexpr int FANOUT = 32;
constexpr int MAX_WIDE_LEVELS = 8;

struct IndexHeader {
    u32 magic;
    int dtype;
    int d;
    int n_levels;          // G: wide levels above the leaves (0 when a single leaf)
    i64 n;
    i64 n_leaves;          // NLf
    i64 count[MAX_WIDE_LEVELS];      // nodes at wide level h (count[0] == n_leaves)
    size_t box_off[MAX_WIDE_LEVELS]; // byte offset of level h boxes
    size_t perm_off;
    size_t tiles_off;
    size_t total_bytes;
    int tile_stride;       // bytes, multiple of 128
    int tile_hdr_bytes;    // bytes before the pair blocks (0)
    int tile_ids_off;      // byte offset of ids within a tile
    int pair_block;        // PB: elements of T per pair block (>= 2d)
    int box_stride;        // BS: elements of T per box row (>= 2d, a multiple of 16 bytes)
};

constexpr u32 INDEX_MAGIC = 0xA5710B20u;

static inline int elem_bytes(int dtype) { return dtype == AL_F64 ? 8 : 4; }

static inline int pair_block_elems(int dtype, int d) {
    int per16 = 16 / elem_bytes(dtype);   // elements per 16 bytes
    return (2 * d + per16 - 1) / per16 * per16;
}

static inline void index_layout(int dtype, i64 n, int d, IndexHeader* h) {
    memset(h, 0, sizeof(*h));
    h->magic = INDEX_MAGIC;
    h->dtype = dtype;
    h->d = d;
    h->n = n;
    h->n_leaves = (n + LEAF - 1) / LEAF;
    int eb = elem_bytes(dtype);
    h->pair_block = pair_block_elems(dtype, d);
    h->box_stride = pair_block_elems(dtype, d);
    h->tile_hdr_bytes = 0;      // (a leaf's own box lives in the level-0 box rows; the tile is coordinates + ids only)
    h->tile_ids_off = h->tile_hdr_bytes + (LEAF / 2) * h->pair_block * eb;
    h->tile_stride = (int)al_align((size_t)h->tile_ids_off + LEAF * 4, 128);
    h->count[0] = h->n_leaves;
    int g = 0;
    while (h->count[g] > 1 && g + 1 < MAX_WIDE_LEVELS) {
        h->count[g + 1] = (h->count[g] + FANOUT - 1) / FANOUT;
        g++;
    }
    h->n_levels = g;
    size_t off = 256;
    h->perm_off = off;
    off += al_align((size_t)h->n_leaves * LEAF * 4);
    h->tiles_off = off;
    off += al_align((size_t)h->n_leaves * h->tile_stride);
    for (int l = 0; l <= g; l++) {
        h->box_off[l] = off;
        off += al_align((size_t)h->box_stride * h->count[l] * eb);
    }
    h->total_bytes = off;
}

// registry of headers of built indices (host side; avoids a device read per query)
bool index_lookup(const void* index, IndexHeader* out);
void index_register(const void* index, const IndexHeader& h);
