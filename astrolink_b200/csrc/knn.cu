// Host entry point of the kNN search (see knn_kernel.cuh for the kernel).
#include "knn_kernel.cuh"
#include <stdlib.h>

#define DECL(T, D) int knn_launch_##T##_d##D(const KnnParams& p, cudaStream_t st);
#define DECL_ALL(T) DECL(T, 1) DECL(T, 2) DECL(T, 3) DECL(T, 4) DECL(T, 5) DECL(T, 6) DECL(T, 7) DECL(T, 8) \
    DECL(T, 9) DECL(T, 10) DECL(T, 11) DECL(T, 12) DECL(T, 13) DECL(T, 14) DECL(T, 15) DECL(T, 16)
DECL_ALL(f32)
DECL_ALL(f64)

constexpr int KNN_MAX_D = 16;
typedef int (*knn_fn)(const KnnParams&, cudaStream_t);
#define TAB(T) {knn_launch_##T##_d1, knn_launch_##T##_d2, knn_launch_##T##_d3, knn_launch_##T##_d4, knn_launch_##T##_d5, \
                knn_launch_##T##_d6, knn_launch_##T##_d7, knn_launch_##T##_d8, knn_launch_##T##_d9, knn_launch_##T##_d10, \
                knn_launch_##T##_d11, knn_launch_##T##_d12, knn_launch_##T##_d13, knn_launch_##T##_d14, knn_launch_##T##_d15, \
                knn_launch_##T##_d16}
static const knn_fn g_knn_f32[KNN_MAX_D] = TAB(f32);
static const knn_fn g_knn_f64[KNN_MAX_D] = TAB(f64);

extern "C" size_t al_knn_workspace_bytes(int dtype, int64_t n, int d, int k) {
    (void)dtype; (void)n; (void)d; (void)k;
    return 256;
}

extern "C" int al_knn(const void* index, int64_t pos_begin, int64_t pos_end, int k, int row_order, uint32_t* idx_out,
                      void* d2_out, unsigned long long* pairs_evaluated, void* ws, size_t ws_bytes, void* stream) {
    return al_knn_link(index, pos_begin, pos_end, k, row_order, idx_out, d2_out, nullptr, 0, pairs_evaluated, ws, ws_bytes, stream);
}

extern "C" int al_knn_link(const void* index, int64_t pos_begin, int64_t pos_end, int k, int row_order, uint32_t* idx_out,
                           void* d2_out, uint32_t* link_out, int link_cols, unsigned long long* pairs_evaluated, void* ws,
                           size_t ws_bytes, void* stream) {
    return al_knn_dealt(index, pos_begin, pos_end, 0, 0, k, row_order, idx_out, d2_out, link_out, link_cols, pairs_evaluated, ws,
                        ws_bytes, stream);
}

extern "C" int64_t al_knn_dealt_rows(int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions) {
    if (pos_end <= pos_begin) return 0;
    if (chunk_positions <= 0) return pos_end - pos_begin;
    return (pos_end - pos_begin + stride_positions - 1) / stride_positions * chunk_positions;
}

extern "C" int al_knn_dealt(const void* index, int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions,
                            int k, int row_order, uint32_t* idx_out, void* d2_out, uint32_t* link_out, int link_cols,
                            unsigned long long* pairs_evaluated, void* ws, size_t ws_bytes, void* stream) {
    AL_REQUIRE(d2_out != nullptr, AL_EINVAL, "al_knn: d2_out is NULL");
    return al_knn_density(index, pos_begin, pos_end, chunk_positions, stride_positions, k, row_order, idx_out, d2_out, link_out, link_cols,
                          nullptr, 0, pairs_evaluated, ws, ws_bytes, stream);
}

extern "C" int al_knn_density(const void* index, int64_t pos_begin, int64_t pos_end, int64_t chunk_positions, int64_t stride_positions,
                              int k, int row_order, uint32_t* idx_out, void* d2_out, uint32_t* link_out, int link_cols,
                              void* logrho_out, int d_intrinsic, unsigned long long* pairs_evaluated, void* ws, size_t ws_bytes,
                              void* stream) {
    (void)ws; (void)ws_bytes;
    AL_REQUIRE(chunk_positions >= 0 && (chunk_positions == 0 || (chunk_positions % (LEAF * WARPS) == 0 && stride_positions >= chunk_positions &&
               stride_positions % (LEAF * WARPS) == 0)), AL_EINVAL,
               "al_knn_dealt: chunk_positions and stride_positions must be multiples of %d, stride_positions >= chunk_positions", LEAF * WARPS);
    AL_REQUIRE(chunk_positions == 0 || row_order == 1, AL_EINVAL, "al_knn_dealt: dealt chunks need row_order 1");
    // link rows and fused densities: dealt launches (the sharded path) write them in index-position order, everything
    // else by original index
    const int link_by_position = chunk_positions > 0 ? 1 : 0;
    AL_REQUIRE(link_out == nullptr || (link_cols >= 1 && link_cols <= k), AL_EINVAL, "al_knn_link: link_cols must be in [1, k]");
    AL_REQUIRE(d2_out != nullptr || logrho_out != nullptr, AL_EINVAL, "al_knn_density: d2_out may only be NULL when logrho_out is given");
    AL_REQUIRE(idx_out != nullptr || link_out != nullptr, AL_EINVAL, "al_knn: idx_out may only be NULL when link_out is given");
    AL_REQUIRE(logrho_out == nullptr || d_intrinsic >= 1, AL_EINVAL, "al_knn_density: d_intrinsic must be >= 1");
    IndexHeader h;
    if (!index_lookup(index, &h)) {
        AL_CUDA(cudaMemcpy(&h, index, sizeof(h), cudaMemcpyDeviceToHost));
        AL_REQUIRE(h.magic == INDEX_MAGIC, AL_EINVAL, "al_knn: not an index built by al_index_build");
        index_register(index, h);
    }
    AL_REQUIRE(k >= 1 && k <= h.n, AL_EINVAL, "al_knn: k must be in [1, n]");
    AL_REQUIRE(k <= 128, AL_EINVAL, "al_knn: k > 128 is not supported");
    AL_REQUIRE(h.d >= 1 && h.d <= KNN_MAX_D, AL_EINVAL, "al_knn: d=%d not supported (1..%d)", h.d, KNN_MAX_D);
    i64 np = h.n_leaves * LEAF;
    AL_REQUIRE(pos_begin >= 0 && pos_begin <= pos_end && pos_end <= np, AL_EINVAL, "al_knn: bad position range");
    AL_REQUIRE(pos_begin % LEAF == 0 && pos_end % LEAF == 0, AL_EINVAL, "al_knn: positions must be multiples of %d", LEAF);
    AL_REQUIRE(row_order == 0 || row_order == 1, AL_EINVAL, "al_knn: row_order must be 0 or 1");
    KnnParams p;
    const char* base = (const char*)index;
    p.tiles = base + h.tiles_off;
    for (int l = 0; l < MAX_WIDE_LEVELS; l++) { p.box[l] = base + h.box_off[l]; p.count[l] = h.count[l]; }
    p.n_levels = h.n_levels;
    p.tile_stride = h.tile_stride;
    p.tile_hdr_bytes = h.tile_hdr_bytes;
    p.tile_ids_off = h.tile_ids_off;
    p.pair_block = h.pair_block;
    p.box_stride = h.box_stride;
    p.leaf_begin = pos_begin / LEAF;
    p.leaf_end = pos_end / LEAF;
    p.chunk_blocks = chunk_positions / (LEAF * WARPS);
    p.chunk_stride_blocks = stride_positions / (LEAF * WARPS);
    p.n = h.n;
    p.k = k;
    p.row_order = row_order;
    p.idx_out = idx_out;
    p.d2_out = d2_out;
    p.link_out = link_out;
    p.link_cols = link_cols;
    p.link_by_position = link_by_position;
    p.logrho_out = logrho_out;
    p.half_d = d_intrinsic / 2.0;
    p.stats = pairs_evaluated;
    // AL_KNN_GENERIC=1 (development): float32 indices are searched by the generic kernel instead of the float32 one
    static const bool generic = getenv("AL_KNN_GENERIC") != nullptr;
    p.use_generic = generic ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (h.dtype == AL_F32) {
        p.warp_smem = knn_warp_smem_bytes<float>(h.tile_stride, k, h.d, h.n_levels);     // generic kernel; the float32 one sizes its own
        return g_knn_f32[h.d - 1](p, st);
    }
    p.warp_smem = knn_warp_smem_bytes<double>(h.tile_stride, k, h.d, h.n_levels);
    return g_knn_f64[h.d - 1](p, st);
}
