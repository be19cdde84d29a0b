// Shared helpers for the astrolink_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/astrolink_b200.h"

typedef int64_t i64;
typedef uint32_t u32;
typedef uint64_t u64;

void al_set_error(const char* fmt, ...);
// number of kernels this library launched (its own kernels; CUB's internal ones are not counted)
void al_count_launches(long n);

#define AL_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            al_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return AL_ECUDA;                                                               \
        }                                                                                  \
    } while (0)

#define AL_CHECK_LAUNCH()            \
    do {                             \
        al_count_launches(1);        \
        AL_CUDA(cudaGetLastError()); \
    } while (0)

#define AL_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            al_set_error(__VA_ARGS__);   \
            return (code);               \
        }                                \
    } while (0)

static inline size_t al_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-owned workspace.
struct Arena {
    char* base;
    size_t cap;
    size_t off;
    bool ok;
    Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), ok(true) {}
    template <typename T>
    T* get(size_t count) {
        size_t bytes = al_align(count * sizeof(T));
        if (off + bytes > cap) { ok = false; return nullptr; }
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};
// Same arithmetic without memory, for *_workspace_bytes().
struct ArenaSizer {
    size_t off = 0;
    template <typename T>
    void add(size_t count) { off += al_align(count * sizeof(T)); }
};

static inline int al_grid(i64 n, int block, int max_blocks = 148 * 16) {
    i64 g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > max_blocks) g = max_blocks;
    return (int)g;
}

// order-preserving float <-> unsigned maps (ascending)
__host__ __device__ static inline u32 f32_to_ordered(float f) {
    u32 b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(f);
#else
    memcpy(&b, &f, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ static inline u64 f64_to_ordered(double f) {
    u64 b;
#ifdef __CUDA_ARCH__
    b = (u64)__double_as_longlong(f);
#else
    memcpy(&b, &f, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

#ifdef __CUDACC__
// a4 raw log-density (astrolink.py:293-297): ln((k - S/c) / c^(d_int/2)) with S the T-precision sequential sum of
// the k squared distances and c the largest one.  One definition shared by logrho_kernel and the kNN epilogues so
// the fused and the stand-alone pass give the same bits.
template <typename T>
__device__ __forceinline__ T al_raw_logrho(T sum, T coreT, int k, double half_d) {
    const double v = ((double)k - (double)(T)(sum / coreT)) / pow((double)coreT, half_d);
    return (T)log(v);
}
#endif
