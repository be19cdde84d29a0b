// CUDA-free basics shared by the aggregation algorithm (aggregate_impl.h), its
// CUDA executor (aggregate.cu) and the test-only serial host executor
// (tests/hostsim/agg_hostsim.cpp).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef int64_t i64;
typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDACC__)
#define AGG_HD __host__ __device__
#define AGG_LAMBDA [=] __host__ __device__
#else
#define AGG_HD
#define AGG_LAMBDA [=]
#endif

constexpr u32 AGG_NONE = 0xffffffffu;

struct AggArena {
    char* base;
    size_t cap;
    size_t off;
    size_t high;
    bool ok;
    AggArena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), high(0), ok(true) {}
    size_t mark() const { return off; }
    void reset(size_t m) { off = m; }
    template <typename T>
    T* get(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
        if (bytes == 0) bytes = 256;
        if (off + bytes > cap) { ok = false; return nullptr; }
        T* r = (T*)(base + off);
        off += bytes;
        if (off > high) high = off;
        return r;
    }
};

AGG_HD inline u32 agg_atomic_min(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
    return atomicMin(p, v);
#else
    u32 o = *p;
    if (v < o) *p = v;
    return o;
#endif
}

AGG_HD inline u32 agg_atomic_or(u32* p, u32 v) {
#if defined(__CUDA_ARCH__)
    return atomicOr(p, v);
#else
    u32 o = *p;
    *p = o | v;
    return o;
#endif
}

// plain load the compiler may not hoist or fuse (values other threads are updating: pointer jumping)
AGG_HD inline u32 agg_load_relaxed(const u32* p) { return *(const volatile u32*)p; }

AGG_HD inline int agg_log2_ceil(u64 x) {
    int b = 0;
    while (((u64)1 << b) < x) b++;
    return b;
}
