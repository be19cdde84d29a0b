// CUDA executor + C entry point of the aggregation pass (algorithm in
// aggregate_impl.h): replaces _aggregate_njit_float{32,64}_uint32
// (reference astrolink/astrolink.py:384-604).
#include <cub/cub.cuh>

#include <iterator>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/discard_iterator.h>

#include "aggregate_impl.h"
#include "common.cuh"
#include <stdlib.h>
#include <time.h>

template <class Tag, class F>
__global__ void __launch_bounds__(256) agg_kernel(i64 n, F f) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) f(i);
}

// the same, skipped as a whole when *flag == 0 (flag == nullptr: always runs)
template <class Tag, class F>
__global__ void __launch_bounds__(256) agg_kernel_if(const u32* flag, i64 n, F f) {
    if (flag != nullptr && *flag == 0u) return;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) f(i);
}

// ---- stable selection through cub::DeviceSelect::If: the "items" are the indices 0..n-1, the output iterator hands
// (index, position among the selected) to the caller's emit functor instead of storing anything itself
template <class Emit>
struct AggEmitProxy {
    Emit emit;
    i64 pos;
    __host__ __device__ const AggEmitProxy& operator=(u32 idx) const { emit((i64)idx, pos); return *this; }
};
template <class Emit>
struct AggEmitIterator {
    typedef std::random_access_iterator_tag iterator_category;
    typedef u32 value_type;
    typedef i64 difference_type;
    typedef void pointer;
    typedef AggEmitProxy<Emit> reference;
    Emit emit;
    i64 base;
    __host__ __device__ AggEmitProxy<Emit> operator[](i64 k) const { return AggEmitProxy<Emit>{emit, base + k}; }
    __host__ __device__ AggEmitProxy<Emit> operator*() const { return AggEmitProxy<Emit>{emit, base}; }
    __host__ __device__ AggEmitIterator operator+(i64 k) const { return AggEmitIterator{emit, base + k}; }
};
template <class Pred>
struct AggPredOp {
    Pred pred;
    __host__ __device__ bool operator()(const u32& i) const { return pred((i64)i); }
};

__global__ void agg_fill_kernel(u32* p, u32 v, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void agg_total_kernel(const u32* last_in, const u32* last_out, u32* tot) { tot[0] = *last_in + *last_out; }

// Everything runs on one stream, so workspace released with ar.reset() may be handed to the next
// kernel without a host synchronisation: stream order already separates the two uses.
struct CudaExec {
    cudaStream_t st;
    cudaError_t err = cudaSuccess;
    long launches = 0;
    long reads = 0;

    void note(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }

    template <class Tag, class F>
    void launch(i64 n, F f) {
        if (n <= 0) return;
        i64 g = (n + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        agg_kernel<Tag, F><<<(unsigned)g, 256, 0, st>>>(n, f);
        note(cudaGetLastError());
        launches++;
    }
    template <class Tag, class F>
    void launch_if(const u32* flag, i64 n, F f) {
        if (n <= 0) return;
        i64 g = (n + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        agg_kernel_if<Tag, F><<<(unsigned)g, 256, 0, st>>>(flag, n, f);
        note(cudaGetLastError());
        launches++;
    }
    // for every i in [0, n) with pred(i), in order: emit(i, rank of i among the selected); the count goes to count_dev[0]
    template <class Tag, class Pred, class Emit>
    void select(i64 n, Pred pred, Emit emit, u32* count_dev, AggArena& ar) {
        if (n <= 0) { note(cudaMemsetAsync(count_dev, 0, 4, st)); return; }
        const size_t mk = ar.mark();
        thrust::counting_iterator<u32> in(0u);
        AggEmitIterator<Emit> out{emit, 0};
        AggPredOp<Pred> op{pred};
        size_t tb = 0;
        note(cub::DeviceSelect::If(nullptr, tb, in, out, count_dev, n, op, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DeviceSelect::If(tmp, tb, in, out, count_dev, n, op, st));
        launches++;
        ar.reset(mk);
    }
    // two selections in ONE pass over [0, n): for every i with pred1(i), in order, emit1(i, rank among those); for every other
    // i with pred2(i), in order, emit2(i, rank among those); the counts go to count_dev[0] and count_dev[1]
    // (cub::DevicePartition::If, three-way: both selected parts keep their relative order; the rest is discarded)
    template <class Tag, class Pred1, class Emit1, class Pred2, class Emit2>
    void select2(i64 n, Pred1 pred1, Emit1 emit1, Pred2 pred2, Emit2 emit2, u32* count_dev, AggArena& ar) {
        if (n <= 0) { note(cudaMemsetAsync(count_dev, 0, 8, st)); return; }
        const size_t mk = ar.mark();
        thrust::counting_iterator<u32> in(0u);
        AggEmitIterator<Emit1> out1{emit1, 0};
        AggEmitIterator<Emit2> out2{emit2, 0};
        thrust::discard_iterator<> rest;
        AggPredOp<Pred1> op1{pred1};
        AggPredOp<Pred2> op2{pred2};
        size_t tb = 0;
        note(cub::DevicePartition::If(nullptr, tb, in, out1, out2, rest, count_dev, n, op1, op2, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DevicePartition::If(tmp, tb, in, out1, out2, rest, count_dev, n, op1, op2, st));
        launches++;
        ar.reset(mk);
    }
    void fill_u32(u32* p, u32 v, i64 n) {
        if (n <= 0) return;
        if (v == 0u || v == 0xffffffffu) note(cudaMemsetAsync(p, v == 0u ? 0 : 0xff, (size_t)n * 4, st));
        else {
            i64 g = (n + 255) / 256;
            if (g > 148 * 16) g = 148 * 16;
            agg_fill_kernel<<<(unsigned)g, 256, 0, st>>>(p, v, n);
            note(cudaGetLastError());
        }
    }
    // one blocking read of `count` consecutive device words (the only host round trips of the pass)
    void read_u32s(const u32* p, u32* host, int count) {
        note(cudaMemcpyAsync(host, p, 4 * (size_t)count, cudaMemcpyDeviceToHost, st));
        note(cudaStreamSynchronize(st));
        reads++;
    }
    // out[i] = sum of in[0..i); the total goes to total_dev[0] on the device (no host synchronisation).  `out` must
    // not alias `in`.
    void exclusive_sum_u32(const u32* in, u32* out, i64 n, AggArena& ar, u32* total_dev) {
        if (n <= 0) { note(cudaMemsetAsync(total_dev, 0, 4, st)); return; }
        const size_t mk = ar.mark();
        size_t tb = 0;
        note(cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, st));
        agg_total_kernel<<<1, 1, 0, st>>>(in + (n - 1), out + (n - 1), total_dev);
        note(cudaGetLastError());
        launches += 1;
        ar.reset(mk);
    }
    void inclusive_sum_u32(const u32* in, u32* out, i64 n, AggArena& ar) {
        if (n <= 0) return;
        const size_t mk = ar.mark();
        size_t tb = 0;
        note(cub::DeviceScan::InclusiveSum(nullptr, tb, in, out, n, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DeviceScan::InclusiveSum(tmp, tb, in, out, n, st));
        ar.reset(mk);
    }
    void sort_pairs_u64_u32(const u64* k_in, u64* k_out, const u32* v_in, u32* v_out, i64 n, int end_bit, AggArena& ar) {
        if (n <= 0) return;
        const size_t mk = ar.mark();
        size_t tb = 0;
        note(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, n, 0, end_bit, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, v_out, n, 0, end_bit, st));
        ar.reset(mk);
    }
    void sort_pairs_u32_u32(const u32* k_in, u32* k_out, const u32* v_in, u32* v_out, i64 n, int end_bit, AggArena& ar) {
        if (n <= 0) return;
        if (end_bit < 1) end_bit = 1;
        if (end_bit > 32) end_bit = 32;
        const size_t mk = ar.mark();
        size_t tb = 0;
        note(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, v_in, v_out, n, 0, end_bit, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DeviceRadixSort::SortPairs(tmp, tb, k_in, k_out, v_in, v_out, n, 0, end_bit, st));
        ar.reset(mk);
    }
    void sync() {}      // stream order is enough between the passes (see above); hosts that need results call read_u32
    void segmented_inclusive_scan(SegItem* items, i64 n, AggArena& ar) {
        if (n <= 0) return;
        const size_t mk = ar.mark();
        size_t tb = 0;
        note(cub::DeviceScan::InclusiveScan(nullptr, tb, items, items, SegOp(), n, st));
        void* tmp = ar.get<char>(tb);
        if (!ar.ok) { ar.reset(mk); return; }
        note(cub::DeviceScan::InclusiveScan(tmp, tb, items, items, SegOp(), n, st));
        ar.reset(mk);
    }
    char lvl_name[64];
    void phase_level(const char* base, int l, i64 m, i64 N) {
        if (!prof) return;
        snprintf(lvl_name, sizeof(lvl_name), "%s-%02d m=%lld N=%lld", base, l, (long long)m, (long long)N);
        phase(lvl_name);
    }
    // optional phase timing (AL_PROFILE=1): wall time between phase markers, stream synchronised
    bool prof = false;
    const char* cur = nullptr;
    double t0 = 0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    void phase(const char* name) {
        if (!prof) return;
        cudaStreamSynchronize(st);
        double t = now();
        if (cur) fprintf(stderr, "[al_aggregate] %-36s %8.3f ms  (launches so far %ld, host reads %ld)\n", cur, t - t0, launches, reads);
        static char keep[64];
        snprintf(keep, sizeof(keep), "%s", name);
        cur = keep;
        t0 = t;
    }
};

extern "C" size_t al_aggregate_workspace_bytes(int dtype, int64_t n, int64_t E) {
    return agg::workspace_bound((size_t)n, (size_t)E, dtype == AL_F64 ? 8 : 4);
}

extern "C" int al_aggregate(const void* logrho, int dtype, const uint32_t* sorted_pairs, int64_t n, int64_t E, uint32_t* ordering_out,
                            uint32_t* groups_out, void* prominences_out, int64_t* n_groups_host, void* ws, size_t ws_bytes,
                            void* stream) {
    AL_REQUIRE(dtype == AL_F32 || dtype == AL_F64, AL_EINVAL, "al_aggregate: bad dtype");
    AL_REQUIRE(n >= 2 && n < 0x7fffffffll && E >= 1 && E < 0xffffffffll, AL_EINVAL, "al_aggregate: n or E out of range");
    AL_REQUIRE(n_groups_host != nullptr, AL_EINVAL, "al_aggregate: n_groups_host is NULL");
    CudaExec x;
    x.st = (cudaStream_t)stream;
    x.prof = getenv("AL_PROFILE") != nullptr;
    AggArena ar(ws, ws_bytes);
    int rc;
    if (dtype == AL_F32)
        rc = agg::run<CudaExec, float>(x, ar, (const float*)logrho, sorted_pairs, n, E, false, ordering_out, groups_out,
                                       (float*)prominences_out, n_groups_host);
    else
        rc = agg::run<CudaExec, double>(x, ar, (const double*)logrho, sorted_pairs, n, E, true, ordering_out, groups_out,
                                        (double*)prominences_out, n_groups_host);
    x.note(cudaStreamSynchronize(x.st));     // surfaces asynchronous kernel errors in this call's return code
    al_count_launches(x.launches);
    if (x.err != cudaSuccess) {
        al_set_error("al_aggregate: CUDA error: %s", cudaGetErrorString(x.err));
        return AL_ECUDA;
    }
    if (rc == -2) {
        al_set_error("al_aggregate: workspace too small (%zu bytes given, %zu used before failing)", ws_bytes, ar.high);
        return AL_ENOMEM;
    }
    if (rc != 0) {
        al_set_error("al_aggregate: internal consistency check %d failed (is every point an endpoint of some edge?)", rc);
        return AL_EINTERNAL;
    }
    return AL_OK;
}
