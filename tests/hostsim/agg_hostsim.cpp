// TEST INFRASTRUCTURE ONLY -- never loaded by the product package.
//
// Serial host executor for the aggregation algorithm in
// astrolink_b200/csrc/aggregate_impl.h.  It runs every data-parallel pass of
// that algorithm as a plain loop (one valid interleaving of the GPU threads), so
// the LOGIC of the parallel formulation can be checked against the oracle on a
// machine without a GPU (`pytest -m "not gpu"`).  It says nothing about races
// or CUDA specifics; those are covered by the -m gpu parity tests.
#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "../../astrolink_b200/csrc/aggregate_impl.h"

struct HostExec {
    long launches = 0;
    // order in which the "threads" of a pass run: 0 ascending, 1 descending, 2 a fixed pseudo-random permutation --
    // three different valid interleavings of the GPU threads; the passes must not depend on which one they get
    int order = 0;
    template <class Tag, class F>
    void launch(i64 n, F f) {
        if (order == 0) {
            for (i64 i = 0; i < n; i++) f(i);
        } else if (order == 1) {
            for (i64 i = n - 1; i >= 0; i--) f(i);
        } else {
            // i -> (a * i + b) mod n with a coprime to n visits every index once
            i64 a = 2654435761ll % (n > 0 ? n : 1);
            if (a == 0) a = 1;
            while (std::gcd(a, n) != 1) a++;
            i64 v = 12345 % (n > 0 ? n : 1);
            for (i64 i = 0; i < n; i++) { f(v); v += a; if (v >= n) v -= n; }
        }
        launches++;
    }
    template <class Tag, class F>
    void launch_if(const u32* flag, i64 n, F f) {
        if (flag != nullptr && *flag == 0u) return;
        launch<Tag>(n, f);
    }
    template <class Tag, class Pred, class Emit>
    void select(i64 n, Pred pred, Emit emit, u32* count, AggArena&) {
        // the predicate of every element is evaluated before any emit runs (on the GPU both happen inside one pass, with
        // no ordering between different elements): emits must not change what a predicate reads
        std::vector<char> keep((size_t)(n > 0 ? n : 0));
        launch<Tag>(n, [&](i64 i) { keep[(size_t)i] = pred(i) ? 1 : 0; });
        i64 pos = 0;
        for (i64 i = 0; i < n; i++) if (keep[(size_t)i]) emit(i, pos++);
        count[0] = (u32)pos;
    }
    template <class Tag, class Pred1, class Emit1, class Pred2, class Emit2>
    void select2(i64 n, Pred1 pred1, Emit1 emit1, Pred2 pred2, Emit2 emit2, u32* count, AggArena&) {
        // as select: all predicates first, then the emits (first part, then second part; on the GPU they interleave)
        std::vector<char> part((size_t)(n > 0 ? n : 0));
        launch<Tag>(n, [&](i64 i) { part[(size_t)i] = pred1(i) ? 1 : (pred2(i) ? 2 : 0); });
        i64 p1 = 0, p2 = 0;
        for (i64 i = 0; i < n; i++) if (part[(size_t)i] == 1) emit1(i, p1++);
        for (i64 i = 0; i < n; i++) if (part[(size_t)i] == 2) emit2(i, p2++);
        count[0] = (u32)p1;
        count[1] = (u32)p2;
    }
    void fill_u32(u32* p, u32 v, i64 n) { for (i64 i = 0; i < n; i++) p[i] = v; }
    void read_u32s(const u32* p, u32* host, int count) { for (int i = 0; i < count; i++) host[i] = p[i]; }
    void exclusive_sum_u32(const u32* in, u32* out, i64 n, AggArena&, u32* total) {
        i64 s = 0;
        for (i64 i = 0; i < n; i++) { u32 v = in[i]; out[i] = (u32)s; s += v; }
        total[0] = (u32)s;
    }
    void inclusive_sum_u32(const u32* in, u32* out, i64 n, AggArena&) {
        u32 s = 0;
        for (i64 i = 0; i < n; i++) { s += in[i]; out[i] = s; }
    }
    void sort_pairs_u64_u32(const u64* k_in, u64* k_out, const u32* v_in, u32* v_out, i64 n, int end_bit, AggArena&) {
        std::vector<i64> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        const u64 mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
        std::stable_sort(idx.begin(), idx.end(), [&](i64 a, i64 b) { return (k_in[a] & mask) < (k_in[b] & mask); });
        for (i64 i = 0; i < n; i++) { k_out[i] = k_in[idx[i]]; v_out[i] = v_in[idx[i]]; }
    }
    void sort_pairs_u32_u32(const u32* k_in, u32* k_out, const u32* v_in, u32* v_out, i64 n, int end_bit, AggArena&) {
        std::vector<i64> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        const u32 mask = end_bit >= 32 ? ~0u : ((1u << end_bit) - 1);
        std::stable_sort(idx.begin(), idx.end(), [&](i64 a, i64 b) { return (k_in[a] & mask) < (k_in[b] & mask); });
        for (i64 i = 0; i < n; i++) { k_out[i] = k_in[idx[i]]; v_out[i] = v_in[idx[i]]; }
    }
    void sync() {}
    void phase(const char*) {}
    void phase_level(const char*, int, i64, i64) {}
    void segmented_inclusive_scan(SegItem* items, i64 n, AggArena&) {
        SegOp op;
        for (i64 i = 1; i < n; i++) items[i] = op(items[i - 1], items[i]);
    }
};

template <typename T>
static i64 run_host(const T* logrho, const u32* sorted_pairs, i64 n, i64 E, bool quirk, u32* ordering, u32* groups, T* prom) {
    size_t bytes = agg::workspace_bound((size_t)n, (size_t)E, sizeof(T));
    void* ws = malloc(bytes);
    AggArena ar(ws, bytes);
    HostExec x;
    if (const char* e = getenv("HOSTSIM_ORDER")) x.order = atoi(e);
    i64 G = 0;
    int rc = agg::run<HostExec, T>(x, ar, logrho, sorted_pairs, n, E, quirk, ordering, groups, prom, &G);
    free(ws);
    return rc == 0 ? G : (i64)rc;
}

extern "C" i64 hostsim_aggregate_f32(const float* logrho, const u32* sorted_pairs, i64 n, i64 E, u32* ordering, u32* groups, float* prom) {
    return run_host<float>(logrho, sorted_pairs, n, E, false, ordering, groups, prom);
}
extern "C" i64 hostsim_aggregate_f64(const double* logrho, const u32* sorted_pairs, i64 n, i64 E, u32* ordering, u32* groups, double* prom) {
    return run_host<double>(logrho, sorted_pairs, n, E, true, ordering, groups, prom);
}
