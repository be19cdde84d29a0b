/*
 * astrolink_oracle.c -- CPU restatement of AstroLink's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA
 * implementation in astrolink_b200/csrc.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product
 * package never does.
 *
 * Every function restates one step of the reference
 * (/root/reference/astrolink/astrolink.py, cited as file:line below).  It is
 * written from the behavioural description of that code, in plain C, with
 * array-based data structures instead of Python lists.
 *
 * Parity status:
 *   - kNN: the reference delegates to the third-party `pykdtree` package
 *     (astrolink.py:16,270,279), unpinned (pyproject.toml:29-36) and absent
 *     from this machine.  Its published behaviour (exact k nearest, squared
 *     distances accumulated dimension by dimension in the data dtype, results
 *     ascending) is restated here with a DEFINED tie-break (index ascending)
 *     and a DEFINED distance formula: acc = fma(q_j - p_j, q_j - p_j, acc) for
 *     j = 0..d-1.  "parity unpinned" for this one step: the reference holds no
 *     golden vectors for it.
 *   - every other step is pinned against the reference's own numba functions
 *     run in this container (oracle/gen_golden.py -> tests/golden/).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef uint32_t u32;

int oracle_abi_version(void) { return 1; }

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU baseline sets its thread count explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* transform_data / _rescale_njit  (astrolink.py:218-243)                    */
/* Each feature divided by its population standard deviation.  A zero-       */
/* variance column is copied through (the reference leaves it uninitialised, */
/* astrolink.py:239-242).                                  */
/* ------------------------------------------------------------------------ */
#define DEFINE_RESCALE(NAME, T)                                               \
    void NAME(const T* P, i64 n, int d, T* out, double* std_out) {            \
        for (int j = 0; j < d; j++) {                                         \
            double mean = 0.0;                                                \
            for (i64 i = 0; i < n; i++) mean += (double)P[i * d + j];         \
            mean /= (double)n;                                                \
            double var = 0.0;                                                 \
            for (i64 i = 0; i < n; i++) {                                     \
                double t = (double)P[i * d + j] - mean;                       \
                var += t * t;                                                 \
            }                                                                 \
            double sd = sqrt(var / (double)n);                                \
            if (std_out) std_out[j] = sd;                                     \
            T sdt = (T)sd;                                                    \
            for (i64 i = 0; i < n; i++)                                       \
                out[i * d + j] = (sd > 0.0) ? (T)(P[i * d + j] / sdt) : P[i * d + j]; \
        }                                                                     \
    }
DEFINE_RESCALE(oracle_rescale_f32, float)
DEFINE_RESCALE(oracle_rescale_f64, double)

/* ------------------------------------------------------------------------ */
/* kNN  (astrolink.py:270,279: KDTree(P).query(P, k, sqr_dists=True))        */
/* ------------------------------------------------------------------------ */
static inline float fma_T_f(float a, float b, float c) { return fmaf(a, b, c); }
static inline double fma_T_d(double a, double b, double c) { return fma(a, b, c); }

/* (d2, idx) lexicographic "a is worse (later) than b" */
#define WORSE(d2a, ia, d2b, ib) ((d2a) > (d2b) || ((d2a) == (d2b) && (ia) > (ib)))

#define DEFINE_KNN(SUF, T, FMA)                                               \
    typedef struct {                                                          \
        const T* P; i64 n; int d; int leaf;                                   \
        i64* perm;            /* tree order -> original index */              \
        i64 n_nodes; i64* lo; i64* hi; i64* left; i64* right;                 \
        T* bmin; T* bmax;     /* per node AABB */                             \
    } kdt_##SUF;                                                              \
                                                                              \
    static inline T dist2_##SUF(const T* q, const T* p, int d) {              \
        T acc = (T)0;                                                         \
        for (int j = 0; j < d; j++) { T df = q[j] - p[j]; acc = FMA(df, df, acc); } \
        return acc;                                                           \
    }                                                                         \
    /* lower bound on dist2 from q to any point of the box, same fma chain  */\
    static inline T boxdist2_##SUF(const T* q, const T* bmin, const T* bmax, int d) { \
        T acc = (T)0;                                                         \
        for (int j = 0; j < d; j++) {                                         \
            T a = bmin[j] - q[j], b = q[j] - bmax[j];                         \
            T g = a > b ? a : b; if (g < (T)0) g = (T)0;                      \
            acc = FMA(g, g, acc);                                             \
        }                                                                     \
        return acc;                                                           \
    }                                                                         \
    static void select_##SUF(const T* P, int d, int dim, i64* a, i64 n, i64 kth) { \
        i64 lo = 0, hi = n - 1;                                               \
        while (lo < hi) {                                                     \
            T pv = P[a[(lo + hi) / 2] * d + dim];                             \
            i64 i = lo, j = hi;                                               \
            while (i <= j) {                                                  \
                while (P[a[i] * d + dim] < pv) i++;                           \
                while (P[a[j] * d + dim] > pv) j--;                           \
                if (i <= j) { i64 t = a[i]; a[i] = a[j]; a[j] = t; i++; j--; }\
            }                                                                 \
            if (kth <= j) hi = j; else if (kth >= i) lo = i; else break;      \
        }                                                                     \
    }                                                                         \
    static i64 build_##SUF(kdt_##SUF* t, i64 lo, i64 hi) {                    \
        i64 id = t->n_nodes++;                                                \
        int d = t->d;                                                         \
        t->lo[id] = lo; t->hi[id] = hi; t->left[id] = -1; t->right[id] = -1;  \
        T* mn = t->bmin + id * d; T* mx = t->bmax + id * d;                   \
        for (int j = 0; j < d; j++) { mn[j] = t->P[t->perm[lo] * d + j]; mx[j] = mn[j]; } \
        for (i64 i = lo + 1; i < hi; i++)                                     \
            for (int j = 0; j < d; j++) {                                     \
                T v = t->P[t->perm[i] * d + j];                               \
                if (v < mn[j]) mn[j] = v; if (v > mx[j]) mx[j] = v;           \
            }                                                                 \
        if (hi - lo > t->leaf) {                                              \
            int dim = 0; T best = mx[0] - mn[0];                              \
            for (int j = 1; j < d; j++) if (mx[j] - mn[j] > best) { best = mx[j] - mn[j]; dim = j; } \
            i64 mid = (hi - lo) / 2;                                          \
            select_##SUF(t->P, d, dim, t->perm + lo, hi - lo, mid);           \
            i64 l = build_##SUF(t, lo, lo + mid);                             \
            i64 r = build_##SUF(t, lo + mid, hi);                             \
            t->left[id] = l; t->right[id] = r;                                \
        }                                                                     \
        return id;                                                            \
    }                                                                         \
    typedef struct { T* d2; i64* ix; int k; int cnt; } heap_##SUF;            \
    static inline void heap_push_##SUF(heap_##SUF* h, T d2, i64 ix) {         \
        /* max-heap on (d2, idx): root = worst of the k kept */               \
        if (h->cnt < h->k) {                                                  \
            int c = h->cnt++;                                                 \
            while (c > 0) {                                                   \
                int p = (c - 1) / 2;                                          \
                if (WORSE(d2, ix, h->d2[p], h->ix[p])) { h->d2[c] = h->d2[p]; h->ix[c] = h->ix[p]; c = p; } \
                else break;                                                   \
            }                                                                 \
            h->d2[c] = d2; h->ix[c] = ix;                                     \
        } else if (WORSE(h->d2[0], h->ix[0], d2, ix)) {                       \
            int c = 0;                                                        \
            for (;;) {                                                        \
                int l = 2 * c + 1, r = l + 1, m = -1;                         \
                if (l < h->k) m = l;                                          \
                if (r < h->k && WORSE(h->d2[r], h->ix[r], h->d2[l], h->ix[l])) m = r; \
                if (m >= 0 && WORSE(h->d2[m], h->ix[m], d2, ix)) { h->d2[c] = h->d2[m]; h->ix[c] = h->ix[m]; c = m; } \
                else break;                                                   \
            }                                                                 \
            h->d2[c] = d2; h->ix[c] = ix;                                     \
        }                                                                     \
    }                                                                         \
    static void search_##SUF(const kdt_##SUF* t, i64 node, const T* q, heap_##SUF* h) { \
        int d = t->d;                                                         \
        if (h->cnt == h->k) {                                                 \
            T bd = boxdist2_##SUF(q, t->bmin + node * d, t->bmax + node * d, d); \
            if (bd > h->d2[0]) return; /* strict: equal distance may still win on index */ \
        }                                                                     \
        if (t->left[node] < 0) {                                              \
            for (i64 i = t->lo[node]; i < t->hi[node]; i++) {                 \
                i64 p = t->perm[i];                                           \
                heap_push_##SUF(h, dist2_##SUF(q, t->P + p * d, d), p);       \
            }                                                                 \
            return;                                                           \
        }                                                                     \
        i64 l = t->left[node], r = t->right[node];                            \
        T dl = boxdist2_##SUF(q, t->bmin + l * d, t->bmax + l * d, d);        \
        T dr = boxdist2_##SUF(q, t->bmin + r * d, t->bmax + r * d, d);        \
        if (dl <= dr) { search_##SUF(t, l, q, h); search_##SUF(t, r, q, h); } \
        else { search_##SUF(t, r, q, h); search_##SUF(t, l, q, h); }          \
    }                                                                         \
    /* exact kNN of queries [q_begin, q_end) of P against all of P          */\
    int oracle_knn_##SUF(const T* P, i64 n, int d, i64 q_begin, i64 q_end, int k, \
                         u32* idx_out, T* d2_out) {                           \
        if (k > n || k < 1) return -1;                                        \
        kdt_##SUF t; t.P = P; t.n = n; t.d = d; t.leaf = 16; t.n_nodes = 0;   \
        i64 maxn = 2 * (n / 8 + 2) + 16;                                      \
        if (maxn < 64) maxn = 64;                                             \
        t.perm = (i64*)malloc(sizeof(i64) * n);                               \
        t.lo = (i64*)malloc(sizeof(i64) * maxn); t.hi = (i64*)malloc(sizeof(i64) * maxn); \
        t.left = (i64*)malloc(sizeof(i64) * maxn); t.right = (i64*)malloc(sizeof(i64) * maxn); \
        t.bmin = (T*)malloc(sizeof(T) * maxn * d); t.bmax = (T*)malloc(sizeof(T) * maxn * d); \
        for (i64 i = 0; i < n; i++) t.perm[i] = i;                            \
        build_##SUF(&t, 0, n);                                                \
        _Pragma("omp parallel")                                               \
        {                                                                     \
            heap_##SUF h; h.k = k;                                            \
            h.d2 = (T*)malloc(sizeof(T) * k); h.ix = (i64*)malloc(sizeof(i64) * k); \
            _Pragma("omp for schedule(dynamic, 256)")                         \
            for (i64 qi = q_begin; qi < q_end; qi++) {                        \
                h.cnt = 0;                                                    \
                search_##SUF(&t, 0, P + qi * d, &h);                          \
                /* heap-sort ascending by (d2, idx) */                        \
                u32* io = idx_out + (qi - q_begin) * k; T* dd = d2_out + (qi - q_begin) * k; \
                for (int m = k - 1; m >= 0; m--) {                            \
                    dd[m] = h.d2[0]; io[m] = (u32)h.ix[0];                    \
                    T ld = h.d2[m]; i64 li = h.ix[m];                         \
                    int c = 0;                                                \
                    for (;;) {                                                \
                        int l = 2 * c + 1, r = l + 1, mm = -1;                \
                        if (l < m) mm = l;                                    \
                        if (r < m && WORSE(h.d2[r], h.ix[r], h.d2[l], h.ix[l])) mm = r; \
                        if (mm >= 0 && WORSE(h.d2[mm], h.ix[mm], ld, li)) { h.d2[c] = h.d2[mm]; h.ix[c] = h.ix[mm]; c = mm; } \
                        else break;                                           \
                    }                                                         \
                    h.d2[c] = ld; h.ix[c] = li;                               \
                }                                                             \
            }                                                                 \
            free(h.d2); free(h.ix);                                           \
        }                                                                     \
        free(t.perm); free(t.lo); free(t.hi); free(t.left); free(t.right); free(t.bmin); free(t.bmax); \
        return 0;                                                             \
    }                                                                         \
    /* brute force, for validating the tree search on small inputs          */\
    int oracle_knn_brute_##SUF(const T* P, i64 n, int d, i64 q_begin, i64 q_end, int k, \
                               u32* idx_out, T* d2_out) {                     \
        if (k > n || k < 1) return -1;                                        \
        _Pragma("omp parallel for schedule(dynamic, 16)")                     \
        for (i64 qi = q_begin; qi < q_end; qi++) {                            \
            u32* io = idx_out + (qi - q_begin) * k; T* dd = d2_out + (qi - q_begin) * k; \
            int cnt = 0;                                                      \
            for (i64 p = 0; p < n; p++) {                                     \
                T v = dist2_##SUF(P + qi * d, P + p * d, d);                  \
                if (cnt == k && !WORSE(dd[k - 1], (i64)io[k - 1], v, p)) continue; \
                int m = cnt < k ? cnt++ : k - 1;                              \
                while (m > 0 && WORSE(dd[m - 1], (i64)io[m - 1], v, p)) { dd[m] = dd[m - 1]; io[m] = io[m - 1]; m--; } \
                dd[m] = v; io[m] = (u32)p;                                    \
            }                                                                 \
        }                                                                     \
        return 0;                                                             \
    }
DEFINE_KNN(f32, float, fma_T_f)
DEFINE_KNN(f64, double, fma_T_d)

/* ------------------------------------------------------------------------ */
/* _compute_logRho_njit / _compute_weighted_logRho_njit (astrolink.py:293-303)*/
/* log((k - sum_j d2_j / d2_K) / d2_K^(d_int/2)); numba returns float64 even */
/* for float32 input; the caller stores into an array of P's dtype (:282).   */
/* weights == NULL -> unweighted.                                            */
/* ------------------------------------------------------------------------ */
#define DEFINE_LOGRHO(NAME, T)                                                \
    void NAME(const T* d2, const u32* idx, const double* weights, i64 nq, int k, \
              int d_intrinsic, T* out) {                                      \
        for (i64 i = 0; i < nq; i++) {                                        \
            const T* r = d2 + i * k;                                          \
            double core = (double)r[k - 1];                                   \
            double v;                                                         \
            if (!weights) {                                                   \
                T s = (T)0;                                                   \
                for (int j = 0; j < k; j++) s += r[j];                        \
                v = ((double)k - (double)(T)(s / r[k - 1])) / pow(core, d_intrinsic / 2.0); \
            } else {                                                          \
                double sw = 0.0, swd = 0.0;                                   \
                for (int j = 0; j < k; j++) {                                 \
                    double w = weights[idx[i * k + j]];                       \
                    sw += w; swd += w * (double)r[j];                         \
                }                                                             \
                v = (sw - swd / core) / pow(core, d_intrinsic / 2.0);         \
            }                                                                 \
            out[i] = (T)log(v);                                               \
        }                                                                     \
    }
DEFINE_LOGRHO(oracle_logrho_f32, float)
DEFINE_LOGRHO(oracle_logrho_f64, double)

/* _normalise_njit (astrolink.py:305-313): (x - min) / (max - min) */
#define DEFINE_NORMALISE(NAME, T)                                             \
    void NAME(T* x, i64 n) {                                                  \
        T mn = x[0], mx = x[0];                                               \
        for (i64 i = 1; i < n; i++) { if (x[i] < mn) mn = x[i]; else if (x[i] > mx) mx = x[i]; } \
        T range = mx - mn;                                                    \
        for (i64 i = 0; i < n; i++) x[i] = (x[i] - mn) / range;               \
    }
DEFINE_NORMALISE(oracle_normalise_f32, float)
DEFINE_NORMALISE(oracle_normalise_f64, double)

/* ------------------------------------------------------------------------ */
/* _make_graph_njit (astrolink.py:355-382).  Row i owns slots                */
/* [i(L-1), (i+1)(L-1)); entries equal to i are skipped; pair = (denser,     */
/* sparser) with ties -> (i, j); weight = the sparser end's logRho.          */
/* Defined behaviour where the reference overruns (self absent from the row, */
/* SURVEY App. B Q7): the last entry of the row is dropped.                  */
/* ------------------------------------------------------------------------ */
#define DEFINE_MAKE_GRAPH(NAME, T)                                            \
    void NAME(const T* logrho, const u32* knn, i64 n, int L, int knn_stride,  \
              T* edges, u32* pairs) {                                         \
        for (i64 i = 0; i < n; i++) {                                         \
            T li = logrho[i];                                                 \
            i64 pos = i * (L - 1), end = (i + 1) * (L - 1);                   \
            for (int r = 0; r < L && pos < end; r++) {                        \
                u32 j = knn[i * knn_stride + r];                              \
                if ((i64)j == i) continue;                                    \
                T lj = logrho[j];                                             \
                if (li >= lj) { edges[pos] = lj; pairs[2 * pos] = (u32)i; pairs[2 * pos + 1] = j; } \
                else { edges[pos] = li; pairs[2 * pos] = j; pairs[2 * pos + 1] = (u32)i; } \
                pos++;                                                        \
            }                                                                 \
        }                                                                     \
    }
DEFINE_MAKE_GRAPH(oracle_make_graph_f32, float)
DEFINE_MAKE_GRAPH(oracle_make_graph_f64, double)

/* ------------------------------------------------------------------------ */
/* Edge order (astrolink.py:341).  The reference uses an UNSTABLE argsort,   */
/* so its tie order is arbitrary.  Canonical order used by the   */
/* oracle and the GPU path: weight descending, slot index ascending          */
/* (== np.argsort(-edges, kind='stable')).  Bottom-up merge sort, stable.    */
/* ------------------------------------------------------------------------ */
#define DEFINE_SORT(NAME, T)                                                  \
    void NAME(const T* edges, i64 E, i64* order) {                            \
        i64* a = order; i64* b = (i64*)malloc(sizeof(i64) * (E > 0 ? E : 1)); \
        for (i64 i = 0; i < E; i++) a[i] = i;                                 \
        i64* src = a; i64* dst = b;                                           \
        for (i64 w = 1; w < E; w *= 2) {                                      \
            for (i64 lo = 0; lo < E; lo += 2 * w) {                           \
                i64 mid = lo + w < E ? lo + w : E, hi = lo + 2 * w < E ? lo + 2 * w : E; \
                i64 i = lo, j = mid, o = lo;                                  \
                while (i < mid && j < hi) {                                   \
                    if (edges[src[j]] > edges[src[i]]) dst[o++] = src[j++];   \
                    else dst[o++] = src[i++];                                 \
                }                                                             \
                while (i < mid) dst[o++] = src[i++];                          \
                while (j < hi) dst[o++] = src[j++];                           \
            }                                                                 \
            i64* t = src; src = dst; dst = t;                                 \
        }                                                                     \
        if (src != order) memcpy(order, src, sizeof(i64) * E);                \
        free(b);                                                              \
    }
DEFINE_SORT(oracle_sort_edges_f32, float)
DEFINE_SORT(oracle_sort_edges_f64, double)

/* ------------------------------------------------------------------------ */
/* _aggregate_njit_float32_uint32 (astrolink.py:384-493) and                 */
/* _aggregate_njit_float64_uint32 (astrolink.py:495-604).                    */
/* Kruskal-style pass over the ordered pairs with relabel-smaller union,     */
/* ordered member lists, per-group [geq_start, leq_start, size] and          */
/* [prom_geq, prom_leq]; disconnected fix-up (:442-461); ordering (:464);    */
/* top-down finalise with noise correction (:471-485) where the float32      */
/* variant corrects column 0 (:484) and the float64 variant column 1 (:595); */
/* rows with size <= 2 dropped, sorted by leq_start, root dropped (:487-493).*/
/* Outputs: ordering[n]; groups[G][3]; prominences[G][2]; returns G, or <0.  */
/* groups_out / prom_out must hold n/2 rows.                                 */
/* ------------------------------------------------------------------------ */
typedef struct { u32 key; i64 id; } keyid_t;
static int cmp_size_id(const void* a, const void* b) {
    const i64* x = (const i64*)a; const i64* y = (const i64*)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}
static int cmp_keyid(const void* a, const void* b) {
    const keyid_t* x = (const keyid_t*)a; const keyid_t* y = (const keyid_t*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return 0;
}

#define DEFINE_AGGREGATE(NAME, T, QUIRK_COL)                                  \
    i64 NAME(const T* logrho, const u32* ordered_pairs, i64 n, i64 E,         \
             u32* ordering, u32* groups_out, T* prom_out) {                   \
        const i64 NONE = -1;                                                  \
        i64 maxg = n / 2 + 2;                                                 \
        i64* ids = (i64*)malloc(sizeof(i64) * n);      /* group of point, n = none */ \
        i64* nextp = (i64*)malloc(sizeof(i64) * n);    /* member list links */\
        i64* head = (i64*)malloc(sizeof(i64) * maxg);                         \
        i64* tail = (i64*)malloc(sizeof(i64) * maxg);                         \
        i64* ga = (i64*)malloc(sizeof(i64) * maxg);                           \
        i64* gb = (i64*)malloc(sizeof(i64) * maxg);                           \
        i64* gs = (i64*)malloc(sizeof(i64) * maxg);                           \
        T* p0 = (T*)malloc(sizeof(T) * maxg);                                 \
        T* p1 = (T*)malloc(sizeof(T) * maxg);                                 \
        i64* chead = (i64*)malloc(sizeof(i64) * maxg); /* children, merge order */ \
        i64* ctail = (i64*)malloc(sizeof(i64) * maxg);                        \
        i64* cnext = (i64*)malloc(sizeof(i64) * maxg);                        \
        char* alive = (char*)malloc(maxg);                                    \
        i64 count = 0;                                                        \
        for (i64 i = 0; i < n; i++) { ids[i] = n; nextp[i] = NONE; }          \
        for (i64 e = 0; e < E; e++) {                                         \
            i64 a = ordered_pairs[2 * e], b = ordered_pairs[2 * e + 1];       \
            i64 id0 = ids[a], id1 = ids[b];                                   \
            if (id0 != n) {                                                   \
                if (id0 == id1) continue;                                     \
                if (id1 == n) {             /* append sparser point b to a's group */ \
                    ids[b] = id0; nextp[tail[id0]] = b; tail[id0] = b; gs[id0] += 1; \
                } else {                    /* merge the smaller group into the larger */ \
                    if (gs[id0] < gs[id1]) { i64 t = id0; id0 = id1; id1 = t; } \
                    for (i64 p = head[id1]; p != NONE; p = nextp[p]) ids[p] = id0; \
                    nextp[tail[id0]] = head[id1]; tail[id0] = tail[id1];      \
                    T cur = logrho[b];                                        \
                    ga[id1] = gb[id0]; gb[id1] = gs[id0];                     \
                    gs[id0] += gs[id1];                                       \
                    p0[id1] = p1[id0] - cur;                                  \
                    if (p1[id1] > p1[id0]) p1[id0] = p1[id1];                 \
                    p1[id1] -= cur;                                           \
                    if (chead[id0] == NONE) chead[id0] = id1; else cnext[ctail[id0]] = id1; \
                    ctail[id0] = id1; alive[id1] = 0;                         \
                }                                                             \
            } else if (id1 == n) {          /* new group from two free points */ \
                i64 g = count++;                                              \
                ids[a] = g; ids[b] = g;                                       \
                head[g] = a; nextp[a] = b; tail[g] = b;                       \
                ga[g] = 0; gb[g] = 0; gs[g] = 2;                              \
                p0[g] = (T)0; p1[g] = logrho[a];                              \
                chead[g] = NONE; ctail[g] = NONE; cnext[g] = NONE; alive[g] = 1; \
            } else {                        /* append denser point a to b's group */ \
                ids[a] = id1; nextp[tail[id1]] = a; tail[id1] = a; gs[id1] += 1; \
                if (logrho[a] > p1[id1]) p1[id1] = logrho[a];                 \
            }                                                                 \
        }                                                                     \
        for (i64 i = 0; i < n; i++) if (ids[i] == n) { count = -2; }          \
        if (count <= 0) goto fail_##NAME;                                     \
        /* disconnected fix-up: live groups ascending by (size, id); largest  \
           keeps its identity, the rest are appended in descending order */   \
        i64 id_final;                                                         \
        {                                                                     \
            i64 nl = 0;                                                       \
            for (i64 g = 0; g < count; g++) nl += alive[g];                   \
            i64* sl = (i64*)malloc(sizeof(i64) * 2 * nl);                     \
            i64 m = 0;                                                        \
            for (i64 g = 0; g < count; g++) if (alive[g]) { sl[2 * m] = gs[g]; sl[2 * m + 1] = g; m++; } \
            qsort(sl, nl, 2 * sizeof(i64), cmp_size_id);                      \
            id_final = sl[2 * (nl - 1) + 1];                                  \
            for (i64 t = nl - 2; t >= 0; t--) {                               \
                i64 sz = sl[2 * t], g = sl[2 * t + 1];                        \
                nextp[tail[id_final]] = head[g]; tail[id_final] = tail[g];    \
                ga[g] = gb[id_final]; gb[g] = gs[id_final];                   \
                p0[g] = p1[id_final];                                         \
                gs[id_final] += sz;                                           \
                if (chead[id_final] == NONE) chead[id_final] = g; else cnext[ctail[id_final]] = g; \
                ctail[id_final] = g;                                          \
            }                                                                 \
            free(sl);                                                         \
        }                                                                     \
        {                                                                     \
            i64 o = 0;                                                        \
            for (i64 p = head[id_final]; p != NONE; p = nextp[p]) ordering[o++] = (u32)p; \
        }                                                                     \
        /* top-down finalise: absolute starts + noise correction */           \
        {                                                                     \
            i64* stack = (i64*)malloc(sizeof(i64) * (count + 1));             \
            i64 sp = 0; stack[sp++] = id_final;                               \
            while (sp > 0) {                                                  \
                i64 g = stack[--sp];                                          \
                if (chead[g] == NONE) continue;                               \
                i64 adj = gb[g];                                              \
                double noise = 0.0; i64 t = 0;                                \
                for (i64 c = chead[g]; c != NONE; c = cnext[c], t++) {        \
                    stack[sp++] = c;                                          \
                    ga[c] += adj; gb[c] += adj;                               \
                    if (t > 0) p0[c] = (T)((double)p0[c] - sqrt(noise / (double)t)); \
                    noise += (double)(T)(p1[c] * p1[c]);                      \
                }                                                             \
                if (QUIRK_COL == 0) p0[g] = (T)((double)p0[g] - sqrt(noise / (double)t)); \
                else p1[g] = (T)((double)p1[g] - sqrt(noise / (double)t));    \
            }                                                                 \
            free(stack);                                                      \
        }                                                                     \
        /* keep size > 2, rows -> [geq_start, leq_start, leq_end], sort by    \
           leq_start, drop the first (root) */                                \
        i64 G;                                                                \
        {                                                                     \
            i64 nk = 0;                                                       \
            for (i64 g = 0; g < count; g++) nk += gs[g] > 2;                  \
            keyid_t* ks = (keyid_t*)malloc(sizeof(keyid_t) * (nk > 0 ? nk : 1)); \
            i64 m = 0;                                                        \
            for (i64 g = 0; g < count; g++) if (gs[g] > 2) { ks[m].key = (u32)gb[g]; ks[m].id = g; m++; } \
            qsort(ks, nk, sizeof(keyid_t), cmp_keyid);                        \
            G = nk > 0 ? nk - 1 : 0;                                          \
            for (i64 r = 0; r < G; r++) {                                     \
                i64 g = ks[r + 1].id;                                         \
                groups_out[3 * r] = (u32)ga[g]; groups_out[3 * r + 1] = (u32)gb[g]; \
                groups_out[3 * r + 2] = (u32)(gb[g] + gs[g]);                 \
                prom_out[2 * r] = p0[g]; prom_out[2 * r + 1] = p1[g];         \
            }                                                                 \
            free(ks);                                                         \
        }                                                                     \
        free(ids); free(nextp); free(head); free(tail); free(ga); free(gb); free(gs); \
        free(p0); free(p1); free(chead); free(ctail); free(cnext); free(alive); \
        return G;                                                             \
    fail_##NAME:                                                              \
        free(ids); free(nextp); free(head); free(tail); free(ga); free(gb); free(gs); \
        free(p0); free(p1); free(chead); free(ctail); free(cnext); free(alive); \
        return -1;                                                            \
    }
DEFINE_AGGREGATE(oracle_aggregate_f32, float, 0)
DEFINE_AGGREGATE(oracle_aggregate_f64, double, 1)
