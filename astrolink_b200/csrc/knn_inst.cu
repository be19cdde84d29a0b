// One translation unit per (dtype, d): compiled with -DKNN_T=float|double
// -DKNN_TNAME=f32|f64 -DKNN_D=<1..16> so the instantiations build in parallel.
// float32 indices are searched by the specialised kernel of knn_fast.cuh, float64 ones by the generic kernel.
#include "knn_fast.cuh"

#define KNN_CAT2(a, b, c) a##b##_d##c
#define KNN_CAT(a, b, c) KNN_CAT2(a, b, c)

template <typename T, int D>
struct KnnDispatch {
    static int run(const KnnParams& p, cudaStream_t st) { return launch_knn<T, D>(p, st); }
};
template <int D>
struct KnnDispatch<float, D> {
    static int run(const KnnParams& p, cudaStream_t st) { return p.use_generic ? launch_knn<float, D>(p, st) : kf::launch<D>(p, st); }
};

int KNN_CAT(knn_launch_, KNN_TNAME, KNN_D)(const KnnParams& p, cudaStream_t st) { return KnnDispatch<KNN_T, KNN_D>::run(p, st); }
