// One translation unit per (dtype, d): compiled with -DKNN_T=float|double
// -DKNN_TNAME=f32|f64 -DKNN_D=<1..8> so the instantiations build in parallel.
#include "knn_kernel.cuh"

#define KNN_CAT2(a, b, c) a##b##_d##c
#define KNN_CAT(a, b, c) KNN_CAT2(a, b, c)

int KNN_CAT(knn_launch_, KNN_TNAME, KNN_D)(const KnnParams& p, cudaStream_t st) { return launch_knn<KNN_T, KNN_D>(p, st); }
