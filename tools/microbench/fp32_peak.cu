// FP32 issue-rate microbenchmark for B200 (sm_100a): scalar FFMA, packed f32x2
// FMA, and the kNN-style SUB+FMA mix. Prints achieved TFLOP/s and the SM clock
// estimated from clock64(). Used to fix the FP32 roofline denominator
// (MEASURED_PEAKS.json has no FP32 CUDA-core figure).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__global__ void k_ffma(float* out, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long x, unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long x, unsigned long long a) {
    unsigned long long r;
    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(a));
    return r;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x[ILP];
    unsigned long long aa = pack(a, a), bb = pack(b, b);
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = pack(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma2(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += __uint_as_float((unsigned)x[i]) + __uint_as_float((unsigned)(x[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// kNN-like: per step diff = q - p ; acc = fma(diff, diff, acc)   (scalar)
__global__ void k_submix(float* out, float q, float p0) {
    float acc[ILP], p[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = 0.f; p[i] = p0 + threadIdx.x * 1e-3f + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) { float d = q - p[i]; acc[i] = fmaf(d, d, acc[i]); p[i] = acc[i] * 1e-9f; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// same, 2-wide: FADD+FFMA only (no extra mul): p fixed
__global__ void k_submix_pure(float* out, float q, float p0) {
    float acc[ILP], p[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = 0.f; p[i] = p0 + threadIdx.x * 1e-3f + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) { float d = q - p[i]; acc[i] = fmaf(d, d, acc[i]); }
        q += 1e-7f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_submix2(float* out, float q, float p0) {
    unsigned long long acc[ILP], p[ILP];
    unsigned long long qq = pack(q, q);
    unsigned long long inc = pack(1e-7f, 1e-7f), one = pack(1.f, 1.f);
#pragma unroll
    for (int i = 0; i < ILP; i++) { acc[i] = pack(0.f, 0.f); p[i] = pack(p0 + threadIdx.x * 1e-3f + i, p0 + threadIdx.x * 2e-3f + i); }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) { unsigned long long d = sub2(qq, p[i]); acc[i] = fma2(d, d, acc[i]); }
        qq = fma2(inc, one, qq);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) f();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms %d clockRate %d kHz\n", prop.name, sms, prop.clockRate);
    int blocks = sms * 8, threads = 256;
    float* out; CK(cudaMalloc(&out, (size_t)blocks * threads * 4));
    double nthr = (double)blocks * threads;
    for (int round = 0; round < 3; round++) {
        float ms;
        ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 20);
        printf("FFMA scalar      : %.3f ms  %.2f TFLOP/s (2 flop/FMA)\n", ms, nthr * ITERS * ILP * 2 / ms / 1e9);
        ms = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 20);
        printf("FFMA2 packed     : %.3f ms  %.2f TFLOP/s\n", ms, nthr * ITERS * ILP * 4 / ms / 1e9);
        ms = timeit([&] { k_submix_pure<<<blocks, threads>>>(out, 0.3f, 0.1f); }, 20);
        printf("SUB+FMA scalar   : %.3f ms  %.2f TFLOP/s (3 flop/pair-dim)  %.2f Tinstr/s\n", ms, nthr * ITERS * ILP * 3 / ms / 1e9, nthr * ITERS * ILP * 2 / ms / 1e9);
        ms = timeit([&] { k_submix2<<<blocks, threads>>>(out, 0.3f, 0.1f); }, 20);
        printf("SUB2+FMA2 packed : %.3f ms  %.2f TFLOP/s (3 flop/pair-dim)  %.2f Tinstr/s\n", ms, nthr * ITERS * ILP * 6 / ms / 1e9, nthr * ITERS * ILP * 2 / ms / 1e9);
    }
    return 0;
}
