// Micro-probe: issue rate of packed FFMA2 (fma.rn.f32x2) against scalar FFMA on sm_100a, alone and mixed with
// integer work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe tools/probes/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters, float a, float b)
{
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = threadIdx.x * 0.001f + j;
    unsigned z = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = __fmaf_rn(x[j], a, b);
        } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 v = __ffma2_rn(make_float2(x[j], x[j + 1]), make_float2(a, a), make_float2(b, b));
                x[j] = v.x;
                x[j + 1] = v.y;
            }
        }
        if (MODE >= 2) {   // 16 dependent-free integer ops per iteration next to the FP work
#pragma unroll
            for (int j = 0; j < 16; ++j) z = (z ^ (z >> 3)) + 0x9e3779b9u * j;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)z;
}

template <int MODE>
float run(float *d, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const double fma = 148.0 * 8 * 256 * 16.0 * iters;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    printf("scalar FFMA      : %.3f ms  %.1f TFLOP/s\n", t0, 2 * fma / t0 * 1e-9);
    printf("packed FFMA2     : %.3f ms  %.1f TFLOP/s\n", t1, 2 * fma / t1 * 1e-9);
    printf("scalar FFMA + int: %.3f ms\n", t2);
    printf("packed FFMA2+ int: %.3f ms\n", t3);
    return cudaGetLastError() != cudaSuccess;
}
