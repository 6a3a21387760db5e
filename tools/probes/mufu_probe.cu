// Micro-probe: throughput of MUFU.TANH (fp32), MUFU.TANH.F16 (scalar half) and tanh.approx.f16x2 (which ptxas lowers
// to TWO MUFU.TANH.F16 + PRMT on sm_100a -- see profiles/r02_sass_markers.txt) per SM, alone and next to FFMA work.
// Answers: can the policy kernel's 256 tanh per env get under the 16 MUFU/clk/SM wall by going to half precision?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mufu_probe tools/probes/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(float *out, int iters)
{
    float x[8];
    unsigned h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        x[j] = threadIdx.x * 0.001f + j * 0.1f;
        h[j] = 0x38003400u + threadIdx.x + j;
    }
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.5f + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0 || MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
            if (MODE == 1) {
                unsigned short s = static_cast<unsigned short>(h[j]);
                asm volatile("tanh.approx.f16 %0, %0;" : "+h"(s));
                h[j] = s;
            }
            if (MODE == 2) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[j]));
            if (MODE == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
        }
        if (MODE == 3) {   // 8 MUFU + 32 FFMA per iteration: does the FMA pipe run beside a saturated MUFU pipe?
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = __fmaf_rn(f[j], 0.999f, 0.001f);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + f[j] + __uint_as_float(h[j]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(float *d, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(d, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(d, iters);
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
    const int iters = 4000;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ops = 148.0 * 8 * 256 * 8.0 * iters;   // PTX-level tanh instructions per launch
    const float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters), t4 = run<4>(d, iters);
    const double per = 1e3 / 148.0 / (clk_khz * 1e3);  // -> per clock per SM at the nominal max clock
    printf("max clock %.0f MHz (rates below assume it; the real clock under load is lower)\n", clk_khz * 1e-3);
    printf("tanh.approx.f32   : %.3f ms  %.2f results/clk/SM\n", t0, ops / t0 * per);
    printf("tanh.approx.f16   : %.3f ms  %.2f results/clk/SM\n", t1, ops / t1 * per);
    printf("tanh.approx.f16x2 : %.3f ms  %.2f results/clk/SM (two per instruction)\n", t2, 2 * ops / t2 * per);
    printf("tanh.f32 + 4 FFMA : %.3f ms  %.2f tanh/clk/SM, %.1f FFMA/clk/SM\n", t3, ops / t3 * per, 4 * ops / t3 * per);
    printf("ex2.approx.f32    : %.3f ms  %.2f results/clk/SM\n", t4, ops / t4 * per);
    return cudaGetLastError() != cudaSuccess;
}
