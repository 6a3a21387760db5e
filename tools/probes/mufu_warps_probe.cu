// Micro-probe: MUFU.TANH throughput per SM as a function of the number of warps per SM sub-partition (1..16) and of
// the number of independent tanh chains per thread (8 or 32).  One CTA per SM (large dynamic shared memory), 148 CTAs.
// Question behind it (r02b): the rollout kernel has 4 warps per sub-partition, of which 1..3 are in a tanh phase at a
// time -- how much of the 16 results/clk/SM can so few warps draw?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mufu_warps_probe tools/probes/mufu_warps_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void probe(float *out, int iters)
{
    float x[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) x[j] = threadIdx.x * 0.001f + j * 0.1f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < CH; ++j) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
float run(float *d, int warps_per_smsp, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int threads = 128 * warps_per_smsp;
    const size_t smem = 120 * 1024;
    cudaFuncSetAttribute(probe<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<CH><<<148, threads, smem>>>(d, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<CH><<<148, threads, smem>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    float *d;
    cudaMalloc(&d, 148 * 1024 * 4);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("max clock %.0f MHz; tanh results per clock per SM (nominal clock)\nwarps/SMSP   8 chains   32 chains\n", clk_khz * 1e-3);
    const int ws[] = {1, 2, 3, 4, 6, 8};
    for (int w : ws) {
        const int it8 = 40000, it32 = 10000;
        const float t8 = run<8>(d, w, it8), t32 = run<32>(d, w, it32);
        const double ops = 128.0 * w * 8.0 * it8;   // per SM
        printf("%5d      %8.2f   %8.2f\n", w, ops / (t8 * 1e-3) / (clk_khz * 1e3), ops / (t32 * 1e-3) / (clk_khz * 1e3));
    }
    return cudaGetLastError() != cudaSuccess;
}
