// Micro-probe: which pipe does F2FP.F16.F32.PACK_AB (cvt.rn.f16x2.f32) use?  Times, per SM with 4 warps per sub-partition:
//   (0) 8 tanh                     (1) 4 F2FP                  (2) 8 tanh + 4 F2FP (the rollout kernel's H1 epilogue mix)
//   (3) 8 tanh + 4 packs done with integer ops on the ALU pipe (round-to-nearest-even fp32 -> fp16 by hand)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/f2fp_probe tools/probes/f2fp_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float *out, int iters)
{
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 0.001f + j * 0.1f;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 1) y[j] = x[j];
            else asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y[j]) : "f"(x[j]));
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t p;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(y[2 * j + 1]), "f"(y[2 * j]));
                acc ^= p;
            }
        }
        if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // |y| <= 1: scale so that the fp32 exponent field lines up with fp16's, round to nearest even on bit 13, shift
                const uint32_t a = __float_as_uint(y[2 * j] * 1.92592994e-34f), b = __float_as_uint(y[2 * j + 1] * 1.92592994e-34f);   // 2^-112
                const uint32_t ha = ((a & 0x7fffffffu) + 0xfffu + ((a >> 13) & 1u)) >> 13, hb = ((b & 0x7fffffffu) + 0xfffu + ((b >> 13) & 1u)) >> 13;
                acc ^= (ha | ((a >> 16) & 0x8000u)) | ((hb | ((b >> 16) & 0x8000u)) << 16);
            }
        }
        if (MODE == 0) acc ^= __float_as_uint(y[0] + y[1] + y[2] + y[3] + y[4] + y[5] + y[6] + y[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float((__float_as_uint(x[j]) ^ (acc & 1u)));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
}

template <int MODE>
float run(float *d, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t smem = 120 * 1024;
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<MODE><<<148, 512, smem>>>(d, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE><<<148, 512, smem>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    float *d;
    cudaMalloc(&d, 148 * 512 * 4);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    const double c = clk_khz * 1e3 * 1e-3 / iters / 4.0;   // clocks per iteration per warp-slot (4 warps per sub-partition)
    printf("clocks per loop iteration per warp (4 warps per sub-partition share the pipes; 8 MUFU alone = 64):\n");
    printf("8 tanh                 : %.1f\n", run<0>(d, iters) * c);
    printf("4 F2FP                 : %.1f\n", run<1>(d, iters) * c);
    printf("8 tanh + 4 F2FP        : %.1f\n", run<2>(d, iters) * c);
    printf("8 tanh + 4 integer packs: %.1f\n", run<3>(d, iters) * c);
    return cudaGetLastError() != cudaSuccess;
}
