// Micro-probe: tcgen05.ld (32x32b.x32: 4 KB per warp instruction) throughput for 1, 4, 8, 16 warps of ONE CTA per SM
// (warp w reads TMEM lanes 32 (w % 4) .. +31, as the hardware requires), with and without 32 MUFU.TANH per load.
// Question (r02b): is the rollout kernel's epilogue bound by the TMEM read port rather than by MUFU?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_ld_probe tools/probes/tmem_ld_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

template <int TANH>
__global__ void probe(float *out, int iters, long long *clk)
{
    __shared__ uint32_t tmem_base;
    const unsigned warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 32u * 4u % 512u;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        tmem_ld32(base + (i & 3) * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (TANH) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float y;
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(r[j])));
                acc += y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) *clk = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
}

int main()
{
    float *d;
    long long *clk, h;
    cudaMalloc(&d, 148 * 1024 * 4);
    cudaMalloc(&clk, 8);
    printf("warps/CTA   clk per LDTM.x32 (+wait)   ... with 32 MUFU.TANH per load   (32 tanh alone = 256 clk of one sub-partition's MUFU)\n");
    const int ws[] = {1, 4, 8, 16};
    for (int w : ws) {
        const int iters = 20000;
        double r[2];
        for (int m = 0; m < 2; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m) probe<1><<<148, 32 * w>>>(d, iters, clk); else probe<0><<<148, 32 * w>>>(d, iters, clk);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
            r[m] = (double)h / iters;
        }
        printf("%5d       %8.1f                    %8.1f\n", w, r[0], r[1]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
