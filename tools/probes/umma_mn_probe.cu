// Probe: does tcgen05.mma accept MN-major, no-swizzle shared-memory operands?  kind::f16 (fp16 operands) and kind::tf32.
// One CTA, M = 128, N = 128, one K step; operands hold small integers so the expected D is exact.
// MN-major canonical no-swizzle layout: 16-byte units of T consecutive MN elements (T = 8 for fp16, 4 for tf32), the
// 8 k's of a core matrix 16 bytes apart, k-groups LBO apart, MN units SBO apart.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_mn_probe tools/probes/umma_mn_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a, uint32_t lbo, uint32_t sbo)
{
    return static_cast<uint64_t>((a & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}
constexpr uint32_t kSbo = 2048 + 16, kLbo = 128;

template <bool F16>
__global__ void __launch_bounds__(128) probe(float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *A = smem, *B = smem + 40 * 1024;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const unsigned t = threadIdx.x;
    constexpr int T = F16 ? 8 : 4, K = F16 ? 16 : 8, ES = F16 ? 2 : 4;
    for (int idx = t; idx < 128 * K; idx += 128) {
        const int mn = idx / K, k = idx % K;
        const float a = static_cast<float>((mn * 3 + k * 7) % 11 - 5), b = static_cast<float>((mn * 5 + k * 3) % 13 - 6);
        const uint32_t off = (mn / T) * kSbo + (k / 8) * kLbo + (k % 8) * 16 + (mn % T) * ES;
        if (F16) {
            *reinterpret_cast<__half *>(A + off) = __float2half(a);
            *reinterpret_cast<__half *>(B + off) = __float2half(b);
        } else {
            *reinterpret_cast<float *>(A + off) = a;
            *reinterpret_cast<float *>(B + off) = b;
        }
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (t == 0) {
        const uint32_t fmt = F16 ? 0u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = umma_desc(smem_u32(A), kLbo, kSbo), db = umma_desc(smem_u32(B), kLbo, kSbo);
        if (F16)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                         "l"(db), "r"(idesc), "r"(0u)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
                         "l"(db), "r"(idesc), "r"(0u)
                         : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\nW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\nD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = tmem + (((t >> 5) * 32u) << 16);
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
            "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(lane_base + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 32; ++c) out[t * 128 + c0 + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128));
}

template <bool F16>
int run(const char *name)
{
    float *d, *h = new float[128 * 128];
    cudaMalloc(&d, 128 * 128 * 4);
    cudaMemset(d, 0xff, 128 * 128 * 4);
    cudaFuncSetAttribute(probe<F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    probe<F16><<<1, 128, 80 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    const int K = F16 ? 16 : 8;
    int bad = 0, zeros = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            float ref = 0;
            for (int k = 0; k < K; ++k) ref += static_cast<float>((m * 3 + k * 7) % 11 - 5) * static_cast<float>((n * 5 + k * 3) % 13 - 6);
            bad += h[m * 128 + n] != ref;
            zeros += h[m * 128 + n] == 0.f;
        }
    printf("%s MN-major no-swizzle: %s, mismatches %d / 16384, zeros %d, D[1][2] = %g\n", name, cudaGetErrorString(e), bad, zeros, h[130]);
    return bad;
}

int main()
{
    run<true>("kind::f16 ");
    run<false>("kind::tf32");
    return 0;
}
