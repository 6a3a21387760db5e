// f16_peaks.cu -- pipe-rate probes: the measured denominators of the per-kernel rooflines that are not in
// MEASURED_PEAKS.json (which holds the HBM copy rate and the cuBLAS bf16 rate only): FP32 FMA, FP64 FMA and MUFU.TANH
// results per second on THIS board at THIS moment's clocks.  bench.py times one launch of each with CUDA events next to
// the kernels it reports on (policy / rollout: MUFU.TANH; trim search: FP64; K-step env kernel: FP32 issue).
// Persistent grid of n_sm x 8 CTAs x 256 threads, 8 independent dependency chains per thread, `iters` trips.
#include <cuda_runtime.h>

#include <cstdint>

#include "f16_launch.h"

namespace f16 {
namespace {

template <int KIND>
__global__ void __launch_bounds__(256) peak_probe_kernel(float *sink, int iters)
{
    float x[8];
    double d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        x[j] = threadIdx.x * 1e-3f + j * 0.1f;
        d[j] = threadIdx.x * 1e-3 + j * 0.1;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0) x[j] = __fmaf_rn(x[j], 0.999f, 0.001f);
                if (KIND == 1) d[j] = __fma_rn(d[j], 0.999, 0.001);
                if (KIND == 2) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
            }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] + static_cast<float>(d[j]);
    if (s == 123.456f) sink[0] = s;   // keeps the chains alive, never true in practice
}

}  // namespace

// ops per launch = n_sm * 8 * 256 threads * 32 ops * iters (an FMA counts as ONE op here; the caller doubles it for flops)
cudaError_t launch_peak_probe(int kind, int iters, int n_sm, float *sink, double *ops, cudaStream_t st)
{
    const dim3 grid(n_sm * 8), block(256);
    if (kind == 0) peak_probe_kernel<0><<<grid, block, 0, st>>>(sink, iters);
    else if (kind == 1) peak_probe_kernel<1><<<grid, block, 0, st>>>(sink, iters);
    else peak_probe_kernel<2><<<grid, block, 0, st>>>(sink, iters);
    if (ops) *ops = static_cast<double>(n_sm) * 8.0 * 256.0 * 32.0 * static_cast<double>(iters);
    return cudaGetLastError();
}

}  // namespace f16
