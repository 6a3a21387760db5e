// f16_policy.cu -- the PPO agent's MLP forward on 5th-generation tensor cores (tcgen05 + TMEM).
//
// Replaces `Agent.get_action_and_value` / `get_value` of the reference
// (control/ppo/ppo_model_gsde.py:37-66,81-99): actor_mean and critic are two tanh MLPs
// 5 -> 64 -> 64 -> 1 evaluated on the observation batch.  This is the one dense contraction on
// the hot path, so it runs on the tensor cores:
//
//   tile = 128 envs = one CTA of 256 threads; threads t and t + 128 share env row t (= TMEM lane t): one
//   handles the actor columns, the other the critic columns.
//   layer 1:  D1[128 x 128] = X[128 x 8] * W1cat^T      one tcgen05.mma (M=128, N=128, K=8, kind::tf32);
//             W1cat stacks actor (rows 0-63) and critic (rows 64-127), K is 5 padded to 8
//   layer 2:  D2[:, 0:64]   = H1[:, 0:64]   * W2actor^T  eight tcgen05.mma (K=64 in steps of 8)
//             D2[:, 64:128] = H1[:, 64:128] * W2critic^T eight more
//   layer 3:  64-term dot products in the epilogue (CUDA cores), fused with the Gaussian policy
//             sample / log-probability.
//
// Operands are fp32 in shared memory, consumed as TF32 (fp32 accumulate in TMEM); both are
// K-major in the canonical no-swizzle layout: element (row r, 16-byte K-chunk c) lives at
// c * LBO + (r / 8) * 128 + (r % 8) * 16 bytes, LBO = rows * 16.  Thread t writes its own row
// chunk by chunk (consecutive threads -> consecutive 16 B: conflict-free).  Accumulators are read
// back with tcgen05.ld (32 lanes x 32 columns per warp), tanh via MUFU.TANH.  The packed weight
// blob (38 KB) is staged once per CTA with a TMA bulk copy; the grid is persistent over tiles.
#include <cuda_runtime.h>

#include <cstdint>

#include "f16_policy.h"

namespace f16 {
namespace {

constexpr int kRows = 128;          // envs per tile == UMMA M
constexpr int kThreads = 256;       // two threads per env row (actor half / critic half)
constexpr int kHid = 64;
constexpr uint32_t kTmemCols = 256;  // D1: columns [0,128), D2: columns [128,256)

struct alignas(16) PolicySmem {
    PolicyBlob w;                    // staged weights (canonical UMMA layouts inside)
    float a1[2][kRows][4];           // X  : 2 K-chunks  x 128 rows x 4 floats
    float a2[32][kRows][4];          // H1 : 32 K-chunks x 128 rows x 4 floats (actor 0-15 | critic 16-31)
    uint64_t bar_w;                  // weights landed
    uint64_t bar_mma;                // tcgen05.commit target
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, sm_100):
// [0,14) start>>4, [16,30) LBO>>4 (between the 16-byte K-chunks), [32,46) SBO>>4 (between 8-row
// groups), [46,48) version = 1, [61,64) layout type = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
           (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
// both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float fast_tanh(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Philox4x32-10 (same generator as the env kernels), used for the policy's exploration noise
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

__global__ void __launch_bounds__(kThreads, 2) policy_forward_kernel(const PolicyArgs P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PolicySmem *sm = reinterpret_cast<PolicySmem *>(smem_raw);
    const unsigned t = threadIdx.x;
    const unsigned warp = t >> 5;

    if (t == 0) {
        mbar_init(&sm->bar_w, 1);
        mbar_init(&sm->bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // one warp allocates the CTA's TMEM columns and publishes the base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm->tmem_base;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (t == 0) {   // weights: one TMA bulk copy per CTA (no kernel writes them: safe ahead of the grid dependency)
        constexpr uint32_t bytes = static_cast<uint32_t>(sizeof(PolicyBlob));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->bar_w)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm->w)),
                     "l"(P.blob), "r"(bytes), "r"(smem_u32(&sm->bar_w))
                     : "memory");
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the observations and env counters come from the previous kernel
    mbar_wait(&sm->bar_w, 0);

    constexpr uint32_t idesc1 = umma_idesc_tf32(128, 128);
    constexpr uint32_t idesc2 = umma_idesc_tf32(128, 64);
    const uint32_t a1 = smem_u32(sm->a1), a2 = smem_u32(sm->a2);
    const uint32_t w1 = smem_u32(sm->w.w1), w2a = smem_u32(sm->w.w2[0]), w2c = smem_u32(sm->w.w2[1]);
    const uint32_t lane_base = tmem + (((warp & 3u) * 32u) << 16);   // this warp's 32 TMEM lanes (quarter = warp % 4)
    uint32_t phase = 0;

    const unsigned n = static_cast<unsigned>(P.n);
    const unsigned n_tiles = (n + kRows - 1) / kRows;
    const float std_dev = P.logstd ? __expf(P.logstd[0]) : 0.f;
    const float log_std = P.logstd ? P.logstd[0] : 0.f;

    // Two threads per env row: thread (row, half) with half 0 = actor columns, half 1 = critic columns.  That
    // halves every thread's epilogue and doubles the resident warps (16 per SM), which is what this
    // latency-bound kernel needs; TMEM lanes are reachable from warps w and w + 4 alike (lane quarter = w % 4).
    const unsigned row = t & (kRows - 1), half = t >> 7;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned i = tile * kRows + row;
        const bool live = i < n;
        // ---- layer-1 operand: the observation row, K padded 5 -> 8 (half 0 writes chunk 0, half 1 chunk 1) ----
        if (half == 0) {
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float *src = P.obs + static_cast<size_t>(i) * 5;
                c0 = make_float4(src[0], src[1], src[2], src[3]);
            }
            *reinterpret_cast<float4 *>(sm->a1[0][row]) = c0;
        } else {
            *reinterpret_cast<float4 *>(sm->a1[1][row]) = make_float4(live ? P.obs[static_cast<size_t>(i) * 5 + 4] : 0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_tf32(tmem + 0, umma_desc(a1, kRows * 16, 128), umma_desc(w1, 128 * 16, 128), idesc1, 0u);
            umma_commit(&sm->bar_mma);
        }
        mbar_wait(&sm->bar_mma, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- hidden layer 1: tanh(D1 + b1) -> layer-2 operand (same canonical layout) ----------
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            float v[32];
            const int col0 = half * 64 + blk * 32;
            tmem_ld32(lane_base + col0, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = col0 + c * 4;
                float4 h;
                h.x = fast_tanh(v[c * 4 + 0] + sm->w.b1[j + 0]);
                h.y = fast_tanh(v[c * 4 + 1] + sm->w.b1[j + 1]);
                h.z = fast_tanh(v[c * 4 + 2] + sm->w.b1[j + 2]);
                h.w = fast_tanh(v[c * 4 + 3] + sm->w.b1[j + 3]);
                *reinterpret_cast<float4 *>(sm->a2[j >> 2][row]) = h;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // actor: K-chunks 0..15, two per MMA
                umma_tf32(tmem + 128, umma_desc(a2 + k * 2 * kRows * 16, kRows * 16, 128),
                          umma_desc(w2a + k * 2 * kHid * 16, kHid * 16, 128), idesc2, k > 0 ? 1u : 0u);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // critic: K-chunks 16..31
                umma_tf32(tmem + 192, umma_desc(a2 + (16 + k * 2) * kRows * 16, kRows * 16, 128),
                          umma_desc(w2c + k * 2 * kHid * 16, kHid * 16, 128), idesc2, k > 0 ? 1u : 0u);
            }
            umma_commit(&sm->bar_mma);
        }
        mbar_wait(&sm->bar_mma, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- hidden layer 2 + output layer: half 0 -> actor mean, half 1 -> critic value ----------------
        float head = sm->w.b3[half];
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            float v[32];
            const int col0 = half * 64 + blk * 32;
            tmem_ld32(lane_base + 128 + col0, v);
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) acc = fmaf(fast_tanh(v[c] + sm->w.b2[col0 + c]), sm->w.w3[col0 + c], acc);
            head += acc;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads done before the next tile's MMAs

        if (live && half == 1) {
            if (P.value) P.value[i] = head;
        } else if (live) {
            const float mean = head;
            if (P.mean) P.mean[i] = mean;
            if (P.action) {
                // Gaussian policy head (ppo_model_gsde.py:81-99 with the state-independent log-std):
                // a = mean + std * eps,  logp = -0.5 eps^2 - log std - 0.5 log(2 pi)
                float eps = 0.f;
                if (P.sample) {
                    // counter = (global env id, episode, step): with the env's own counters the stream is unique per
                    // env-step without any host-side step argument (CUDA-graph replays stay fresh)
                    const int64_t gid = P.id_base + i;
                    const uint32_t c2 = P.episode_no ? P.episode_no[i] : 0u;
                    const uint32_t c3 = P.ep_len ? static_cast<uint32_t>(P.ep_len[i]) : P.step;
                    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32) ^ 0x50504F31u, c2, c3),
                                               make_uint2(static_cast<uint32_t>(P.seed), static_cast<uint32_t>(P.seed >> 32)));
                    const float u1 = (static_cast<float>(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
                    const float u2 = static_cast<float>(r.y >> 8) * (1.0f / 16777216.0f);
                    eps = sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
                }
                P.action[i] = mean + std_dev * eps;
                if (P.logprob) P.logprob[i] = -0.5f * eps * eps - log_std - 0.91893853320467274178f;
            }
        }
    }

    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
}

// Generalised advantage estimation, the reverse scan of ppo_model_gsde.py:244-268 (gae=True):
//   delta_t = r_t + gamma * V_{t+1} * (1 - d_{t+1}) - V_t ;  A_t = delta_t + gamma * lambda * (1 - d_{t+1}) * A_{t+1}
// where d_{t+1} = done flag returned BY step t (the reference's dones[t+1] / next_done) = done_after[t],
// and V_T = the bootstrap value after the last step.  One thread per env walks its T steps backwards; every
// access of a warp is one contiguous row segment of the [T][n] buffers.
__global__ void __launch_bounds__(256) gae_kernel(int T, int64_t n, const float *__restrict__ rewards, const float *__restrict__ values,
                                                  const uint8_t *__restrict__ done_after, const float *__restrict__ next_value,
                                                  float gamma, float lam, float *__restrict__ advantages, float *__restrict__ returns)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float nextvalues = next_value[i];
        float lastgaelam = 0.f;
        for (int t = T - 1; t >= 0; --t) {
            const int64_t k = (int64_t)t * n + i;
            const float nextnonterminal = done_after[k] ? 0.f : 1.f;
            const float v = values[k];
            const float delta = rewards[k] + gamma * nextvalues * nextnonterminal - v;
            lastgaelam = delta + gamma * lam * nextnonterminal * lastgaelam;
            advantages[k] = lastgaelam;
            returns[k] = lastgaelam + v;
            nextvalues = v;
        }
    }
}

}  // namespace

cudaError_t launch_gae(int T, int64_t n, const float *rewards, const float *values, const uint8_t *done_after, const float *next_value,
                       float gamma, float lam, float *advantages, float *returns, int n_sm, cudaStream_t st)
{
    const int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)n_sm * 8;
    gae_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(T, n, rewards, values, done_after, next_value, gamma, lam, advantages,
                                                                   returns);
    return cudaGetLastError();
}

size_t policy_smem_bytes() { return sizeof(PolicySmem); }

cudaError_t launch_policy_forward(const PolicyArgs &P, int n_sm, int pdl, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(policy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(sizeof(PolicySmem)));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (P.n + kRows - 1) / kRows;
    const int64_t cap = static_cast<int64_t>(n_sm) * 2;   // 2 CTAs per SM: 2 x 256 TMEM columns, 2 x 106 KB shared memory
    // programmatic dependent launch: TMEM allocation, barrier init and the weight copy overlap the tail of the
    // env-step kernel that precedes this one in a rollout
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(static_cast<unsigned>(tiles < cap ? tiles : cap));
    lc.blockDim = dim3(kThreads);
    lc.dynamicSmemBytes = sizeof(PolicySmem);
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&lc, policy_forward_kernel, P);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace f16
