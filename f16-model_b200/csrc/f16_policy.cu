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
#include <cstdlib>

#include "f16_policy.h"
#include "f16_umma.cuh"

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

// same load without the wait: issue several, then tmem_wait_ld() once
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
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
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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


// ---------------------------------------------------------------------------
// Fused PPO minibatch gradient: one launch = forward of actor_mean and critic on the minibatch rows, the clipped-PPO
// loss of Agent.train (control/ppo/ppo_model_gsde.py:303-364: ratio clipping, clipped value loss, entropy bonus,
// advantage normalisation), and the backward pass of both MLPs -- nothing but the 9.3 k gradient floats leaves the SM.
//
//   tile = 128 minibatch rows, one CTA of 256 threads: thread (row, half), half 0 = actor, 1 = critic.
//   tensor cores (tcgen05, TF32, fp32 accumulate in TMEM):
//       D1 = X  * W1cat^T                 (layer 1, as in the rollout kernel)
//       D2 = H1 * W2^T       per net      (layer 2)
//       D3 = dZ2 * W2        per net      (backward data: B operand = W2^T, K-major over the OUTPUT features)
//   CUDA cores: tanh, heads, loss and its derivative per row, dZ2 = dy w3 (1 - H2^2), dZ1 = D3 (1 - H1^2), and the
//   weight gradients -- dW2 += dZ2^T H1 as register-tiled outer products (thread = 4 x 8 patch of one net's 64 x 64,
//   rows streamed from the shared-memory operands the MMAs already use; it runs while D3 is in flight), dW1 likewise,
//   dW3 / biases as per-thread partial sums reduced through shared memory once per CTA.  Gradients are accumulated over
//   the CTA's tiles in registers and added to global memory with one atomic per element per CTA.
//   H1 and dZ2 are kept in fp32 in the canonical K-major layout with a padded chunk stride (LBO = 2048 + 16 bytes), so the
//   strided reads of the outer-product loop do not collide in the banks.
//   Tried and measured: the same [4-feature chunk][row][4 floats] buffers are, byte for byte, the canonical MN-major
//   no-swizzle operands of dW = dZ^T * H (16-byte units SBO apart, 8 rows 16 bytes apart, LBO = 128), which would put the
//   weight gradients on the tensor cores for free (kernel 223 -> 134 us).  With the MN-major bits set in the instruction
//   descriptor, kind::tf32 MMAs on such operands write zeros: MN-major TF32 operands are only defined in the
//   128B-swizzle / 32-byte-atom layout.  fp16 operands (where MN-major no-swizzle exists) are the way to keep that idea.
// ---------------------------------------------------------------------------
constexpr uint32_t kLbo2 = kRows * 16 + 16;   // bytes between K-chunks of the H1 / dZ operands

struct alignas(16) TrainSmem {
    TrainBlob w;
    float a1[2][kRows][4];            // X, K-major
    unsigned char a2[32 * kLbo2];     // H1   [chunk j/4][row][4 floats], padded chunk stride
    unsigned char d2[32 * kLbo2];     // dZ2, later dZ1, same layout; scratch for the final reductions
    uint64_t bar_w;
    uint64_t bar_mma;
    uint32_t tmem_base;
    float red[16];
};

__device__ __forceinline__ float *opnd(unsigned char *base, unsigned chunk, unsigned row)
{
    return reinterpret_cast<float *>(base + chunk * kLbo2 + row * 16u);
}

__global__ void __launch_bounds__(kThreads, 1) ppo_grad_kernel(const PpoGradArgs P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem *sm = reinterpret_cast<TrainSmem *>(smem_raw);
    const unsigned t = threadIdx.x, warp = t >> 5, lane = t & 31u;

    if (t == 0) {
        mbar_init(&sm->bar_w, 1);
        mbar_init(&sm->bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 16) sm->red[t] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm->tmem_base;
    if (t == 0) {
        constexpr uint32_t bytes = static_cast<uint32_t>(sizeof(TrainBlob));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->bar_w)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm->w)),
                     "l"(P.blob), "r"(bytes), "r"(smem_u32(&sm->bar_w))
                     : "memory");
    }
    mbar_wait(&sm->bar_w, 0);

    const PolicyBlob &W = sm->w.fwd;
    constexpr uint32_t idesc1 = umma_idesc_tf32(128, 128);
    constexpr uint32_t idesc2 = umma_idesc_tf32(128, 64);
    const uint32_t a1 = smem_u32(sm->a1), a2 = smem_u32(sm->a2), d2 = smem_u32(sm->d2);
    const uint32_t w1 = smem_u32(W.w1), w2a = smem_u32(W.w2[0]), w2c = smem_u32(W.w2[1]);
    const uint32_t w2ta = smem_u32(sm->w.w2t[0]), w2tc = smem_u32(sm->w.w2t[1]);
    const uint32_t lane_base = tmem + (((warp & 3u) * 32u) << 16);
    uint32_t phase = 0;

    const unsigned n = static_cast<unsigned>(P.n);
    const unsigned n_tiles = (n + kRows - 1) / kRows;
    const float inv_n = 1.0f / static_cast<float>(n);
    const float log_std = P.logstd[0], inv_std = __expf(-log_std);
    const float adv_mean = P.norm_adv ? P.adv_stats[0] : 0.f;
    const float adv_scale = P.norm_adv ? 1.0f / (P.adv_stats[1] + 1e-8f) : 1.f;
    const float eps = P.clip_coef;
    const unsigned row = t & (kRows - 1), half = t >> 7;

    // accumulators that live across the CTA's tiles
    float2 acc_w2[4][4];    // patch of dW2[net = half]: rows j = jb*4.., columns i in chunks {ib, ib + 8}, as FFMA2 pairs
    float acc_b2[4] = {0.f, 0.f, 0.f, 0.f};
    float acc_w3[64];       // this thread's rows only; reduced over the CTA at the end
    float acc_b3 = 0.f, acc_logstd = 0.f;
    float acc_w1[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, acc_b1 = 0.f;
    float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f, st4 = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc_w2[a][b] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 64; ++c) acc_w3[c] = 0.f;
    const unsigned jb = (t & 127u) >> 3, ib = t & 7u;      // dW2 patch of this thread
    const unsigned j1 = t >> 1, kb = t & 1u;               // dW1: output feature j1 (0..127 = actor | critic), row half kb

    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned i = tile * kRows + row;
        const bool live = i < n;
        const int64_t src = live ? P.mb_idx[i] : 0;
        // ---- gather: observation row -> layer-1 operand; the row's scalars for the loss ----
        float s_a = 0.f, s_b = 0.f, s_c = 0.f;
        if (half == 0) {
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float *o = P.obs + src * 5;
                c0 = make_float4(o[0], o[1], o[2], o[3]);
                s_a = P.actions[src];
                s_b = P.logprobs[src];
                s_c = P.adv[src];
            }
            *reinterpret_cast<float4 *>(sm->a1[0][row]) = c0;
        } else {
            float o4 = 0.f;
            if (live) {
                o4 = P.obs[src * 5 + 4];
                s_a = P.ret[src];
                s_b = P.val[src];
            }
            *reinterpret_cast<float4 *>(sm->a1[1][row]) = make_float4(o4, 0.f, 0.f, 0.f);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

        // ---- H1 = tanh(D1 + b1) -> a2 ----
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            float v[32];
            const int col0 = half * 64 + blk * 32;
            tmem_ld32(lane_base + col0, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = col0 + c * 4;
                float4 h;
                h.x = fast_tanh(v[c * 4 + 0] + W.b1[j + 0]);
                h.y = fast_tanh(v[c * 4 + 1] + W.b1[j + 1]);
                h.z = fast_tanh(v[c * 4 + 2] + W.b1[j + 2]);
                h.w = fast_tanh(v[c * 4 + 3] + W.b1[j + 3]);
                *reinterpret_cast<float4 *>(opnd(sm->a2, j >> 2, row)) = h;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_tf32(tmem + 128, umma_desc(a2 + k * 2 * kLbo2, kLbo2, 128), umma_desc(w2a + k * 2 * kHid * 16, kHid * 16, 128), idesc2,
                          k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_tf32(tmem + 192, umma_desc(a2 + (16 + k * 2) * kLbo2, kLbo2, 128), umma_desc(w2c + k * 2 * kHid * 16, kHid * 16, 128),
                          idesc2, k > 0 ? 1u : 0u);
            umma_commit(&sm->bar_mma);
        }
        mbar_wait(&sm->bar_mma, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- H2 (registers), head ----
        float h2[64];
        float head = W.b3[half];
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            float v[32];
            const int col0 = half * 64 + blk * 32;
            tmem_ld32(lane_base + 128 + col0, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float h = fast_tanh(v[c] + W.b2[col0 + c]);
                h2[blk * 32 + c] = h;
                head = fmaf(h, W.w3[col0 + c], head);
            }
        }

        // ---- clipped-PPO loss and d loss / d head for this row ----
        float dy = 0.f;
        if (live) {
            if (half == 0) {
                const float z = (s_a - head) * inv_std;
                const float newlp = -0.5f * z * z - log_std - 0.91893853320467274178f;
                const float logratio = newlp - s_b;
                const float ratio = __expf(logratio);
                const float A = (s_c - adv_mean) * adv_scale;
                const float lo = 1.0f - eps, hi = 1.0f + eps;
                const float rc = fminf(fmaxf(ratio, lo), hi);
                const float pg1 = -A * ratio, pg2 = -A * rc;
                const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
                const float g_ratio = pg1 > pg2 ? -A : (pg2 > pg1 ? -A * in_range : -A * (0.5f + 0.5f * in_range));
                const float g_lp = g_ratio * ratio * inv_n;
                dy = g_lp * z * inv_std;
                acc_logstd += g_lp * (z * z - 1.0f) - P.ent_coef * inv_n;
                st0 -= logratio;
                st1 += (ratio - 1.0f) - logratio;
                st2 += fabsf(ratio - 1.0f) > eps ? 1.f : 0.f;
                st4 += fmaxf(pg1, pg2);
            } else {
                const float du = head - s_a;
                float g = du, vl = du * du;
                if (P.clip_vloss) {
                    const float dv = head - s_b;
                    const float vc = s_b + fminf(fmaxf(dv, -eps), eps);
                    const float dc = vc - s_a, vcl = dc * dc;
                    const float in_range = (dv >= -eps && dv <= eps) ? 1.f : 0.f;
                    g = vl > vcl ? du : (vcl > vl ? dc * in_range : 0.5f * du + 0.5f * dc * in_range);
                    vl = fmaxf(vl, vcl);
                }
                dy = P.vf_coef * g * inv_n;
                st3 += 0.5f * vl;
            }
        }
        acc_b3 += dy;
        // ---- dZ2 = dy * w3 * (1 - H2^2) -> d2; dW3 partial sums ----
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float4 g;
            const int j = half * 64 + c * 4;
            acc_w3[c * 4 + 0] = fmaf(dy, h2[c * 4 + 0], acc_w3[c * 4 + 0]);
            acc_w3[c * 4 + 1] = fmaf(dy, h2[c * 4 + 1], acc_w3[c * 4 + 1]);
            acc_w3[c * 4 + 2] = fmaf(dy, h2[c * 4 + 2], acc_w3[c * 4 + 2]);
            acc_w3[c * 4 + 3] = fmaf(dy, h2[c * 4 + 3], acc_w3[c * 4 + 3]);
            g.x = dy * W.w3[j + 0] * (1.0f - h2[c * 4 + 0] * h2[c * 4 + 0]);
            g.y = dy * W.w3[j + 1] * (1.0f - h2[c * 4 + 1] * h2[c * 4 + 1]);
            g.z = dy * W.w3[j + 2] * (1.0f - h2[c * 4 + 2] * h2[c * 4 + 2]);
            g.w = dy * W.w3[j + 3] * (1.0f - h2[c * 4 + 3] * h2[c * 4 + 3]);
            *reinterpret_cast<float4 *>(opnd(sm->d2, j >> 2, row)) = g;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D1 has been read by everyone: D3 may overwrite it
        __syncthreads();
        if (t == 0) {   // D3 = dZ2 * W2 (per net), into D1's columns
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_tf32(tmem + 0, umma_desc(d2 + k * 2 * kLbo2, kLbo2, 128), umma_desc(w2ta + k * 2 * kHid * 16, kHid * 16, 128), idesc2,
                          k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_tf32(tmem + 64, umma_desc(d2 + (16 + k * 2) * kLbo2, kLbo2, 128), umma_desc(w2tc + k * 2 * kHid * 16, kHid * 16, 128),
                          idesc2, k > 0 ? 1u : 0u);
            umma_commit(&sm->bar_mma);
        }
        // ---- while D3 is in flight: dW2[net][j][i] += sum_r dZ2[r][j] * H1[r][i] (outer products on the CUDA cores) ----
        {
            const unsigned cg = half * 16 + jb;                 // dZ2 chunk holding j = jb*4 .. jb*4+3 of this net
            const unsigned ch0 = half * 16 + ib, ch1 = ch0 + 8;  // H1 chunks: i = ib*4 .. +3 and 32 + ib*4 .. +3
#pragma unroll 4
            for (unsigned r = 0; r < kRows; ++r) {
                const float4 g = *reinterpret_cast<const float4 *>(opnd(sm->d2, cg, r));
                const float4 ha = *reinterpret_cast<const float4 *>(opnd(sm->a2, ch0, r));
                const float4 hb = *reinterpret_cast<const float4 *>(opnd(sm->a2, ch1, r));
                const float gv[4] = {g.x, g.y, g.z, g.w};
                const float2 hp[4] = {make_float2(ha.x, ha.y), make_float2(ha.z, ha.w), make_float2(hb.x, hb.y), make_float2(hb.z, hb.w)};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float2 g2 = make_float2(gv[a], gv[a]);
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc_w2[a][b] = __ffma2_rn(g2, hp[b], acc_w2[a][b]);   // packed FFMA2 (sm_100): half the issue slots
                }
                acc_b2[0] += g.x;   // every thread of a jb group keeps the sums; one of them flushes
                acc_b2[1] += g.y;
                acc_b2[2] += g.z;
                acc_b2[3] += g.w;
            }
        }
        mbar_wait(&sm->bar_mma, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        __syncthreads();   // every thread has finished reading dZ2: d2 may be overwritten with dZ1

        // ---- dZ1 = D3 * (1 - H1^2) -> d2 ----
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            float v[32];
            const int col0 = half * 64 + blk * 32;
            tmem_ld32(lane_base + col0, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = col0 + c * 4;
                const float4 h = *reinterpret_cast<const float4 *>(opnd(sm->a2, j >> 2, row));
                float4 g;
                g.x = v[c * 4 + 0] * (1.0f - h.x * h.x);
                g.y = v[c * 4 + 1] * (1.0f - h.y * h.y);
                g.z = v[c * 4 + 2] * (1.0f - h.z * h.z);
                g.w = v[c * 4 + 3] * (1.0f - h.w * h.w);
                *reinterpret_cast<float4 *>(opnd(sm->d2, j >> 2, row)) = g;
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // ---- dW1[j][k] += sum_r dZ1[r][j] * X[r][k]; db1[j] += sum_r dZ1[r][j] ----
        {   // thread = (output feature j1, half of the rows): one scalar dZ1 and one broadcast X row per step, 5 FMAs
            const unsigned cj = j1 >> 2, ej = j1 & 3u;
            const unsigned r0 = kb * (kRows / 2);
#pragma unroll 8
            for (unsigned r = r0; r < r0 + kRows / 2; ++r) {
                const float dz = opnd(sm->d2, cj, r)[ej];
                const float4 x0 = *reinterpret_cast<const float4 *>(sm->a1[0][r]);
                const float x4 = sm->a1[1][r][0];
                acc_w1[0] = fmaf(dz, x0.x, acc_w1[0]);
                acc_w1[1] = fmaf(dz, x0.y, acc_w1[1]);
                acc_w1[2] = fmaf(dz, x0.z, acc_w1[2]);
                acc_w1[3] = fmaf(dz, x0.w, acc_w1[3]);
                acc_w1[4] = fmaf(dz, x4, acc_w1[4]);
                acc_b1 += dz;
            }
        }
        __syncthreads();   // operands are rewritten by the next tile
    }

    // ---- flush: register accumulators -> global (atomics), per-thread partials reduced through shared memory ----
    PolicyGrad *G = P.grad;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const unsigned j = jb * 4 + a;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const unsigned ii = (b < 4 ? ib * 4 + b : 32 + ib * 4 + (b - 4));
            atomicAdd(&G->w2[half][j][ii], (b & 1) ? acc_w2[a][b >> 1].y : acc_w2[a][b >> 1].x);
        }
        if (ib == 0) atomicAdd(&G->b2[half][j], acc_b2[a]);
    }
    {
        const unsigned net = j1 >> 6, jn = j1 & 63u;
#pragma unroll
        for (int k = 0; k < 5; ++k) atomicAdd(&G->w1[net][jn][k], acc_w1[k]);
        atomicAdd(&G->b1[net][jn], acc_b1);
    }
    // dW3: [half][row][64] partials -> d2 scratch (64 KB), then 128 threads sum over the 128 rows
    {
        float *scr = reinterpret_cast<float *>(sm->d2);
#pragma unroll
        for (int c = 0; c < 64; ++c) scr[(half * kRows + row) * 64 + c] = acc_w3[c];
        __syncthreads();
        if (t < 128) {
            const unsigned net = t >> 6, c = t & 63u;
            float s = 0.f;
            for (unsigned r = 0; r < kRows; ++r) s += scr[(net * kRows + r) * 64 + c];
            atomicAdd(&G->w3[net][c], s);
        }
    }
    // scalars: warp shuffle, then shared-memory atomics, then one global atomic each
    float sc[8] = {st0, st1, st2, st3, st4, acc_logstd, half == 0 ? acc_b3 : 0.f, half == 1 ? acc_b3 : 0.f};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float v = sc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(&sm->red[q], v);
    }
    __syncthreads();
    if (t < 5) atomicAdd(&G->stats[t], sm->red[t]);
    if (t == 5) atomicAdd(&G->logstd, sm->red[5]);
    if (t == 6) atomicAdd(&G->b3[0], sm->red[6]);
    if (t == 7) atomicAdd(&G->b3[1], sm->red[7]);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
}


// ---------------------------------------------------------------------------
// fp16-operand training kernel: same structure as ppo_grad_kernel, with the weight gradients on the tensor cores.
//   D1 = X H1 ... as before (kind::f16, K = 16 per MMA);  D3 = dZ2 * W2;
//   dW2 | .  += dZ2^T * H1      M = 128 (both nets' outputs), N = 128 (both nets' inputs; the cross-net blocks are ignored)
//   db2      += dZ2^T * [X | 1] N = 16: column 5 of the X operand is 1
//   dW1 | db1 += dZ1^T * [X | 1]
// all accumulated in TMEM over the CTA's tiles ([256,384), [400,416), [384,400)) and read back once.  dZ is carried
// multiplied by the minibatch size so that it sits in fp16's normal range; the accumulators are rescaled at the flush.
// ---------------------------------------------------------------------------
constexpr uint32_t kLboH = kRows * 16 + 16;   // bytes between 8-feature chunks of the fp16 H1 / dZ operands

struct alignas(16) TrainSmem16 {
    TrainBlob16 w;
    // two tiles are in flight per CTA (slots 0 / 1, software-pipelined: while one slot's MMAs run, the CTA works on the other's epilogue)
    __half a1[2][2][kRows][8];        // X: chunk 0 = obs[0..4], 1, 0, 0; chunk 1 = 0
    unsigned char a2[2][16 * kLboH];  // H1  [chunk j/8][row][8 halfs]  (a2 doubles as the 64 KB scratch of the final dW3 reduction)
    unsigned char d2[2][16 * kLboH];  // dZ2, later dZ1
    float stage[2][2][kThreads][8];   // gathered minibatch rows [parity of the tile pair][slot][thread]: x[5], a, b, c
    int64_t next_idx[2][kThreads];    // minibatch indices of the NEXT pair's rows [slot][thread] (-1: dead row)
    uint64_t bar_w;
    uint64_t bar_mma[2];
    uint64_t bar_mma2[2];             // ppo_grad16w_kernel: the slot's dW2 | db2 MMAs have completed (its dZ2 / H1 operands may be overwritten)
    uint32_t tmem_base;
    float red[16];
    // ppo_grad16w_kernel: the biases ride in the MMAs, as in rollout_kernel: b = hi + lo in fp16 against two columns of ones (layer 1: X columns
    // 5, 6 and the CTA's copy of W1; layer 2: one more K = 16 MMA per net, A = `ones`, B = `wb2`)
    alignas(16) __half ones[2][kRows][8];
    alignas(16) __half wb2[2][2][kHid][8];
    alignas(16) __half2 w3h[64];      // ppo_grad16w_kernel: the heads' weights as packed halves (actor 0-31 | critic 32-63), for the half2 backward arithmetic
};
static_assert(sizeof(TrainSmem16) <= 227 * 1024, "shared memory per CTA");

__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool mn_major)
{
    return (1u << 4) | (mn_major ? (1u << 15) | (1u << 16) : 0u) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint4 pack8(const float (&v)[8])
{
    __half2 h[4] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3]), __floats2half2_rn(v[4], v[5]), __floats2half2_rn(v[6], v[7])};
    return *reinterpret_cast<const uint4 *>(h);
}

#ifdef F16_GRAD_TRACE
// debug build only (make EXTRA=-DF16_GRAD_TRACE): clock64 at the stage boundaries of CTA 0's 8 warps, tile pairs 2..5
__device__ unsigned long long g_grad_trace[8][4][10];   // [warp][pair 2..5][slot]
#define GTRACE(slot)                                                                                          \
    do {                                                                                                      \
        if (blockIdx.x == 0 && lane == 0 && tile_no >= 2 && tile_no < 6) g_grad_trace[warp][tile_no - 2][slot] = clock64(); \
    } while (0)
#else
#define GTRACE(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1) ppo_grad16_kernel(const PpoGradArgs P, const TrainBlob16 *blob16, float *partials)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem16 *sm = reinterpret_cast<TrainSmem16 *>(smem_raw);
    const unsigned t = threadIdx.x, lane = t & 31u;
    const unsigned warp = __shfl_sync(0xffffffffu, t >> 5, 0);   // warp-uniform for the compiler: the MMA-issuing branch stays on the uniform datapath

    if (t == 0) {
        mbar_init(&sm->bar_w, 1);
        mbar_init(&sm->bar_mma[0], 1);
        mbar_init(&sm->bar_mma[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (t < 16) sm->red[t] = 0.f;
    if (warp == 0) {   // all 512 TMEM columns: slot s: D1 -> D2 -> D3 in [128 s, 128 s + 128); dW2 [256,384), dW1|db1 [384,400), db2 [400,416)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const unsigned row = t & (kRows - 1), half = t >> 7;
    if (half == 1) {   // K padding 8..15 of both slots' X operands, written once
        *reinterpret_cast<uint4 *>(sm->a1[0][1][row]) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(sm->a1[1][1][row]) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm->tmem_base;
    // programmatic dependent launch: everything above overlapped the previous kernel's tail; the weights, log-std and
    // statistics read below are that kernel's (the optimiser's) output
    const TrainBlob16 &W = sm->w;
    constexpr uint32_t id_l1 = umma_idesc_f16(128, 128, false), id_l2 = umma_idesc_f16(128, 64, false);
    constexpr uint32_t id_w2 = umma_idesc_f16(128, 128, true), id_w1 = umma_idesc_f16(128, 16, true);
    const uint32_t w1 = smem_u32(W.w1), w2a = smem_u32(W.w2[0]), w2c = smem_u32(W.w2[1]);
    const uint32_t w2ta = smem_u32(W.w2t[0]), w2tc = smem_u32(W.w2t[1]);
    const uint32_t lane_base = tmem + (((warp & 3u) * 32u) << 16);

    const unsigned n = static_cast<unsigned>(P.n);
    const unsigned n_tiles = (n + kRows - 1) / kRows;
    const float n_f = static_cast<float>(n), inv_n = 1.0f / n_f;
    const float eps = P.clip_coef;

    // dW3: per tile the 64 products dy * H2[row][c] of a thread are summed over the warp's 32 rows by a shuffle butterfly that halves
    // the value count at every step (62 shuffles), leaving lane l with columns 2l, 2l+1 -- two persistent registers instead of 64
    // (with 64 the two-tile pipeline spilled: 255 registers).
    float acc_w3a = 0.f, acc_w3b = 0.f;
    float acc_b3 = 0.f, acc_logstd = 0.f;
    float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f, st4 = 0.f;

    // One row of the minibatch = a random 20-byte row of the rollout buffer plus three scalars behind an index: two dependent DRAM
    // round trips.  Both are taken off the critical path AND out of the long-lived register set (the kernel sits at 255 registers:
    // rows held across the loss stage spilled, and a spill store waits for its load): after stage 1 the rows of the next pair are
    // requested with the indices fetched one round earlier, the indices of the pair after that are requested too, stage 2 runs on
    // top of those loads, and before stage 3 everything is parked in shared memory.  (`cp.async` straight into shared memory was
    // tried first: every `fence.proxy.async` of the stages then waits for the outstanding copies.)
    auto load_index = [&](unsigned tile_) -> int64_t {
        const unsigned i_ = tile_ * kRows + row;
        return (tile_ < n_tiles && i_ < n) ? P.mb_idx[i_] : int64_t(-1);
    };
    struct RowIn { float v[8]; };   // x[5], a, b, c
    auto load_row = [&](int64_t src) {
        RowIn g;
#pragma unroll
        for (int k = 0; k < 8; ++k) g.v[k] = 0.f;
        if (src >= 0) {
            if (P.packed) {            // one aligned 32-byte sector (half 0) / 8 bytes (half 1) per thread
                if (half == 0) {
                    const float4 *r = reinterpret_cast<const float4 *>(P.obs + src * 8);
                    const float4 lo = __ldg(r), hi = __ldg(r + 1);
                    g.v[0] = lo.x, g.v[1] = lo.y, g.v[2] = lo.z, g.v[3] = lo.w;
                    g.v[4] = hi.x, g.v[5] = hi.y, g.v[6] = hi.z, g.v[7] = hi.w;
                } else {
                    const float2 rv = __ldg(reinterpret_cast<const float2 *>(P.ret + src * 2));
                    g.v[5] = rv.x, g.v[6] = rv.y;
                }
            } else if (half == 0) {
                const float *o = P.obs + src * 5;
#pragma unroll
                for (int k = 0; k < 5; ++k) g.v[k] = o[k];
                g.v[5] = P.actions[src];
                g.v[6] = P.logprobs[src];
                g.v[7] = P.adv[src];
            } else {
                g.v[5] = P.ret[src];
                g.v[6] = P.val[src];
            }
        }
        return g;
    };
    auto park_row = [&](float *dst, const RowIn &g) {
        *reinterpret_cast<float4 *>(dst) = make_float4(g.v[0], g.v[1], g.v[2], g.v[3]);
        *reinterpret_cast<float4 *>(dst + 4) = make_float4(g.v[4], g.v[5], g.v[6], g.v[7]);
    };
    auto row_live = [&](unsigned tile_) { return tile_ < n_tiles && tile_ * kRows + row < n; };

    // The first two pairs' gathers (index, then row: two dependent DRAM round trips) are requested BEFORE the grid dependency resolves:
    // the minibatch indices and the packed rows are written once per epoch / update by plain (fully ordered) launches, never by the
    // optimiser kernel this launch is chained to with PDL, whose single CTA leaves the other SMs idle for its ~14 us anyway.
    park_row(sm->stage[0][0][t], load_row(load_index(blockIdx.x)));
    park_row(sm->stage[0][1][t], load_row(load_index(blockIdx.x + gridDim.x)));
    sm->next_idx[0][t] = load_index(blockIdx.x + 2 * gridDim.x);
    sm->next_idx[1][t] = load_index(blockIdx.x + 3 * gridDim.x);
    // the weights, log-std and advantage statistics are the previous kernels' (optimiser / adv_stats) output
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // dependents are released only now: whatever the next kernel does ahead of ITS wait then runs after everything older than this
    // kernel has completed (ppo_reduce_adam_kernel pre-loads the optimiser state its previous call wrote)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (t == 0) {
        constexpr uint32_t bytes = static_cast<uint32_t>(sizeof(TrainBlob16));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->bar_w)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm->w)),
                     "l"(blob16), "r"(bytes), "r"(smem_u32(&sm->bar_w))
                     : "memory");
    }
    const float log_std = P.logstd[0], inv_std = __expf(-log_std);
    const float adv_mean = P.norm_adv ? P.adv_stats[0] : 0.f;
    const float adv_scale = P.norm_adv ? 1.0f / (P.adv_stats[1] + 1e-8f) : 1.f;

    // ---- the four stages of a tile; `s` is the slot (0 / 1): shared-memory buffers, accumulator columns and mbarrier of that slot ----
    // Every stage ends by handing a batch of MMAs to the tensor core; the wait for them opens the tile's NEXT stage, which runs after a
    // stage of the other slot: the round trips (r02 trace: 2.5 k of 9.8 k clocks per tile) and the CUDA-core work overlap.
    // S1: observation row -> X operand (with a 1 in column 5: bias gradients = that column of dZ^T [X | 1]); issue layer 1
    auto stage1 = [&](const int s, const float *in, const bool live) {
        if (half == 0) {
            float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f};
            if (live) {
#pragma unroll
                for (int k = 0; k < 5; ++k) x[k] = in[k];
            }
            *reinterpret_cast<uint4 *>(sm->a1[s][0][row]) = pack8(x);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0 && umma::elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_f16(tmem + s * 128, umma_desc(smem_u32(sm->a1[s]), kRows * 16, 128), umma_desc(w1, 128 * 16, 128), id_l1, 0u);
            umma_commit(&sm->bar_mma[s]);
        }
    };
    // S2: H1 = tanh(D1 + b1) -> a2 (fp16); issue layer 2 (into D1's columns: every thread has read them)
    auto stage2 = [&](const int s) {
        mbar_wait(&sm->bar_mma[s], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        unsigned char *a2p = sm->a2[s];
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(lane_base + s * 128 + half * 64, ra);
        tmem_ld32_issue(lane_base + s * 128 + half * 64 + 32, rb);
        tmem_wait_ld();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            const int col0 = half * 64 + blk * 32;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = col0 + c * 8;
                float h[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) h[q] = fast_tanh(__uint_as_float(blk ? rb[c * 8 + q] : ra[c * 8 + q]) + W.b1[j + q]);
                *reinterpret_cast<uint4 *>(a2p + (j >> 3) * kLboH + row * 16u) = pack8(h);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0 && umma::elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a2 = smem_u32(a2p);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(tmem + s * 128, umma_desc(a2 + k * 2 * kLboH, kLboH, 128), umma_desc(w2a + k * 2 * kHid * 16, kHid * 16, 128), id_l2,
                         k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(tmem + s * 128 + 64, umma_desc(a2 + (8 + k * 2) * kLboH, kLboH, 128), umma_desc(w2c + k * 2 * kHid * 16, kHid * 16, 128), id_l2,
                         k > 0 ? 1u : 0u);
            umma_commit(&sm->bar_mma[s]);
        }
    };
    // S3: H2, head, clipped-PPO loss, dZ2 -> d2; issue D3 = dZ2 W2 (into D2's columns) and the weight gradients dW2, db2
    auto stage3 = [&](const int s, const float *in, const bool live, const bool first) {   // first: this batch initialises the dW2 / db2 accumulators
        mbar_wait(&sm->bar_mma[s], 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const float s_a = live ? in[5] : 0.f, s_b = live ? in[6] : 0.f, s_c = live ? in[7] : 0.f;
        unsigned char *d2p = sm->d2[s];
        uint32_t ra[32], rb[32];
        float h2[64];
        float head = W.b3[half];
        tmem_ld32_issue(lane_base + s * 128 + half * 64, ra);
        tmem_ld32_issue(lane_base + s * 128 + half * 64 + 32, rb);
        tmem_wait_ld();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            const int col0 = half * 64 + blk * 32;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const float h = fast_tanh(__uint_as_float(blk ? rb[c] : ra[c]) + W.b2[col0 + c]);
                h2[blk * 32 + c] = h;
                head = fmaf(h, W.w3[col0 + c], head);
            }
        }

        // ---- clipped-PPO loss; g = d(n * loss) / d head for this row ----
        float g_head = 0.f;
        if (live) {
            if (half == 0) {
                const float z = (s_a - head) * inv_std;
                const float newlp = -0.5f * z * z - log_std - 0.91893853320467274178f;
                const float logratio = newlp - s_b;
                const float ratio = __expf(logratio);
                const float A = (s_c - adv_mean) * adv_scale;
                const float lo = 1.0f - eps, hi = 1.0f + eps;
                const float rc = fminf(fmaxf(ratio, lo), hi);
                const float pg1 = -A * ratio, pg2 = -A * rc;
                const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
                const float g_ratio = pg1 > pg2 ? -A : (pg2 > pg1 ? -A * in_range : -A * (0.5f + 0.5f * in_range));
                const float g_lp = g_ratio * ratio;
                g_head = g_lp * z * inv_std;
                acc_logstd += (g_lp * (z * z - 1.0f) - P.ent_coef) * inv_n;
                st0 -= logratio;
                st1 += (ratio - 1.0f) - logratio;
                st2 += fabsf(ratio - 1.0f) > eps ? 1.f : 0.f;
                st4 += fmaxf(pg1, pg2);
            } else {
                const float du = head - s_a;
                float g = du, vl = du * du;
                if (P.clip_vloss) {
                    const float dv = head - s_b;
                    const float vc = s_b + fminf(fmaxf(dv, -eps), eps);
                    const float dc = vc - s_a, vcl = dc * dc;
                    const float in_range = (dv >= -eps && dv <= eps) ? 1.f : 0.f;
                    g = vl > vcl ? du : (vcl > vl ? dc * in_range : 0.5f * du + 0.5f * dc * in_range);
                    vl = fmaxf(vl, vcl);
                }
                g_head = P.vf_coef * g;
                st3 += 0.5f * vl;
            }
        }
        const float dy = g_head * inv_n;
        acc_b3 += dy;
        // ---- n * dZ2 = g * w3 * (1 - H2^2) -> d2 (fp16); dW3 partial sums in fp32 ----
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = half * 64 + c * 8;
            float g[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float h = h2[c * 8 + q];
                g[q] = g_head * W.w3[j + q] * (1.0f - h * h);
                h2[c * 8 + q] = dy * h;
            }
            *reinterpret_cast<uint4 *>(d2p + (j >> 3) * kLboH + row * 16u) = pack8(g);
        }
#pragma unroll
        for (int sz = 32; sz >= 2; sz >>= 1) {   // 64 -> 32 -> 16 -> 8 -> 4 -> 2 values; lanes whose bit sz/2 is set keep the upper half
            const unsigned o = static_cast<unsigned>(sz) >> 1;
            const bool upper = (lane & o) != 0u;
#pragma unroll
            for (int i = 0; i < sz; ++i) {
                const float lo = h2[i], hi = h2[i + sz];
                h2[i] = (upper ? hi : lo) + __shfl_xor_sync(0xffffffffu, upper ? lo : hi, o);
            }
        }
        acc_w3a += h2[0];
        acc_w3b += h2[1];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D2 has been read by everyone: D3 may overwrite it
        __syncthreads();
        if (warp == 0 && umma::elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a1 = smem_u32(sm->a1[s]), a2 = smem_u32(sm->a2[s]), d2 = smem_u32(d2p);
            // D3 = dZ2 * W2 (per net): dZ2 read K-major (rows x features)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(tmem + s * 128, umma_desc(d2 + k * 2 * kLboH, kLboH, 128), umma_desc(w2ta + k * 2 * kHid * 16, kHid * 16, 128), id_l2,
                         k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16(tmem + s * 128 + 64, umma_desc(d2 + (8 + k * 2) * kLboH, kLboH, 128), umma_desc(w2tc + k * 2 * kHid * 16, kHid * 16, 128), id_l2,
                         k > 0 ? 1u : 0u);
            // dW2 += dZ2^T * H1, db2 += dZ2^T * [X | 1]: the same buffers read MN-major (features x rows), 16 rows per MMA
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_f16(tmem + 256, umma_desc(d2 + k * 256, 128, kLboH), umma_desc(a2 + k * 256, 128, kLboH), id_w2, (first && k == 0) ? 0u : 1u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_f16(tmem + 400, umma_desc(d2 + k * 256, 128, kLboH), umma_desc(a1 + k * 256, 128, kRows * 16), id_w1,
                         (first && k == 0) ? 0u : 1u);
            umma_commit(&sm->bar_mma[s]);
        }
    };
    // S4: n * dZ1 = D3 * (1 - H1^2) -> d2 (every MMA that read dZ2 has completed); issue dW1 | db1 += dZ1^T [X | 1]
    auto stage4 = [&](const int s, const bool first) {
        mbar_wait(&sm->bar_mma[s], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        unsigned char *a2p = sm->a2[s], *d2p = sm->d2[s];
        uint32_t ra[32], rb[32];
        tmem_ld32_issue(lane_base + s * 128 + half * 64, ra);
        tmem_ld32_issue(lane_base + s * 128 + half * 64 + 32, rb);
        tmem_wait_ld();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
            const int col0 = half * 64 + blk * 32;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = col0 + c * 8;
                const uint4 raw = *reinterpret_cast<const uint4 *>(a2p + (j >> 3) * kLboH + row * 16u);
                const __half2 *hh = reinterpret_cast<const __half2 *>(&raw);
                float g[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 h = __half22float2(hh[q]);
                    g[2 * q] = __uint_as_float(blk ? rb[c * 8 + 2 * q] : ra[c * 8 + 2 * q]) * (1.0f - h.x * h.x);
                    g[2 * q + 1] = __uint_as_float(blk ? rb[c * 8 + 2 * q + 1] : ra[c * 8 + 2 * q + 1]) * (1.0f - h.y * h.y);
                }
                *reinterpret_cast<uint4 *>(d2p + (j >> 3) * kLboH + row * 16u) = pack8(g);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0 && umma::elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a1 = smem_u32(sm->a1[s]), d2 = smem_u32(d2p);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                umma_f16(tmem + 384, umma_desc(d2 + k * 256, 128, kLboH), umma_desc(a1 + k * 256, 128, kRows * 16), id_w1,
                         (first && k == 0) ? 0u : 1u);
            umma_commit(&sm->bar_mma[s]);
        }
    };
    // the slot's operands are rewritten by its next tile: its last batch (dW1) must have completed
    auto retire = [&](const int s) {
        mbar_wait(&sm->bar_mma[s], 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    // a CTA's tiles: blockIdx.x + j gridDim.x, taken two at a time (slot = j & 1); a missing second tile runs as 128 dead rows (zero gradient)
    const unsigned stride = gridDim.x;
    mbar_wait(&sm->bar_w, 0);
    int tile_no = 0;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += 2 * stride) {
        GTRACE(0);
        const int par = tile_no & 1;
        const float *in0 = sm->stage[par][0][t], *in1 = sm->stage[par][1][t];
        const bool live0 = row_live(tile), live1 = row_live(tile + stride);
        if (tile_no) retire(0);
        stage1(0, in0, live0);
        if (tile_no) retire(1);
        stage1(1, in1, live1);
        GTRACE(1);
        stage2(0);
        GTRACE(2);
        stage2(1);
        // Rows of the next pair (indices fetched one round ago) and the indices of the pair after that.  Requested HERE because every
        // stage ends with fence.proxy.async (= FENCE.VIEW.ASYNC + MEMBAR.ALL.CTA, which waits for the thread's outstanding loads):
        // the loss stage of slot 0 is the longest stretch without one (~3.5 k clocks), so the DRAM round trip hides under it.
        const RowIn r0 = load_row(sm->next_idx[0][t]), r1 = load_row(sm->next_idx[1][t]);
        const int64_t i0 = load_index(tile + 4 * stride), i1 = load_index(tile + 5 * stride);
        GTRACE(3);
        stage3(0, in0, live0, tile_no == 0);
        park_row(sm->stage[par ^ 1][0][t], r0);
        park_row(sm->stage[par ^ 1][1][t], r1);
        sm->next_idx[0][t] = i0;
        sm->next_idx[1][t] = i1;
        GTRACE(4);
        stage3(1, in1, live1, false);
        GTRACE(5);
        stage4(0, tile_no == 0);
        GTRACE(6);
        stage4(1, false);
        GTRACE(7);
        ++tile_no;
    }
    retire(0);
    retire(1);

    // ---- flush ----
    // with `partials` every CTA owns a private copy of the gradient and writes each element exactly once; otherwise atomics
    const bool own = partials != nullptr;
    PolicyGrad *G = own ? reinterpret_cast<PolicyGrad *>(partials) + blockIdx.x : P.grad;
    auto put = [&](float *dst, float val) {
        if (own) *dst = val;
        else atomicAdd(dst, val);
    };
    {
        // weight-gradient accumulators: TMEM lane = output feature j (actor 0-63 | critic 64-127), scaled by n.  dW2: the block of
        // the lane's own net, 64 columns split between the lane's two threads; dW1 (384..388), db1 (389), db2 (405): half 0.
        const unsigned net = row >> 6, jn = row & 63u;
        float v[32];
        tmem_ld32(lane_base + 256 + net * 64 + half * 32, v);
        // (148 CTAs add into the same 8 k addresses here: 1.2 M atomics cost ~21 us of the kernel whatever the order each
        // CTA walks its columns in -- measured; per-CTA partials + a reduction kernel would be the next step)
        if (own) {   // a thread's 32 columns are 128 contiguous bytes of its CTA's private gradient: eight 16-byte stores instead of 32 scalar ones
            float4 *dst = reinterpret_cast<float4 *>(&G->w2[net][jn][half * 32]);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c] * inv_n, v[4 * c + 1] * inv_n, v[4 * c + 2] * inv_n, v[4 * c + 3] * inv_n);
        } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) put(&G->w2[net][jn][half * 32 + c], v[c] * inv_n);
        }
        if (half == 0) {
            tmem_ld32(lane_base + 384, v);
#pragma unroll
            for (int k = 0; k < 5; ++k) put(&G->w1[net][jn][k], v[k] * inv_n);
            put(&G->b1[net][jn], v[5] * inv_n);
            put(&G->b2[net][jn], v[16 + 5] * inv_n);
        }
    }
    __syncthreads();
    // dW3: lane l of every warp holds columns 2l, 2l+1 of its net, summed over the warp's rows and the CTA's tiles: the four warps
    // of a net meet in shared memory and are added in a fixed order (the gradient stays bit-reproducible)
    {
        float *scr = reinterpret_cast<float *>(sm->a2);
        scr[warp * 64 + 2 * lane] = acc_w3a;
        scr[warp * 64 + 2 * lane + 1] = acc_w3b;
        __syncthreads();
        if (t < 128) {
            const unsigned net = t >> 6, c = t & 63u;
            put(&G->w3[net][c], (scr[(net * 4 + 0) * 64 + c] + scr[(net * 4 + 1) * 64 + c]) + (scr[(net * 4 + 2) * 64 + c] + scr[(net * 4 + 3) * 64 + c]));
        }
    }
    float sc[8] = {st0, st1, st2, st3, st4, acc_logstd, half == 0 ? acc_b3 : 0.f, half == 1 ? acc_b3 : 0.f};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float v = sc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(&sm->red[q], v);
    }
    __syncthreads();
    if (t < 5) put(&G->stats[t], sm->red[t]);
    if (t == 5) put(&G->logstd, sm->red[5]);
    if (t == 6) put(&G->b3[0], sm->red[6]);
    if (t == 7) put(&G->b3[1], sm->red[7]);

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// ---------------------------------------------------------------------------
// ppo_grad16w_kernel: the same tile program as ppo_grad16_kernel with the two slots on their OWN warps -- 512 threads, slot s =
// warps 8s .. 8s+7, each slot a free-running 256-thread group (named barrier 1 + s) that walks its own tiles through the four
// stages.  ppo_grad16_kernel interleaves the slots in program order on the same 8 warps: every round trip is hidden, but the SM
// never holds more than two warps per sub-partition and all of them sit in the same phase (ncu: issue slots 28 % busy, stall
// samples spread evenly -- latency-bound).  Here a sub-partition holds four warps, two of each slot, and the slots are out of
// phase by construction: the weight-gradient accumulators (dW2 | db2, dW1 | db1) are shared, so the two issuing warps pass a
// token -- slot 0's S3 batch of pair j, then slot 1's, then slot 0's of pair j + 1 ... and the same for the S4 batches -- which
// keeps the accumulation order, and with it every bit of the gradient, fixed, and makes slot 1 trail slot 0 by a fraction of a tile.
// 128 registers per thread (ppo_grad16_kernel: 226): a thread carries one slot, and the gathered row of its next tile only across S4.
// H2 is held as packed halves, dZ2 / dZ1 are computed in packed-half arithmetic (they are fp16 tensor-core operands), dW3 goes through a
// half2 shuffle butterfly, and the biases ride in the MMAs (hi + lo fp16 split against two columns of ones, as in rollout_kernel).
// ---------------------------------------------------------------------------
constexpr int kThreadsW = 512;

__device__ __forceinline__ void group_sync(int s) { asm volatile("bar.sync %0, 256;" ::"r"(s + 1) : "memory"); }
__device__ __forceinline__ void token_wait(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void token_pass(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

#ifdef F16_GRAD_TRACE
// debug build only: clock64 at the stage boundaries of CTA 0, warps 0-3 of each slot (trace rows 4 s + (warp & 3)), rounds 2..5
#define WTRACE(pt)                                                                                                                  \
    do {                                                                                                                            \
        if (blockIdx.x == 0 && lane == 0 && (warp & 7u) < 4u && j >= 2 && j < 6) g_grad_trace[s * 4 + (warp & 3u)][j - 2][pt] = clock64(); \
    } while (0)
__device__ unsigned long long g_epoch_cta[160][12];   // globaltimer (ns) per CTA at the same points, thread 0, minibatch 2
__device__ unsigned long long g_epoch_trace[4][12];   // [CTA 0 / last CTA][thread 0 / 256]... rows: (blockIdx.x == 0 ? 0 : 2) + (t == 256), minibatch 2
#define ETRACE(pt)                                                                                                   \
    do {                                                                                                             \
        if (EPOCH && mbi == 2 && (t & 255u) == 0u && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))                \
            g_epoch_trace[(blockIdx.x == 0 ? 0 : 2) + (t >> 8)][pt] = clock64();                                      \
        if (EPOCH && mbi == 2 && t == 0 && blockIdx.x < 160) {                                                       \
            unsigned long long now_;                                                                                 \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                                                 \
            g_epoch_cta[blockIdx.x][pt] = now_;                                                                      \
        }                                                                                                            \
    } while (0)
#else
#define WTRACE(pt) do { } while (0)
#define ETRACE(pt) do { } while (0)
#endif
// EPOCH = true: the kernel is persistent over the n_mb minibatches of an epoch and takes the optimiser step itself between two
// minibatches (cooperative launch, one CTA per SM): after the flush the CTAs meet at a grid-wide counter, every CTA sums the partial
// gradients of the parameters it owns (n_out / gridDim.x, in 64-parameter passes) and posts its share of the squared norm, a second
// meeting makes the clip coefficient known everywhere (shares added in a fixed order: bit-identical in every CTA), the owners apply
// clip_grad_norm_ + Adam(amsgrad) and scatter the new parameters into the three weight blobs (as ppo_reduce_adam_kernel does), and
// after a third meeting every CTA re-reads the fp16 training blob from L2.  Per minibatch this replaces a kernel boundary, the
// 289-CTA optimiser launch, the next gradient launch's prologue (TMEM allocation, barrier set-up, weight TMA) and the drain / fill
// of the SMs around them; the gather pipeline numbers its rounds through the whole launch, so the first rows of the next minibatch
// travel under the last rounds of this one.  With A.bufs the reduction phase also takes the ranks' mean over NVLink peer memory.
template <bool EPOCH>
__global__ void __launch_bounds__(kThreadsW, 1) ppo_grad16w_kernel(const PpoGradArgs P, const TrainBlob16 *blob16, float *partials, const ReduceAdamArgs A,
                                                                   const int n_mb)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem16 *sm = reinterpret_cast<TrainSmem16 *>(smem_raw);
    const unsigned t = threadIdx.x, lane = t & 31u;
    const unsigned warp = __shfl_sync(0xffffffffu, t >> 5, 0);   // warp-uniform for the compiler (MMA issue on the uniform datapath)
    const int s = static_cast<int>(warp >> 3);                   // slot = warp group
    const bool issuer = (warp & 7u) == 0u;                       // warps 0 and 8
    const unsigned tg = t & 255u, row = tg & (kRows - 1), half = tg >> 7;

    if (t == 0) {
        mbar_init(&sm->bar_w, 1);
        mbar_init(&sm->bar_mma[0], 1);
        mbar_init(&sm->bar_mma[1], 1);
        mbar_init(&sm->bar_mma2[0], 1);
        mbar_init(&sm->bar_mma2[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // all 512 TMEM columns, as in ppo_grad16_kernel
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (half == 1) *reinterpret_cast<uint4 *>(sm->a1[s][1][row]) = make_uint4(0u, 0u, 0u, 0u);   // K padding 8..15 of the slot's X operand
    if (t < kRows) {   // the layer-2 bias MMA's A operand: k = 0, 1 are 1.0 in every row
        *reinterpret_cast<uint4 *>(sm->ones[0][t]) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(sm->ones[1][t]) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm->tmem_base;
    const TrainBlob16 &W = sm->w;
    constexpr uint32_t id_l1 = umma_idesc_f16(128, 128, false), id_l2 = umma_idesc_f16(128, 64, false);
    constexpr uint32_t id_w2 = umma_idesc_f16(128, 128, true), id_w1 = umma_idesc_f16(128, 16, true);
    const uint32_t w1 = smem_u32(W.w1), w2a = smem_u32(W.w2[0]), w2c = smem_u32(W.w2[1]);
    const uint32_t w2ta = smem_u32(W.w2t[0]), w2tc = smem_u32(W.w2t[1]);
    const uint32_t tcol = tmem + s * 128;                                       // the slot's D1 -> D2 -> D3 columns
    const uint32_t lane_base = tcol + (((warp & 3u) * 32u) << 16) + half * 64;  // this thread's row and net in them
    uint64_t *bar = &sm->bar_mma[s];
    unsigned char *a2p = sm->a2[s], *d2p = sm->d2[s];
    const uint32_t a1u = smem_u32(sm->a1[s]), a2u = smem_u32(a2p), d2u = smem_u32(d2p);

    const unsigned n = static_cast<unsigned>(P.n);
    const unsigned n_tiles = (n + kRows - 1) / kRows;
    const float n_f = static_cast<float>(n), inv_n = 1.0f / n_f;
    const float eps = P.clip_coef;

    // the slot's tiles: blockIdx.x + (2 j + s) gridDim.x, j = 0 .. n_iter - 1; both slots run the same number of rounds (a missing tile = 128
    // dead rows).  The gather pipeline numbers its rounds through the whole launch: round r = minibatch r / n_iter, tile round r % n_iter, so
    // with EPOCH the first rows of the next minibatch travel under the last rounds of this one (they do not depend on the optimiser step)
    const unsigned stride = gridDim.x;
    const unsigned n_iter = (n_tiles - blockIdx.x + 2 * stride - 1) / (2 * stride);
    const unsigned tile0 = blockIdx.x + s * stride;
    // where the index of this thread's row in tile round jj of minibatch mb_ lives; jj may run up to n_iter + 1 (two rounds ahead of the last
    // round of a minibatch): it is carried into the following minibatch(es) without a division
    auto index_addr = [&](unsigned mb_, unsigned jj, bool &ok) -> const int64_t * {
        if (jj >= n_iter) {
            jj -= n_iter, ++mb_;
            if (jj >= n_iter) jj -= n_iter, ++mb_;
        }
        const unsigned tile_ = tile0 + 2 * stride * jj;
        const unsigned i_ = tile_ * kRows + row;
        ok = mb_ < static_cast<unsigned>(n_mb) && tile_ < n_tiles && i_ < n;
        return P.mb_idx + static_cast<size_t>(mb_) * n + i_;
    };
    auto load_index = [&](unsigned mb_, unsigned jj) -> int64_t {
        bool ok;
        const int64_t *a = index_addr(mb_, jj, ok);
        return ok ? *a : int64_t(-1);
    };
    // L2 prefetches one round ahead of the gathers (fire and forget: unlike a load they are not waited for by the fences that end every
    // stage): the row a thread will gather in S4 of this round, and the line of indices it will read there.  Without them a group waits,
    // once per tile, for the slowest of its 256 random DRAM accesses within the ~1 us of S4 (per-CTA tile loops spread 45 .. 55 us)
    auto prefetch_row = [&](int64_t src) {
        if (P.packed && src >= 0) {
            const void *a = half == 0 ? static_cast<const void *>(P.obs + src * 8) : static_cast<const void *>(P.ret + src * 2);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    };
    auto prefetch_index = [&](unsigned mb_, unsigned jj) {
        bool ok;
        const int64_t *a = index_addr(mb_, jj, ok);
        if (ok && (tg & 15u) == 0u && tg < 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
    };
    // A gathered row: (x[0..3]), (x[4], a, b, c) for the actor half's thread, (ret, val) for the critic half's.  The two halves' loads stay
    // in SEPARATE registers until they are parked: merged into one array, the compiler joins the two branches with moves that wait for the
    // load right where it was issued (ncu: the top long-scoreboard stall of the kernel), and the gather no longer travels under S4.
    struct RowIn { float4 lo, hi; float2 rv; };
    auto load_row = [&](int64_t src) {
        // No zero-initialisation and no `src >= 0` guard around the loads: a value that is either loaded or zero is a predicated move behind the
        // load, and the thread then waits for the gather where the move sits (ncu: the first instruction of S4).  A dead row (src < 0) gathers
        // row 0 instead; nothing reads what it parks (S1 and S3 mask dead rows themselves).  Only the fields of the thread's own half are set.
        RowIn g;
        const int64_t row_ = src >= 0 ? src : 0;
        if (P.packed) {            // one aligned 32-byte sector (half 0) / 8 bytes (half 1) per thread
            if (half == 0) {
                const float4 *r = reinterpret_cast<const float4 *>(P.obs + row_ * 8);
                g.lo = __ldg(r), g.hi = __ldg(r + 1);
            } else {
                g.rv = __ldg(reinterpret_cast<const float2 *>(P.ret + row_ * 2));
            }
        } else if (half == 0) {
            const float *o = P.obs + row_ * 5;
            g.lo = make_float4(o[0], o[1], o[2], o[3]);
            g.hi = make_float4(o[4], P.actions[row_], P.logprobs[row_], P.adv[row_]);
        } else {
            g.rv = make_float2(P.ret[row_], P.val[row_]);
        }
        return g;
    };
    auto park_row = [&](float *dst, const RowIn &g) {   // [0..4] x (read by the actor half only), [5] a | ret, [6] logprob | val, [7] advantage
        if (half == 0) {
            *reinterpret_cast<float4 *>(dst) = g.lo;
            *reinterpret_cast<float4 *>(dst + 4) = g.hi;
        } else {
            *reinterpret_cast<float4 *>(dst + 4) = make_float4(0.f, g.rv.x, g.rv.y, 0.f);
        }
    };

    // first round's row and second round's indices: requested BEFORE the grid dependency resolves (see ppo_grad16_kernel)
    park_row(sm->stage[0][s][tg], load_row(load_index(0, 0)));
    sm->next_idx[s][tg] = load_index(0, 1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (t == 0) {
        constexpr uint32_t bytes = static_cast<uint32_t>(sizeof(TrainBlob16));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->bar_w)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm->w)),
                     "l"(blob16), "r"(bytes), "r"(smem_u32(&sm->bar_w))
                     : "memory");
    }
    mbar_wait(&sm->bar_w, 0);
    // derived operands of a freshly loaded blob: biases as (hi, lo) halves into W1's columns 5, 6 and into wb2, w3 as packed halves
    auto derive_weights = [&]() {
        if (t < 128) {
            const float b1v = W.b1[t], b2v = W.b2[t];
            const __half h1 = __float2half_rn(b1v), h2v = __float2half_rn(b2v);
            sm->w.w1[0][t][5] = h1;
            sm->w.w1[0][t][6] = __float2half_rn(b1v - __half2float(h1));
            const __half2 b2p = __halves2half2(h2v, __float2half_rn(b2v - __half2float(h2v)));
            *reinterpret_cast<uint4 *>(sm->wb2[t >> 6][0][t & 63u]) = make_uint4(*reinterpret_cast<const uint32_t *>(&b2p), 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(sm->wb2[t >> 6][1][t & 63u]) = make_uint4(0u, 0u, 0u, 0u);
        } else if (t < 192) {
            const unsigned c = t - 128;
            sm->w3h[c] = __floats2half2_rn(W.w3[2 * c], W.w3[2 * c + 1]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    };
    derive_weights();

    // ---- EPOCH: grid-wide meetings at one counter (zeroed by the launcher before every launch; a cooperative launch: all CTAs are resident) ----
    // (the cooperative-groups pattern: the CTA's writes are ordered before thread 0's fence by the CTA barrier, the fence is cumulative;
    // everything read after a meeting is read with ld.global.cg, past the non-coherent L1)
    auto grid_meet = [&](unsigned n_meet) {   // n_meet: the number of this meeting in the launch, from 1
        __syncthreads();
        if (t == 0) {
            unsigned long long *arrive = reinterpret_cast<unsigned long long *>(A.sync);
            __threadfence();
            atomicAdd(arrive, 1ull);
            const unsigned long long want = static_cast<unsigned long long>(n_meet) * gridDim.x;
            unsigned long long t0 = 0ull;
            for (unsigned spin = 0;; ++spin) {
                unsigned long long seen;
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(arrive) : "memory");
                if (seen >= want) break;
                if ((spin & 1023u) == 1023u) {   // a CTA that never arrives must not hang the GPU: fail loudly after ~10 s
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > 10000000000ull) __trap();
                }
            }
        }
        __syncthreads();
    };
    const float step0 = EPOCH ? A.step[0] : 0.f;   // read before anybody's first meeting; CTA 0 writes the new value after a step's second meeting
    unsigned round = 0;                            // rounds of the gather pipeline done so far (all minibatches)

  for (int mbi = 0; mbi < n_mb; ++mbi) {
    const float log_std = EPOCH ? __ldcg(P.logstd) : P.logstd[0], inv_std = __expf(-log_std);   // EPOCH: a parameter another CTA has just updated
    const float adv_mean = P.norm_adv ? P.adv_stats[2 * mbi] : 0.f;
    const float adv_scale = P.norm_adv ? 1.0f / (P.adv_stats[2 * mbi + 1] + 1e-8f) : 1.f;
    float acc_w3a = 0.f, acc_w3b = 0.f;   // dW3 columns 2 lane, 2 lane + 1 of this warp's net (transposing butterfly, see ppo_grad16_kernel)
    float acc_b3 = 0.f, acc_logstd = 0.f;
    float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f, st4 = 0.f;
    unsigned tile = tile0;
    ETRACE(0);
    if (s == 1) asm volatile("bar.sync 7, 288;" ::: "memory");   // slot 1 starts once slot 0 has issued its first layer 2: half a tile behind

    for (unsigned j = 0; j < n_iter; ++j, tile += 2 * stride, ++round) {
        WTRACE(0);
        const float *in = sm->stage[round & 1u][s][tg];
        const bool live = tile < n_tiles && tile * kRows + row < n;
        const bool first = s == 0 && j == 0;   // this tile's batches initialise the weight-gradient accumulators
        prefetch_row(sm->next_idx[s][tg]);
        prefetch_index(static_cast<unsigned>(mbi), j + 2);
        if (j) {   // the slot's operands are rewritten below: its last batch (dW1) must have completed
            mbar_wait(bar, 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        // ---- S1: observation row -> X operand (1 in column 5: bias gradients = that column of dZ^T [X | 1]); issue layer 1 ----
        if (half == 0) {
            float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 1.f, 1.f, 0.f};   // columns 5, 6: the ones of the bias (hi, lo) rows of W1; column 5 also yields db
            if (live) {
#pragma unroll
                for (int k = 0; k < 5; ++k) x[k] = in[k];
            }
            *reinterpret_cast<uint4 *>(sm->a1[s][0][row]) = pack8(x);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(s);
        if (issuer && umma::elect_one()) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_f16(tcol, umma_desc(a1u, kRows * 16, 128), umma_desc(w1, 128 * 16, 128), id_l1, 0u);
            umma_commit(bar);
        }
        WTRACE(1);
        // ---- S2: H1 = tanh(D1 + b1) -> a2 (fp16); issue layer 2 (into D1's columns) ----
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WTRACE(2);
        {
            uint32_t ra[32], rb[32];
            tmem_ld32_issue(lane_base, ra);
            tmem_ld32_issue(lane_base + 32, rb);
            tmem_wait_ld();
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int col0 = half * 64 + blk * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int jj = col0 + c * 8;
                    float h[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) h[q] = fast_tanh(__uint_as_float(blk ? rb[c * 8 + q] : ra[c * 8 + q]));   // (b1 came with the MMA)
                    *reinterpret_cast<uint4 *>(a2p + (jj >> 3) * kLboH + row * 16u) = pack8(h);
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(s);
        if (issuer) {
            if (umma::elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tcol, umma_desc(a2u + k * 2 * kLboH, kLboH, 128), umma_desc(w2a + k * 2 * kHid * 16, kHid * 16, 128), id_l2, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tcol + 64, umma_desc(a2u + (8 + k * 2) * kLboH, kLboH, 128), umma_desc(w2c + k * 2 * kHid * 16, kHid * 16, 128), id_l2,
                             k > 0 ? 1u : 0u);
#pragma unroll
                for (int net = 0; net < 2; ++net)   // + b2: ones (k = 0, 1) times (hi, lo)
                    umma_f16(tcol + net * 64, umma_desc(smem_u32(sm->ones), kRows * 16, 128), umma_desc(smem_u32(sm->wb2[net]), kHid * 16, 128), id_l2, 1u);
                umma_commit(bar);
            }
            if (s == 0 && j == 0) {
                __syncwarp();
                asm volatile("bar.arrive 7, 288;" ::: "memory");
            }
        }
        WTRACE(3);
        // ---- S3: H2, head, clipped-PPO loss, dZ2 -> d2; issue D3 = dZ2 W2 (into D2's columns) and the weight gradients dW2, db2 ----
        mbar_wait(bar, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WTRACE(4);
        {
            const float s_a = live ? in[5] : 0.f, s_b = live ? in[6] : 0.f, s_c = live ? in[7] : 0.f;
            // H2 is kept as packed halves between the head and the backward pass (32 registers instead of 64: the slot's whole state must fit
            // 128 registers; the backward pass reads H1 in fp16 too), the head itself is summed from the fp32 values
            uint32_t hp[32];
            float head = W.b3[half];
            {
                uint32_t ra[32], rb[32];
                tmem_ld32_issue(lane_base, ra);
                tmem_ld32_issue(lane_base + 32, rb);
                tmem_wait_ld();
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const int col0 = half * 64 + blk * 32;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float h0 = fast_tanh(__uint_as_float(blk ? rb[2 * c] : ra[2 * c]));   // (b2 came with the MMAs)
                        const float h1 = fast_tanh(__uint_as_float(blk ? rb[2 * c + 1] : ra[2 * c + 1]));
                        head = fmaf(h0, W.w3[col0 + 2 * c], head);
                        head = fmaf(h1, W.w3[col0 + 2 * c + 1], head);
                        const __half2 hh = __floats2half2_rn(h0, h1);
                        hp[blk * 16 + c] = *reinterpret_cast<const uint32_t *>(&hh);
                    }
                }
            }
            // clipped-PPO loss; g = d(n * loss) / d head for this row (ppo_model_gsde.py:303-364, as in ppo_grad16_kernel)
            float g_head = 0.f;
            if (live) {
                if (half == 0) {
                    const float z = (s_a - head) * inv_std;
                    const float newlp = -0.5f * z * z - log_std - 0.91893853320467274178f;
                    const float logratio = newlp - s_b;
                    const float ratio = __expf(logratio);
                    const float A = (s_c - adv_mean) * adv_scale;
                    const float lo = 1.0f - eps, hi = 1.0f + eps;
                    const float rc = fminf(fmaxf(ratio, lo), hi);
                    const float pg1 = -A * ratio, pg2 = -A * rc;
                    const float in_range = (ratio >= lo && ratio <= hi) ? 1.f : 0.f;
                    const float g_ratio = pg1 > pg2 ? -A : (pg2 > pg1 ? -A * in_range : -A * (0.5f + 0.5f * in_range));
                    const float g_lp = g_ratio * ratio;
                    g_head = g_lp * z * inv_std;
                    acc_logstd += (g_lp * (z * z - 1.0f) - P.ent_coef) * inv_n;
                    st0 -= logratio;
                    st1 += (ratio - 1.0f) - logratio;
                    st2 += fabsf(ratio - 1.0f) > eps ? 1.f : 0.f;
                    st4 += fmaxf(pg1, pg2);
                } else {
                    const float du = head - s_a;
                    float g = du, vl = du * du;
                    if (P.clip_vloss) {
                        const float dv = head - s_b;
                        const float vc = s_b + fminf(fmaxf(dv, -eps), eps);
                        const float dc = vc - s_a, vcl = dc * dc;
                        const float in_range = (dv >= -eps && dv <= eps) ? 1.f : 0.f;
                        g = vl > vcl ? du : (vcl > vl ? dc * in_range : 0.5f * du + 0.5f * dc * in_range);
                        vl = fmaxf(vl, vcl);
                    }
                    g_head = P.vf_coef * g;
                    st3 += 0.5f * vl;
                }
            }
            const float dy = g_head * inv_n;
            acc_b3 += dy;
            // n * dZ2 = g * w3 * (1 - H2^2) -> d2 (fp16), in packed-half arithmetic (dZ2 is an fp16 tensor-core operand and H2 is held in fp16): three
            // HFMA2-pipe instructions per PAIR of columns (1 - h^2, g_head * w3, their product) instead of eight fp32 ones and a conversion.
            // dW3 = sum over rows of dy * H2: the thread's 32 packed products (g_head / 32) * h -- 1 / 32: the sum of a warp's 32 rows stays
            // within fp16's range whenever g_head itself does -- are summed over the warp's rows by a transposing shuffle butterfly on half2 words
            // (each level halves the word count and pairs lanes o apart; lane l ends with word l = columns 2l, 2l + 1) and leave the butterfly
            // as fp32: 31 shuffles per tile instead of the 62 of an fp32 butterfly, no conversions.
            uint32_t pw[16];
            {
                const bool upper = (lane & 16u) != 0u;
                const __half2 gh2 = __float2half2_rn(g_head), gs2 = __float2half2_rn(g_head * (1.0f / 32.0f)), one2 = __float2half2_rn(1.0f);
                auto h2u = [](const __half2 v) { return *reinterpret_cast<const uint32_t *>(&v); };
                auto u2h = [](const uint32_t v) { return *reinterpret_cast<const __half2 *>(&v); };
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pl[4], ph[4];
#pragma unroll
                    for (int hi = 0; hi < 2; ++hi) {
                        const int jj = half * 64 + hi * 32 + c * 8;
                        const uint4 w3q = *reinterpret_cast<const uint4 *>(&sm->w3h[jj >> 1]);   // 8 columns of w3 (same address in every lane)
                        const __half2 *w3p = reinterpret_cast<const __half2 *>(&w3q);
                        uint4 gq;
                        __half2 *gp = reinterpret_cast<__half2 *>(&gq);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const __half2 hq = u2h(hp[hi * 16 + c * 4 + q]);
                            gp[q] = __hmul2(__hmul2(gh2, w3p[q]), __hfma2(__hneg2(hq), hq, one2));
                            (hi ? ph : pl)[q] = h2u(__hmul2(gs2, hq));
                        }
                        *reinterpret_cast<uint4 *>(d2p + (jj >> 3) * kLboH + row * 16u) = gq;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q)   // first level: word e with word e + 16 (columns i and i + 32)
                        pw[c * 4 + q] = h2u(__hadd2(u2h(upper ? ph[q] : pl[q]), u2h(__shfl_xor_sync(0xffffffffu, upper ? pl[q] : ph[q], 16))));
                }
#pragma unroll
                for (int sz = 8; sz >= 1; sz >>= 1) {   // 16 -> 8 -> 4 -> 2 -> 1 words; lanes whose bit sz is set keep the upper half
                    const bool up = (lane & static_cast<unsigned>(sz)) != 0u;
#pragma unroll
                    for (int i = 0; i < sz; ++i) {
                        const uint32_t lo = pw[i], hi = pw[i + sz];
                        pw[i] = h2u(__hadd2(u2h(up ? hi : lo), u2h(__shfl_xor_sync(0xffffffffu, up ? lo : hi, sz))));
                    }
                }
                const float2 w3sum = __half22float2(u2h(pw[0]));
                acc_w3a += w3sum.x * (32.0f * inv_n);
                acc_w3b += w3sum.y * (32.0f * inv_n);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // D2 has been read by the whole group: D3 may overwrite it
        group_sync(s);
        WTRACE(5);
        if (issuer) {
            // token: the S3 batches of the two slots accumulate into the same dW2 / db2 columns, in the order 0, 1, 0, 1, ...
            if (s == 0) {
                if (j) token_wait(4);
            } else
                token_wait(3);
            if (umma::elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 4; ++k)   // D3 = dZ2 * W2 (per net): dZ2 read K-major (rows x features)
                    umma_f16(tcol, umma_desc(d2u + k * 2 * kLboH, kLboH, 128), umma_desc(w2ta + k * 2 * kHid * 16, kHid * 16, 128), id_l2, k > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_f16(tcol + 64, umma_desc(d2u + (8 + k * 2) * kLboH, kLboH, 128), umma_desc(w2tc + k * 2 * kHid * 16, kHid * 16, 128), id_l2,
                             k > 0 ? 1u : 0u);
                umma_commit(bar);   // S4's arithmetic starts on D3; its stores into d2 wait for the weight-gradient MMAs below (bar_mma2)
#pragma unroll
                for (int k = 0; k < 8; ++k)   // dW2 += dZ2^T * H1: the same buffers read MN-major (features x rows), 16 rows per MMA
                    umma_f16(tmem + 256, umma_desc(d2u + k * 256, 128, kLboH), umma_desc(a2u + k * 256, 128, kLboH), id_w2, (first && k == 0) ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < 8; ++k)   // db2 += dZ2^T * [X | 1]
                    umma_f16(tmem + 400, umma_desc(d2u + k * 256, 128, kLboH), umma_desc(a1u + k * 256, 128, kRows * 16), id_w1,
                             (first && k == 0) ? 0u : 1u);
                umma_commit(&sm->bar_mma2[s]);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            __syncwarp();
            if (s == 0)
                token_pass(3);
            else if (j + 1 < n_iter)
                token_pass(4);
        }
        WTRACE(6);
        // the next tile's row (index fetched one round ago) and the index of the tile after it travel under S4 -- every stage ends with
        // fence.proxy.async, which waits for the thread's outstanding loads, and S3 has no registers to spare -- and are parked before its fence
        const RowIn nr = load_row(sm->next_idx[s][tg]);
        const int64_t ni = load_index(static_cast<unsigned>(mbi), j + 2);
        // ---- S4: n * dZ1 = D3 * (1 - H1^2) -> d2 (every MMA that read dZ2 has completed); issue dW1 | db1 += dZ1^T [X | 1] ----
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WTRACE(7);
        {
            uint32_t ra[32], rb[32];
            uint4 gq[8];   // the thread's 64 columns of dZ1 as packed halves
            tmem_ld32_issue(lane_base, ra);
            tmem_ld32_issue(lane_base + 32, rb);
            tmem_wait_ld();
            const __half2 one2 = __float2half2_rn(1.0f);
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int col0 = half * 64 + blk * 32;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int jj = col0 + c * 8;
                    const uint4 raw = *reinterpret_cast<const uint4 *>(a2p + (jj >> 3) * kLboH + row * 16u);
                    const __half2 *hh = reinterpret_cast<const __half2 *>(&raw);
                    __half2 *gp = reinterpret_cast<__half2 *>(&gq[blk * 4 + c]);   // packed-half arithmetic, as in S3: D3 -> fp16, times (1 - h^2)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __half2 d = __floats2half2_rn(__uint_as_float(blk ? rb[c * 8 + 2 * q] : ra[c * 8 + 2 * q]),
                                                            __uint_as_float(blk ? rb[c * 8 + 2 * q + 1] : ra[c * 8 + 2 * q + 1]));
                        gp[q] = __hmul2(d, __hfma2(__hneg2(hh[q]), hh[q], one2));
                    }
                }
            }
            // dZ1 overwrites dZ2, which the slot's dW2 | db2 MMAs (issued behind D3 in the same batch) are still reading: they were given the time of
            // the arithmetic above; now they must have completed
            mbar_wait(&sm->bar_mma2[s], round & 1u);
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int jj = half * 64 + blk * 32 + c * 8;
                    *reinterpret_cast<uint4 *>(d2p + (jj >> 3) * kLboH + row * 16u) = gq[blk * 4 + c];
                }
        }
        park_row(sm->stage[(round & 1u) ^ 1u][s][tg], nr);
        sm->next_idx[s][tg] = ni;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(s);
        if (issuer) {
            if (s == 0) {
                if (j) token_wait(6);
            } else
                token_wait(5);
            if (umma::elect_one()) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_f16(tmem + 384, umma_desc(d2u + k * 256, 128, kLboH), umma_desc(a1u + k * 256, 128, kRows * 16), id_w1,
                             (first && k == 0) ? 0u : 1u);
                umma_commit(bar);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            __syncwarp();
            if (s == 0)
                token_pass(5);
            else if (j + 1 < n_iter)
                token_pass(6);
        }
        WTRACE(8);
    }
    mbar_wait(bar, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    ETRACE(1);
    __syncthreads();   // both slots have retired: the accumulators are final, the operand buffers are free
    ETRACE(2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- flush (per-CTA private gradient when `partials` is given, atomics otherwise) ----
    const bool own = partials != nullptr;
    PolicyGrad *G = own ? reinterpret_cast<PolicyGrad *>(partials) + blockIdx.x : P.grad;
    auto put = [&](float *dst, float val) {
        if (own) *dst = val;
        else atomicAdd(dst, val);
    };
    {
        // weight-gradient accumulators: TMEM lane = output feature (actor 0-63 | critic 64-127), scaled by n.  dW2: the block of the lane's own
        // net, 64 columns in four 16-column pieces (one per thread that can reach the lane: 2 slots x 2 halves); dW1 | db1 | db2: slot 0, half 0
        const unsigned net = row >> 6, jn = row & 63u;
        const uint32_t lane_acc = tmem + (((warp & 3u) * 32u) << 16);
        const unsigned piece = static_cast<unsigned>(s) * 2u + half;
        float v[32];
        tmem_ld32(lane_acc + 256 + net * 64 + (piece >> 1) * 32, v);   // 32 columns, this thread flushes 16 of them
        const unsigned c0 = (piece & 1u) * 16u;
        if (own) {
            float4 *dst = reinterpret_cast<float4 *>(&G->w2[net][jn][(piece >> 1) * 32 + c0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float4 o;
                o.x = (c0 ? v[16 + 4 * c] : v[4 * c]) * inv_n;
                o.y = (c0 ? v[17 + 4 * c] : v[4 * c + 1]) * inv_n;
                o.z = (c0 ? v[18 + 4 * c] : v[4 * c + 2]) * inv_n;
                o.w = (c0 ? v[19 + 4 * c] : v[4 * c + 3]) * inv_n;
                dst[c] = o;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) put(&G->w2[net][jn][(piece >> 1) * 32 + c0 + c], (c0 ? v[16 + c] : v[c]) * inv_n);
        }
        if (piece == 0) {
            tmem_ld32(lane_acc + 384, v);
#pragma unroll
            for (int k = 0; k < 5; ++k) put(&G->w1[net][jn][k], v[k] * inv_n);
            put(&G->b1[net][jn], v[5] * inv_n);
            put(&G->b2[net][jn], v[16 + 5] * inv_n);
        }
    }
    // dW3 and the scalar sums: the 16 warps meet in shared memory and are added in a fixed order (bit-reproducible gradient)
    {
        float *scr = reinterpret_cast<float *>(sm->a2);   // [16 warps][64] dW3 columns, then [16 warps][8] scalars
        scr[warp * 64 + 2 * lane] = acc_w3a;
        scr[warp * 64 + 2 * lane + 1] = acc_w3b;
        float sc[8] = {st0, st1, st2, st3, st4, acc_logstd, half == 0 ? acc_b3 : 0.f, half == 1 ? acc_b3 : 0.f};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float v = sc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) scr[16 * 64 + warp * 8 + q] = v;
        }
        __syncthreads();
        if (t < 128) {   // net of warp w: (w >> 2) & 1
            const unsigned net = t >> 6, c = t & 63u;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) a += scr[((k >> 2) * 8 + net * 4 + (k & 3)) * 64 + c];
            put(&G->w3[net][c], a);
        } else if (t < 136) {
            const unsigned q = t - 128;
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 16; ++w) a += scr[16 * 64 + w * 8 + q];
            if (q < 5) put(&G->stats[q], a);
            else if (q == 5) put(&G->logstd, a);
            else put(&G->b3[q - 6], a);
        }
    }
    if constexpr (EPOCH) {
        // ---- the optimiser step of this minibatch, inside the kernel ----
        float *scr = reinterpret_cast<float *>(sm->d2);   // both slots have retired: [0, per) reduced gradients, [per, per + 512) staging, 16 more
        const unsigned nG = gridDim.x;
        const int per = ((A.n_out + static_cast<int>(nG) - 1) / static_cast<int>(nG) + 63) & ~63;   // parameters (and statistics) a CTA owns
        const int base = static_cast<int>(blockIdx.x) * per;
        // Adam's bias corrections (as torch.optim.Adam computes them) -- two powf that depend on the step number only
        const float step = step0 + static_cast<float>(mbi + 1);
        float step_size = 0.f, inv_sqrt_bc2 = 0.f;
        if (static_cast<int>(t) < per) {   // only the warps that own parameters pay for the two powf
            const float bc1 = 1.0f - powf(A.beta1, step), bc2 = 1.0f - powf(A.beta2, step);
            step_size = A.lr[0] / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
        }
        // the owners' optimiser state (this CTA wrote it one step ago) is requested before the meeting
        float m = 0.f, v = 0.f, vm = 0.f, prm = 0.f;
        int4 dst = make_int4(-1, -1, -1, -1);
        if (static_cast<int>(t) < per && base + static_cast<int>(t) < A.n) {
            const int p = base + static_cast<int>(t);
            m = A.exp_avg[p], v = A.exp_avg_sq[p], vm = A.max_exp_avg_sq[p], prm = A.param[p];
            dst = reinterpret_cast<const int4 *>(A.scatter)[p];
        }
        ETRACE(3);
        grid_meet(3u * mbi + 1u);   // every CTA's private gradient is in L2
        ETRACE(4);
        for (int c0 = 0; c0 < per; c0 += 64) {   // 64 parameters x 8 groups of partial rows per pass, groups added in a fixed order
            const int p = base + c0 + static_cast<int>(t & 63u);
            const unsigned g = t >> 6;
            float sum = 0.f;
            if (p < A.n_out) {   // the group's rows b = g, g + 8, ...: all loads of a batch of 20 in flight, then added in row order
                const float *src = partials + p;
                for (unsigned b0 = g; b0 < nG; b0 += 160) {
                    float vals[20];
#pragma unroll
                    for (int k = 0; k < 20; ++k) {
                        const unsigned b = b0 + 8u * k;
                        vals[k] = b < nG ? __ldcg(src + static_cast<size_t>(b) * A.n_out) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 20; ++k) sum += vals[k];
                }
            }
            scr[per + t] = sum;
            __syncthreads();
            if (t < 64) {
                float gsum = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k) gsum += scr[per + k * 64 + t];
                if (A.bufs && p < A.n_pad) {
                    // the ranks' mean over NVLink peer memory: the exchange of ppo_reduce_adam_kernel<true> ((value, call) words pushed to every
                    // rank's buffer, own buffer polled, rows added in rank order), with the same call numbering per 32-parameter block (A.seq), so
                    // f16_ppo_step16 and f16_ppo_epoch16 calls can be mixed on one set of exchange buffers.  (p >> 5 is warp-uniform.)
                    const int blk = p >> 5;
                    const uint32_t call = A.seq[blk] + 1u;
                    const size_t slot_at = static_cast<size_t>(call & 1u) * A.world * A.n_pad;
                    const size_t mine = slot_at + static_cast<size_t>(A.rank) * A.n_pad + p;
                    for (int r = 1; r < A.world; ++r) {
                        uint2 *to = A.bufs[(A.rank + r) % A.world] + mine;
                        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(to), "r"(__float_as_uint(gsum)), "r"(call) : "memory");
                    }
                    const uint2 *own_buf = A.bufs[A.rank] + slot_at + p;
                    float rsum = 0.f;
                    unsigned long long t0 = 0ull;
                    for (int r = 0; r < A.world; ++r) {
                        float val = gsum;
                        if (r != A.rank) {
                            const uint2 *src = own_buf + static_cast<size_t>(r) * A.n_pad;
                            uint32_t bits, flag;
                            for (unsigned spin = 0;; ++spin) {
                                asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(flag) : "l"(src) : "memory");
                                if (flag == call) break;
                                if ((spin & 1023u) == 1023u) {   // a rank that never arrives: fail loudly after ~10 s
                                    unsigned long long now;
                                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                                    if (t0 == 0ull) t0 = now;
                                    else if (now - t0 > 10000000000ull) __trap();
                                }
                            }
                            val = __uint_as_float(bits);
                        }
                        rsum += val;
                    }
                    gsum = rsum * (1.0f / static_cast<float>(A.world));
                    __syncwarp();
                    if (lane == 0) A.seq[blk] = call;
                }
                scr[c0 + t] = gsum;
            }
            __syncthreads();
        }
        {   // clip_grad_norm_: this CTA's share of the squared norm
            float sq = 0.f;
            for (int i = static_cast<int>(t); i < per; i += kThreadsW)
                if (base + i < A.n) sq += scr[i] * scr[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) scr[per + 512 + warp] = sq;
            __syncthreads();
            if (t == 0) {
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < 16; ++w) tot += scr[per + 512 + w];
                reinterpret_cast<float *>(A.sync + 8)[blockIdx.x] = tot;
            }
        }
        ETRACE(5);
        grid_meet(3u * mbi + 2u);   // every CTA's share is posted
        const float *sq_all = reinterpret_cast<const float *>(A.sync + 8);
        ETRACE(6);
        // (only the warps that own parameters: every warp of every CTA reading the same five lines of L2 took 1.5 us)
        float tot = 0.f;
        if (static_cast<int>(warp) * 32 < per)
        for (unsigned c0 = lane; c0 < nG; c0 += 160) {   // five loads in flight per lane, added in a fixed order
            float sv[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) sv[k] = c0 + 32u * k < nG ? __ldcg(sq_all + c0 + 32u * k) : 0.f;
#pragma unroll
            for (int k = 0; k < 5; ++k) tot += sv[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        float coef = A.max_grad_norm / (sqrtf(tot) + 1e-6f);
        coef = coef < 1.0f ? coef : 1.0f;
        // Adam(amsgrad), as torch.optim.Adam computes it (the arithmetic of ppo_reduce_adam_kernel)
        for (int i = static_cast<int>(t); i < per; i += kThreadsW) {
            const int p = base + i;
            if (p >= A.n_out) break;
            const float grad = scr[i];
            A.grad[p] = grad;
            float pn = 0.f;
            if (p < A.n) {
                if (i >= kThreadsW) {   // (only when a CTA owns more than 512 parameters: small grids)
                    m = A.exp_avg[p], v = A.exp_avg_sq[p], vm = A.max_exp_avg_sq[p], prm = A.param[p];
                    dst = reinterpret_cast<const int4 *>(A.scatter)[p];
                }
                const float gc = grad * coef;
                const float mn = m + (gc - m) * (1.0f - A.beta1);
                const float vn = v * A.beta2 + (1.0f - A.beta2) * gc * gc;
                const float vmax = fmaxf(vm, vn);
                const float denom = sqrtf(vmax) * inv_sqrt_bc2 + A.eps;
                pn = prm - step_size * (mn / denom);
                A.exp_avg[p] = mn;
                A.exp_avg_sq[p] = vn;
                A.max_exp_avg_sq[p] = vmax;
                A.param[p] = pn;
                if (dst.x >= 0) A.fwd_blob[dst.x] = pn;
                if (dst.y >= 0) A.tail_blob[dst.y] = pn;
                if (dst.z >= 0) A.half_blob[dst.z] = __float2half_rn(pn);
                if (dst.w >= 0) A.half_blob[dst.w] = __float2half_rn(pn);
            }
            if (A.diag) {   // the five loss statistics sit right behind the parameters, log-std is the last parameter
                const int k = p - A.n;
                if (k >= 0 && k < 5) {
                    if (k == 2) A.diag[2] += grad * A.inv_rows;
                    else A.diag[k] = grad * A.inv_rows;
                }
                if (p == A.n - 1) A.diag[5] = pn + 1.4189385332046727f;   // Gaussian entropy 0.5 + 0.5 log(2 pi) + log std (after the step)
            }
        }
        if (blockIdx.x == 0 && t == 0) A.step[0] = step;
        ETRACE(7);
        grid_meet(3u * mbi + 3u);   // the blobs are complete
        ETRACE(8);
        if (mbi + 1 < n_mb) {   // the new weights: L2 -> shared memory (plain loads: the blob was written through the generic proxy a moment ago)
            const uint4 *src = reinterpret_cast<const uint4 *>(blob16);
            uint4 *dstw = reinterpret_cast<uint4 *>(&sm->w);
            constexpr unsigned n16 = sizeof(TrainBlob16) / 16, n_it = (n16 + kThreadsW - 1) / kThreadsW;
            uint4 wv[n_it];
#pragma unroll
            for (unsigned k = 0; k < n_it; ++k)
                if (t + k * kThreadsW < n16) wv[k] = __ldcg(src + t + k * kThreadsW);
#pragma unroll
            for (unsigned k = 0; k < n_it; ++k)
                if (t + k * kThreadsW < n16) dstw[t + k * kThreadsW] = wv[k];
            __syncthreads();
            derive_weights();
        }
        ETRACE(9);
    }
  }   // minibatches

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// keyed bijection of [0, n): see f16_ppo_randperm in the header.  Every round is invertible on `bits` bits (odd multiplier, right
// xor-shift, addition), so the composition is a permutation of [0, 2^bits); values >= n walk the cycle until they fall below n.
struct PermKey { uint64_t mul[4], add[4]; int shift[4]; int bits; };
__global__ void __launch_bounds__(256) ppo_randperm_kernel(int64_t n, PermKey K, int64_t *__restrict__ out)
{
    const uint64_t mask = (K.bits >= 64) ? ~0ull : ((1ull << K.bits) - 1ull);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        uint64_t x = static_cast<uint64_t>(i);
        do {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                x = (x * K.mul[r]) & mask;
                x ^= x >> K.shift[r];
                x = (x + K.add[r]) & mask;
            }
        } while (x >= static_cast<uint64_t>(n));
        out[i] = static_cast<int64_t>(x);
    }
}

// once per update: the six arrays a minibatch row is gathered from -> one 32-byte record + one 8-byte record per rollout row
__global__ void __launch_bounds__(256) ppo_pack_rows_kernel(int64_t n, const float *__restrict__ obs, const float *__restrict__ actions,
                                                            const float *__restrict__ logprobs, const float *__restrict__ adv,
                                                            const float *__restrict__ ret, const float *__restrict__ val,
                                                            float4 *__restrict__ rows, float2 *__restrict__ rv)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float *o = obs + i * 5;
        rows[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
        rows[2 * i + 1] = make_float4(o[4], actions[i], logprobs[i], adv[i]);
        rv[i] = make_float2(ret[i], val[i]);
    }
}

// sum the CTAs' private gradients: thread (output p, group g of 8) adds every 8th row, the 8 groups meet in shared memory
__global__ void __launch_bounds__(256) ppo_grad_reduce_kernel(const float *__restrict__ partials, int n_rows, int n_out, float *__restrict__ out)
{
    __shared__ float part[8][32];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned lane = threadIdx.x & 31u, g = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (p < n_out)
        for (int b = g; b < n_rows; b += 8) s += partials[static_cast<size_t>(b) * n_out + p];
    part[g][lane] = s;
    __syncthreads();
    if (g == 0 && p < n_out) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += part[k][lane];
        out[p] = tot;
    }
}

// The same reduction fused with the cross-rank gradient all-reduce (the trainer's only collective: ~37 KB per minibatch, pure latency
// for NCCL: 20-29 us per call on 2-8 B200s, a fifth of the update).  Every rank owns an exchange buffer that all ranks of the node have
// mapped (CUDA IPC over NVLink / NVSwitch peer memory): words[2 slots][world][n_pad], a word = (gradient element, call number) in ONE
// 8-byte store -- the "LL" idea: the payload carries its own flag, an aligned 8-byte store arrives whole, so there is no fence, no
// separate flag and no round trip.  CTA c reduces its 32 gradient elements over the local CTAs' partials, PUSHES (value, call) into
// slot (call & 1), row `rank`, of every rank's buffer (posted NVLink writes), then polls the `world` rows of its OWN buffer until each
// word shows this call and adds the values in rank order -- every rank adds the same numbers in the same order, so the parameters
// stay bit-identical across ranks -- times 1 / world.  One NVLink one-way latency per call.  The call counters live in device memory,
// so the launch replays from a CUDA graph.  Two slots make reuse safe: a rank starts call k + 2 only after it has received every
// peer's call k + 1, which a peer sends after it has finished reading call k.
// (First version, measured on 2 B200s: separate flags behind __threadfence_system() + st.release.sys = two fabric round trips,
// +16.7 us per call -- no better than NCCL's +12.4 us.)
__global__ void __launch_bounds__(256) ppo_grad_reduce_peer_kernel(const float *__restrict__ partials, int n_rows, int n_out, float *__restrict__ out,
                                                                   uint2 *const *__restrict__ bufs, uint32_t *__restrict__ seq, int rank, int world,
                                                                   int n_pad)
{
    __shared__ float part[8][32];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned lane = threadIdx.x & 31u, g = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;           // p < n_pad always: the buffers are padded to whole CTAs
    float s = 0.f;
    if (p < n_out)
        for (int b = g; b < n_rows; b += 8) s += partials[static_cast<size_t>(b) * n_out + p];
    part[g][lane] = s;
    __syncthreads();
    if (g != 0) return;
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][lane];
    const uint32_t call = seq[blockIdx.x] + 1u;
    const size_t slot_at = static_cast<size_t>(call & 1u) * world * n_pad;
    const size_t mine = slot_at + static_cast<size_t>(rank) * n_pad + p;
    for (int r = 1; r < world; ++r) {               // the peers, each rank starting with a different one
        uint2 *dst = bufs[(rank + r) % world] + mine;
        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(tot)), "r"(call) : "memory");
    }
    const uint2 *own = bufs[rank] + slot_at + p;
    float sum = 0.f;
    unsigned long long t0 = 0ull;
    for (int r = 0; r < world; ++r) {
        float v = tot;
        if (r != rank) {
            const uint2 *src = own + static_cast<size_t>(r) * n_pad;
            uint32_t bits, flag;
            for (unsigned spin = 0;; ++spin) {
                asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(flag) : "l"(src) : "memory");
                if (flag == call) break;
                if ((spin & 1023u) == 1023u) {      // a rank that never arrives must not hang the GPU: fail loudly after ~10 s
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t0 == 0ull) t0 = now;
                    else if (now - t0 > 10000000000ull) __trap();
                }
            }
            v = __uint_as_float(bits);
        }
        sum += v;
    }
    if (p < n_out) out[p] = sum * (1.0f / static_cast<float>(world));
    __syncwarp();
    if (lane == 0) seq[blockIdx.x] = call;
}

// ---------------------------------------------------------------------------
// ppo_reduce_adam_kernel: the gradient reduction (and, with PEER, the cross-rank exchange above) FUSED with the optimiser step --
// clip_grad_norm_ + Adam(amsgrad) + re-packing of the weight blobs -- in one multi-CTA launch.  The single-CTA ppo_adam_kernel is a chain
// of dependent latencies (gradient loads, CTA-wide norm, state loads, ~45 k gathered blob words: 13 us) behind a launch of its own; here
// every CTA owns 32 parameters end to end: it reduces their gradients, contributes its partial sum of squares, meets the other CTAs at ONE
// grid-wide arrival counter (the grid is 289 small CTAs, all resident once the gradient kernel has left), adds the partials in a fixed
// order (the norm is bit-identical in every CTA and on every rank), updates its parameters and Adam state (loaded BEFORE the grid
// dependency resolved), and SCATTERS each new parameter to its <= 4 places in the blobs through an inverse table instead of gathering.
// ---------------------------------------------------------------------------
template <bool PEER>
__global__ void __launch_bounds__(256) ppo_reduce_adam_kernel(const ReduceAdamArgs A)
{
    __shared__ float part[8][32];
    const unsigned lane = threadIdx.x & 31u, g = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const bool is_param = p < A.n;
    // ---- before the grid dependency resolves: this slice's optimiser state and scatter table (written by the PREVIOUS call of this kernel,
    // which had completed before the gradient kernel in between released this launch: that kernel triggers its dependents after its own wait)
    float m = 0.f, v = 0.f, vm = 0.f, prm = 0.f;
    int4 dst = make_int4(-1, -1, -1, -1);
    if (g == 0 && is_param) {
        m = A.exp_avg[p];
        v = A.exp_avg_sq[p];
        vm = A.max_exp_avg_sq[p];
        prm = A.param[p];
        dst = reinterpret_cast<const int4 *>(A.scatter)[p];
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float s = 0.f;
    if (p < A.n_out)
        for (int b = g; b < A.n_rows; b += 8) s += A.partials[static_cast<size_t>(b) * A.n_out + p];
    part[g][lane] = s;
    __syncthreads();
    if (g != 0) return;
    float grad = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) grad += part[k][lane];
    uint32_t *calls = reinterpret_cast<uint32_t *>(A.sync + 8 + sizeof(float) * gridDim.x);
    const uint32_t call = calls[blockIdx.x] + 1u;
    unsigned long long t0 = 0ull;
    auto impatient = [&](unsigned spin) {   // a CTA / rank that never arrives must not hang the GPU: fail loudly after ~10 s
        if ((spin & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0ull) t0 = now;
            else if (now - t0 > 10000000000ull) __trap();
        }
    };
    if constexpr (PEER) {   // see ppo_grad_reduce_peer_kernel; the exchange numbers its calls in peer->seq, shared with that kernel
        const uint32_t call = A.seq[blockIdx.x] + 1u;
        const size_t slot_at = static_cast<size_t>(call & 1u) * A.world * A.n_pad;
        const size_t mine = slot_at + static_cast<size_t>(A.rank) * A.n_pad + p;
        for (int r = 1; r < A.world; ++r) {
            uint2 *to = A.bufs[(A.rank + r) % A.world] + mine;
            asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(to), "r"(__float_as_uint(grad)), "r"(call) : "memory");
        }
        const uint2 *own = A.bufs[A.rank] + slot_at + p;
        float sum = 0.f;
        for (int r = 0; r < A.world; ++r) {
            float val = grad;
            if (r != A.rank) {
                const uint2 *src = own + static_cast<size_t>(r) * A.n_pad;
                uint32_t bits, flag;
                for (unsigned spin = 0;; ++spin) {
                    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(flag) : "l"(src) : "memory");
                    if (flag == call) break;
                    impatient(spin);
                }
                val = __uint_as_float(bits);
            }
            sum += val;
        }
        grad = sum * (1.0f / static_cast<float>(A.world));
        __syncwarp();
        if (lane == 0) A.seq[blockIdx.x] = call;
    }
    if (p < A.n_out) A.grad[p] = grad;
    // ---- clip_grad_norm_: this CTA's share of the squared norm, then the grid-wide meeting ----
    float sq = is_param ? grad * grad : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    float *sq_all = reinterpret_cast<float *>(A.sync + 8);
    unsigned long long *arrive = reinterpret_cast<unsigned long long *>(A.sync);
    const float step = A.step[0] + 1.0f;           // read before arriving: CTA 0 writes the new value after the meeting
    const float lr = A.lr[0];
    if (lane == 0) {
        sq_all[blockIdx.x] = sq;
        __threadfence();
        atomicAdd(arrive, 1ull);
        const unsigned long long want = static_cast<unsigned long long>(call) * gridDim.x;
        for (unsigned spin = 0;; ++spin) {
            unsigned long long seen;
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(arrive) : "memory");
            if (seen >= want) break;
            impatient(spin);
        }
    }
    __syncwarp();
    float tot = 0.f;
    for (unsigned c = lane; c < gridDim.x; c += 32) tot += __ldcg(sq_all + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    float coef = A.max_grad_norm / (sqrtf(tot) + 1e-6f);
    coef = coef < 1.0f ? coef : 1.0f;
    // ---- Adam(amsgrad), as torch.optim.Adam computes it ----
    const float bc1 = 1.0f - powf(A.beta1, step), bc2 = 1.0f - powf(A.beta2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    float pn = prm;
    if (is_param) {
        const float gc = grad * coef;
        const float mn = m + (gc - m) * (1.0f - A.beta1);
        const float vn = v * A.beta2 + (1.0f - A.beta2) * gc * gc;
        const float vmax = fmaxf(vm, vn);
        const float denom = sqrtf(vmax) * inv_sqrt_bc2 + A.eps;
        pn = prm - step_size * (mn / denom);
        A.exp_avg[p] = mn;
        A.exp_avg_sq[p] = vn;
        A.max_exp_avg_sq[p] = vmax;
        A.param[p] = pn;
        if (dst.x >= 0) A.fwd_blob[dst.x] = pn;
        if (dst.y >= 0) A.tail_blob[dst.y] = pn;
        if (dst.z >= 0) A.half_blob[dst.z] = __float2half_rn(pn);
        if (dst.w >= 0) A.half_blob[dst.w] = __float2half_rn(pn);
    }
    if (A.diag) {   // the five loss statistics sit right behind the parameters, log-std is the last parameter
        const int k = p - A.n;
        if (k >= 0 && k < 5) {
            if (k == 2) A.diag[2] += grad * A.inv_rows;
            else A.diag[k] = grad * A.inv_rows;
        }
        if (p == A.n - 1) A.diag[5] = pn + 1.4189385332046727f;   // Gaussian entropy 0.5 + 0.5 log(2 pi) + log std (after the step)
    }
    __syncwarp();
    if (lane == 0) {
        calls[blockIdx.x] = call;
        if (blockIdx.x == 0) A.step[0] = step;
    }
}

// torch.nn.utils.clip_grad_norm_ (L2, max_norm) followed by torch.optim.Adam(amsgrad=True) on the flat parameter vector,
// then both weight blobs are re-packed from the updated parameters: what torch runs as ~60 launches per minibatch.
constexpr int kAdamMaxParams = 10240;   // updated parameters are staged in shared memory for the blob gathers (40 KB)

// The gather permutations are constants of the training run: with cache_perms they are staged in shared memory as 16-bit indices BEFORE
// the grid dependency resolves (i.e. under the tail of the gradient reduction this launch is chained to), so that the re-packing after the
// optimiser step is a shared-memory gather instead of ~45 k dependent 8-byte loads from L2 by one SM (it was half of the kernel's 14 us).
__global__ void __launch_bounds__(1024) ppo_adam_kernel(const PpoAdamArgs A, const int cache_perms)
{
    __shared__ float warp_sum[32];
    __shared__ float s_coef;
    extern __shared__ __align__(16) unsigned char adam_smem[];
    float *s_param = reinterpret_cast<float *>(adam_smem);                                  // [kAdamMaxParams]
    uint16_t *s_perm = reinterpret_cast<uint16_t *>(adam_smem + kAdamMaxParams * sizeof(float));   // fwd | train | half, when cache_perms
    const unsigned t = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (cache_perms) {
        for (int k = t; k < A.n_fwd; k += 1024) s_perm[k] = static_cast<uint16_t>(A.fwd_perm[k]);
        for (int k = t; k < A.n_train; k += 1024) s_perm[A.n_fwd + k] = static_cast<uint16_t>(A.train_perm[k]);
        for (int k = t; k < A.n_half; k += 1024) s_perm[A.n_fwd + A.n_train + k] = static_cast<uint16_t>(A.half_perm[k]);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float ss = 0.f;
    for (int i = t; i < A.n; i += 1024) {
        const float g = A.grad[i];
        ss = fmaf(g, g, ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((t & 31u) == 0) warp_sum[t >> 5] = ss;
    __syncthreads();
    if (t < 32) {
        float v = warp_sum[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (t == 0) {
            const float total = sqrtf(v);
            const float c = A.max_grad_norm / (total + 1e-6f);      // clip_grad_norm_: clip_coef clamped to 1
            s_coef = c < 1.0f ? c : 1.0f;
        }
    }
    __syncthreads();
    const float coef = s_coef;
    const float step = A.step[0] + 1.0f;
    const float lr = A.lr[0];
    const float bc1 = 1.0f - powf(A.beta1, step), bc2 = 1.0f - powf(A.beta2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const bool staged = A.n <= kAdamMaxParams;
    // batches of 4 independent elements per thread: all loads of a batch are issued before its first store (without the
    // explicit batching the stores serialise the iterations on global-memory latency: 29 us for 9.3 k parameters)
    for (int i0 = t; i0 < A.n; i0 += 4 * 1024) {
        float g[4], m[4], v[4], vm[4], p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 1024;
            const bool ok = i < A.n;
            g[u] = ok ? A.grad[i] : 0.f;
            m[u] = ok ? A.exp_avg[i] : 0.f;
            v[u] = ok ? A.exp_avg_sq[i] : 0.f;
            vm[u] = ok ? A.max_exp_avg_sq[i] : 0.f;
            p[u] = ok ? A.param[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 1024;
            if (i < A.n) {
                const float gc = g[u] * coef;
                const float mn = m[u] + (gc - m[u]) * (1.0f - A.beta1);                     // lerp, as torch does
                const float vn = v[u] * A.beta2 + (1.0f - A.beta2) * gc * gc;
                const float vmax = fmaxf(vm[u], vn);
                const float denom = sqrtf(vmax) * inv_sqrt_bc2 + A.eps;
                const float pn = p[u] - step_size * (mn / denom);
                A.exp_avg[i] = mn;
                A.exp_avg_sq[i] = vn;
                A.max_exp_avg_sq[i] = vmax;
                A.param[i] = pn;
                if (staged) s_param[i] = pn;
            }
        }
    }
    __syncthreads();   // the CTA's own writes (shared and global) are visible to it after the barrier
    if (t == 0) A.step[0] = step;
    const float *src = staged ? s_param : A.param;
    auto repack = [&](const int64_t *perm, const uint16_t *cached, float *blob, int count) {
        for (int k0 = t; k0 < count; k0 += 8 * 1024) {
            int q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u * 1024;
                q[u] = k < count ? (cache_perms ? static_cast<int>(cached[k]) : static_cast<int>(perm[k])) : -1;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int k = k0 + u * 1024;
                if (q[u] >= 0) blob[k] = q[u] ? src[q[u] - 1] : 0.f;
            }
        }
    };
    repack(A.fwd_perm, s_perm, A.fwd_blob, A.n_fwd);
    repack(A.train_perm, s_perm + A.n_fwd, A.train_blob, A.n_train);
    {
        const uint16_t *cached = s_perm + A.n_fwd + A.n_train;
        __half2 *dst2 = reinterpret_cast<__half2 *>(A.half_blob);
        if (cache_perms && (A.n_half & 1) == 0 && (reinterpret_cast<uintptr_t>(A.half_blob) & 3u) == 0) {
            // two fp16 words per store
            for (int k = t; k < A.n_half / 2; k += 1024) {
                const int q0 = cached[2 * k], q1 = cached[2 * k + 1];
                dst2[k] = __floats2half2_rn(q0 ? src[q0 - 1] : 0.f, q1 ? src[q1 - 1] : 0.f);
            }
        } else {
            for (int k0 = t; k0 < A.n_half; k0 += 8 * 1024) {
                int q[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 + u * 1024;
                    q[u] = k < A.n_half ? (cache_perms ? static_cast<int>(cached[k]) : static_cast<int>(A.half_perm[k])) : -1;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 + u * 1024;
                    if (q[u] >= 0) A.half_blob[k] = __float2half_rn(q[u] ? src[q[u] - 1] : 0.f);
                }
            }
        }
    }
    if (A.diag && t == 0) {
        const float *st = A.grad + A.n;
        A.diag[0] = st[0] * A.inv_rows;
        A.diag[1] = st[1] * A.inv_rows;
        A.diag[2] += st[2] * A.inv_rows;
        A.diag[3] = st[3] * A.inv_rows;
        A.diag[4] = st[4] * A.inv_rows;
        A.diag[5] = A.logstd[0] + 1.4189385332046727f;   // Gaussian entropy 0.5 + 0.5 log(2 pi) + log std (after the step)
    }
    if (A.zero_grad) {
        __syncthreads();   // thread 0 has read the statistics
        for (int i = t; i < A.n + 5; i += 1024) A.grad_rw[i] = 0.f;
    }
}

// Advantage normalisation statistics of one minibatch (ppo_model_gsde.py:318-321: mb_advantages.mean(), .std()): gather + sum +
// sum of squares in one pass (double accumulators), the last CTA to finish turns them into (mean, unbiased std) and re-arms the
// scratch.  Replaces torch's index + std_mean + two copies.
__global__ void __launch_bounds__(256) adv_stats_kernel(const float *__restrict__ adv, const int64_t *__restrict__ mb_idx, int64_t n,
                                                        float *out, double *scratch)
{
    // blockIdx.y = minibatch: rows [y * n, (y + 1) * n) of mb_idx, out[y][2], scratch[y][4] (a whole epoch in one launch)
    mb_idx += static_cast<int64_t>(blockIdx.y) * n;
    out += blockIdx.y * 2;
    scratch += blockIdx.y * 4;
    double s = 0.0, ss = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = static_cast<double>(adv[mb_idx[i]]);
        s += a;
        ss += a * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    __shared__ double ws[8], wss[8];
    __shared__ bool last;
    if ((threadIdx.x & 31u) == 0) {
        ws[threadIdx.x >> 5] = s;
        wss[threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            a += ws[w];
            b += wss[w];
        }
        atomicAdd(&scratch[0], a);
        atomicAdd(&scratch[1], b);
        __threadfence();
        const unsigned done = atomicAdd(reinterpret_cast<unsigned *>(&scratch[2]), 1u);
        last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const double sum = atomicAdd(&scratch[0], 0.0), sq = atomicAdd(&scratch[1], 0.0);
        const double mean = sum / static_cast<double>(n);
        const double var = n > 1 ? (sq - sum * mean) / static_cast<double>(n - 1) : 0.0;
        out[0] = static_cast<float>(mean);
        out[1] = static_cast<float>(sqrt(var > 0.0 ? var : 0.0));
        scratch[0] = 0.0;
        scratch[1] = 0.0;
        *reinterpret_cast<unsigned *>(&scratch[2]) = 0u;
    }
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
        // the scan is serial in t but its loads are not: eight steps' rows are requested together (24 independent loads in flight per
        // thread: with one thread per env -- 65,536 in config 4 -- the kernel is bound by memory latency, not by the dependent FMAs)
        constexpr int U = 8;
        int t = T - 1;
        for (; t >= U - 1; t -= U) {
            float r[U], v[U];
            uint8_t d[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t k = (int64_t)(t - u) * n + i;
                r[u] = __ldg(rewards + k);
                v[u] = __ldg(values + k);
                d[u] = __ldg(done_after + k);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t k = (int64_t)(t - u) * n + i;
                const float nextnonterminal = d[u] ? 0.f : 1.f;
                const float delta = r[u] + gamma * nextvalues * nextnonterminal - v[u];
                lastgaelam = delta + gamma * lam * nextnonterminal * lastgaelam;
                advantages[k] = lastgaelam;
                returns[k] = lastgaelam + v[u];
                nextvalues = v[u];
            }
        }
        for (; t >= 0; --t) {
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

#ifdef F16_GRAD_TRACE
extern "C" int f16_debug_epoch_cta(unsigned long long *out)
{
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_epoch_cta, sizeof(g_epoch_cta)));
}
extern "C" int f16_debug_epoch_trace(unsigned long long *out)
{
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_epoch_trace, sizeof(g_epoch_trace)));
}
extern "C" int f16_debug_grad_trace(unsigned long long *out)
{
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_grad_trace, sizeof(g_grad_trace)));
}
#endif


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

cudaError_t launch_adv_stats(const float *adv, const int64_t *mb_idx, int64_t n, int n_minibatches, float *out, double *scratch, int n_sm,
                             cudaStream_t st)
{
    const int64_t blocks = (n + 255) / 256;
    int per_mb = (2 * n_sm + n_minibatches - 1) / n_minibatches;   // about two waves in total
    if (per_mb < 1) per_mb = 1;
    const dim3 grid(static_cast<unsigned>(blocks < per_mb ? blocks : per_mb), static_cast<unsigned>(n_minibatches));
    adv_stats_kernel<<<grid, 256, 0, st>>>(adv, mb_idx, n, out, scratch);
    return cudaGetLastError();
}

cudaError_t launch_ppo_adam(const PpoAdamArgs &A, cudaStream_t st)
{
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(1);
    lc.blockDim = dim3(1024);
    lc.stream = st;
    lc.attrs = attr;
    lc.numAttrs = 1;
    // shared memory: the updated parameters (40 KB) + the three gather permutations as 16-bit indices, when they fit
    const size_t n_perm = static_cast<size_t>(A.n_fwd) + static_cast<size_t>(A.n_train) + static_cast<size_t>(A.n_half);
    const size_t base = kAdamMaxParams * sizeof(float);
    const int cache = (A.n < 65535 && base + n_perm * 2 <= 200 * 1024) ? 1 : 0;
    lc.dynamicSmemBytes = base + (cache ? (n_perm * 2 + 15) / 16 * 16 : 0);
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e0 = cudaFuncSetAttribute(ppo_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 16);
        if (e0 != cudaSuccess) return e0;
        configured[dev] = true;
    }
    cudaError_t e = cudaLaunchKernelEx(&lc, ppo_adam_kernel, A, cache);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_ppo_randperm(int64_t n, uint64_t seed, int64_t *out, int n_sm, cudaStream_t st)
{
    PermKey K;
    K.bits = 1;
    while (K.bits < 63 && (1ll << K.bits) < n) ++K.bits;
    uint64_t z = seed;
    auto next = [&z]() {   // splitmix64
        z += 0x9E3779B97F4A7C15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        return x ^ (x >> 31);
    };
    for (int r = 0; r < 4; ++r) {
        K.mul[r] = next() | 1ull;
        K.add[r] = next();
        const int half = K.bits / 2 > 0 ? K.bits / 2 : 1;
        K.shift[r] = half + static_cast<int>(next() % 3) - (half > 1 ? 1 : 0);   // around half the width
        if (K.shift[r] < 1) K.shift[r] = 1;
    }
    const int64_t want = (n + 255) / 256;
    const int grid = static_cast<int>(want < static_cast<int64_t>(n_sm) * 8 ? want : static_cast<int64_t>(n_sm) * 8);
    ppo_randperm_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(n, K, out);
    return cudaGetLastError();
}

cudaError_t launch_ppo_pack_rows(int64_t n, const float *obs, const float *actions, const float *logprobs, const float *adv, const float *ret,
                                 const float *val, float *rows, float *rv, int n_sm, cudaStream_t st)
{
    const int64_t want = (n + 255) / 256;
    const int grid = static_cast<int>(want < static_cast<int64_t>(n_sm) * 8 ? want : static_cast<int64_t>(n_sm) * 8);
    ppo_pack_rows_kernel<<<grid > 0 ? grid : 1, 256, 0, st>>>(n, obs, actions, logprobs, adv, ret, val, reinterpret_cast<float4 *>(rows),
                                                             reinterpret_cast<float2 *>(rv));
    return cudaGetLastError();
}

int ppo_grad_floats_padded() { return (static_cast<int>(sizeof(PolicyGrad) / sizeof(float)) + 31) / 32 * 32; }
int ppo_grad_reduce_ctas() { return (static_cast<int>(sizeof(PolicyGrad) / sizeof(float)) + 31) / 32; }

// the gradient kernel of a minibatch step: ppo_grad16w_kernel (512 threads, one warp group per tile slot); F16B200_GRAD16=-1 selects
// the 256-thread ppo_grad16_kernel (slots interleaved in program order) for A/B measurements
static bool grad16_interleaved()
{
    static const bool v = [] {
        const char *e = std::getenv("F16B200_GRAD16");
        return e && std::atoi(e) < 0;
    }();
    return v;
}
static cudaError_t configure_grad16()
{
    const int bytes = static_cast<int>(sizeof(TrainSmem16));
    cudaError_t e = cudaFuncSetAttribute(ppo_grad16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ppo_grad16w_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ppo_grad16w_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
}
static cudaError_t launch_grad16_variant(cudaLaunchConfig_t &lc, const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials)
{
    if (grad16_interleaved()) {
        lc.blockDim = dim3(kThreads);
        return cudaLaunchKernelEx(&lc, ppo_grad16_kernel, P, blob16, partials);
    }
    lc.blockDim = dim3(kThreadsW);
    ReduceAdamArgs none = {};
    return cudaLaunchKernelEx(&lc, ppo_grad16w_kernel<false>, P, blob16, partials, none, 1);
}

size_t ppo_epoch16_sync_bytes() { return 8 + 1024 * sizeof(float); }   // meeting counter + one squared-norm share per CTA

// One launch = the n_mb minibatch steps of an epoch (gradient + reduction + clip + Adam + blob scatter each): ppo_grad16w_kernel<true>.
// P.mb_idx: n_mb * P.n indices, P.adv_stats: n_mb pairs.  Cooperative launch (the CTAs meet at a grid-wide counter three times per step).
cudaError_t launch_ppo_epoch16(const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials, ReduceAdamArgs RA, int n_mb, int n_sm, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = configure_grad16();
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (P.n + kRows - 1) / kRows;
    const int grid = static_cast<int>(tiles < n_sm ? tiles : n_sm);
    if (grid > 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(RA.sync, 0, 8, st);   // the meeting counter counts from zero in every launch
    if (e != cudaSuccess) return e;
    RA.partials = partials;
    RA.n_rows = grid;
    RA.n_out = static_cast<int>(sizeof(PolicyGrad) / sizeof(float));
    RA.grad = reinterpret_cast<float *>(P.grad);
    RA.n_pad = ppo_grad_floats_padded();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kThreadsW);
    lc.dynamicSmemBytes = sizeof(TrainSmem16);
    lc.stream = st;
    lc.attrs = attr;
    lc.numAttrs = 1;
    e = cudaLaunchKernelEx(&lc, ppo_grad16w_kernel<true>, P, blob16, partials, RA, n_mb);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_ppo_grad16(const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials, int n_sm, cudaStream_t st, const PeerArgs *peer)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = configure_grad16();
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (P.n + kRows - 1) / kRows;
    const int grid = static_cast<int>(tiles < n_sm ? tiles : n_sm);   // one CTA per SM: it owns all 512 TMEM columns
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the minibatch chain gradient -> reduction -> optimiser -> gradient ...
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kThreads);
    lc.dynamicSmemBytes = sizeof(TrainSmem16);
    lc.stream = st;
    lc.attrs = attr;
    lc.numAttrs = 1;
    cudaError_t e = launch_grad16_variant(lc, P, blob16, partials);
    if (e != cudaSuccess) return e;
    if (partials) {
        constexpr int n_out = static_cast<int>(sizeof(PolicyGrad) / sizeof(float));
        lc.gridDim = dim3((n_out + 31) / 32);
        lc.blockDim = dim3(256);
        lc.dynamicSmemBytes = 0;
        if (peer)
            e = cudaLaunchKernelEx(&lc, ppo_grad_reduce_peer_kernel, static_cast<const float *>(partials), grid, n_out, reinterpret_cast<float *>(P.grad),
                                   peer->bufs, peer->seq, peer->rank, peer->world, ppo_grad_floats_padded());
        else
            e = cudaLaunchKernelEx(&lc, ppo_grad_reduce_kernel, static_cast<const float *>(partials), grid, n_out, reinterpret_cast<float *>(P.grad));
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

size_t ppo_step16_sync_bytes() { return 8 + static_cast<size_t>(ppo_grad_reduce_ctas()) * (sizeof(float) + sizeof(uint32_t)); }

cudaError_t launch_ppo_step16(const PpoGradArgs &P, const TrainBlob16 *blob16, float *partials, ReduceAdamArgs RA, int n_sm, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = configure_grad16();
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (P.n + kRows - 1) / kRows;
    const int grid = static_cast<int>(tiles < n_sm ? tiles : n_sm);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kThreads);
    lc.dynamicSmemBytes = sizeof(TrainSmem16);
    lc.stream = st;
    lc.attrs = attr;
    lc.numAttrs = 1;
    cudaError_t e = launch_grad16_variant(lc, P, blob16, partials);
    if (e != cudaSuccess) return e;
    RA.partials = partials;
    RA.n_rows = grid;
    RA.n_out = static_cast<int>(sizeof(PolicyGrad) / sizeof(float));
    RA.grad = reinterpret_cast<float *>(P.grad);
    RA.n_pad = ppo_grad_floats_padded();
    lc.gridDim = dim3(ppo_grad_reduce_ctas());
    lc.blockDim = dim3(256);
    lc.dynamicSmemBytes = 0;
    e = RA.bufs ? cudaLaunchKernelEx(&lc, ppo_reduce_adam_kernel<true>, RA) : cudaLaunchKernelEx(&lc, ppo_reduce_adam_kernel<false>, RA);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

cudaError_t launch_ppo_grad(const PpoGradArgs &P, int n_sm, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ppo_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(TrainSmem)));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t tiles = (P.n + kRows - 1) / kRows;
    const int grid = static_cast<int>(tiles < n_sm ? tiles : n_sm);   // one CTA per SM (206 KB of shared memory)
    ppo_grad_kernel<<<grid, kThreads, sizeof(TrainSmem), st>>>(P);
    return cudaGetLastError();
}

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
