// f16_rollout.cu -- the closed-loop rollout in ONE persistent kernel: policy MLPs (tcgen05) -> Gaussian head -> env step,
// K times per launch, state and observation never leaving the SM.
//
// Replaces the inner loop of the reference's trainer (control/ppo/ppo_model_gsde.py:216-230):
//     for step in range(num_steps):
//         action, logprob, _, value = agent.get_action_and_value(next_obs)       # two 5-64-64-1 tanh MLPs + Gaussian head
//         next_obs, reward, terminated, truncated, infos = envs.step(action)     # F16.step for every env (+ autoreset)
// which the two-kernel path runs as policy_forward_kernel -> env_step2_kernel per step (2 launches, the observation,
// action and state round-tripping through L2 / HBM every step).  Here a sub-tile of 128 aircraft stays resident in one
// group of four warps for all K steps: thread r of the group owns aircraft r -- its nine state values, counters and
// current observation live in registers -- and row r of the group's tensor-core tiles (TMEM lane r):
//
//   per step   X[128 x 16]  (obs, fp16)  --tcgen05.mma-->  D[128 x 128] (TMEM)   layer 1, actor | critic stacked
//              H1 = tanh(D + b1) -> shared memory (fp16, canonical K-major layout)
//              H1[:, net]  --4 x tcgen05.mma per net-->  D[128 x 128]              layer 2 (overwrites layer 1's columns)
//              mean / value = w3 . tanh(D + b2) + b3                                layer 3 in the epilogue (CUDA cores)
//              action = mean + std * eps (Philox, keyed by the env's own counters), log-probability
//              F16.step with that action: same arithmetic as env_step_kernel (bit-identical, tested), autoreset included
//   streams out only the rollout rows: obs[k+1], action, logprob, value, reward, done  (37 B per env-step)
//
// A CTA is four such groups (512 threads, all 512 TMEM columns, one copy of the weights and of the aero tables in shared
// memory); the groups synchronise among their own 128 threads only (named barriers, one mbarrier each for MMA completion),
// so they drift apart and one group's tanh stream (MUFU) overlaps another's MMA round trip or env step.  The kernel is
// bound by MUFU.TANH: 256 tanh + 12 other MUFU per env-step at 16 per clock per SM (profiles/r02_mufu_probe.md).
// Operands are fp16 (11 significant bits, like TF32) -- the layout and weights blob of the training kernel
// (TrainBlob16), so the rollout's forward pass and the update's forward pass round the same way.
#include <cuda_fp16.h>

#include "f16_kernels.cuh"
#include "f16_policy.h"
#include "f16_umma.cuh"

namespace f16 {
namespace {

constexpr int kSub = 4;                       // 128-env groups per CTA
constexpr int kRowsR = 128;                   // envs per group == UMMA M == TMEM lanes
constexpr int kThreadsR = kSub * kRowsR;
constexpr uint32_t kLbo = kRowsR * 16;        // bytes between the 16-byte K-chunks of the X / H1 operands

struct alignas(16) RolloutSmem {
    Tables<float> tbl;
    // the first 20 KB and the fp32 tail of TrainBlob16 (its W2^T block is not needed here)
    __half w1[2][128][8];
    __half w2[2][8][64][8];
    float b1[128], b2[128], w3[128], b3[4];
    __half a1[kSub][2][kRowsR][8];            // X : K = 16 (obs[0..4], 1, 0 ...)
    __half a2[kSub][16][kRowsR][8];           // H1: 16 K-chunks (actor 0-7 | critic 8-15) x 128 rows x 8 halfs
    uint64_t bar_tbl, bar_w;
    uint64_t bar_mma[kSub];
    uint32_t tmem_base;
    CtaStats stats;
};
static_assert(offsetof(RolloutSmem, w2) - offsetof(RolloutSmem, w1) == sizeof(TrainBlob16::w1), "w1 | w2 must be contiguous");
static_assert(offsetof(RolloutSmem, b1) - offsetof(RolloutSmem, w1) == offsetof(TrainBlob16, w2t), "fp32 tail follows w2");
static_assert(sizeof(RolloutSmem) <= 227 * 1024, "shared memory per CTA");

__device__ __forceinline__ uint4 pack8h(const float (&v)[8])
{
    __half2 h[4] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3]), __floats2half2_rn(v[4], v[5]), __floats2half2_rn(v[6], v[7])};
    return *reinterpret_cast<const uint4 *>(h);
}

__device__ __forceinline__ void group_sync(unsigned sub)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1u + sub), "n"(kRowsR) : "memory");
}

template <bool FAST>
__global__ void __launch_bounds__(kThreadsR, 1) rollout_kernel(const EnvArgs<float> A, const RolloutArgs R)
{
    namespace u = umma;
    using Rl = float;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RolloutSmem *sm = reinterpret_cast<RolloutSmem *>(smem_raw);
    const Tables<float> &tbl = sm->tbl;
    const unsigned t = threadIdx.x, warp = t >> 5, sub = t >> 7, row = t & (kRowsR - 1);

    if (t == 0) {
        mbar_init(&sm->bar_tbl, 1);
        mbar_init(&sm->bar_w, 1);
#pragma unroll
        for (int s = 0; s < kSub; ++s) mbar_init(&sm->bar_mma[s], 1);
        sm->stats.count = 0;
        sm->stats.length = 0;
        sm->stats.ret = 0.0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // the CTA owns the SM: all 512 TMEM columns, 128 per group
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    *reinterpret_cast<uint4 *>(sm->a1[sub][1][row]) = make_uint4(0u, 0u, 0u, 0u);   // K padding 8..15 of X, written once
    u::fence_before_sync();
    __syncthreads();
    u::fence_after_sync();
    const uint32_t tmem = sm->tmem_base;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (t == 0) tma_bulk_g2s(&sm->tbl, A.tables, static_cast<uint32_t>(sizeof(Tables<float>)), &sm->bar_tbl);
    // the weights are the optimiser kernel's output, the state and obs0 the reset's / the previous rollout's
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (t == 0) {
        constexpr uint32_t head = static_cast<uint32_t>(offsetof(TrainBlob16, w2t));
        constexpr uint32_t tail = static_cast<uint32_t>(sizeof(TrainBlob16) - offsetof(TrainBlob16, b1));
        const unsigned char *src = reinterpret_cast<const unsigned char *>(R.blob);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->bar_w)), "r"(head + tail) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm->w1)),
                     "l"(src), "r"(head), "r"(smem_u32(&sm->bar_w))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm->b1)),
                     "l"(src + offsetof(TrainBlob16, b1)), "r"(tail), "r"(smem_u32(&sm->bar_w))
                     : "memory");
    }
    mbar_wait(&sm->bar_tbl, 0);
    mbar_wait(&sm->bar_w, 0);

    constexpr uint32_t id_l1 = u::idesc_f16(128, 128), id_l2 = u::idesc_f16(128, 64);
    const uint32_t a1 = smem_u32(sm->a1[sub]), a2 = smem_u32(sm->a2[sub]);
    const uint32_t w1 = smem_u32(sm->w1), w2a = smem_u32(sm->w2[0]), w2c = smem_u32(sm->w2[1]);
    const uint32_t tcol = tmem + sub * 128u;                                  // this group's accumulator columns
    const uint32_t lane_base = tcol + (((warp & 3u) * 32u) << 16);           // this warp's 32 TMEM lanes
    uint64_t *bar = &sm->bar_mma[sub];
    uint32_t phase = 0;

    const unsigned n = static_cast<unsigned>(A.n);
    const unsigned n_tiles = (n + kRowsR - 1) / kRowsR;
    const float dt = A.dt;
    const float log_std = R.logstd[0], std_dev = __expf(log_std);
    const uint2 pkey = make_uint2(static_cast<uint32_t>(R.policy_seed), static_cast<uint32_t>(R.policy_seed >> 32));
    const int K = R.K;

    for (unsigned tile = blockIdx.x * kSub + sub; tile < n_tiles; tile += gridDim.x * kSub) {
        const unsigned i = tile * kRowsR + row;
        const bool live = i < n;
        // ---- this aircraft: state, counters, current observation -> registers for the whole launch ----
        Rl x[9] = {0.f, 3000.f, 0.f, 0.05f, 200.f, 0.05f, 0.f, 0.f, 0.f};   // a harmless flight condition for the rows past n
        Rl o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        int ep_len = 0, ref_idx = 0;
        uint32_t episode = 0;
        Rl ret = 0.f;
        if (live) {
            const Rl *p = A.state + i;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                x[j] = *p;
                p += n;
            }
            ep_len = A.ep_len[i];
            ref_idx = A.ref_idx[i];
            ret = A.total_return[i];
            episode = A.episode_no[i];
            const Rl *po = R.obs0 + static_cast<size_t>(i) * 5u;
#pragma unroll
            for (int j = 0; j < 5; ++j) o[j] = po[j];
        } else if (A.ref_codes) {
            ref_idx = static_cast<int>(A.ref_codes[0]);
        }
        const Rl pa_in = x[8];
        bool was_reset = false;

        for (int k = 0; k <= K; ++k) {
            // ================= policy forward on the current observation =================
            {
                const float xin[8] = {o[0], o[1], o[2], o[3], o[4], 1.f, 0.f, 0.f};
                *reinterpret_cast<uint4 *>(sm->a1[sub][0][row]) = pack8h(xin);
            }
            u::fence_async_smem();
            u::fence_before_sync();        // this thread's TMEM loads of the previous step are complete (wait_ld) and ordered
            group_sync(sub);
            if (row == 0) {
                u::fence_after_sync();
                u::mma_f16(tcol, u::desc(a1, kLbo, 128), u::desc(w1, 128 * 16, 128), id_l1, 0u);
                u::commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            u::fence_after_sync();
            // ---- H1 = tanh(D + b1) -> a2 (fp16) ----
#pragma unroll
            for (int pair = 0; pair < 2; ++pair) {
                uint32_t ra[32], rb[32];
                u::tmem_ld32_issue(lane_base + pair * 64, ra);
                u::tmem_ld32_issue(lane_base + pair * 64 + 32, rb);
                u::wait_ld();
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
                    const int col0 = pair * 64 + blk * 32;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = col0 + c * 8;
                        float h[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) h[q] = u::tanh_approx(__uint_as_float(blk ? rb[c * 8 + q] : ra[c * 8 + q]) + sm->b1[j + q]);
                        *reinterpret_cast<uint4 *>(sm->a2[sub][j >> 3][row]) = pack8h(h);
                    }
                }
            }
            u::fence_async_smem();
            u::fence_before_sync();
            group_sync(sub);
            if (row == 0) {
                u::fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)   // actor: H1 chunks 0-7, two per MMA (K = 16)
                    u::mma_f16(tcol, u::desc(a2 + kk * 2 * kLbo, kLbo, 128), u::desc(w2a + kk * 2 * 64 * 16, 64 * 16, 128), id_l2, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)   // critic: chunks 8-15
                    u::mma_f16(tcol + 64, u::desc(a2 + (8 + kk * 2) * kLbo, kLbo, 128), u::desc(w2c + kk * 2 * 64 * 16, 64 * 16, 128), id_l2,
                            kk > 0 ? 1u : 0u);
                u::commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
            u::fence_after_sync();
            // ---- heads: mean = w3a . tanh(D[:, :64] + b2a) + b3a, value likewise from the critic columns ----
            float head[2] = {sm->b3[0], sm->b3[1]};
#pragma unroll
            for (int pair = 0; pair < 2; ++pair) {
                uint32_t ra[32], rb[32];
                u::tmem_ld32_issue(lane_base + pair * 64, ra);
                u::tmem_ld32_issue(lane_base + pair * 64 + 32, rb);
                u::wait_ld();
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    acc0 = fmaf(u::tanh_approx(__uint_as_float(ra[c]) + sm->b2[pair * 64 + c]), sm->w3[pair * 64 + c], acc0);
                    acc1 = fmaf(u::tanh_approx(__uint_as_float(rb[c]) + sm->b2[pair * 64 + 32 + c]), sm->w3[pair * 64 + 32 + c], acc1);
                }
                head[pair] += acc0 + acc1;
            }
            const float mean = head[0], value = head[1];
            if (k == K) {   // one forward past the last step: the GAE bootstrap value of the final observation
                if (live && R.next_value) R.next_value[i] = value;
                break;
            }
            // ================= Gaussian policy head (same counter stream as policy_forward_kernel) =================
            float eps = 0.f;
            if (R.sample) {
                const int64_t gid = A.id_base + i;
                const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32) ^ 0x50504F31u, episode,
                                                      static_cast<uint32_t>(ep_len)), pkey);
                const float u1 = (static_cast<float>(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
                const float u2 = static_cast<float>(r.y >> 8) * (1.0f / 16777216.0f);
                eps = sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
            }
            const float action = mean + std_dev * eps;
            const unsigned kn = static_cast<unsigned>(k) * n + i;
            if (live) {
                R.actions[kn] = action;
                R.logprobs[kn] = -0.5f * eps * eps - log_std - 0.91893853320467274178f;
                R.values[kn] = value;
            }

            // ================= F16.step (env/env.py:46-70), the arithmetic of env_step_kernel<float, FAST, autoreset> =================
            const Rl ref = ref_sample<Rl>(A, ref_idx, ep_len);
            const Rl a = clampr(action, Rl(-1), Rl(1));
            const Rl u_stab = Rl(K::act_lo) + Rl(0.5) * (a + Rl(1.0)) * Rl(K::act_span);
            Rl lo0 = Rl(0), lo1 = Rl(0);
            euler_step<Rl, FAST, true>(tbl, x, u_stab, Rl(0), dt, false, lo0, lo1);
            Rl pen = Rl(0);
            uint32_t done_now = 0;
            if (!(x[1] > Rl(300))) { pen += Rl(-1000); done_now = 1; }
            if (x[1] >= Rl(30000)) { pen += Rl(-1000); done_now = 1; }
            if (!(Mth<Rl, FAST>::fabs(x[2]) < Rl(K::wz_term))) { pen += Rl(-1000); done_now = 1; }
            if (ep_len == A.horizon) done_now = 1;
            const Rl reward = pen + tracking_reward<Rl, FAST>(ref, x[2]);
            make_obs<Rl, FAST>(x[1], x[2], x[4], ref, o);
            ep_len += 1;
            ret += reward;
            if (done_now && live) {
                if (R.stats) episode_stats_add(&sm->stats, static_cast<double>(ret), ep_len);
                if (R.final_obs) store_obs<Rl>(R.final_obs, 0, n, i, o);
                if (R.final_ep_len) R.final_ep_len[i] = ep_len;
                if (R.final_return) R.final_return[i] = ret;
                episode += 1u;
                init_episode<Rl>(A, i, episode, false, x, ref_idx);
                was_reset = true;
                ep_len = 0;
                ret = Rl(0);
                make_obs<Rl, FAST>(x[1], x[2], x[4], ref_sample<Rl>(A, ref_idx, 0), o);
            }
            if (live) {
                R.rewards[kn] = reward;
                R.dones[kn] = static_cast<uint8_t>(done_now);
                Rl *po = R.obs_next + (static_cast<size_t>(k) * n + i) * 5u;
#pragma unroll
                for (int j = 0; j < 5; ++j) po[j] = o[j];
            }
        }

        // ---- write back ----
        if (live) {
            Rl *p = A.state + i;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                if (j == 4) {
                    if (was_reset) *p = x[4];
                } else if (j == 8) {
                    if (x[8] != pa_in) *p = x[8];
                } else {
                    *p = x[j];
                }
                p += n;
            }
            A.ep_len[i] = ep_len;
            A.total_return[i] = ret;
            if (was_reset) {
                A.ref_idx[i] = ref_idx;
                A.episode_no[i] = episode;
            }
        }
    }

    u::fence_before_sync();
    __syncthreads();
    if (R.stats && t == 0 && sm->stats.count) {
        atomicAdd(R.stats + 0, static_cast<double>(sm->stats.count));
        atomicAdd(R.stats + 1, sm->stats.ret);
        atomicAdd(R.stats + 2, static_cast<double>(sm->stats.length));
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

}  // namespace

cudaError_t launch_rollout(const EnvArgs<float> &A, const RolloutArgs &R, bool fast, int n_sm, int pdl, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(RolloutSmem)));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(RolloutSmem)));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t groups = (A.n + kRowsR - 1) / kRowsR;
    const int64_t ctas = (groups + kSub - 1) / kSub;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(static_cast<unsigned>(ctas < n_sm ? ctas : n_sm));   // one CTA per SM: it owns all 512 TMEM columns
    lc.blockDim = dim3(kThreadsR);
    lc.dynamicSmemBytes = sizeof(RolloutSmem);
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    cudaError_t e = fast ? cudaLaunchKernelEx(&lc, rollout_kernel<true>, A, R) : cudaLaunchKernelEx(&lc, rollout_kernel<false>, A, R);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace f16
