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
//              H1 = tanh(D) -> shared memory (fp16, canonical K-major layout); the biases ride in the MMAs (see below)
//              H1[:, net]  --4 x tcgen05.mma per net-->  D[128 x 128]              layer 2 (overwrites layer 1's columns)
//              mean / value = w3 . tanh(D) + b3                                     layer 3 in the epilogue (CUDA cores)
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

#include <cstdlib>

#include "f16_kernels.cuh"
#include "f16_policy.h"
#include "f16_umma.cuh"

namespace f16 {
#ifdef F16_ROLLOUT_TRACE
// debug build only (make EXTRA=-DF16_ROLLOUT_TRACE=8): clock64 at the phase boundaries of CTA 0's 16 warps, steps 8..15
__device__ unsigned long long g_rollout_trace[16][8][12];
__device__ unsigned long long g_rollout_starts[16][128];   // start of every step k < 128
#define RTRACE(slot)                                                                                  \
    do {                                                                                              \
        if (blockIdx.x == 0 && lane == 0 && k >= F16_ROLLOUT_TRACE && k < F16_ROLLOUT_TRACE + 8)              \
            g_rollout_trace[warp][k - F16_ROLLOUT_TRACE][slot] = clock64();                                   \
    } while (0)
#else
#define RTRACE(slot) do { } while (0)
#endif
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
    __half wb2[2][2][64][8];                  // layer-2 bias as one more K = 16 slice of W2: k = 0 fp16(b2), k = 1 the fp16 remainder
    __half ones[2][kRowsR][8];                // its A operand: k = 0, 1 are 1.0 for every row (shared by the four groups)
    __half a1[kSub][2][kRowsR][8];            // X : K = 16 (obs[0..4], 1, 1, 0 ...)
    __half a2[kSub][16][kRowsR][8];           // H1: 16 K-chunks (actor 0-7 | critic 8-15) x 128 rows x 8 halfs
    uint64_t bar_tbl, bar_w;
    uint64_t bar_mma[kSub][3];                // per group: layer 1, layer 2 actor, layer 2 critic (one phase flip per step each)
    uint64_t bar_adm[kSub];                   // per group: admission to the tanh section (pacing, see below)
    uint32_t adm_tickets, adm_warps_done;     // FIFO tickets taken by groups / warps that have left their tanh section
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
    const unsigned t = threadIdx.x, row = t & (kRowsR - 1), lane = t & 31u;
    const unsigned warp = __shfl_sync(0xffffffffu, t >> 5, 0), sub = warp >> 2;   // warp-uniform for the compiler, too

    if (t == 0) {
        mbar_init(&sm->bar_tbl, 1);
        mbar_init(&sm->bar_w, 1);
#pragma unroll
        for (int s = 0; s < kSub; ++s) {
            for (int j = 0; j < 3; ++j) mbar_init(&sm->bar_mma[s][j], 1);
            mbar_init(&sm->bar_adm[s], 1);
        }
        sm->adm_tickets = 0;
        sm->adm_warps_done = 0;
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
    // The biases ride in the MMAs (no shared-memory loads or adds in the tanh epilogues): b = hi + lo with hi = fp16(b),
    // lo = fp16(b - hi) (22 significant bits), multiplied by two columns of ones and accumulated in fp32 by the tensor core.
    // Layer 1: X carries the ones in k = 5, 6 and this CTA's copy of W1 gets (hi, lo) there (zero in the blob);
    // layer 2: one extra K = 16 MMA per net with A = `ones`, B = `wb2`.
    if (t < 128) {
        const float b = sm->b1[t];
        const __half hi = __float2half_rn(b);
        sm->w1[0][t][5] = hi;
        sm->w1[0][t][6] = __float2half_rn(b - __half2float(hi));
    } else if (t < 256) {
        const unsigned j = t - 128u;
        const float b = sm->b2[j];
        const __half hi = __float2half_rn(b), z = __float2half_rn(0.f);
        __half *d = sm->wb2[j >> 6][0][j & 63u];
        d[0] = hi;
        d[1] = __float2half_rn(b - __half2float(hi));
#pragma unroll
        for (int q = 2; q < 8; ++q) d[q] = z;
        *reinterpret_cast<uint4 *>(sm->wb2[j >> 6][1][j & 63u]) = make_uint4(0u, 0u, 0u, 0u);
    } else if (t < 384) {
        const unsigned r = t - 256u;
        *reinterpret_cast<uint4 *>(sm->ones[0][r]) = make_uint4(0x3C003C00u, 0u, 0u, 0u);   // {1.0h, 1.0h, 0 ...}
        *reinterpret_cast<uint4 *>(sm->ones[1][r]) = make_uint4(0u, 0u, 0u, 0u);
    }
    u::fence_async_smem();
    __syncthreads();

    constexpr uint32_t id_l1 = u::idesc_f16(128, 128), id_l2 = u::idesc_f16(128, 64);
    const uint32_t a1 = smem_u32(sm->a1[sub]), a2 = smem_u32(sm->a2[sub]);
    const uint32_t w1 = smem_u32(sm->w1), w2a = smem_u32(sm->w2[0]), w2c = smem_u32(sm->w2[1]);
    const uint32_t ones = smem_u32(sm->ones), wb2 = smem_u32(sm->wb2);
    const uint32_t tcol = tmem + sub * 128u;                                  // this group's accumulator columns
    const uint32_t lane_base = tcol + (((warp & 3u) * 32u) << 16);           // this warp's 32 TMEM lanes
    uint64_t *bar1 = &sm->bar_mma[sub][0], *bar2a = &sm->bar_mma[sub][1], *bar2c = &sm->bar_mma[sub][2];
    uint32_t phase = 0;
    // One thread per group issues its MMAs; group g's issuer sits in warp g of the group, i.e. on SM sub-partition g, so
    // that the four issuers (which block while the tensor-core queue drains the other groups' MMAs) are spread over the
    // four schedulers instead of all living on sub-partition 0 (r02c trace: that sub-partition's warps ran every phase
    // 15-40 % longer and the other three waited for them at the group barriers).
    // The issuing WARP takes the branch as a whole and elects one lane inside it: with warp-uniform descriptors the MMA issue
    // is a handful of uniform-datapath instructions (a per-thread `if (row == ...)` made ptxas wrap every tcgen05.mma in a
    // lane-waterfall loop: ~90 clocks per MMA, r02e trace).
    const bool issuer_warp = (warp & 3u) == sub;

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
            // Split-phase schedule: the two nets' layer-2 products are issued as soon as "their" half of H1 is in shared
            // memory, so the tensor-core round trip of the actor half runs under the critic half's tanh work and the
            // critic's under the actor head (r02b trace: sync + wait of a single layer-2 commit cost 1.3-1.5 k of the 12 k
            // clocks of a step).  Accumulator columns [0, 64) are the actor's in both layers, [64, 128) the critic's.
            RTRACE(0);
#ifdef F16_ROLLOUT_TRACE
            if (blockIdx.x == 0 && lane == 0 && k < 128) g_rollout_starts[warp][k] = clock64();
#endif
            {
                const float xin[8] = {o[0], o[1], o[2], o[3], o[4], 1.f, 1.f, 0.f};
                *reinterpret_cast<uint4 *>(sm->a1[sub][0][row]) = pack8h(xin);
            }
            u::fence_async_smem();
            u::fence_before_sync();        // this thread's TMEM loads of the previous step are complete (wait_ld) and ordered
            group_sync(sub);
            if (issuer_warp && u::elect_one()) {
                u::fence_after_sync();
                u::mma_f16(tcol, u::desc(a1, kLbo, 128), u::desc(w1, 128 * 16, 128), id_l1, 0u);
                u::commit(bar1);
                if (R.pace) {
                    // Pacing: at most R.pace of the CTA's four groups are inside their tanh section (layer-1 epilogue .. heads) at a time,
                    // admitted first come, first served.  Left alone, the four warps of a sub-partition -- one of each group -- fall into
                    // lockstep: all four share the MUFU unit through the tanh phases, then all four sit in the env step with the unit idle
                    // (period = MUFU work + the rest).  With two groups inside, two warps per sub-partition keep the unit saturated while the
                    // other two run their env step / MMA round trip (period -> max of the two).  A group is admitted as a whole (its four warps
                    // wait on bar_adm), so the group barriers inside the section cannot wait for a warp that is still queued.
                    const uint32_t my = atomicAdd(&sm->adm_tickets, 1u);
                    const uint32_t need = my >= static_cast<uint32_t>(R.pace) ? 4u * (my - static_cast<uint32_t>(R.pace) + 1u) : 0u;   // warps of the groups ahead that must be out
                    for (unsigned spin = 0;; ++spin) {
                        const uint32_t done_w = *reinterpret_cast<volatile uint32_t *>(&sm->adm_warps_done);
                        if (static_cast<int32_t>(done_w - need) >= 0) break;
                        if (spin > (1u << 26)) __trap();
                        __nanosleep(32);
                    }
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm->bar_adm[sub])) : "memory");
                }
            }
            RTRACE(1);
            mbar_wait(bar1, phase);
            if (R.pace) mbar_wait(&sm->bar_adm[sub], phase);
            u::fence_after_sync();
            RTRACE(2);
#pragma unroll
            for (int net = 0; net < 2; ++net) {
                // ---- H1[:, net] = tanh(D[:, 64 net .. 64 net + 63]) -> a2 chunks 8 net .. 8 net + 7 (fp16) ----
                uint32_t ra[32], rb[32];
                u::tmem_ld32_issue(lane_base + net * 64, ra);
                u::tmem_ld32_issue(lane_base + net * 64 + 32, rb);
                u::wait_ld();
#pragma unroll
                for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = net * 64 + blk * 32 + c * 8;
                        float h[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) h[q] = u::tanh_approx(__uint_as_float(blk ? rb[c * 8 + q] : ra[c * 8 + q]));
                        *reinterpret_cast<uint4 *>(sm->a2[sub][j >> 3][row]) = pack8h(h);
                    }
                }
                if (net == 0) RTRACE(3);
                u::fence_async_smem();
                u::fence_before_sync();
                group_sync(sub);
                if (issuer_warp && u::elect_one()) {   // layer 2 of this net: 4 x (K = 16) over its H1 chunks + the bias slice; overwrites its layer-1 columns
                    u::fence_after_sync();
                    const uint32_t w2n = net ? w2c : w2a;
                    if (net == 0) RTRACE(8);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        u::mma_f16(tcol + net * 64, u::desc(a2 + (net * 8 + kk * 2) * kLbo, kLbo, 128), u::desc(w2n + kk * 2 * 64 * 16, 64 * 16, 128),
                                   id_l2, kk > 0 ? 1u : 0u);
                    if (net == 0) RTRACE(9);
                    u::mma_f16(tcol + net * 64, u::desc(ones, kLbo, 128), u::desc(wb2 + net * 2 * 64 * 16, 64 * 16, 128), id_l2, 1u);
                    if (net == 0) RTRACE(10);
                    u::commit(net ? bar2c : bar2a);
                    if (net == 0) RTRACE(11);
                }
            }
            RTRACE(4);
            // ---- heads: mean = w3a . tanh(D[:, :64]) + b3a, value likewise from the critic columns ----
            float head[2] = {sm->b3[0], sm->b3[1]};
#pragma unroll
            for (int net = 0; net < 2; ++net) {
                mbar_wait(net ? bar2c : bar2a, phase);
                u::fence_after_sync();
                if (net == 0) RTRACE(5);
                uint32_t ra[32], rb[32];
                u::tmem_ld32_issue(lane_base + net * 64, ra);
                u::tmem_ld32_issue(lane_base + net * 64 + 32, rb);
                u::wait_ld();
                float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    acc0 = fmaf(u::tanh_approx(__uint_as_float(ra[c])), sm->w3[net * 64 + c], acc0);
                    acc1 = fmaf(u::tanh_approx(__uint_as_float(rb[c])), sm->w3[net * 64 + 32 + c], acc1);
                }
                head[net] += acc0 + acc1;
            }
            phase ^= 1u;
            if (R.pace) {   // this warp has left the tanh section
                __syncwarp();
                if (lane == 0) atomicAdd(&sm->adm_warps_done, 1u);
            }
            RTRACE(6);
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
            RTRACE(7);
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


// ---------------------------------------------------------------------------
// policy_forward16_kernel -- Agent.get_action_and_value (control/ppo/ppo_model_gsde.py:81-99,158-160) for a batch of
// observations with the rollout kernel's schedule: fp16 tensor-core operands (the TrainBlob16 the optimiser keeps current),
// biases inside the MMAs, four independent 128-row groups per CTA (16 warps, all 512 TMEM columns), split-phase layer 2.
// The TF32 policy_forward_kernel (f16_policy.cu) runs one 128-row tile per 256-thread CTA with a single MMA round trip in
// flight and reaches 0.54 of the MUFU.TANH rate at 2^20 rows; here a group's MMA round trips and shared-memory stores hide
// under the other three groups' tanh streams.  Same counter-RNG keys as policy_forward_kernel: the sampled noise is identical.
// ---------------------------------------------------------------------------
struct alignas(16) Policy16Smem {
    __half w1[2][128][8];
    __half w2[2][8][64][8];
    float b1[128], b2[128], w3[128], b3[4];
    __half wb2[2][2][64][8];
    __half ones[2][kRowsR][8];
    __half a1[kSub][2][kRowsR][8];
    __half a2[kSub][16][kRowsR][8];
    uint64_t bar_w;
    uint64_t bar_mma[kSub][3];
    uint64_t bar_adm[kSub];
    uint32_t adm_tickets, adm_warps_done;
    uint32_t tmem_base;
};
static_assert(offsetof(Policy16Smem, w2) - offsetof(Policy16Smem, w1) == sizeof(TrainBlob16::w1), "w1 | w2 must be contiguous");
static_assert(offsetof(Policy16Smem, b1) - offsetof(Policy16Smem, w1) == offsetof(TrainBlob16, w2t), "fp32 tail follows w2");

template <int PACE>   // groups admitted to their tanh section at a time (0 = no pacing), see rollout_kernel
__global__ void __launch_bounds__(kThreadsR, 1) policy_forward16_kernel(const PolicyArgs P, const TrainBlob16 *blob16)
{
    constexpr int pace = PACE;
    namespace u = umma;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Policy16Smem *sm = reinterpret_cast<Policy16Smem *>(smem_raw);
    const unsigned t = threadIdx.x, row = t & (kRowsR - 1);
    const unsigned warp = __shfl_sync(0xffffffffu, t >> 5, 0), sub = warp >> 2;

    if (t == 0) {
        mbar_init(&sm->bar_w, 1);
#pragma unroll
        for (int s = 0; s < kSub; ++s) {
            for (int j = 0; j < 3; ++j) mbar_init(&sm->bar_mma[s][j], 1);
            mbar_init(&sm->bar_adm[s], 1);
        }
        sm->adm_tickets = 0;
        sm->adm_warps_done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm->tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    *reinterpret_cast<uint4 *>(sm->a1[sub][1][row]) = make_uint4(0u, 0u, 0u, 0u);   // K padding 8..15 of X, written once
    u::fence_before_sync();
    __syncthreads();
    u::fence_after_sync();
    const uint32_t tmem = sm->tmem_base;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");   // weights: the optimiser's output; observations: the env step's
    if (t == 0) {
        constexpr uint32_t head = static_cast<uint32_t>(offsetof(TrainBlob16, w2t));
        constexpr uint32_t tail = static_cast<uint32_t>(sizeof(TrainBlob16) - offsetof(TrainBlob16, b1));
        const unsigned char *src = reinterpret_cast<const unsigned char *>(blob16);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm->bar_w)), "r"(head + tail) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm->w1)),
                     "l"(src), "r"(head), "r"(smem_u32(&sm->bar_w))
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm->b1)),
                     "l"(src + offsetof(TrainBlob16, b1)), "r"(tail), "r"(smem_u32(&sm->bar_w))
                     : "memory");
    }
    const unsigned n = static_cast<unsigned>(P.n);
    const unsigned n_tiles = (n + kRowsR - 1) / kRowsR;
    // first tile's observation row: requested before the weights land
    unsigned tile = blockIdx.x * kSub + sub;
    float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    auto load_obs = [&](unsigned tile_, float (&dst)[5]) {
        const unsigned i_ = tile_ * kRowsR + row;
        if (tile_ < n_tiles && i_ < n) {
            const float *po = P.obs + static_cast<size_t>(i_) * 5u;
#pragma unroll
            for (int j = 0; j < 5; ++j) dst[j] = __ldg(po + j);
        }
    };
    load_obs(tile, o);
    mbar_wait(&sm->bar_w, 0);
    // biases into the MMAs: see rollout_kernel
    if (t < 128) {
        const float b = sm->b1[t];
        const __half hi = __float2half_rn(b);
        sm->w1[0][t][5] = hi;
        sm->w1[0][t][6] = __float2half_rn(b - __half2float(hi));
    } else if (t < 256) {
        const unsigned j = t - 128u;
        const float b = sm->b2[j];
        const __half hi = __float2half_rn(b), z = __float2half_rn(0.f);
        __half *d = sm->wb2[j >> 6][0][j & 63u];
        d[0] = hi;
        d[1] = __float2half_rn(b - __half2float(hi));
#pragma unroll
        for (int q = 2; q < 8; ++q) d[q] = z;
        *reinterpret_cast<uint4 *>(sm->wb2[j >> 6][1][j & 63u]) = make_uint4(0u, 0u, 0u, 0u);
    } else if (t < 384) {
        const unsigned r = t - 256u;
        *reinterpret_cast<uint4 *>(sm->ones[0][r]) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(sm->ones[1][r]) = make_uint4(0u, 0u, 0u, 0u);
    }
    u::fence_async_smem();
    __syncthreads();

    constexpr uint32_t id_l1 = u::idesc_f16(128, 128), id_l2 = u::idesc_f16(128, 64);
    const uint32_t a1 = smem_u32(sm->a1[sub]), a2 = smem_u32(sm->a2[sub]);
    const uint32_t w1 = smem_u32(sm->w1), w2a = smem_u32(sm->w2[0]), w2c = smem_u32(sm->w2[1]);
    const uint32_t ones = smem_u32(sm->ones), wb2 = smem_u32(sm->wb2);
    const uint32_t tcol = tmem + sub * 128u;
    const uint32_t lane_base = tcol + (((warp & 3u) * 32u) << 16);
    uint64_t *bar1 = &sm->bar_mma[sub][0], *bar2a = &sm->bar_mma[sub][1], *bar2c = &sm->bar_mma[sub][2];
    uint32_t phase = 0;
    const bool issuer_warp = (warp & 3u) == sub;
    const float log_std = P.logstd ? P.logstd[0] : 0.f, std_dev = __expf(log_std);
    const uint2 pkey = make_uint2(static_cast<uint32_t>(P.seed), static_cast<uint32_t>(P.seed >> 32));

    for (; tile < n_tiles; tile += gridDim.x * kSub) {
        const unsigned i = tile * kRowsR + row;
        const bool live = i < n;
        {
            const float xin[8] = {o[0], o[1], o[2], o[3], o[4], 1.f, 1.f, 0.f};
            *reinterpret_cast<uint4 *>(sm->a1[sub][0][row]) = pack8h(xin);
        }
        u::fence_async_smem();
        u::fence_before_sync();
        group_sync(sub);
        if (issuer_warp && u::elect_one()) {
            u::fence_after_sync();
            u::mma_f16(tcol, u::desc(a1, kLbo, 128), u::desc(w1, 128 * 16, 128), id_l1, 0u);
            u::commit(bar1);
            if (pace) {   // pacing of the tanh section, as in rollout_kernel
                const uint32_t my = atomicAdd(&sm->adm_tickets, 1u);
                const uint32_t need = my >= static_cast<uint32_t>(pace) ? 4u * (my - static_cast<uint32_t>(pace) + 1u) : 0u;
                for (unsigned spin = 0;; ++spin) {
                    const uint32_t done_w = *reinterpret_cast<volatile uint32_t *>(&sm->adm_warps_done);
                    if (static_cast<int32_t>(done_w - need) >= 0) break;
                    if (spin > (1u << 26)) __trap();
                    __nanosleep(32);
                }
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm->bar_adm[sub])) : "memory");
            }
        }
        // the next tile's observation row travels under this tile's work
#pragma unroll
        for (int j = 0; j < 5; ++j) o[j] = 0.f;
        load_obs(tile + gridDim.x * kSub, o);
        // RNG counters of this row (read early: the loads hide under the MMA round trips)
        uint32_t c2 = 0u, c3 = P.step;
        if (live && P.action && P.sample) {
            if (P.episode_no) c2 = P.episode_no[i];
            if (P.ep_len) c3 = static_cast<uint32_t>(P.ep_len[i]);
        }
        mbar_wait(bar1, phase);
        if (pace) mbar_wait(&sm->bar_adm[sub], phase);
        u::fence_after_sync();
#pragma unroll
        for (int net = 0; net < 2; ++net) {
            uint32_t ra[32], rb[32];
            u::tmem_ld32_issue(lane_base + net * 64, ra);
            u::tmem_ld32_issue(lane_base + net * 64 + 32, rb);
            u::wait_ld();
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int j = net * 64 + blk * 32 + c * 8;
                    float h[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) h[q] = u::tanh_approx(__uint_as_float(blk ? rb[c * 8 + q] : ra[c * 8 + q]));
                    *reinterpret_cast<uint4 *>(sm->a2[sub][j >> 3][row]) = pack8h(h);
                }
            }
            u::fence_async_smem();
            u::fence_before_sync();
            group_sync(sub);
            if (issuer_warp && u::elect_one()) {
                u::fence_after_sync();
                const uint32_t w2n = net ? w2c : w2a;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    u::mma_f16(tcol + net * 64, u::desc(a2 + (net * 8 + kk * 2) * kLbo, kLbo, 128), u::desc(w2n + kk * 2 * 64 * 16, 64 * 16, 128),
                               id_l2, kk > 0 ? 1u : 0u);
                u::mma_f16(tcol + net * 64, u::desc(ones, kLbo, 128), u::desc(wb2 + net * 2 * 64 * 16, 64 * 16, 128), id_l2, 1u);
                u::commit(net ? bar2c : bar2a);
            }
        }
        float head[2] = {sm->b3[0], sm->b3[1]};
#pragma unroll
        for (int net = 0; net < 2; ++net) {
            mbar_wait(net ? bar2c : bar2a, phase);
            u::fence_after_sync();
            uint32_t ra[32], rb[32];
            u::tmem_ld32_issue(lane_base + net * 64, ra);
            u::tmem_ld32_issue(lane_base + net * 64 + 32, rb);
            u::wait_ld();
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                acc0 = fmaf(u::tanh_approx(__uint_as_float(ra[c])), sm->w3[net * 64 + c], acc0);
                acc1 = fmaf(u::tanh_approx(__uint_as_float(rb[c])), sm->w3[net * 64 + 32 + c], acc1);
            }
            head[net] += acc0 + acc1;
        }
        phase ^= 1u;
        if (pace) {
            __syncwarp();
            if ((t & 31u) == 0) atomicAdd(&sm->adm_warps_done, 1u);
        }
        if (live) {
            const float mean = head[0];
            if (P.value) P.value[i] = head[1];
            if (P.mean) P.mean[i] = mean;
            if (P.action) {
                float eps = 0.f;
                if (P.sample) {
                    const int64_t gid = P.id_base + i;
                    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32) ^ 0x50504F31u, c2, c3), pkey);
                    const float u1 = (static_cast<float>(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
                    const float u2 = static_cast<float>(r.y >> 8) * (1.0f / 16777216.0f);
                    eps = sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
                }
                P.action[i] = mean + std_dev * eps;
                if (P.logprob) P.logprob[i] = -0.5f * eps * eps - log_std - 0.91893853320467274178f;
            }
        }
    }

    u::fence_before_sync();
    __syncthreads();
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
    RolloutArgs Rp = R;
    static const int pace_env = [] {
        const char *e = std::getenv("F16B200_ROLLOUT_PACE");   // tuning knob: groups admitted to the tanh section at a time, 0 = no pacing
        return e ? std::atoi(e) : -1;
    }();
    if (pace_env >= 0) Rp.pace = pace_env > 4 ? 4 : pace_env;
    cudaError_t e = fast ? cudaLaunchKernelEx(&lc, rollout_kernel<true>, A, Rp) : cudaLaunchKernelEx(&lc, rollout_kernel<false>, A, Rp);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}


cudaError_t launch_policy_forward16(const PolicyArgs &P, const TrainBlob16 *blob16, int n_sm, int pdl, cudaStream_t st)
{
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(policy_forward16_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Policy16Smem)));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(policy_forward16_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Policy16Smem)));
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    const int64_t groups = (P.n + kRowsR - 1) / kRowsR;
    const int64_t ctas = (groups + kSub - 1) / kSub;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(static_cast<unsigned>(ctas < n_sm ? ctas : n_sm));
    lc.blockDim = dim3(kThreadsR);
    lc.dynamicSmemBytes = sizeof(Policy16Smem);
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = pdl ? 1 : 0;
    // pacing pays once every group has several tiles (2^20 rows: 84.3 -> 77.2 us with three of the four groups admitted at a time); with one
    // or two tiles per group it only delays (65,536 rows: 8.3 -> 9.2 us)
    const bool paced = groups >= static_cast<int64_t>(n_sm) * kSub * 4;
    cudaError_t e = paced ? cudaLaunchKernelEx(&lc, policy_forward16_kernel<3>, P, blob16) : cudaLaunchKernelEx(&lc, policy_forward16_kernel<0>, P, blob16);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

#ifdef F16_ROLLOUT_TRACE
extern "C" int f16_debug_rollout_trace(unsigned long long *out)
{
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_rollout_trace, sizeof(g_rollout_trace)));
}
extern "C" int f16_debug_rollout_starts(unsigned long long *out)
{
    return static_cast<int>(cudaMemcpyFromSymbol(out, g_rollout_starts, sizeof(g_rollout_starts)));
}
#endif

}  // namespace f16
