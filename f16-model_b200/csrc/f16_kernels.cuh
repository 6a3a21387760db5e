// f16_kernels.cuh -- the CUDA kernels (sm_100a) and their typed launchers.
//
// Included by exactly two translation units: f16_kernels_f32.cu (FMA contraction on) and
// f16_kernels_f64.cu (-fmad=false: the parity build).  Launchers are declared in
// f16_launch.h and called from the C ABI (f16_capi.cu).
//
// Kernel shape (SURVEY.md section 7 / BASELINE north_star):
//   * one thread per aircraft, state SoA real[9][n] -> every global access of a warp is one
//     contiguous 128 B (fp32) / 256 B (fp64) segment;
//   * persistent grid (SMs x resident CTAs, grid-stride over 256-env tiles);
//   * the packed table image (8.4 KB fp32 / 16.7 KB fp64) is staged ONCE per CTA into shared
//     memory by a single TMA bulk copy (cp.async.bulk + mbarrier);
//   * the per-env inputs of every full tile (state rows, first action row, counters) are staged
//     by TMA bulk copies into a 2-stage shared-memory ring two tiles ahead of the compute; the
//     last warp to drain a stage re-arms it (no CTA-wide barrier, no fixed producer warp);
//   * the nine state values, the episode counter and the return live in registers across the K
//     env-steps of a launch; only action / reward / done stream per step;
//   * episode termination, gymnasium-style autoreset (counter-based Philox RNG), the observation
//     and hierarchical episode statistics (shuffle -> shared atomics -> one flush per CTA) are
//     fused into the same kernel;
//   * launched with programmatic dependent launch: the next launch's prologue overlaps this tail.
#pragma once
#include <cstdint>

#include "f16_device.cuh"
#include "f16_launch.h"

namespace f16 {

#ifndef F16_CTA_THREADS
#define F16_CTA_THREADS 256
#endif
constexpr int kCtaThreads = F16_CTA_THREADS;

// global stores of the step kernel's outputs.  F16_STREAMING_STORES=1 marks them evict-first (st.global.cs):
// +0.9 % in the bench's streaming steady state, but it would evict state that a single L2-resident batch
// re-reads on the next step, so it is off by default.
#if defined(F16_STREAMING_STORES) && F16_STREAMING_STORES
#define F16_ST(ptr, val) __stcs((ptr), (val))
#else
#define F16_ST(ptr, val) (*(ptr) = (val))
#endif

// ---------------------------------------------------------------------------
// TMA bulk copy global -> shared with mbarrier completion (sm_90+/sm_100a)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(2000u)   // suspend-time hint [ns]: a waiting warp sleeps instead of burning issue slots
        : "memory");
}

// Stage the table image into this CTA's shared memory.  Thread 0 issues the copy; the caller
// may issue independent global loads before calling tables_wait().
template <typename R>
__device__ __forceinline__ void tables_issue(Tables<R> *dst, const Tables<R> *src, uint64_t *bar)
{
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) tma_bulk_g2s(dst, src, static_cast<uint32_t>(sizeof(Tables<R>)), bar);
}
__device__ __forceinline__ void tables_wait(uint64_t *bar) { mbar_wait(bar, 0); }

// ---------------------------------------------------------------------------
// Philox4x32-10 counter RNG (Salmon et al. 2011): key (seed), counter (env, episode, stream, step)
// ---------------------------------------------------------------------------
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
// uniform in [0,1) with 24 (float) / 32 (double) random bits
template <typename R>
__device__ __forceinline__ R u01(uint32_t v);
template <>
__device__ __forceinline__ float u01<float>(uint32_t v) { return static_cast<float>(v >> 8) * (1.0f / 16777216.0f); }
template <>
__device__ __forceinline__ double u01<double>(uint32_t v) { return static_cast<double>(v) * (1.0 / 4294967296.0); }

// ---------------------------------------------------------------------------
// env-level helpers (F16model/env/env.py)
// ---------------------------------------------------------------------------
// state_transform + normalize (env.py:83-101,119-123; utils/calc.py:14-20)
template <typename R, bool FAST>
__device__ __forceinline__ void make_obs(R Oy, R wz_obs, R V, R ref, R (&o)[5])
{
    using M = Mth<R, FAST>;
    if constexpr (sizeof(R) == 4) {   // same affine maps, one FMA / FMUL each
        o[0] = Oy * R(1.0 / 35000.0);
        o[1] = wz_obs * R(1.0 / (2.0 * K::wz_bound)) + R(0.5);
        o[2] = V * R(1.0 / 500.0);
        o[3] = ref;
        o[4] = R(0.5) * ref - R(0.5) * wz_obs;
    } else {
        o[0] = M::divc(Oy - R(0), R(35000.0), R(1.0 / 35000.0));
        o[1] = M::divc(wz_obs - R(-K::wz_bound), R(K::wz_bound - (-K::wz_bound)), R(1.0 / (2.0 * K::wz_bound)));
        o[2] = M::divc(V - R(0), R(500.0), R(1.0 / 500.0));
        o[3] = ref;
        o[4] = R(0.5) * (ref - wz_obs);
    }
}

template <typename R>
__device__ __forceinline__ void store_obs(R *obs, int layout, unsigned n, unsigned i, const R (&o)[5])
{
    if (layout == 0) {
        R *p = obs + static_cast<size_t>(i) * 5;   // one address, five immediate-offset stores (L2 merges the sectors)
#pragma unroll
        for (int j = 0; j < 5; ++j) F16_ST(p + j, o[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 5; ++j) F16_ST(obs + static_cast<size_t>(j) * n + i, o[j]);
    }
}

// compute_reward (env.py:72-81), k = 1
template <typename R, bool FAST>
__device__ __forceinline__ R tracking_reward(R ref, R wz)
{
    using M = Mth<R, FAST>;
    const R e = (ref - wz) * R(K::rad2deg);
    const R ae = M::fabs(e);
    if constexpr (sizeof(R) == 4) {
        // 1 - |e|/(1+|e|) == 1/(1+|e|) in (0, 1]; 1 - e^2 <= 1: only the lower clip of the second term can act
        const R asym = M::div(R(1), R(1) + ae);
        const R lin = M::fmax(R(1) - e * e, R(0));
        return asym + R(0.04) * lin;
    } else {
        const R asym = clampr(R(1) - M::div(ae, R(1) + ae), R(0), R(1));
        const R lin = clampr(R(1) - e * e, R(0), R(1));
        return asym + R(0.04) * lin;
    }
}

// Episode statistics with warp primitives, called from the rare termination branch only: the lanes that
// are here together (__activemask) elect the lowest one, it gathers their returns and lengths by shuffle and
// adds them to the CTA's shared-memory accumulators; the CTA flushes to global memory once, at kernel end
// (hierarchical reduction: shuffle -> shared atomics -> 3 global atomics per CTA).  Out of line so that it
// costs the hot path no registers.
struct CtaStats {
    double ret;
    int count, length;
};
static __device__ __noinline__ void episode_stats_add(CtaStats *cta, double ret, int ep_len)
{
    const unsigned grp = __activemask();
    const unsigned leader = __ffs(grp) - 1;
    double sum_ret = 0.0;
    int sum_len = 0;
    for (unsigned m = grp; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        sum_ret += __shfl_sync(grp, ret, src);
        sum_len += __shfl_sync(grp, ep_len, src);
    }
    if ((threadIdx.x & 31u) == leader) {
        atomicAdd(&cta->count, __popc(grp));
        atomicAdd(&cta->length, sum_len);
        atomicAdd(&cta->ret, sum_ret);
    }
}

// Reference sample wz_ref[t] of one env (env/task.py: ReferenceSignal.wz_ref[episode_length]).
//   row mode      ref_bank is real[n_ref][ref_len] and ref_idx a row of it;
//   segment mode  (seg_len > 0) the signal is a concatenation of up to four equal-length segments followed by zeros -- the
//                 reference's random `combo` scenario: a rectangular pulse, three seconds of sine and a raised-cosine bump in
//                 shuffled order (env/task.py:65-103).  ref_bank is then the SEGMENT bank real[seg_len][seg_stride], stored
//                 sample-major, and ref_idx packs four 8-bit rows of it (segment s in bits [8s, 8s+8)).  The 6480 distinct combo
//                 signals (26 MB as rows) become 43 rows of 300 samples (52 KB, L1-resident), and the 32 envs of a warp, which
//                 sit at the same sample of different signals, read ONE 172-byte run instead of 32 different sectors.
template <typename R>
__device__ __forceinline__ R ref_sample(const EnvArgs<R> &A, int ref_idx, int t)
{
    t = min(t, A.ref_len - 1);
    if (A.seg_len > 0) {
        const int L = A.seg_len;
        const int seg = (t >= L) + (t >= 2 * L) + (t >= 3 * L);
        const unsigned row = (static_cast<unsigned>(ref_idx) >> (8 * seg)) & 0xffu;
        return A.ref_bank[static_cast<unsigned>(t - seg * L) * static_cast<unsigned>(A.seg_stride) + row];
    }
    return A.ref_bank[static_cast<unsigned>(ref_idx) * static_cast<unsigned>(A.ref_len) + static_cast<unsigned>(t)];
}

// Draw / copy an episode's initial state (env.py:172-195,221-243) and pick its reference signal.
template <typename R>
__device__ __forceinline__ void init_episode(const EnvArgs<R> &A, int64_t i, uint32_t episode, bool keep_ref, R (&x)[9],
                                             int &ref_idx)
{
    const uint64_t seed = A.seed_dev ? *A.seed_dev : A.seed;
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const int64_t gid = A.id_base + i;   // global env id: sharding-invariant streams
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), episode, 0u), key);
    if (A.init_mode == 0) {
        // get_random_state: alpha=theta~U(+-5deg), Oy~U(2000,4000), wz~U(+-5deg/s), V~U(90,250)
        constexpr double d5 = 5.0 * (K::kPi / 180.0);
        const R al = R(-d5) + R(2.0 * d5) * u01<R>(r.x);
        x[0] = R(0);
        x[1] = R(2000) + R(2000) * u01<R>(r.y);
        x[2] = R(-d5) + R(2.0 * d5) * u01<R>(r.z);
        x[3] = al;
        x[4] = R(90) + R(160) * u01<R>(r.w);
        x[5] = al;
        x[6] = R(0);
        x[7] = R(0);
        x[8] = R(0);   // find_correct_thrust_position(0) == 0
    } else if (A.init_mode == 1) {
#pragma unroll
        for (int j = 0; j < 9; ++j) x[j] = A.x0[j];
    } else {
#pragma unroll
        for (int j = 0; j < 9; ++j) x[j] = A.init_state[j * A.n + i];
    }
    if (!keep_ref && A.n_ref > 1) {
        const uint4 q = philox4x32(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), episode, 1u), key);
        const uint32_t r = __umulhi(q.x, static_cast<uint32_t>(A.n_ref));
        ref_idx = A.ref_codes ? static_cast<int>(A.ref_codes[r]) : static_cast<int>(r);
    } else if (A.n_ref <= 1) {
        ref_idx = A.ref_codes ? static_cast<int>(A.ref_codes[0]) : 0;
    }
}

// observation noise (env.py:139-152): radians(N(0, 0.5)) on wz, Box-Muller on a Philox draw
template <typename R>
__device__ __forceinline__ R obs_noise(const EnvArgs<R> &A, int64_t i, uint32_t episode, int k)
{
    const uint64_t seed = A.seed_dev ? *A.seed_dev : A.seed;
    const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const int64_t gid = A.id_base + i;
    const uint4 r = philox4x32(make_uint4(static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32), episode,
                                          2u + static_cast<uint32_t>(k)), key);
    const double u1 = (static_cast<double>(r.x) + 1.0) * (1.0 / 4294967296.0);   // (0,1]
    const double u2 = static_cast<double>(r.y) * (1.0 / 4294967296.0);
    const double z = ::sqrt(-2.0 * ::log(u1)) * ::cospi(2.0 * u2);
    return R(0.5 * z * (K::kPi / 180.0));
}

// ---------------------------------------------------------------------------
// THE hot kernel: K x F16.step for n envs
// ---------------------------------------------------------------------------
// Resident CTAs per SM the register allocator is asked to make room for: fp32 64 regs/thread
// (4 x 256 threads = 32 warps/SM hide the HBM latency of the K = 1 gym step), fp64 128 regs.
template <typename R> struct MinCtas;
#ifndef F16_MINCTAS_F32
#define F16_MINCTAS_F32 4
#endif
template <> struct MinCtas<float> { static constexpr int v = F16_MINCTAS_F32; };
template <> struct MinCtas<double> { static constexpr int v = 2; };

// Input staging.  A tile is kCtaThreads consecutive envs; its per-env inputs are up to 16
// contiguous SoA rows (9 state rows, the first action row, ep_len, ref_idx, total_return, the two
// position carries, done).  One elected thread streams them global -> shared with TMA bulk
// copies two tiles ahead of the compute (2-stage ring, one mbarrier per stage), so HBM latency
// is hidden by the pipeline instead of by occupancy, and nobody spends registers or address
// arithmetic on the loads.  Tiles that are partial or not 16-byte aligned fall back to plain
// coalesced loads.
constexpr int kInRows = 16;
enum { ROW_ACTION = 9, ROW_EPLEN = 10, ROW_REFIDX = 11, ROW_RET = 12, ROW_OXLO = 13, ROW_OYLO = 14, ROW_DONE = 15 };

template <typename R>
struct alignas(16) StepSmem {
    Tables<R> tbl;
    R in[2][kInRows][kCtaThreads];
    uint64_t bar_tbl;
    uint64_t bar_in[2];      // "full": TMA bytes of a stage have landed
    unsigned drained[2];     // warps that have copied a stage into registers; the last one refills it
    CtaStats stats;          // episodes finished by this CTA (flushed to global once, at kernel end)
};

template <typename R, bool AUTORESET, bool COMP>
__device__ __forceinline__ void issue_tile(const EnvArgs<R> &A, const StepArgs<R> &S, StepSmem<R> *sm, int stage, unsigned i0)
{
    // called by ONE thread; i0 = first env of a full, aligned tile
    const unsigned n = static_cast<unsigned>(A.n);
    constexpr uint32_t rb = kCtaThreads * sizeof(R), ib = kCtaThreads * 4u;
    constexpr uint32_t total = 11u * rb + 2u * ib + (COMP ? 2u * rb : 0u) + (AUTORESET ? 0u : kCtaThreads);
    uint64_t *bar = &sm->bar_in[stage];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total) : "memory");
    auto cp = [&](void *dst, const void *src, uint32_t bytes) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                     "l"(src), "r"(bytes), "r"(smem_u32(bar))
                     : "memory");
    };
    R(*in)[kCtaThreads] = sm->in[stage];
#pragma unroll
    for (int j = 0; j < 9; ++j) cp(in[j], A.state + static_cast<size_t>(j) * n + i0, rb);
    cp(in[ROW_ACTION], S.actions + i0, rb);
    cp(in[ROW_EPLEN], A.ep_len + i0, ib);
    cp(in[ROW_REFIDX], A.ref_idx + i0, ib);
    cp(in[ROW_RET], A.total_return + i0, rb);
    if (COMP) {
        cp(in[ROW_OXLO], A.pos_comp + i0, rb);
        cp(in[ROW_OYLO], A.pos_comp + static_cast<size_t>(n) + i0, rb);
    }
    if (!AUTORESET) cp(in[ROW_DONE], A.done + i0, kCtaThreads);
}

// SINGLE: the K = 1 gym step (one action row, one observation) with the K-loop bookkeeping compiled out.
// COMP: Kahan carries for the position integrals (fp32 only).
template <typename R, bool FAST, bool AUTORESET, bool SINGLE, bool COMP>
__global__ void __launch_bounds__(kCtaThreads, MinCtas<R>::v) env_step_kernel(const EnvArgs<R> A, const StepArgs<R> S)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StepSmem<R> *sm = reinterpret_cast<StepSmem<R> *>(smem_raw);
    const Tables<R> &tbl = sm->tbl;

    // 32-bit element indices (the C ABI rejects batches whose arrays exceed 2^32 elements)
    const unsigned n = static_cast<unsigned>(A.n);
    const unsigned i_begin = static_cast<unsigned>(S.i_begin), i_end = static_cast<unsigned>(S.i_end);
    const unsigned n_tiles = (i_end - i_begin + kCtaThreads - 1) / kCtaThreads;
    const unsigned n_full = (i_end - i_begin) / kCtaThreads;     // tiles [0, n_full) are full
    const bool tma = S.tma_ok != 0;
    const R dt = A.dt;
    constexpr bool comp = COMP;

    if (threadIdx.x == 0) {
        mbar_init(&sm->bar_tbl, 1);
        mbar_init(&sm->bar_in[0], 1);
        mbar_init(&sm->bar_in[1], 1);
        sm->drained[0] = 0;
        sm->drained[1] = 0;
        sm->stats.count = 0;
        sm->stats.length = 0;
        sm->stats.ret = 0.0;
    }
    __syncthreads();
    // Programmatic dependent launch: let the NEXT launch in the stream start its own prologue (barrier
    // init, table staging) while this grid is still running ...
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) tma_bulk_g2s(&sm->tbl, A.tables, static_cast<uint32_t>(sizeof(Tables<R>)), &sm->bar_tbl);
    // ... and do not touch anything the PREVIOUS launch may still be writing (state, counters, actions)
    // until it has completed.  (No-op when the launch carries no programmatic dependency.)
    // pdl_mode 2 = this launch is the ragged tail of a step whose full tiles run in the two-env kernel just before it: that kernel
    // released us only after ITS wait (everything older is complete) and touches other envs, so the tail runs alongside it and
    // waits at its end instead -- its completion then still implies the completion of everything before it.
    if (S.pdl_mode != 2) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) {
        if (tma) {   // prologue: only the FIRST tile of every CTA -- the opening burst is half as deep, so it lands sooner
            const unsigned t0 = blockIdx.x;
            if (t0 < n_full) issue_tile<R, AUTORESET, COMP>(A, S, sm, 0, i_begin + t0 * kCtaThreads);
        }
    }
    bool tables_ready = false;

    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const unsigned i = i_begin + tile * kCtaThreads + threadIdx.x;
        const bool live = i < i_end;
        const bool staged = tma && tile < n_full;              // CTA-uniform
        const int stage = it & 1;
        R x[9];
        R ox_lo = R(0), oy_lo = R(0), a0 = R(0);
        int ep_len = 0, ref_idx = 0;
        uint32_t done_sticky = 0;
        bool was_reset = false;
        R ret = R(0);
        if (staged) {
            mbar_wait(&sm->bar_in[stage], (it >> 1) & 1u);
            if (it == 0 && threadIdx.x == 0) {   // first tile has landed: now put the second one in flight
                const unsigned t1 = tile + gridDim.x;
                if (t1 < n_full) issue_tile<R, AUTORESET, COMP>(A, S, sm, 1, i_begin + t1 * kCtaThreads);
            }
            const R(*in)[kCtaThreads] = sm->in[stage];
            const unsigned t = threadIdx.x;
#pragma unroll
            for (int j = 0; j < 9; ++j) x[j] = in[j][t];
            a0 = in[ROW_ACTION][t];
            ep_len = reinterpret_cast<const int *>(in[ROW_EPLEN])[t];
            ref_idx = reinterpret_cast<const int *>(in[ROW_REFIDX])[t];
            ret = in[ROW_RET][t];
            if (comp) {
                ox_lo = in[ROW_OXLO][t];
                oy_lo = in[ROW_OYLO][t];
            }
            if (!AUTORESET) done_sticky = reinterpret_cast<const uint8_t *>(in[ROW_DONE])[t];
            // This warp has drained the stage.  The LAST warp of the CTA to get here refills the stage
            // with the tile two ahead -- at the earliest possible moment and with no CTA-wide barrier.
            // (A fixed producer thread does not work: the SM's scheduler favours high warp ids, so
            // warp 0 runs last and every other warp ends up waiting for its refill.)
            __syncwarp();
            if ((threadIdx.x & 31u) == 0) {
                unsigned prior;   // plain atom.shared (atomicAdd would be wrapped in warp-aggregation code)
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(prior) : "r"(smem_u32(&sm->drained[stage])) : "memory");
                if (prior == kCtaThreads / 32 - 1) {
                    sm->drained[stage] = 0;
                    const unsigned next = tile + 2 * gridDim.x;
                    if (next < n_full) issue_tile<R, AUTORESET, COMP>(A, S, sm, stage, i_begin + next * kCtaThreads);
                }
            }
        } else if (live) {
            const R *p = A.state + i;   // pointer bumps, not j * n + i: no per-row induction variables in the tile loop
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                x[j] = *p;
                p += n;
            }
            a0 = S.actions[i];
            ep_len = A.ep_len[i];
            ref_idx = A.ref_idx[i];
            done_sticky = AUTORESET ? 0u : A.done[i];
            ret = A.total_return[i];
            if (comp) {
                ox_lo = A.pos_comp[i];
                oy_lo = A.pos_comp[n + i];
            }
        }
        const R pa_in = x[8];
        if (!tables_ready) {   // first tile only: the input loads above overlapped the table copy
            mbar_wait(&sm->bar_tbl, 0);
            tables_ready = true;
        }
        if (!live) continue;

        R o[5];
        bool have_obs = false;
        R a_k = a0;
        const int K_steps = SINGLE ? 1 : S.K;
        const bool every = !SINGLE && S.obs_every_step;
        for (int k = 0; k < K_steps; ++k) {
            const unsigned kn = static_cast<unsigned>(k) * n + i;
            // issue next step's loads before this step's arithmetic: the reference sample (a gather
            // that depends on ref_idx / ep_len) and the next action row
            const R ref = ref_sample<R>(A, ref_idx, ep_len);
            R a_next = R(0);
            if (!SINGLE && k + 1 < K_steps) a_next = S.actions[kn + n];
            R reward = R(0);
            uint32_t done_now = done_sticky;
            if (!done_sticky) {
                // F16.rescale_action (env.py:200-211): clip to [-1,1], float32-cast bounds
                const R a = clampr(a_k, R(-1), R(1));
                const R u_stab = R(K::act_lo) + R(0.5) * (a + R(1.0)) * R(K::act_span);
                euler_step<R, FAST, true>(tbl, x, u_stab, R(0), dt, comp, ox_lo, oy_lo);   // Control(action, 0), env.py:49
                // check_state (env.py:103-117) on the post-step state, pre-increment counter
                R pen = R(0);
                if (!(x[1] > R(300))) { pen += R(-1000); done_now = 1; }
                if (x[1] >= R(30000)) { pen += R(-1000); done_now = 1; }
                if (!(Mth<R, FAST>::fabs(x[2]) < R(K::wz_term))) { pen += R(-1000); done_now = 1; }   // >= for finite wz; NaN terminates too
                if (ep_len == A.horizon) done_now = 1;
                reward = pen + tracking_reward<R, FAST>(ref, x[2]);
                R wz_obs = x[2];
                if (A.noise) wz_obs += obs_noise<R>(A, i, A.episode_no[i], ep_len);
                make_obs<R, FAST>(x[1], wz_obs, x[4], ref, o);
                have_obs = true;
                ep_len += 1;
                ret += reward;
                if (done_now) {
                    if (S.stats) episode_stats_add(&sm->stats, static_cast<double>(ret), ep_len);
                    if (S.final_obs) store_obs<R>(S.final_obs, A.obs_layout, n, i, o);
                    if (S.final_ep_len) S.final_ep_len[i] = ep_len;
                    if (S.final_return) S.final_return[i] = ret;
                    if (AUTORESET) {
                        // gymnasium SyncVectorEnv: reset immediately, hand back the reset observation.
                        // Everything only a reset touches (episode counter, reference row, V) is read and
                        // written here, off the per-step path.
                        const uint32_t episode = A.episode_no[i] + 1u;
                        A.episode_no[i] = episode;
                        init_episode<R>(A, i, episode, false, x, ref_idx);
                        A.ref_idx[i] = ref_idx;
                        was_reset = true;
                        ep_len = 0;
                        ret = R(0);
                        ox_lo = R(0);
                        oy_lo = R(0);
                        make_obs<R, FAST>(x[1], x[2], x[4], ref_sample<R>(A, ref_idx, 0), o);
                    } else {
                        done_sticky = 1;
                    }
                }
            }
            a_k = a_next;
            F16_ST(S.reward + kn, reward);
            F16_ST(S.done + kn, static_cast<uint8_t>(done_now));
            if (every) {
                if (!have_obs) {   // env was already done when the launch started
                    make_obs<R, FAST>(x[1], x[2], x[4], ref_sample<R>(A, ref_idx, max(ep_len - 1, 0)), o);
                    have_obs = true;
                }
                store_obs<R>(S.obs + static_cast<size_t>(k) * n * 5, A.obs_layout, n, i, o);
            }
        }
        // write back
        {
            R *p = A.state + i;   // row pointers by repeated "+= n": one IMAD.WIDE per row, no per-row induction variables
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                if (j == 4) {
                    if (AUTORESET && was_reset) F16_ST(p, x[4]);      // V only changes at a reset
                } else if (j == 8) {
                    if (x[8] != pa_in) F16_ST(p, x[8]);               // engine power sits at exactly 0 with the throttle closed: don't rewrite it
                } else {
                    F16_ST(p, x[j]);
                }
                p += n;
            }
        }
        F16_ST(A.ep_len + i, ep_len);
        F16_ST(A.total_return + i, ret);
        if (comp) {
            A.pos_comp[i] = ox_lo;
            A.pos_comp[n + i] = oy_lo;
        }
        if (!AUTORESET) A.done[i] = static_cast<uint8_t>(done_sticky);
        if (!every && S.obs) {
            if (!have_obs) make_obs<R, FAST>(x[1], x[2], x[4], ref_sample<R>(A, ref_idx, max(ep_len - 1, 0)), o);
            store_obs<R>(S.obs, A.obs_layout, n, i, o);
        }
    }
    if (!tables_ready) mbar_wait(&sm->bar_tbl, 0);   // CTA had no tile: still drain the copy before exit
    if (S.stats) {
        __syncthreads();
        if (threadIdx.x == 0 && sm->stats.count) {
            atomicAdd(S.stats + 0, static_cast<double>(sm->stats.count));
            atomicAdd(S.stats + 1, sm->stats.ret);
            atomicAdd(S.stats + 2, static_cast<double>(sm->stats.length));
        }
    }
    if (S.pdl_mode == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---------------------------------------------------------------------------
// Two aircraft per thread: the gym step (K = 1, autoreset, fp32) with instruction-level parallelism.
//
// The one-env kernel above is bound by dependency stalls (short scoreboard + fixed-latency waits are 37 % of
// its stall samples, issue slots 64 % busy) and by per-tile bookkeeping.  Here every thread advances two envs of a
// 256-env tile (128-thread CTAs, four per SM; a warp owns 64 consecutive envs, see the kernel); the common path of a step is ONE branch-free basic block holding both envs'
// arithmetic, so the scheduler interleaves two independent dependency chains, and the tile bookkeeping
// (barrier wait, ring refill, parameter loads) is paid once per two envs.  Rare events -- an altitude outside
// the troposphere, a termination with its autoreset -- are patched after the common path.  Handles full,
// 16-byte aligned tiles only; the caller runs the one-env kernel on the remainder.
// ---------------------------------------------------------------------------
#ifndef F16_CTA2_THREADS
#define F16_CTA2_THREADS 128
#endif
#ifndef F16_MINCTAS2S
#define F16_MINCTAS2S 4
#endif
#ifndef F16_CTA2K_THREADS
#define F16_CTA2K_THREADS 256
#endif
// CTA size by mode (a tile is two envs per thread).  Gym step (K = 1): 128 threads, 4 CTAs per SM -- measured 2 % faster
// than 256 x 2 on one stream and 5 % faster with overlapping launches (finer tiles shorten the fill and drain of every
// launch; 64 x 8 is 15 % slower).  K-step launches are issue-bound and 5 % faster with 256 x 2.
template <bool SINGLE>
struct Cta2 {
    static constexpr int threads = SINGLE ? F16_CTA2_THREADS : F16_CTA2K_THREADS;
    static constexpr int min_ctas = SINGLE ? F16_MINCTAS2S : 2;
    static constexpr int tile = 2 * threads;
};
constexpr int kIn2Rows = 13;   // 9 state rows, action, ep_len, ref_idx, total_return
#ifndef F16_STAGES2
#define F16_STAGES2 2
#endif
#ifndef F16_PREFETCH2
#define F16_PREFETCH2 1
#endif
constexpr int kStages2 = F16_STAGES2;      // depth of the gym-step kernel's input ring
#ifndef F16_STAGES2K
#define F16_STAGES2K 2
#endif
constexpr int kStages2K = F16_STAGES2K;    // ... of the K-step kernel's (its loads are amortised over K steps)
constexpr int kPrefetch2 = F16_PREFETCH2;  // tiles whose rows are prefetched into L2 before the grid dependency resolves

template <typename R, int CT, bool OBS_STAGING>
struct alignas(16) Step2Smem {
    static constexpr int TILE = 2 * CT;
    static constexpr int STAGES = OBS_STAGING ? kStages2 : kStages2K;   // OBS_STAGING == gym-step kernel
    Tables<R> tbl;
    R in[STAGES][kIn2Rows][TILE];
    uint64_t bar_tbl;
    uint64_t bar_in[STAGES];
    unsigned drained[STAGES];
    CtaStats stats;
    // AoS observations of the tile, written back by per-warp bulk stores (gym-step kernel only: in the K-step kernel the
    // 10 KB would push the shared-memory carve-out up a notch and cost L1, which its reference-row reads live in: -5 %)
    alignas(16) R obs_out[OBS_STAGING ? TILE * 5 : 4];
};

// L2 prefetch of one tile's input rows: legal before griddepcontrol.wait (it moves no data into the SM and L2 is
// the coherence point, so lines the previous grid is still writing are simply updated in place)
template <typename R, int CT>
__device__ __forceinline__ void prefetch_tile2(const EnvArgs<R> &A, const StepArgs<R> &S, unsigned i0)
{
    constexpr int TILE = 2 * CT;
    const unsigned n = static_cast<unsigned>(A.n);
    constexpr uint32_t rb = TILE * sizeof(R), ib = TILE * 4u;
    auto pf = [&](const void *src, uint32_t bytes) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
    };
#pragma unroll
    for (int j = 0; j < 9; ++j) pf(A.state + static_cast<size_t>(j) * n + i0, rb);
    pf(S.actions + i0, rb);
    pf(A.ep_len + i0, ib);
    pf(A.ref_idx + i0, ib);
    pf(A.total_return + i0, rb);
}

template <typename R, int CT, bool OBS_STAGING>
__device__ __forceinline__ void issue_tile2(const EnvArgs<R> &A, const StepArgs<R> &S, Step2Smem<R, CT, OBS_STAGING> *sm, int stage, unsigned i0)
{
    constexpr int TILE = 2 * CT;
    const unsigned n = static_cast<unsigned>(A.n);
    constexpr uint32_t rb = TILE * sizeof(R), ib = TILE * 4u;
    constexpr uint32_t total = 11u * rb + 2u * ib;
    uint64_t *bar = &sm->bar_in[stage];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(total) : "memory");
    auto cp = [&](void *dst, const void *src, uint32_t bytes) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                     "l"(src), "r"(bytes), "r"(smem_u32(bar))
                     : "memory");
    };
    R(*in)[TILE] = sm->in[stage];
#pragma unroll
    for (int j = 0; j < 9; ++j) cp(in[j], A.state + static_cast<size_t>(j) * n + i0, rb);
    cp(in[9], S.actions + i0, rb);
    cp(in[10], A.ep_len + i0, ib);
    cp(in[11], A.ref_idx + i0, ib);
    cp(in[12], A.total_return + i0, rb);
}

template <typename R, bool FAST, bool SINGLE>
__global__ void __launch_bounds__(Cta2<SINGLE>::threads, Cta2<SINGLE>::min_ctas) env_step2_kernel(const EnvArgs<R> A, const StepArgs<R> S)
{
    constexpr int CT = Cta2<SINGLE>::threads, TILE = 2 * CT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Smem = Step2Smem<R, CT, SINGLE>;
    Smem *sm = reinterpret_cast<Smem *>(smem_raw);
    const Tables<R> &tbl = sm->tbl;
    using M = Mth<R, FAST>;

    const unsigned n = static_cast<unsigned>(A.n);
    const unsigned i_begin = static_cast<unsigned>(S.i_begin);
    const unsigned n_tiles = static_cast<unsigned>(S.i_end - S.i_begin) / TILE;   // full tiles only
    const R dt = A.dt;
    const unsigned t = threadIdx.x;
    // A warp owns 64 CONSECUTIVE envs of the tile: lane l advances envs 64 w + l and 64 w + 32 + l.  Every global store of the
    // warp still covers 32 consecutive envs, and the warp's [64, 5] block of AoS observations (1280 B) as well as its 64
    // rewards / done flags are contiguous -- one bulk store each on the host route (full-size PCIe writes).
#ifndef F16_X2_WARP64
#define F16_X2_WARP64 1
#endif
#if F16_X2_WARP64
    constexpr unsigned E2 = 32;
    const unsigned tl = ((t & ~31u) << 1) | (t & 31u);
#else
    constexpr unsigned E2 = CT;
    const unsigned tl = t;
#endif

    if (t == 0) {
        mbar_init(&sm->bar_tbl, 1);
#pragma unroll
        for (int s = 0; s < Smem::STAGES; ++s) {
            mbar_init(&sm->bar_in[s], 1);
            sm->drained[s] = 0;
        }
        sm->stats.count = 0;
        sm->stats.length = 0;
        sm->stats.ret = 0.0;
    }
    __syncthreads();
    // (pdl_mode 1: a ragged-tail launch follows and must not start before everything older than this grid is complete)
    if (S.pdl_mode != 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (t == 0) {
        tma_bulk_g2s(&sm->tbl, A.tables, static_cast<uint32_t>(sizeof(Tables<R>)), &sm->bar_tbl);
#pragma unroll
        for (int p = 0; p < kPrefetch2; ++p) {
            const unsigned tp = blockIdx.x + p * gridDim.x;
            if (tp < n_tiles) prefetch_tile2<R, CT>(A, S, i_begin + tp * TILE);
        }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (S.pdl_mode == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (t == 0 && blockIdx.x < n_tiles) issue_tile2<R, CT, SINGLE>(A, S, sm, 0, i_begin + blockIdx.x * TILE);
    bool tables_ready = false;

    constexpr int NST = Smem::STAGES;
    static_assert(NST == 1 || NST == 2 || NST == 4, "ring position is derived from the tile counter");
    unsigned it = 0;
    for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const unsigned stage = it & (NST - 1), phase = (it / NST) & 1u;   // nothing ring-related stays live across the K loop
        const unsigned base = i_begin + tile * TILE;
        R x[2][9], a_in[2], ret[2];
        int ep_len[2], ref_idx[2];
        mbar_wait(&sm->bar_in[stage], phase);
        if (it == 0 && t == 0) {   // the rest of the opening burst goes out once the first tile has landed
#pragma unroll
            for (int s = 1; s < NST; ++s) {
                const unsigned ts = tile + s * gridDim.x;
                if (ts < n_tiles) issue_tile2<R, CT, SINGLE>(A, S, sm, s, i_begin + ts * TILE);
            }
        }
        {
            const R(*in)[TILE] = sm->in[stage];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const unsigned c = tl + e * E2;
#pragma unroll
                for (int j = 0; j < 9; ++j) x[e][j] = in[j][c];
                a_in[e] = in[9][c];
                ep_len[e] = reinterpret_cast<const int *>(in[10])[c];
                ref_idx[e] = reinterpret_cast<const int *>(in[11])[c];
                ret[e] = in[12][c];
            }
        }
        __syncwarp();
        if ((t & 31u) == 0) {
            unsigned prior;
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(prior) : "r"(smem_u32(&sm->drained[stage])) : "memory");
            if (prior == CT / 32 - 1) {
                sm->drained[stage] = 0;
                const unsigned next = tile + NST * gridDim.x;
                if (next < n_tiles) issue_tile2<R, CT, SINGLE>(A, S, sm, stage, i_begin + next * TILE);
            }
        }
        if (!tables_ready) {
            mbar_wait(&sm->bar_tbl, 0);
            tables_ready = true;
        }

        R pa_in[2] = {x[0][8], x[1][8]};
        bool was_reset[2] = {false, false};
        R o[2][5];
        const int K_steps = SINGLE ? 1 : S.K;
        for (int k = 0; k < K_steps; ++k) {
            // ---- common path: both envs in one basic block ------------------------------------------------
            R ref[2], Hg[2], Tk[2], pk[2], rho[2], asnd[2], a_next[2] = {R(0), R(0)};
            bool slow[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                ref[e] = ref_sample<R>(A, ref_idx[e], ep_len[e]);
                if (!SINGLE && k + 1 < K_steps) a_next[e] = S.actions[static_cast<unsigned>(k + 1) * n + base + tl + e * E2];
                Hg[e] = isa_geopotential<R, FAST>(x[e][1]);
                slow[e] = !(Hg[e] >= R(0) && Hg[e] < R(11000));
                if constexpr (FAST) isa_troposphere_direct<R>(Hg[e], rho[e], asnd[e]);
                else isa_troposphere<R, FAST>(Hg[e], Tk[e], pk[e]);
            }
            if (slow[0] | slow[1]) {   // rare: some other layer of the standard atmosphere
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (slow[e]) {
                        isa_layer<R, FAST>(Hg[e], Tk[e], pk[e]);
                        if constexpr (FAST) isa_finish<R, FAST>(Tk[e], pk[e], rho[e], asnd[e]);
                    }
            }
            R reward[2], ret_in[2];
            bool done[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                R dx[9];
                if constexpr (!FAST) isa_finish<R, FAST>(Tk[e], pk[e], rho[e], asnd[e]);
                const R a = clampr(a_in[e], R(-1), R(1));
                const R u_stab = R(K::act_lo) + R(0.5) * (a + R(1.0)) * R(K::act_span);
#ifdef F16_PROBE_NOMATH   // probe build: the memory pipeline alone (inputs -> outputs with a trivial right-hand side)
#pragma unroll
                for (int j = 0; j < 9; ++j) dx[j] = x[e][j] * R(1e-6) + u_stab * rho[e] * asnd[e] * R(1e-9);
#else
                rhs_atm<R, FAST, false, true>(tbl, x[e], u_stab, R(0), rho[e], asnd[e], dx);
#endif
#pragma unroll
                for (int j = 0; j < 9; ++j)
                    if (j != 4) x[e][j] = dx[j] * dt + x[e][j];
                x[e][2] = clampr(x[e][2], R(-K::wz_clip), R(K::wz_clip));
                // check_state (env.py:103-117): the flags here, the -1000 penalties in the rare branch below (0 + r == r)
                const bool d = !(x[e][1] > R(300)) | (x[e][1] >= R(30000)) | !(M::fabs(x[e][2]) < R(K::wz_term)) | (ep_len[e] == A.horizon);
                reward[e] = tracking_reward<R, FAST>(ref[e], x[e][2]);
                make_obs<R, FAST>(x[e][1], x[e][2], x[e][4], ref[e], o[e]);
                ep_len[e] += 1;
                ret_in[e] = ret[e];
                ret[e] += reward[e];
                done[e] = d;
                a_in[e] = a_next[e];
            }

            // ---- rare path: terminations (final_* outputs, statistics, autoreset) ---------------------------
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (done[e]) {
                    const unsigned i = base + tl + e * E2;
                    R pen = R(0);   // penalties stack (env.py:106-114); evaluated on the post-step state, like the flags above
                    if (!(x[e][1] > R(300))) pen += R(-1000);
                    if (x[e][1] >= R(30000)) pen += R(-1000);
                    if (!(M::fabs(x[e][2]) < R(K::wz_term))) pen += R(-1000);
                    if (pen != R(0)) {
                        reward[e] = pen + reward[e];
                        ret[e] = ret_in[e] + reward[e];
                    }
                    if (S.stats) episode_stats_add(&sm->stats, static_cast<double>(ret[e]), ep_len[e]);
                    if (S.final_obs) store_obs<R>(S.final_obs, A.obs_layout, n, i, o[e]);
                    if (S.final_ep_len) S.final_ep_len[i] = ep_len[e];
                    if (S.final_return) S.final_return[i] = ret[e];
                    const uint32_t episode = A.episode_no[i] + 1u;
                    A.episode_no[i] = episode;
                    init_episode<R>(A, i, episode, false, x[e], ref_idx[e]);
                    A.ref_idx[i] = ref_idx[e];
                    was_reset[e] = true;
                    ep_len[e] = 0;
                    ret[e] = R(0);
                    make_obs<R, FAST>(x[e][1], x[e][2], x[e][4], ref_sample<R>(A, ref_idx[e], 0), o[e]);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const unsigned kn = static_cast<unsigned>(k) * n + base + tl + e * E2;
                S.reward[kn] = reward[e];
                S.done[kn] = static_cast<uint8_t>(done[e]);
            }
        }

        // ---- write back ---------------------------------------------------------------------------------------
        {
            const unsigned i = base + tl;
            R *p = A.state + i;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                if (j == 4) {
                    if (was_reset[0]) p[0] = x[0][4];
                    if (was_reset[1]) p[E2] = x[1][4];
                } else if (j == 8) {
                    if (x[0][8] != pa_in[0]) p[0] = x[0][8];
                    if (x[1][8] != pa_in[1]) p[E2] = x[1][8];
                } else {
                    p[0] = x[0][j];
                    p[E2] = x[1][j];
                }
                p += n;
            }
            int32_t *pl = A.ep_len + i;
            pl[0] = ep_len[0];
            pl[E2] = ep_len[1];
            R *pt = A.total_return + i;
            pt[0] = ret[0];
            pt[E2] = ret[1];
            if (S.obs && !(SINGLE && S.obs_bulk)) {
                if (A.obs_layout == 0) {   // F16_OBS_AOS
                    R *po = S.obs + static_cast<size_t>(i) * 5u;
#if defined(F16_OBS_INTERLEAVED)
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        po[j] = o[0][j];
                        po[E2 * 5 + j] = o[1][j];
                    }
#else
                    // env-major: the five 4-byte stores of one env's row go out back to back (they share sectors)
#pragma unroll
                    for (int j = 0; j < 5; ++j) po[j] = o[0][j];
#pragma unroll
                    for (int j = 0; j < 5; ++j) po[E2 * 5 + j] = o[1][j];
#endif
                } else {
                    R *po = S.obs + i;
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        po[0] = o[0][j];
                        po[E2] = o[1][j];
                        po += n;
                    }
                }
            }
        }
        if (SINGLE && S.obs_bulk) {   // the host route is a gym step: compiled out of the K-step kernel
            // [N,5] observations: each warp stages its 64 envs' rows (1280 B, stride-5 words = conflict-free) and lane 0
            // writes them with one bulk store -- full lines instead of 20-byte-strided scalar stores, which is what
            // makes the same kernel usable on mapped host memory (PCIe sees large writes)
            const unsigned lane = t & 31u, w0 = tl & ~31u;   // first env of this warp's first run
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // last tile's stores have read the staging area
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                R *dst = sm->obs_out + (tl + e * E2) * 5u;
#pragma unroll
                for (int j = 0; j < 5; ++j) dst[j] = o[e][j];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                constexpr int RUNS = (E2 == 32) ? 1 : 2;           // one [64, 5] block, or two [32, 5] runs CT envs apart
                constexpr uint32_t run_bytes = (2 / RUNS) * 32u * 5u * sizeof(R);
#pragma unroll
                for (int e = 0; e < RUNS; ++e) {
                    const unsigned c0 = w0 + e * E2;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(S.obs + static_cast<size_t>(base + c0) * 5u),
                                 "r"(smem_u32(sm->obs_out + c0 * 5u)), "r"(run_bytes)
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (SINGLE && S.obs_bulk && (t & 31u) == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (!tables_ready) mbar_wait(&sm->bar_tbl, 0);
    if (S.stats) {
        __syncthreads();
        if (t == 0 && sm->stats.count) {
            atomicAdd(S.stats + 0, static_cast<double>(sm->stats.count));
            atomicAdd(S.stats + 1, sm->stats.ret);
            atomicAdd(S.stats + 2, static_cast<double>(sm->stats.length));
        }
    }
}

// ---------------------------------------------------------------------------
// F16.reset for masked envs
// ---------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kCtaThreads) env_reset_kernel(const EnvArgs<R> A, const uint8_t *mask, int keep_ref, R *obs)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        R x[9];
        int ref_idx = A.ref_idx[i];
        const uint32_t episode = A.episode_no[i] + 1u;
        if (keep_ref & 2) {   // F16_RESET_KEEP_STATE: the caller has just written the state; only counters + observation
#pragma unroll
            for (int j = 0; j < 9; ++j) x[j] = A.state[j * A.n + i];
        } else {
            init_episode<R>(A, i, episode, (keep_ref & 1) != 0, x, ref_idx);
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) A.state[j * A.n + i] = x[j];
        A.ep_len[i] = 0;
        if (A.total_return) A.total_return[i] = R(0);
        if (A.pos_comp) {
            A.pos_comp[i] = R(0);
            A.pos_comp[A.n + i] = R(0);
        }
        A.ref_idx[i] = ref_idx;
        A.done[i] = 0;
        A.episode_no[i] = episode;
        if (obs) {
            R o[5];
            make_obs<R, false>(x[1], x[2], x[4], ref_sample<R>(A, ref_idx, 0), o);
            store_obs<R>(obs, A.obs_layout, A.n, i, o);
        }
    }
}

// ---------------------------------------------------------------------------
// model- and data-level kernels (trim search, parity probes)
// ---------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kCtaThreads) rhs_kernel(const Tables<R> *tables, int64_t n, const R *x, const R *u, R *dx)
{
    __shared__ Tables<R> tbl;
    __shared__ alignas(8) uint64_t bar;
    tables_issue(&tbl, tables, &bar);
    tables_wait(&bar);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        R xs[9], d[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) xs[j] = x[j * n + i];
        rhs<R, false, true>(tbl, xs, u[i], u[n + i], d);
#pragma unroll
        for (int j = 0; j < 9; ++j) dx[j * n + i] = d[j];
    }
}

template <typename R>
__global__ void __launch_bounds__(kCtaThreads) model_step_kernel(const Tables<R> *tables, int64_t n, R *x, const R *u, R dt, int K)
{
    __shared__ Tables<R> tbl;
    __shared__ alignas(8) uint64_t bar;
    tables_issue(&tbl, tables, &bar);
    tables_wait(&bar);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        R xs[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) xs[j] = x[j * n + i];
        const R us = u[i], ut = u[n + i];
        for (int k = 0; k < K; ++k) euler_step<R, false>(tbl, xs, us, ut, dt);
#pragma unroll
        for (int j = 0; j < 9; ++j) x[j * n + i] = xs[j];
    }
}

template <typename R>
__global__ void __launch_bounds__(kCtaThreads) coeffs_kernel(const Tables<R> *tables, int64_t n, const R *alpha, const R *fi, const R *wz,
                                                             const R *V, R *out)
{
    __shared__ Tables<R> tbl;
    __shared__ alignas(8) uint64_t bar;
    tables_issue(&tbl, tables, &bar);
    tables_wait(&bar);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const R qhat = Mth<R, false>::div(wz[i] * R(K::ba), R(2) * V[i]);
        R cx, cy, mz;
        aero_coeffs<R, false>(tbl, alpha[i], fi[i], qhat, cx, cy, mz);
        out[i] = cx;
        out[n + i] = cy;
        out[2 * n + i] = mz;
    }
}

template <typename R>
__global__ void __launch_bounds__(kCtaThreads) thrust_kernel(const Tables<R> *tables, int64_t n, const R *H, const R *Mach, const R *Pa, R *out)
{
    __shared__ Tables<R> tbl;
    __shared__ alignas(8) uint64_t bar;
    tables_issue(&tbl, tables, &bar);
    tables_wait(&bar);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = thrust<R, false>(tbl, H[i], Mach[i], Pa[i]);
}

template <typename R>
__global__ void __launch_bounds__(kCtaThreads) isa_kernel(int64_t n, const R *h, R *rho, R *a)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        R r, s;
        isa<R, false>(h[i], r, s);
        rho[i] = r;
        a[i] = s;
    }
}

// ---------------------------------------------------------------------------
// Batched trim search (F16model/utils/find_trim_position.py:11-113): one thread runs one
// Nelder-Mead search (scipy.optimize.fmin semantics: rho=1, chi=2, psi=sigma=0.5, 5 % initial
// simplex, xtol=ftol=1e-10, maxiter=1000, maxfun=5000) over u = (stab, throttle, alpha) for one
// (V, H) cell from one start, minimising the reference's cost
//   10*Oy_dot^2 + 10*wz_dot^2 + 15*theta_dot^2 + 30*V_dot^2 + 20*alpha_dot^2
// of the RHS at x0 = (0, H, 0, alpha, V, alpha, stab, 0, Pc(throttle)).  fp64 only.
// ---------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ R trim_cost(const Tables<R> &T, R V, R H, const R (&u)[3])
{
    const R thr = u[1];
    const R Pa = thr <= R(0.77) ? R(64.94) * thr : R(217.38) * thr - R(117.38);   // find_correct_thrust_position
    const R x[9] = {R(0), H, R(0), u[2], V, u[2], u[0], R(0), Pa};
    R d[9];
    rhs<R, false, true>(T, x, u[0], thr, d);
    // np.matmul(weight, step**2) over [Ox,Oy,wz,theta,V,alpha,stab,dstab,Pa] with weight (0,10,10,15,30,20,0,0,0)
    R c = R(0) * (d[0] * d[0]);
    c += R(10) * (d[1] * d[1]);
    c += R(10) * (d[2] * d[2]);
    c += R(15) * (d[3] * d[3]);
    c += R(30) * (d[4] * d[4]);
    c += R(20) * (d[5] * d[5]);
    return c;
}

// One thread = one Nelder-Mead search at a time, written as a STATE MACHINE around a single cost evaluation: every trip of the loop
// evaluates the cost at one query point, and the bookkeeping between two evaluations (accept / expand / contract / shrink, the stable sort,
// the convergence test, writing a finished search out and loading the thread's next one) is a handful of cheap divergent branches.  The
// expensive part -- the right-hand side with its table lookups, ~300 FP64 instructions -- is therefore executed by ALL lanes of a warp on
// every trip, whatever iteration, branch or search each lane is in.  (The first version ran one search per thread as straight-line code with
// seven inlined cost evaluations: a warp stayed until its slowest search converged and lanes in different branches took turns -- ncu: 30 % of
// the FP64 lanes active.)  The arithmetic of a search is unchanged, so results are bit-identical.
template <typename R>
__global__ void __launch_bounds__(128) trim_kernel(const Tables<R> *tables, int n_cells, int n_starts, const R *Vc, const R *Hc,
                                                    const R *starts /*[S][3]*/, R *out_x /*[cells][S][3]*/, R *out_cost /*[cells][S]*/,
                                                    int *out_iters /*[cells][S] or null*/)
{
    __shared__ Tables<R> tbl;
    __shared__ alignas(8) uint64_t bar;
    tables_issue(&tbl, tables, &bar);
    tables_wait(&bar);
    const int64_t total = (int64_t)n_cells * n_starts;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = id < total;

    enum { PH_INIT, PH_REFLECT, PH_EXPAND, PH_CONTRACT_OUT, PH_CONTRACT_IN, PH_SHRINK };
    R V = R(200), H = R(3000), sim[4][3], fs[4] = {R(0), R(0), R(0), R(0)}, xr[3] = {R(0), R(0), R(0)}, fxr = R(0);
    R xq[3] = {R(-0.05), R(0.3), R(0.05)};                   // a harmless query for lanes without work
    int phase = PH_INIT, k = 0, fcalls = 0, iters = 1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 3; ++j) sim[q][j] = xq[j];

    auto start_search = [&]() {
        const int cell = (int)(id / n_starts), st = (int)(id % n_starts);
        V = Vc[cell];
        H = Hc[cell];
#pragma unroll
        for (int j = 0; j < 3; ++j) sim[0][j] = starts[st * 3 + j];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
            for (int j = 0; j < 3; ++j) sim[kk + 1][j] = sim[0][j];
            sim[kk + 1][kk] = sim[0][kk] != R(0) ? (R(1) + R(0.05)) * sim[0][kk] : R(0.00025);
        }
        fcalls = 0;
        iters = 1;
        phase = PH_INIT;
        k = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) xq[j] = sim[0][j];
    };
    auto sort4 = [&]() {   // stable insertion sort == np.argsort(kind='stable') on 4 keys
#pragma unroll
        for (int a = 1; a < 4; ++a)
#pragma unroll
            for (int b = a; b > 0; --b) {
                if (fs[b] < fs[b - 1]) {
                    const R tf = fs[b]; fs[b] = fs[b - 1]; fs[b - 1] = tf;
#pragma unroll
                    for (int j = 0; j < 3; ++j) { const R tx = sim[b][j]; sim[b][j] = sim[b - 1][j]; sim[b - 1][j] = tx; }
                } else {
                    break;
                }
            }
    };
    auto vertex_to_query = [&](int kk) {   // xq = sim[kk] without a dynamically indexed register array
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (kk == q) {
#pragma unroll
                for (int j = 0; j < 3; ++j) xq[j] = sim[q][j];
            }
    };
    auto store_f = [&](int kk, R f) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (kk == q) fs[q] = f;
    };
    auto centroid_query = [&](R a, R b) {   // xq = a * xbar + b * worst, xbar = centroid of the three best (scipy's expression order)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const R xbar = ((sim[0][j] + sim[1][j]) + sim[2][j]) / R(3);
            xq[j] = a * xbar + b * sim[3][j];
        }
    };
    auto accept = [&](const R (&x)[3], R f) {
#pragma unroll
        for (int j = 0; j < 3; ++j) sim[3][j] = x[j];
        fs[3] = f;
    };

    if (active) start_search();
    // The loop is warp-synchronous on purpose: __any_sync at its head re-converges the warp before EVERY cost evaluation (with `continue`s
    // out of the divergent bookkeeping the lanes drifted apart again: 13 of 32 lanes active per instruction, measured), and lanes that have
    // run out of searches keep evaluating a dummy point until the whole warp is done.
    while (__any_sync(0xffffffffu, active)) {
        const R f = trim_cost<R>(tbl, V, H, xq);
        if (!active) continue;
        ++fcalls;
        // ---- what the value means; `next`: 0 = the query for the next evaluation is set, 1 = iteration finished, 2 = shrink, 3 = simplex ready (after INIT)
        int next = 0;
        if (phase == PH_INIT || phase == PH_SHRINK) {
            store_f(k, f);
            ++k;
            if (k < 4) vertex_to_query(k);
            else next = phase == PH_INIT ? 3 : 1;
        } else if (phase == PH_REFLECT) {
            fxr = f;
            if (fxr < fs[0]) {
                centroid_query(R(1) + R(2), -R(2));
                phase = PH_EXPAND;
            } else if (fxr < fs[2]) {
                accept(xr, fxr);
                next = 1;
            } else if (fxr < fs[3]) {
                centroid_query(R(1) + R(0.5), -R(0.5));
                phase = PH_CONTRACT_OUT;
            } else {
                centroid_query(R(1) - R(0.5), R(0.5));
                phase = PH_CONTRACT_IN;
            }
        } else if (phase == PH_EXPAND) {
            if (f < fxr) accept(xq, f);
            else accept(xr, fxr);
            next = 1;
        } else if (phase == PH_CONTRACT_OUT) {
            if (f <= fxr) {
                accept(xq, f);
                next = 1;
            } else {
                next = 2;
            }
        } else {   // PH_CONTRACT_IN
            if (f < fs[3]) {
                accept(xq, f);
                next = 1;
            } else {
                next = 2;
            }
        }
        if (next == 2) {
#pragma unroll
            for (int q = 1; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 3; ++j) sim[q][j] = sim[0][j] + R(0.5) * (sim[q][j] - sim[0][j]);
            phase = PH_SHRINK;
            k = 1;
            vertex_to_query(1);
        } else if (next != 0) {
            if (next == 1) ++iters;
            sort4();
            // ---- top of scipy's loop: limits, convergence test, or the next reflection ----
            bool done = !(fcalls < 5000 && iters < 1000);
            if (!done) {
                R dx = R(0), df = R(0);
#pragma unroll
                for (int q = 1; q < 4; ++q) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) dx = fmax(dx, fabs(sim[q][j] - sim[0][j]));
                    df = fmax(df, fabs(fs[0] - fs[q]));
                }
                done = dx <= R(1e-10) && df <= R(1e-10);
            }
            if (done) {
#pragma unroll
                for (int j = 0; j < 3; ++j) out_x[id * 3 + j] = sim[0][j];
                out_cost[id] = fs[0];
                if (out_iters) out_iters[id] = iters;
                id += stride;
                active = id < total;
                if (active) start_search();
            } else {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const R xbar = ((sim[0][j] + sim[1][j]) + sim[2][j]) / R(3);
                    xr[j] = (R(1) + R(1)) * xbar - R(1) * sim[3][j];
                    xq[j] = xr[j];
                }
                phase = PH_REFLECT;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// typed launchers
// ---------------------------------------------------------------------------
inline int grid_for(int64_t n, const LaunchGeom &g)
{
    const int64_t tiles = (n + kCtaThreads - 1) / kCtaThreads;
    const int64_t cap = (int64_t)g.n_sm * g.ctas_per_sm;
    return (int)(tiles < cap ? tiles : cap);
}

template <typename R, bool FAST>
cudaError_t launch_env_step_t(const EnvArgs<R> &A, const StepArgs<R> &S, const LaunchGeom &g, cudaStream_t st)
{
    const int grid = grid_for(S.i_end - S.i_begin, g);
    const size_t smem = sizeof(StepSmem<R>);
    const bool single = S.K == 1;
    const bool comp = A.pos_comp != nullptr && sizeof(R) == 4;
    if constexpr (sizeof(R) == 4) {
        // the gym step of the throughput mode: two envs per thread on the full 256-env tiles, one-env kernel on the rest
        if (g.two_per_thread && A.autoreset && !comp && !A.noise && S.tma_ok && !(S.obs_every_step && !single) &&
            (S.i_end - S.i_begin) >= (single ? Cta2<true>::tile : Cta2<false>::tile)) {
            const int tile2 = single ? Cta2<true>::tile : Cta2<false>::tile;
            const int64_t n2 = (S.i_end - S.i_begin) / tile2;
            const int64_t cap2 = static_cast<int64_t>(g.n_sm) * (single ? g.ctas2_per_sm : g.ctas2k_per_sm);
            cudaLaunchConfig_t lc2 = {};
            lc2.gridDim = dim3(static_cast<unsigned>(n2 < cap2 ? n2 : cap2));
            lc2.blockDim = dim3(tile2 / 2);
            lc2.dynamicSmemBytes = single ? sizeof(Step2Smem<R, Cta2<true>::threads, true>) : sizeof(Step2Smem<R, Cta2<false>::threads, false>);
            lc2.stream = st;
            cudaLaunchAttribute at2[1];
            at2[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at2[0].val.programmaticStreamSerializationAllowed = 1;
            lc2.attrs = at2;
            lc2.numAttrs = g.pdl ? 1 : 0;
            StepArgs<R> head = S, rest = S;
            rest.i_begin = S.i_begin + n2 * tile2;
            const bool split = rest.i_begin < rest.i_end && g.pdl && S.pdl_mode == 0;
            if (split) {   // the ragged tail runs alongside the full tiles instead of after them
                head.pdl_mode = 1;
                rest.pdl_mode = 2;
            }
            cudaError_t e2 = single ? cudaLaunchKernelEx(&lc2, env_step2_kernel<R, FAST, true>, A, head)
                                    : cudaLaunchKernelEx(&lc2, env_step2_kernel<R, FAST, false>, A, head);
            if (e2 != cudaSuccess) return e2;
            if (rest.i_begin >= rest.i_end) return cudaGetLastError();
            LaunchGeom g1 = g;
            g1.two_per_thread = 0;
            return launch_env_step_t<R, FAST>(A, rest, g1, st);
        }
    }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kCtaThreads);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap this launch's prologue with the previous tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = g.pdl ? 1 : 0;
    cudaError_t le = cudaSuccess;
#define F16_LAUNCH(AR, SG, CP) le = cudaLaunchKernelEx(&lc, env_step_kernel<R, FAST, AR, SG, CP>, A, S)
    if (comp) {
        if constexpr (sizeof(R) == 4) {
            if (A.autoreset) { if (single) F16_LAUNCH(true, true, true); else F16_LAUNCH(true, false, true); }
            else             { if (single) F16_LAUNCH(false, true, true); else F16_LAUNCH(false, false, true); }
        }
    } else {
        if (A.autoreset) { if (single) F16_LAUNCH(true, true, false); else F16_LAUNCH(true, false, false); }
        else             { if (single) F16_LAUNCH(false, true, false); else F16_LAUNCH(false, false, false); }
    }
#undef F16_LAUNCH
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}

// Opt the step kernels into > 48 KB of dynamic shared memory and report resident CTAs per SM.
template <typename R, bool FAST, bool AUTORESET>
int prepare_env_step()
{
    int b1 = 0, bk = 0;
    auto prep = [](auto fn, int &b) {
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(StepSmem<R>)));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fn, kCtaThreads, sizeof(StepSmem<R>));
    };
    prep(env_step_kernel<R, FAST, AUTORESET, true, false>, b1);
    prep(env_step_kernel<R, FAST, AUTORESET, false, false>, bk);
    if constexpr (sizeof(R) == 4) {
        int c1 = 0, ck = 0;
        prep(env_step_kernel<R, FAST, AUTORESET, true, true>, c1);
        prep(env_step_kernel<R, FAST, AUTORESET, false, true>, ck);
    }
    return b1 < bk ? b1 : bk;
}

// two-envs-per-thread gym-step kernel: opt into its shared memory, report resident CTAs per SM
// (returns the gym-step kernel's count in the low byte, the K-step kernel's in the next one)
template <typename R, bool FAST>
int prepare_env_step2()
{
    auto fn1 = env_step2_kernel<R, FAST, true>;
    auto fnk = env_step2_kernel<R, FAST, false>;
    constexpr int s1 = static_cast<int>(sizeof(Step2Smem<R, Cta2<true>::threads, true>)), sk = static_cast<int>(sizeof(Step2Smem<R, Cta2<false>::threads, false>));
    cudaFuncSetAttribute(fn1, cudaFuncAttributeMaxDynamicSharedMemorySize, s1);
    cudaFuncSetAttribute(fnk, cudaFuncAttributeMaxDynamicSharedMemorySize, sk);
    int b1 = 0, bk = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, fn1, Cta2<true>::threads, s1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bk, fnk, Cta2<false>::threads, sk);
    return (b1 & 0xff) | ((bk & 0xff) << 8);
}

}  // namespace f16
