// f16_kernels.cuh -- the CUDA kernels (sm_100a) and their typed launchers.
//
// Included by exactly two translation units: f16_kernels_f32.cu (FMA contraction on) and
// f16_kernels_f64.cu (-fmad=false: the parity build).  Launchers are declared in
// f16_launch.h and called from the C ABI (f16_capi.cu).
//
// Kernel shape (SURVEY.md section 7 / BASELINE north_star):
//   * one thread per aircraft, state SoA real[9][n] -> every global access of a warp is one
//     contiguous 128 B (fp32) / 256 B (fp64) segment;
//   * the packed table image (8.4 KB fp32 / 16.7 KB fp64) is staged ONCE per CTA into
//     shared memory by a single TMA bulk copy (cp.async.bulk + mbarrier) that overlaps the
//     CTA's first state loads; the grid is persistent (SMs x resident CTAs, grid-stride over
//     256-env tiles) so the staging cost is paid ~10^3 times per launch, not per tile;
//   * the nine state values, the episode counter and the return live in registers across
//     the K env-steps of a launch; only action / reward / done stream per step;
//   * episode termination, gymnasium-style autoreset (counter-based Philox RNG) and the
//     observation are fused into the same kernel.
#pragma once
#include <cstdint>

#include "f16_device.cuh"
#include "f16_launch.h"

namespace f16 {

constexpr int kCtaThreads = 256;

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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
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
    o[0] = M::divc(Oy - R(0), R(35000.0), R(1.0 / 35000.0));
    o[1] = M::divc(wz_obs - R(-K::wz_bound), R(K::wz_bound - (-K::wz_bound)), R(1.0 / (2.0 * K::wz_bound)));
    o[2] = M::divc(V - R(0), R(500.0), R(1.0 / 500.0));
    o[3] = ref;
    o[4] = R(0.5) * (ref - wz_obs);
}

template <typename R>
__device__ __forceinline__ void store_obs(R *obs, int layout, int64_t n, int64_t i, const R (&o)[5])
{
    if (layout == 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) obs[i * 5 + j] = o[j];
    } else {
#pragma unroll
        for (int j = 0; j < 5; ++j) obs[j * n + i] = o[j];
    }
}

// compute_reward (env.py:72-81), k = 1
template <typename R, bool FAST>
__device__ __forceinline__ R tracking_reward(R ref, R wz)
{
    using M = Mth<R, FAST>;
    const R e = (ref - wz) * R(K::rad2deg);
    const R ae = M::fabs(e);
    const R asym = clampr(R(1) - M::div(ae, R(1) + ae), R(0), R(1));
    const R lin = clampr(R(1) - e * e, R(0), R(1));
    return asym + R(0.04) * lin;
}

// Draw / copy an episode's initial state (env.py:172-195,221-243) and pick its reference signal.
template <typename R>
__device__ __forceinline__ void init_episode(const EnvArgs<R> &A, int64_t i, uint32_t episode, bool keep_ref, R (&x)[9],
                                             int &ref_idx)
{
    const uint2 key = make_uint2(static_cast<uint32_t>(A.seed), static_cast<uint32_t>(A.seed >> 32));
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
        ref_idx = static_cast<int>(__umulhi(q.x, static_cast<uint32_t>(A.n_ref)));
    } else if (A.n_ref <= 1) {
        ref_idx = 0;
    }
}

// observation noise (env.py:139-152): radians(N(0, 0.5)) on wz, Box-Muller on a Philox draw
template <typename R>
__device__ __forceinline__ R obs_noise(const EnvArgs<R> &A, int64_t i, uint32_t episode, int k)
{
    const uint2 key = make_uint2(static_cast<uint32_t>(A.seed), static_cast<uint32_t>(A.seed >> 32));
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
template <typename R, bool FAST, bool AUTORESET>
__global__ void __launch_bounds__(kCtaThreads) env_step_kernel(const EnvArgs<R> A, const StepArgs<R> S)
{
    __shared__ Tables<R> tbl;
    __shared__ alignas(8) uint64_t bar;
    tables_issue(&tbl, A.tables, &bar);
    bool tables_ready = false;

    const int64_t n = A.n;
    const int64_t n_tiles = (S.i_end - S.i_begin + kCtaThreads - 1) / kCtaThreads;
    const R dt = A.dt;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = S.i_begin + tile * kCtaThreads + threadIdx.x;
        const bool live = i < S.i_end;
        R x[9];
        int ep_len = 0, ref_idx = 0;
        uint32_t done_sticky = 0, episode = 0;
        R ret = R(0);
        if (live) {
#pragma unroll
            for (int j = 0; j < 9; ++j) x[j] = A.state[j * n + i];
            ep_len = A.ep_len[i];
            ref_idx = A.ref_idx[i];
            done_sticky = AUTORESET ? 0u : A.done[i];
            if (A.total_return) ret = A.total_return[i];
            if (AUTORESET || A.noise) episode = A.episode_no[i];
        }
        if (!tables_ready) {   // first tile only: the state loads above overlapped the TMA copy
            tables_wait(&bar);
            tables_ready = true;
        }
        if (!live) continue;

        R o[5];
        bool have_obs = false;
        for (int k = 0; k < S.K; ++k) {
            R reward = R(0);
            uint32_t done_now = done_sticky;
            if (!done_sticky) {
                // F16.rescale_action (env.py:200-211): clip to [-1,1], float32-cast bounds
                const R a = clampr(S.actions[(int64_t)k * n + i], R(-1), R(1));
                const R u_stab = R(K::act_lo) + R(0.5) * (a + R(1.0)) * R(K::act_span);
                euler_step<R, FAST>(tbl, x, u_stab, R(0), dt);   // Control(action, 0), env.py:49
                // check_state (env.py:103-117) on the post-step state, pre-increment counter
                R pen = R(0);
                if (!(x[1] > R(300))) { pen += R(-1000); done_now = 1; }
                if (x[1] >= R(30000)) { pen += R(-1000); done_now = 1; }
                if (Mth<R, FAST>::fabs(x[2]) >= R(K::wz_term)) { pen += R(-1000); done_now = 1; }
                if (ep_len == A.horizon) done_now = 1;
                const R ref = A.ref_bank[(int64_t)ref_idx * A.ref_len + min(ep_len, A.ref_len - 1)];
                reward = pen + tracking_reward<R, FAST>(ref, x[2]);
                R wz_obs = x[2];
                if (A.noise) wz_obs += obs_noise<R>(A, i, episode, ep_len);
                make_obs<R, FAST>(x[1], wz_obs, x[4], ref, o);
                have_obs = true;
                ep_len += 1;
                ret += reward;
                if (done_now) {
                    if (S.final_obs) store_obs<R>(S.final_obs, A.obs_layout, n, i, o);
                    if (S.final_ep_len) S.final_ep_len[i] = ep_len;
                    if (S.final_return) S.final_return[i] = ret;
                    if (AUTORESET) {
                        // gymnasium SyncVectorEnv: reset immediately, hand back the reset observation
                        episode += 1;
                        init_episode<R>(A, i, episode, false, x, ref_idx);
                        ep_len = 0;
                        ret = R(0);
                        const R ref0 = A.ref_bank[(int64_t)ref_idx * A.ref_len];
                        make_obs<R, FAST>(x[1], x[2], x[4], ref0, o);
                    } else {
                        done_sticky = 1;
                    }
                }
            }
            S.reward[(int64_t)k * n + i] = reward;
            S.done[(int64_t)k * n + i] = static_cast<uint8_t>(done_now);
            if (S.obs_every_step) {
                if (!have_obs) {   // env was already done when the launch started
                    const R ref = A.ref_bank[(int64_t)ref_idx * A.ref_len + min(max(ep_len - 1, 0), A.ref_len - 1)];
                    make_obs<R, FAST>(x[1], x[2], x[4], ref, o);
                    have_obs = true;
                }
                store_obs<R>(S.obs + (int64_t)k * n * 5, A.obs_layout, n, i, o);
            }
        }
        // write back
#pragma unroll
        for (int j = 0; j < 9; ++j)
            if (j != 4 || AUTORESET) A.state[j * n + i] = x[j];
        A.ep_len[i] = ep_len;
        if (A.total_return) A.total_return[i] = ret;
        if (AUTORESET) {
            A.ref_idx[i] = ref_idx;
            A.episode_no[i] = episode;
        } else {
            A.done[i] = static_cast<uint8_t>(done_sticky);
        }
        if (!S.obs_every_step && S.obs) {
            if (!have_obs) {
                const R ref = A.ref_bank[(int64_t)ref_idx * A.ref_len + min(max(ep_len - 1, 0), A.ref_len - 1)];
                make_obs<R, FAST>(x[1], x[2], x[4], ref, o);
            }
            store_obs<R>(S.obs, A.obs_layout, n, i, o);
        }
    }
    if (!tables_ready) tables_wait(&bar);   // CTA had no tile: still drain the copy before exit
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
        init_episode<R>(A, i, episode, keep_ref != 0, x, ref_idx);
#pragma unroll
        for (int j = 0; j < 9; ++j) A.state[j * A.n + i] = x[j];
        A.ep_len[i] = 0;
        if (A.total_return) A.total_return[i] = R(0);
        A.ref_idx[i] = ref_idx;
        A.done[i] = 0;
        A.episode_no[i] = episode;
        if (obs) {
            R o[5];
            make_obs<R, false>(x[1], x[2], x[4], A.ref_bank[(int64_t)ref_idx * A.ref_len], o);
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
    if (A.autoreset) env_step_kernel<R, FAST, true><<<grid, kCtaThreads, 0, st>>>(A, S);
    else env_step_kernel<R, FAST, false><<<grid, kCtaThreads, 0, st>>>(A, S);
    return cudaGetLastError();
}

}  // namespace f16
