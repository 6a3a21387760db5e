// f16_capi.cu -- the C ABI (include/f16_b200.h): argument checking, per-device table images,
// launch geometry, and the host-buffer (end-to-end) step pipeline.  No torch types, no
// exceptions across the boundary.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>

#include "../../include/f16_b200.h"
#include "f16_launch.h"
#include "f16_policy.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) return fail(F16_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

constexpr int kMaxDev = 64;

// Makes `device` current for the scope and restores the caller's device on every exit path.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

struct HostPipe {   // scratch of f16_env_step_host, one per device; `mu` serialises the host threads that share it
    std::mutex mu;
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t done_ev[3] = {nullptr, nullptr, nullptr};
    void *actions = nullptr, *obs = nullptr, *reward = nullptr;
    uint8_t *done = nullptr;
    size_t cap_actions = 0, cap_obs = 0, cap_reward = 0, cap_done = 0;
};

struct DeviceCtx {
    bool ready = false;
    int n_sm = 0;
    f16::Tables<float> *t32 = nullptr;
    f16::Tables<double> *t64 = nullptr;
    int occ[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};   // [dtype][fast][autoreset]
    int occ2[2] = {0, 0};                                     // two-envs-per-thread kernel, [fast]
    HostPipe pipe;
};

DeviceCtx g_ctx[kMaxDev];
std::mutex g_mu;

int ensure_device(int device, DeviceCtx **out)
{
    if (device < 0 || device >= kMaxDev) return fail(F16_EINVAL, "device %d out of range", device);
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceCtx &c = g_ctx[device];
    if (!c.ready) {
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            cudaGetLastError();
            return fail(F16_ENODEV, "no CUDA device visible: the F-16 simulator has no CPU fallback");
        }
        if (device >= count) return fail(F16_ENODEV, "device %d not present (%d visible)", device, count);
        DeviceGuard guard(device);   // the caller's current device is restored on every return below
        CUDA_TRY(guard.err);
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            return fail(F16_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                        prop.minor);
        c.n_sm = prop.multiProcessorCount;
        std::unique_ptr<f16::Tables<float>> h32(new (std::nothrow) f16::Tables<float>);
        std::unique_ptr<f16::Tables<double>> h64(new (std::nothrow) f16::Tables<double>);
        if (!h32 || !h64) return fail(F16_EINVAL, "out of host memory");
        f16::build_tables(*h32);
        f16::build_tables(*h64);
        CUDA_TRY(cudaMalloc(&c.t32, sizeof *h32));
        CUDA_TRY(cudaMalloc(&c.t64, sizeof *h64));
        CUDA_TRY(cudaMemcpy(c.t32, h32.get(), sizeof *h32, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(c.t64, h64.get(), sizeof *h64, cudaMemcpyHostToDevice));
        for (int f = 0; f < 2; ++f)
            for (int a = 0; a < 2; ++a) {
                c.occ[F16_F32][f][a] = f16::env_step_occupancy_f32(f != 0, a != 0);
                c.occ[F16_F64][f][a] = f16::env_step_occupancy_f64(a != 0);
            }
        c.occ2[0] = f16::env_step2_occupancy_f32(false);
        c.occ2[1] = f16::env_step2_occupancy_f32(true);
        CUDA_TRY(cudaGetLastError());
        c.ready = true;
    }
    *out = &c;
    return F16_OK;
}

int current_device(int *dev)
{
    cudaError_t e = cudaGetDevice(dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(F16_ENODEV, "cudaGetDevice: %s (the F-16 simulator has no CPU fallback)", cudaGetErrorString(e));
    }
    return F16_OK;
}

// flags = f16_config.flags: kernel selection is a per-call property of the config, not process state
f16::LaunchGeom geom(const DeviceCtx &c, int dtype, bool fast, bool autoreset, int flags)
{
    int occ = c.occ[dtype][fast ? 1 : 0][autoreset ? 1 : 0];
    if (occ < 1) occ = 1;
    f16::LaunchGeom g{c.n_sm, occ, (flags & F16_FLAG_NO_PDL) ? 0 : 1};
    g.two_per_thread = (dtype == F16_F32 && !(flags & F16_FLAG_ONE_ENV_PER_THREAD)) ? 1 : 0;
    const int o2 = c.occ2[fast ? 1 : 0];
    g.ctas2_per_sm = (o2 & 0xff) > 0 ? (o2 & 0xff) : 1;
    g.ctas2k_per_sm = ((o2 >> 8) & 0xff) > 0 ? ((o2 >> 8) & 0xff) : 1;
    return g;
}

int check_cfg(const f16_config *cfg)
{
    if (!cfg) return fail(F16_EINVAL, "cfg is NULL");
    if (cfg->dtype != F16_F32 && cfg->dtype != F16_F64) return fail(F16_EINVAL, "dtype must be F16_F32 or F16_F64");
    if (cfg->n_envs <= 0) return fail(F16_EINVAL, "n_envs must be positive");
    if (!(cfg->dt > 0)) return fail(F16_EINVAL, "dt must be positive");
    if (cfg->ref_len <= 0 || cfg->n_ref <= 0) return fail(F16_EINVAL, "reference bank must hold >= 1 signal of >= 1 sample");
    if (cfg->init_mode < 0 || cfg->init_mode > 2) return fail(F16_EINVAL, "bad init_mode");
    if (cfg->obs_layout != F16_OBS_AOS && cfg->obs_layout != F16_OBS_SOA) return fail(F16_EINVAL, "bad obs_layout");
    if (cfg->seg_len < 0 || (cfg->seg_len > 0 && (cfg->seg_stride < 1 || cfg->seg_stride > 256 || cfg->ref_len > 4 * cfg->seg_len)))
        return fail(F16_EINVAL, "segment mode needs 1 <= seg_stride <= 256 and ref_len <= 4 * seg_len");
    if (cfg->flags & ~(F16_FLAG_ONE_ENV_PER_THREAD | F16_FLAG_NO_PDL | F16_FLAG_HOST_STAGED)) return fail(F16_EINVAL, "unknown bits in flags");
    return F16_OK;
}

int check_env(const f16_config *cfg, const f16_env_buffers *env)
{
    if (!env) return fail(F16_EINVAL, "env buffers are NULL");
    if (!env->state || !env->ep_len || !env->total_return || !env->ref_idx || !env->done || !env->episode_no || !env->ref_bank)
        return fail(F16_EINVAL, "state, ep_len, total_return, ref_idx, done, episode_no and ref_bank are required");
    if (env->pos_comp && cfg->dtype != F16_F32) return fail(F16_EINVAL, "pos_comp is an fp32-mode feature; leave it NULL in fp64");
    if ((cfg->seg_len > 0) != (env->ref_codes != nullptr)) return fail(F16_EINVAL, "ref_codes goes with segment mode (seg_len > 0), and only with it");
    if (cfg->init_mode == F16_INIT_PER_ENV && !env->init_state)
        return fail(F16_EINVAL, "init_mode F16_INIT_PER_ENV needs env->init_state");
    return F16_OK;
}

template <typename R>
f16::EnvArgs<R> make_env_args(const f16_config *cfg, const f16_env_buffers *env, const f16::Tables<R> *tables)
{
    f16::EnvArgs<R> A;
    A.tables = tables;
    A.n = cfg->n_envs;
    A.dt = static_cast<R>(cfg->dt);
    A.horizon = cfg->horizon;
    A.ref_len = cfg->ref_len;
    A.n_ref = cfg->n_ref;
    A.seg_len = cfg->seg_len;
    A.seg_stride = cfg->seg_stride;
    A.ref_codes = env->ref_codes;
    A.autoreset = cfg->autoreset;
    A.init_mode = cfg->init_mode;
    A.obs_layout = cfg->obs_layout;
    A.noise = cfg->noise;
    A.seed = cfg->seed;
    A.seed_dev = env->seed_dev;
    A.id_base = cfg->env_id_base;
    for (int j = 0; j < 9; ++j) A.x0[j] = static_cast<R>(cfg->x0[j]);
    A.state = static_cast<R *>(env->state);
    A.ep_len = env->ep_len;
    A.total_return = static_cast<R *>(env->total_return);
    A.ref_idx = env->ref_idx;
    A.done = env->done;
    A.episode_no = env->episode_no;
    A.ref_bank = static_cast<const R *>(env->ref_bank);
    A.init_state = static_cast<const R *>(env->init_state);
    A.pos_comp = static_cast<R *>(env->pos_comp);
    return A;
}

template <typename R>
f16::StepArgs<R> make_step_args(const f16_config *cfg, const f16_step_io *io, int K)
{
    f16::StepArgs<R> S;
    S.actions = static_cast<const R *>(io->actions);
    S.obs = static_cast<R *>(io->obs);
    S.reward = static_cast<R *>(io->reward);
    S.done = io->done;
    S.final_obs = static_cast<R *>(io->final_obs);
    S.final_ep_len = io->final_ep_len;
    S.final_return = static_cast<R *>(io->final_return);
    S.stats = io->stats;
    S.obs_every_step = io->obs_every_step;
    S.K = K;
    S.tma_ok = 0;   // set by step_range once the env buffers are known
    S.obs_bulk = 0;
    S.pdl_mode = 0;
    S.i_begin = 0;
    S.i_end = cfg->n_envs;
    return S;
}

int check_step(const f16_config *cfg, const f16_step_io *io, int K)
{
    if (!io) return fail(F16_EINVAL, "io is NULL");
    if (K <= 0) return fail(F16_EINVAL, "K must be >= 1");
    // the kernels index with 32 bits: every array must stay below 2^32 elements
    const double lim = 4294967296.0;
    const double n = static_cast<double>(cfg->n_envs);
    if (9.0 * n >= lim || static_cast<double>(K) * n * (io->obs_every_step ? 5.0 : 1.0) >= lim || 5.0 * n >= lim ||
        (cfg->seg_len > 0 ? 0.0 : static_cast<double>(cfg->n_ref) * cfg->ref_len) >= lim)
        return fail(F16_EINVAL, "batch too large for 32-bit indexing (n_envs * max(9, K, 5K with obs_every_step) must be < 2^32): "
                                "shard the envs or lower K");
    if (!io->actions || !io->reward || !io->done) return fail(F16_EINVAL, "actions, reward and done are required");
    if (io->obs_every_step && !io->obs) return fail(F16_EINVAL, "obs_every_step needs obs");
    return F16_OK;
}

// TMA bulk copies need 16-byte aligned global addresses and sizes: every SoA row of every full tile
// starts at base + (j * n + i_begin + 256 * t) * elem, so bases, n and i_begin must be aligned.
int tma_aligned(const f16_config *cfg, const f16_env_buffers *env, const f16_step_io *io, int64_t i0)
{
    const size_t es = cfg->dtype == F16_F32 ? 4 : 8;
    auto ok = [](const void *p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    if (!ok(env->state) || !ok(env->ep_len) || !ok(env->ref_idx) || !ok(env->done) || !ok(io->actions)) return 0;
    if (!ok(env->total_return)) return 0;
    if (env->pos_comp && !ok(env->pos_comp)) return 0;
    if ((static_cast<size_t>(cfg->n_envs) * es) % 16 != 0) return 0;
    if (i0 % 16 != 0) return 0;
    return 1;
}

// bulk_obs: stage the [N,5] observations in shared memory and write them with bulk stores.  Worth it only when obs
// lives in mapped host memory (full-line PCIe writes); on device memory plain stores are 1.5-4 % faster (measured).
int step_range(DeviceCtx *ctx, const f16_config *cfg, const f16_env_buffers *env, const f16_step_io *io, int K, int64_t i0, int64_t i1,
               cudaStream_t st, bool bulk_obs = false)
{
    const bool fast = cfg->dtype == F16_F32 && cfg->fast_math != 0;
    const f16::LaunchGeom g = geom(*ctx, cfg->dtype, fast, cfg->autoreset != 0, cfg->flags);
    if (cfg->dtype == F16_F32) {
        auto A = make_env_args<float>(cfg, env, ctx->t32);
        auto S = make_step_args<float>(cfg, io, K);
        S.i_begin = i0;
        S.i_end = i1;
        S.tma_ok = tma_aligned(cfg, env, io, i0);
        static const bool force_bulk = [] {   // probing knob: bulk-stored observations on the device route too
            const char *e = std::getenv("F16B200_BULK_OBS");
            return e && e[0] == '1';
        }();
        S.obs_bulk = (bulk_obs || force_bulk) && S.tma_ok && io->obs && cfg->obs_layout == F16_OBS_AOS && reinterpret_cast<uintptr_t>(io->obs) % 16 == 0;
        CUDA_TRY(f16::launch_env_step(A, S, fast, g, st));
    } else {
        auto A = make_env_args<double>(cfg, env, ctx->t64);
        auto S = make_step_args<double>(cfg, io, K);
        S.i_begin = i0;
        S.i_end = i1;
        S.tma_ok = tma_aligned(cfg, env, io, i0);
        CUDA_TRY(f16::launch_env_step(A, S, false, g, st));
    }
    return F16_OK;
}

thread_local int t_host_route = F16_HOST_ROUTE_STAGED;

// pinned + mapped host allocation?  -> the device-side alias of the pointer
bool mapped_host(const void *p, void **dev)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev = a.devicePointer;
    return true;
}

int grow(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return F16_OK;
    if (*p) CUDA_TRY(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CUDA_TRY(cudaMalloc(p, need));
    *cap = need;
    return F16_OK;
}

}  // namespace

extern "C" {

int f16_abi_version(void) { return F16_ABI_VERSION; }

const char *f16_last_error(void) { return g_err; }

int f16_init(int device)
{
    DeviceCtx *c;
    return ensure_device(device, &c);
}

int f16_launch_geometry(int device, int dtype, int *n_sm, int *cta_threads, int *ctas_per_sm)
{
    DeviceCtx *c;
    int rc = ensure_device(device, &c);
    if (rc) return rc;
    if (dtype != F16_F32 && dtype != F16_F64) return fail(F16_EINVAL, "bad dtype");
    if (n_sm) *n_sm = c->n_sm;
    if (cta_threads) *cta_threads = 256;
    if (ctas_per_sm) *ctas_per_sm = c->occ[dtype][dtype == F16_F32 ? 1 : 0][0];
    return F16_OK;
}

int f16_peak_probe(int kind, int iters, double *ops_per_launch, void *scratch, void *stream)
{
    if (kind < F16_PEAK_FP32_FMA || kind > F16_PEAK_MUFU_TANH || iters <= 0 || !scratch) return fail(F16_EINVAL, "bad kind / iters / scratch");
    int dev, rc;
    if ((rc = current_device(&dev))) return rc;
    DeviceCtx *ctx;
    if ((rc = ensure_device(dev, &ctx))) return rc;
    CUDA_TRY(f16::launch_peak_probe(kind, iters, ctx->n_sm, static_cast<float *>(scratch), ops_per_launch, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_env_reset(const f16_config *cfg, const f16_env_buffers *env, const uint8_t *mask, int keep_ref, void *obs, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if ((rc = check_env(cfg, env))) return rc;
    int dev;
    if ((rc = current_device(&dev))) return rc;
    DeviceCtx *ctx;
    if ((rc = ensure_device(dev, &ctx))) return rc;
    const f16::LaunchGeom g{ctx->n_sm, 8};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cfg->dtype == F16_F32) {
        auto A = make_env_args<float>(cfg, env, ctx->t32);
        CUDA_TRY(f16::launch_env_reset(A, mask, keep_ref, static_cast<float *>(obs), g, st));
    } else {
        auto A = make_env_args<double>(cfg, env, ctx->t64);
        CUDA_TRY(f16::launch_env_reset(A, mask, keep_ref, static_cast<double *>(obs), g, st));
    }
    return F16_OK;
}

int f16_env_step(const f16_config *cfg, const f16_env_buffers *env, const f16_step_io *io, int K, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if ((rc = check_env(cfg, env))) return rc;
    if ((rc = check_step(cfg, io, K))) return rc;
    int dev;
    if ((rc = current_device(&dev))) return rc;
    DeviceCtx *ctx;
    if ((rc = ensure_device(dev, &ctx))) return rc;
    return step_range(ctx, cfg, env, io, K, 0, cfg->n_envs, static_cast<cudaStream_t>(stream));
}

int f16_env_step_host(const f16_config *cfg, const f16_env_buffers *env, const void *actions_host, void *obs_host, void *reward_host,
                      uint8_t *done_host, int K, int device)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if ((rc = check_env(cfg, env))) return rc;
    if (K <= 0) return fail(F16_EINVAL, "K must be >= 1");
    if (!actions_host || !reward_host || !done_host) return fail(F16_EINVAL, "actions_host, reward_host and done_host are required");
    {   // the same 32-bit index limits as the device route (check_step)
        f16_step_io probe;
        std::memset(&probe, 0, sizeof probe);
        probe.actions = actions_host;
        probe.reward = reward_host;
        probe.done = done_host;
        if ((rc = check_step(cfg, &probe, K))) return rc;
    }
    DeviceCtx *ctx;
    if ((rc = ensure_device(device, &ctx))) return rc;
    DeviceGuard guard(device);   // restores the caller's current device when the call returns
    CUDA_TRY(guard.err);
    HostPipe &P = ctx->pipe;
    std::lock_guard<std::mutex> pipe_lock(P.mu);   // staging buffers and streams are per device: one host step at a time
    const int64_t n = cfg->n_envs;
    const size_t es = cfg->dtype == F16_F32 ? 4 : 8;
    if (!P.streams[0])
        for (int s = 0; s < 3; ++s) {
            CUDA_TRY(cudaStreamCreateWithFlags(&P.streams[s], cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&P.done_ev[s], cudaEventDisableTiming));
        }
    f16_step_io io;
    std::memset(&io, 0, sizeof io);

    // Zero-copy route: when every host buffer is pinned AND mapped (cudaHostAlloc / cudaHostRegister, e.g. torch's
    // pin_memory()), the step kernel reads the actions from and writes obs / reward / done to host memory itself --
    // one launch, no staging copies, transfers overlapped with the arithmetic tile by tile.  It needs the kernel
    // whose stores are full lines (two-envs-per-thread kernel: bulk-stored [N,5] observations, or the SoA layout);
    // 20-byte-strided scalar stores over PCIe are 2.5x slower than the staged pipeline (tools/probes/zero_copy_probe.py).
    {
        const bool zero_copy = !(cfg->flags & F16_FLAG_HOST_STAGED);
        void *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr, *d_done = nullptr;
        const f16::LaunchGeom g = geom(*ctx, cfg->dtype, cfg->fast_math != 0, cfg->autoreset != 0, cfg->flags);
        const bool wide_stores = cfg->dtype == F16_F32 && g.two_per_thread && cfg->autoreset && !cfg->noise && !env->pos_comp && n >= 512;
        if (zero_copy && K == 1 && wide_stores && mapped_host(actions_host, &d_act) && mapped_host(reward_host, &d_rew) &&
            mapped_host(done_host, &d_done) && (!obs_host || mapped_host(obs_host, &d_obs))) {
            io.actions = d_act;
            io.obs = d_obs;
            io.reward = d_rew;
            io.done = static_cast<uint8_t *>(d_done);
            if ((rc = step_range(ctx, cfg, env, &io, K, 0, n, P.streams[0], true))) return rc;
            CUDA_TRY(cudaStreamSynchronize(P.streams[0]));
            t_host_route = F16_HOST_ROUTE_ZERO_COPY;
            return F16_OK;
        }
    }

    // Staged route (pageable or unmapped host memory, K > 1, other kernels):
    t_host_route = F16_HOST_ROUTE_STAGED;
    if ((rc = grow(&P.actions, &P.cap_actions, es * K * n))) return rc;
    if ((rc = grow(&P.obs, &P.cap_obs, es * 5 * n))) return rc;
    if ((rc = grow(&P.reward, &P.cap_reward, es * K * n))) return rc;
    if ((rc = grow(reinterpret_cast<void **>(&P.done), &P.cap_done, (size_t)K * n))) return rc;
    io.actions = P.actions;
    io.obs = obs_host ? P.obs : nullptr;
    io.reward = P.reward;
    io.done = P.done;
    // chunk the env range so that H2D(c+1), kernel(c) and D2H(c-1) overlap on three streams; the transfer is
    // D2H-bound (25 B per env-step out), so few large copies beat many small ones
    int64_t chunks = n >= (1 << 17) ? 3 : 1;   // measured on B200 / PCIe Gen5 (tools/e2e_probe.py): 3 chunks 1.78e9, 1: 1.72e9, 8: 1.60e9 env-steps/s
    if (const char *e = std::getenv("F16B200_HOST_CHUNKS")) {   // tuning knob
        const long v = std::atol(e);
        if (v >= 1 && v <= 64) chunks = v;
    }
    // chunk c holds a share proportional to ratio^c: a small first chunk gets the device->host engine going early
    double ratio = 1.0;
    if (const char *e = std::getenv("F16B200_HOST_RATIO")) {
        const double v = std::atof(e);
        if (v >= 1.0 && v <= 16.0) ratio = v;
    }
    int64_t bounds[65];
    {
        double total = 0.0, w = 1.0;
        for (int64_t c = 0; c < chunks; ++c, w *= ratio) total += w;
        double acc = 0.0;
        w = 1.0;
        bounds[0] = 0;
        for (int64_t c = 0; c < chunks; ++c, w *= ratio) {
            acc += w;
            int64_t b = static_cast<int64_t>(static_cast<double>(n) * (acc / total));
            b = (b + 255) / 256 * 256;
            if (b > n || c == chunks - 1) b = n;
            if (b < bounds[c]) b = bounds[c];
            bounds[c + 1] = b;
        }
    }
    for (int64_t c = 0; c < chunks; ++c) {
        const int64_t i0 = bounds[c], i1 = bounds[c + 1];
        if (i1 <= i0) continue;
        const int64_t cnt = i1 - i0;
        cudaStream_t st = P.streams[c % 3];
        for (int k = 0; k < K; ++k)
            CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(P.actions) + es * ((size_t)k * n + i0),
                                     static_cast<const char *>(actions_host) + es * ((size_t)k * n + i0), es * cnt,
                                     cudaMemcpyHostToDevice, st));
        if ((rc = step_range(ctx, cfg, env, &io, K, i0, i1, st))) return rc;
        for (int k = 0; k < K; ++k) {
            CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(reward_host) + es * ((size_t)k * n + i0),
                                     static_cast<const char *>(P.reward) + es * ((size_t)k * n + i0), es * cnt,
                                     cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(done_host + (size_t)k * n + i0, P.done + (size_t)k * n + i0, (size_t)cnt, cudaMemcpyDeviceToHost,
                                     st));
        }
        if (obs_host) {
            if (cfg->obs_layout == F16_OBS_AOS) {
                CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(obs_host) + es * 5 * (size_t)i0,
                                         static_cast<const char *>(P.obs) + es * 5 * (size_t)i0, es * 5 * cnt, cudaMemcpyDeviceToHost,
                                         st));
            } else {
                for (int j = 0; j < 5; ++j)
                    CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(obs_host) + es * ((size_t)j * n + i0),
                                             static_cast<const char *>(P.obs) + es * ((size_t)j * n + i0), es * cnt,
                                             cudaMemcpyDeviceToHost, st));
            }
        }
    }
    for (int s = 0; s < 3; ++s) CUDA_TRY(cudaStreamSynchronize(P.streams[s]));
    return F16_OK;
}

int f16_host_route(void) { return t_host_route; }

#define DISPATCH_SIMPLE(call32, call64)                 \
    do {                                                \
        if (dtype == F16_F32) CUDA_TRY(call32);         \
        else if (dtype == F16_F64) CUDA_TRY(call64);    \
        else return fail(F16_EINVAL, "bad dtype");      \
    } while (0)

static int prep(int64_t n, DeviceCtx **ctx, f16::LaunchGeom *g)
{
    if (n <= 0) return fail(F16_EINVAL, "n must be positive");
    int dev, rc;
    if ((rc = current_device(&dev))) return rc;
    if ((rc = ensure_device(dev, ctx))) return rc;
    *g = f16::LaunchGeom{(*ctx)->n_sm, 4};
    return F16_OK;
}

int f16_rhs(int dtype, int64_t n, const void *x, const void *u, void *dx, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!x || !u || !dx) return fail(F16_EINVAL, "NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DISPATCH_SIMPLE(f16::launch_rhs(c->t32, n, (const float *)x, (const float *)u, (float *)dx, g, st),
                    f16::launch_rhs(c->t64, n, (const double *)x, (const double *)u, (double *)dx, g, st));
    return F16_OK;
}

int f16_model_step(int dtype, int64_t n, void *x, const void *u, double dt, int K, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!x || !u || K <= 0) return fail(F16_EINVAL, "NULL pointer or K <= 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DISPATCH_SIMPLE(f16::launch_model_step(c->t32, n, (float *)x, (const float *)u, (float)dt, K, g, st),
                    f16::launch_model_step(c->t64, n, (double *)x, (const double *)u, dt, K, g, st));
    return F16_OK;
}

int f16_coeffs(int dtype, int64_t n, const void *alpha, const void *fi, const void *wz, const void *V, void *out, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!alpha || !fi || !wz || !V || !out) return fail(F16_EINVAL, "NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DISPATCH_SIMPLE(
        f16::launch_coeffs(c->t32, n, (const float *)alpha, (const float *)fi, (const float *)wz, (const float *)V, (float *)out, g, st),
        f16::launch_coeffs(c->t64, n, (const double *)alpha, (const double *)fi, (const double *)wz, (const double *)V, (double *)out, g, st));
    return F16_OK;
}

int f16_thrust(int dtype, int64_t n, const void *H, const void *M, const void *Pa, void *out, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!H || !M || !Pa || !out) return fail(F16_EINVAL, "NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DISPATCH_SIMPLE(f16::launch_thrust(c->t32, n, (const float *)H, (const float *)M, (const float *)Pa, (float *)out, g, st),
                    f16::launch_thrust(c->t64, n, (const double *)H, (const double *)M, (const double *)Pa, (double *)out, g, st));
    return F16_OK;
}

int f16_atmosphere(int dtype, int64_t n, const void *h, void *rho, void *a, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!h || !rho || !a) return fail(F16_EINVAL, "NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DISPATCH_SIMPLE(f16::launch_isa(n, (const float *)h, (float *)rho, (float *)a, g, st),
                    f16::launch_isa(n, (const double *)h, (double *)rho, (double *)a, g, st));
    return F16_OK;
}

int f16_trim(int n_cells, const void *V, const void *H, int n_starts, const void *starts, void *out_x, void *out_cost, void *out_iters,
             void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n_cells, &c, &g);
    if (rc) return rc;
    if (n_starts <= 0 || !V || !H || !starts || !out_x || !out_cost) return fail(F16_EINVAL, "NULL pointer or n_starts <= 0");
    CUDA_TRY(f16::launch_trim(c->t64, n_cells, n_starts, (const double *)V, (const double *)H, (const double *)starts, (double *)out_x,
                              (double *)out_cost, (int *)out_iters, g, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_policy_blob_bytes(void) { return static_cast<int>(sizeof(f16::PolicyBlob)); }

int f16_policy_forward(int64_t n, const void *obs, const void *blob, const void *logstd, void *mean, void *value, void *action,
                       void *logprob, int sample, uint64_t seed, uint32_t step, const int32_t *ep_len, const uint32_t *episode_no,
                       int64_t env_id_base, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!obs || !blob) return fail(F16_EINVAL, "obs and blob are required");
    if (reinterpret_cast<uintptr_t>(blob) % 16 != 0) return fail(F16_EINVAL, "blob must be 16-byte aligned");
    if (action && !logstd) return fail(F16_EINVAL, "sampling an action needs logstd");
    f16::PolicyArgs P;
    P.blob = static_cast<const f16::PolicyBlob *>(blob);
    P.obs = static_cast<const float *>(obs);
    P.logstd = static_cast<const float *>(logstd);
    P.mean = static_cast<float *>(mean);
    P.value = static_cast<float *>(value);
    P.action = static_cast<float *>(action);
    P.logprob = static_cast<float *>(logprob);
    P.n = n;
    P.sample = sample;
    P.step = step;
    P.ep_len = ep_len;
    P.episode_no = episode_no;
    P.seed = seed;
    P.id_base = env_id_base;
    CUDA_TRY(f16::launch_policy_forward(P, c->n_sm, geom(*c, F16_F32, true, true, 0).pdl, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_policy_forward16(int64_t n, const void *obs, const void *train_blob16, const void *logstd, void *mean, void *value, void *action,
                         void *logprob, int sample, uint64_t seed, uint32_t step, const int32_t *ep_len, const uint32_t *episode_no,
                         int64_t env_id_base, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!obs || !train_blob16) return fail(F16_EINVAL, "obs and train_blob16 are required");
    if (reinterpret_cast<uintptr_t>(train_blob16) % 16 != 0) return fail(F16_EINVAL, "train_blob16 must be 16-byte aligned");
    if (action && !logstd) return fail(F16_EINVAL, "sampling an action needs logstd");
    if (static_cast<double>(n) * 5.0 >= 4294967296.0) return fail(F16_EINVAL, "batch too large for 32-bit indexing");
    f16::PolicyArgs P;
    P.blob = nullptr;
    P.obs = static_cast<const float *>(obs);
    P.logstd = static_cast<const float *>(logstd);
    P.mean = static_cast<float *>(mean);
    P.value = static_cast<float *>(value);
    P.action = static_cast<float *>(action);
    P.logprob = static_cast<float *>(logprob);
    P.n = n;
    P.sample = sample;
    P.step = step;
    P.ep_len = ep_len;
    P.episode_no = episode_no;
    P.seed = seed;
    P.id_base = env_id_base;
    CUDA_TRY(f16::launch_policy_forward16(P, static_cast<const f16::TrainBlob16 *>(train_blob16), c->n_sm, geom(*c, F16_F32, true, true, 0).pdl,
                                          static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_rollout(const f16_config *cfg, const f16_env_buffers *env, const f16_rollout_io *io, int K, void *stream)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if ((rc = check_env(cfg, env))) return rc;
    if (!io) return fail(F16_EINVAL, "io is NULL");
    if (K <= 0) return fail(F16_EINVAL, "K must be >= 1");
    if (cfg->dtype != F16_F32 || !cfg->autoreset || cfg->obs_layout != F16_OBS_AOS || cfg->noise || env->pos_comp)
        return fail(F16_EINVAL, "f16_rollout runs the fp32 autoreset env with [n][5] observations, without observation noise or position carries");
    if (!io->train_blob16 || !io->logstd || !io->obs0 || !io->obs_next || !io->actions || !io->logprobs || !io->values || !io->rewards || !io->dones)
        return fail(F16_EINVAL, "train_blob16, logstd, obs0, obs_next, actions, logprobs, values, rewards and dones are required");
    if (reinterpret_cast<uintptr_t>(io->train_blob16) % 16 != 0) return fail(F16_EINVAL, "train_blob16 must be 16-byte aligned");
    if (5.0 * static_cast<double>(K) * static_cast<double>(cfg->n_envs) >= 4294967296.0 || 9.0 * static_cast<double>(cfg->n_envs) >= 4294967296.0)
        return fail(F16_EINVAL, "batch too large for 32-bit indexing (5 * K * n_envs must be < 2^32): shard the envs or lower K");
    int dev;
    if ((rc = current_device(&dev))) return rc;
    DeviceCtx *ctx;
    if ((rc = ensure_device(dev, &ctx))) return rc;
    auto A = make_env_args<float>(cfg, env, ctx->t32);
    f16::RolloutArgs R;
    R.blob = static_cast<const f16::TrainBlob16 *>(io->train_blob16);
    R.logstd = static_cast<const float *>(io->logstd);
    R.obs0 = static_cast<const float *>(io->obs0);
    R.obs_next = static_cast<float *>(io->obs_next);
    R.actions = static_cast<float *>(io->actions);
    R.logprobs = static_cast<float *>(io->logprobs);
    R.values = static_cast<float *>(io->values);
    R.rewards = static_cast<float *>(io->rewards);
    R.dones = io->dones;
    R.next_value = static_cast<float *>(io->next_value);
    R.final_obs = static_cast<float *>(io->final_obs);
    R.final_ep_len = io->final_ep_len;
    R.final_return = static_cast<float *>(io->final_return);
    R.stats = io->stats;
    R.K = K;
    R.sample = io->sample;
    R.policy_seed = io->policy_seed;
    R.pace = 3;   // groups of a CTA inside their tanh section at a time (measured: 0 / 2 / 3 / 4 -> 6.71 / 6.77 / 5.98 / 6.90 us per step at 65,536 envs)
    CUDA_TRY(f16::launch_rollout(A, R, cfg->fast_math != 0, ctx->n_sm, (cfg->flags & F16_FLAG_NO_PDL) ? 0 : 1, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_ppo_adv_stats(int64_t n, int n_minibatches, const void *adv, const int64_t *mb_idx, void *out, void *scratch, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!adv || !mb_idx || !out || !scratch) return fail(F16_EINVAL, "NULL pointer");
    if (n_minibatches < 1 || n_minibatches > 65535) return fail(F16_EINVAL, "n_minibatches must be in 1..65535");
    CUDA_TRY(f16::launch_adv_stats(static_cast<const float *>(adv), mb_idx, n, n_minibatches, static_cast<float *>(out), static_cast<double *>(scratch),
                                   c->n_sm, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_ppo_train_blob_bytes(void) { return static_cast<int>(sizeof(f16::TrainBlob)); }
int f16_ppo_grad_floats(void) { return static_cast<int>(sizeof(f16::PolicyGrad) / sizeof(float)); }

static int ppo_grad_impl(bool fp16, void *partials, const f16_peer_group *peer, int64_t n, const void *train_blob, const void *obs, const void *actions, const void *logprobs, const void *adv,
                 const void *ret, const void *val, const int64_t *mb_idx, const void *logstd, const void *adv_stats, void *grad,
                 double clip_coef, double vf_coef, double ent_coef, int norm_adv, int clip_vloss, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    const bool packed = fp16 && !actions;   // rows float[N][8] in `obs`, (ret, val) float[N][2] in `ret` (f16_ppo_pack_rows)
    if (!train_blob || !obs || !ret || !mb_idx || !logstd || !grad || (!packed && (!actions || !logprobs || !adv || !val)))
        return fail(F16_EINVAL, "NULL pointer");
    if (packed && (reinterpret_cast<uintptr_t>(obs) % 32 != 0 || reinterpret_cast<uintptr_t>(ret) % 8 != 0))
        return fail(F16_EINVAL, "packed rows must be 32-byte aligned, (ret, val) pairs 8-byte aligned");
    if (norm_adv && !adv_stats) return fail(F16_EINVAL, "norm_adv needs adv_stats (mean, std of the minibatch advantages)");
    if (reinterpret_cast<uintptr_t>(train_blob) % 16 != 0) return fail(F16_EINVAL, "train_blob must be 16-byte aligned");
    f16::PpoGradArgs P;
    P.blob = static_cast<const f16::TrainBlob *>(train_blob);
    P.obs = static_cast<const float *>(obs);
    P.actions = static_cast<const float *>(actions);
    P.logprobs = static_cast<const float *>(logprobs);
    P.adv = static_cast<const float *>(adv);
    P.ret = static_cast<const float *>(ret);
    P.val = static_cast<const float *>(val);
    P.mb_idx = mb_idx;
    P.logstd = static_cast<const float *>(logstd);
    P.adv_stats = static_cast<const float *>(adv_stats);
    P.grad = static_cast<f16::PolicyGrad *>(grad);
    P.n = n;
    P.clip_coef = static_cast<float>(clip_coef);
    P.vf_coef = static_cast<float>(vf_coef);
    P.ent_coef = static_cast<float>(ent_coef);
    P.norm_adv = norm_adv;
    P.clip_vloss = clip_vloss;
    P.packed = packed ? 1 : 0;
    f16::PeerArgs pa;
    if (peer) {
        if (!fp16 || !partials) return fail(F16_EINVAL, "the peer all-reduce rides in the reduction kernel of f16_ppo_grad16: scratch is required");
        if (peer->world < 1 || peer->world > 32 || peer->rank < 0 || peer->rank >= peer->world || !peer->bufs || !peer->seq)
            return fail(F16_EINVAL, "bad f16_peer_group (1 <= world <= 32, 0 <= rank < world, bufs and seq set)");
        pa.bufs = reinterpret_cast<uint2 *const *>(peer->bufs);
        pa.seq = peer->seq;
        pa.rank = peer->rank;
        pa.world = peer->world;
    }
    if (fp16) CUDA_TRY(f16::launch_ppo_grad16(P, static_cast<const f16::TrainBlob16 *>(train_blob), static_cast<float *>(partials), c->n_sm, static_cast<cudaStream_t>(stream), peer ? &pa : nullptr));
    else CUDA_TRY(f16::launch_ppo_grad(P, c->n_sm, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_ppo_grad(int64_t n, const void *train_blob, const void *obs, const void *actions, const void *logprobs, const void *adv,
                 const void *ret, const void *val, const int64_t *mb_idx, const void *logstd, const void *adv_stats, void *grad,
                 double clip_coef, double vf_coef, double ent_coef, int norm_adv, int clip_vloss, void *stream)
{
    return ppo_grad_impl(false, nullptr, nullptr, n, train_blob, obs, actions, logprobs, adv, ret, val, mb_idx, logstd, adv_stats, grad, clip_coef, vf_coef,
                         ent_coef, norm_adv, clip_vloss, stream);
}

int f16_ppo_randperm(int64_t n, uint64_t seed, int64_t *out, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (!out) return fail(F16_EINVAL, "NULL pointer");
    CUDA_TRY(f16::launch_ppo_randperm(n, seed, out, c->n_sm, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_ppo_pack_rows(int64_t N, const void *obs, const void *actions, const void *logprobs, const void *adv, const void *ret, const void *val,
                      void *rows, void *rv, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(N, &c, &g);
    if (rc) return rc;
    if (!obs || !actions || !logprobs || !adv || !ret || !val || !rows || !rv) return fail(F16_EINVAL, "NULL pointer");
    if (reinterpret_cast<uintptr_t>(rows) % 32 != 0 || reinterpret_cast<uintptr_t>(rv) % 8 != 0)
        return fail(F16_EINVAL, "rows must be 32-byte aligned, rv 8-byte aligned");
    CUDA_TRY(f16::launch_ppo_pack_rows(N, static_cast<const float *>(obs), static_cast<const float *>(actions), static_cast<const float *>(logprobs),
                                       static_cast<const float *>(adv), static_cast<const float *>(ret), static_cast<const float *>(val),
                                       static_cast<float *>(rows), static_cast<float *>(rv), c->n_sm, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_ppo_train_blob16_bytes(void) { return static_cast<int>(sizeof(f16::TrainBlob16)); }

int f16_ppo_grad16_scratch_floats(int device)
{
    DeviceCtx *c;
    if (ensure_device(device, &c)) return -1;
    return c->n_sm * static_cast<int>(sizeof(f16::PolicyGrad) / sizeof(float));
}

int f16_ppo_grad16(int64_t n, const void *train_blob16, const void *obs, const void *actions, const void *logprobs, const void *adv,
                   const void *ret, const void *val, const int64_t *mb_idx, const void *logstd, const void *adv_stats, void *grad,
                   double clip_coef, double vf_coef, double ent_coef, int norm_adv, int clip_vloss, void *scratch, void *stream)
{
    return ppo_grad_impl(true, scratch, nullptr, n, train_blob16, obs, actions, logprobs, adv, ret, val, mb_idx, logstd, adv_stats, grad, clip_coef, vf_coef,
                         ent_coef, norm_adv, clip_vloss, stream);
}

int f16_ppo_grad16_peer(int64_t n, const void *train_blob16, const void *obs, const void *actions, const void *logprobs, const void *adv,
                        const void *ret, const void *val, const int64_t *mb_idx, const void *logstd, const void *adv_stats, void *grad,
                        double clip_coef, double vf_coef, double ent_coef, int norm_adv, int clip_vloss, void *scratch,
                        const f16_peer_group *peer, void *stream)
{
    if (!peer) return fail(F16_EINVAL, "peer is NULL");
    return ppo_grad_impl(true, scratch, peer, n, train_blob16, obs, actions, logprobs, adv, ret, val, mb_idx, logstd, adv_stats, grad, clip_coef, vf_coef,
                         ent_coef, norm_adv, clip_vloss, stream);
}

int f16_ppo_step16_sync_bytes(void) { return static_cast<int>((f16::ppo_step16_sync_bytes() + 15) / 16 * 16); }

static int ppo_step16_impl(const f16_ppo_step16_args *a, int n_minibatches, bool epoch, void *stream)
{
    if (!a) return fail(F16_EINVAL, "args is NULL");
    if (epoch && n_minibatches < 1) return fail(F16_EINVAL, "f16_ppo_epoch16: n_minibatches >= 1");
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(a->n, &c, &g);
    if (rc) return rc;
    const bool packed = !a->actions;
    if (!a->train_blob16 || !a->obs || !a->ret || !a->mb_idx || !a->logstd || !a->grad || !a->scratch ||
        (!packed && (!a->actions || !a->logprobs || !a->adv || !a->val)))
        return fail(F16_EINVAL, "NULL pointer");
    if (packed && (reinterpret_cast<uintptr_t>(a->obs) % 32 != 0 || reinterpret_cast<uintptr_t>(a->ret) % 8 != 0))
        return fail(F16_EINVAL, "packed rows must be 32-byte aligned, (ret, val) pairs 8-byte aligned");
    if (a->norm_adv && !a->adv_stats) return fail(F16_EINVAL, "norm_adv needs adv_stats");
    if (reinterpret_cast<uintptr_t>(a->train_blob16) % 16 != 0) return fail(F16_EINVAL, "train_blob16 must be 16-byte aligned");
    if (a->n_params != f16_ppo_grad_floats() - 5) return fail(F16_EINVAL, "n_params must be f16_ppo_grad_floats() - 5");
    if (!a->param || !a->exp_avg || !a->exp_avg_sq || !a->max_exp_avg_sq || !a->step || !a->lr || !a->scatter || !a->fwd_blob || !a->tail_blob ||
        !a->half_blob || !a->sync)
        return fail(F16_EINVAL, "NULL pointer (optimiser state, scatter table, blobs or sync)");
    if (reinterpret_cast<uintptr_t>(a->scatter) % 16 != 0 || reinterpret_cast<uintptr_t>(a->sync) % 8 != 0)
        return fail(F16_EINVAL, "scatter must be 16-byte aligned, sync 8-byte aligned");
    f16::PpoGradArgs P;
    P.blob = nullptr;
    P.obs = static_cast<const float *>(a->obs);
    P.actions = static_cast<const float *>(a->actions);
    P.logprobs = static_cast<const float *>(a->logprobs);
    P.adv = static_cast<const float *>(a->adv);
    P.ret = static_cast<const float *>(a->ret);
    P.val = static_cast<const float *>(a->val);
    P.mb_idx = a->mb_idx;
    P.logstd = static_cast<const float *>(a->logstd);
    P.adv_stats = static_cast<const float *>(a->adv_stats);
    P.grad = static_cast<f16::PolicyGrad *>(a->grad);
    P.n = a->n;
    P.clip_coef = static_cast<float>(a->clip_coef);
    P.vf_coef = static_cast<float>(a->vf_coef);
    P.ent_coef = static_cast<float>(a->ent_coef);
    P.norm_adv = a->norm_adv;
    P.clip_vloss = a->clip_vloss;
    P.packed = packed ? 1 : 0;
    f16::ReduceAdamArgs R;
    std::memset(&R, 0, sizeof R);
    if (a->peer) {
        const f16_peer_group *peer = a->peer;
        if (peer->world < 1 || peer->world > 32 || peer->rank < 0 || peer->rank >= peer->world || !peer->bufs || !peer->seq)
            return fail(F16_EINVAL, "bad f16_peer_group (1 <= world <= 32, 0 <= rank < world, bufs and seq set)");
        R.bufs = reinterpret_cast<uint2 *const *>(peer->bufs);
        R.seq = peer->seq;
        R.rank = peer->rank;
        R.world = peer->world;
    }
    R.n = a->n_params;
    R.param = static_cast<float *>(a->param);
    R.exp_avg = static_cast<float *>(a->exp_avg);
    R.exp_avg_sq = static_cast<float *>(a->exp_avg_sq);
    R.max_exp_avg_sq = static_cast<float *>(a->max_exp_avg_sq);
    R.step = static_cast<float *>(a->step);
    R.lr = static_cast<const float *>(a->lr);
    R.beta1 = static_cast<float>(a->beta1);
    R.beta2 = static_cast<float>(a->beta2);
    R.eps = static_cast<float>(a->eps);
    R.max_grad_norm = static_cast<float>(a->max_grad_norm);
    R.scatter = a->scatter;
    R.fwd_blob = static_cast<float *>(a->fwd_blob);
    R.tail_blob = static_cast<float *>(a->tail_blob);
    R.half_blob = static_cast<__half *>(a->half_blob);
    R.diag = static_cast<float *>(a->diag);
    R.inv_rows = static_cast<float>(a->inv_rows);
    R.sync = static_cast<unsigned char *>(a->sync);
    if (epoch)
        CUDA_TRY(f16::launch_ppo_epoch16(P, static_cast<const f16::TrainBlob16 *>(a->train_blob16), static_cast<float *>(a->scratch), R, n_minibatches,
                                         c->n_sm, static_cast<cudaStream_t>(stream)));
    else
        CUDA_TRY(f16::launch_ppo_step16(P, static_cast<const f16::TrainBlob16 *>(a->train_blob16), static_cast<float *>(a->scratch), R, c->n_sm,
                                        static_cast<cudaStream_t>(stream)));
    return F16_OK;
}
int f16_ppo_step16(const f16_ppo_step16_args *a, void *stream) { return ppo_step16_impl(a, 1, false, stream); }
int f16_ppo_epoch16_sync_bytes(void) { return static_cast<int>(f16::ppo_epoch16_sync_bytes()); }
int f16_ppo_epoch16(const f16_ppo_step16_args *a, int n_minibatches, void *stream) { return ppo_step16_impl(a, n_minibatches, true, stream); }

// ---- exchange buffers in peer memory (CUDA IPC; the ranks of one node) ----
int64_t f16_peer_buffer_bytes(int world)
{
    if (world < 1 || world > 32) return -1;
    return static_cast<int64_t>(2) * world * f16::ppo_grad_floats_padded() * 8;   // (value, call) words [2 slots][world][padded gradient]
}
int f16_peer_seq_words(void) { return f16::ppo_grad_reduce_ctas(); }

int f16_peer_alloc(int64_t bytes, void **ptr, void *ipc_handle)
{
    if (bytes <= 0 || !ptr || !ipc_handle) return fail(F16_EINVAL, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle crosses the ABI as 64 opaque bytes");
    void *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, static_cast<size_t>(bytes)));
    cudaError_t e = cudaMemset(p, 0, static_cast<size_t>(bytes));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(F16_ECUDA, cudaGetErrorString(e));
    }
    std::memcpy(ipc_handle, &h, sizeof(h));
    *ptr = p;
    return F16_OK;
}

int f16_peer_open(const void *ipc_handle, void **ptr)
{
    if (!ipc_handle || !ptr) return fail(F16_EINVAL, "bad argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof(h));
    CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return F16_OK;
}

int f16_peer_close(void *ptr)
{
    if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return F16_OK;
}

int f16_peer_free(void *ptr)
{
    if (ptr) CUDA_TRY(cudaFree(ptr));
    return F16_OK;
}

int f16_ppo_adam(int n_params, void *param, const void *grad, void *exp_avg, void *exp_avg_sq, void *max_exp_avg_sq, void *step,
                 const void *lr, double beta1, double beta2, double eps, double max_grad_norm, const int64_t *fwd_perm, void *fwd_blob,
                 int n_fwd, const int64_t *train_perm, void *train_blob, int n_train, const int64_t *half_perm, void *half_blob, int n_half,
                 void *diag, const void *logstd, double inv_rows, int zero_grad, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n_params, &c, &g);
    if (rc) return rc;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !max_exp_avg_sq || !step || !lr) return fail(F16_EINVAL, "NULL pointer");
    if ((n_fwd > 0 && (!fwd_perm || !fwd_blob)) || (n_train > 0 && (!train_perm || !train_blob)) || n_fwd < 0 || n_train < 0)
        return fail(F16_EINVAL, "blob permutation / destination missing");
    f16::PpoAdamArgs A;
    A.n = n_params;
    A.param = static_cast<float *>(param);
    A.grad = static_cast<const float *>(grad);
    A.exp_avg = static_cast<float *>(exp_avg);
    A.exp_avg_sq = static_cast<float *>(exp_avg_sq);
    A.max_exp_avg_sq = static_cast<float *>(max_exp_avg_sq);
    A.step = static_cast<float *>(step);
    A.lr = static_cast<const float *>(lr);
    A.beta1 = static_cast<float>(beta1);
    A.beta2 = static_cast<float>(beta2);
    A.eps = static_cast<float>(eps);
    A.max_grad_norm = static_cast<float>(max_grad_norm);
    A.fwd_perm = fwd_perm;
    A.fwd_blob = static_cast<float *>(fwd_blob);
    A.n_fwd = n_fwd;
    A.train_perm = train_perm;
    A.train_blob = static_cast<float *>(train_blob);
    A.n_train = n_train;
    if (n_half < 0 || (n_half > 0 && (!half_perm || !half_blob))) return fail(F16_EINVAL, "half blob permutation / destination missing");
    A.half_perm = half_perm;
    A.half_blob = static_cast<__half *>(half_blob);
    A.n_half = n_half;
    if (diag && !logstd) return fail(F16_EINVAL, "diag needs logstd");
    A.diag = static_cast<float *>(diag);
    A.grad_rw = static_cast<float *>(const_cast<void *>(grad));
    A.logstd = static_cast<const float *>(logstd);
    A.inv_rows = static_cast<float>(inv_rows);
    A.zero_grad = zero_grad;
    CUDA_TRY(f16::launch_ppo_adam(A, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

int f16_gae(int T, int64_t n, const void *rewards, const void *values, const uint8_t *done_after, const void *next_value, double gamma,
            double gae_lambda, void *advantages, void *returns, void *stream)
{
    DeviceCtx *c;
    f16::LaunchGeom g;
    int rc = prep(n, &c, &g);
    if (rc) return rc;
    if (T <= 0 || !rewards || !values || !done_after || !next_value || !advantages || !returns)
        return fail(F16_EINVAL, "NULL pointer or T <= 0");
    CUDA_TRY(f16::launch_gae(T, n, (const float *)rewards, (const float *)values, done_after, (const float *)next_value, (float)gamma,
                             (float)gae_lambda, (float *)advantages, (float *)returns, c->n_sm, static_cast<cudaStream_t>(stream)));
    return F16_OK;
}

}  // extern "C"
