// f16_launch.h -- typed kernel arguments and launcher prototypes shared by the two kernel
// translation units (f32 / f64) and the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "f16_tables.h"

namespace f16 {

struct LaunchGeom {
    int n_sm;
    int ctas_per_sm;
    int pdl = 1;   // launch the step kernel with programmatic stream serialization (F16_FLAG_NO_PDL disables)
    int two_per_thread = 0;   // fp32 gym step: use the two-envs-per-thread kernel on full tiles
    int ctas2_per_sm = 4;    // resident CTAs of the two-env gym-step kernel (128 threads)
    int ctas2k_per_sm = 2;   // ... of its K-step variant (256 threads)
};

// typed view of f16_config + f16_env_buffers (include/f16_b200.h)
template <typename R>
struct EnvArgs {
    const Tables<R> *tables;
    int64_t n;
    R dt;
    int horizon, ref_len, n_ref;
    int seg_len, seg_stride;    // segment mode of the reference bank (seg_len > 0), see ref_sample() / include/f16_b200.h
    int autoreset, init_mode, obs_layout, noise;
    uint64_t seed;
    const uint64_t *seed_dev;   // optional device copy of the seed (read on the rare reset / noise paths): re-keying reaches captured CUDA graphs
    int64_t id_base;
    R x0[9];
    R *state;
    int32_t *ep_len;
    R *total_return;
    int32_t *ref_idx;
    uint8_t *done;
    uint32_t *episode_no;
    const R *ref_bank;
    const uint32_t *ref_codes;  // segment mode: packed segment rows of each of the n_ref signals (read by resets only)
    const R *init_state;
    R *pos_comp;   // [2][n] Kahan carry of Ox, Oy (fp32 mode), or nullptr
};

// typed view of f16_step_io
template <typename R>
struct StepArgs {
    const R *actions;
    R *obs;
    R *reward;
    uint8_t *done;
    R *final_obs;
    int32_t *final_ep_len;
    R *final_return;
    double *stats;            // [4] accumulators: episodes finished, sum of their returns, sum of their lengths, spare
    int obs_every_step;
    int K;
    int tma_ok;               // every input row of a full tile is 16-byte aligned: TMA-staged loads allowed
    int obs_bulk;             // AoS obs rows of a full tile are 16-byte aligned: staged in shared memory, written by bulk stores
    int pdl_mode;             // 0: normal; 1: head of a split step (dependents released after the own wait); 2: tail (waits at its END)
    int64_t i_begin, i_end;   // env sub-range advanced by this launch (host-buffer path pipelines chunks)
};

// f16_kernels_f32.cu
cudaError_t launch_env_step(const EnvArgs<float> &, const StepArgs<float> &, bool fast, const LaunchGeom &, cudaStream_t);
cudaError_t launch_env_reset(const EnvArgs<float> &, const uint8_t *mask, int keep_ref, float *obs, const LaunchGeom &, cudaStream_t);
cudaError_t launch_rhs(const Tables<float> *, int64_t n, const float *x, const float *u, float *dx, const LaunchGeom &, cudaStream_t);
cudaError_t launch_model_step(const Tables<float> *, int64_t n, float *x, const float *u, float dt, int K, const LaunchGeom &, cudaStream_t);
cudaError_t launch_coeffs(const Tables<float> *, int64_t n, const float *, const float *, const float *, const float *, float *, const LaunchGeom &, cudaStream_t);
cudaError_t launch_thrust(const Tables<float> *, int64_t n, const float *, const float *, const float *, float *, const LaunchGeom &, cudaStream_t);
cudaError_t launch_isa(int64_t n, const float *, float *, float *, const LaunchGeom &, cudaStream_t);
int env_step_occupancy_f32(bool fast, bool autoreset);
int env_step2_occupancy_f32(bool fast);

// f16_kernels_f64.cu
cudaError_t launch_env_step(const EnvArgs<double> &, const StepArgs<double> &, bool fast, const LaunchGeom &, cudaStream_t);
cudaError_t launch_env_reset(const EnvArgs<double> &, const uint8_t *mask, int keep_ref, double *obs, const LaunchGeom &, cudaStream_t);
cudaError_t launch_rhs(const Tables<double> *, int64_t n, const double *x, const double *u, double *dx, const LaunchGeom &, cudaStream_t);
cudaError_t launch_model_step(const Tables<double> *, int64_t n, double *x, const double *u, double dt, int K, const LaunchGeom &, cudaStream_t);
cudaError_t launch_coeffs(const Tables<double> *, int64_t n, const double *, const double *, const double *, const double *, double *, const LaunchGeom &, cudaStream_t);
cudaError_t launch_thrust(const Tables<double> *, int64_t n, const double *, const double *, const double *, double *, const LaunchGeom &, cudaStream_t);
cudaError_t launch_isa(int64_t n, const double *, double *, double *, const LaunchGeom &, cudaStream_t);
int env_step_occupancy_f64(bool autoreset);
cudaError_t launch_trim(const Tables<double> *, int n_cells, int n_starts, const double *V, const double *H, const double *starts,
                        double *out_x, double *out_cost, int *out_iters, const LaunchGeom &, cudaStream_t);

// f16_peaks.cu: kind 0 = FP32 FMA, 1 = FP64 FMA, 2 = MUFU.TANH; *ops = operations one launch performs
cudaError_t launch_peak_probe(int kind, int iters, int n_sm, float *sink, double *ops, cudaStream_t);

}  // namespace f16
