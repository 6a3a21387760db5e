// fp32 instantiation of the kernels (throughput mode; FMA contraction on).
#include "f16_kernels.cuh"

namespace f16 {

cudaError_t launch_env_step(const EnvArgs<float> &A, const StepArgs<float> &S, bool fast, const LaunchGeom &g, cudaStream_t st)
{
    return fast ? launch_env_step_t<float, true>(A, S, g, st) : launch_env_step_t<float, false>(A, S, g, st);
}
cudaError_t launch_env_reset(const EnvArgs<float> &A, const uint8_t *mask, int keep_ref, float *obs, const LaunchGeom &g, cudaStream_t st)
{
    env_reset_kernel<float><<<grid_for(A.n, g), kCtaThreads, 0, st>>>(A, mask, keep_ref, obs);
    return cudaGetLastError();
}
cudaError_t launch_rhs(const Tables<float> *T, int64_t n, const float *x, const float *u, float *dx, const LaunchGeom &g, cudaStream_t st)
{
    rhs_kernel<float><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, x, u, dx);
    return cudaGetLastError();
}
cudaError_t launch_model_step(const Tables<float> *T, int64_t n, float *x, const float *u, float dt, int K, const LaunchGeom &g, cudaStream_t st)
{
    model_step_kernel<float><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, x, u, dt, K);
    return cudaGetLastError();
}
cudaError_t launch_coeffs(const Tables<float> *T, int64_t n, const float *a, const float *f, const float *w, const float *V, float *out,
                          const LaunchGeom &g, cudaStream_t st)
{
    coeffs_kernel<float><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, a, f, w, V, out);
    return cudaGetLastError();
}
cudaError_t launch_thrust(const Tables<float> *T, int64_t n, const float *H, const float *M, const float *Pa, float *out, const LaunchGeom &g,
                          cudaStream_t st)
{
    thrust_kernel<float><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, H, M, Pa, out);
    return cudaGetLastError();
}
cudaError_t launch_isa(int64_t n, const float *h, float *rho, float *a, const LaunchGeom &g, cudaStream_t st)
{
    isa_kernel<float><<<grid_for(n, g), kCtaThreads, 0, st>>>(n, h, rho, a);
    return cudaGetLastError();
}
int env_step_occupancy_f32(bool fast, bool autoreset)
{
    if (fast) return autoreset ? prepare_env_step<float, true, true>() : prepare_env_step<float, true, false>();
    return autoreset ? prepare_env_step<float, false, true>() : prepare_env_step<float, false, false>();
}

int env_step2_occupancy_f32(bool fast) { return fast ? prepare_env_step2<float, true>() : prepare_env_step2<float, false>(); }

}  // namespace f16
