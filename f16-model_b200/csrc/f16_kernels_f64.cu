// fp64 instantiation of the kernels: the parity build.  Compiled with -fmad=false so that
// a*b+c rounds twice, exactly like the reference's CPython/NumPy scalar arithmetic.
#include "f16_kernels.cuh"

namespace f16 {

cudaError_t launch_env_step(const EnvArgs<double> &A, const StepArgs<double> &S, bool /*fast*/, const LaunchGeom &g, cudaStream_t st)
{
    return launch_env_step_t<double, false>(A, S, g, st);
}
cudaError_t launch_env_reset(const EnvArgs<double> &A, const uint8_t *mask, int keep_ref, double *obs, const LaunchGeom &g, cudaStream_t st)
{
    env_reset_kernel<double><<<grid_for(A.n, g), kCtaThreads, 0, st>>>(A, mask, keep_ref, obs);
    return cudaGetLastError();
}
cudaError_t launch_rhs(const Tables<double> *T, int64_t n, const double *x, const double *u, double *dx, const LaunchGeom &g, cudaStream_t st)
{
    rhs_kernel<double><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, x, u, dx);
    return cudaGetLastError();
}
cudaError_t launch_model_step(const Tables<double> *T, int64_t n, double *x, const double *u, double dt, int K, const LaunchGeom &g, cudaStream_t st)
{
    model_step_kernel<double><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, x, u, dt, K);
    return cudaGetLastError();
}
cudaError_t launch_coeffs(const Tables<double> *T, int64_t n, const double *a, const double *f, const double *w, const double *V, double *out,
                          const LaunchGeom &g, cudaStream_t st)
{
    coeffs_kernel<double><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, a, f, w, V, out);
    return cudaGetLastError();
}
cudaError_t launch_thrust(const Tables<double> *T, int64_t n, const double *H, const double *M, const double *Pa, double *out, const LaunchGeom &g,
                          cudaStream_t st)
{
    thrust_kernel<double><<<grid_for(n, g), kCtaThreads, 0, st>>>(T, n, H, M, Pa, out);
    return cudaGetLastError();
}
cudaError_t launch_isa(int64_t n, const double *h, double *rho, double *a, const LaunchGeom &g, cudaStream_t st)
{
    isa_kernel<double><<<grid_for(n, g), kCtaThreads, 0, st>>>(n, h, rho, a);
    return cudaGetLastError();
}
cudaError_t launch_trim(const Tables<double> *T, int n_cells, int n_starts, const double *V, const double *H, const double *starts,
                        double *out_x, double *out_cost, int *out_iters, const LaunchGeom &g, cudaStream_t st)
{
    const int64_t total = (int64_t)n_cells * n_starts;
    const int64_t blocks = (total + 127) / 128;
    const int64_t cap = (int64_t)g.n_sm * 8;
    trim_kernel<double><<<(int)(blocks < cap ? blocks : cap), 128, 0, st>>>(T, n_cells, n_starts, V, H, starts, out_x, out_cost, out_iters);
    return cudaGetLastError();
}
int env_step_occupancy_f64(bool autoreset)
{
    return autoreset ? prepare_env_step<double, false, true>() : prepare_env_step<double, false, false>();
}

}  // namespace f16
