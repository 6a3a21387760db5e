// f16_device.cuh -- device-side physics of the F-16 longitudinal model, templated on
// the arithmetic type.  One thread advances one aircraft; everything here is
// register-resident scalar code plus table reads from the CTA's shared-memory image.
//
// Reference (what each routine replaces; paths relative to the reference root):
//   isa()          F16model/data/atmosphere.py:5-11  (ambiance.Atmosphere, ICAO 1993)
//   aero_coeffs()  F16model/data/coeff.py:8-63,225-262 via utils/interpolator.py:21-44
//   thrust()       F16model/data/thrust.py:6-14
//   power_level()  F16model/model/engine.py:4-32
//   rhs()          F16model/model/ODE_3DoF.py:8-60
//   euler_step()   F16model/model/_interface.py:116-122
//
// Precision policy.  double: the parity mode -- reference operation order, true
// divisions, libm sin/cos/pow; this TU is compiled with -fmad=false so products and
// sums round exactly like CPython/NumPy scalars.  float: the throughput mode --
// FMA contraction on, divisions by constants folded into reciprocals, and with
// FAST=true the MUFU intrinsics (sin/cos/lg2/ex2/rcp/rsq).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "f16_tables.h"

namespace f16 {

// ---------------------------------------------------------------------------
// constants (F16model/data/plane.py:3-18, data/atmosphere.py:3)
// ---------------------------------------------------------------------------
namespace K {
constexpr double kPi = 3.141592653589793;
constexpr double m = 9295.44, S = 27.87, ba = 3.45, Jz = 75673.6;
constexpr double rcgx = -0.05 * 3.45;
constexpr double Tstab = 0.03, Xistab = 0.707;
constexpr double maxabsstab = 25.0 * (kPi / 180.0);    // np.radians(25)
constexpr double maxabsdstab = 60.0 * (kPi / 180.0);   // np.radians(60)
constexpr double wz_clip = 60.0 * (kPi / 180.0);       // _interface.py:118
constexpr double wz_term = 50.0 * (kPi / 180.0);       // env.py:112
constexpr double wz_bound = 150.0 * (kPi / 180.0);     // plane.state_bound["wz"]
constexpr double g = 9.80665;
constexpr double rad2deg = 180.0 / kPi;
// normalize_value() casts the bounds to float32 and subtracts in float32 (utils/calc.py:8-10):
// low = float32(-radians(25)) and (high - low) as exactly-representable doubles.
constexpr double act_lo = -0.4363323152065277;
constexpr double act_span = 0.8726646304130554;
// ISA (ambiance 1.3.1 CONST)
constexpr double isa_g0 = 9.80665, isa_R = 287.05287, isa_kappa = 1.4, isa_r = 6356766.0;
// alpha1 grid: -20..60 deg in 5 deg steps (17 nodes), then 70, 80, 90
constexpr double alpha_min = -20.0 * (kPi / 180.0);
constexpr double alpha_max = 90.0 * (kPi / 180.0);
constexpr double inv_dalpha = 1.0 / (5.0 * (kPi / 180.0));
constexpr double fi_m25 = -25.0 * (kPi / 180.0), fi_m10 = -10.0 * (kPi / 180.0), fi_p10 = 10.0 * (kPi / 180.0);
constexpr double fi_p15 = 15.0 * (kPi / 180.0), fi_p20 = 20.0 * (kPi / 180.0), fi_p25 = 25.0 * (kPi / 180.0);
constexpr double thrust_dH = 10000.0 * 0.3048, thrust_dM = 0.2;
}  // namespace K

// ISA layers above/below the 0-11 km troposphere (slow path only)
struct IsaLayer { double Hb, Tb, beta, pb; };
static __constant__ IsaLayer c_isa_layers[8] = {
    {-5.0e3, 320.65, -6.5e-3, 1.77687e+5}, {0.0, 288.15, -6.5e-3, 1.01325e+5},
    {11.0e3, 216.65, 0.0, 2.26320e+4},     {20.0e3, 216.65, 1.0e-3, 5.47487e+3},
    {32.0e3, 228.65, 2.8e-3, 8.68014e+2},  {47.0e3, 270.65, 0.0, 1.10906e+2},
    {51.0e3, 270.65, -2.8e-3, 6.69384e+1}, {71.0e3, 214.65, -2.0e-3, 3.95639e+0},
};

// ---------------------------------------------------------------------------
// math policies
// ---------------------------------------------------------------------------
template <typename R, bool FAST>
struct Mth;

template <>
struct Mth<double, false> {
    static __device__ __forceinline__ void sincos(double x, double &s, double &c) { ::sincos(x, &s, &c); }
    static __device__ __forceinline__ double div(double a, double b) { return a / b; }
    // division by a compile-time constant: keep the true division (reference order)
    static __device__ __forceinline__ double divc(double a, double c, double /*rc*/) { return a / c; }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double pow(double b, double e) { return ::pow(b, e); }
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double fabs(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double fmin(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double fmax(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ int f2i(double x) { return __double2int_rd(x); }
};

template <>
struct Mth<float, false> {
    static __device__ __forceinline__ void sincos(float x, float &s, float &c) { ::sincosf(x, &s, &c); }
    static __device__ __forceinline__ float div(float a, float b) { return a / b; }
    static __device__ __forceinline__ float divc(float a, float /*c*/, float rc) { return a * rc; }
    static __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
    static __device__ __forceinline__ float pow(float b, float e) { return ::powf(b, e); }
    static __device__ __forceinline__ float exp(float x) { return ::expf(x); }
    static __device__ __forceinline__ float log(float x) { return ::logf(x); }
    static __device__ __forceinline__ float fabs(float x) { return ::fabsf(x); }
    static __device__ __forceinline__ float fmin(float a, float b) { return ::fminf(a, b); }
    static __device__ __forceinline__ float fmax(float a, float b) { return ::fmaxf(a, b); }
    static __device__ __forceinline__ int f2i(float x) { return __float2int_rd(x); }
};

template <>
struct Mth<float, true> {
    static __device__ __forceinline__ void sincos(float x, float &s, float &c) { __sincosf(x, &s, &c); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
    static __device__ __forceinline__ float divc(float a, float /*c*/, float rc) { return a * rc; }
    static __device__ __forceinline__ float sqrt(float x) { return x * rsqrtf(x); }
    static __device__ __forceinline__ float pow(float b, float e) { return exp2f(e * __log2f(b)); }
    static __device__ __forceinline__ float exp(float x) { return __expf(x); }
    static __device__ __forceinline__ float log(float x) { return __logf(x); }
    static __device__ __forceinline__ float fabs(float x) { return ::fabsf(x); }
    static __device__ __forceinline__ float fmin(float a, float b) { return ::fminf(a, b); }
    static __device__ __forceinline__ float fmax(float a, float b) { return ::fmaxf(a, b); }
    static __device__ __forceinline__ int f2i(float x) { return __float2int_rd(x); }
};

template <typename R>
__device__ __forceinline__ R clampr(R x, R lo, R hi)
{
    return x < lo ? lo : (x > hi ? hi : x);   // np.clip / min(max()) semantics for finite x
}

// 16-byte shared-memory vector loads of 4 table entries
template <typename R>
struct Vec4;
template <>
struct Vec4<float> {
    float x, y, z, w;
    static __device__ __forceinline__ Vec4 ld(const float *p)
    {
        float4 v = *reinterpret_cast<const float4 *>(p);
        return {v.x, v.y, v.z, v.w};
    }
};
template <>
struct Vec4<double> {
    double x, y, z, w;
    static __device__ __forceinline__ Vec4 ld(const double *p)
    {
        double2 a = *reinterpret_cast<const double2 *>(p);
        double2 b = *reinterpret_cast<const double2 *>(p + 2);
        return {a.x, a.y, b.x, b.y};
    }
};

// ---------------------------------------------------------------------------
// ISA atmosphere: density and speed of sound at geometric height h
// ---------------------------------------------------------------------------
// troposphere branch (0 <= H < 11 km): every BASELINE workload lives here
template <typename R, bool FAST>
__device__ __forceinline__ void isa_troposphere(R H, R &T, R &p)
{
    using M = Mth<R, FAST>;
    constexpr double Tb = 288.15, beta = -6.5e-3, pb = 1.01325e5;
    constexpr double expo = (1.0 / beta) * (-K::isa_g0 / K::isa_R);
    T = R(Tb) + R(beta) * (H - R(0));
    p = R(pb) * M::pow(R(1) + R(beta / Tb) * (H - R(0)), R(expo));
}

// any other layer of the table (slow path)
template <typename R, bool FAST>
__device__ __forceinline__ void isa_layer(R H, R &T, R &p)
{
    using M = Mth<R, FAST>;
    int L = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) L = (H >= R(c_isa_layers[i].Hb)) ? i : L;
    const R Hb = R(c_isa_layers[L].Hb), Tb = R(c_isa_layers[L].Tb), beta = R(c_isa_layers[L].beta);
    const R pb = R(c_isa_layers[L].pb);
    T = Tb + beta * (H - Hb);
    if (beta == R(0)) {
        p = pb * M::exp(M::div(R(-K::isa_g0), R(K::isa_R) * T) * (H - Hb));
    } else {
        const R expo = M::div(R(1), beta) * R(-K::isa_g0 / K::isa_R);
        p = pb * M::pow(R(1) + M::div(beta, Tb) * (H - Hb), expo);
    }
}

template <typename R, bool FAST>
__device__ __forceinline__ R isa_geopotential(R h)
{
    if constexpr (FAST) return __fdividef(h, h * R(1.0 / K::isa_r) + R(1));   // h / (1 + h/r): one FMA, one MUFU.RCP, one FMUL
    else return Mth<R, FAST>::div(R(K::isa_r) * h, R(K::isa_r) + h);
}

// FAST fp32, troposphere: density and speed of sound straight from u = T/Tb = 1 + (beta/Tb) H, without forming T and p:
//   rho = p / (R T) = pb/(R Tb) * u^(expo - 1),   a = sqrt(kappa R Tb) * u^(1/2)
// -- one MUFU.LG2 shared by two MUFU.EX2 (3 FMA/FMUL + 3 MUFU instead of 8 + 4).  Both exponents stay O(1), so the
// approximate lg2/ex2 keep their ~2^-22 relative accuracy.  expo = g0 / (-beta R) = 5.2558798...
template <typename R>
__device__ __forceinline__ void isa_troposphere_direct(R H, R &rho, R &a)
{
    const R L = __log2f(H * R(-6.5e-3 / 288.15) + R(1));
    rho = exp2f(L * R(4.255879812716676) + R(0.29278177057300514));   // lg2(pb / (R Tb)) = lg2(1.2250)
    a = R(340.29398803459594) * exp2f(R(0.5) * L);                    // sqrt(kappa R Tb)
}

template <typename R, bool FAST>
__device__ __forceinline__ void isa_finish(R T, R p, R &rho, R &a)
{
    using M = Mth<R, FAST>;
    rho = M::div(p, R(K::isa_R) * T);
    a = M::sqrt(R(K::isa_kappa) * R(K::isa_R) * T);
}

template <typename R, bool FAST>
__device__ __forceinline__ void isa(R h, R &rho, R &a)
{
    const R H = isa_geopotential<R, FAST>(h);
    R T, p;
    if constexpr (FAST) {
        if (H >= R(0) && H < R(11000)) {   // warp-uniform in practice
            isa_troposphere_direct<R>(H, rho, a);
            return;
        }
        isa_layer<R, FAST>(H, T, p);
    } else {
        if (H >= R(0) && H < R(11000)) isa_troposphere<R, FAST>(H, T, p);
        else isa_layer<R, FAST>(H, T, p);
    }
    isa_finish<R, FAST>(T, p, rho, a);
}

// ---------------------------------------------------------------------------
// closed-form cell indices
// ---------------------------------------------------------------------------
// alpha1 interval of a value already clipped to [alpha_min, alpha_max]
template <typename R, bool FAST>
__device__ __forceinline__ int alpha_cell(R ac)
{
    int j = Mth<R, FAST>::f2i((ac - R(K::alpha_min)) * R(K::inv_dalpha));   // 0..22
    j = j < 16 ? j : 16 + ((j - 16) >> 1);
    return min(max(j, 0), kNA - 1);
}

// ---------------------------------------------------------------------------
// aerodynamic coefficients at beta=0, lef=0, sb=0
// ---------------------------------------------------------------------------
template <typename R, bool FAST>
__device__ __forceinline__ void aero_coeffs(const Tables<R> &T, R alpha, R fi, R qhat, R &cx, R &cy, R &mz)
{
    // Interpolator clips the query to the grid for the linear tables (utils/interpolator.py:22-24);
    // pchip_interpolate does NOT clip: it extrapolates the edge cubic.
    const R ac = clampr(alpha, R(K::alpha_min), R(K::alpha_max));
    const R fc = clampr(fi, R(K::fi_m25), R(K::fi_p25));
    const int ia = alpha_cell<R, FAST>(ac);
    const int i1 = (fc >= R(K::fi_m10)) + (fc >= R(0)) + (fc >= R(K::fi_p10));          // fi1 interval 0..3
    const int i3 = i1 + (fc >= R(K::fi_p15)) + (fc >= R(K::fi_p20));                    // fi3 interval 0..5
    // interval origins
    const R *row = T.arow[ia];
    const Vec4<R> r3 = Vec4<R>::ld(row + 12);                                            // dmz1 y0, slope, alpha_i
    const R a_i = r3.z;
    // interval origins as select chains on the same predicates (a switch on i1 compiles to a jump table)
    const R f1_i = fc >= R(K::fi_p10) ? R(K::fi_p10) : (fc >= R(0) ? R(0) : (fc >= R(K::fi_m10) ? R(K::fi_m10) : R(K::fi_m25)));
    const R f3_i = fc >= R(K::fi_p20) ? R(K::fi_p20) : (fc >= R(K::fi_p15) ? R(K::fi_p15) : f1_i);
    const R sa = ac - a_i;      // linear tables: clipped alpha
    const R s = alpha - a_i;    // PCHIP: raw alpha
    const R sf = fc - f1_i;
    const R sf3 = fc - f3_i;
    const R se = fi - f1_i;     // eta_fi PCHIP: raw fi

    const R *c3 = T.cell3[i1][ia];
    const Vec4<R> px = Vec4<R>::ld(c3), py = Vec4<R>::ld(c3 + 4), pm = Vec4<R>::ld(c3 + 8);
    const R Cx = (px.x + px.y * sf) + sa * (px.z + px.w * sf);
    const R Cy = (py.x + py.y * sf) + sa * (py.z + py.w * sf);
    const R m0 = (pm.x + pm.y * sf) + sa * (pm.z + pm.w * sf);

    const Vec4<R> kx = Vec4<R>::ld(row), ky = Vec4<R>::ld(row + 4), km = Vec4<R>::ld(row + 8);
    const R Cxwz = ((kx.x * s + kx.y) * s + kx.z) * s + kx.w;
    const R Cywz = ((ky.x * s + ky.y) * s + ky.z) * s + ky.w;
    const R mzwz = ((km.x * s + km.y) * s + km.z) * s + km.w;

    const R dmz = r3.y * sa + r3.x;
    const Vec4<R> pd = Vec4<R>::ld(T.dmzds[ia][i3]);
    const R dmz_ds = (pd.x + pd.y * sf3) + sa * (pd.z + pd.w * sf3);
    const Vec4<R> ke = Vec4<R>::ld(T.eta[i1]);
    const R eta = ((ke.x * se + ke.y) * se + ke.z) * se + ke.w;

    cx = Cx + Cxwz * qhat;                                   // coeff.py:57-62
    cy = Cy + Cywz * qhat;                                   // coeff.py:29-34
    mz = ((m0 * eta + dmz) + mzwz * qhat) + dmz_ds;          // coeff.py:255-262
}

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
template <typename R, bool FAST>
__device__ __forceinline__ R thrust(const Tables<R> &T, R H, R Mach, R Pa)
{
    using M = Mth<R, FAST>;
    // edge cells extrapolate linearly (RegularGridInterpolator fill_value=None)
    const int ip = Pa >= R(50) ? 1 : 0;
    const int ih = min(max(M::f2i(H * R(1.0 / K::thrust_dH)), 0), 4);
    const int im = min(max(M::f2i(Mach * R(1.0 / K::thrust_dM)), 0), 4);
    const R p = Pa - R(50) * R(ip);
    const R h = H - R(K::thrust_dH) * R(ih);
    const R m = Mach - R(K::thrust_dM) * R(im);
    const R *t = T.thrust[ip][ih] + 8 * im;
    const Vec4<R> lo = Vec4<R>::ld(t), hi = Vec4<R>::ld(t + 4);
    return ((lo.x + lo.y * m) + h * (lo.z + lo.w * m)) + p * ((hi.x + hi.y * m) + h * (hi.z + hi.w * m));
}

template <typename R>
__device__ __forceinline__ R power_level(R Pa, R thr)
{
    R Pc = thr <= R(0.77) ? R(64.94) * thr : R(217.38) * thr - R(117.38);
    if (Pc >= R(50) && Pa < R(50)) Pc = R(60);
    else if (Pc < R(50) && Pa >= R(50)) Pc = R(40);
    const R dP = Pc - Pa;
    R w = dP <= R(25) ? R(1.0) : (dP >= R(50) ? R(0.1) : R(1.9) - R(0.036) * dP);
    if (Pa >= R(50)) w = R(5);
    return w * dP;
}

// ---------------------------------------------------------------------------
// right-hand side.  x = [Ox,Oy,wz,theta,V,alpha,stab,dstab,Pa]; u = (stab cmd, throttle).
// WANT_VDOT: also compute V_dot (the integrator discards it; trim search needs it).
// ---------------------------------------------------------------------------
// rhs_atm: the right-hand side with the atmosphere (rho, speed of sound at x[1]) supplied by the caller.
// U_CLAMPED: the caller's stab command is already inside +-maxabsstab (the env's rescaled action in fp32, where
// float(act_lo) == -float(maxabsstab) and act_lo + act_span == float(maxabsstab)): the clip of ODE_3DoF.py:37 is a no-op.
// fp32 only: constant factors are folded (0.5 S into q S, 1/Jz and ba into the moment, ba/2 into qhat) -- the fp64
// parity mode keeps the reference's operation order.
template <typename R, bool FAST, bool WANT_VDOT, bool U_CLAMPED = false>
__device__ __forceinline__ void rhs_atm(const Tables<R> &T, const R (&x)[9], R u_stab, R u_thr, R rho, R asnd, R (&dx)[9])
{
    constexpr bool F32 = sizeof(R) == 4;
    using M = Mth<R, FAST>;
    const R Oy = x[1], wz = x[2], theta = x[3], V = x[4], alpha = x[5], stab = x[6], dstab = x[7], Pa = x[8];
    R sa, ca, st, ct;
    M::sincos(alpha, sa, ca);
    M::sincos(theta, st, ct);
    const R Vx = V * ca;       // wind2body with beta = 0 (utils/cs_transform.py:11-15)
    const R Vy = -V * sa;

    const R Mach = M::div(V, asnd);

    dx[0] = ct * Vx - st * Vy;
    dx[1] = st * Vx + ct * Vy;
    dx[3] = wz;

    R qhat;
    if constexpr (F32) qhat = M::div(wz * R(0.5 * K::ba), V);
    else qhat = M::div(wz * R(K::ba), R(2) * V);
    R cx, cy, mz;
    aero_coeffs<R, FAST>(T, alpha, stab, qhat, cx, cy, mz);

    R qS;
    if constexpr (F32) qS = (rho * (V * V)) * R(0.5 * K::S);
    else qS = ((rho * (V * V)) * R(0.5)) * R(K::S);
    const R X = -qS * cx;
    const R Y = qS * cy;
    const R Px = thrust<R, FAST>(T, Oy, Mach, Pa);
    const R Rx = X + Px;
    const R Ry = Y;

    const R cs = (U_CLAMPED && F32) ? u_stab : M::fmin(M::fmax(u_stab, R(-K::maxabsstab)), R(K::maxabsstab));
    const R cth = M::fmin(M::fmax(u_thr, R(0)), R(1));

    const R Vx_dot = (wz * Vy - R(K::g) * st) + M::divc(Rx, R(K::m), R(1.0 / K::m));
    const R Vy_dot = (-wz * Vx - R(K::g) * ct) + M::divc(Ry, R(K::m), R(1.0 / K::m));

    dx[4] = WANT_VDOT ? M::div(Vx * Vx_dot - Vy * Vy_dot, V) : R(0);
    dx[5] = M::div(-Vx * Vy_dot + Vy * Vx_dot, Vx * Vx + Vy * Vy);
    if constexpr (F32) {
        dx[2] = (qS * R(1.0 / K::Jz)) * (R(K::ba) * mz + R(K::rcgx) * cy);   // (Mz + rcgx Ry) / Jz with q S factored out
    } else {
        const R Mz = qS * R(K::ba) * mz;
        const R MRz = Mz + R(K::rcgx) * Ry;
        dx[2] = M::divc(MRz, R(K::Jz), R(1.0 / K::Jz));
    }
    dx[6] = M::fmin(M::fmax(dstab, R(-K::maxabsdstab)), R(K::maxabsdstab));
    dx[7] = M::divc(R(-2 * K::Tstab) * R(K::Xistab) * dstab - stab + cs, R(K::Tstab * K::Tstab),
                    R(1.0 / (K::Tstab * K::Tstab)));
    dx[8] = power_level<R>(Pa, cth);
}

template <typename R, bool FAST, bool WANT_VDOT, bool U_CLAMPED = false>
__device__ __forceinline__ void rhs(const Tables<R> &T, const R (&x)[9], R u_stab, R u_thr, R (&dx)[9])
{
    R rho, asnd;
    isa<R, FAST>(x[1], rho, asnd);
    rhs_atm<R, FAST, WANT_VDOT, U_CLAMPED>(T, x, u_stab, u_thr, rho, asnd, dx);
}

// F16model.step: explicit Euler, clip wz, V frozen (dx[4] ignored).
// comp: Kahan-compensated accumulation of the two position integrals (fp32 mode): Ox, Oy are pure
// accumulators of ~1e-4-relative increments, where plain fp32 drifts by up to 1 ulp per step; the
// carries (ox_lo, oy_lo) persist across steps and launches.
template <typename R, bool FAST, bool U_CLAMPED = false>
__device__ __forceinline__ void euler_step(const Tables<R> &T, R (&x)[9], R u_stab, R u_thr, R dt, bool comp, R &ox_lo, R &oy_lo)
{
    R dx[9];
    rhs<R, FAST, false, U_CLAMPED>(T, x, u_stab, u_thr, dx);
    if (comp) {
        R y = dx[0] * dt + ox_lo;
        R t = x[0] + y;
        ox_lo = y - (t - x[0]);
        x[0] = t;
        y = dx[1] * dt + oy_lo;
        t = x[1] + y;
        oy_lo = y - (t - x[1]);
        x[1] = t;
    } else {
        x[0] = dx[0] * dt + x[0];
        x[1] = dx[1] * dt + x[1];
    }
#pragma unroll
    for (int i = 2; i < 9; ++i)
        if (i != 4) x[i] = dx[i] * dt + x[i];
    x[2] = clampr(x[2], R(-K::wz_clip), R(K::wz_clip));
}

template <typename R, bool FAST>
__device__ __forceinline__ void euler_step(const Tables<R> &T, R (&x)[9], R u_stab, R u_thr, R dt)
{
    R a = R(0), b = R(0);
    euler_step<R, FAST>(T, x, u_stab, u_thr, dt, false, a, b);
}

}  // namespace f16
