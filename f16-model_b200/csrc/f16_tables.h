// f16_tables.h -- packed table image shared by host (builder) and device (kernels).
//
// The reference rebuilds ten PCHIP splines and eleven grid interpolators per
// step from the raw tables (data/coeff.py, data/thrust.py).  Here
// every table is turned ONCE, on the host in fp64, into per-cell polynomial
// coefficients so that a lookup is "closed-form cell index -> one or two 16-byte
// shared-memory loads -> a few FMAs":
//
//   bilinear cell   f = p0 + p1*sf + sa*(p2 + p3*sf)          sa = a - a_i, sf = fi - fi_i
//   PCHIP interval  f = ((c0*s + c1)*s + c2)*s + c3           s  = a - a_i  (unclipped: cubic extrapolation)
//   trilinear cell  f = t0 + t1*m + h*(t2 + t3*m) + p*(t4 + t5*m + h*(t6 + t7*m))
//
// which are algebraically the interpolants SciPy evaluates (RegularGridInterpolator
// linear / interp1d / PPoly), to rounding.  One image per precision lives in global
// memory; each CTA stages it into shared memory with a single TMA bulk copy.
#pragma once
#include <cstdint>

namespace f16 {

constexpr int kNA = 19;   // alpha1 intervals (20 nodes, non-uniform: 5 deg to 60 deg, then 10 deg)
constexpr int kNF1 = 4;   // fi1 intervals (5 nodes)
constexpr int kNF3 = 6;   // fi3 intervals (7 nodes)

// Shared-memory bank layout.  A lookup is an LDS.128 whose address depends on the lane's env: a quarter-warp (8 lanes x
// 16 B) is served in one wavefront only if the 16-byte chunks it touches fall into different bank groups (chunk index
// mod 8) or coincide (broadcast).  Neighbouring cells are what a quarter-warp of similar envs reads, so the strides (in
// fp32 chunks) are padded to spread neighbours over the groups: alpha rows 5 chunks apart (was 4: every second row
// collided), fi planes of the 3-D tables 60 = 4 (mod 8) chunks apart, thrust altitude rows 11 chunks apart.  Measured
// with the stationary state distribution of the bench (tests/studies/bank_conflict_model.py): 17.3 -> 11.3 wavefronts
// per quarter-warp for the 11 loads (11 = conflict-free).
constexpr int kArowStride = 20;     // 16 used + 4 pad
constexpr int kThrustRow = 5 * 8 + 4;   // five Mach cells of 8 + 4 pad

template <typename R>
struct alignas(16) Tables {
    // [ifi][ia] -> {Cx p0..p3, Cy p0..p3, mz p0..p3}          (data/coeff.py: Cx1, Cy1, mz1 at beta=0); row kNA is padding
    R cell3[kNF1][kNA + 1][12];
    // [ia] -> {Cxwz c0..c3, Cywz c0..c3, mzwz c0..c3, dmz1 y0, dmz1 slope, alpha_i, 0, pad x 4}
    R arow[kNA][kArowStride];
    // [ia][ifi3] -> {p0..p3} of dmz_ds1 (axes alpha1 x fi3)
    R dmzds[kNA][kNF3][4];
    // [ifi] -> eta_fi1 PCHIP c0..c3 on fi1
    R eta[kNF1][4];
    // [ip][ih][8 * im + k] -> t0..t7, offsets measured from the cell origin (Pa_i, H_i, M_i)
    R thrust[2][5][kThrustRow];
};

static_assert(sizeof(Tables<float>) % 16 == 0, "TMA bulk copy needs a 16-byte multiple");
static_assert(sizeof(Tables<double>) % 16 == 0, "TMA bulk copy needs a 16-byte multiple");

// Host: build the image from the raw tables in f16_tables_data.h (fp64 arithmetic, rounded once to R).
void build_tables(Tables<double> &out);
void build_tables(Tables<float> &out);

}  // namespace f16
