// f16_tables.cpp -- host-side construction of the packed table image (see f16_tables.h).
#include "f16_tables.h"

#include <cstring>

#include "f16_tables_data.h"

namespace f16 {

namespace {

// bilinear cell -> polynomial in the offsets from the cell origin
inline void bilinear_poly(double v00, double v01, double v10, double v11, double d0, double d1, double *p)
{
    // axis 0 offset s0 (width d0), axis 1 offset s1 (width d1):
    //   f = v00 + (v10-v00)/d0*s0 + s1*((v01-v00)/d1 + (v11-v10-v01+v00)/(d0*d1)*s0)
    p[0] = v00;
    p[1] = (v10 - v00) / d0;
    p[2] = (v01 - v00) / d1;
    p[3] = ((v11 - v10) - (v01 - v00)) / (d0 * d1);
}

}  // namespace

void build_tables(Tables<double> &T)
{
    std::memset(&T, 0, sizeof(T));
    // 3-D tables at beta = 0: axis 0 = fi1 (offset sf), axis 1 = alpha1 (offset sa)
    const double(*src3[3])[20] = {F16T_CX0, F16T_CY0, F16T_MZ0};
    for (int f = 0; f < kNF1; ++f)
        for (int a = 0; a < kNA; ++a) {
            const double df = F16T_FI1[f + 1] - F16T_FI1[f];
            const double da = F16T_ALPHA1[a + 1] - F16T_ALPHA1[a];
            for (int t = 0; t < 3; ++t)
                bilinear_poly(src3[t][f][a], src3[t][f][a + 1], src3[t][f + 1][a], src3[t][f + 1][a + 1], df, da,
                              &T.cell3[f][a][4 * t]);
        }
    // per-alpha-interval row
    const double(*pp[3])[19] = {F16T_CXWZ_PP, F16T_CYWZ_PP, F16T_MZWZ_PP};
    for (int a = 0; a < kNA; ++a) {
        for (int t = 0; t < 3; ++t)
            for (int k = 0; k < 4; ++k) T.arow[a][4 * t + k] = pp[t][k][a];
        const double da = F16T_ALPHA1[a + 1] - F16T_ALPHA1[a];
        T.arow[a][12] = F16T_DMZ1[a];
        T.arow[a][13] = (F16T_DMZ1[a + 1] - F16T_DMZ1[a]) / da;
        T.arow[a][14] = F16T_ALPHA1[a];
        T.arow[a][15] = 0.0;
    }
    // dmz_ds1: axis 0 = alpha1 (offset sa), axis 1 = fi3 (offset sf3).  Stored so that
    //   f = p0 + p1*sf3 + sa*(p2 + p3*sf3)
    for (int a = 0; a < kNA; ++a)
        for (int f = 0; f < kNF3; ++f) {
            const double da = F16T_ALPHA1[a + 1] - F16T_ALPHA1[a];
            const double df = F16T_FI3[f + 1] - F16T_FI3[f];
            double p[4];
            // treat fi3 as "axis 0" (offset sf3) and alpha as "axis 1" (offset sa)
            bilinear_poly(F16T_DMZ_DS[a][f], F16T_DMZ_DS[a + 1][f], F16T_DMZ_DS[a][f + 1], F16T_DMZ_DS[a + 1][f + 1],
                          df, da, p);
            for (int k = 0; k < 4; ++k) T.dmzds[a][f][k] = p[k];
        }
    for (int f = 0; f < kNF1; ++f)
        for (int k = 0; k < 4; ++k) T.eta[f][k] = F16T_ETA_FI_PP[k][f];
    // thrust cube, axes (Pa, H, M)
    for (int ip = 0; ip < 2; ++ip)
        for (int ih = 0; ih < 5; ++ih)
            for (int im = 0; im < 5; ++im) {
                const double dp = F16T_THRUST_PA[ip + 1] - F16T_THRUST_PA[ip];
                const double dh = F16T_THRUST_H[ih + 1] - F16T_THRUST_H[ih];
                const double dm = F16T_THRUST_M[im + 1] - F16T_THRUST_M[im];
                double lo[4], hi[4];  // bilinear in (h, m) on the two Pa planes: axis0 = M, axis1 = H
                const double(*P0)[6] = F16T_THRUST[ip];
                const double(*P1)[6] = F16T_THRUST[ip + 1];
                bilinear_poly(P0[ih][im], P0[ih + 1][im], P0[ih][im + 1], P0[ih + 1][im + 1], dm, dh, lo);
                bilinear_poly(P1[ih][im], P1[ih + 1][im], P1[ih][im + 1], P1[ih + 1][im + 1], dm, dh, hi);
                double *t = T.thrust[ip][ih] + 8 * im;
                for (int k = 0; k < 4; ++k) {
                    t[k] = lo[k];                       // t0 + t1*m + h*(t2 + t3*m)
                    t[4 + k] = (hi[k] - lo[k]) / dp;    // + p*(t4 + t5*m + h*(t6 + t7*m))
                }
            }
}

void build_tables(Tables<float> &out)
{
    static_assert(sizeof(Tables<double>) == 2 * sizeof(Tables<float>), "same shape in both precisions");
    Tables<double> T;
    build_tables(T);
    const double *s = reinterpret_cast<const double *>(&T);
    float *d = reinterpret_cast<float *>(&out);
    for (size_t i = 0; i < sizeof(Tables<float>) / sizeof(float); ++i) d[i] = static_cast<float>(s[i]);
}

}  // namespace f16
