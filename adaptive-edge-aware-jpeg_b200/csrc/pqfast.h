// pqfast.h -- the transfer-function chains of the PQ / sRGB / OKLAB colour spaces on top of powtab.h, each returning its value
// together with a bound on the relative distance to what the exact float64 path computes (host + device).
//
// "Exact path" = the reference's float64 arithmetic (common.py:94-159) with a pow of at most 1 ulp (fpow in color.cu, libm on
// the host).  A bound `rel` returned here covers BOTH the table error and the exact path's own roundings, so
//      | fast - exact |  <=  rel * |value|
// and the caller may keep (float)fast whenever every double within that distance rounds to the same float32 (round_is_safe).
#pragma once
#include "powtab.h"

#define PQF_U 2.220446049250313e-16          /* 2^-52: bound on the relative error of one exact-path operation (pow: < 1 ulp) */

struct PqTabs {
    PowTabView m1;          // x^(2610/16384), x = c / 10000
    PowTabView m2[2];       // r^m2, r in [c1, 1]:   [0] m2 = 2523/32 (ICtCp, ICaCb),  [1] 1.7 * 2523/32 (JzAzBz)
    PowTabView im2[2];      // y^(1/m2)
    PowTabView im1;         // w^(16384/2610)
    PowTabView isrgb;       // d^(1/2.4)
    PowTabView cbrt32;      // l^((double)(float)(1/3))   (oklab.py:73: numpy casts the exponent to float32)
    PowTabView cube;        // |l'|^3
    // the whole PQ curves, one evaluation each (outside their domains the two-stage evaluation above is used, then the exact path)
    FnTabView enc[2];       // c -> ((c1 + c2 t) / (1 + c3 t))^m2, t = (c / 10000)^m1
    FnTabView dec[2];       // y -> 10000 ((t - c1) / (c2 - c3 t))^(1/m1), t = y^(1/m2)
    double enc_rel[2], dec_rel[2];   // table bound + the exact path's own rounding noise over the domain (constants, pqtabs_build.h)
};

AEAJ_HD double pqf_abs(double x) { return x < 0.0 ? -x : x; }

// 1 / a to within 2 ulp for a normal a > 0, without the IEEE division sequence: the fast path needs bounded errors, not the exact
// path's bits (device: MUFU.RCP64H seed + two Newton steps, 5 instructions instead of ~20)
AEAJ_HD double pqf_rcp(double a) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = __fma_rn(-a, y, 1.0);
    y = __fma_rn(y, e, y);
    e = __fma_rn(-a, y, 1.0);
    return __fma_rn(y, e, y);
#else
    return 1.0 / a;
#endif
}

// every double within E of v rounds to the same float32  (rounding is monotone, so the two ends decide)
AEAJ_HD bool round_is_safe(double v, double E) { return (float)(v + E) == (float)(v - E); }

// common.py:131-159 -- ((c1 + c2 t) / (1 + c3 t))^m2,  t = (c / 10000)^m1
AEAJ_HD double pqf_inv_eotf(const PqTabs& Q, int which, double m2, double c, double& rel, bool& ok) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0;
    {
        bool okc = ok;
        const double f = fntab_eval(Q.enc[which], c, okc);
        if (okc) { rel = Q.enc_rel[which]; return f; }
    }
    const double x = c * 1e-4;                                        // 1 ulp from the exact path's c / 10000: covered by the bound below
    const bool zero = (c == 0.0);                                     // black: 0^m1 = 0 exactly in both paths
    bool okt = ok;
    double t = powtab_eval<8>(Q.m1, x, okt);
    ok = zero ? ok : okt;
    t = zero ? 0.0 : t;
    const double et = zero ? 0.0 : Q.m1.eps + 2.0 * PQF_U;
    const double ct = c2 * t, dt = c3 * t;
    const double num = c1 + ct, den = 1.0 + dt;
    const double rnum = pqf_rcp(num), rden = pqf_rcp(den);
    const double r = num * rden;
    // d ln r = (ct/num - dt/den) d ln t: the two terms nearly cancel (0.03 .. 0.1 for image values); plus the roundings of both paths
    const double er = (pqf_abs(ct * rnum - dt * rden) + 1e-3) * et + 8.0 * PQF_U;
    const double f = powtab_eval<12>(Q.m2[which], r, ok);
    rel = m2 * er + Q.m2[which].eps + PQF_U;
    return f;
}

// common.py:94-129 -- 10000 ((t - c1) / (c2 - c3 t))^(1/m1),  t = y^(1/m2);  negative numerator -> 0, non-positive denominator -> 1e-12
AEAJ_HD double pqf_eotf(const PqTabs& Q, int which, double y, double& rel, bool& ok) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0, m1 = 2610.0 / 16384.0;
    {
        bool okc = ok;
        const double f = fntab_eval(Q.dec[which], y, okc);
        if (okc) { rel = Q.dec_rel[which]; return f; }
    }
    const double t = powtab_eval<8>(Q.im2[which], y, ok);
    const double et = (Q.im2[which].eps + PQF_U) * t;                // absolute
    const double num = t - c1, den = c2 - c3 * t;
    rel = 0.0;
    if (!ok) return 0.0;
    if (num < -2.0 * et) return 0.0;                                  // both paths clamp the numerator to 0: the result is exactly 0
    ok = ok && num > 4.0 * et && den > 4.0 * c3 * et + 1e-9;          // too close to one of the two clamps to know which side the exact path takes
    if (!ok) return 0.0;
    const double rnum = pqf_rcp(num), rden = pqf_rcp(den);
    const double w = num * rden;
    const double ew = et * rnum * 1.000001 + c3 * et * rden * 1.000001 + 8.0 * PQF_U;   // relative
    const double p = powtab_eval<8>(Q.im1, w, ok);
    rel = ew / m1 + Q.im1.eps + 2.0 * PQF_U;
    return 10000.0 * p;
}

// common.py:62-92 without the final clamp: v <= 0.0031308 ? 12.92 v : 1.055 v^(1/2.4) - 0.055;  E = absolute bound
AEAJ_HD double pqf_linear_to_srgb(const PqTabs& Q, double d, double& E, bool& ok) {
    if (d <= 0.0031308) { E = 0.0; ok = ok && (d == d); return d * 12.92; }   // the same single multiplication as the exact path
    const double p = powtab_eval<8>(Q.isrgb, d, ok);
    const double a = 1.055 * p;
    E = a * (Q.isrgb.eps + 3.0 * PQF_U);
    return a - 0.055;
}
