// pqtabs_build.h -- host side: the nine fixed-exponent tables of pqfast.h, built once per handle (long double on the host).
#pragma once
#include "pqfast.h"
#include <math.h>
#include <stdlib.h>
#include <vector>
// Host-side builder (long double = x87 extended on the x86-64 hosts this runs on: 64-bit mantissa, errors ~2^-63).
struct PowTabHost {
    std::vector<double> te, coef;
    PowTabView v;
    double measured;       // largest relative error seen on the validation points
};
// Chebyshev-node interpolation of (1 + (j + (1 + s) / 2) / nseg)^m on s in [-1, 1], converted to monomials in s
inline void powtab_build(PowTabHost& H, double m, int emin, int nexp, int lg_nseg, int deg, double fmin = 0.0) {
    if (deg != 8 && deg != 12) abort();                       // the evaluators are instantiated for these two degrees (pqfast.h)
    const int nseg = 1 << lg_nseg, jmin = (int)floor(fmin * nseg);
    H.te.resize(nexp);
    for (int k = 0; k < nexp; k++) H.te[k] = (double)powl(2.0L, (long double)(emin + k) * (long double)m);
    const int n = deg + 1, stride = (n + 1) & ~1;
    std::vector<long double> node(n), V((size_t)n * n), y(n), a(n);
    for (int i = 0; i < n; i++) node[i] = cosl(3.14159265358979323846264338327950288L * (2 * i + 1) / (2.0L * n));
    H.coef.assign((size_t)(nseg - jmin) * stride, 0.0);
    for (int j = jmin; j < nseg; j++) {
        for (int i = 0; i < n; i++) {
            y[i] = powl(1.0L + ((long double)j + (1.0L + node[i]) / 2.0L) / nseg, (long double)m);
            long double pw = 1.0L;
            for (int d = 0; d < n; d++) { V[(size_t)i * n + d] = pw; pw *= node[i]; }
        }
        // solve V a = y (Gaussian elimination with partial pivoting, n <= 9)
        std::vector<long double> M(V);
        std::vector<long double> b(y);
        for (int c = 0; c < n; c++) {
            int piv = c;
            for (int r = c + 1; r < n; r++) if (fabsl(M[(size_t)r * n + c]) > fabsl(M[(size_t)piv * n + c])) piv = r;
            for (int d = 0; d < n; d++) std::swap(M[(size_t)c * n + d], M[(size_t)piv * n + d]);
            std::swap(b[c], b[piv]);
            for (int r = c + 1; r < n; r++) {
                const long double f = M[(size_t)r * n + c] / M[(size_t)c * n + c];
                for (int d = c; d < n; d++) M[(size_t)r * n + d] -= f * M[(size_t)c * n + d];
                b[r] -= f * b[c];
            }
        }
        for (int c = n - 1; c >= 0; c--) {
            long double t = b[c];
            for (int d = c + 1; d < n; d++) t -= M[(size_t)c * n + d] * a[d];
            a[c] = t / M[(size_t)c * n + c];
        }
        for (int d = 0; d < n; d++) H.coef[(size_t)(j - jmin) * stride + d] = (double)a[deg - d];   // highest degree first
    }
    H.v.te = H.te.data(); H.v.coef = H.coef.data(); H.v.emin = emin; H.v.nexp = nexp; H.v.lg_nseg = lg_nseg; H.v.jmin = jmin; H.v.deg = deg; H.v.stride = stride;
    H.v.eps = 0.0;
    // validation: segment ends, Chebyshev extrema and random points of every segment, three exponents
    double worst = 0.0;
    uint64_t rng = 0x9e3779b97f4a7c15ull;
    for (int j = jmin; j < nseg; j++)
        for (int t = 0; t < 12; t++) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            long double pos = (t == 0) ? 0.0L : (t == 1) ? 0.99999999L : (long double)((rng >> 11) & 0xfffffffffffffull) / 4503599627370496.0L;
            const int kk = (t % 3 == 0) ? 0 : (t % 3 == 1) ? nexp - 1 : nexp / 2;
            const double x = ldexp(1.0 + (double)(((long double)j + pos) / nseg), emin + kk);
            bool ok = true;
            const double got = (deg == 12) ? powtab_eval<12>(H.v, x, ok) : powtab_eval<8>(H.v, x, ok);
            const long double want = powl((long double)x, (long double)m);
            if (!ok) continue;
            const double rel = (double)fabsl(((long double)got - want) / want);
            if (rel > worst) worst = rel;
        }
    H.measured = worst;
    H.v.eps = 2.0 * worst + 4.0 * 1.1102230246251565e-16;
}


struct FnTabHost {
    std::vector<double> coef;
    FnTabView v;
    double measured;
};
// f fitted per (binade, segment) at 9 Chebyshev nodes (long double), validated on random points like powtab_build
template <class F>
inline void fntab_build(FnTabHost& H, F f, int emin, int nexp, int lg_nseg) {
    const int n = 9, nseg = 1 << lg_nseg;
    std::vector<long double> node(n), y(n), a(n);
    for (int i = 0; i < n; i++) node[i] = cosl(3.14159265358979323846264338327950288L * (2 * i + 1) / (2.0L * n));
    H.coef.assign((size_t)nexp * nseg * 10, 0.0);
    for (int k = 0; k < nexp; k++)
        for (int j = 0; j < nseg; j++) {
            std::vector<long double> M((size_t)n * n), b(n);
            for (int i = 0; i < n; i++) {
                b[i] = f(ldexpl(1.0L + ((long double)j + (1.0L + node[i]) / 2.0L) / nseg, emin + k));
                long double pw = 1.0L;
                for (int d = 0; d < n; d++) { M[(size_t)i * n + d] = pw; pw *= node[i]; }
            }
            for (int c = 0; c < n; c++) {
                int piv = c;
                for (int r = c + 1; r < n; r++) if (fabsl(M[(size_t)r * n + c]) > fabsl(M[(size_t)piv * n + c])) piv = r;
                for (int d = 0; d < n; d++) std::swap(M[(size_t)c * n + d], M[(size_t)piv * n + d]);
                std::swap(b[c], b[piv]);
                for (int r = c + 1; r < n; r++) {
                    const long double q = M[(size_t)r * n + c] / M[(size_t)c * n + c];
                    for (int d = c; d < n; d++) M[(size_t)r * n + d] -= q * M[(size_t)c * n + d];
                    b[r] -= q * b[c];
                }
            }
            for (int c = n - 1; c >= 0; c--) {
                long double t = b[c];
                for (int d = c + 1; d < n; d++) t -= M[(size_t)c * n + d] * a[d];
                a[c] = t / M[(size_t)c * n + c];
            }
            for (int d = 0; d < n; d++) H.coef[((size_t)k * nseg + j) * 10 + d] = (double)a[8 - d];
        }
    H.v.coef = H.coef.data(); H.v.emin = emin; H.v.nexp = nexp; H.v.lg_nseg = lg_nseg; H.v.eps = 0.0;
    double worst = 0.0;
    uint64_t rng = 0x2545f4914f6cdd1dull;
    for (int k = 0; k < nexp; k++)
        for (int j = 0; j < nseg; j++)
            for (int t = 0; t < 6; t++) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                const long double pos = (t == 0) ? 0.0L : (t == 1) ? 0.99999999L : (long double)((rng >> 11) & 0xfffffffffffffull) / 4503599627370496.0L;
                const double x = ldexp(1.0 + (double)(((long double)j + pos) / nseg), emin + k);
                bool ok = true;
                const double got = fntab_eval(H.v, x, ok);
                const long double want = f((long double)x);
                const double rel = (double)fabsl(((long double)got - want) / want);
                if (ok && rel > worst) worst = rel;
            }
    H.measured = worst;
    H.v.eps = 2.0 * worst + 4.0 * 1.1102230246251565e-16;
}

struct PqTabsHost {
    PowTabHost m1, m2[2], im2[2], im1, isrgb, cbrt32, cube;
    FnTabHost enc[2], dec[2];
    PqTabs view;                                      // views over the host vectors
};

// Domains: whatever 8-bit sources and their decoded neighbourhood produce; anything outside takes the exact path.
// Few segments, high degree: a DFMA is cheap next to the exact path's ~70 FP64 instructions per power, while the gathers of the
// coefficients must hit L1 -- all nine tables together are 30 KB (the first version, 4096 segments of degree 6 for m2, was 150 KB,
// lived in L2, and made the "fast" path slower than the exact one).
//   m1:     x = LMS / 10000, LMS in (2^-24 .. 4)          -> x in [2^-38, 2^-11)
//   m2:     r = (c1 + c2 t)/(1 + c3 t) in [c1 = 0.836, 1)  -> one binade [0.5, 1), segments from f = 0.67 on; steep (m2 up to 134)
//   1/m2:   y = PQ-encoded L'M'S' in [2^-40, 2)
//   1/m1:   w = (t - c1)/(c2 - c3 t) in [2^-64, 2)
//   1/2.4:  linear RGB in [2^-9, 4)                        (below 0.0031308 the encode is linear)
//   f32(1/3), cube: OKLAB's LMS in [2^-24, 4), L'M'S' in [2^-10, 4)
inline void pqtabs_build(PqTabsHost& H) {
    powtab_build(H.m1, 2610.0 / 16384.0, -38, 27, 5, 8);
    powtab_build(H.m2[0], 2523.0 / 32.0, -1, 1, 8, 12, 0.67);
    powtab_build(H.m2[1], 1.7 * 2523.0 / 32.0, -1, 1, 8, 12, 0.67);
    powtab_build(H.im2[0], 32.0 / 2523.0, -40, 41, 5, 8);
    powtab_build(H.im2[1], 1.0 / (1.7 * 2523.0 / 32.0), -40, 41, 5, 8);
    powtab_build(H.im1, 16384.0 / 2610.0, -64, 65, 6, 8);
    powtab_build(H.isrgb, 1.0 / 2.4, -9, 11, 5, 8);
    powtab_build(H.cbrt32, (double)(float)(1.0 / 3.0), -24, 26, 5, 8);
    powtab_build(H.cube, 3.0, -10, 12, 5, 8);
    // the whole curves.  Their constant bounds add the exact path's own rounding noise (u = 2^-52 per operation, pow <= 1 ulp):
    //   encode: t carries 1.1 u, r = num / den 3.9 u at most, amplified by m2;   decode over the tabulated domain: w = num / den carries
    //   at most 31 u (cancellation in t - c1 at the low end of the domain), amplified by 1 / m1 = 6.28
    const long double c1 = 3424.0L / 4096.0L, c2 = 2413.0L / 128.0L, c3 = 2392.0L / 128.0L, m1 = 2610.0L / 16384.0L;
    for (int w = 0; w < 2; w++) {
        const long double m2 = w ? 1.7L * 2523.0L / 32.0L : 2523.0L / 32.0L;
        fntab_build(H.enc[w], [=](long double c) { const long double t = powl(c / 10000.0L, m1); return powl((c1 + c2 * t) / (1.0L + c3 * t), m2); },
                    -24, 26, 5);
        fntab_build(H.dec[w], [=](long double y) { const long double t = powl(y, 1.0L / m2); return 10000.0L * powl((t - c1) / (c2 - c3 * t), 1.0L / m1); },
                    w ? -24 : -12, w ? 24 : 12, 6);          // y < 1: the curve has a pole at y = (c2 / c3)^m2 > 1
        H.view.enc[w] = H.enc[w].v; H.view.dec[w] = H.dec[w].v;
        H.view.enc_rel[w] = H.enc[w].v.eps + (3.9 * (double)m2 + 1.0) * PQF_U;
        H.view.dec_rel[w] = H.dec[w].v.eps + (46.0 / (double)m1 + 2.0) * PQF_U;     // 31 u by the analysis, 1.5 x margin
    }
    H.view.m1 = H.m1.v; H.view.m2[0] = H.m2[0].v; H.view.m2[1] = H.m2[1].v; H.view.im2[0] = H.im2[0].v; H.view.im2[1] = H.im2[1].v;
    H.view.im1 = H.im1.v; H.view.isrgb = H.isrgb.v; H.view.cbrt32 = H.cbrt32.v; H.view.cube = H.cube.v;
}
