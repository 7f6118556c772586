// powtab.h -- x^m for a FIXED exponent m by table + short polynomial, with a proven-at-build-time error bound (host + device).
//
// The PQ / sRGB / OKLAB transfer functions of the reference (common.py:34-159, oklab.py:73,94) raise to seven fixed powers in
// float64; the general pow (color.cu: fpow, a restatement of the libm algorithm, ~70 FP64 instructions) is what bounds those
// colour kernels.  For a fixed m,
//     x = 2^e (1 + f),   x^m = 2^(e m) . (1 + f)^m  =  Te[e] . P_j(s)
// with j = the top `lg_nseg` bits of f, s in [-1, 1] the position inside segment j (computed exactly from the bits of x), P_j
// a polynomial of degree `deg` fitted on the host in long double at Chebyshev nodes, Te[e] = 2^(e m) rounded to double.
// ~35 instructions with the degree as a compile-time constant.  The builder measures the largest relative error against powl on random points of every segment and
// stores a bound `eps` (twice the measured maximum plus the roundings of the evaluation); callers carry that bound through
// the few operations that follow, and when the final float32 rounding of a result is closer to a rounding boundary than the
// bound they recompute that pixel with the exact path.  Results are therefore bit-identical to the exact path, always --
// the tables only decide how often the slow path runs (measured on image data: 2 .. 6 pixels in 1000 for the PQ encoders, whose
// chroma outputs sit near zero where float32 is dense; 1e-4 and less elsewhere).
#pragma once
#include <stdint.h>
#include <string.h>
#ifndef __CUDACC__
#define AEAJ_HD
#else
#define AEAJ_HD __host__ __device__ __forceinline__
#endif

struct PowTabView {
    const double* te;      // [nexp]  2^((emin + k) m)
    const double* coef;    // [nseg - jmin][deg + 1], highest degree first (Horner)
    double eps;            // bound on |approx - x^m| / x^m inside the domain
    int emin, nexp;        // exponents covered: x in [2^emin, 2^(emin + nexp))
    int lg_nseg, jmin;     // segments per binade = 1 << lg_nseg; only segments j >= jmin are stored (a domain inside one binade)
    int deg, stride;       // polynomial degree; doubles per segment (deg + 1 rounded up to even: 16-byte aligned pairs)
};

AEAJ_HD uint64_t powtab_bits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
AEAJ_HD double powtab_from_bits(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
AEAJ_HD double powtab_ld(const double* p) {                          // read-only path: the tables are small enough to live in L1
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
AEAJ_HD double powtab_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// x^m for x inside the table's domain; ok = false (and the result is meaningless) outside it or for non-finite / non-positive x.
// DEG is the table's degree as a compile-time constant: the Horner chain is straight-line code and the coefficients come in
// 16-byte pairs (a rolled loop with scalar loads cost ~140 instructions per evaluation; this is ~35).
template <int DEG>
AEAJ_HD double powtab_eval(const PowTabView& T, double x, bool& ok) {
    const uint64_t u = powtab_bits(x);
    const int e = (int)((u >> 52) & 0x7ff) - 1023;                 // sign bit set -> huge e -> rejected below
    const int k = e - T.emin;
    const int j = (int)((u >> (52 - T.lg_nseg)) & ((1u << T.lg_nseg) - 1u));
    ok = ok && ((u >> 63) == 0) && k >= 0 && k < T.nexp && j >= T.jmin && T.deg == DEG;
    if (!ok) return 0.0;
    // s = 2 (f . nseg - j) - 1, exact: the low mantissa bits of x below the segment index, as a double in [1, 2), minus 1.5, times 2
    const uint64_t low = (u << T.lg_nseg) & 0x000fffffffffffffull;   // f . nseg - j in [0, 1) as a 52-bit fraction (low bits shifted out are zero-filled)
    const double frac1 = powtab_from_bits(0x3ff0000000000000ull | low);   // 1 + (f . nseg - j)
    const double s = (frac1 - 1.5) * 2.0;                                 // exact
    const double* c = T.coef + (size_t)(j - T.jmin) * T.stride;
    double cc[DEG + 2];
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int d = 0; d < (DEG + 2) / 2; d++) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(c) + d);
        cc[2 * d] = v.x; cc[2 * d + 1] = v.y;
    }
#else
    for (int d = 0; d <= DEG; d++) cc[d] = c[d];
#endif
    double p = cc[0];
#pragma unroll
    for (int d = 1; d <= DEG; d++) p = powtab_fma(p, s, cc[d]);
    return p * powtab_ld(T.te + k);
}

// A whole transfer curve f(x) (not a pure power) tabulated the same way: one degree-8 polynomial per (binade, segment), no 2^(e m)
// factor.  One evaluation then replaces two powers, a division and the error arithmetic between them.
struct FnTabView {
    const double* coef;    // [nexp << lg_nseg][10]  (9 coefficients, highest degree first, padded to 16-byte pairs)
    double eps;            // bound on |approx - f| / |f| inside the domain
    int emin, nexp, lg_nseg;
};
AEAJ_HD double fntab_eval(const FnTabView& T, double x, bool& ok) {
    const uint64_t u = powtab_bits(x);
    const int k = (int)((u >> 52) & 0x7ff) - 1023 - T.emin;
    ok = ok && ((u >> 63) == 0) && k >= 0 && k < T.nexp;
    if (!ok) return 0.0;
    const int j = (int)((u >> (52 - T.lg_nseg)) & ((1u << T.lg_nseg) - 1u));
    const uint64_t low = (u << T.lg_nseg) & 0x000fffffffffffffull;
    const double s = (powtab_from_bits(0x3ff0000000000000ull | low) - 1.5) * 2.0;          // exact, in [-1, 1)
    const double* c = T.coef + ((size_t)((k << T.lg_nseg) + j)) * 10;
    double cc[10];
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int d = 0; d < 5; d++) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(c) + d);
        cc[2 * d] = v.x; cc[2 * d + 1] = v.y;
    }
#else
    for (int d = 0; d < 9; d++) cc[d] = c[d];
#endif
    double p = cc[0];
#pragma unroll
    for (int d = 1; d <= 8; d++) p = powtab_fma(p, s, cc[d]);
    return p;
}
