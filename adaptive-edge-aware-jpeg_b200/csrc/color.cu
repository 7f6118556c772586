// color.cu -- colour transforms (src/color/*.py), chroma resampling (jpeg.py:323-354), u8 cast
// (edge_detection.py:70) for sm_100a.  HBM-bound elementwise work: one fused, coalesced,
// float4-vectorised kernel per direction.
//
// Arithmetic contract (validated against the reference by the CPU oracle):
//   * np.dot((N,3),M.T) == fma(x2,m2, fma(x1,m1, x0*m0)) in f32  -> dot3()
//   * sRGB transfer functions and PQ are evaluated in f64 and stored as f32 (numba, Python-float consts)
//   * INTER_AREA 2x2: ((a+b)+(c+d))*0.25f ; 1x4: (((a+b)+c)+d)*0.25f ; other ratios: OpenCV's table path
//   * INTER_LINEAR: half-pixel centres, edge clamp, horizontal then vertical, a*(1-t)+b*t, no fma
// The file is compiled with -fmad=false; every fused multiply-add below is explicit.
#include "aeaj_internal.cuh"
#include "pqtabs_build.h"

namespace {

__device__ __forceinline__ void dot3(const float* m, float x0, float x1, float x2, float& o0, float& o1, float& o2) {
    o0 = __fmaf_rn(x2, m[2], __fmaf_rn(x1, m[1], __fmul_rn(x0, m[0])));
    o1 = __fmaf_rn(x2, m[5], __fmaf_rn(x1, m[4], __fmul_rn(x0, m[3])));
    o2 = __fmaf_rn(x2, m[8], __fmaf_rn(x1, m[7], __fmul_rn(x0, m[6])));
}

__constant__ float c_rgb2xyz[9] = {0.4124564f, 0.3575761f, 0.1804375f, 0.2126729f, 0.7151522f, 0.0721750f,
                                   0.0193339f, 0.1191920f, 0.9503041f};               // xyz.py:27-32
__constant__ float c_xyz2rgb[9] = {3.2404542f, -1.5371385f, -0.4985314f, -0.9692660f, 1.8760108f, 0.0415560f,
                                   0.0556434f, -0.2040259f, 1.0572252f};              // xyz.py:35-40

// ---------------------------------------------------------------------------------------------
// pow(double, double) for the transfer functions.  The reference evaluates them in float64 through numba / libm
// (common.py, jzazbz.py); CUDA's generic pow() costs ~200 instructions and made the PQ spaces 3x slower than YCbCr.
// This is the table-driven algorithm of modern libms restated for positive finite bases: log(x) = k ln2 + log(c_i) +
// log1p(x / c_i - 1) with a 128-entry table whose 1/c_i have 8 fractional bits (r = fma(z, 1/c, -1) is exact or within
// 2^-62), the sum kept as hi + lo; y log(x) in double-double; exp through a 128-entry table of 2^(j/128) and a degree-5
// polynomial.  Tables: pow_tables.inc, derived from first principles by tools/gen_pow_tables.py (mpmath).  Checked on
// the CPU against glibc's pow: 1.1e-3 of results differ by 1 ulp (double), none of 2e8 after rounding to float, and the
// whole colour chains built on it differ from the libm-based oracle in ~1 of 24 M float outputs (the CUDA pow did in 2e-5).
// Everything else (zero, negative, subnormal, inf, NaN, results near the double range limits) takes CUDA's pow.
// ---------------------------------------------------------------------------------------------
#include "pow_tables.inc"
__device__ const double d_pow_log_tab[128][3] = POW_LOG_TAB_INIT;
__device__ const double d_pow_exp_tab[128][2] = POW_EXP_TAB_INIT;
struct PowTabs { const double (*lg)[3]; const double (*ex)[2]; };      // shared-memory copies (data-dependent indices)
constexpr int POW_TAB_DOUBLES = 128 * 3 + 128 * 2;

__device__ __forceinline__ PowTabs load_pow_tabs(double* smem) {
    const int nthr = blockDim.x * blockDim.y, t = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = t; i < 128 * 3; i += nthr) smem[i] = (&d_pow_log_tab[0][0])[i];
    for (int i = t; i < 128 * 2; i += nthr) smem[128 * 3 + i] = (&d_pow_exp_tab[0][0])[i];
    __syncthreads();
    PowTabs T;
    T.lg = reinterpret_cast<const double (*)[3]>(smem);
    T.ex = reinterpret_cast<const double (*)[2]>(smem + 128 * 3);
    return T;
}

__device__ __noinline__ double pow_generic(double x, double y) { return pow(x, y); }

// y > 0 at every call site (1/3 as float, 3, 2.4, 1/2.4, the PQ exponents and their reciprocals)
__device__ __forceinline__ double fpow(const PowTabs& T, double xs, double y) {
    // black pixels and slightly negative L'M'S' values are common enough to keep them off the slow path:
    // pow(+-0, y > 0) = 0 (+-0 for the odd integer 3), pow(x < 0, 3) = -pow(-x, 3), pow(x < 0, non-integer) = NaN (class T-NAN)
    if (xs == 0.0) return (y == 3.0) ? xs : 0.0;
    const bool neg = xs < 0.0;
    if (neg && y != 3.0) return __longlong_as_double(0x7ff8000000000000ll);
    const double x = fabs(xs);
    if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return pow_generic(xs, y);   // subnormal, inf, NaN
    const unsigned long long ix = (unsigned long long)__double_as_longlong(x);
    const unsigned long long tmp = ix - 0x3fe6955500000000ull;
    const int i = (int)((tmp >> 45) & 127);
    const long long k = (long long)tmp >> 52;
    const double z = __longlong_as_double((long long)(ix - (tmp & 0xfff0000000000000ull)));
    const double kd = (double)k;
    const double invc = T.lg[i][0], logc = T.lg[i][1], logctail = T.lg[i][2];
    const double r = __fma_rn(z, invc, -1.0);
    const double t1 = __dadd_rn(__dmul_rn(kd, POW_LN2HI), logc);
    const double t2 = __dadd_rn(t1, r);
    const double lo1 = __dadd_rn(__dmul_rn(kd, POW_LN2LO), logctail);
    const double lo2 = __dadd_rn(__dsub_rn(t1, t2), r);
    const double ar = __dmul_rn(-0.5, r), ar2 = __dmul_rn(r, ar), ar3 = __dmul_rn(r, ar2);
    const double hi = __dadd_rn(t2, ar2);
    const double lo3 = __fma_rn(ar, r, -ar2);
    const double lo4 = __dadd_rn(__dsub_rn(t2, hi), ar2);
    // p = ar3 * (A1 + r A2 + ar2 (A3 + r A4 + ar2 (A5 + r A6))): the Taylor coefficients of log1p scaled by the powers of -1/2 in ar2, ar3
    const double p5 = __dadd_rn(-0x1.2492492492492p+0, __dmul_rn(r, 1.0));
    const double p3 = __dadd_rn(__dadd_rn(0x1.999999999999ap-1, __dmul_rn(r, -0x1.5555555555555p-1)), __dmul_rn(ar2, p5));
    const double p1 = __dadd_rn(__dadd_rn(-0x1.5555555555555p-1, __dmul_rn(r, 0.5)), __dmul_rn(ar2, p3));
    const double p = __dmul_rn(ar3, p1);
    const double lo = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(lo1, lo2), lo3), lo4), p);
    const double lhi = __dadd_rn(hi, lo), llo = __dadd_rn(__dsub_rn(hi, lhi), lo);
    const double ehi = __dmul_rn(y, lhi), elo = __dadd_rn(__dmul_rn(y, llo), __fma_rn(y, lhi, -ehi));
    if (!(fabs(ehi) < 700.0)) return pow_generic(xs, y);
    // exp(ehi + elo)
    const double kd2 = rint(__dmul_rn(POW_INVLN2N, ehi));
    const long long ki = (long long)kd2;
    double r2 = __dadd_rn(__dadd_rn(ehi, __dmul_rn(kd2, POW_NEGLN2HIN)), __dmul_rn(kd2, POW_NEGLN2LON));
    r2 = __dadd_rn(r2, elo);
    const int idx = (int)(ki & 127);
    const long long top = ki >> 7;
    const double tail = T.ex[idx][0];
    const double scale = __longlong_as_double(__double_as_longlong(T.ex[idx][1]) + (top << 52));
    const double rr = __dmul_rn(r2, r2);
    const double q1 = __dmul_rn(rr, __dadd_rn(0.5, __dmul_rn(r2, 0x1.5555555555555p-3)));
    const double q2 = __dmul_rn(__dmul_rn(rr, rr), __dadd_rn(0x1.5555555555555p-5, __dmul_rn(r2, 0x1.1111111111111p-7)));
    const double t = __dadd_rn(__dadd_rn(__dadd_rn(tail, r2), q1), q2);
    const double res = __dadd_rn(scale, __dmul_rn(scale, t));
    return neg ? -res : res;
}

// common.py:34-60.  lut (shared memory, 256 entries computed by the host libm) is exact for the
// 8-bit-sourced inputs Image.load produces (image.py:80); anything else takes the f64 path.
__device__ __forceinline__ float srgb_to_linear(const PowTabs& T, float v, const float* lut) {
    if (lut) {
        float k = rintf(__fmul_rn(v, 255.0f));
        if (k >= 0.0f && k <= 255.0f && __fdiv_rn(k, 255.0f) == v) return lut[(int)k];
    }
    double d = (double)v;
    if (d <= 0.04045) return (float)(d / 12.92);
    return (float)fpow(T, (d + 0.055) / 1.055, 2.4);
}
// common.py:62-92.  NaN -> 1.0 exactly like the reference's fastmath select chain (class T-NAN).
__device__ __forceinline__ float linear_to_srgb(const PowTabs& T, float v) {
    double d = (double)v, r;
    if (d <= 0.0031308) r = d * 12.92;
    else r = 1.055 * fpow(T, d, 1.0 / 2.4) - 0.055;
    float f = (float)r;
    f = (f < 1.0f) ? f : 1.0f;
    f = (f > 0.0f) ? f : 0.0f;
    return f;
}
// common.py:131-159
__device__ __forceinline__ double pq_inv_eotf(const PowTabs& T, double c, double m2) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0, m1 = 2610.0 / 16384.0;
    double t = fpow(T, c / 10000.0, m1);
    return fpow(T, (c1 + c2 * t) / (1.0 + c3 * t), m2);
}
// common.py:94-129
__device__ __forceinline__ double pq_eotf(const PowTabs& T, double c, double m2) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0, m1 = 2610.0 / 16384.0;
    double t = fpow(T, c, 1.0 / m2);
    double num = t - c1, den = c2 - c3 * t;
    if (num < 0.0) num = 0.0;
    if (den <= 0.0) den = 1e-12;
    return 10000.0 * fpow(T, num / den, 1.0 / m1);
}
#define PQ_M2 (2523.0 / 32.0)
#define JZ_P (1.7 * 2523.0 / 32.0)
#define JZ_B 1.15
#define JZ_G 0.66
#define JZ_D (-0.56)
#define JZ_D0 1.6295499532821566e-11

// no-contraction f64 helpers (nvcc would otherwise be free to fuse under -fmad=true; we compile with
// -fmad=false, these just make the intent explicit)
__device__ __forceinline__ double mul3add(double a0, double b0, double a1, double b1, double a2, double b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}


// ---------------------------------------------------------------------------------------------
// Fast transfer functions (powtab.h / pqfast.h): fixed-exponent powers by table + polynomial (~12 FP64 instructions instead of
// ~70) with a bound on the distance to the exact float64 path carried to every float32 rounding.  Each helper returns true
// only if ALL of its float32 results are provably the exact path's (every double within the bound rounds to the same
// float); otherwise the caller recomputes the pixel with the exact code below it.  Bit-identical by construction; the tables
// only decide how often the slow path runs (a few pixels in 10^5; more around the PQ decoder's clamp at black).
// ---------------------------------------------------------------------------------------------
#ifdef AEAJ_FAST_STATS
__device__ unsigned long long g_fast_stats[8];       // [0] forward pixels, [1] forward fallbacks, [2] inverse pixels, [3] inverse XYZ fallbacks, [4] sRGB fallbacks
#define FAST_STAT(i) atomicAdd(&g_fast_stats[i], 1ull)
#define FAST_STAT_INV(i) (void)atomicAdd(&g_fast_stats[i], 1ull)
extern "C" AEAJ_API int aeaj_debug_fast_stats(unsigned long long* out8, int reset) {
    if (out8) cudaMemcpyFromSymbol(out8, g_fast_stats, sizeof g_fast_stats);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_fast_stats, z, sizeof z); }
    return 0;
}
#else
#define FAST_STAT(i) do { } while (0)
#define FAST_STAT_INV(i) (void)0
#endif
__device__ __forceinline__ bool row3_fast(const float* q, double p0, double p1, double p2, double r0, double r1, double r2, double& o, double& E) {
    const double a0 = fabs((double)q[0]) * p0, a1 = fabs((double)q[1]) * p1, a2 = fabs((double)q[2]) * p2;   // p >= 0
    o = __dadd_rn(__dadd_rn(__dmul_rn((double)q[0], p0), __dmul_rn((double)q[1], p1)), __dmul_rn((double)q[2], p2));   // = mul3add
    E = a0 * r0 + a1 * r1 + a2 * r2 + 8.0 * PQF_U * (a0 + a1 + a2);
    return round_is_safe(o, E);
}

template <int SPACE>
__device__ __forceinline__ bool color_fwd_fast(const ColorConsts& C, float X, float Y, float Z, float& o0, float& o1, float& o2) {
    const PqTabs& Q = C.pq;
    bool ok = true;
    if (SPACE == AEAJ_OKLAB) {                                   // (float) l^((double)(float)(1/3)), oklab.py:71-75
        float v[3], w[3];
        dot3(C.fwd1, X, Y, Z, v[0], v[1], v[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (v[k] == 0.0f) { w[k] = 0.0f; continue; }
            const double p = powtab_eval<8>(Q.cbrt32, (double)v[k], ok);
            if (!ok || !round_is_safe(p, p * (Q.cbrt32.eps + 2.0 * PQF_U))) return false;
            w[k] = (float)p;
        }
        dot3(C.fwd2, w[0], w[1], w[2], o0, o1, o2);
        return true;
    }
    double p[3], r[3];
    if (SPACE == AEAJ_ICACB || SPACE == AEAJ_ICTCP) {
        float l, m, s;
        dot3(C.fwd1, X, Y, Z, l, m, s);
        p[0] = pqf_inv_eotf(Q, 0, PQ_M2, (double)l, r[0], ok);
        p[1] = pqf_inv_eotf(Q, 0, PQ_M2, (double)m, r[1], ok);
        p[2] = pqf_inv_eotf(Q, 0, PQ_M2, (double)s, r[2], ok);
        if (!ok) return false;
        double o, E;
        if (!row3_fast(C.fwd2, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
        o0 = (float)o;
        if (!row3_fast(C.fwd2 + 3, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
        o1 = (float)o;
        if (!row3_fast(C.fwd2 + 6, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
        o2 = (float)o;
        return true;
    }
    // JzAzBz (jzazbz.py:57-97): the operations before and after the three PQ encodes are the exact path's own
    const double Xd = (double)X, Yd = (double)Y;
    const double Xp = __dsub_rn(__dmul_rn(JZ_B, Xd), __dmul_rn(JZ_B - 1.0, (double)Z));
    const double Yp = __dsub_rn(__dmul_rn(JZ_G, Yd), __dmul_rn(JZ_G - 1.0, Xd));
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float* m = C.fwd1 + 3 * k;
        const float mz = __fmul_rn(m[2], Z);
        const double L = __dadd_rn(__dadd_rn(__dmul_rn((double)m[0], Xp), __dmul_rn((double)m[1], Yp)), (double)mz);
        p[k] = pqf_inv_eotf(Q, 1, JZ_P, L, r[k], ok);
    }
    if (!ok) return false;
    double Iz, Ei, o, E;
    row3_fast(C.fwd2, p[0], p[1], p[2], r[0], r[1], r[2], Iz, Ei);          // Iz stays a double: judged after the Jz formula
    if (!row3_fast(C.fwd2 + 3, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
    o1 = (float)o;
    if (!row3_fast(C.fwd2 + 6, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
    o2 = (float)o;
    const double den = __dadd_rn(1.0, __dmul_rn(JZ_D, Iz));
    const double Jz = __dsub_rn(__ddiv_rn(__dmul_rn(1.0 + JZ_D, Iz), den), JZ_D0);
    const double rd = pqf_rcp(den);
    const double Ej = Ei * ((1.0 + JZ_D) * rd * rd) * 1.001 + 8.0 * PQF_U * (fabs(Jz) + JZ_D0);      // d/dIz of (1 + d) Iz / (1 + d Iz)
    if (!(den > 0.25) || !round_is_safe(Jz, Ej)) return false;
    o0 = (float)Jz;
    return true;
}

// space -> XYZ (the part of color_inv before the sRGB encode)
template <int SPACE>
__device__ __forceinline__ bool color_inv_xyz_fast(const ColorConsts& C, float a, float b, float c, float& X, float& Y, float& Z) {
    const PqTabs& Q = C.pq;
    bool ok = true;
    if (SPACE == AEAJ_OKLAB) {                                   // (float) l'^3, oklab.py:93-96
        float v[3], w[3];
        dot3(C.inv1, a, b, c, v[0], v[1], v[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (v[k] == 0.0f) { w[k] = v[k]; continue; }
            const double p = powtab_eval<8>(Q.cube, fabs((double)v[k]), ok);
            if (!ok || !round_is_safe(p, p * (Q.cube.eps + 2.0 * PQF_U))) return false;
            w[k] = v[k] < 0.0f ? -(float)p : (float)p;
        }
        dot3(C.inv2, w[0], w[1], w[2], X, Y, Z);
        return true;
    }
    double p[3], r[3], o, E;
    if (SPACE == AEAJ_ICACB || SPACE == AEAJ_ICTCP) {
        float lp, mp, sp;
        dot3(C.inv1, a, b, c, lp, mp, sp);
        p[0] = pqf_eotf(Q, 0, (double)lp, r[0], ok);
        p[1] = pqf_eotf(Q, 0, (double)mp, r[1], ok);
        p[2] = pqf_eotf(Q, 0, (double)sp, r[2], ok);
        if (!ok) return false;
        if (!row3_fast(C.inv2, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
        X = (float)o;
        if (!row3_fast(C.inv2 + 3, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
        Y = (float)o;
        if (!row3_fast(C.inv2 + 6, p[0], p[1], p[2], r[0], r[1], r[2], o, E)) return false;
        Z = (float)o;
        return true;
    }
    // JzAzBz (jzazbz.py:131-171)
    const double jd = __dadd_rn((double)a, JZ_D0);
    const double Iz = __ddiv_rn(jd, __dsub_rn(1.0 + JZ_D, __dmul_rn(JZ_D, jd)));
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float* m = C.inv1 + 3 * k;
        const float ta = __fmul_rn(m[1], b), tb = __fmul_rn(m[2], c);
        const double lp = __dadd_rn(__dadd_rn(__dmul_rn((double)m[0], Iz), (double)ta), (double)tb);
        p[k] = pqf_eotf(Q, 1, lp, r[k], ok);
    }
    if (!ok) return false;
    double Xp, Yp, Zp, Ex, Ey, Ez;
    row3_fast(C.inv2, p[0], p[1], p[2], r[0], r[1], r[2], Xp, Ex);
    row3_fast(C.inv2 + 3, p[0], p[1], p[2], r[0], r[1], r[2], Yp, Ey);
    if (!row3_fast(C.inv2 + 6, p[0], p[1], p[2], r[0], r[1], r[2], Zp, Ez)) return false;
    const double Xd = __ddiv_rn(__dadd_rn(Xp, __dmul_rn(JZ_B - 1.0, Zp)), JZ_B);
    const double Exd = (Ex + (JZ_B - 1.0) * Ez) / JZ_B + 8.0 * PQF_U * (fabs(Xp) + fabs(Zp));
    const double Yd = __ddiv_rn(__dadd_rn(Yp, __dmul_rn(JZ_G - 1.0, Xd)), JZ_G);
    const double Eyd = (Ey + (1.0 - JZ_G) * Exd) / JZ_G + 8.0 * PQF_U * (fabs(Yp) + fabs(Xd)) / JZ_G;
    if (!round_is_safe(Xd, Exd) || !round_is_safe(Yd, Eyd)) return false;
    X = (float)Xd; Y = (float)Yd; Z = (float)Zp;
    return true;
}

// common.py:62-92 on one channel, clamp included
__device__ __forceinline__ bool linear_to_srgb_fast(const ColorConsts& C, float v, float& out) {
    bool ok = true;
    double E;
    const double r = pqf_linear_to_srgb(C.pq, (double)v, E, ok);
    if (!ok || !round_is_safe(r, E)) return false;
    float f = (float)r;
    f = (f < 1.0f) ? f : 1.0f;
    f = (f > 0.0f) ? f : 0.0f;
    out = f;
    return true;
}

template <int SPACE>
__device__ __forceinline__ void color_fwd(const ColorConsts& C, const PowTabs& T, const float* lut, float r, float g, float b,
                                          float& o0, float& o1, float& o2) {
    if (SPACE <= AEAJ_YCOCG_R) { dot3(C.fwd1, r, g, b, o0, o1, o2); return; }
    float lr = srgb_to_linear(T, r, lut), lg = srgb_to_linear(T, g, lut), lb = srgb_to_linear(T, b, lut);
    float X, Y, Z;
    dot3(c_rgb2xyz, lr, lg, lb, X, Y, Z);
    if (SPACE == AEAJ_XYZ) { o0 = X; o1 = Y; o2 = Z; return; }
    if (C.fast) { FAST_STAT(0); if (color_fwd_fast<SPACE>(C, X, Y, Z, o0, o1, o2)) return; FAST_STAT(1); }
#ifdef AEAJ_FAST_ONLY                                   /* timing experiment only: wrong results for the pixels that need the exact path */
    if (C.fast) { o0 = o1 = o2 = 0.0f; return; }
#endif
    if (SPACE == AEAJ_OKLAB) {                                   // oklab.py:71-75
        float l, m, s;
        dot3(C.fwd1, X, Y, Z, l, m, s);
        const double e = (double)(float)(1.0 / 3.0);             // numpy casts the exponent to float32
        float lp = (float)fpow(T, (double)l, e), mp = (float)fpow(T, (double)m, e), sp = (float)fpow(T, (double)s, e);
        dot3(C.fwd2, lp, mp, sp, o0, o1, o2);
    } else if (SPACE == AEAJ_ICACB || SPACE == AEAJ_ICTCP) {     // ictcp.py:47-79
        float l, m, s;
        dot3(C.fwd1, X, Y, Z, l, m, s);
        double lp = pq_inv_eotf(T, (double)l, PQ_M2), mp = pq_inv_eotf(T, (double)m, PQ_M2), sp = pq_inv_eotf(T, (double)s, PQ_M2);
        const float* q = C.fwd2;
        o0 = (float)mul3add((double)q[0], lp, (double)q[1], mp, (double)q[2], sp);
        o1 = (float)mul3add((double)q[3], lp, (double)q[4], mp, (double)q[5], sp);
        o2 = (float)mul3add((double)q[6], lp, (double)q[7], mp, (double)q[8], sp);
    } else {                                                     // JzAzBz: jzazbz.py:57-97
        double Xd = (double)X, Yd = (double)Y;
        double Xp = __dsub_rn(__dmul_rn(JZ_B, Xd), __dmul_rn(JZ_B - 1.0, (double)Z));
        double Yp = __dsub_rn(__dmul_rn(JZ_G, Yd), __dmul_rn(JZ_G - 1.0, Xd));
        double lp[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float* m = C.fwd1 + 3 * k;
            float mz = __fmul_rn(m[2], Z);                       // Z' stays f32 (jzazbz.py:63)
            double L = __dadd_rn(__dadd_rn(__dmul_rn((double)m[0], Xp), __dmul_rn((double)m[1], Yp)), (double)mz);
            lp[k] = pq_inv_eotf(T, L, JZ_P);
        }
        const float* q = C.fwd2;
        double Iz = mul3add((double)q[0], lp[0], (double)q[1], lp[1], (double)q[2], lp[2]);
        double Az = mul3add((double)q[3], lp[0], (double)q[4], lp[1], (double)q[5], lp[2]);
        double Bz = mul3add((double)q[6], lp[0], (double)q[7], lp[1], (double)q[8], lp[2]);
        double Jz = __dsub_rn(__ddiv_rn(__dmul_rn(1.0 + JZ_D, Iz), __dadd_rn(1.0, __dmul_rn(JZ_D, Iz))), JZ_D0);
        o0 = (float)Jz; o1 = (float)Az; o2 = (float)Bz;
    }
}

template <int SPACE>
__device__ __forceinline__ void color_inv(const ColorConsts& C, const PowTabs& T, float a, float b, float c, float& r, float& g, float& bl) {
    if (SPACE <= AEAJ_YCOCG_R) {                                 // ycbcr.py:79-82: dot then np.clip
        float t0, t1, t2;
        dot3(C.inv1, a, b, c, t0, t1, t2);
        r = fminf(fmaxf(t0, 0.0f), 1.0f); g = fminf(fmaxf(t1, 0.0f), 1.0f); bl = fminf(fmaxf(t2, 0.0f), 1.0f);
        return;
    }
    float X, Y, Z;
    if (SPACE == AEAJ_XYZ) { X = a; Y = b; Z = c; }
    else if (C.fast && (FAST_STAT_INV(2), color_inv_xyz_fast<SPACE>(C, a, b, c, X, Y, Z))) { }
    else if ((FAST_STAT_INV(3), SPACE == AEAJ_OKLAB)) {          // oklab.py:93-96
        float lp, mp, sp;
        dot3(C.inv1, a, b, c, lp, mp, sp);
        float l = (float)fpow(T, (double)lp, 3.0), m = (float)fpow(T, (double)mp, 3.0), s = (float)fpow(T, (double)sp, 3.0);
        dot3(C.inv2, l, m, s, X, Y, Z);
    } else if (SPACE == AEAJ_ICACB || SPACE == AEAJ_ICTCP) {     // ictcp.py:103-137
        float lp, mp, sp;
        dot3(C.inv1, a, b, c, lp, mp, sp);
        double l = pq_eotf(T, (double)lp, PQ_M2), m = pq_eotf(T, (double)mp, PQ_M2), s = pq_eotf(T, (double)sp, PQ_M2);
        const float* q = C.inv2;
        X = (float)mul3add((double)q[0], l, (double)q[1], m, (double)q[2], s);
        Y = (float)mul3add((double)q[3], l, (double)q[4], m, (double)q[5], s);
        Z = (float)mul3add((double)q[6], l, (double)q[7], m, (double)q[8], s);
    } else {                                                     // jzazbz.py:131-171
        double jd = __dadd_rn((double)a, JZ_D0);
        double Iz = __ddiv_rn(jd, __dsub_rn(1.0 + JZ_D, __dmul_rn(JZ_D, jd)));
        double l[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float* m = C.inv1 + 3 * k;
            float ta = __fmul_rn(m[1], b), tb = __fmul_rn(m[2], c);
            double lp = __dadd_rn(__dadd_rn(__dmul_rn((double)m[0], Iz), (double)ta), (double)tb);
            l[k] = pq_eotf(T, lp, JZ_P);
        }
        const float* q = C.inv2;
        double Xp = mul3add((double)q[0], l[0], (double)q[1], l[1], (double)q[2], l[2]);
        double Yp = mul3add((double)q[3], l[0], (double)q[4], l[1], (double)q[5], l[2]);
        double Zp = mul3add((double)q[6], l[0], (double)q[7], l[1], (double)q[8], l[2]);
        double Xd = __ddiv_rn(__dadd_rn(Xp, __dmul_rn(JZ_B - 1.0, Zp)), JZ_B);
        double Yd = __ddiv_rn(__dadd_rn(Yp, __dmul_rn(JZ_G - 1.0, Xd)), JZ_G);
        X = (float)Xd; Y = (float)Yd; Z = (float)Zp;
    }
    float lr, lg, lb;
    dot3(c_xyz2rgb, X, Y, Z, lr, lg, lb);
    if (!(C.fast && linear_to_srgb_fast(C, lr, r))) { if (C.fast) FAST_STAT(4); r = linear_to_srgb(T, lr); }
    if (!(C.fast && linear_to_srgb_fast(C, lg, g))) g = linear_to_srgb(T, lg);
    if (!(C.fast && linear_to_srgb_fast(C, lb, bl))) bl = linear_to_srgb(T, lb);
}

__device__ __forceinline__ uint8_t cast_u8(float v) {            // (img*255).astype(np.uint8)
    return (uint8_t)(__float2int_rz(__fmul_rn(v, 255.0f)) & 0xff);
}

__device__ __forceinline__ const float* load_lut(const float* lut_g, float* lut_s) {
    if (!lut_g) return nullptr;
    for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 256; i += blockDim.x * blockDim.y) lut_s[i] = lut_g[i];
    __syncthreads();
    return lut_s;
}

// ---------------------------------------------------------------------------------------------
// (N,3) -> (N,3) pixel kernels: the `convert()` drop-in (conversion.py:95-124)
// ---------------------------------------------------------------------------------------------
template <int SPACE, bool INVERSE>
__global__ void __launch_bounds__(256) k_color_pixels(const __grid_constant__ ColorConsts C, const float* __restrict__ lut_g,
                                                      const float* __restrict__ in, float* __restrict__ out, size_t n) {
    __shared__ float lut_s[256];
    __shared__ double pow_s[SPACE > AEAJ_YCOCG_R ? POW_TAB_DOUBLES : 1];
    const float* lut = (SPACE > AEAJ_YCOCG_R && !INVERSE) ? load_lut(lut_g, lut_s) : nullptr;
    PowTabs T = {nullptr, nullptr};
    if (SPACE > AEAJ_YCOCG_R) T = load_pow_tabs(pow_s);
    // 4 pixels (12 floats = 3 x float4) per thread when possible
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t nq = n / 4;
    for (; q < nq; q += (size_t)gridDim.x * blockDim.x) {
        const float4* p = reinterpret_cast<const float4*>(in) + q * 3;
        float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
        float x[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w}, y[12];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (INVERSE) color_inv<SPACE>(C, T, x[3 * k], x[3 * k + 1], x[3 * k + 2], y[3 * k], y[3 * k + 1], y[3 * k + 2]);
            else color_fwd<SPACE>(C, T, lut, x[3 * k], x[3 * k + 1], x[3 * k + 2], y[3 * k], y[3 * k + 1], y[3 * k + 2]);
        }
        float4* o = reinterpret_cast<float4*>(out) + q * 3;
        o[0] = make_float4(y[0], y[1], y[2], y[3]); o[1] = make_float4(y[4], y[5], y[6], y[7]); o[2] = make_float4(y[8], y[9], y[10], y[11]);
    }
    // tail
    size_t t = nq * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        float y0, y1, y2;
        if (INVERSE) color_inv<SPACE>(C, T, in[3 * t], in[3 * t + 1], in[3 * t + 2], y0, y1, y2);
        else color_fwd<SPACE>(C, T, lut, in[3 * t], in[3 * t + 1], in[3 * t + 2], y0, y1, y2);
        out[3 * t] = y0; out[3 * t + 1] = y1; out[3 * t + 2] = y2;
    }
}

// ---------------------------------------------------------------------------------------------
// fused forward: RGB (HWC) -> planar layers (+ chroma INTER_AREA) + u8 cast planes
// MODE 0: chroma 2x2, needs H%2==0, W%4==0; thread = 4 px x 2 rows
// MODE 1: chroma 1x4, needs W%4==0;          thread = 4 px x 1 row
// MODE 2: anything: thread = 1 px, chroma written full-res to scratch (area kernel follows)
// ---------------------------------------------------------------------------------------------
struct FwdOut {
    float* y; float* c1; float* c2;          // per-image strides below
    uint8_t* y8; uint8_t* c18; uint8_t* c28;
    size_t sy, sc;                           // elements per image for luma / chroma outputs
};

// 8-bit source pixels: Image.load's imread(path).astype(np.float32) / 255.0 (image.py:84), one correctly rounded division
__device__ __forceinline__ float u8_to_unit(unsigned v) { return __fdiv_rn((float)v, 255.0f); }
// Image.get_uint8 / Image.save: (data * 255).astype(np.uint8) (image.py:112,127); data is clipped to [0,1] by the inverse colour
__device__ __forceinline__ unsigned unit_to_u8(float v) { return (unsigned)(int)__fmul_rn(v, 255.0f) & 0xffu; }

template <int SPACE, int MODE, bool U8>
__global__ void __launch_bounds__(256) k_color_forward_planar(const __grid_constant__ ColorConsts C, const float* __restrict__ lut_g,
                                                              const void* __restrict__ rgb_any, int H, int W, FwdOut o) {
    __shared__ float lut_s[256];
    __shared__ double pow_s[SPACE > AEAJ_YCOCG_R ? POW_TAB_DOUBLES : 1];
    const float* lut = (SPACE > AEAJ_YCOCG_R) ? load_lut(lut_g, lut_s) : nullptr;
    PowTabs T = {nullptr, nullptr};
    if (SPACE > AEAJ_YCOCG_R) T = load_pow_tabs(pow_s);
    const int b = blockIdx.z;
    const float* img = reinterpret_cast<const float*>(rgb_any) + (U8 ? 0 : (size_t)b * H * W * 3);
    const uint8_t* img8 = reinterpret_cast<const uint8_t*>(rgb_any) + (U8 ? (size_t)b * H * W * 3 : 0);
    if (MODE == 2) {
        int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
        if (x >= W || y >= H) return;
        size_t i = (size_t)y * W + x;
        float v0, v1, v2;
        if (U8) color_fwd<SPACE>(C, T, lut, u8_to_unit(img8[3 * i]), u8_to_unit(img8[3 * i + 1]), u8_to_unit(img8[3 * i + 2]), v0, v1, v2);
        else color_fwd<SPACE>(C, T, lut, img[3 * i], img[3 * i + 1], img[3 * i + 2], v0, v1, v2);
        o.y[b * o.sy + i] = v0; o.y8[b * o.sy + i] = cast_u8(v0);
        o.c1[b * o.sc + i] = v1; o.c2[b * o.sc + i] = v2;     // full-res scratch (sc == H*W here)
        return;
    }
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int ROWS = (MODE == 0) ? 2 : 1;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * ROWS;
    if (x0 >= W || y0 >= H) return;
    float c1v[2][4], c2v[2][4];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        float x[12];
        if (U8) {
            // 4 px = 12 bytes = three aligned words (W % 4 == 0 in the fused modes)
            const uint32_t* p8 = reinterpret_cast<const uint32_t*>(img8 + ((size_t)(y0 + r) * W + x0) * 3);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const uint32_t wv = __ldg(p8 + k);
#pragma unroll
                for (int j = 0; j < 4; j++) x[4 * k + j] = u8_to_unit((wv >> (8 * j)) & 0xffu);
            }
        } else {
            const float4* p = reinterpret_cast<const float4*>(img + ((size_t)(y0 + r) * W + x0) * 3);
            float4 a = __ldg(p), bq = __ldg(p + 1), c = __ldg(p + 2);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = bq.x; x[5] = bq.y; x[6] = bq.z; x[7] = bq.w; x[8] = c.x; x[9] = c.y; x[10] = c.z; x[11] = c.w;
        }
        float yv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) color_fwd<SPACE>(C, T, lut, x[3 * k], x[3 * k + 1], x[3 * k + 2], yv[k], c1v[r][k], c2v[r][k]);
        size_t i = (size_t)b * o.sy + (size_t)(y0 + r) * W + x0;
        *reinterpret_cast<float4*>(o.y + i) = make_float4(yv[0], yv[1], yv[2], yv[3]);
        *reinterpret_cast<uchar4*>(o.y8 + i) = make_uchar4(cast_u8(yv[0]), cast_u8(yv[1]), cast_u8(yv[2]), cast_u8(yv[3]));
    }
    if (MODE == 0) {
        // cv.resize INTER_AREA 2x2: ((a+b)+(c+d))*0.25f
        const int cw = W >> 1;
        size_t i = (size_t)b * o.sc + (size_t)(y0 >> 1) * cw + (x0 >> 1);
        float u0 = __fmul_rn(__fadd_rn(__fadd_rn(c1v[0][0], c1v[0][1]), __fadd_rn(c1v[1][0], c1v[1][1])), 0.25f);
        float u1 = __fmul_rn(__fadd_rn(__fadd_rn(c1v[0][2], c1v[0][3]), __fadd_rn(c1v[1][2], c1v[1][3])), 0.25f);
        float w0 = __fmul_rn(__fadd_rn(__fadd_rn(c2v[0][0], c2v[0][1]), __fadd_rn(c2v[1][0], c2v[1][1])), 0.25f);
        float w1 = __fmul_rn(__fadd_rn(__fadd_rn(c2v[0][2], c2v[0][3]), __fadd_rn(c2v[1][2], c2v[1][3])), 0.25f);
        *reinterpret_cast<float2*>(o.c1 + i) = make_float2(u0, u1);
        *reinterpret_cast<float2*>(o.c2 + i) = make_float2(w0, w1);
        *reinterpret_cast<uchar2*>(o.c18 + i) = make_uchar2(cast_u8(u0), cast_u8(u1));
        *reinterpret_cast<uchar2*>(o.c28 + i) = make_uchar2(cast_u8(w0), cast_u8(w1));
    } else {
        // 1x4: (((a+b)+c)+d)*0.25f
        const int cw = W >> 2;
        size_t i = (size_t)b * o.sc + (size_t)y0 * cw + (x0 >> 2);
        float u = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(c1v[0][0], c1v[0][1]), c1v[0][2]), c1v[0][3]), 0.25f);
        float w = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(c2v[0][0], c2v[0][1]), c2v[0][2]), c2v[0][3]), 0.25f);
        o.c1[i] = u; o.c2[i] = w; o.c18[i] = cast_u8(u); o.c28[i] = cast_u8(w);
    }
}

// ---------------------------------------------------------------------------------------------
// general INTER_AREA (OpenCV ResizeArea_Invoker with computeResizeAreaTab), one thread per output
// sample; used when a dimension is not divisible by the subsampling ratio (jpeg.py:686 floors).
// ---------------------------------------------------------------------------------------------
struct AreaAxis { int s1, s2; float aL, aM, aR; int hasL, hasR; };
__device__ __forceinline__ AreaAxis area_axis(int d, int ssize, double scale) {
    AreaAxis a;
    double f1 = d * scale, f2 = f1 + scale;
    double cell = fmin(scale, ssize - f1);
    int s1 = (int)ceil(f1), s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    a.s1 = s1; a.s2 = s2;
    a.hasL = (s1 - f1 > 1e-3); a.aL = (float)((s1 - f1) / cell);
    a.aM = (float)(1.0 / cell);
    a.hasR = (f2 - s2 > 1e-3); a.aR = (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell);
    return a;
}
__global__ void __launch_bounds__(256) k_area(const float* __restrict__ src, int H, int W, float* __restrict__ dst, int dh, int dw,
                                              uint8_t* __restrict__ u8, size_t src_stride, size_t dst_stride) {
    int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const float* S = src + blockIdx.z * src_stride;
    AreaAxis ax = area_axis(dx, W, (double)W / dw), ay = area_axis(dy, H, (double)H / dh);
    float sum = 0.0f;
    bool first = true;
    int j0 = ay.hasL ? ay.s1 - 1 : ay.s1, j1 = ay.hasR ? ay.s2 : ay.s2 - 1;
    for (int j = j0; j <= j1; j++) {
        float beta = (j < ay.s1) ? ay.aL : ((j >= ay.s2) ? ay.aR : ay.aM);
        const float* row = S + (size_t)j * W;
        float buf = 0.0f;
        if (ax.hasL) buf = __fadd_rn(buf, __fmul_rn(row[ax.s1 - 1], ax.aL));
        for (int k = ax.s1; k < ax.s2; k++) buf = __fadd_rn(buf, __fmul_rn(row[k], ax.aM));
        if (ax.hasR) buf = __fadd_rn(buf, __fmul_rn(row[ax.s2], ax.aR));
        float t = __fmul_rn(beta, buf);
        sum = first ? t : __fadd_rn(sum, t);
        first = false;
    }
    size_t i = blockIdx.z * dst_stride + (size_t)dy * dw + dx;
    dst[i] = sum;
    if (u8) u8[i] = cast_u8(sum);
}

// ---------------------------------------------------------------------------------------------
// INTER_LINEAR sample (half-pixel centres, edge clamp; horizontal then vertical; no fma)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_coord(int d, double scale, int ssize, int& s0, int& s1, float& f) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (s < 0) { fx = 0.0f; s = 0; }
    if (s >= ssize - 1) { fx = 0.0f; s = ssize - 1; }
    s0 = s; s1 = min(s + 1, ssize - 1); f = fx;
}
__device__ __forceinline__ float ldf(const float* p, bool remote) { return remote ? __ldcv(p) : __ldg(p); }   // remote: a peer GPU's row
__device__ __forceinline__ float lin_sample_rows(const float* r0, const float* r1, bool rem0, bool rem1, int x0, int x1, float fx, float fy) {
    float a0 = __fsub_rn(1.0f, fx), b0 = __fsub_rn(1.0f, fy);
    float t0 = __fadd_rn(__fmul_rn(ldf(r0 + x0, rem0), a0), __fmul_rn(ldf(r0 + x1, rem0), fx));
    float t1 = __fadd_rn(__fmul_rn(ldf(r1 + x0, rem1), a0), __fmul_rn(ldf(r1 + x1, rem1), fx));
    return __fadd_rn(__fmul_rn(t0, b0), __fmul_rn(t1, fy));
}
__device__ __forceinline__ float lin_sample(const float* __restrict__ p, int sw, int x0, int x1, float fx, int y0, int y1, float fy) {
    return lin_sample_rows(p + (size_t)y0 * sw, p + (size_t)y1 * sw, false, false, x0, x1, fx, fy);
}
__global__ void __launch_bounds__(256) k_resize_linear(const float* __restrict__ src, int sh, int sw, float* __restrict__ dst, int H, int W) {
    int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= W || dy >= H) return;
    if (sh == H && sw == W) { dst[(size_t)dy * W + dx] = src[(size_t)dy * W + dx]; return; }
    int x0, x1, y0, y1; float fx, fy;
    lin_coord(dx, (double)sw / W, sw, x0, x1, fx);
    lin_coord(dy, (double)sh / H, sh, y0, y1, fy);
    dst[(size_t)dy * W + dx] = lin_sample(src, sw, x0, x1, fx, y0, y1, fy);
}

// fused decode tail: per-layer INTER_LINEAR upsample (jpeg.py:340-354) + stack + inverse colour
// (jpeg.py:290-297) -> RGB HWC
struct UpIn {
    const float* p[3]; int h[3], w[3]; size_t stride[3];
    int lo[3], hi[3];                  // halo-split: rows of every layer held by this GPU ...
    long long peer_up, peer_dn;        // ... the row above / below comes from the neighbour rank's copy (byte deltas; 0: local)
};
// row `y` of layer l (batch image b): this GPU's copy or the neighbour's.  PEER = false (every launch but those of a
// multi-GPU halo-split rank) compiles to the plain address: no band tests, no second load flavour.
template <bool PEER>
__device__ __forceinline__ const float* up_row(const UpIn& in, int l, int b, int y, bool& remote) {
    remote = false;
    if (!PEER) return in.p[l] + (size_t)b * in.stride[l] + (size_t)y * in.w[l];
    const long long d = (y < in.lo[l]) ? in.peer_up : ((y >= in.hi[l]) ? in.peer_dn : 0ll);
    remote = d != 0;
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(in.p[l] + (size_t)b * in.stride[l]) + d) + (size_t)y * in.w[l];
}

template <int SPACE>
__global__ void __launch_bounds__(256) k_upsample_color_inverse(const __grid_constant__ ColorConsts C, UpIn in, int H, int W, float* __restrict__ rgb,
                                                                uint8_t* __restrict__ rgb8, int y_lo, int y_hi) {
    __shared__ double pow_s[SPACE > AEAJ_YCOCG_R ? POW_TAB_DOUBLES : 1];
    PowTabs T = {nullptr, nullptr};
    if (SPACE > AEAJ_YCOCG_R) T = load_pow_tabs(pow_s);
    int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = y_lo + blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= W || dy >= y_hi) return;
    int b = blockIdx.z;
    float v[3];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        const float* p = in.p[l] + (size_t)b * in.stride[l];
        if (in.h[l] == H && in.w[l] == W) v[l] = __ldg(p + (size_t)dy * W + dx);
        else {
            int x0, x1, y0, y1; float fx, fy;
            lin_coord(dx, (double)in.w[l] / W, in.w[l], x0, x1, fx);
            lin_coord(dy, (double)in.h[l] / H, in.h[l], y0, y1, fy);
            bool rem0, rem1;
            const float* r0 = up_row<true>(in, l, b, y0, rem0);
            const float* r1 = up_row<true>(in, l, b, y1, rem1);
            v[l] = lin_sample_rows(r0, r1, rem0, rem1, x0, x1, fx, fy);
        }
    }
    float r, g, bl;
    color_inv<SPACE>(C, T, v[0], v[1], v[2], r, g, bl);
    const size_t oi = ((size_t)b * H * W + (size_t)dy * W + dx) * 3;
    if (rgb) { rgb[oi] = r; rgb[oi + 1] = g; rgb[oi + 2] = bl; }
    if (rgb8) { rgb8[oi] = (uint8_t)unit_to_u8(r); rgb8[oi + 1] = (uint8_t)unit_to_u8(g); rgb8[oi + 2] = (uint8_t)unit_to_u8(bl); }
}

// 2x2 chroma specialisation (H == 2h, W == 2w, W % 4 == 0): 4 output px per thread, float4 stores.
// The half-pixel coordinates are exact dyadic numbers here -- even dx: source k-1 with t = .75, odd dx:
// source k with t = .25, clamped with t = 0 at the borders -- so this reproduces lin_coord() bit for bit.
__device__ __forceinline__ void up2_coord(int d, int ssize, int& s0, int& s1, float& f) {
    int k = d >> 1;
    int s = (d & 1) ? k : k - 1;
    float fx = (d & 1) ? 0.25f : 0.75f;
    if (s < 0) { fx = 0.0f; s = 0; }
    if (s >= ssize - 1) { fx = 0.0f; s = ssize - 1; }
    s0 = s; s1 = min(s + 1, ssize - 1); f = fx;
}
template <int SPACE, bool PEER>
__global__ void __launch_bounds__(256) k_upsample2x_color_inverse(const __grid_constant__ ColorConsts C, UpIn in, int H, int W, float* __restrict__ rgb,
                                                                  uint8_t* __restrict__ rgb8, int y_lo, int y_hi) {
    __shared__ double pow_s[SPACE > AEAJ_YCOCG_R ? POW_TAB_DOUBLES : 1];
    PowTabs T = {nullptr, nullptr};
    if (SPACE > AEAJ_YCOCG_R) T = load_pow_tabs(pow_s);
    const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, dy = y_lo + blockIdx.y * blockDim.y + threadIdx.y;
    if (dx0 >= W || dy >= y_hi) return;
    const int b = blockIdx.z;
    const int cw = in.w[1], chh = in.h[1];
    const float4 yv = __ldg(reinterpret_cast<const float4*>(in.p[0] + (size_t)b * in.stride[0] + (size_t)dy * W + dx0));
    const float lum[4] = {yv.x, yv.y, yv.z, yv.w};
    int y0, y1; float fy;
    up2_coord(dy, chh, y0, y1, fy);
    const float b0 = __fsub_rn(1.0f, fy);
    // The 4 px of a thread (dx0 = 4m) read chroma columns 2m-1 .. 2m+2: px0 = (2m-1, 2m; .75), px1 = (2m, 2m+1; .25),
    // px2 = (2m, 2m+1; .75), px3 = (2m+1, 2m+2; .25), with up2_coord's clamps at the two borders (t = 0).  Same operands
    // and the same a*(1-t) + b*t expressions as the per-pixel form, with 3 loads per row instead of 8.
    const int m2 = dx0 >> 1;
    const bool first = (m2 == 0), last = (m2 + 1 >= cw - 1);
    const float f0 = first ? 0.0f : 0.75f, f3 = last ? 0.0f : 0.25f;
    const float a0 = __fsub_rn(1.0f, f0), a3 = __fsub_rn(1.0f, f3);
    float cv[2][4];
#pragma unroll
    for (int l = 1; l <= 2; l++) {
        float t[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            bool rem;
            const float* row = up_row<PEER>(in, l, b, r ? y1 : y0, rem);
            const float2 mid = (PEER && rem) ? __ldcv(reinterpret_cast<const float2*>(row + m2)) : __ldg(reinterpret_cast<const float2*>(row + m2));
            const float vm = ldf(row + max(m2 - 1, 0), PEER && rem), vp = ldf(row + min(m2 + 2, cw - 1), PEER && rem);
            const float p0a = first ? mid.x : vm, p0b = first ? mid.y : mid.x;        // (x0, x1) of px0
            const float p3b = last ? mid.y : vp;                                       // x1 of px3 (x0 = 2m+1)
            t[r][0] = __fadd_rn(__fmul_rn(p0a, a0), __fmul_rn(p0b, f0));
            t[r][1] = __fadd_rn(__fmul_rn(mid.x, 0.75f), __fmul_rn(mid.y, 0.25f));
            t[r][2] = __fadd_rn(__fmul_rn(mid.x, 0.25f), __fmul_rn(mid.y, 0.75f));
            t[r][3] = __fadd_rn(__fmul_rn(mid.y, a3), __fmul_rn(p3b, f3));
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cv[l - 1][k] = __fadd_rn(__fmul_rn(t[0][k], b0), __fmul_rn(t[1][k], fy));
    }
    float o[12];
#pragma unroll
    for (int k = 0; k < 4; k++) color_inv<SPACE>(C, T, lum[k], cv[0][k], cv[1][k], o[3 * k], o[3 * k + 1], o[3 * k + 2]);
    const size_t oi = ((size_t)b * H * W + (size_t)dy * W + dx0) * 3;
    if (rgb) {
        float4* out = reinterpret_cast<float4*>(rgb + oi);
        out[0] = make_float4(o[0], o[1], o[2], o[3]); out[1] = make_float4(o[4], o[5], o[6], o[7]); out[2] = make_float4(o[8], o[9], o[10], o[11]);
    }
    if (rgb8) {
        uint32_t* out8 = reinterpret_cast<uint32_t*>(rgb8 + oi);          // 12 bytes, word aligned (W % 4 == 0)
#pragma unroll
        for (int k = 0; k < 3; k++)
            out8[k] = unit_to_u8(o[4 * k]) | (unit_to_u8(o[4 * k + 1]) << 8) | (unit_to_u8(o[4 * k + 2]) << 16) | (unit_to_u8(o[4 * k + 3]) << 24);
    }
}

__global__ void __launch_bounds__(256) k_normalize(const float* __restrict__ in, float* __restrict__ out, size_t n, float mid, float scale, int inverse) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = inverse ? __fadd_rn(__fdiv_rn(in[i], scale), mid) : __fmul_rn(__fsub_rn(in[i], mid), scale);
}
__global__ void __launch_bounds__(256) k_cast_u8(const float* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = cast_u8(in[i]);
}

template <typename F>
int dispatch_space(int space, F&& f) {
    switch (space) {
        case AEAJ_YCBCR: return f(std::integral_constant<int, AEAJ_YCBCR>());
        case AEAJ_YCOCG: return f(std::integral_constant<int, AEAJ_YCOCG>());
        case AEAJ_YCOCG_R: return f(std::integral_constant<int, AEAJ_YCOCG_R>());
        case AEAJ_OKLAB: return f(std::integral_constant<int, AEAJ_OKLAB>());
        case AEAJ_ICACB: return f(std::integral_constant<int, AEAJ_ICACB>());
        case AEAJ_ICTCP: return f(std::integral_constant<int, AEAJ_ICTCP>());
        case AEAJ_JZAZBZ: return f(std::integral_constant<int, AEAJ_JZAZBZ>());
        case AEAJ_XYZ: return f(std::integral_constant<int, AEAJ_XYZ>());
    }
    aeaj_set_error("unknown colour space id %d", space);
    return AEAJ_EINVAL;
}

}  // namespace

// builds the power tables on the host (long double), uploads them once and hands every colour space a view (pqtabs_build.h)
int aeaj_color_init(aeaj_handle* h) {
    PqTabsHost H;
    pqtabs_build(H);
    const PowTabHost* tabs[9] = {&H.m1, &H.m2[0], &H.m2[1], &H.im2[0], &H.im2[1], &H.im1, &H.isrgb, &H.cbrt32, &H.cube};
    PowTabView* views[9] = {&H.view.m1, &H.view.m2[0], &H.view.m2[1], &H.view.im2[0], &H.view.im2[1], &H.view.im1, &H.view.isrgb, &H.view.cbrt32, &H.view.cube};
    size_t total = 0;
    double worst = 0.0;
    for (int i = 0; i < 9; i++) { total += tabs[i]->te.size() + tabs[i]->coef.size(); worst = std::max(worst, tabs[i]->v.eps); }
    AEAJ_REQUIRE(worst < 4e-15, "power tables: the fitted polynomials miss their accuracy target (host long double too short?)");
    std::vector<double> flat;
    flat.reserve(total + 18);
    size_t off_te[9], off_coef[9];
    for (int i = 0; i < 9; i++) {
        off_te[i] = flat.size();
        flat.insert(flat.end(), tabs[i]->te.begin(), tabs[i]->te.end());
        if (flat.size() & 1) flat.push_back(0.0);                  // coefficient pairs are fetched with 16-byte loads
        off_coef[i] = flat.size();
        flat.insert(flat.end(), tabs[i]->coef.begin(), tabs[i]->coef.end());
    }
    FnTabHost* curves[4] = {&H.enc[0], &H.enc[1], &H.dec[0], &H.dec[1]};
    FnTabView* cviews[4] = {&H.view.enc[0], &H.view.enc[1], &H.view.dec[0], &H.view.dec[1]};
    size_t off_curve[4];
    for (int i = 0; i < 4; i++) {
        if (flat.size() & 1) flat.push_back(0.0);
        off_curve[i] = flat.size();
        flat.insert(flat.end(), curves[i]->coef.begin(), curves[i]->coef.end());
        worst = std::max(worst, curves[i]->v.eps);
    }
    AEAJ_REQUIRE(worst < 4e-15, "transfer-curve tables: the fitted polynomials miss their accuracy target");
    AEAJ_CUDA(cudaMalloc(&h->pq_tabs_dev, flat.size() * sizeof(double)));
    for (int i = 0; i < 9; i++) { views[i]->te = h->pq_tabs_dev + off_te[i]; views[i]->coef = h->pq_tabs_dev + off_coef[i]; }
    for (int i = 0; i < 4; i++) cviews[i]->coef = h->pq_tabs_dev + off_curve[i];
    AEAJ_CUDA(cudaMemcpy(h->pq_tabs_dev, flat.data(), flat.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->pq_worst_eps = worst;
    for (int sp = 0; sp < 8; sp++) {
        h->colors_host[sp].pq = H.view;
        h->colors_host[sp].fast = (sp == AEAJ_OKLAB || sp == AEAJ_ICACB || sp == AEAJ_ICTCP || sp == AEAJ_JZAZBZ) ? 1 : 0;
    }
    return 0;
}

int launch_color_pixels(aeaj_handle* h, int space, int inverse, const float* in, float* out, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    const ColorConsts& C = h->colors_host[space];
    const float* lut = h->has_srgb_lut ? h->srgb_lut_dev : nullptr;
    size_t nq = n / 4 > 0 ? n / 4 : 1;
    int blocks = (int)std::min<size_t>((nq + 255) / 256, (size_t)h->sm_count * 16);
    blocks = std::max(blocks, 1);
    return dispatch_space(space, [&](auto S) {
        constexpr int SP = decltype(S)::value;
        if (inverse) k_color_pixels<SP, true><<<blocks, 256, 0, st>>>(C, lut, in, out, n);
        else k_color_pixels<SP, false><<<blocks, 256, 0, st>>>(C, lut, in, out, n);
        AEAJ_LAUNCH_CHECK();
        return 0;
    });
}

int launch_area(const float* src, int H, int W, float* dst, int dh, int dw, uint8_t* u8_out, int planes,
                size_t src_stride, size_t dst_stride, cudaStream_t st) {
    dim3 blk(32, 8), grd(aeaj_cdiv(dw, 32), aeaj_cdiv(dh, 8), planes);
    k_area<<<grd, blk, 0, st>>>(src, H, W, dst, dh, dw, u8_out, src_stride, dst_stride);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int launch_resize_linear(const float* src, int sh, int sw, float* dst, int H, int W, cudaStream_t st) {
    dim3 blk(32, 8), grd(aeaj_cdiv(W, 32), aeaj_cdiv(H, 8));
    k_resize_linear<<<grd, blk, 0, st>>>(src, sh, sw, dst, H, W);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int launch_normalize(const float* in, float* out, size_t n, float mid, float scale, int inverse, cudaStream_t st) {
    if (n == 0) return 0;
    int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    k_normalize<<<blocks, 256, 0, st>>>(in, out, n, mid, scale, inverse);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int launch_cast_u8(const float* in, uint8_t* out, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    k_cast_u8<<<blocks, 256, 0, st>>>(in, out, n);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

// planes_host: [B*3] plane descriptors, plane index = b*3 + layer; layers of one kind are contiguous
// across the batch (stride = h*w), which is what FwdOut's per-image strides assume.
int launch_color_forward_planar(aeaj_handle* h, int space, const float* rgb, const uint8_t* rgb8, int B, int H, int W,
                                const PlaneDesc* planes_dev, const PlaneDesc* P, float* full_c1, float* full_c2,
                                cudaStream_t st, int* launches, int band0, int band1) {
    (void)planes_dev;
    if (band1 < 0) band1 = H;
    const bool banded = (band0 != 0 || band1 != H);
    const ColorConsts& C = h->colors_host[space];
    const float* lut = h->has_srgb_lut ? h->srgb_lut_dev : nullptr;
    const int ch = P[1].h, cw = P[1].w;
    int mode;
    if (ch * 2 == H && cw * 2 == W && (W % 4) == 0) mode = 0;
    else if (ch == H && cw * 4 == W) mode = 1;
    else mode = 2;
    if (banded) {
        // a band [band0, band1) of full-res rows is converted as an image of its own; the outputs land at the
        // band's row offset inside the full-size planes (needs the fused modes: no chroma cell straddles a band)
        if (mode == 2 || B != 1 || (mode == 0 && ((band0 | band1) & 1))) { aeaj_set_error("banded colour conversion needs batch 1, even band rows and W %% 4 == 0"); return AEAJ_EINVAL; }
    }
    FwdOut o;
    o.y = P[0].layer_f32; o.y8 = P[0].u8a; o.sy = (size_t)H * W;
    if (mode == 2) { o.c1 = full_c1; o.c2 = full_c2; o.c18 = nullptr; o.c28 = nullptr; o.sc = (size_t)H * W; }
    else { o.c1 = P[1].layer_f32; o.c2 = P[2].layer_f32; o.c18 = P[1].u8a; o.c28 = P[2].u8a; o.sc = (size_t)ch * cw; }
    int Hk = H;                                         // rows the kernel sees
    if (banded) {
        const size_t yo = (size_t)band0 * W, co = (mode == 0) ? (size_t)(band0 / 2) * cw : (size_t)band0 * cw;
        if (rgb) rgb += yo * 3; else rgb8 += yo * 3;
        o.y += yo; o.y8 += yo; o.c1 += co; o.c2 += co; o.c18 += co; o.c28 += co;
        Hk = band1 - band0;
    }
    int rc = dispatch_space(space, [&](auto S) {
        constexpr int SP = decltype(S)::value;
        if (mode == 0) {
            dim3 blk(32, 8), grd(aeaj_cdiv(W / 4, 32), aeaj_cdiv(Hk / 2, 8), B);
            if (rgb) k_color_forward_planar<SP, 0, false><<<grd, blk, 0, st>>>(C, lut, rgb, Hk, W, o);
            else k_color_forward_planar<SP, 0, true><<<grd, blk, 0, st>>>(C, lut, rgb8, Hk, W, o);
        } else if (mode == 1) {
            dim3 blk(32, 8), grd(aeaj_cdiv(W / 4, 32), aeaj_cdiv(Hk, 8), B);
            if (rgb) k_color_forward_planar<SP, 1, false><<<grd, blk, 0, st>>>(C, lut, rgb, Hk, W, o);
            else k_color_forward_planar<SP, 1, true><<<grd, blk, 0, st>>>(C, lut, rgb8, Hk, W, o);
        } else {
            dim3 blk(32, 8), grd(aeaj_cdiv(W, 32), aeaj_cdiv(H, 8), B);
            if (rgb) k_color_forward_planar<SP, 2, false><<<grd, blk, 0, st>>>(C, lut, rgb, H, W, o);
            else k_color_forward_planar<SP, 2, true><<<grd, blk, 0, st>>>(C, lut, rgb8, H, W, o);
        }
        AEAJ_LAUNCH_CHECK();
        return 0;
    });
    if (rc) return rc;
    (*launches)++;
    if (mode == 2) {
        rc = launch_area(full_c1, H, W, P[1].layer_f32, ch, cw, P[1].u8a, B, (size_t)H * W, (size_t)ch * cw, st);
        if (rc) return rc;
        rc = launch_area(full_c2, H, W, P[2].layer_f32, ch, cw, P[2].u8a, B, (size_t)H * W, (size_t)ch * cw, st);
        if (rc) return rc;
        (*launches) += 2;
    }
    return 0;
}

int launch_upsample_color_inverse(aeaj_handle* h, int space, const PlaneDesc* P, int B, int H, int W, float* rgb, uint8_t* rgb8, cudaStream_t st, int band0, int band1) {
    const ColorConsts& C = h->colors_host[space];
    if (band1 < 0) band1 = H;
    const int Hb = band1 - band0;
    UpIn in;
    for (int l = 0; l < 3; l++) {
        in.p[l] = P[l].layer_f32; in.h[l] = P[l].h; in.w[l] = P[l].w; in.stride[l] = (size_t)P[l].h * P[l].w;
        in.lo[l] = P[l].ry0; in.hi[l] = P[l].ry1;
    }
    in.peer_up = P[0].peer_up; in.peer_dn = P[0].peer_dn;
    const bool fast2x = (in.h[1] * 2 == H && in.w[1] * 2 == W && in.h[2] == in.h[1] && in.w[2] == in.w[1] && (W % 4) == 0 &&
                         in.h[0] == H && in.w[0] == W);
    return dispatch_space(space, [&](auto S) {
        constexpr int SP = decltype(S)::value;
        if (fast2x) {
            dim3 blk(32, 8), grd(aeaj_cdiv(W / 4, 32), aeaj_cdiv(Hb, 8), B);
            if (in.peer_up != 0 || in.peer_dn != 0) k_upsample2x_color_inverse<SP, true><<<grd, blk, 0, st>>>(C, in, H, W, rgb, rgb8, band0, band1);
            else k_upsample2x_color_inverse<SP, false><<<grd, blk, 0, st>>>(C, in, H, W, rgb, rgb8, band0, band1);
        } else {
            dim3 blk(32, 8), grd(aeaj_cdiv(W, 32), aeaj_cdiv(Hb, 8), B);
            k_upsample_color_inverse<SP><<<grd, blk, 0, st>>>(C, in, H, W, rgb, rgb8, band0, band1);
        }
        AEAJ_LAUNCH_CHECK();
        return 0;
    });
}
