/*
 * aeaj_oracle.c -- CPU restatement ("oracle") of the adaptive edge-aware JPEG hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (adaptive-edge-aware-jpeg_b200/)
 * may link, import or call this file.  It is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py as the checker and the timed CPU port.
 *
 * Parity status: PINNED.  Every stage below is checked against outputs of the reference
 * itself (run in the build container through oracle/ref_import.py) by
 * tests/golden/make_golden.py; the committed fixtures under tests/golden/ are those
 * reference outputs.  Oracle mode "S" (cv2.ipp.setUseIPP(False)) is the bit-exact target;
 * mode "D" (IPP on) differs only in the bilateral filter / DCT / bilinear resize, see DESIGN.md.
 *
 * The reference is pure Python; its arithmetic lives in un-vendored third-party wheels
 * (requirements.txt: opencv-python 4.11.0.86, numpy 2.1.3, numba 0.61.0).  Each function
 * cites the reference call site (file:line under /root/reference) and restates the published
 * algorithm of the library routine that call site invokes.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -mfma -fopenmp).  FMA is used only
 * where written explicitly (fmaf/fma), because the reference's results depend on it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define AO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * colour spaces (src/color)
 * ---------------------------------------------------------------------------------------- */
enum { AO_YCBCR = 0, AO_YCOCG = 1, AO_YCOCG_R = 2, AO_OKLAB = 3, AO_ICACB = 4, AO_ICTCP = 5,
       AO_JZAZBZ = 6, AO_XYZ = 7 };

/* All 3x3 matrices are passed in from Python (oracle/tables.py) as float32 so that the inverse
 * matrices are *exactly* the np.linalg.inv results the reference computes at import
 * (oklab.py:35,47; icacb.py:149,159; ictcp.py:149,159; jzazbz.py:196,206). */
typedef struct {
    float m1[9];      /* first 3x3  (forward: XYZ->LMS ; inverse: space->LMS')          */
    float m2[9];      /* second 3x3 (forward: LMS'->space ; inverse: LMS->XYZ)           */
    float rgb2xyz[9]; /* xyz.py:27-32 */
    float xyz2rgb[9]; /* xyz.py:35-40 */
    float lin[9];     /* linear spaces: the single 3x3 (ycbcr.py:25-38, ycocg.py:25-55)  */
} ao_color_tables;

AO_API void ao_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
AO_API int ao_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* np.dot((N,3) f32, M.T) == OpenBLAS sgemm with K=3: fma(x2,m2, fma(x1,m1, x0*m0)) in f32
 * (ycbcr.py:61,79; ycocg.py:82,100,121,139; xyz.py:64,82; oklab.py:72,74,93,95). */
static inline void dot3(const float *m, float x0, float x1, float x2, float *o) {
    for (int k = 0; k < 3; k++)
        o[k] = fmaf(x2, m[3 * k + 2], fmaf(x1, m[3 * k + 1], x0 * m[3 * k + 0]));
}

/* numba kernels (icacb.py:50-59, ictcp.py:50-59): f32*f32 products contracted by LLVM as
 * fma(M2,Z, fma(M1,Y, M0*X)) -- same association as dot3. */

/* common.py:34-60: constants are Python floats, so the arithmetic is f64, stored as f32. */
static inline float srgb_to_linear(float v) {
    double d = (double)v;
    if (d <= 0.04045) return (float)(d / 12.92);
    return (float)pow((d + 0.055) / 1.055, 2.4);
}
/* common.py:62-92 */
static inline float linear_to_srgb(float v) {
    double d = (double)v, r;
    if (d <= 0.0031308) r = d * 12.92;
    else r = 1.055 * pow(d, 1.0 / 2.4) - 0.055;
    float f = (float)r;
    /* max(0.0, min(1.0, srgb)) compiled under fastmath as selects: min(1,x) = x<1 ? x : 1, so a
     * NaN (negative base in a pow upstream) becomes 1.0 -- measured on the reference: a NaN in any
     * channel of the PQ stage turns the whole pixel white (class T-NAN). */
    f = (f < 1.0f) ? f : 1.0f;
    f = (f > 0.0f) ? f : 0.0f;
    return f;
}

/* common.py:131-159 (f64) */
static inline double pq_inv_eotf(double c, double m2) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0;
    const double m1 = 2610.0 / 16384.0;
    double t = pow(c / 10000.0, m1);
    return pow((c1 + c2 * t) / (1.0 + c3 * t), m2);
}
/* common.py:94-129 (f64) */
static inline double pq_eotf(double c, double m2) {
    const double c1 = 3424.0 / 4096.0, c2 = 2413.0 / 128.0, c3 = 2392.0 / 128.0;
    const double m1 = 2610.0 / 16384.0;
    double t = pow(c, 1.0 / m2); /* negative base -> NaN, propagates (T-NAN) */
    double num = t - c1, den = c2 - c3 * t;
    if (num < 0.0) num = 0.0;
    if (den <= 0.0) den = 1e-12;
    return 10000.0 * pow(num / den, 1.0 / m1);
}

static const double PQ_M2 = 2523.0 / 32.0;
static const double JZ_P = 1.7 * 2523.0 / 32.0;
static const double JZ_B = 1.15, JZ_G = 0.66, JZ_D = -0.56, JZ_D0 = 1.6295499532821566e-11;

/* numpy float32 power(x, 1/3) (oklab.py:73): the exponent is the f32 value 0.33333334f.
 * Restated as correctly rounded pow in f64 (tie class T-POW where numpy's SIMD powf is 1 ULP off). */
static inline float np_powf_third(float x) { return (float)pow((double)x, (double)(float)(1.0 / 3.0)); }
static inline float np_powf_cube(float x) { return (float)pow((double)x, 3.0); }

AO_API void ao_color_forward(int space, const ao_color_tables *T, const float *rgb, float *out,
                             size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const float *p = rgb + 3 * i;
        float *o = out + 3 * i;
        if (space <= AO_YCOCG_R) { dot3(T->lin, p[0], p[1], p[2], o); continue; }
        float lin[3], xyz[3];
        for (int k = 0; k < 3; k++) lin[k] = srgb_to_linear(p[k]);
        dot3(T->rgb2xyz, lin[0], lin[1], lin[2], xyz);
        if (space == AO_XYZ) { o[0] = xyz[0]; o[1] = xyz[1]; o[2] = xyz[2]; continue; }
        if (space == AO_OKLAB) { /* oklab.py:71-75 */
            float lms[3], lp[3];
            dot3(T->m1, xyz[0], xyz[1], xyz[2], lms);
            for (int k = 0; k < 3; k++) lp[k] = np_powf_third(lms[k]);
            dot3(T->m2, lp[0], lp[1], lp[2], o);
        } else if (space == AO_ICACB || space == AO_ICTCP) { /* ictcp.py:47-79 */
            float lms[3];
            double lp[3];
            dot3(T->m1, xyz[0], xyz[1], xyz[2], lms);
            for (int k = 0; k < 3; k++) lp[k] = pq_inv_eotf((double)lms[k], PQ_M2);
            for (int k = 0; k < 3; k++) {
                const float *m = T->m2 + 3 * k; /* f32 entries widened, f64 accumulate */
                o[k] = (float)((double)m[0] * lp[0] + (double)m[1] * lp[1] + (double)m[2] * lp[2]);
            }
        } else { /* JzAzBz: jzazbz.py:57-97 */
            double X = xyz[0], Y = xyz[1];
            double Xp = JZ_B * X - (JZ_B - 1.0) * (double)xyz[2];
            double Yp = JZ_G * Y - (JZ_G - 1.0) * X;
            float Zp = xyz[2]; /* stays f32 (jzazbz.py:63) */
            double lp[3];
            for (int k = 0; k < 3; k++) {
                const float *m = T->m1 + 3 * k;
                float mz = m[2] * Zp; /* f32 x f32 product rounded to f32 */
                double L = (double)m[0] * Xp + (double)m[1] * Yp + (double)mz;
                lp[k] = pq_inv_eotf(L, JZ_P);
            }
            double v[3];
            for (int k = 0; k < 3; k++) {
                const float *m = T->m2 + 3 * k;
                v[k] = (double)m[0] * lp[0] + (double)m[1] * lp[1] + (double)m[2] * lp[2];
            }
            double Jz = ((1.0 + JZ_D) * v[0]) / (1.0 + JZ_D * v[0]) - JZ_D0;
            o[0] = (float)Jz; o[1] = (float)v[1]; o[2] = (float)v[2];
        }
    }
}

AO_API void ao_color_inverse(int space, const ao_color_tables *T, const float *in, float *rgb,
                             size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const float *p = in + 3 * i;
        float *o = rgb + 3 * i;
        if (space <= AO_YCOCG_R) { /* ycbcr.py:79-82: dot then np.clip */
            float t[3];
            dot3(T->lin, p[0], p[1], p[2], t);
            for (int k = 0; k < 3; k++) o[k] = t[k] < 0.0f ? 0.0f : (t[k] > 1.0f ? 1.0f : t[k]);
            continue;
        }
        float xyz[3];
        if (space == AO_XYZ) { xyz[0] = p[0]; xyz[1] = p[1]; xyz[2] = p[2]; }
        else if (space == AO_OKLAB) { /* oklab.py:93-96 */
            float lp[3], lms[3];
            dot3(T->m1, p[0], p[1], p[2], lp);
            for (int k = 0; k < 3; k++) lms[k] = np_powf_cube(lp[k]);
            dot3(T->m2, lms[0], lms[1], lms[2], xyz);
        } else if (space == AO_ICACB || space == AO_ICTCP) { /* ictcp.py:103-137 */
            float lp[3];
            double l[3];
            dot3(T->m1, p[0], p[1], p[2], lp);
            for (int k = 0; k < 3; k++) l[k] = pq_eotf((double)lp[k], PQ_M2);
            for (int k = 0; k < 3; k++) {
                const float *m = T->m2 + 3 * k;
                xyz[k] = (float)((double)m[0] * l[0] + (double)m[1] * l[1] + (double)m[2] * l[2]);
            }
        } else { /* jzazbz.py:131-171 */
            double jd = (double)p[0] + JZ_D0;
            double Iz = jd / (1.0 + JZ_D - JZ_D * jd);
            double l[3];
            for (int k = 0; k < 3; k++) {
                const float *m = T->m1 + 3 * k;
                float ta = m[1] * p[1], tb = m[2] * p[2]; /* f32 x f32 products */
                double lp = (double)m[0] * Iz + (double)ta + (double)tb;
                l[k] = pq_eotf(lp, JZ_P);
            }
            double v[3];
            for (int k = 0; k < 3; k++) {
                const float *m = T->m2 + 3 * k;
                v[k] = (double)m[0] * l[0] + (double)m[1] * l[1] + (double)m[2] * l[2];
            }
            double X = (v[0] + (JZ_B - 1.0) * v[2]) / JZ_B;
            double Y = (v[1] + (JZ_G - 1.0) * X) / JZ_G;
            xyz[0] = (float)X; xyz[1] = (float)Y; xyz[2] = (float)v[2];
        }
        float lin[3];
        dot3(T->xyz2rgb, xyz[0], xyz[1], xyz[2], lin);
        for (int k = 0; k < 3; k++) o[k] = linear_to_srgb(lin[k]);
    }
}

/* common.py:161-189 via conversion.py:126-157 (call sites jpeg.py:387-390, 452-455):
 * normalise (d-m)*s ; denormalise n/s + m ; plain f32 two-op, no fma, true divide. */
AO_API void ao_normalize(const float *in, float *out, size_t n, float mid, float scale, int inverse) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        if (inverse) { float q = in[i] / scale; out[i] = q + mid; }
        else { float d = in[i] - mid; out[i] = d * scale; }
    }
}

/* interleaved (N,3) -> 3 planes (jpeg.py:263-264 reshape+transpose) */
AO_API void ao_deinterleave(const float *in, float *p0, float *p1, float *p2, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { p0[i] = in[3 * i]; p1[i] = in[3 * i + 1]; p2[i] = in[3 * i + 2]; }
}
AO_API void ao_interleave(const float *p0, const float *p1, const float *p2, float *out, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { out[3 * i] = p0[i]; out[3 * i + 1] = p1[i]; out[3 * i + 2] = p2[i]; }
}

/* ------------------------------------------------------------------------------------------
 * resampling (jpeg.py:323-354 -> cv.resize INTER_AREA / INTER_LINEAR)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int di, si; float alpha; } ao_dectab;

/* OpenCV computeResizeAreaTab (imgproc/resize.cpp) */
static int area_tab(int ssize, int dsize, double scale, ao_dectab *tab) {
    int k = 0;
    for (int dx = 0; dx < dsize; dx++) {
        double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        double cellWidth = fmin(scale, ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cellWidth); }
        for (int sx = sx1; sx < sx2; sx++) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cellWidth); }
        if (fsx2 - sx2 > 1e-3) {
            tab[k].di = dx; tab[k].si = sx2;
            tab[k++].alpha = (float)(fmin(fmin(fsx2 - sx2, 1.), cellWidth) / cellWidth);
        }
    }
    return k;
}

AO_API void ao_downsample_area(const float *src, int H, int W, float *dst, int h, int w) {
    if (h == H && w == W) { memcpy(dst, src, sizeof(float) * (size_t)H * W); return; }
    int ry = H / h, rx = W / w;
    if (ry * h == H && rx * w == W) {
        /* integer ratio: ResizeAreaFast -- sum in raster order over the ry x rx cell, * 1/(area) */
        float sc = 1.0f / (float)(rx * ry);
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const float *s = src + (size_t)(y * ry) * W + x * rx;
                float acc;
                if (rx == 2 && ry == 2) acc = (s[0] + s[1]) + (s[W] + s[W + 1]);
                else {
                    acc = 0.0f;
                    for (int j = 0; j < ry; j++)
                        for (int i = 0; i < rx; i++) acc += s[(size_t)j * W + i];
                }
                dst[(size_t)y * w + x] = acc * sc;
            }
        return;
    }
    /* general area path (ResizeArea_Invoker) */
    double sx = (double)W / w, sy = (double)H / h;
    ao_dectab *xt = malloc(sizeof(ao_dectab) * (size_t)(W * 2 + 4)), *yt = malloc(sizeof(ao_dectab) * (size_t)(H * 2 + 4));
    int nx = area_tab(W, w, sx, xt), ny = area_tab(H, h, sy, yt);
    float *buf = malloc(sizeof(float) * w), *sum = malloc(sizeof(float) * w);
    int prev_dy = yt[0].di;
    for (int x = 0; x < w; x++) sum[x] = 0.0f;
    for (int j = 0; j < ny; j++) {
        int dy = yt[j].di, syi = yt[j].si;
        float beta = yt[j].alpha;
        const float *S = src + (size_t)syi * W;
        for (int x = 0; x < w; x++) buf[x] = 0.0f;
        for (int k = 0; k < nx; k++) buf[xt[k].di] += S[xt[k].si] * xt[k].alpha;
        if (dy != prev_dy) {
            for (int x = 0; x < w; x++) { dst[(size_t)prev_dy * w + x] = sum[x]; sum[x] = beta * buf[x]; }
            prev_dy = dy;
        } else {
            for (int x = 0; x < w; x++) sum[x] += beta * buf[x];
        }
    }
    for (int x = 0; x < w; x++) dst[(size_t)prev_dy * w + x] = sum[x];
    free(xt); free(yt); free(buf); free(sum);
}

/* cv.resize INTER_LINEAR, float32, half-pixel centres, edge clamp; horizontal then vertical;
 * a*(1-t) + b*t in f32 (OpenCV open-source path == cv2.ipp.setUseIPP(False)). */
AO_API void ao_resize_linear(const float *src, int h, int w, float *dst, int H, int W) {
    if (h == H && w == W) { memcpy(dst, src, sizeof(float) * (size_t)H * W); return; }
    double scx = (double)w / W, scy = (double)h / H;
    int *xo = malloc(sizeof(int) * W); float *xa = malloc(sizeof(float) * W);
    for (int dx = 0; dx < W; dx++) {
        float fx = (float)((dx + 0.5) * scx - 0.5);
        int s = (int)floorf(fx);
        fx -= s;
        if (s < 0) { fx = 0; s = 0; }
        if (s >= w - 1) { fx = 0; s = w - 1; }
        xo[dx] = s; xa[dx] = fx;
    }
#pragma omp parallel for schedule(static)
    for (int dy = 0; dy < H; dy++) {
        float fy = (float)((dy + 0.5) * scy - 0.5);
        int s = (int)floorf(fy);
        fy -= s;
        if (s < 0) { fy = 0; s = 0; }
        if (s >= h - 1) { fy = 0; s = h - 1; }
        int s1 = s + 1 < h ? s + 1 : h - 1;
        const float *r0 = src + (size_t)s * w, *r1 = src + (size_t)s1 * w;
        float b0 = 1.f - fy, b1 = fy;
        for (int dx = 0; dx < W; dx++) {
            int x0 = xo[dx], x1 = x0 + 1 < w ? x0 + 1 : w - 1;
            float a0 = 1.f - xa[dx], a1 = xa[dx];
            float t0 = r0[x0] * a0 + r0[x1] * a1;
            float t1 = r1[x0] * a0 + r1[x1] * a1;
            dst[(size_t)dy * W + dx] = t0 * b0 + t1 * b1;
        }
    }
    free(xo); free(xa);
}

/* ------------------------------------------------------------------------------------------
 * Canny pipeline (edge_detection.py:70-86)
 * ---------------------------------------------------------------------------------------- */
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * (len - 1) - p; }
    return p;
}
static inline int cv_round_f(float v) { return (int)lrintf(v); } /* half-even */
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* edge_detection.py:70  (img*255).astype(np.uint8): f32 multiply, truncate toward zero, wrap mod 256 */
AO_API void ao_cast_u8(const float *layer, uint8_t *out, size_t n) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        float v = layer[i] * 255.0f;
        out[i] = (uint8_t)(int32_t)v; /* |v| < 2^31 on this path */
    }
}

/* edge_detection.py:73-74  cv.createCLAHE(0.75,(4,4)).apply  (OpenCV imgproc/clahe.cpp) */
AO_API void ao_clahe(const uint8_t *src, int h, int w, uint8_t *dst) {
    const int TX = 4, TY = 4;
    int eh = h, ew = w;
    if (w % TX != 0 || h % TY != 0) { eh = h + (TY - h % TY); ew = w + (TX - w % TX); }
    int th = eh / TY, tw = ew / TX, area = th * tw;
    float lutScale = 255.0f / (float)area;
    int clip = (int)(0.75 * area / 256);
    if (clip < 1) clip = 1;
    static _Thread_local uint8_t lut[16][256];
    for (int ty = 0; ty < TY; ty++)
        for (int tx = 0; tx < TX; tx++) {
            int hist[256];
            memset(hist, 0, sizeof hist);
            for (int y = ty * th; y < (ty + 1) * th; y++) {
                int sy = reflect101(y, h);
                for (int x = tx * tw; x < (tx + 1) * tw; x++) hist[src[(size_t)sy * w + reflect101(x, w)]]++;
            }
            int clipped = 0;
            for (int i = 0; i < 256; i++) if (hist[i] > clip) { clipped += hist[i] - clip; hist[i] = clip; }
            int redist = clipped / 256, residual = clipped - redist * 256;
            for (int i = 0; i < 256; i++) hist[i] += redist;
            if (residual != 0) {
                int step = 256 / residual; if (step < 1) step = 1;
                for (int i = 0; i < 256 && residual > 0; i += step, residual--) hist[i]++;
            }
            int sum = 0;
            for (int i = 0; i < 256; i++) { sum += hist[i]; lut[ty * TX + tx][i] = sat_u8(cv_round_f((float)sum * lutScale)); }
        }
    float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    for (int y = 0; y < h; y++) {
        float tyf = (float)y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > TY - 1) ty2 = TY - 1;
        for (int x = 0; x < w; x++) {
            float txf = (float)x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            float xa = txf - (float)tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > TX - 1) tx2 = TX - 1;
            int v = src[(size_t)y * w + x];
            float a = (float)lut[ty1 * TX + tx1][v] * xa1, b = (float)lut[ty1 * TX + tx2][v] * xa;
            float c = (float)lut[ty2 * TX + tx1][v] * xa1, d = (float)lut[ty2 * TX + tx2][v] * xa;
            float r0 = (a + b) * ya1, r1 = (c + d) * ya;
            dst[(size_t)y * w + x] = sat_u8(cv_round_f(r0 + r1));
        }
    }
}

/* edge_detection.py:77  cv.GaussianBlur(u8,(3,3),0): [1 2 1]x[1 2 1], REFLECT_101, (s+8)>>4 */
AO_API void ao_gauss3(const uint8_t *src, int h, int w, uint8_t *dst) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = src + (size_t)reflect101(y - 1, h) * w, *r1 = src + (size_t)y * w,
                      *r2 = src + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int s = (r0[xm] + 2 * r0[x] + r0[xp]) + 2 * (r1[xm] + 2 * r1[x] + r1[xp]) + (r2[xm] + 2 * r2[x] + r2[xp]);
            dst[(size_t)y * w + x] = (uint8_t)((s + 8) >> 4);
        }
    }
}

/* edge_detection.py:78  cv.bilateralFilter(u8, 5, 75, 75) -- OpenCV open-source path
 * (imgproc/bilateral_filter.dispatch.cpp + .simd.hpp), i.e. oracle mode S:
 *   radius 2, 13 taps with sqrt(i^2+j^2) <= 2 in raster order, REFLECT_101,
 *   w = space_w[k] * color_w[|v - v0|] (f32), sum = fma(v, w, sum), wsum += w, cvRound(sum / wsum). */
static float g_bil_color[256];
static float g_bil_space[13];
static int g_bil_dy[13], g_bil_dx[13], g_bil_init = 0;
static void bil_init(void) {
    if (g_bil_init) return;
    double cc = -0.5 / (75.0 * 75.0), sc = -0.5 / (75.0 * 75.0);
    for (int i = 0; i < 256; i++) g_bil_color[i] = (float)exp((double)(i * i) * cc);
    int k = 0;
    for (int i = -2; i <= 2; i++)
        for (int j = -2; j <= 2; j++) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > 2) continue;
            g_bil_space[k] = (float)exp(r * r * sc);
            g_bil_dy[k] = i; g_bil_dx[k] = j; k++;
        }
    g_bil_init = 1;
}
AO_API void ao_bilateral5(const uint8_t *src, int h, int w, uint8_t *dst, int use_fma) {
    bil_init();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int v0 = src[(size_t)y * w + x];
            float sum = 0.0f, wsum = 0.0f;
            for (int k = 0; k < 13; k++) {
                int v = src[(size_t)reflect101(y + g_bil_dy[k], h) * w + reflect101(x + g_bil_dx[k], w)];
                int d = v - v0; if (d < 0) d = -d;
                float wgt = g_bil_space[k] * g_bil_color[d];
                if (use_fma) sum = fmaf((float)v, wgt, sum);
                else { float t = (float)v * wgt; sum = sum + t; }
                wsum = wsum + wgt;
            }
            dst[(size_t)y * w + x] = sat_u8(cv_round_f(sum / wsum));
        }
}

/* edge_detection.py:81-82  np.percentile(u8, q) (method 'linear') from a 256-bin histogram */
static double percentile_from_hist(const uint64_t *hist, uint64_t n, double q) {
    double v = (double)(n - 1) * q;
    double lo = floor(v), g = v - lo;
    uint64_t ilo = (uint64_t)lo, ihi = ilo + 1 < n ? ilo + 1 : n - 1;
    int a = 0, b = 0; uint64_t c = 0; int fa = 0;
    for (int i = 0; i < 256; i++) {
        c += hist[i];
        if (!fa && c > ilo) { a = i; fa = 1; }
        if (c > ihi) { b = i; break; }
    }
    double d = (double)(b - a);
    return g < 0.5 ? (double)a + d * g : (double)b - d * (1.0 - g);
}
AO_API void ao_percentile_thresholds(const uint8_t *src, size_t n, double *lo, double *hi) {
    uint64_t hist[256];
    memset(hist, 0, sizeof hist);
    for (size_t i = 0; i < n; i++) hist[src[i]]++;
    /* canny_low_ratio*100 = 10.0, canny_high_ratio*100 = 30.0 -> q = 10/100, 30/100 */
    *lo = percentile_from_hist(hist, n, 10.0 / 100.0);
    *hi = percentile_from_hist(hist, n, 30.0 / 100.0);
}

/* edge_detection.py:85  cv.Canny(u8, lo, hi, apertureSize=3, L2gradient=True) (imgproc/canny.cpp).
 * Output 0/1 (the reference divides 255 by 255, edge_detection.py:86). */
AO_API void ao_canny(const uint8_t *src, int h, int w, double lo, double hi, uint8_t *edge) {
    if (lo > hi) { double t = lo; lo = hi; hi = t; }
    lo = fmin(32767.0, lo); hi = fmin(32767.0, hi);
    if (lo > 0) lo *= lo;
    if (hi > 0) hi *= hi;
    int low = (int)floor(lo), high = (int)floor(hi);
    size_t n = (size_t)h * w;
    int16_t *dx = malloc(n * 2), *dy = malloc(n * 2);
    int32_t *mag = malloc(n * 4);
    uint8_t *cls = malloc(n); /* 0 none, 1 weak candidate, 2 strong */
#define PX(yy, xx) ((int)src[(size_t)((yy) < 0 ? 0 : ((yy) >= h ? h - 1 : (yy))) * w + ((xx) < 0 ? 0 : ((xx) >= w ? w - 1 : (xx)))])
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int gx = (PX(y - 1, x + 1) + 2 * PX(y, x + 1) + PX(y + 1, x + 1)) - (PX(y - 1, x - 1) + 2 * PX(y, x - 1) + PX(y + 1, x - 1));
            int gy = (PX(y + 1, x - 1) + 2 * PX(y + 1, x) + PX(y + 1, x + 1)) - (PX(y - 1, x - 1) + 2 * PX(y - 1, x) + PX(y - 1, x + 1));
            size_t i = (size_t)y * w + x;
            dx[i] = (int16_t)gx; dy[i] = (int16_t)gy; mag[i] = gx * gx + gy * gy;
        }
#undef PX
#define MG(yy, xx) (((yy) < 0 || (yy) >= h || (xx) < 0 || (xx) >= w) ? 0 : mag[(size_t)(yy) * w + (xx)])
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = (size_t)y * w + x;
            int m = mag[i], keep = 0;
            if (m > low) {
                int xs = dx[i], ys = dy[i];
                int ax = abs(xs), ay = abs(ys) << 15;
                int tg22x = ax * 13573;
                if (ay < tg22x) keep = (m > MG(y, x - 1) && m >= MG(y, x + 1));
                else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) keep = (m > MG(y - 1, x) && m >= MG(y + 1, x));
                    else {
                        int s = (xs ^ ys) < 0 ? -1 : 1;
                        keep = (m > MG(y - 1, x - s) && m > MG(y + 1, x + s));
                    }
                }
            }
            cls[i] = keep ? (m > high ? 2 : 1) : 0;
        }
#undef MG
    /* hysteresis: flood from strong through weak, 8-connected */
    size_t *stack = malloc(sizeof(size_t) * (n + 1));
    size_t sp = 0;
    for (size_t i = 0; i < n; i++) if (cls[i] == 2) stack[sp++] = i;
    while (sp) {
        size_t i = stack[--sp];
        int y = (int)(i / w), x = (int)(i % w);
        for (int j = -1; j <= 1; j++)
            for (int k = -1; k <= 1; k++) {
                int yy = y + j, xx = x + k;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                size_t q = (size_t)yy * w + xx;
                if (cls[q] == 1) { cls[q] = 2; stack[sp++] = q; }
            }
    }
    for (size_t i = 0; i < n; i++) edge[i] = cls[i] == 2;
    free(dx); free(dy); free(mag); free(cls); free(stack);
}

/* the whole of EdgeDetection.canny on one layer; optional taps for stage-isolated parity */
AO_API void ao_canny_pipeline(const float *layer, int h, int w, uint8_t *edge, uint8_t *tap_u8,
                              uint8_t *tap_clahe, uint8_t *tap_gauss, uint8_t *tap_bil, double *tap_thr) {
    size_t n = (size_t)h * w;
    uint8_t *a = malloc(n), *b = malloc(n);
    ao_cast_u8(layer, a, n);
    if (tap_u8) memcpy(tap_u8, a, n);
    ao_clahe(a, h, w, b);
    if (tap_clahe) memcpy(tap_clahe, b, n);
    ao_gauss3(b, h, w, a);
    if (tap_gauss) memcpy(tap_gauss, a, n);
    ao_bilateral5(a, h, w, b, 1);
    if (tap_bil) memcpy(tap_bil, b, n);
    double lo, hi;
    ao_percentile_thresholds(b, n, &lo, &hi);
    if (tap_thr) { tap_thr[0] = lo; tap_thr[1] = hi; }
    ao_canny(b, h, w, lo, hi, edge);
    free(a); free(b);
}

/* ------------------------------------------------------------------------------------------
 * quadtree (quadtree.py:68-165, utils.py:24-41)
 * ---------------------------------------------------------------------------------------- */
AO_API int ao_root_size(int h, int w) {
    int n = h > w ? h : w;
    if (n <= 2) return n * 2;
    int p = 1;
    while (p * 2 < n) p *= 2; /* largest power of two strictly below n (utils.py:40-41) */
    return p * 2;
}

static int region_has_edge(const uint8_t *edge, int h, int w, int x, int y, int size) {
    int y1 = y + size < h ? y + size : h, x1 = x + size < w ? x + size : w;
    for (int yy = y; yy < y1; yy++)
        for (int xx = x; xx < x1; xx++)
            if (edge[(size_t)yy * w + xx]) return 1;
    return 0;
}

/* DFS pre-order, children TL,TR,BL,BR (quadtree.py:123-131,136-165).
 * leaves: (x,y,size) triples; states: 0 leaf, 1 split, 2 absent child. Returns root size. */
typedef struct { int x, y, size; } ao_node;
AO_API int ao_quadtree(const uint8_t *edge, int h, int w, int min_size, int max_size, int *leaves,
                       int *n_leaves, uint8_t *states, int *n_states) {
    int root = ao_root_size(h, w);
    size_t cap = 64, sp = 0;
    ao_node *st = malloc(sizeof(ao_node) * cap);
    st[sp++] = (ao_node){0, 0, root};
    int nl = 0, ns = 0;
    while (sp) {
        ao_node nd = st[--sp];
        if (nd.x >= w || nd.y >= h) { states[ns++] = 2; continue; }
        /* NB the root is always in bounds; out-of-bounds children are recorded as None (quadtree.py:109) */
        if (nd.size > max_size || (nd.size > min_size && region_has_edge(edge, h, w, nd.x, nd.y, nd.size))) {
            states[ns++] = 1;
            int hs = nd.size / 2;
            if (sp + 4 > cap) { cap *= 2; st = realloc(st, sizeof(ao_node) * cap); }
            st[sp++] = (ao_node){nd.x + hs, nd.y + hs, hs};
            st[sp++] = (ao_node){nd.x, nd.y + hs, hs};
            st[sp++] = (ao_node){nd.x + hs, nd.y, hs};
            st[sp++] = (ao_node){nd.x, nd.y, hs};
        } else {
            states[ns++] = 0;
            leaves[3 * nl] = nd.x; leaves[3 * nl + 1] = nd.y; leaves[3 * nl + 2] = nd.size; nl++;
        }
    }
    free(st);
    *n_leaves = nl; *n_states = ns;
    return root;
}

/* jpeg.py:768-800 + 428-448: states -> leaf (x,y,size) by size-matching DFS */
AO_API int ao_states_to_leaves(const uint8_t *states, int n_states, int root, int h, int w, int *leaves) {
    size_t cap = 64, sp = 0;
    ao_node *st = malloc(sizeof(ao_node) * cap);
    st[sp++] = (ao_node){0, 0, root};
    int nl = 0, si = 0;
    while (sp && si < n_states) {
        ao_node nd = st[--sp];
        int s = states[si++];
        if (s == 0) { leaves[3 * nl] = nd.x; leaves[3 * nl + 1] = nd.y; leaves[3 * nl + 2] = nd.size; nl++; }
        else if (s == 1) {
            int hs = nd.size / 2;
            if (sp + 4 > cap) { cap *= 2; st = realloc(st, sizeof(ao_node) * cap); }
            st[sp++] = (ao_node){nd.x + hs, nd.y + hs, hs};
            st[sp++] = (ao_node){nd.x, nd.y + hs, hs};
            st[sp++] = (ao_node){nd.x + hs, nd.y, hs};
            st[sp++] = (ao_node){nd.x, nd.y, hs};
        }
    }
    (void)h; (void)w;
    free(st);
    return nl;
}

/* ------------------------------------------------------------------------------------------
 * DCT / quantiser (jpeg.py:461-529)
 * ---------------------------------------------------------------------------------------- */
static double *dct_matrix(int s) { /* C[k][i] = sqrt(2/s) cos(pi (2i+1) k / 2s), row 0 / sqrt 2 */
    double *c = malloc(sizeof(double) * s * s);
    for (int k = 0; k < s; k++)
        for (int i = 0; i < s; i++) {
            double v = sqrt(2.0 / s) * cos(M_PI * (2 * i + 1) * k / (2.0 * s));
            if (k == 0) v *= sqrt(0.5);
            c[k * s + i] = v;
        }
    return c;
}
static double *g_dct[16];
static double *get_dct(int s) {
    int l = 0; while ((1 << l) < s) l++;
    double *c;
#pragma omp critical(ao_dct_tab)
    { if (!g_dct[l]) g_dct[l] = dct_matrix(s); c = g_dct[l]; }
    return c;
}
/* cv.dct(block) (jpeg.py:471): orthonormal 2-D DCT-II = C X C^T; f64 accumulate, f32 result */
AO_API void ao_dct2d(const float *x, int s, float *y, int inverse) {
    const double *c = get_dct(s);
    double *t = malloc(sizeof(double) * s * s);
    if (!inverse) {
        for (int k = 0; k < s; k++) for (int j = 0; j < s; j++) { double a = 0; for (int i = 0; i < s; i++) a += c[k * s + i] * (double)x[i * s + j]; t[k * s + j] = a; }
        for (int k = 0; k < s; k++) for (int l = 0; l < s; l++) { double a = 0; for (int j = 0; j < s; j++) a += t[k * s + j] * c[l * s + j]; y[k * s + l] = (float)a; }
    } else { /* cv.idct (jpeg.py:483): C^T Y C */
        for (int i = 0; i < s; i++) for (int l = 0; l < s; l++) { double a = 0; for (int k = 0; k < s; k++) a += c[k * s + i] * (double)x[k * s + l]; t[i * s + l] = a; }
        for (int i = 0; i < s; i++) for (int j = 0; j < s; j++) { double a = 0; for (int l = 0; l < s; l++) a += t[i * s + l] * c[l * s + j]; y[i * s + j] = (float)a; }
    }
    free(t);
}

/* jpeg.py:393-404: slice + np.pad(mode='reflect') of partial leaves (period 2(n-1); n==1 replicates) */
static inline int pad_reflect(int p, int n) {
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    p %= period;
    return p < n ? p : period - p;
}

/* encode one layer given leaves: normalise -> extract -> DCT -> quantise
 * (jpeg.py:387-404, 471, 497-504). qtabs[k] is the s x s int32 table for size 2^k. */
AO_API void ao_encode_blocks(const float *layer, int h, int w, float mid, float scale, const int *leaves,
                             int n_leaves, const int32_t *const *qtabs, int32_t *coef, float *dct_tap) {
    size_t *off = malloc(sizeof(size_t) * (size_t)(n_leaves + 1));
    off[0] = 0;
    for (int i = 0; i < n_leaves; i++) off[i + 1] = off[i] + (size_t)leaves[3 * i + 2] * leaves[3 * i + 2];
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n_leaves; i++) {
        int x = leaves[3 * i], y = leaves[3 * i + 1], s = leaves[3 * i + 2];
        int bh = h - y < s ? h - y : s, bw = w - x < s ? w - x : s;
        int lg = 0; while ((1 << lg) < s) lg++;
        float *blk = malloc(sizeof(float) * s * s * 2), *d = blk + s * s;
        for (int r = 0; r < s; r++)
            for (int c = 0; c < s; c++) {
                float v = layer[(size_t)(y + pad_reflect(r, bh)) * w + (x + pad_reflect(c, bw))];
                float t = v - mid;
                blk[r * s + c] = t * scale;
            }
        ao_dct2d(blk, s, d, 0);
        const int32_t *q = qtabs[lg];
        for (int k = 0; k < s * s; k++) {
            if (dct_tap) dct_tap[off[i] + k] = d[k];
            coef[off[i] + k] = (int32_t)rint((double)d[k] / (double)q[k]); /* f64 divide, half-even */
        }
        free(blk);
    }
    free(off);
}

/* decode one layer: dequantise -> IDCT -> place -> crop -> denormalise (jpeg.py:508-529, 483, 410-459) */
AO_API void ao_decode_blocks(const int32_t *coef, const int *leaves, int n_leaves, const int32_t *const *qtabs,
                             int h, int w, float mid, float scale, float *layer) {
    size_t *off = malloc(sizeof(size_t) * (size_t)(n_leaves + 1));
    off[0] = 0;
    for (int i = 0; i < n_leaves; i++) off[i + 1] = off[i] + (size_t)leaves[3 * i + 2] * leaves[3 * i + 2];
    /* zero canvas (jpeg.py:425-426) denormalised; leaves tile the in-bounds area so this is
     * overwritten everywhere, kept for fidelity */
    { float z = 0.0f / scale; z = z + mid; for (size_t k = 0; k < (size_t)h * w; k++) layer[k] = z; }
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < n_leaves; i++) {
        int x = leaves[3 * i], y = leaves[3 * i + 1], s = leaves[3 * i + 2];
        int lg = 0; while ((1 << lg) < s) lg++;
        const int32_t *q = qtabs[lg];
        float *blk = malloc(sizeof(float) * s * s * 2), *d = blk + s * s;
        for (int k = 0; k < s * s; k++) blk[k] = (float)(coef[off[i] + k] * q[k]);
        ao_dct2d(blk, s, d, 1);
        for (int r = 0; r < s && y + r < h; r++)
            for (int c = 0; c < s && x + c < w; c++) {
                float t = d[r * s + c] / scale;
                layer[(size_t)(y + r) * w + (x + c)] = t + mid;
            }
        free(blk);
    }
    free(off);
}
