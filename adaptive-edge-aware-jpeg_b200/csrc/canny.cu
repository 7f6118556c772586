// canny.cu -- EdgeDetection.canny (edge_detection.py:28-86) for sm_100a.
//
//   u8 cast (color.cu) -> CLAHE tile histograms -> CLAHE LUTs -> [CLAHE apply + Gaussian 3x3 +
//   bilateral 5x5] fused, shared-memory tiled with halo staging -> 256-bin histogram -> percentile
//   thresholds -> Sobel/magnitude/NMS/double threshold -> bit-packed strong/weak maps (warp ballot)
//   -> hysteresis by iterative bitmap frontier propagation (cooperative grid, tile-local convergence
//   in shared memory, border exchange through global bitmaps between rounds).
//
// All of it is integer / u8 work bounded by HBM (or L2) bandwidth and launch latency, not by math.
// Arithmetic follows SURVEY.md App. A3 as validated by the CPU oracle against the reference.
#include <cooperative_groups.h>
#include "aeaj_internal.cuh"

namespace cg = cooperative_groups;

namespace {

// bilateral tables: cv.bilateralFilter(d=5, sigmaColor=75, sigmaSpace=75) -- OpenCV builds them with
// std::exp in double and stores float; done on the host with the same libm (aeaj_create).
__constant__ float c_bil_color[256];
__constant__ float c_bil_space[13];
__constant__ int c_bil_dy[13];
__constant__ int c_bil_dx[13];

// ---------------------------------------------------------------------------------------------
// CLAHE (OpenCV imgproc/clahe.cpp), clip 0.75, 4x4 tiles
// ---------------------------------------------------------------------------------------------
struct ClaheGeom { int th, tw, clip; float lutScale, inv_th, inv_tw; };
__host__ __device__ inline ClaheGeom clahe_geom(int h, int w) {
    int eh = h, ew = w;
    if ((w % 4) != 0 || (h % 4) != 0) { eh = h + (4 - h % 4); ew = w + (4 - w % 4); }
    ClaheGeom g;
    g.th = eh / 4; g.tw = ew / 4;
    int area = g.th * g.tw;
    g.lutScale = 255.0f / (float)area;
    int clip = (int)(0.75 * area / 256);
    g.clip = clip < 1 ? 1 : clip;
    g.inv_th = 1.0f / (float)g.th; g.inv_tw = 1.0f / (float)g.tw;
    return g;
}

constexpr int HIST_ROWS = 32;   // rows of one CLAHE tile handled by one block
// 16 consecutive pixels from one 128-bit load, equal neighbours merged into one shared-memory atomic
__device__ __forceinline__ void hist_add16(unsigned int* wh, uint4 q) {
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
    int prev = w[0] & 0xff, cnt = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int v = (w[k >> 2] >> (8 * (k & 3))) & 0xff;
        if (v == prev) cnt++;
        else { atomicAdd(&wh[prev], (unsigned)cnt); prev = v; cnt = 1; }
    }
    atomicAdd(&wh[prev], (unsigned)cnt);
}
// grid: (row chunks, 16 tiles, planes).  Each warp walks rows of the tile with 128-bit loads over the
// 16-byte-aligned middle of the row segment; head / tail / reflected padding go byte by byte.
__global__ void __launch_bounds__(256) k_clahe_hist(const PlaneDesc* __restrict__ planes) {
    const PlaneDesc& P = planes[blockIdx.z];
    ClaheGeom g = clahe_geom(P.h, P.w);
    const int tile = blockIdx.y, ty = tile >> 2, tx = tile & 3;
    const int r0 = blockIdx.x * HIST_ROWS;
    if (r0 >= g.th) return;
    __shared__ unsigned int wh[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&wh[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r1 = min(r0 + HIST_ROWS, g.th);
    const uint8_t* src = P.u8a;
    unsigned int* my = wh[warp];
    const int cx0 = tx * g.tw, cx1 = cx0 + g.tw;
    for (int r = r0 + warp; r < r1; r += 8) {
        const int sy = reflect101(ty * g.th + r, P.h);
        const uint8_t* rowp = src + (size_t)sy * P.w;
        const int xa = cx0, xb = min(cx1, P.w);
        if (xa < xb) {
            const int head = min((int)((16 - ((uintptr_t)(rowp + xa) & 15)) & 15), xb - xa);
            const int nvec = (xb - xa - head) >> 4;
            const int tail0 = xa + head + nvec * 16;
            for (int x = xa + lane; x < xa + head; x += 32) atomicAdd(&my[rowp[x]], 1u);
            for (int x = tail0 + lane; x < xb; x += 32) atomicAdd(&my[rowp[x]], 1u);
            const uint4* vp = reinterpret_cast<const uint4*>(rowp + xa + head);
            for (int c = lane; c < nvec; c += 32) hist_add16(my, __ldg(vp + c));
        }
        for (int x = max(cx0, P.w) + lane; x < cx1; x += 32) atomicAdd(&my[rowp[reflect101(x, P.w)]], 1u);
    }
    __syncthreads();
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += wh[k][threadIdx.x];
    if (s) atomicAdd(&P.clahe_hist[tile * 256 + threadIdx.x], s);
}

// grid: (16 tiles, planes), 256 threads = bins
__global__ void __launch_bounds__(256) k_clahe_lut(const PlaneDesc* __restrict__ planes) {
    const PlaneDesc& P = planes[blockIdx.y];
    ClaheGeom g = clahe_geom(P.h, P.w);
    const int tile = blockIdx.x, i = threadIdx.x;
    __shared__ int red[256];
    int hv = (int)P.clahe_hist[tile * 256 + i];
    P.clahe_hist[tile * 256 + i] = 0;                 // leave the accumulator clean for the next call
    int excess = hv > g.clip ? hv - g.clip : 0;
    hv = hv > g.clip ? g.clip : hv;
    red[i] = excess;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (i < s) red[i] += red[i + s]; __syncthreads(); }
    int clipped = red[0];
    __syncthreads();
    int redist = clipped / 256, residual = clipped - redist * 256;
    hv += redist;
    if (residual != 0) {
        int step = 256 / residual; if (step < 1) step = 1;
        if ((i % step) == 0 && (i / step) < residual) hv++;
    }
    // inclusive scan
    red[i] = hv;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        int t = (i >= off) ? red[i - off] : 0;
        __syncthreads();
        red[i] += t;
        __syncthreads();
    }
    int v = __float2int_rn(__fmul_rn((float)red[i], g.lutScale));
    P.clahe_lut[tile * 256 + i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__device__ __forceinline__ uint8_t clahe_apply_px(const uint8_t (*lut)[256], const ClaheGeom& g, int y, int x, int v) {
    float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0); ty2 = min(ty2, 3);
    float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
    int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
    float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0); tx2 = min(tx2, 3);
    float a = __fmul_rn((float)lut[ty1 * 4 + tx1][v], xa1), b = __fmul_rn((float)lut[ty1 * 4 + tx2][v], xa);
    float c = __fmul_rn((float)lut[ty2 * 4 + tx1][v], xa1), d = __fmul_rn((float)lut[ty2 * 4 + tx2][v], xa);
    float r = __fadd_rn(__fmul_rn(__fadd_rn(a, b), ya1), __fmul_rn(__fadd_rn(c, d), ya));
    int iv = __float2int_rn(r);
    return (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
}

// ---------------------------------------------------------------------------------------------
// fused pre-filter: [CLAHE apply] -> [Gaussian 3x3] -> [bilateral d=5] on a 64x32 tile.
// Halo cells hold the value at the REFLECT_101-folded coordinate, so each stage sees exactly the
// border OpenCV builds for it (proof sketch in DESIGN.md: the fold is a local isometry and the
// Gaussian is symmetric; the bilateral is evaluated at in-image pixels only).
// stages: bit0 CLAHE apply, bit1 Gaussian, bit2 bilateral.   grid: (tiles x, tiles y, planes)
// ---------------------------------------------------------------------------------------------
constexpr int PF_TW = 64, PF_TH = 32;
constexpr int PF_AS = PF_TW + 8;                     // row stride of both staging tiles (4-byte aligned rows)
// tap k of the 13-tap bilateral window in raster order -> (dy, dx, weight class r^2 in {0,1,2,4} -> 0..3)
#define BIL_TAPS(X) X(0,-2,0,3) X(1,-1,-1,2) X(2,-1,0,1) X(3,-1,1,2) X(4,0,-2,3) X(5,0,-1,1) X(6,0,0,0) X(7,0,1,1) X(8,0,2,3) \
                    X(9,1,-1,2) X(10,1,0,1) X(11,1,1,2) X(12,2,0,3)

__global__ void __launch_bounds__(256) k_prefilter(const PlaneDesc* __restrict__ planes, int stages, int do_hist) {
    const PlaneDesc& P = planes[blockIdx.z];
    const int X0 = blockIdx.x * PF_TW, Y0 = blockIdx.y * PF_TH;
    if (X0 >= P.w || Y0 >= P.h) return;
    __shared__ __align__(16) uint8_t sA[PF_TH + 6][PF_AS];     // CLAHE output (or source), halo 3
    __shared__ __align__(16) uint8_t sG[PF_TH + 4][PF_AS];     // Gaussian output, halo 2
    __shared__ __align__(16) uint8_t sLut[16][256];
    __shared__ float sW[4][256];                               // space weight (r^2 = 0,1,2,4) x colour weight, rounded product
    __shared__ unsigned int sHist[256];
    __shared__ float sXa[PF_AS], sYa[PF_TH + 6];               // CLAHE interpolation weights per tile column / row
    __shared__ uint8_t sTx[PF_AS][2], sTy[PF_TH + 6][2];       // CLAHE tile indices per tile column / row
    __shared__ int sFx[PF_AS], sFy[PF_TH + 6];                 // REFLECT_101-folded source coordinates
    const int tid = threadIdx.x;
    const ClaheGeom g = clahe_geom(P.h, P.w);
    if (tid < PF_AS) {
        const int x = reflect101(X0 + tid - 3, P.w);
        sFx[tid] = x;
        const float txf = __fsub_rn(__fmul_rn((float)x, g.inv_tw), 0.5f);
        const int t1 = (int)floorf(txf);
        sXa[tid] = __fsub_rn(txf, (float)t1);
        sTx[tid][0] = (uint8_t)max(t1, 0); sTx[tid][1] = (uint8_t)min(t1 + 1, 3);
    } else if (tid >= 128 && tid < 128 + PF_TH + 6) {
        const int r = tid - 128;
        const int y = reflect101(Y0 + r - 3, P.h);
        sFy[r] = y;
        const float tyf = __fsub_rn(__fmul_rn((float)y, g.inv_th), 0.5f);
        const int t1 = (int)floorf(tyf);
        sYa[r] = __fsub_rn(tyf, (float)t1);
        sTy[r][0] = (uint8_t)(max(t1, 0) * 4); sTy[r][1] = (uint8_t)(min(t1 + 1, 3) * 4);
    }
    if (stages & 1)
        for (int i = tid; i < 16 * 256 / 4; i += 256) reinterpret_cast<uint32_t*>(&sLut[0][0])[i] = reinterpret_cast<const uint32_t*>(P.clahe_lut)[i];
    {
        float cw = c_bil_color[tid];
        sW[0][tid] = cw;                                       // centre tap: space weight exp(0) = 1
        sW[1][tid] = __fmul_rn(c_bil_space[2], cw);            // r^2 = 1
        sW[2][tid] = __fmul_rn(c_bil_space[1], cw);            // r^2 = 2
        sW[3][tid] = __fmul_rn(c_bil_space[0], cw);            // r^2 = 4
        sHist[tid] = 0;
    }
    __syncthreads();
    const uint8_t* src = P.u8a;
    // stage A: source (folded coordinates) -> CLAHE; the staging region is 38 rows x 72 columns (2 spare
    // columns keep the index arithmetic to shifts; they hold valid folded pixels and are never read)
    for (int i = tid; i < (PF_TH + 6) * PF_AS; i += 256) {
        const int ry = i / PF_AS, rx = i - ry * PF_AS;
        const int v = src[(size_t)sFy[ry] * P.w + sFx[rx]];
        int outv = v;
        if (stages & 1) {
            const float xa = sXa[rx], xa1 = __fsub_rn(1.0f, xa), ya = sYa[ry], ya1 = __fsub_rn(1.0f, ya);
            const uint8_t* l0 = &sLut[sTy[ry][0]][v];
            const uint8_t* l1 = &sLut[sTy[ry][1]][v];
            const int c0 = sTx[rx][0] * 256, c1 = sTx[rx][1] * 256;
            const float a = __fmul_rn((float)l0[c0], xa1), b = __fmul_rn((float)l0[c1], xa);
            const float c = __fmul_rn((float)l1[c0], xa1), d = __fmul_rn((float)l1[c1], xa);
            const float r = __fadd_rn(__fmul_rn(__fadd_rn(a, b), ya1), __fmul_rn(__fadd_rn(c, d), ya));
            outv = min(max(__float2int_rn(r), 0), 255);
        }
        sA[ry][rx] = (uint8_t)outv;
    }
    __syncthreads();
    // stage B: Gaussian [1 2 1]x[1 2 1], 4 outputs per thread sharing the 6 column sums
    for (int i = tid; i < (PF_TH + 4) * 17; i += 256) {
        const int ry = i / 17, gx = (i - ry * 17) * 4;            // output cells (ry, gx..gx+3) of the halo-2 region
        uchar4 o;
        if (stages & 2) {
            int cs[6];
#pragma unroll
            for (int j = 0; j < 6; j++) cs[j] = sA[ry][gx + j] + 2 * sA[ry + 1][gx + j] + sA[ry + 2][gx + j];
            o.x = (uint8_t)((cs[0] + 2 * cs[1] + cs[2] + 8) >> 4); o.y = (uint8_t)((cs[1] + 2 * cs[2] + cs[3] + 8) >> 4);
            o.z = (uint8_t)((cs[2] + 2 * cs[3] + cs[4] + 8) >> 4); o.w = (uint8_t)((cs[3] + 2 * cs[4] + cs[5] + 8) >> 4);
        } else {
            o = make_uchar4(sA[ry + 1][gx + 1], sA[ry + 1][gx + 2], sA[ry + 1][gx + 3], sA[ry + 1][gx + 4]);
        }
        *reinterpret_cast<uchar4*>(&sG[ry][gx]) = o;
    }
    __syncthreads();
    // stage C: bilateral, thread -> 4 consecutive px in rows (tid/16) and (tid/16 + 16)
    uint8_t* dst = P.u8b;
    const int tx = (tid & 15) * 4;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int ty = (tid >> 4) + half * 16;
        const int y = Y0 + ty;
        if (y >= P.h) continue;
        uint8_t outv[4];
        if (stages & 4) {
            // 5 rows x 8 bytes window: rows ty..ty+4, cols tx..tx+7 of sG
            int win[5][8];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uchar4 a = *reinterpret_cast<const uchar4*>(&sG[ty + r][tx]);
                const uchar4 b = *reinterpret_cast<const uchar4*>(&sG[ty + r][tx + 4]);
                win[r][0] = a.x; win[r][1] = a.y; win[r][2] = a.z; win[r][3] = a.w;
                win[r][4] = b.x; win[r][5] = b.y; win[r][6] = b.z; win[r][7] = b.w;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int v0 = win[2][k + 2];
                float sum = 0.0f, wsum = 0.0f;
#define BIL_ONE(idx, dy, dx, cls) { const int v = win[2 + (dy)][k + 2 + (dx)]; const float wgt = sW[cls][__sad(v, v0, 0u)]; \
                                    sum = __fmaf_rn((float)v, wgt, sum); wsum = __fadd_rn(wsum, wgt); }
                BIL_TAPS(BIL_ONE)
#undef BIL_ONE
                const int res = __float2int_rn(__fdiv_rn(sum, wsum));
                outv[k] = (uint8_t)min(max(res, 0), 255);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) outv[k] = sG[ty + 2][tx + k + 2];
        }
        const int x = X0 + tx;
        uint8_t* drow = dst + (size_t)y * P.w;
        if (x + 3 < P.w && ((P.w & 3) == 0)) *reinterpret_cast<uchar4*>(drow + x) = make_uchar4(outv[0], outv[1], outv[2], outv[3]);
        else {
#pragma unroll
            for (int k = 0; k < 4; k++) if (x + k < P.w) drow[x + k] = outv[k];
        }
        if (do_hist) {
            int prev = -1, cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (x + k < P.w) {
                    if (outv[k] == prev) cnt++;
                    else { if (cnt) atomicAdd(&sHist[prev], (unsigned)cnt); prev = outv[k]; cnt = 1; }
                }
            }
            if (cnt) atomicAdd(&sHist[prev], (unsigned)cnt);
        }
    }
    if (do_hist) {
        __syncthreads();
        if (sHist[tid]) atomicAdd(&P.hist[tid], sHist[tid]);
    }
}

// plain 256-bin histogram of a u8 plane (stage API: aeaj_percentile_thresholds on its own)
__global__ void __launch_bounds__(256) k_hist_u8(const uint8_t* __restrict__ src, size_t n, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) atomicAdd(&sh[src[i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// np.percentile(img, 10 / 30) (method 'linear') from the histogram, then cv.Canny's threshold prep
// (imgproc/canny.cpp: L2gradient -> squared, clamped to 32767, floor).  One thread per plane.
__device__ double percentile_from_hist(const unsigned int* hist, unsigned long long n, double q) {
    double v = (double)(n - 1) * q;
    double lo = floor(v), gfrac = v - lo;
    unsigned long long ilo = (unsigned long long)lo, ihi = ilo + 1 < n ? ilo + 1 : n - 1;
    int a = 0, b = 0, fa = 0;
    unsigned long long c = 0;
    for (int i = 0; i < 256; i++) {
        c += hist[i];
        if (!fa && c > ilo) { a = i; fa = 1; }
        if (c > ihi) { b = i; break; }
    }
    double d = (double)(b - a);
    return gfrac < 0.5 ? (double)a + d * gfrac : (double)b - d * (1.0 - gfrac);
}
__device__ void canny_prepare_thresholds(double lo, double hi, int* thr) {
    if (lo > hi) { double t = lo; lo = hi; hi = t; }
    lo = fmin(32767.0, lo); hi = fmin(32767.0, hi);
    if (lo > 0) lo *= lo;
    if (hi > 0) hi *= hi;
    thr[0] = (int)floor(lo); thr[1] = (int)floor(hi);
}
__global__ void k_thresholds(const PlaneDesc* __restrict__ planes, int nplanes) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nplanes) return;
    const PlaneDesc& P = planes[p];
    unsigned long long n = (unsigned long long)P.h * P.w;
    double lo = percentile_from_hist(P.hist, n, 10.0 / 100.0), hi = percentile_from_hist(P.hist, n, 30.0 / 100.0);
    if (P.thr_d) { P.thr_d[0] = lo; P.thr_d[1] = hi; }
    canny_prepare_thresholds(lo, hi, P.thr);
    for (int i = 0; i < 256; i++) P.hist[i] = 0;     // clean accumulator for the next call
}
__global__ void k_thresholds_from_double(const double* thr_d, int* thr) { canny_prepare_thresholds(thr_d[0], thr_d[1], thr); }

// ---------------------------------------------------------------------------------------------
// Sobel 3x3 (BORDER_REPLICATE) + L2 magnitude + NMS + double threshold -> strong/weak bitmaps.
// tile 64x32, 8 warps; each warp classifies 32 consecutive pixels of a row and ballots them into
// one bitmap word.   grid: (tiles x, tiles y, planes)
// ---------------------------------------------------------------------------------------------
constexpr int NM_TW = 64, NM_TH = 32;
__global__ void __launch_bounds__(256) k_canny_nms(const PlaneDesc* __restrict__ planes) {
    const PlaneDesc& P = planes[blockIdx.z];
    const int X0 = blockIdx.x * NM_TW, Y0 = blockIdx.y * NM_TH;
    if (X0 >= P.w || Y0 >= P.h) return;
    __shared__ uint8_t sS[NM_TH + 4][NM_TW + 4];
    __shared__ int sM[NM_TH + 2][NM_TW + 2];
    __shared__ short sDx[NM_TH + 2][NM_TW + 2];
    __shared__ short sDy[NM_TH + 2][NM_TW + 2];
    const int tid = threadIdx.x;
    const uint8_t* src = P.u8b;
    for (int i = tid; i < (NM_TH + 4) * (NM_TW + 4); i += 256) {
        int ry = i / (NM_TW + 4), rx = i - ry * (NM_TW + 4);
        int y = clampi(Y0 + ry - 2, 0, P.h - 1), x = clampi(X0 + rx - 2, 0, P.w - 1);
        sS[ry][rx] = src[(size_t)y * P.w + x];
    }
    __syncthreads();
    for (int i = tid; i < (NM_TH + 2) * (NM_TW + 2); i += 256) {
        int ry = i / (NM_TW + 2), rx = i - ry * (NM_TW + 2);
        int y = Y0 + ry - 1, x = X0 + rx - 1;
        int gx = 0, gy = 0, m = 0;
        if (y >= 0 && y < P.h && x >= 0 && x < P.w) {
            const uint8_t* r0 = &sS[ry][rx]; const uint8_t* r1 = &sS[ry + 1][rx]; const uint8_t* r2 = &sS[ry + 2][rx];
            gx = (r0[2] + 2 * r1[2] + r2[2]) - (r0[0] + 2 * r1[0] + r2[0]);
            gy = (r2[0] + 2 * r2[1] + r2[2]) - (r0[0] + 2 * r0[1] + r0[2]);
            m = gx * gx + gy * gy;
        }
        sM[ry][rx] = m; sDx[ry][rx] = (short)gx; sDy[ry][rx] = (short)gy;
    }
    __syncthreads();
    const int low = P.thr[0], high = P.thr[1];
    const int warp = tid >> 5, lane = tid & 31;
    for (int seg = warp; seg < NM_TH * 2; seg += 8) {
        const int ty = seg >> 1, hx = (seg & 1) * 32;
        const int y = Y0 + ty, x = X0 + hx + lane;
        int cls = 0;
        if (y < P.h && x < P.w) {
            const int ry = ty + 1, rx = hx + lane + 1;
            const int m = sM[ry][rx];
            if (m > low) {
                const int xs = sDx[ry][rx], ys = sDy[ry][rx];
                const int ax = abs(xs), ay = abs(ys) << 15;
                const int tg22x = ax * 13573;
                bool keep;
                if (ay < tg22x) keep = (m > sM[ry][rx - 1]) && (m >= sM[ry][rx + 1]);
                else {
                    const int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) keep = (m > sM[ry - 1][rx]) && (m >= sM[ry + 1][rx]);
                    else {
                        const int s = ((xs ^ ys) < 0) ? -1 : 1;
                        keep = (m > sM[ry - 1][rx - s]) && (m > sM[ry + 1][rx + s]);
                    }
                }
                if (keep) cls = (m > high) ? 2 : 1;
            }
        }
        unsigned bs = __ballot_sync(0xffffffffu, cls == 2), bw = __ballot_sync(0xffffffffu, cls == 1);
        const int word = (X0 + hx) >> 5;
        if (lane == 0 && y < P.h && word < P.wpr) {
            P.strong[(size_t)y * P.wpr + word] = bs;
            P.weak[(size_t)y * P.wpr + word] = bw;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// hysteresis: strong |= weak pixels 8-connected (through weak pixels) to a strong pixel.
// Bit-packed frontier propagation.  A block owns a tile of 8 words x 32 rows (256 x 32 px), one
// word per thread; it iterates to local convergence in shared memory.  A cooperative grid loop
// repeats until no tile's border changed; tiles whose neighbours did not change are skipped.
// ---------------------------------------------------------------------------------------------
constexpr int HY_WW = 8, HY_TR = 32;                 // tile = 256 px x 32 rows, one word per thread (256 threads)
constexpr int HY_THREADS = HY_WW * HY_TR;
struct HystTileMap { int nplanes; int ntiles; };

__device__ __forceinline__ unsigned spread3(unsigned L, unsigned Cw, unsigned R) {
    return Cw | (Cw << 1) | (Cw >> 1) | (L >> 31) | (R << 31);
}

__global__ void __launch_bounds__(HY_THREADS) k_hysteresis(const PlaneDesc* __restrict__ planes, int nplanes, const int* __restrict__ tile_base,
                                                        int ntiles, int* __restrict__ flags, int* __restrict__ ctrl, int* __restrict__ status, int max_rounds) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned sS[HY_TR + 2][HY_WW + 2];
    __shared__ int sPlane, sAnyBorder;
    const int tid = threadIdx.x, tr = tid / HY_WW, tc = tid % HY_WW;
    int round = 0;
    for (;; round++) {
        int* fl_cur = flags + (size_t)(round & 1) * ntiles;
        int* fl_nxt = flags + (size_t)((round + 1) & 1) * ntiles;
        int* c_nxt = ctrl + (round + 1) % 3;
        if (blockIdx.x == 0 && tid == 0) ctrl[(round + 2) % 3] = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            if (round > 0 && __ldcg(fl_cur + tile) == 0) continue;   // uniform per block
            __syncthreads();
            if (tid == 0) {
                fl_cur[tile] = 0;
                int p = 0;
                while (p + 1 < nplanes && tile_base[p + 1] <= tile) p++;
                sPlane = p; sAnyBorder = 0;
            }
            __syncthreads();
            const PlaneDesc& P = planes[sPlane];
            const int local = tile - tile_base[sPlane];
            const int ntx = aeaj_cdiv(P.wpr, HY_WW);
            const int tyi = local / ntx, txi = local - tyi * ntx;
            const int gy0 = tyi * HY_TR, gw0 = txi * HY_WW;
            // load strong words incl. 1-row / 1-word halo
            for (int i = tid; i < (HY_TR + 2) * (HY_WW + 2); i += HY_THREADS) {
                int ry = i / (HY_WW + 2), rw = i - ry * (HY_WW + 2);
                int gy = gy0 + ry - 1, gw = gw0 + rw - 1;
                unsigned v = 0;
                if (gy >= 0 && gy < P.h && gw >= 0 && gw < P.wpr) v = __ldcg(P.strong + (size_t)gy * P.wpr + gw);
                sS[ry][rw] = v;
            }
            const int gy = gy0 + tr, gw = gw0 + tc;
            const bool valid = gy < P.h && gw < P.wpr;
            const unsigned wk = valid ? P.weak[(size_t)gy * P.wpr + gw] : 0u;
            __syncthreads();
            const unsigned s_init = sS[tr + 1][tc + 1];
            unsigned s = s_init;
            for (;;) {
                unsigned n = spread3(sS[tr][tc], sS[tr][tc + 1], sS[tr][tc + 2]) |
                             spread3(sS[tr + 1][tc], s, sS[tr + 1][tc + 2]) |
                             spread3(sS[tr + 2][tc], sS[tr + 2][tc + 1], sS[tr + 2][tc + 2]);
                unsigned add = wk & ~s & n;
                unsigned s_new = s | add;
                // flood along the row inside the word
                while (add) { add = wk & ~s_new & ((s_new << 1) | (s_new >> 1)); s_new |= add; }
                int ch = (s_new != s);
                s = s_new;
                __syncthreads();
                if (ch) sS[tr + 1][tc + 1] = s;
                if (!__syncthreads_or(ch)) break;
            }
            if (valid && s != s_init) {
                P.strong[(size_t)gy * P.wpr + gw] = s;
                unsigned diff = s ^ s_init;
                bool border = (tr == 0) || (tr == HY_TR - 1) || (gy == P.h - 1) || (tc == 0 && (diff & 1u)) || (tc == HY_WW - 1 && (diff >> 31));
                if (border) sAnyBorder = 1;
            }
            __syncthreads();
            if (sAnyBorder && tid < 9 && tid != 4) {
                int dy = tid / 3 - 1, dx = tid % 3 - 1;
                int nty = aeaj_cdiv(P.h, HY_TR);
                int ny = tyi + dy, nx = txi + dx;
                if (ny >= 0 && ny < nty && nx >= 0 && nx < ntx) {
                    fl_nxt[tile_base[sPlane] + ny * ntx + nx] = 1;
                    atomicAdd(c_nxt, 1);
                }
            }
        }
        __threadfence();
        grid.sync();
        int pending = *((volatile int*)c_nxt);
        if (pending == 0 || round + 1 >= max_rounds) {
            if (blockIdx.x == 0 && tid == 0 && status) { status[0] = round + 1; status[1] = (pending == 0); }
            break;
        }
    }
}

// final strong bitmap -> uint8 {0,1} map (API / taps) and the reverse (stage API)
__global__ void __launch_bounds__(256) k_bitmap_to_u8(const PlaneDesc* __restrict__ planes, uint8_t* const* __restrict__ outs) {
    const PlaneDesc& P = planes[blockIdx.z];
    uint8_t* out = outs[blockIdx.z];
    int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (!out || y >= P.h || x >= P.w) return;
    out[(size_t)y * P.w + x] = (P.strong[(size_t)y * P.wpr + (x >> 5)] >> (x & 31)) & 1u;
}
__global__ void __launch_bounds__(256) k_u8_to_bitmap(const uint8_t* __restrict__ edge, int h, int w, int wpr, uint32_t* __restrict__ bits) {
    int word = blockIdx.x * 8 + (threadIdx.x >> 5), y = blockIdx.y, lane = threadIdx.x & 31;
    if (word >= wpr) return;
    int x = word * 32 + lane;
    unsigned b = __ballot_sync(0xffffffffu, x < w && edge[(size_t)y * w + x] == 1);
    if (lane == 0) bits[(size_t)y * wpr + word] = b;
}

int max_dim(const PlaneDesc* P, int n, bool width) {
    int m = 0;
    for (int i = 0; i < n; i++) m = std::max(m, width ? P[i].w : P[i].h);
    return m;
}

}  // namespace

int aeaj_canny_init_constants() {
    float color[256], space[13];
    int dy[13], dx[13];
    double cc = -0.5 / (75.0 * 75.0), sc = -0.5 / (75.0 * 75.0);
    for (int i = 0; i < 256; i++) color[i] = (float)exp((double)(i * i) * cc);
    int k = 0;
    for (int i = -2; i <= 2; i++)
        for (int j = -2; j <= 2; j++) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > 2) continue;
            space[k] = (float)exp(r * r * sc); dy[k] = i; dx[k] = j; k++;
        }
    AEAJ_CUDA(cudaMemcpyToSymbol(c_bil_color, color, sizeof color));
    AEAJ_CUDA(cudaMemcpyToSymbol(c_bil_space, space, sizeof space));
    AEAJ_CUDA(cudaMemcpyToSymbol(c_bil_dy, dy, sizeof dy));
    AEAJ_CUDA(cudaMemcpyToSymbol(c_bil_dx, dx, sizeof dx));
    return 0;
}

int launch_clahe_hist(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, cudaStream_t st) {
    int maxth = 0;
    for (int i = 0; i < nplanes; i++) maxth = std::max(maxth, clahe_geom(P[i].h, P[i].w).th);
    dim3 grd(aeaj_cdiv(maxth, HIST_ROWS), 16, nplanes);
    k_clahe_hist<<<grd, 256, 0, st>>>(planes_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_clahe_lut(const PlaneDesc* planes_dev, int nplanes, cudaStream_t st) {
    k_clahe_lut<<<dim3(16, nplanes), 256, 0, st>>>(planes_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_prefilter(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, int stages, int do_hist, cudaStream_t st) {
    dim3 grd(aeaj_cdiv(max_dim(P, nplanes, true), PF_TW), aeaj_cdiv(max_dim(P, nplanes, false), PF_TH), nplanes);
    k_prefilter<<<grd, 256, 0, st>>>(planes_dev, stages, do_hist);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_hist_u8(const uint8_t* src, size_t n, unsigned int* hist, cudaStream_t st) {
    int blocks = (int)std::min<size_t>((n + 4095) / 4096, 148 * 8);
    k_hist_u8<<<std::max(blocks, 1), 256, 0, st>>>(src, n, hist);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_thresholds(const PlaneDesc* planes_dev, int nplanes, cudaStream_t st) {
    k_thresholds<<<aeaj_cdiv(nplanes, 64), 64, 0, st>>>(planes_dev, nplanes);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_thresholds_from_double(const double* thr_d, int* thr, cudaStream_t st) {
    k_thresholds_from_double<<<1, 1, 0, st>>>(thr_d, thr);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_canny_nms(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, cudaStream_t st) {
    dim3 grd(aeaj_cdiv(max_dim(P, nplanes, true), NM_TW), aeaj_cdiv(max_dim(P, nplanes, false), NM_TH), nplanes);
    k_canny_nms<<<grd, 256, 0, st>>>(planes_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int hysteresis_tiles(const PlaneDesc* P, int nplanes, int* tile_base_host) {
    int n = 0;
    for (int i = 0; i < nplanes; i++) { tile_base_host[i] = n; n += aeaj_cdiv(P[i].wpr, HY_WW) * aeaj_cdiv(P[i].h, HY_TR); }
    return n;
}

// flags: int[2*ntiles]; ctrl: int[3]; both zeroed here.  tile_base_dev: int[nplanes]
int launch_hysteresis(aeaj_handle* h, const PlaneDesc* planes_dev, int nplanes, const int* tile_base_dev, int ntiles,
                      int* flags, int* ctrl, int* status, cudaStream_t st) {
    AEAJ_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * 2 * (size_t)ntiles, st));
    AEAJ_CUDA(cudaMemsetAsync(ctrl, 0, sizeof(int) * 3, st));
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        AEAJ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_hysteresis, HY_THREADS, 0));
        if (blocks_per_sm < 1) { aeaj_set_error("hysteresis kernel cannot be resident"); return AEAJ_EINVAL; }
    }
    int grid = std::min(ntiles, blocks_per_sm * h->sm_count);
    int max_rounds = 1 << 20;
    void* args[] = {(void*)&planes_dev, (void*)&nplanes, (void*)&tile_base_dev, (void*)&ntiles, (void*)&flags, (void*)&ctrl, (void*)&status, (void*)&max_rounds};
    AEAJ_CUDA(cudaLaunchCooperativeKernel((void*)k_hysteresis, dim3(grid), dim3(HY_THREADS), args, 0, st));
    return 0;
}

int launch_bitmap_to_u8(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, uint8_t* const* outs_dev, cudaStream_t st) {
    dim3 grd(aeaj_cdiv(max_dim(P, nplanes, true), 256), max_dim(P, nplanes, false), nplanes);
    k_bitmap_to_u8<<<grd, 256, 0, st>>>(planes_dev, outs_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
int launch_u8_to_bitmap(const uint8_t* edge, int h, int w, uint32_t* bits, cudaStream_t st) {
    int wpr = aeaj_cdiv(w, 32);
    k_u8_to_bitmap<<<dim3(aeaj_cdiv(wpr, 8), h), 256, 0, st>>>(edge, h, w, wpr, bits);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
