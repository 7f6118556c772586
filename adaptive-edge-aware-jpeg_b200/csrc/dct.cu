// dct.cu -- variable-size 2-D DCT/IDCT with fused (de)normalisation, reflect-padded leaf
// extraction, quantisation and dequantisation (jpeg.py:387-404, 461-529, 410-459) for sm_100a.
//
// Work arrives as per-size-class lists of leaves (quadtree.cu).  Both directions are the same
// product  Out = M . In . M^T  with M = C (forward, cv.dct) or M = C^T (inverse, cv.idct), where
// C[k][i] = sqrt(2/s) cos(pi (2i+1) k / 2s), row 0 scaled by 1/sqrt 2, evaluated in f64 on the host
// and stored as f32.  FP32 FMA accumulation: SURVEY.md App. A5 measured 0 quantiser flips against
// cv.dct for an f32 matrix DCT (tie class T-DCT), which is the parity bar; tensor-core (tf32) tiles do
// not hold that bar without error compensation and are not used here.
//
//   s <= 32 : "row" kernel -- s lanes per leaf, lane r owns row r; pass 1 in registers against
//             broadcast reads of M from shared memory, exchange through a per-warp shared tile,
//             pass 2 with the lane's own M row in registers; stores are fully coalesced rows.
//   s >= 64 : "cta" kernel -- one CTA per leaf, two register-blocked shared-memory GEMMs
//             (V = M.In, Out = V.M^T) with the intermediate kept transposed in shared memory.
// Quantise: rint(double(coef)/double(q)) half-even (np.round of an f64 quotient, jpeg.py:501);
// dequantise: float(int32*int32) (jpeg.py:524); normalise (v-m)*s, denormalise n/s+m, no fma.
#include "aeaj_internal.cuh"

namespace {

__device__ __forceinline__ int quantize(float z, int q) {
    return __double2int_rn(__ddiv_rn((double)z, (double)q));
}

// ---------------------------------------------------------------------------------------------
// row kernel, S in {2,4,8,16,32}; block = 256 threads = 8 warps; each warp handles 32/S leaves.
// ---------------------------------------------------------------------------------------------
template <int S, bool INVERSE>
__global__ void __launch_bounds__(256) k_dct_rows(const PlaneDesc* __restrict__ planes, const ClassEntry* __restrict__ list,
                                                  const int* __restrict__ count_ptr, const float* __restrict__ Mg, const float* __restrict__ Mtg) {
    constexpr int LPW = 32 / S;                        // leaves per warp
    constexpr int LG = (S == 2) ? 1 : (S == 4) ? 2 : (S == 8) ? 3 : (S == 16) ? 4 : 5;
    __shared__ __align__(16) float sM[S * S];          // M   (row-major)
    __shared__ __align__(16) float sMt[S * S];         // M^T (row-major)
    // exchange tiles: row stride TS (16 B aligned, spreads the pass-1 column stores over banks) and
    // a per-leaf stride GS chosen so that the LPW leaves of a warp start in different bank groups
    constexpr int TS = (S >= 8) ? S + 4 : S;
    constexpr int GS = S * TS + ((LPW > 1) ? ((32 / LPW) - (S * TS) % 32 + 32) % 32 : 0);
    __shared__ __align__(16) float sT[8][LPW * GS + 4];
    for (int i = threadIdx.x; i < S * S; i += 256) { sM[i] = Mg[i]; sMt[i] = Mtg[i]; }
    __syncthreads();
    const int count = *count_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / S, r = lane % S;            // leaf slot in warp, row in leaf
    float mrow[S];                                     // M[r][*] for pass 2
#pragma unroll
    for (int i = 0; i < S; i++) mrow[i] = sMt[i * S + r];
    const int groups = (count + LPW - 1) / LPW;        // warp-sized groups of leaves
    for (int g = blockIdx.x * 8 + warp; g < groups; g += gridDim.x * 8) {
        const int li = g * LPW + sub;
        const bool act = li < count;
        ClassEntry e = {0, 0, 0, 0};
        if (act) e = list[li];
        const PlaneDesc& P = planes[e.plane];
        float in[S];
        if (act) {
            if (!INVERSE) {
                const int bh = min(S, P.h - e.y), bw = min(S, P.w - e.x);
                const float* row = P.layer_f32 + (size_t)(e.y + pad_reflect(r, bh)) * P.w + e.x;
                const float mid = P.mid, sc = P.scale;
                if (bw == S && S >= 4 && ((P.w & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < S; j += 4) {
                        float4 v = __ldg(reinterpret_cast<const float4*>(row + j));
                        in[j] = v.x; in[j + 1] = v.y; in[j + 2] = v.z; in[j + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < S; j++) in[j] = __ldg(row + pad_reflect(j, bw));
                }
#pragma unroll
                for (int j = 0; j < S; j++) in[j] = __fmul_rn(__fsub_rn(in[j], mid), sc);
            } else {
                const int* cf = P.coef + (size_t)e.coef_off + r * S;
                const int* qt = P.qtab[LG] + r * S;
                if (S >= 4) {
#pragma unroll
                    for (int j = 0; j < S; j += 4) {
                        int4 cv = __ldg(reinterpret_cast<const int4*>(cf + j)), qv = __ldg(reinterpret_cast<const int4*>(qt + j));
                        in[j] = (float)(cv.x * qv.x); in[j + 1] = (float)(cv.y * qv.y); in[j + 2] = (float)(cv.z * qv.z); in[j + 3] = (float)(cv.w * qv.w);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < S; j++) in[j] = (float)(__ldg(cf + j) * __ldg(qt + j));
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < S; j++) in[j] = 0.0f;
        }
        // pass 1: t[l] = sum_j in[j] * M[l][j]
        float* T = &sT[warp][sub * GS];
#pragma unroll
        for (int l = 0; l < S; l++) {
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < S; j++) acc = __fmaf_rn(in[j], sM[l * S + j], acc);
            T[r * TS + l] = acc;
        }
        __syncwarp();
        // pass 2: out[l] = sum_i M[r][i] * T[i][l]
        float out[S];
#pragma unroll
        for (int l = 0; l < S; l++) out[l] = 0.0f;
#pragma unroll
        for (int i = 0; i < S; i++) {
#pragma unroll
            for (int l = 0; l < S; l++) out[l] = __fmaf_rn(mrow[i], T[i * TS + l], out[l]);
        }
        __syncwarp();
        if (act) {
            if (!INVERSE) {
                int* cf = P.coef + (size_t)e.coef_off + r * S;
                const int* qt = P.qtab[LG] + r * S;
                if (S >= 4) {
#pragma unroll
                    for (int l = 0; l < S; l += 4) {
                        int4 qv = __ldg(reinterpret_cast<const int4*>(qt + l));
                        *reinterpret_cast<int4*>(cf + l) = make_int4(quantize(out[l], qv.x), quantize(out[l + 1], qv.y), quantize(out[l + 2], qv.z), quantize(out[l + 3], qv.w));
                    }
                } else {
#pragma unroll
                    for (int l = 0; l < S; l++) cf[l] = quantize(out[l], __ldg(qt + l));
                }
            } else {
                const int y = e.y + r;
                if (y < P.h) {
                    float* row = P.layer_f32 + (size_t)y * P.w + e.x;
                    const int bw = min(S, P.w - e.x);
                    const float mid = P.mid, sc = P.scale;
                    if (bw == S && S >= 4 && ((P.w & 3) == 0)) {
#pragma unroll
                        for (int l = 0; l < S; l += 4)
                            *reinterpret_cast<float4*>(row + l) = make_float4(__fadd_rn(__fdiv_rn(out[l], sc), mid), __fadd_rn(__fdiv_rn(out[l + 1], sc), mid),
                                                                              __fadd_rn(__fdiv_rn(out[l + 2], sc), mid), __fadd_rn(__fdiv_rn(out[l + 3], sc), mid));
                    } else {
#pragma unroll
                        for (int l = 0; l < S; l++)
                            if (l < bw) row[l] = __fadd_rn(__fdiv_rn(out[l], sc), mid);
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// CTA kernel, S in {64,128}: 256 threads, (S/16 x S/16) register tile per thread.
//   pass 1: V = A1 . In   with A1^T staged in sA (sA[i][k] = M[k][i] = M^T row-major)
//   pass 2: Out = V . M^T with V^T staged in sB (aliasing In) and B2 = M^T = sA again
// ---------------------------------------------------------------------------------------------
template <int S, bool INVERSE>
__global__ void __launch_bounds__(256) k_dct_cta(const PlaneDesc* __restrict__ planes, const ClassEntry* __restrict__ list,
                                                 const int* __restrict__ count_ptr, const float* __restrict__ Mtg) {
    constexpr int TM = S / 16, TN = S / 16;
    constexpr int LG = (S == 64) ? 6 : 7;
    extern __shared__ __align__(16) float smem[];
    float* sA = smem;                                  // M^T row-major: sA[i*S + k] = M[k][i]
    float* sB = smem + S * S;                          // In (row-major), later V^T
    const int tid = threadIdx.x;
    for (int i = tid; i < S * S / 4; i += 256) reinterpret_cast<float4*>(sA)[i] = __ldg(reinterpret_cast<const float4*>(Mtg) + i);
    const int count = *count_ptr;
    const int tr = tid / 16, tc = tid % 16;            // thread tile: rows tr*4 + 64*c + (0..3), cols tc*4 + 64*c + (0..3)
    auto ridx = [&](int a) { return tr * 4 + (a & 3) + (a >> 2) * 64; };
    auto cidx = [&](int b) { return tc * 4 + (b & 3) + (b >> 2) * 64; };
    for (int li = blockIdx.x; li < count; li += gridDim.x) {
        const ClassEntry e = list[li];
        const PlaneDesc& P = planes[e.plane];
        __syncthreads();                               // previous leaf done with sB
        if (!INVERSE) {
            const int bh = min(S, P.h - e.y), bw = min(S, P.w - e.x);
            const float mid = P.mid, sc = P.scale;
            for (int i = tid; i < S * S; i += 256) {
                int rr = i / S, cc = i - rr * S;
                float v = __ldg(P.layer_f32 + (size_t)(e.y + pad_reflect(rr, bh)) * P.w + e.x + pad_reflect(cc, bw));
                sB[i] = __fmul_rn(__fsub_rn(v, mid), sc);
            }
        } else {
            const int* cf = P.coef + (size_t)e.coef_off;
            const int* qt = P.qtab[LG];
            for (int i = tid; i < S * S; i += 256) sB[i] = (float)(__ldg(cf + i) * __ldg(qt + i));
        }
        __syncthreads();
        float acc[TM][TN];
#pragma unroll
        for (int a = 0; a < TM; a++)
#pragma unroll
            for (int b = 0; b < TN; b++) acc[a][b] = 0.0f;
        // pass 1: V[k][j] = sum_i M[k][i] In[i][j]
        for (int i = 0; i < S; i++) {
            float av[TM], bv[TN];
#pragma unroll
            for (int a = 0; a < TM; a += 4) { float4 t = *reinterpret_cast<const float4*>(sA + i * S + ridx(a)); av[a] = t.x; av[a + 1] = t.y; av[a + 2] = t.z; av[a + 3] = t.w; }
#pragma unroll
            for (int b = 0; b < TN; b += 4) { float4 t = *reinterpret_cast<const float4*>(sB + i * S + cidx(b)); bv[b] = t.x; bv[b + 1] = t.y; bv[b + 2] = t.z; bv[b + 3] = t.w; }
#pragma unroll
            for (int a = 0; a < TM; a++)
#pragma unroll
                for (int b = 0; b < TN; b++) acc[a][b] = __fmaf_rn(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();                               // everyone finished reading In
        // store V transposed: sB[j*S + k] = V[k][j]
#pragma unroll
        for (int b = 0; b < TN; b++)
#pragma unroll
            for (int a = 0; a < TM; a += 4)
                *reinterpret_cast<float4*>(sB + cidx(b) * S + ridx(a)) = make_float4(acc[a][b], acc[a + 1][b], acc[a + 2][b], acc[a + 3][b]);
        __syncthreads();
#pragma unroll
        for (int a = 0; a < TM; a++)
#pragma unroll
            for (int b = 0; b < TN; b++) acc[a][b] = 0.0f;
        // pass 2: Out[k][l] = sum_j V[k][j] M[l][j] = sum_j sB[j][k] * sA[j][l]
        for (int j = 0; j < S; j++) {
            float av[TM], bv[TN];
#pragma unroll
            for (int a = 0; a < TM; a += 4) { float4 t = *reinterpret_cast<const float4*>(sB + j * S + ridx(a)); av[a] = t.x; av[a + 1] = t.y; av[a + 2] = t.z; av[a + 3] = t.w; }
#pragma unroll
            for (int b = 0; b < TN; b += 4) { float4 t = *reinterpret_cast<const float4*>(sA + j * S + cidx(b)); bv[b] = t.x; bv[b + 1] = t.y; bv[b + 2] = t.z; bv[b + 3] = t.w; }
#pragma unroll
            for (int a = 0; a < TM; a++)
#pragma unroll
                for (int b = 0; b < TN; b++) acc[a][b] = __fmaf_rn(av[a], bv[b], acc[a][b]);
        }
        if (!INVERSE) {
            int* cf = P.coef + (size_t)e.coef_off;
            const int* qt = P.qtab[LG];
#pragma unroll
            for (int a = 0; a < TM; a++) {
                const int k = ridx(a);
#pragma unroll
                for (int b = 0; b < TN; b += 4) {
                    const int l = cidx(b);
                    int4 qv = __ldg(reinterpret_cast<const int4*>(qt + k * S + l));
                    int4 o = make_int4(quantize(acc[a][b], qv.x), quantize(acc[a][b + 1], qv.y), quantize(acc[a][b + 2], qv.z), quantize(acc[a][b + 3], qv.w));
                    *reinterpret_cast<int4*>(cf + k * S + l) = o;
                }
            }
        } else {
            const float mid = P.mid, sc = P.scale;
#pragma unroll
            for (int a = 0; a < TM; a++) {
                const int y = e.y + ridx(a);
                if (y >= P.h) continue;
#pragma unroll
                for (int b = 0; b < TN; b++) {
                    const int x = e.x + cidx(b);
                    if (x < P.w) P.layer_f32[(size_t)y * P.w + x] = __fadd_rn(__fdiv_rn(acc[a][b], sc), mid);
                }
            }
        }
    }
}

template <int S, bool INV>
int launch_rows(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* list, const int* count, int64_t cap, cudaStream_t st) {
    constexpr int LPW = 32 / S;
    int64_t groups = aeaj_cdiv64(cap, LPW);
    int blocks = (int)std::min<int64_t>(aeaj_cdiv64(groups, 8), (int64_t)h->sm_count * 8);
    if (blocks < 1) blocks = 1;
    const int lg = ilog2i(S);
    k_dct_rows<S, INV><<<blocks, 256, 0, st>>>(planes_dev, list, count, INV ? h->dct_dev[lg] + S * S : h->dct_dev[lg],
                                               INV ? h->dct_dev[lg] : h->dct_dev[lg] + S * S);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
template <int S, bool INV>
int launch_cta(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* list, const int* count, int64_t cap, cudaStream_t st) {
    const size_t smem = 2 * (size_t)S * S * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        AEAJ_CUDA(cudaFuncSetAttribute(k_dct_cta<S, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int per_sm = (smem > 100 * 1024) ? 1 : 4;
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cap, 1), (int64_t)h->sm_count * per_sm);
    const int lg = ilog2i(S);
    // forward: M = C  -> M^T = C^T (second half of the table); inverse: M = C^T -> M^T = C (first half)
    k_dct_cta<S, INV><<<blocks, 256, smem, st>>>(planes_dev, list, count, INV ? h->dct_dev[lg] : h->dct_dev[lg] + S * S);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

template <bool INV>
int launch_all(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* class_lists, const int* class_counts,
               const int64_t* off, const int64_t* caps, int lg_min, int lg_max, cudaStream_t st, int* launches,
               void (*mark)(void*, const char*), void* mark_ctx) {
    static const char* fwd_names[9] = {"", "dct_quant_2", "dct_quant_4", "dct_quant_8", "dct_quant_16", "dct_quant_32", "dct_quant_64", "dct_quant_128", ""};
    static const char* inv_names[9] = {"", "dequant_idct_2", "dequant_idct_4", "dequant_idct_8", "dequant_idct_16", "dequant_idct_32", "dequant_idct_64", "dequant_idct_128", ""};
    for (int lg = lg_min; lg <= lg_max; lg++) {
        if (caps[lg] <= 0) continue;
        const ClassEntry* list = class_lists + off[lg];
        const int* cnt = class_counts + lg;
        int rc = 0;
        switch (lg) {
            case 1: rc = launch_rows<2, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            case 2: rc = launch_rows<4, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            case 3: rc = launch_rows<8, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            case 4: rc = launch_rows<16, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            case 5: rc = launch_rows<32, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            case 6: rc = launch_cta<64, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            case 7: rc = launch_cta<128, INV>(h, planes_dev, list, cnt, caps[lg], st); break;
            default: aeaj_set_error("block size %d not supported (2..128)", 1 << lg); return AEAJ_EINVAL;
        }
        if (rc) return rc;
        if (launches) (*launches)++;
        if (mark) mark(mark_ctx, INV ? inv_names[lg] : fwd_names[lg]);
    }
    return 0;
}

}  // namespace

// DCT-II matrices C and C^T per size, f64 on the host, stored f32: table for size s at dct_dev[log2 s],
// layout [C (s*s)][C^T (s*s)].
int aeaj_dct_init(aeaj_handle* h) {
    size_t total = 0;
    for (int lg = 1; lg <= 7; lg++) total += 2 * ((size_t)1 << (2 * lg));
    float* host = (float*)malloc(total * sizeof(float));
    if (!host) return AEAJ_ENOMEM;
    AEAJ_CUDA(cudaMalloc(&h->dct_all_dev, total * sizeof(float)));
    size_t o = 0;
    for (int lg = 1; lg <= 7; lg++) {
        const int s = 1 << lg;
        for (int k = 0; k < s; k++)
            for (int i = 0; i < s; i++) {
                double v = sqrt(2.0 / s) * cos(M_PI * (2 * i + 1) * k / (2.0 * s));
                if (k == 0) v *= sqrt(0.5);
                host[o + (size_t)k * s + i] = (float)v;
                host[o + (size_t)s * s + (size_t)i * s + k] = (float)v;
            }
        h->dct_dev[lg] = h->dct_all_dev + o;
        o += 2 * (size_t)s * s;
    }
    h->dct_dev[0] = nullptr; h->dct_dev[8] = nullptr;
    AEAJ_CUDA(cudaMemcpy(h->dct_all_dev, host, total * sizeof(float), cudaMemcpyHostToDevice));
    free(host);
    return 0;
}

int launch_dct_quant(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* class_lists, const int* class_counts,
                     const int64_t* off, const int64_t* caps, int lg_min, int lg_max, cudaStream_t st, int* launches,
                     void (*mark)(void*, const char*), void* mark_ctx) {
    return launch_all<false>(h, planes_dev, class_lists, class_counts, off, caps, lg_min, lg_max, st, launches, mark, mark_ctx);
}
int launch_dequant_idct(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* class_lists, const int* class_counts,
                        const int64_t* off, const int64_t* caps, int lg_min, int lg_max, cudaStream_t st, int* launches,
                        void (*mark)(void*, const char*), void* mark_ctx) {
    return launch_all<true>(h, planes_dev, class_lists, class_counts, off, caps, lg_min, lg_max, st, launches, mark, mark_ctx);
}
