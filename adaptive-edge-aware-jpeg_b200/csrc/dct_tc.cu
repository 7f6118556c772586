// dct_tc.cu -- 128x128 forward DCT + quantise on the 5th-generation tensor cores (tcgen05, TMEM),
// error-compensated 3xTF32:  Out = C.X.C^T  with every operand split as v = hi + lo (hi = top 11 mantissa
// bits, exactly representable in TF32) and  A.B ~= Ah.Bh + Ah.Bl + Al.Bh  accumulated in FP32 in tensor memory.
//
// OPT-IN (aeaj_plan_set_tensor_dct): the parity bar for the DCT is the quantiser (tie class T-DCT); the
// FP32-FMA kernels in dct.cu are the default, this path must show the same flip count before it replaces them
// (tests/test_gpu_parity.py::test_tensor_core_dct_parity reports it).
//
// One CTA (256 threads) per 128x128 leaf, persistent over the size-128 work list:
//   GEMM1  W = C . X      A = C   (smem, K-major, hi/lo)        B = X^T (smem, K-major, hi/lo; two 64-column halves)
//   split  W -> Wh, Wl    TMEM -> registers -> TMEM (tcgen05.ld / tcgen05.st)
//   GEMM2  Out = W . C^T  A = W   (TMEM, hi/lo)                 B = C   (the same shared-memory tiles as GEMM1's A)
//   epilogue: tcgen05.ld -> exact float32 quantiser -> int32 coefficients
// Shared-memory operands use the canonical no-swizzle K-major layout: [K/4 chunks][rows][4 floats], i.e. 8x16-byte
// core matrices, stride-byte-offset 128 B (next 8 rows), leading-byte-offset rows*16 B (next K chunk).
#include "aeaj_internal.cuh"

namespace {

constexpr int TC_N = 128;
constexpr uint32_t TC_IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 4) << 24);   // F32 accum, TF32 x TF32, K-major, M=128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
// bounded wait: a mis-programmed MMA must not hang the GPU -- give up after ~1 s and raise the error flag
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity, int* err) {
    for (long long it = 0; it < 400000000ll; it++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(mbar), "r"(parity) : "memory");
        if (ok) return true;
    }
    if (err) atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
                 "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                 "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, "
                 "%25, %26, %27, %28, %29, %30, %31, %32};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                    "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
                    "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
                    "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// c_tiles: [Ch | Cl], each TC_N*TC_N floats in the canonical layout [K/4][128][4] (built on the host)
__global__ void __launch_bounds__(256, 1) k_dct_tc128(const PlaneDesc* __restrict__ planes, const ClassEntry* __restrict__ list,
                                                      const int* __restrict__ count_ptr, const float* __restrict__ c_tiles, int* __restrict__ err, int dbg_mode) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* sCh = reinterpret_cast<float*>(smem_raw);                     // 64 KB
    float* sCl = sCh + TC_N * TC_N;                                      // 64 KB
    float* sXh = sCl + TC_N * TC_N;                                      // 32 KB  (64 columns of X, hi)
    float* sXl = sXh + TC_N * 64;                                        // 32 KB
    __shared__ __align__(8) unsigned long long mbar_storage;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 2 * TC_N * TC_N / 4; i += 256) reinterpret_cast<float4*>(sCh)[i] = __ldg(reinterpret_cast<const float4*>(c_tiles) + i);
    const uint32_t mbar = smem_u32(&mbar_storage);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&tmem_base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t D1 = tbase, WH = tbase + 128, WL = tbase + 256, D2 = tbase + 384;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;          // this warp's 32 TMEM lanes (warps w and w+4 share a quadrant)
    const int row = (warp & 3) * 32 + (tid & 31);                         // the matrix row this thread reads from TMEM
    const int cbeg = (warp >> 2) * 2;                                     // ... and its two 32-column chunks
    const uint32_t aCh = smem_u32(sCh), aCl = smem_u32(sCl), aXh = smem_u32(sXh), aXl = smem_u32(sXl);
    uint32_t phase = 0;
    bool alive = true;
    const int count = *count_ptr;
    // the 64 quantiser steps this thread applies (the same positions for every leaf) stay in registers; they are
    // reloaded only when the work list moves to a plane with another table (luma <-> chroma)
    float fq[16][4];
    const int* cur_qt = nullptr;
    for (int li = blockIdx.x; li < count && alive; li += gridDim.x) {
        const ClassEntry e = list[li];
        const PlaneDesc& P = planes[e.plane];
        if (e.y < P.ry0 || e.y >= P.ry1) continue;
        if (P.qtab[7] != cur_qt) {
            cur_qt = P.qtab[7];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int4 qv = __ldg(reinterpret_cast<const int4*>(cur_qt) + tid + 256 * i);
                fq[i][0] = (float)qv.x; fq[i][1] = (float)qv.y; fq[i][2] = (float)qv.z; fq[i][3] = (float)qv.w;
            }
        }
        const int bh = min(TC_N, P.h - e.y), bw = min(TC_N, P.w - e.x);
        const float mid = P.mid, sc = P.scale;
        // ---- GEMM1: W = C . X, two halves of 64 columns ------------------------------------------------
        const bool fast = (bh == TC_N && bw == TC_N && (P.w & 3) == 0);
        const bool dbg = (blockIdx.x == 0 && tid == 0 && li == blockIdx.x + 2 * gridDim.x);
        long long t0 = clock64(); int ti = 1;
#define TC_STAMP() do { if (dbg) { long long t1 = clock64(); err[ti++] = (int)(t1 - t0); t0 = t1; } } while (0)
        // 4x4 patches: rows 4c..4c+3 (one K chunk), columns 4g..4g+3 of a 64-column half; 128-bit loads (issued one
        // half ahead, so that they are in flight while the tensor core works), register transpose, one 16-byte store
        // per column into the K-major tile
        float v[2][4][4];
        auto load_half = [&](int half) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int pidx = tid + 256 * u, c = pidx >> 4, g = pidx & 15;
                if (fast) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(P.layer_f32 + (size_t)(e.y + 4 * c + r) * P.w + e.x + half * 64 + 4 * g));
                        v[u][r][0] = t.x; v[u][r][1] = t.y; v[u][r][2] = t.z; v[u][r][3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            v[u][r][q] = __ldg(P.layer_f32 + (size_t)(e.y + pad_reflect(4 * c + r, bh)) * P.w + e.x + pad_reflect(half * 64 + 4 * g + q, bw));
                }
            }
        };
        load_half(0);
#pragma unroll
        for (int half = 0; half < 2; half++) {
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int pidx = tid + 256 * u, c = pidx >> 4, g = pidx & 15;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    float hi[4], lo[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float x = __fmul_rn(__fsub_rn(v[u][r][q], mid), sc);
                        hi[r] = tf32_hi(x); lo[r] = __fsub_rn(x, hi[r]);
                    }
                    const int jj = 4 * g + q;
                    *reinterpret_cast<float4*>(sXh + (c * 64 + jj) * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(sXl + (c * 64 + jj) * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            TC_STAMP();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t idesc = TC_IDESC_BASE | ((64u >> 3) << 17);
                for (int s = 0; s < TC_N / 8; s++) {
                    const uint64_t ah = make_desc(aCh + s * (TC_N * 32), TC_N * 16, 128), al = make_desc(aCl + s * (TC_N * 32), TC_N * 16, 128);
                    const uint64_t bh_ = make_desc(aXh + s * (64 * 32), 64 * 16, 128), bl = make_desc(aXl + s * (64 * 32), 64 * 16, 128);
                    mma_ss(D1 + half * 64, ah, bh_, idesc, s > 0);
                    mma_ss(D1 + half * 64, ah, bl, idesc, 1);
                    mma_ss(D1 + half * 64, al, bh_, idesc, 1);
                }
                mma_commit(mbar);
            }
            TC_STAMP();
            if (half == 0) load_half(1);
            alive = mbar_wait(mbar, phase, err);
            phase ^= 1;
            TC_STAMP();
            if (!alive) break;
        }
        if (!alive) break;
        // ---- split W = Wh + Wl inside tensor memory ------------------------------------------------------
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c = cbeg; c < cbeg + 2; c++) {
            uint32_t v[32], h[32], l[32];
            tmem_ld32(D1 + lane_sel + c * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const float w = __uint_as_float(v[i]), wh = tf32_hi(w);
                h[i] = __float_as_uint(wh); l[i] = __float_as_uint(__fsub_rn(w, wh));
            }
            tmem_st32(WH + lane_sel + c * 32, h);
            tmem_st32(WL + lane_sel + c * 32, l);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        TC_STAMP();
        // ---- GEMM2: Out = W . C^T -------------------------------------------------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t idesc = TC_IDESC_BASE | ((128u >> 3) << 17);
            for (int s = 0; s < TC_N / 8; s++) {
                const uint64_t bh_ = make_desc(aCh + s * (TC_N * 32), TC_N * 16, 128), bl = make_desc(aCl + s * (TC_N * 32), TC_N * 16, 128);
                mma_ts(D2, WH + s * 8, bh_, idesc, s > 0);
                mma_ts(D2, WH + s * 8, bl, idesc, 1);
                mma_ts(D2, WL + s * 8, bh_, idesc, 1);
            }
            mma_commit(mbar);
        }
        TC_STAMP();
        alive = mbar_wait(mbar, phase, err);
        phase ^= 1;
        TC_STAMP();
        if (!alive) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: TMEM -> shared staging (row-major, 16-byte groups XOR-swizzled by the row so that the
        //      row-per-thread writes are bank-conflict free) -> coalesced quantise + store by all threads
        float* stage = sXh;                                                // 128 x 128 floats: the X tiles are free now
#pragma unroll 1
        for (int c = cbeg; c < cbeg + 2; c++) {
            uint32_t v[32];
            tmem_ld32(D2 + lane_sel + c * 32, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; i++)
                *reinterpret_cast<float4*>(stage + row * TC_N + c * 32 + ((i ^ (row & 7)) * 4)) =
                    make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
        TC_STAMP();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        TC_STAMP();
        {
            int* cf = P.coef + (size_t)e.coef_off;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int idx = tid + 256 * i, r = idx >> 5, j4 = idx & 31;
                const float4 z = *reinterpret_cast<const float4*>(stage + r * TC_N + (((j4 & ~7) | ((j4 & 7) ^ (r & 7))) * 4));
                const int4 o = (dbg_mode & 4) ? make_int4(__float_as_int(z.x) + __float_as_int(fq[i][0]), __float_as_int(z.y) + __float_as_int(fq[i][1]), __float_as_int(z.z), __float_as_int(z.w) + __float_as_int(fq[i][3]))
                                              : make_int4(quantize_f(z.x, fq[i][0]), quantize_f(z.y, fq[i][1]), quantize_f(z.z, fq[i][2]), quantize_f(z.w, fq[i][3]));
                if (!(dbg_mode & 2) || o.x == 0x7fffffff) reinterpret_cast<int4*>(cf)[idx] = o;
            }
        }
        TC_STAMP();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                   // TMEM / smem tiles are free for the next leaf
        TC_STAMP();
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

}  // namespace

// [Ch | Cl] in the canonical K-major layout: element (row r, k) at (k/4)*(128*4) + r*4 + (k%4)
int aeaj_dct_tc_init(aeaj_handle* h) {
    const int N = TC_N;
    std::vector<float> host(2 * (size_t)N * N);
    for (int r = 0; r < N; r++)
        for (int k = 0; k < N; k++) {
            double v = sqrt(2.0 / N) * cos(M_PI * (2 * k + 1) * r / (2.0 * N));
            if (r == 0) v *= sqrt(0.5);
            const float c = (float)v;
            uint32_t bits; memcpy(&bits, &c, 4); bits &= 0xffffe000u;
            float hi; memcpy(&hi, &bits, 4);
            const size_t o = (size_t)(k / 4) * (N * 4) + (size_t)r * 4 + (k % 4);
            host[o] = hi; host[(size_t)N * N + o] = c - hi;
        }
    AEAJ_CUDA(cudaMalloc(&h->dct_tc_tiles_dev, host.size() * sizeof(float)));
    AEAJ_CUDA(cudaMemcpy(h->dct_tc_tiles_dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    AEAJ_CUDA(cudaMalloc(&h->tc_err_dev, 32 * sizeof(int)));
    AEAJ_CUDA(cudaMemset(h->tc_err_dev, 0, 32 * sizeof(int)));
    return 0;
}

int launch_dct_tc128(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* list, const int* count, int64_t cap, cudaStream_t st) {
    const size_t smem = (size_t)(2 * TC_N * TC_N + 2 * TC_N * 64) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        AEAJ_CUDA(cudaFuncSetAttribute(k_dct_tc128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cap, 1), (int64_t)h->sm_count);
    static int dbg_mode = -1;
    if (dbg_mode < 0) { const char* e = getenv("AEAJ_TC_DBG"); dbg_mode = e ? atoi(e) : 0; }
    k_dct_tc128<<<blocks, 256, smem, st>>>(planes_dev, list, count, h->dct_tc_tiles_dev, h->tc_err_dev, dbg_mode);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
