// dct_tc.cu -- 128x128 DCT + quantise / dequantise + IDCT on the 5th-generation tensor cores (tcgen05, TMEM),
// error-compensated 3xTF32:  Out = A.X.A^T  (A = C forward, C^T inverse) with every operand split as v = hi + lo
// (hi = top 11 mantissa bits, exactly representable in TF32) and  P.Q ~= Ph.Qh + Ph.Ql + Pl.Qh  accumulated in
// FP32 in tensor memory.
//
// Default path for 128x128 leaves (aeaj_plan_set_tensor_dct(plan, 0) selects the FP32-FMA kernels of dct.cu): the
// parity bar for the forward DCT is the quantiser (tie class T-DCT) and for the inverse the <= 1 LSB / 3e-6 bound of
// the decoded samples; tests/test_gpu_parity.py::test_tensor_core_dct_parity compares both paths against the oracle.
//
// One CTA (256 threads) per 128x128 leaf, persistent over the size-128 work list:
//   GEMM1  W = A . X      A tile (smem, K-major, hi/lo)         B = X^T (smem, K-major, hi/lo; four 32-row K chunks)
//   split  W -> Wh, Wl    TMEM -> registers -> TMEM (tcgen05.ld / tcgen05.st)
//   GEMM2  Out = W . A^T  A operand = W (TMEM, hi/lo)           B = the same shared-memory tiles as GEMM1's A
//   epilogue: tcgen05.ld -> smem staging -> exact float32 quantiser -> int32 coefficients   (forward)
//                                        -> de-normalise, crop -> float32 layer samples    (inverse)
// Shared-memory operands use the canonical no-swizzle K-major layout: [K/4 chunks][rows][4 floats], i.e. 8x16-byte
// core matrices, stride-byte-offset 128 B (next 8 rows), leading-byte-offset rows*16 B (next K chunk).
#include "aeaj_internal.cuh"

namespace {

constexpr int TC_N = 128;
constexpr size_t TC_SMEM_BYTES = (size_t)(2 * TC_N * TC_N + 2 * TC_N * 64) * sizeof(float);
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
// round to the nearest TF32 value (10 explicit mantissa bits; ties away from zero).  The tensor core truncates the low 13
// bits of an FP32 operand, so both parts are rounded here: hi = rn(v), lo = rn(v - hi) leaves |v - hi - lo| <= 2^-24 |v|
// and |lo| <= 2^-12 |v|, i.e. the dropped lo.lo product and the representation error are both at FP32 rounding level.
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) { hi = tf32_rn(v); lo = tf32_rn(__fsub_rn(v, hi)); }

struct LeafGeo {
    float* base;            // layer + e.y * w + e.x
    int* cf;                // coefficient block of this leaf
    const int* qt;          // quantiser table (128 x 128)
    const int* zz;          // zigzag order, or nullptr for the row-major stream
    int w, bh, bw;
    float mid, sc;
    int col[4];             // the four columns (g + 32 q) this thread reads, reflect-padded
    bool fast;
};

// c_tiles: [Ch | Cl], each TC_N*TC_N floats in the canonical layout [K/4][128][4] (built on the host)
//
// Pipeline of one leaf (all 256 threads unless noted):
//   GEMM1 runs over K (= rows of X) in four chunks of 32 rows.  A chunk is 32 x 128 samples: thread (warp kc, lane g)
//   holds rows 4kc..4kc+3 of columns g, g+32, g+64, g+96 in registers (coalesced 128-byte loads, issued two chunks
//   ahead -- the first two chunks of the NEXT leaf are fetched while this leaf is in GEMM2 / the epilogue), splits them
//   into hi/lo and stores them into one of two K-major shared-memory buffers (16-byte stores, conflict free); thread 0
//   then issues the 12 MMAs (N = 128) of that chunk and commits to the buffer's mbarrier, which the stores of chunk c+2
//   wait for.  split / GEMM2 / epilogue as described in the file header.
template <bool INV>
__global__ void __launch_bounds__(256, 1) k_dct_tc128(const PlaneDesc* __restrict__ planes, const ClassEntry* __restrict__ list,
                                                      const int* __restrict__ count_ptr, const float* __restrict__ c_tiles,
                                                      const int* __restrict__ izz, int* __restrict__ err) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* sCh = reinterpret_cast<float*>(smem_raw);                     // 64 KB
    float* sCl = sCh + TC_N * TC_N;                                      // 64 KB
    float* sX = sCl + TC_N * TC_N;                                       // 2 buffers x (hi 16 KB + lo 16 KB); the epilogue's staging tile
    __shared__ __align__(8) unsigned long long mbar_storage[3];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2 * TC_N * TC_N / 4; i += 256) reinterpret_cast<float4*>(sCh)[i] = __ldg(reinterpret_cast<const float4*>(c_tiles) + i);
    const uint32_t bar0 = smem_u32(&mbar_storage[0]), bar1 = smem_u32(&mbar_storage[1]), bar2 = smem_u32(&mbar_storage[2]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar2) : "memory");
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
    const int row = (warp & 3) * 32 + lane;                               // the matrix row this thread reads from TMEM
    const int cbeg = (warp >> 2) * 2;                                     // ... and its two 32-column chunks
    const uint32_t aCh = smem_u32(sCh), aCl = smem_u32(sCl), aX = smem_u32(sX);
    const int count = *count_ptr;

    auto next_valid = [&](int li) {
        while (li < count) {
            const ClassEntry e = list[li];
            const PlaneDesc& P = planes[e.plane];
            if (e.y >= P.ry0 && e.y < P.ry1) break;
            li += gridDim.x;
        }
        return li;
    };
    auto geo_of = [&](int li) {
        const ClassEntry e = list[li];
        const PlaneDesc& P = planes[e.plane];
        LeafGeo G;
        G.w = P.w; G.bh = min(TC_N, P.h - e.y); G.bw = min(TC_N, P.w - e.x);
        G.base = P.layer_f32 + (size_t)e.y * P.w + e.x;
        G.cf = P.coef + (size_t)e.coef_off;
        G.qt = P.qtab[7];
        G.zz = P.zigzag ? P.zz[7] : nullptr;
        G.mid = P.mid; G.sc = P.scale;
        G.fast = (G.bh == TC_N && G.bw == TC_N);
#pragma unroll
        for (int q = 0; q < 4; q++) G.col[q] = G.fast ? lane + 32 * q : pad_reflect(lane + 32 * q, G.bw);
        return G;
    };
    uint32_t xr[2][4][4];                                                 // two chunks in flight (raw bits): [slot][row r][column q]
    // quantiser steps kept in registers for the whole run of leaves that share a table (luma <-> chroma changes only):
    // forward: the 64 positions this thread quantises in the epilogue; inverse: the 64 positions it loads
    // (packed with the position of the coefficient in the block's stream: row-major, or zigzag through izz: pos << 16 | q)
    float fq[16][4];
    uint32_t iq[4][4][4];
    auto load_chunk = [&](const LeafGeo& G, int c, int slot) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = 32 * c + 4 * warp + r;
            if (INV) {
#pragma unroll
                for (int q = 0; q < 4; q++) xr[slot][r][q] = (uint32_t)__ldg(G.cf + (iq[c][r][q] >> 16));
            } else {
                const float* rp = G.base + (size_t)(G.fast ? i : pad_reflect(i, G.bh)) * G.w;
#pragma unroll
                for (int q = 0; q < 4; q++) xr[slot][r][q] = __float_as_uint(__ldg(rp + G.col[q]));
            }
        }
    };
    auto store_chunk = [&](const LeafGeo& G, int c, int slot) {         // registers -> hi/lo K-major tiles of buffer `slot`
        float* xh = sX + slot * 8192;
        float* xl = xh + 4096;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            float hi[4], lo[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const float x = INV ? (float)((int)xr[slot][r][q] * (int)(iq[c][r][q] & 0xffffu))      // jpeg.py:524
                                    : __fmul_rn(__fsub_rn(__uint_as_float(xr[slot][r][q]), G.mid), G.sc);
                tf32_split(x, hi[r], lo[r]);
            }
            const int o = (warp * TC_N + lane + 32 * q) * 4;
            *reinterpret_cast<float4*>(xh + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(xl + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
    };

    const int* cur_qt = nullptr;
    const int* cur_zz = nullptr;
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0;
    bool alive = true;
    int li = next_valid(blockIdx.x);
    LeafGeo cur, nxt;
    auto load_tables = [&](const LeafGeo& G) {                          // on a change of plane type only
        cur_qt = G.qt; cur_zz = G.zz;
        if (INV) {
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int nat = (32 * c + 4 * warp + r) * TC_N + lane + 32 * q;
                        const int pos = G.zz ? __ldg(izz + nat) : nat;
                        iq[c][r][q] = ((uint32_t)pos << 16) | (uint32_t)__ldg(G.qt + nat);
                    }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int4 qv = __ldg(reinterpret_cast<const int4*>(G.qt) + tid + 256 * i);
                fq[i][0] = (float)qv.x; fq[i][1] = (float)qv.y; fq[i][2] = (float)qv.z; fq[i][3] = (float)qv.w;
            }
        }
    };
    if (li < count) { cur = geo_of(li); load_tables(cur); load_chunk(cur, 0, 0); load_chunk(cur, 1, 1); }
    int leaf_no = 0;
    while (li < count && alive) {
        const int nli = next_valid(li + gridDim.x);
        const bool has_next = nli < count;
        if (has_next) nxt = geo_of(nli);
        if (cur.qt != cur_qt || cur.zz != cur_zz) load_tables(cur);
        const bool dbg = (blockIdx.x == 0 && tid == 0 && leaf_no == 2);   // phase clocks of one steady-state leaf -> err[1..]
        long long t0 = clock64(); int ti = 1;
#define TC_STAMP() do { if (dbg) { long long t1 = clock64(); err[ti++] = (int)(t1 - t0); t0 = t1; } } while (0)
        // ---- GEMM1: W = C . X over four K chunks -----------------------------------------------------------
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int b = c & 1;
            if (c >= 2 && alive) {                                         // the MMAs of chunk c-2 have released buffer b
                if (b == 0) { alive = mbar_wait(bar0, ph0, err); ph0 ^= 1; } else { alive = mbar_wait(bar1, ph1, err); ph1 ^= 1; }
            }
            store_chunk(cur, c, b);
            if (c < 2) load_chunk(cur, c + 2, b);
            else if (has_next) load_chunk(nxt, c - 2, b);              // (stream positions: the layout is the same for all planes of a plan)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tid == 0 && alive) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t idesc = TC_IDESC_BASE | ((128u >> 3) << 17);
                const uint32_t xh = aX + b * 32768, xl = xh + 16384;
#pragma unroll
                for (int s = 0; s < 4; s++) {
                    const uint32_t ao = (c * 4 + s) * (TC_N * 32);
                    const uint64_t ah = make_desc(aCh + ao, TC_N * 16, 128), al = make_desc(aCl + ao, TC_N * 16, 128);
                    const uint64_t bh_ = make_desc(xh + s * (TC_N * 32), TC_N * 16, 128), bl = make_desc(xl + s * (TC_N * 32), TC_N * 16, 128);
                    mma_ss(D1, ah, bh_, idesc, (c | s) != 0);
                    mma_ss(D1, ah, bl, idesc, 1);
                    mma_ss(D1, al, bh_, idesc, 1);
                }
                mma_commit(b == 0 ? bar0 : bar1);
            }
            TC_STAMP();
        }
        if (alive) { alive = mbar_wait(bar0, ph0, err); ph0 ^= 1; }
        if (alive) { alive = mbar_wait(bar1, ph1, err); ph1 ^= 1; }
        TC_STAMP();
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
                float wh, wl;
                tf32_split(__uint_as_float(v[i]), wh, wl);
                h[i] = __float_as_uint(wh); l[i] = __float_as_uint(wl);
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
            mma_commit(bar2);
        }
        TC_STAMP();
        alive = mbar_wait(bar2, ph2, err);
        ph2 ^= 1;
        TC_STAMP();
        if (!alive) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- epilogue: TMEM -> shared staging (row-major, 16-byte groups XOR-swizzled by the row so that the
        //      row-per-thread writes are bank-conflict free) -> coalesced quantise + store by all threads
        float* stage = sX;                                                 // 128 x 128 floats: the X buffers are free now
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
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        TC_STAMP();
        if (INV) {
            const bool vec = (cur.w & 3) == 0;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int idx = tid + 256 * i, r = idx >> 5, j4 = idx & 31;
                const float4 z = *reinterpret_cast<const float4*>(stage + r * TC_N + (((j4 & ~7) | ((j4 & 7) ^ (r & 7))) * 4));
                const float v[4] = {__fadd_rn(__fdiv_rn(z.x, cur.sc), cur.mid), __fadd_rn(__fdiv_rn(z.y, cur.sc), cur.mid),
                                    __fadd_rn(__fdiv_rn(z.z, cur.sc), cur.mid), __fadd_rn(__fdiv_rn(z.w, cur.sc), cur.mid)};
                if (r < cur.bh) {
                    float* rp = cur.base + (size_t)r * cur.w + 4 * j4;
                    if (vec && 4 * j4 + 3 < cur.bw) *reinterpret_cast<float4*>(rp) = make_float4(v[0], v[1], v[2], v[3]);
                    else {
#pragma unroll
                        for (int k = 0; k < 4; k++) if (4 * j4 + k < cur.bw) rp[k] = v[k];
                    }
                }
            }
        } else {
            const bool zig = cur.zz != nullptr;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int idx = tid + 256 * i, r = idx >> 5, j4 = idx & 31;
                float* sp = stage + r * TC_N + (((j4 & ~7) | ((j4 & 7) ^ (r & 7))) * 4);
                const float4 z = *reinterpret_cast<const float4*>(sp);
                const int4 o = make_int4(quantize_f(z.x, fq[i][0]), quantize_f(z.y, fq[i][1]), quantize_f(z.z, fq[i][2]), quantize_f(z.w, fq[i][3]));
                if (zig) *reinterpret_cast<int4*>(sp) = o;                 // in place; gathered in stream order below
                else reinterpret_cast<int4*>(cur.cf)[idx] = o;
            }
            if (zig) {
                __syncthreads();
                const int* si = reinterpret_cast<const int*>(stage);
#pragma unroll 4
                for (int i = tid; i < TC_N * TC_N; i += 256) {
                    const int z = __ldg(cur.zz + i), r = z >> 7, j = z & 127;
                    cur.cf[i] = si[r * TC_N + ((((j >> 2) & ~7) | (((j >> 2) & 7) ^ (r & 7))) * 4) + (j & 3)];
                }
            }
        }
        TC_STAMP();
        __syncthreads();                                                   // the staging tile / TMEM are free for the next leaf
        li = nli; cur = nxt; leaf_no++;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

}  // namespace

// tiles [fwd: Ch | Cl][inv: Ch | Cl] in the canonical K-major layout: element (row r, k) at (k/4)*(128*4) + r*4 + (k%4);
// forward tile element (r, k) = C[r][k], inverse tile element (r, k) = C[k][r]
int aeaj_dct_tc_init(aeaj_handle* h) {
    const int N = TC_N;
    AEAJ_CUDA(cudaFuncSetAttribute(k_dct_tc128<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    AEAJ_CUDA(cudaFuncSetAttribute(k_dct_tc128<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    std::vector<float> host(4 * (size_t)N * N);
    for (int inv = 0; inv < 2; inv++)
        for (int r = 0; r < N; r++)
            for (int k = 0; k < N; k++) {
                const int u = inv ? k : r, x = inv ? r : k;                  // C[u][x] = a(u) cos(pi (2x+1) u / 2N)
                double v = sqrt(2.0 / N) * cos(M_PI * (2 * x + 1) * u / (2.0 * N));
                if (u == 0) v *= sqrt(0.5);
                const float c = (float)v;
                auto rn = [](float f) { uint32_t b; memcpy(&b, &f, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&f, &b, 4); return f; };
                const float hi = rn(c);
                const size_t o = (size_t)inv * 2 * N * N + (size_t)(k / 4) * (N * 4) + (size_t)r * 4 + (k % 4);
                host[o] = hi; host[(size_t)N * N + o] = rn(c - hi);
            }
    AEAJ_CUDA(cudaMalloc(&h->dct_tc_tiles_dev, host.size() * sizeof(float)));
    AEAJ_CUDA(cudaMemcpy(h->dct_tc_tiles_dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    // inverse zigzag permutation for 128x128 (row-major index -> stream position), from the table aeaj_dct_init built
    std::vector<int32_t> zz((size_t)N * N), izz((size_t)N * N);
    AEAJ_CUDA(cudaMemcpy(zz.data(), h->zz_dev[7], zz.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < N * N; i++) izz[zz[i]] = i;
    AEAJ_CUDA(cudaMalloc(&h->tc_izz_dev, izz.size() * sizeof(int32_t)));
    AEAJ_CUDA(cudaMemcpy(h->tc_izz_dev, izz.data(), izz.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    AEAJ_CUDA(cudaMalloc(&h->tc_err_dev, 32 * sizeof(int)));
    AEAJ_CUDA(cudaMemset(h->tc_err_dev, 0, 32 * sizeof(int)));
    return 0;
}

int launch_dct_tc128(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* list, const int* count, int64_t cap, int inverse, cudaStream_t st) {
    const size_t smem = TC_SMEM_BYTES;                  // opted in per device by aeaj_dct_tc_init
    int blocks = (int)std::min<int64_t>(std::max<int64_t>(cap, 1), (int64_t)h->sm_count);
    if (inverse) k_dct_tc128<true><<<blocks, 256, smem, st>>>(planes_dev, list, count, h->dct_tc_tiles_dev + 2 * TC_N * TC_N, h->tc_izz_dev, h->tc_err_dev);
    else k_dct_tc128<false><<<blocks, 256, smem, st>>>(planes_dev, list, count, h->dct_tc_tiles_dev, h->tc_izz_dev, h->tc_err_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
