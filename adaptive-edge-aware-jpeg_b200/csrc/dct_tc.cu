// dct_tc.cu -- DCT + quantise / dequantise + IDCT for the size classes 16 .. 128 on the 5th-generation tensor cores
// (tcgen05, TMEM), error-compensated 3xTF32.
//
// One kernel shape for every class: a CTA transforms a 128 x 128 SUPER-TILE that holds (128/S)^2 leaves of size S
// (S = 128: one leaf) as   Out = A . X . A^T   with A = blockdiag(C_S, .., C_S) forward and blockdiag(C_S^T, ..) inverse:
// the (p, q) sub-block of Out is C_S . X_pq . C_S^T, i.e. the 2-D DCT of leaf (p, q).  Small leaves therefore run at
// the same per-sample cost as 128 x 128 ones, on the tensor pipe instead of through shared-memory-bound FP32 loops.
// Every operand is split as v = hi + lo (both rounded to nearest TF32) and  P.Q ~= Ph.Qh + Ph.Ql + Pl.Qh  accumulates
// in FP32 in tensor memory: the error is at FP32 rounding level, the parity bar is the exact quantiser (tie class T-DCT)
// forward and the <= 1 LSB / 3e-6 bound of the decoded samples inverse (tests/test_gpu_parity.py).
//
// One persistent, warp-specialised CTA per SM over the super-tiles of a class (roles and mbarriers: see k_dct_tc):
//   GEMM1  W = A . X      A tiles (smem, K-major, hi/lo; 128 KB, loaded once)   B = X^T (smem, K-major, hi/lo), streamed
//                         in four K chunks of 32 super-rows through two buffers (full / empty mbarriers, tcgen05.commit)
//   split  W -> Wh, Wl    TMEM -> registers -> TMEM (tcgen05.ld / tcgen05.st), in place: Wh over D1, Wl next to it
//   GEMM2  Out = W . A^T  A operand = W (TMEM, hi/lo)   B = the same shared-memory tiles as GEMM1's A; A is block
//                         diagonal, so every K step only needs the N = S columns of its own block
//   epilogue: tcgen05.ld -> registers -> exact float32 quantiser -> 256-bit stores of int32 coefficients     (forward)
//                                     -> de-normalise, crop -> 256-bit stores of float32 layer samples       (inverse)
// No shared-memory staging of the result, and two Out buffers in tensor memory: the epilogue of tile i reads one while GEMM2 of
// tile i+1 fills the other.
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
// the same descriptor `units` 16-byte units further on: only the 14-bit address field in the low word changes (no carry out of
// it: shared memory is < 256 KB), so stepping through a tile costs one 32-bit add per operand instead of rebuilding 64 bits
__device__ __forceinline__ uint64_t desc_at(uint64_t base, uint32_t units) {
    return (base & 0xffffffff00000000ull) | (uint64_t)((uint32_t)base + units);
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
// 256-bit global loads / stores (sm_100): one full 32-byte sector per instruction and thread
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t* v) {
    // volatile: stays where it is written relative to the other volatile statements, i.e. ABOVE the tensor-memory load that
    // follows it in the epilogue (a plain asm is sunk to its first use and its L2 latency exposed once per group)
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// round to the nearest TF32 value (10 explicit mantissa bits; ties away from zero).  The tensor core truncates the low 13
// bits of an FP32 operand, so both parts are rounded here: hi = rn(v), lo = rn(v - hi) leaves |v - hi - lo| <= 2^-24 |v|
// and |lo| <= 2^-12 |v|, i.e. the dropped lo.lo product and the representation error are both at FP32 rounding level.
__device__ __forceinline__ float tf32_rn(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void tf32_split(float v, float& hi, float& lo) { hi = tf32_rn(v); lo = tf32_rn(__fsub_rn(v, hi)); }

// one leaf of a super-tile (shared memory)
struct TcLeaf {
    float* base;            // layer + y * w + x (forward: source, inverse: destination); nullptr: empty slot / leaf of another band
    int* cf;                // coefficient block of this leaf
    const float* qf;        // quantiser steps as float in the epilogue layout [S/8 column groups][S rows][8] (forward)
    const int* qi;          // quantiser steps, S x S (inverse)
    int w, bh, bw, zig;
    float mid, sc;
};

__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
                    "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}

#ifdef AEAJ_TC_STAMPS
#define TC_STAMP(slot) do { if (blockIdx.x == 0 && it == 3 && lane == 0) err[(slot)] = (int)(clock64() & 0x7fffffff); } while (0)
#else
#define TC_STAMP(slot) do { } while (0)
#endif
constexpr int TC_CONSUMERS = 256, TC_PRODUCERS = 256, TC_THREADS = TC_CONSUMERS + TC_PRODUCERS + 128;
constexpr int TC_CW = TC_CONSUMERS / 32, TC_PW = TC_PRODUCERS / 32;       // consumer / producer warps; the MMA warp is warp TC_CW + TC_PW
// Registers: 20 warps = 5 per SM sub-partition, whose register file holds 512 per thread slot.  The kernel is compiled for 96
// (launch bound); the role branches then trade with setmaxnreg: the MMA warpgroup (one issuing thread, three idle warps)
// shrinks to 40 and the consumer / producer warpgroups grow to 112 / 104.  The trade happens inside the CTA's own allocation
// (20 x 32 x 96 registers): 256 x 112 + 256 x 104 + 128 x 40 = 60416 <= 61440 -- asking for more than the pool holds never returns.
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(N)); }
constexpr int TC_TABS = 8;                                                // leaf tables in flight (ring): the producers run up to 4 tiles ahead of the epilogue

// Warp-specialised, persistent: one CTA per SM loops over the super-tiles blockIdx.x, blockIdx.x + gridDim.x, ..
//   warps 0-7    CONSUMERS  (two per TMEM lane quadrant: a thread owns one super-tile row and half of its columns) split W in
//                           place inside tensor memory, then the epilogue straight out of tensor memory
//   warps 8-15   PRODUCERS  global -> registers (one chunk ahead) -> normalise / dequantise -> hi/lo split -> K-major smem tiles
//   warp 16      MMA        one thread issues G1(0), then per tile G2(i), G1(i+1) -- GEMM1 chunk by chunk as the X buffers fill
//   warps 17-19  idle       (setmaxnreg works on whole warpgroups; they give their registers away and wait for the end)
// mbarriers: full[b] (one arrival per producer warp) / empty[b] (tcgen05.commit) per X buffer; d1_full (commit: W complete),
// w_ready (one arrival per consumer warp: Wh, Wl written), d2_full[k] (commit: Out buffer k complete), d2_free[k] (Out buffer k read).
// Two Out buffers: the consumers split W(i+1) BEFORE the epilogue of tile i, so that G2(i+1) and G1(i+2) run on the tensor pipe
// during that epilogue.  Measured history (profiles/r2_tc_variants.md): 13 warps with 4 consumer warps 0.215 / 0.203 ms (128 class,
// forward / inverse, 8 4K frames); 17 warps capped every thread at 96 registers (5 warps on one SM sub-partition) and spilled;
// this layout with the setmaxnreg trade 0.20 / 0.166 ms.  What bounds it now is the tensor side itself: an MMA of K = 8 reads
// 8 KB of operands from shared memory and issues at ~100 cycles, 96 of them per tile, and the two X buffers (all the shared
// memory that is left next to the 128 KB of C tiles) let the producers run only half a tile ahead of GEMM1.
// Every wait is bounded: a protocol error raises the error flag instead of hanging the GPU.
//
// a_tiles: [Ah | Al], each 128 x 128 floats in the canonical layout [K/4][128][4] (built on the host): blockdiag(C_S) or its transpose
// izz: inverse zigzag permutation of an S x S block (row-major index -> stream position)
template <int S, bool INV>
__global__ void __launch_bounds__(TC_THREADS, 1) k_dct_tc(const PlaneDesc* __restrict__ planes, const ClassEntry* __restrict__ list,
                                                          const int* __restrict__ count_ptr, const float* __restrict__ a_tiles,
                                                          const int* __restrict__ izz, int* __restrict__ err) {
    constexpr int NB = TC_N / S;                                          // leaves per super-tile side
    constexpr int NL = NB * NB;                                           // leaves per super-tile
    constexpr int LG = (S == 16) ? 4 : (S == 32) ? 5 : (S == 64) ? 6 : 7;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* sCh = reinterpret_cast<float*>(smem_raw);                     // 64 KB
    float* sCl = sCh + TC_N * TC_N;                                      // 64 KB
    float* sX = sCl + TC_N * TC_N;                                       // 2 buffers x (hi 16 KB + lo 16 KB)
    __shared__ __align__(8) unsigned long long mbar_storage[10];
    __shared__ uint32_t tmem_base_s;
    __shared__ TcLeaf sLeaf[TC_TABS][NL];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 2 * TC_N * TC_N / 4; i += TC_THREADS) reinterpret_cast<float4*>(sCh)[i] = __ldg(reinterpret_cast<const float4*>(a_tiles) + i);
    const uint32_t bar_full[2] = {smem_u32(&mbar_storage[0]), smem_u32(&mbar_storage[1])};
    const uint32_t bar_empty[2] = {smem_u32(&mbar_storage[2]), smem_u32(&mbar_storage[3])};
    const uint32_t bar_d1full = smem_u32(&mbar_storage[4]), bar_wready = smem_u32(&mbar_storage[5]);
    const uint32_t bar_d2full[2] = {smem_u32(&mbar_storage[6]), smem_u32(&mbar_storage[7])};      // per Out buffer
    const uint32_t bar_d2free[2] = {smem_u32(&mbar_storage[8]), smem_u32(&mbar_storage[9])};
    if (tid == 0) {
        auto init = [](uint32_t b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(b), "r"(n) : "memory"); };
        // producer / consumer warps arrive once per warp (lane 0, after __syncwarp): 8 and 4 arrivals instead of 256 and 128
        init(bar_full[0], TC_PRODUCERS / 32); init(bar_full[1], TC_PRODUCERS / 32); init(bar_empty[0], 1); init(bar_empty[1], 1);
        init(bar_d1full, 1); init(bar_wready, TC_CONSUMERS / 32);
        init(bar_d2full[0], 1); init(bar_d2full[1], 1); init(bar_d2free[0], TC_CONSUMERS / 32); init(bar_d2free[1], TC_CONSUMERS / 32);
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
    // tensor memory, 512 columns: W = A.X is split into hi + lo IN PLACE (hi over D1, lo next to it), which leaves room for two
    // Out buffers -- GEMM2 of tile i+1 runs while the consumers still read tile i
    const uint32_t D1 = tbase, WH = tbase, WL = tbase + 128, D2A = tbase + 256;
    const uint32_t aCh = smem_u32(sCh), aCl = smem_u32(sCl), aX = smem_u32(sX);
    const int count = *count_ptr;
    const int ntiles = (count + NL - 1) / NL;
    const int my_tiles = (blockIdx.x < ntiles) ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp >= TC_CW && warp < TC_CW + TC_PW) {
        // =============================================== PRODUCERS ===============================================
        reg_inc<104>();
        const int pt = tid - TC_CONSUMERS, pw = pt >> 5;                  // producer thread / warp (0 .. 7)
        auto make_leaf = [&](int tile) {                                  // list entry tile * NL + pt -> leaf descriptor (registers)
            TcLeaf L;
            L.base = nullptr; L.cf = nullptr; L.qf = nullptr; L.qi = nullptr; L.w = 0; L.bh = 0; L.bw = 0; L.zig = 0; L.mid = 0.0f; L.sc = 1.0f;
            const int li = tile * NL + pt;
            if (pt < NL && tile < ntiles && li < count) {
                const ClassEntry e = list[li];
                const PlaneDesc& P = planes[e.plane];
                if (e.y >= P.ry0 && e.y < P.ry1) {                        // halo-split: only the leaves of this call's band
                    L.base = P.layer_f32 + (size_t)e.y * P.w + e.x;
                    L.cf = P.coef + (size_t)e.coef_off;
                    L.qf = P.qtabf[LG]; L.qi = P.qtab[LG];
                    L.w = P.w; L.bh = min(S, P.h - e.y); L.bw = min(S, P.w - e.x);
                    L.zig = P.zigzag; L.mid = P.mid; L.sc = P.scale;
                }
            }
            return L;
        };
        auto producers_sync = [] { asm volatile("bar.sync 1, 256;" ::: "memory"); };
        // registers of two chunks in flight: thread (warp kc, lane g) holds super-rows 32c + 4kc .. +3 of columns g, g+32, g+64, g+96
        uint32_t xr[2][4][4];
        int xq[INV ? 2 : 1][4][4];                                        // inverse: the quantiser steps of the same positions
        auto load_chunk = [&](const TcLeaf* T, int c, int slot) {
            const int r0 = 32 * c + 4 * pw;                               // super-row of r = 0; the four rows stay inside one leaf row block
            const int p = r0 / S, li0 = r0 % S;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = lane + 32 * q;
                const TcLeaf& L = T[p * NB + j / S];
                const int lj = j % S;
                if (L.base == nullptr) {
#pragma unroll
                    for (int r = 0; r < 4; r++) { xr[slot][r][q] = 0u; if (INV) xq[INV ? slot : 0][r][q] = 0; }
                    continue;
                }
                if (INV) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int nat = (li0 + r) * S + lj;
                        xr[slot][r][q] = (uint32_t)__ldg(L.cf + (L.zig ? __ldg(izz + nat) : nat));
                        xq[INV ? slot : 0][r][q] = __ldg(L.qi + nat);
                    }
                } else if (L.bh == S && L.bw == S) {
                    const float* src = L.base + (size_t)li0 * L.w + lj;
#pragma unroll
                    for (int r = 0; r < 4; r++) xr[slot][r][q] = __float_as_uint(__ldg(src + (size_t)r * L.w));
                } else {
                    // partial leaf (np.pad 'reflect', jpeg.py:402): one modulo for the column and one for the first row, the other
                    // three rows follow the triangle wave
                    const float* src = L.base + pad_reflect(lj, L.bw);
                    int rr = pad_reflect(li0, L.bh);
                    int dir = (L.bh > 1 && (li0 % (2 * (L.bh - 1))) >= L.bh) ? -1 : 1;
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        xr[slot][r][q] = __float_as_uint(__ldg(src + (size_t)rr * L.w));
                        if (L.bh > 1) {
                            int nx = rr + dir;
                            if (nx >= L.bh) { dir = -1; nx = L.bh - 2; } else if (nx < 0) { dir = 1; nx = 1; }
                            rr = nx;
                        }
                    }
                }
            }
        };
        auto store_chunk = [&](const TcLeaf* T, int c, int slot) {       // registers -> hi/lo K-major tiles of buffer `slot`
            float* xh = sX + slot * 8192;
            float* xl = xh + 4096;
            const int p = (32 * c + 4 * pw) / S;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const TcLeaf& L = T[p * NB + (lane + 32 * q) / S];
                const bool live = (L.base != nullptr);
                const float mid = L.mid, sc = L.sc;
                float hi[4], lo[4];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    float x = INV ? (float)((int)xr[slot][r][q] * xq[INV ? slot : 0][r][q])          // jpeg.py:524
                                  : __fmul_rn(__fsub_rn(__uint_as_float(xr[slot][r][q]), mid), sc);   // jpeg.py:387-390
                    x = live ? x : 0.0f;
                    tf32_split(x, hi[r], lo[r]);
                }
                const int o = (pw * TC_N + lane + 32 * q) * 4;
                *reinterpret_cast<float4*>(xh + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(xl + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
        };
        bool alive = true;
        if (my_tiles > 0) {
            if (pt < NL) sLeaf[0][pt] = make_leaf(blockIdx.x);
            producers_sync();
            load_chunk(sLeaf[0], 0, 0);
            load_chunk(sLeaf[0], 1, 1);
        }
        TcLeaf Lp = make_leaf(blockIdx.x + gridDim.x);                   // the next tile's entry, prefetched into registers
        uint32_t uses = 0;                                                // chunk stores so far: buffer b has been used (uses + 1 - b) / 2 times
        for (int it = 0; it < my_tiles && alive; it++) {
            const TcLeaf* Tc = sLeaf[it & (TC_TABS - 1)];
            const TcLeaf* Tn = sLeaf[(it + 1) & (TC_TABS - 1)];
            const bool has_next = it + 1 < my_tiles;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int b = c & 1;
                if (warp == TC_CW && c == 1) TC_STAMP(25);
                if (uses >= 2) alive = alive && mbar_wait(bar_empty[b], ((uses >> 1) - 1) & 1, err);   // the MMAs that read this buffer last are done
                if (warp == TC_CW) TC_STAMP(2 + 2 * c);
                if (c == 0) {
                    // the table ring slot of tile it + 1 belonged to tile it - 3, whose epilogue is over (its consumers went on to
                    // split tile it - 2 before the MMAs just waited for could be issued)
                    if (pt < NL) sLeaf[(it + 1) & (TC_TABS - 1)][pt] = Lp;
                    producers_sync();
                    Lp = make_leaf(blockIdx.x + (it + 2) * gridDim.x);
                }
                store_chunk(Tc, c, b);
                if (warp == TC_CW && c == 1) TC_STAMP(26);
                // publish the chunk BEFORE issuing the next global loads: the proxy fence waits for every memory operation of the
                // thread that is still in flight, so behind the loads it exposed their whole latency (~1.4 k cycles per chunk, half of
                // the producers' time, measured with the cycle stamps)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full[b]);
                if (warp == TC_CW) TC_STAMP(3 + 2 * c);
                if (c < 2) load_chunk(Tc, c + 2, b);
                else if (has_next) load_chunk(Tn, c - 2, b);
                if (warp == TC_CW && c == 1) TC_STAMP(27);
                uses++;
            }
        }
    } else if (warp >= TC_CW + TC_PW) {
        // =============================================== MMA ISSUER ===============================================
        reg_dec<40>();
        if (warp == TC_CW + TC_PW && lane == 0) {
            bool alive = true;
            uint32_t nfull = 0;                                           // chunks consumed so far
            const uint32_t idesc1 = TC_IDESC_BASE | ((128u >> 3) << 17);
            const uint32_t idesc2 = TC_IDESC_BASE | (((uint32_t)S >> 3) << 17);
            const uint64_t dCh = make_desc(aCh, TC_N * 16, 128), dCl = make_desc(aCl, TC_N * 16, 128), dX = make_desc(aX, TC_N * 16, 128);
            // GEMM1 of one tile: chunk by chunk as the producers fill the X buffers
            auto gemm1 = [&](int it) {
#pragma unroll 1
                for (int c = 0; c < 4 && alive; c++) {
                    const int b = c & 1;
                    alive = mbar_wait(bar_full[b], (nfull >> 1) & 1, err);
                    TC_STAMP(10 + c);
                    nfull++;
                    if (!alive) break;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // K step s of chunk c: A tile 256 units (TC_N * 32 bytes) per step, X buffer b = 2048 units, lo half 1024 units in
                    const uint32_t au = (uint32_t)c * 1024u, xu = (uint32_t)b * 2048u;
#pragma unroll
                    for (int s = 0; s < 4; s++) {
                        const uint64_t ah = desc_at(dCh, au + s * 256u), al = desc_at(dCl, au + s * 256u);
                        const uint64_t bh_ = desc_at(dX, xu + s * 256u), bl = desc_at(dX, xu + 1024u + s * 256u);
                        mma_ss(D1, ah, bh_, idesc1, (c | s) != 0);
                        mma_ss(D1, ah, bl, idesc1, 1);
                        mma_ss(D1, al, bh_, idesc1, 1);
                    }
                    mma_commit(bar_empty[b]);
                }
                if (alive) mma_commit(bar_d1full);                         // W = A . X complete (and every MMA issued before it)
                TC_STAMP(14);
            };
            // Order of issue = order of execution: G1(0), then per tile G2(i), G1(i+1).  G1(i+1) overwrites D1 = Wh(i), which G2(i)
            // -- issued just before it -- still reads: the tensor pipe executes a thread's MMAs in order.  The consumers split
            // W(i+1) only after d1_full(i+1), a commit that also covers G2(i).
            if (my_tiles > 0) gemm1(0);
            for (int it = 0; it < my_tiles && alive; it++) {
                alive = mbar_wait(bar_wready, it & 1, err);                // Wh / Wl of this tile written
                TC_STAMP(15);
                if (alive && it >= 2) alive = mbar_wait(bar_d2free[it & 1], ((it >> 1) - 1) & 1, err);   // Out buffer read by tile it - 2
                if (!alive) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // GEMM2: Out = W . A^T.  A is block diagonal: K step s (columns 8s .. 8s+7 of W) only feeds the S output columns of its block
                const uint32_t D2 = D2A + (it & 1) * 128;
#pragma unroll 4
                for (int s = 0; s < TC_N / 8; s++) {
                    const int n0 = (8 * s / S) * S;                        // first output column of the block
                    const uint32_t bu = (uint32_t)s * 256u + (uint32_t)n0; // K chunk pair s, tile rows n0 .. (16-byte units)
                    const uint64_t bh_ = desc_at(dCh, bu), bl = desc_at(dCl, bu);
                    const uint32_t first = (8 * s % S) == 0 ? 0u : 1u;     // first K step of a block overwrites its columns
                    mma_ts(D2 + n0, WH + s * 8, bh_, idesc2, first);
                    mma_ts(D2 + n0, WH + s * 8, bl, idesc2, 1);
                    mma_ts(D2 + n0, WL + s * 8, bh_, idesc2, 1);
                }
                mma_commit(bar_d2full[it & 1]);
                TC_STAMP(16);
                if (it + 1 < my_tiles) gemm1(it + 1);
            }
        }
    } else {
        // =============================================== CONSUMERS ===============================================
        reg_inc<112>();
        const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;      // this warp's 32 TMEM lanes (hardware: quadrant warp % 4)
        const int row = (warp & 3) * 32 + lane;                           // the super-tile row this thread owns in tensor memory
        const int cbeg = (warp >> 2) * (TC_N / 2);                        // ... and the half of its columns: two warps share a row
        const int p = row / S, li = row % S;
        bool alive = true;
        // W = Wh + Wl of tile `t`, in place inside tensor memory (hi over D1, lo in WL); then tell the MMA thread
        auto split_w = [&](int it) {
            alive = mbar_wait(bar_d1full, it & 1, err);                    // GEMM1(t) complete -- and GEMM2(t - 1), the last reader of Wh / Wl
            if (warp == 0) TC_STAMP(21);
            if (!alive) return;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                uint32_t v[16], h[16], l[16];
                tmem_ld16(D1 + lane_sel + cbeg + c * 16, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    float wh, wl;
                    tf32_split(__uint_as_float(v[i]), wh, wl);
                    h[i] = __float_as_uint(wh); l[i] = __float_as_uint(wl);
                }
                tmem_st16(WH + lane_sel + cbeg + c * 16, h);
                tmem_st16(WL + lane_sel + cbeg + c * 16, l);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_wready);
        };
        if (my_tiles > 0) split_w(0);
        for (int it = 0; it < my_tiles && alive; it++) {
            const TcLeaf* Tc = sLeaf[it & (TC_TABS - 1)];
            if (warp == 0) TC_STAMP(20);
            // the next tile's split comes BEFORE this tile's epilogue: GEMM2(it + 1) and GEMM1(it + 2) then run on the tensor pipe
            // while the consumers are busy with the epilogue, instead of the two sides waiting for each other in turn
            if (it + 1 < my_tiles) split_w(it + 1);
            if (!alive) break;
            if (warp == 0) TC_STAMP(22);
            alive = mbar_wait(bar_d2full[it & 1], (it >> 1) & 1, err);
            if (warp == 0) TC_STAMP(23);
            if (!alive) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t D2 = D2A + (it & 1) * 128;
            // epilogue straight from tensor memory: super-row `row`, 32 columns at a time
#pragma unroll 1
            for (int c = 0; c < 2; c++) {
                // forward: the quantiser steps of this thread's 32 positions first -- they do not depend on tensor memory, and their
                // L2 latency would otherwise be paid once per group of 8, serially (table layout [column group][row][8]: a warp reads
                // one contiguous 1 KB run per group)
                uint32_t qv[INV ? 1 : 4][8];
                if (!INV) {
#pragma unroll
                    for (int g = 0; g < 4; g++) {
                        const int j = cbeg + c * 32 + g * 8;
                        const TcLeaf& L = Tc[p * NB + j / S];
                        if (L.base != nullptr) ld_global_v8(L.qf + ((size_t)((j % S) >> 3) * S + li) * 8, qv[INV ? 0 : g]);
                    }
                }
                uint32_t v[32];
                tmem_ld32(D2 + lane_sel + cbeg + c * 32, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const int j = cbeg + c * 32 + g * 8;                   // 8 consecutive columns: inside one leaf (8 | S)
                    const TcLeaf& L = Tc[p * NB + j / S];
                    const int lj = j % S;
                    if (L.base == nullptr) continue;
                    if (INV) {
                        if (li >= L.bh) continue;
                        float* rp = L.base + (size_t)li * L.w + lj;
                        uint32_t o[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) o[k] = __float_as_uint(__fadd_rn(__fdiv_rn(__uint_as_float(v[g * 8 + k]), L.sc), L.mid));   // jpeg.py:452-455
                        if (lj + 7 < L.bw && (reinterpret_cast<uintptr_t>(rp) & 31) == 0) st_global_v8(rp, o);
                        else {
#pragma unroll
                            for (int k = 0; k < 8; k++) if (lj + k < L.bw) rp[k] = __uint_as_float(o[k]);
                        }
                    } else {
                        const int nat = li * S + lj;
                        uint32_t o[8];
#pragma unroll
                        for (int k = 0; k < 8; k++) o[k] = (uint32_t)quantize_f(__uint_as_float(v[g * 8 + k]), __uint_as_float(qv[INV ? 0 : g][k]));
                        if (!L.zig) {
                            int* dst = L.cf + nat;                             // 32-byte aligned unless 2 x 2 leaves precede the block
                            if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) st_global_v8(dst, o);
                            else {
                                *reinterpret_cast<int4*>(dst) = make_int4((int)o[0], (int)o[1], (int)o[2], (int)o[3]);
                                *reinterpret_cast<int4*>(dst + 4) = make_int4((int)o[4], (int)o[5], (int)o[6], (int)o[7]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; k++) L.cf[__ldg(izz + nat + k)] = (int)o[k];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_d2free[it & 1]);
            if (warp == 0) TC_STAMP(24);
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tbase) : "memory");
}

constexpr int tc_slot(int S) { return S == 16 ? 0 : S == 32 ? 1 : S == 64 ? 2 : 3; }

template <int S>
int set_attrs() {
    AEAJ_CUDA(cudaFuncSetAttribute(k_dct_tc<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    AEAJ_CUDA(cudaFuncSetAttribute(k_dct_tc<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    return 0;
}

template <int S>
int launch_one(aeaj_handle* h, const PlaneDesc* planes_dev, const ClassEntry* list, const int* count, int64_t cap, int inverse, cudaStream_t st) {
    constexpr int NL = (TC_N / S) * (TC_N / S);
    const int64_t tiles = std::max<int64_t>((cap + NL - 1) / NL, 1);
    const int blocks = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count);
    const float* a = h->dct_tc_tiles_dev + (size_t)(tc_slot(S) * 2 + (inverse ? 1 : 0)) * 2 * TC_N * TC_N;
    const int* izz = h->tc_izz_dev[tc_slot(S)];
    if (inverse) k_dct_tc<S, true><<<blocks, TC_THREADS, TC_SMEM_BYTES, st>>>(planes_dev, list, count, a, izz, h->tc_err_dev);
    else k_dct_tc<S, false><<<blocks, TC_THREADS, TC_SMEM_BYTES, st>>>(planes_dev, list, count, a, izz, h->tc_err_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// tiles per class S in {16, 32, 64, 128} and direction: [Ah | Al] in the canonical K-major layout, element (row r, k) at
// (k/4)*(128*4) + r*4 + (k%4); forward A = blockdiag(C_S), inverse A = blockdiag(C_S^T)
int aeaj_dct_tc_init(aeaj_handle* h) {
    const int N = TC_N;
    int rc;
    if ((rc = set_attrs<16>()) || (rc = set_attrs<32>()) || (rc = set_attrs<64>()) || (rc = set_attrs<128>())) return rc;
    std::vector<float> host((size_t)4 * 2 * 2 * N * N, 0.0f);
    auto rn = [](float f) { uint32_t b; memcpy(&b, &f, 4); b = (b + 0x1000u) & 0xffffe000u; memcpy(&f, &b, 4); return f; };
    for (int slot = 0; slot < 4; slot++) {
        const int S = 16 << slot;
        for (int inv = 0; inv < 2; inv++) {
            float* dst = host.data() + (size_t)(slot * 2 + inv) * 2 * N * N;
            for (int r = 0; r < N; r++)
                for (int k = 0; k < N; k++) {
                    if (r / S != k / S) continue;                            // off-diagonal blocks stay zero
                    const int rr = r % S, kk = k % S;
                    const int u = inv ? kk : rr, x = inv ? rr : kk;          // C[u][x] = a(u) cos(pi (2x+1) u / 2S)
                    double v = sqrt(2.0 / S) * cos(M_PI * (2 * x + 1) * u / (2.0 * S));
                    if (u == 0) v *= sqrt(0.5);
                    const float c = (float)v;
                    const float hi = rn(c);
                    const size_t o = (size_t)(k / 4) * (N * 4) + (size_t)r * 4 + (k % 4);
                    dst[o] = hi; dst[(size_t)N * N + o] = rn(c - hi);
                }
        }
    }
    AEAJ_CUDA(cudaMalloc(&h->dct_tc_tiles_dev, host.size() * sizeof(float)));
    AEAJ_CUDA(cudaMemcpy(h->dct_tc_tiles_dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    // inverse zigzag permutations (row-major index -> stream position), from the tables aeaj_dct_init built
    size_t total = 0;
    for (int slot = 0; slot < 4; slot++) total += (size_t)(16 << slot) * (16 << slot);
    AEAJ_CUDA(cudaMalloc(&h->tc_izz_all_dev, total * sizeof(int32_t)));
    size_t off = 0;
    for (int slot = 0; slot < 4; slot++) {
        const int S = 16 << slot, n = S * S;
        std::vector<int32_t> zz((size_t)n), izz((size_t)n);
        AEAJ_CUDA(cudaMemcpy(zz.data(), h->zz_dev[4 + slot], zz.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) izz[zz[i]] = i;
        h->tc_izz_dev[slot] = h->tc_izz_all_dev + off;
        AEAJ_CUDA(cudaMemcpy(h->tc_izz_dev[slot], izz.data(), izz.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        off += (size_t)n;
    }
    AEAJ_CUDA(cudaMalloc(&h->tc_err_dev, 32 * sizeof(int)));
    AEAJ_CUDA(cudaMemset(h->tc_err_dev, 0, 32 * sizeof(int)));
    return 0;
}

// size in {16, 32, 64, 128}
int launch_dct_tc(aeaj_handle* h, int size, const PlaneDesc* planes_dev, const ClassEntry* list, const int* count, int64_t cap, int inverse, cudaStream_t st) {
    switch (size) {
        case 16: return launch_one<16>(h, planes_dev, list, count, cap, inverse, st);
        case 32: return launch_one<32>(h, planes_dev, list, count, cap, inverse, st);
        case 64: return launch_one<64>(h, planes_dev, list, count, cap, inverse, st);
        case 128: return launch_one<128>(h, planes_dev, list, count, cap, inverse, st);
    }
    aeaj_set_error("tensor-core DCT: unsupported size %d", size);
    return AEAJ_EINVAL;
}
