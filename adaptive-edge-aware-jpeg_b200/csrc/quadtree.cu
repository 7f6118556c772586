// quadtree.cu -- QuadTree(edge, max, min).get_leaves_and_states() (quadtree.py:68-165) for sm_100a.
//
// The reference walks the tree with a Python stack, calling a numba `np.any` per node.  Here the
// tree is never materialised:
//   1. per "top block" (T = min(max_size, root) pixels square) an OR-pyramid of edge presence over
//      min_size cells is reduced bottom-up in shared memory from the bit-packed edge map;
//   2. split(node) = size > max || (size > min && any_edge(node)); a node exists iff its parent
//      splits; DFS pre-order == ascending Morton order of the node's first cell, larger nodes first;
//   3. exclusive scans in Morton order (inside a block: warp-shuffle scan; across top blocks: one
//      block per plane) give every node its position in the state stream and every leaf its index
//      and coefficient offset; leaves are also appended to per-size-class work lists for the DCT.
// Equivalence with the reference's traversal (incl. '10' states for out-of-bounds children and the
// all-split levels above max_size) is validated by the oracle tests on degenerate shapes.
#include "aeaj_internal.cuh"

namespace {

constexpr int QT_THREADS = 256;
constexpr int QT_OCC_BYTES = 22016;                  // T / min_size <= 128 cells per side: sum_{l} (128 >> l)^2 = 21845, padded

struct QtParams { int min_size, lg_min, max_size; };

struct Scan3 { int a, b, c; };
__device__ __forceinline__ Scan3 block_excl_scan3(Scan3 v, Scan3& total, int* smem /* 3*8 + 3 ints */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Scan3 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int ta = __shfl_up_sync(0xffffffffu, inc.a, o), tb = __shfl_up_sync(0xffffffffu, inc.b, o), tc = __shfl_up_sync(0xffffffffu, inc.c, o);
        if (lane >= o) { inc.a += ta; inc.b += tb; inc.c += tc; }
    }
    __syncthreads();                                   // protect smem reuse across calls
    if (lane == 31) { smem[warp * 3] = inc.a; smem[warp * 3 + 1] = inc.b; smem[warp * 3 + 2] = inc.c; }
    __syncthreads();
    Scan3 off = {0, 0, 0}, tot = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < QT_THREADS / 32; k++) {
        int a = smem[k * 3], b = smem[k * 3 + 1], c = smem[k * 3 + 2];
        if (k < warp) { off.a += a; off.b += b; off.c += c; }
        tot.a += a; tot.b += b; tot.c += c;
    }
    total = tot;
    Scan3 ex = {off.a + inc.a - v.a, off.b + inc.b - v.b, off.c + inc.c - v.c};
    return ex;
}

__device__ __forceinline__ bool cell_has_edge(const uint32_t* __restrict__ bits, int wpr, int h, int w, int x0, int y0, int c) {
    int y1 = min(y0 + c, h), x1 = min(x0 + c, w);
    if (x0 >= w || y0 >= h) return false;
    int w0 = x0 >> 5, w1 = (x1 - 1) >> 5;
    for (int y = y0; y < y1; y++)
        for (int wi = w0; wi <= w1; wi++) {
            unsigned v = __ldg(bits + (size_t)y * wpr + wi);
            int lo = max(x0 - wi * 32, 0), hi = min(x1 - wi * 32, 32);   // bit range [lo, hi)
            unsigned m = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            if (v & m) return true;
        }
    return false;
}

// phase 0: totals per top block.  phase 1: emit states / leaves / class entries.
// grid: (max top blocks over planes, nplanes)
__global__ void __launch_bounds__(QT_THREADS) k_qt_blocks(const PlaneDesc* __restrict__ planes, const __grid_constant__ TileMap tm, QtParams q, int phase,
                                                          ClassEntry* __restrict__ class_lists, int* __restrict__ class_counts,
                                                          const long long* __restrict__ class_offsets) {
    int plane_i, bx, by;
    tile_decode(tm, blockIdx.x, plane_i, bx, by);
    const PlaneDesc& P = planes[plane_i];
    const int tb = by * P.ntx + bx;
    const int T = P.top, c = q.min_size;
    if (by * T < P.ry0 || by * T >= P.ry1) return;     // halo-split: only the top blocks of this call's band (bands are multiples of T rows)
    const int n = T / c;                               // cells per side (power of two, >= 1)
    int L = 0; while ((1 << L) < n) L++;
    const int X0 = bx * T, Y0 = by * T;
    __shared__ uint8_t occ[QT_OCC_BYTES];
    __shared__ int lvl_off[9];
    __shared__ int s_scan[3 * 8 + 4];
    __shared__ int s_cls[9], s_cls_base[9];
    const int tid = threadIdx.x;
    if (tid == 0) { int o = 0; for (int l = 0; l <= L; l++) { lvl_off[l] = o; o += (n >> l) * (n >> l); } }
    if (tid < 9) { s_cls[tid] = 0; s_cls_base[tid] = 0; }
    __syncthreads();
    // level 0 from the bitmap.  For cells narrower than a word, one thread ORs the c rows of a
    // (cell row, bitmap word) pair and splits the word into 32/c cell flags.
    if (c <= 32) {
        const int cpw = 32 / c;                                    // cells per word
        const int wpt = (T + 31) / 32;                             // words per top-block row (a power of two)
        const int lg_wpt = 31 - __clz(wpt);
        const int w0 = X0 >> 5;                                    // X0 is a multiple of T; if T < 32, bits are offset
        const int bit0 = X0 & 31;
        for (int i = tid; i < n * wpt; i += QT_THREADS) {
            const int cy = i >> lg_wpt, wi = i & (wpt - 1);
            const int y0 = Y0 + cy * c, y1 = min(y0 + c, P.h);
            unsigned acc = 0;
            if (w0 + wi < P.wpr)
                for (int y = y0; y < y1; y++) acc |= __ldg(P.strong + (size_t)y * P.wpr + w0 + wi);
            acc >>= bit0;                                          // only non-zero when T < 32 (then wpt == 1)
            const int ncell = min(cpw, n - wi * cpw);
            const unsigned m = (c == 32) ? 0xffffffffu : ((1u << c) - 1u);
            for (int k = 0; k < ncell; k++) occ[cy * n + wi * cpw + k] = ((acc >> (k * c)) & m) != 0;
        }
    } else {
        for (int i = tid; i < n * n; i += QT_THREADS) {
            int cy = i >> L, cx = i & (n - 1);
            occ[i] = cell_has_edge(P.strong, P.wpr, P.h, P.w, X0 + cx * c, Y0 + cy * c, c);
        }
    }
    __syncthreads();
    for (int l = 1; l <= L; l++) {
        const int nl = n >> l, np = n >> (l - 1);
        const uint8_t* prev = occ + lvl_off[l - 1];
        uint8_t* cur = occ + lvl_off[l];
        for (int i = tid; i < nl * nl; i += QT_THREADS) {
            int j = i >> (L - l), k = i & (nl - 1);
            cur[i] = prev[(2 * j) * np + 2 * k] | prev[(2 * j) * np + 2 * k + 1] | prev[(2 * j + 1) * np + 2 * k] | prev[(2 * j + 1) * np + 2 * k + 1];
        }
        __syncthreads();
    }
    // node predicates ---------------------------------------------------------------------------
    auto node_split = [&](int l, int j, int k) -> bool { return l > 0 && occ[lvl_off[l] + (j << (L - l)) + k]; };
    auto node_exists = [&](int l, int j, int k) -> bool { return l == L ? true : node_split(l + 1, j >> 1, k >> 1); };
    auto node_inb = [&](int l, int j, int k) -> bool { return (X0 + (k << l) * c) < P.w && (Y0 + (j << l) * c) < P.h; };

    // Every thread owns `per` consecutive Morton positions (a 2x2 quad of cells for the usual 32x32 block), so one
    // block scan serves the whole top block; a node of level l starts at z iff z has 2l trailing zero bits.
    const int ncell = n * n;
    const int per = (ncell + QT_THREADS - 1) / QT_THREADS;
    const int zb = tid * per;
    Scan3 cnt = {0, 0, 0};
    for (int t = 0; t < per; t++) {
        const int z = zb + t;
        if (z >= ncell) break;
        const int cx = (int)compact1by1((uint32_t)z), cy = (int)compact1by1((uint32_t)z >> 1);
        const int lmax = z ? min(L, (__ffs(z) - 1) >> 1) : L;
        for (int l = lmax; l >= 0; l--) {
            const int j = cy >> l, k = cx >> l;
            if (!node_exists(l, j, k)) continue;
            cnt.a++;
            if (node_inb(l, j, k) && !node_split(l, j, k)) {
                cnt.b++; const int sz = c << l; cnt.c += sz * sz;
                if (phase == 1) atomicAdd(&s_cls[l], 1);           // per-class leaf counts of this block
            }
        }
    }
    Scan3 tot;
    const Scan3 ex = block_excl_scan3(cnt, tot, s_scan);
    if (phase == 0) {
        if (tid == 0) { P.tb_tot[tb] = make_int2(tot.a, tot.b); P.tb_coef[tb] = tot.c; }
        return;
    }
    // phase 1: reserve this block's ranges in the global class lists, then emit in Morton order
    __syncthreads();
    if (tid <= L && s_cls[tid] > 0) s_cls_base[tid] = atomicAdd(&class_counts[q.lg_min + tid], s_cls[tid]);
    __syncthreads();
    if (tid < 9) s_cls[tid] = 0;
    __syncthreads();
    const int4 base = P.tb_base[tb];
    int spos = base.x + ex.a, li = base.y + ex.b, co = base.z + ex.c;
    for (int t = 0; t < per; t++) {
        const int z = zb + t;
        if (z >= ncell) break;
        const int cx = (int)compact1by1((uint32_t)z), cy = (int)compact1by1((uint32_t)z >> 1);
        const int lmax = z ? min(L, (__ffs(z) - 1) >> 1) : L;
        for (int l = lmax; l >= 0; l--) {
            const int j = cy >> l, k = cx >> l;
            if (!node_exists(l, j, k)) continue;
            const bool inb = node_inb(l, j, k), sp = node_split(l, j, k);
            P.states[spos++] = inb ? (sp ? 1 : 0) : 2;
            if (inb && !sp) {
                const int sz = c << l, x = X0 + (k << l) * c, y = Y0 + (j << l) * c;
                reinterpret_cast<int4*>(P.leaves)[li] = make_int4(x, y, sz, co);
                const int r = s_cls_base[l] + atomicAdd(&s_cls[l], 1);
                ClassEntry e; e.x = x; e.y = y; e.plane = plane_i; e.coef_off = co;
                class_lists[class_offsets[q.lg_min + l] + r] = e;
                li++; co += sz * sz;
            }
        }
    }
}

// one block per plane: Morton-order scan over all (root/T)^2 top positions, emitting the states of
// the all-split levels above T and the '10' states of out-of-bounds positions.
__global__ void __launch_bounds__(QT_THREADS) k_qt_scan(const PlaneDesc* __restrict__ planes) {
    const PlaneDesc& P = planes[blockIdx.x];
    __shared__ int s_scan[3 * 8 + 4];
    const int T = P.top, R = P.root;
    const int side = R / T;                            // power of two
    int U = 0; while ((1 << U) < side) U++;
    const int M = side * side;
    Scan3 carry = {0, 0, 0};
    for (int m0 = 0; m0 < M; m0 += QT_THREADS) {
        const int m = m0 + threadIdx.x;
        Scan3 cnt = {0, 0, 0};
        int n_upper = 0, top_kind = 0;                 // 0 none, 1 in-bounds subtree, 2 single '10'
        int bx = 0, by = 0;
        if (m < M) {
            bx = (int)compact1by1((uint32_t)m); by = (int)compact1by1((uint32_t)m >> 1);
            const int x = bx * T, y = by * T;
            for (int u = U; u >= 1; u--) {
                if (m & ((1 << (2 * u)) - 1)) continue;
                long long size = (long long)T << u;
                bool inb = x < P.w && y < P.h;
                bool pinb = true;
                if (u < U) { long long ps = size * 2; long long px = (x / ps) * ps, py = (y / ps) * ps; pinb = px < P.w && py < P.h; }
                if (inb || pinb) n_upper++;
            }
            bool inb = x < P.w && y < P.h, pinb = true;
            if (U >= 1) { long long ps = (long long)T * 2; long long px = (x / ps) * ps, py = (y / ps) * ps; pinb = px < P.w && py < P.h; }
            if (inb) {
                int tb = by * P.ntx + bx;
                int2 t = P.tb_tot[tb];
                cnt.a = t.x; cnt.b = t.y; cnt.c = P.tb_coef[tb];
                top_kind = 1;
            } else if (pinb) { cnt.a = 1; top_kind = 2; }
            cnt.a += n_upper;
        }
        Scan3 tot;
        Scan3 ex = block_excl_scan3(cnt, tot, s_scan);
        if (m < M) {
            int pos = carry.a + ex.a;
            const int x = bx * T, y = by * T;
            for (int u = U; u >= 1; u--) {
                if (m & ((1 << (2 * u)) - 1)) continue;
                long long size = (long long)T << u;
                bool inb = x < P.w && y < P.h;
                bool pinb = true;
                if (u < U) { long long ps = size * 2; long long px = (x / ps) * ps, py = (y / ps) * ps; pinb = px < P.w && py < P.h; }
                if (inb) P.states[pos++] = 1;          // size > max: always split (quadtree.py:118)
                else if (pinb) P.states[pos++] = 2;
            }
            if (top_kind == 1) P.tb_base[by * P.ntx + bx] = make_int4(pos, carry.b + ex.b, carry.c + ex.c, 0);
            else if (top_kind == 2) P.states[pos] = 2;
        }
        carry.a += tot.a; carry.b += tot.b; carry.c += tot.c;
    }
    if (threadIdx.x == 0) { P.counts[0] = carry.b; P.counts[1] = carry.a; P.counts[2] = carry.c; P.counts[3] = R; }
}

// decode side: leaves (x,y,size,coef_off) -> per-size-class work lists
// A leaf list that did not come from aeaj_quadtree / aeaj_states_to_leaves_host is not trusted: leaves whose size is not a
// power of two inside [2^lg_min, 2^lg_max], whose origin lies outside the layer or whose coefficient block leaves the
// plane's buffer are skipped and counted in class_counts[15] (reported through aeaj_decode_io.status).
// class_offsets: [0..8] first slot of every size class in class_lists, [9..17] capacity of the class.
__global__ void __launch_bounds__(256) k_bucket_leaves(const PlaneDesc* __restrict__ planes, ClassEntry* __restrict__ class_lists,
                                                       int* __restrict__ class_counts, const long long* __restrict__ class_offsets,
                                                       int lg_min, int lg_max) {
    const PlaneDesc& P = planes[blockIdx.y];
    const int nl = (int)min((long long)max(P.counts[0], 0), (long long)P.cap_leaves);
    __shared__ int s_cls[9], s_base[9];
    if (threadIdx.x < 9) s_cls[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    int4 lf = make_int4(0, 0, 0, 0);
    int lg = -1, rank = 0;
    if (i < nl) {
        lf = reinterpret_cast<const int4*>(P.leaves)[i];
        lg = lf.z > 0 ? 31 - __clz(lf.z) : -1;
        const bool ok = lg >= lg_min && lg <= lg_max && lf.z == (1 << lg) && lf.x >= 0 && lf.y >= 0 && lf.x < P.w && lf.y < P.h &&
                        lf.w >= 0 && (long long)lf.w + (long long)lf.z * lf.z <= (long long)P.cap_coef;
        // multi-GPU halo-split: the leaf slots other ranks own are zero-filled on this rank (api.cu, AEAJ_PHASE_COLOR): not an error
        const bool absent = lf.z == 0 && (P.ry0 > 0 || P.ry1 < P.h);
        if (ok) rank = atomicAdd(&s_cls[lg], 1);
        else { lg = -1; if (!absent) atomicAdd(&class_counts[15], 1); }
    }
    __syncthreads();
    if (threadIdx.x < 9 && s_cls[threadIdx.x] > 0) s_base[threadIdx.x] = atomicAdd(&class_counts[threadIdx.x], s_cls[threadIdx.x]);
    __syncthreads();
    if (i < nl && lg >= 0) {
        ClassEntry e; e.x = lf.x; e.y = lf.y; e.plane = blockIdx.y; e.coef_off = lf.w;
        const long long slot = (long long)s_base[lg] + rank;
        if (slot < class_offsets[9 + lg]) class_lists[class_offsets[lg] + slot] = e;      // more leaves of a size than can tile the planes: overlapping leaves
        else atomicAdd(&class_counts[15], 1);
    }
}

// 2 bits per state, MSB first, zero padded (jpeg.py:563-571); grid: (byte chunks, planes)
__global__ void __launch_bounds__(256) k_pack_states(const PlaneDesc* __restrict__ planes) {
    const PlaneDesc& P = planes[blockIdx.y];
    if (!P.packed_states) return;
    const int n = P.counts[1], nb = (n + 3) >> 2;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nb; i += gridDim.x * 256) {
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const int idx = 4 * i + k; v = (v << 2) | (idx < n ? (P.states[idx] & 3u) : 0u); }
        P.packed_states[i] = (uint8_t)v;
    }
}

}  // namespace

int launch_pack_states(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, cudaStream_t st) {
    int64_t maxs = 1;
    for (int i = 0; i < nplanes; i++) maxs = std::max<int64_t>(maxs, P[i].cap_states);
    dim3 grd((unsigned)std::min<int64_t>(aeaj_cdiv64(aeaj_cdiv64(maxs, 4), 256), 64), nplanes);
    k_pack_states<<<grd, 256, 0, st>>>(planes_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int launch_quadtree(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, int min_size, int max_size,
                    ClassEntry* class_lists, int* class_counts, const long long* class_offsets_dev, cudaStream_t st, int* launches, int parts) {
    QtParams q; q.min_size = min_size; q.lg_min = ilog2i(min_size); q.max_size = max_size;
    const TileMap tm = make_tile_map(P, nplanes, 0, 0, true);
    const int grd = tile_map_total(tm, nplanes);
    if (parts & 1) {
        k_qt_blocks<<<grd, QT_THREADS, 0, st>>>(planes_dev, tm, q, 0, class_lists, class_counts, class_offsets_dev);
        AEAJ_LAUNCH_CHECK();
        if (launches) (*launches)++;
    }
    if (parts & 2) {
        k_qt_scan<<<nplanes, QT_THREADS, 0, st>>>(planes_dev);
        AEAJ_LAUNCH_CHECK();
        if (launches) (*launches)++;
    }
    if (parts & 4) {
        k_qt_blocks<<<grd, QT_THREADS, 0, st>>>(planes_dev, tm, q, 1, class_lists, class_counts, class_offsets_dev);
        AEAJ_LAUNCH_CHECK();
        if (launches) (*launches)++;
    }
    return 0;
}

int launch_bucket_leaves(const PlaneDesc* planes_dev, const PlaneDesc* P, int nplanes, ClassEntry* class_lists,
                         int* class_counts, const long long* class_offsets_dev, int lg_min, int lg_max, cudaStream_t st) {
    int64_t maxl = 1;
    for (int i = 0; i < nplanes; i++) maxl = std::max<int64_t>(maxl, P[i].cap_leaves);
    dim3 grd((unsigned)aeaj_cdiv64(maxl, 256), nplanes);
    k_bucket_leaves<<<grd, 256, 0, st>>>(planes_dev, class_lists, class_counts, class_offsets_dev, lg_min, lg_max);
    AEAJ_LAUNCH_CHECK();
    return 0;
}
