// peer.cu -- one image over several GPUs of one node (BASELINE config C4): shared device memory, a
// device-side barrier and gathers over NVLink, without NCCL on the data path.
//
// Every rank allocates its plan workspace with aeaj_peer_alloc and exports it (CUDA IPC); the ranks open each other's
// workspaces, so a buffer that lives at offset o of this rank's workspace lives at offset o of every peer's: kernels
// reach a neighbour's copy of a plane by adding a byte delta to their own pointer.  The stencil kernels read the few
// halo rows outside their band straight from the neighbour (prefilter 3 rows, Sobel / NMS 2, chroma upsampling 1),
// the CLAHE / percentile kernels sum the per-rank partial histograms on the fly, and two small gathers copy the other
// ranks' rows of the strong / weak bitmaps (hysteresis runs replicated) and of the quadtree block totals.  Between the
// phases the ranks meet at k_peer_barrier: rank r stores the epoch into slot r of every peer's flag array and spins on
// its own array -- one tiny kernel per rank, a few microseconds over NVLink, ordered with the data by the stream.
// All waits are bounded (error flag instead of a hang).  Peer data is read with ld.cv: never from a stale cache line.
#include "aeaj_internal.cuh"

namespace {

struct PeerFlags { int* f[AEAJ_MAX_PEERS]; };

__global__ void k_peer_barrier(PeerFlags F, int rank, int world, int epoch, int* err) {
    const int t = threadIdx.x;
    if (t < world && t != rank) {
        __threadfence_system();                                            // everything this GPU wrote before is visible to the peers
        *reinterpret_cast<volatile int*>(F.f[t] + rank) = epoch;           // "rank has reached `epoch`", into rank t's array
        const volatile int* mine = F.f[rank] + t;
        const long long t0 = clock64();
        while (*mine < epoch)
            if (clock64() - t0 > 6000000000ll) { if (err) atomicExch(err, 2); break; }    // ~3 s: a missing rank must not hang the GPU
        __threadfence_system();
    }
}

// one byte range, by the whole grid row: 128-bit accesses when both ends are 16-byte aligned, words when 4-byte aligned, bytes
// otherwise (row b > 0 of a uint8 [B][cap] buffer need not be word aligned).  CV: the source is peer memory (never a stale line).
template <bool CV>
__device__ __forceinline__ void copy_range(const PeerSeg& s) {
    const long long stride = (long long)gridDim.x * 256, first = (long long)blockIdx.x * 256 + threadIdx.x;
    const uintptr_t both = (uintptr_t)s.src | (uintptr_t)s.dst;
    if ((both & 3) == 0) {
        const long long n4 = s.bytes >> 2;
        const unsigned* src = reinterpret_cast<const unsigned*>(s.src);
        unsigned* dst = reinterpret_cast<unsigned*>(s.dst);
        long long done4 = 0;
        if ((both & 15) == 0) {
            const long long n16 = n4 >> 2;
            for (long long i = first; i < n16; i += stride)
                reinterpret_cast<uint4*>(dst)[i] = CV ? __ldcv(reinterpret_cast<const uint4*>(src) + i) : reinterpret_cast<const uint4*>(src)[i];
            done4 = n16 * 4;
        }
        for (long long i = done4 + first; i < n4; i += stride) dst[i] = CV ? __ldcv(src + i) : src[i];
        for (long long i = n4 * 4 + first; i < s.bytes; i += stride)
            reinterpret_cast<unsigned char*>(s.dst)[i] = reinterpret_cast<const unsigned char*>(s.src)[i];
    } else {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(s.src);
        unsigned char* dst = reinterpret_cast<unsigned char*>(s.dst);
        for (long long i = first; i < s.bytes; i += stride) dst[i] = CV ? __ldcv(src + i) : src[i];
    }
}

// copies `nseg` byte ranges from peer memory into local memory; grid: (blocks, nseg)
__global__ void __launch_bounds__(256) k_peer_gather(const PeerSeg* __restrict__ segs) { copy_range<true>(segs[blockIdx.y]); }

// the same copy with the table in the kernel parameters (<= 64 ranges, 1.5 KB): nothing to upload, so nothing that could synchronise
struct SegParams { PeerSeg s[AEAJ_SEGS_BY_PARAM]; };
__global__ void __launch_bounds__(256) k_copy_segments(const __grid_constant__ SegParams T) { copy_range<false>(T.s[blockIdx.y]); }

}  // namespace

int launch_copy_segments_param(const PeerSeg* segs_host, int nseg, long long max_bytes, cudaStream_t st) {
    if (nseg <= 0) return 0;
    SegParams T;
    memset(&T, 0, sizeof T);
    for (int i = 0; i < nseg && i < AEAJ_SEGS_BY_PARAM; i++) T.s[i] = segs_host[i];
    const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((max_bytes / 16 + 255) / 256, 148));
    k_copy_segments<<<dim3(blocks, nseg), 256, 0, st>>>(T);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int launch_peer_barrier(int* const* flags_host, int rank, int world, int epoch, int* err_dev, cudaStream_t st) {
    PeerFlags F;
    for (int i = 0; i < AEAJ_MAX_PEERS; i++) F.f[i] = i < world ? flags_host[i] : nullptr;
    k_peer_barrier<<<1, 32, 0, st>>>(F, rank, world, epoch, err_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

int launch_peer_gather(const PeerSeg* segs_dev, int nseg, long long max_bytes, cudaStream_t st) {
    if (nseg <= 0) return 0;
    const unsigned blocks = (unsigned)std::max<long long>(1, std::min<long long>((max_bytes / 16 + 255) / 256, 64));
    k_peer_gather<<<dim3(blocks, nseg), 256, 0, st>>>(segs_dev);
    AEAJ_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// shared allocations (CUDA IPC)
// ---------------------------------------------------------------------------------------------
extern "C" int aeaj_peer_alloc(size_t bytes, void** ptr) {
    AEAJ_REQUIRE(ptr && bytes > 0, "aeaj_peer_alloc: bad arguments");
    AEAJ_CUDA(cudaMalloc(ptr, bytes));                                     // a plain cudaMalloc block: exportable with cudaIpcGetMemHandle
    AEAJ_CUDA(cudaMemset(*ptr, 0, bytes));
    return 0;
}
extern "C" int aeaj_peer_free(void* ptr) {
    if (ptr) AEAJ_CUDA(cudaFree(ptr));
    return 0;
}
extern "C" int aeaj_peer_export(void* ptr, void* handle64_host) {
    AEAJ_REQUIRE(ptr && handle64_host, "aeaj_peer_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    AEAJ_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64_host, &h, 64);
    return 0;
}
extern "C" int aeaj_peer_open(const void* handle64_host, void** ptr) {
    AEAJ_REQUIRE(ptr && handle64_host, "aeaj_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    AEAJ_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int aeaj_peer_close(void* ptr) {
    if (ptr) AEAJ_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}
