// peer.cu -- peer memory for the tile-sharded multi-GPU mode (one process per GPU on one NVLink / NVSwitch box):
// exportable allocations (CUDA IPC), and a device-side barrier among the ranks that is an ordinary kernel on the
// caller's stream -- no host synchronisation, capturable in a CUDA graph.
//
// The blend kernels use such memory to fuse their collectives into the compute: the forward stores every finished
// pixel straight into the image buffer of EVERY rank (the "all-gather of rendered tiles" happens tile by tile while
// other tiles are still blending), the backward adds each (tile, Gaussian) gradient row straight into the accumulator
// of the rank that owns the Gaussian (the "reduce-scatter of per-Gaussian gradients" is the red.global.add itself,
// carried by NVLink).  What remains of the collectives is the barrier below.
#include "common.cuh"

namespace h3dgs {

// Flag block of one rank: slot r is written by rank r (its barrier epoch), slot kMaxPeers is this rank's own epoch
// counter, slot kMaxPeers + 1 a sticky timeout marker.
constexpr int kFlagWords = H3DGS_MAX_PEERS + 2;

struct FlagPtrs { uint32_t* p[H3DGS_MAX_PEERS]; };

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One CTA, one thread per rank.  Everything this rank enqueued before the barrier (kernel boundaries order it) is
// released to the system scope together with the epoch; the epochs of all ranks are acquired before anything after
// the barrier runs.  A rank that never arrives would hang every other GPU, so the wait gives up after ~4 s and leaves
// a sticky marker (h3dgs_peer_barrier_status) instead.
__global__ void __launch_bounds__(32)
peer_barrier_kernel(int world, int rank, uint32_t* local, FlagPtrs peers)
{
    __shared__ uint32_t s_epoch;
    const int tid = threadIdx.x;
    if (tid == 0) { s_epoch = local[H3DGS_MAX_PEERS] + 1u; local[H3DGS_MAX_PEERS] = s_epoch; }
    __syncthreads();
    const uint32_t e = s_epoch;
    __threadfence_system();
    if (tid < world) {
        st_release_sys(peers.p[tid] + rank, e);
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(local + tid) - e) < 0) {
            if (clock64() - t0 > 8000000000ll) { local[H3DGS_MAX_PEERS + 1] = 1u; break; }
        }
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace h3dgs

using namespace h3dgs;

extern "C" size_t h3dgs_peer_flag_bytes(void) { return align_up(kFlagWords * sizeof(uint32_t)); }

extern "C" int h3dgs_peer_alloc(size_t bytes, void** ptr)
{
    if (!ptr || bytes == 0) { set_error("peer_alloc: bad arguments"); return H3DGS_EINVAL; }
    H3_CUDA(cudaMalloc(ptr, bytes));
    H3_CUDA(cudaMemset(*ptr, 0, bytes));
    return H3DGS_OK;
}
extern "C" int h3dgs_peer_free(void* ptr) { H3_CUDA(cudaFree(ptr)); return H3DGS_OK; }

extern "C" int h3dgs_peer_export(const void* ptr, void* handle)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == H3DGS_IPC_HANDLE_BYTES, "IPC handle size");
    if (!ptr || !handle) { set_error("peer_export: bad arguments"); return H3DGS_EINVAL; }
    cudaIpcMemHandle_t h;
    H3_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    memcpy(handle, &h, sizeof(h));
    return H3DGS_OK;
}
extern "C" int h3dgs_peer_open(const void* handle, void** ptr)
{
    if (!ptr || !handle) { set_error("peer_open: bad arguments"); return H3DGS_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    H3_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return H3DGS_OK;
}
extern "C" int h3dgs_peer_close(void* ptr) { H3_CUDA(cudaIpcCloseMemHandle(ptr)); return H3DGS_OK; }

extern "C" int h3dgs_peer_barrier(int32_t world, int32_t rank, uint32_t* flags, uint32_t* const* peer_flags, void* stream)
{
    if (world < 1 || world > H3DGS_MAX_PEERS || rank < 0 || rank >= world || !flags || !peer_flags) {
        set_error("peer_barrier: bad arguments (world %d, rank %d)", world, rank); return H3DGS_EINVAL;
    }
    FlagPtrs fp;
    for (int r = 0; r < H3DGS_MAX_PEERS; r++) fp.p[r] = r < world ? peer_flags[r] : nullptr;
    if (fp.p[rank] != flags) { set_error("peer_barrier: peer_flags[rank] must be the local flag block"); return H3DGS_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    peer_barrier_kernel<<<1, 32, 0, s>>>(world, rank, flags, fp);
    H3_LAUNCHED("peer_barrier", 0, s);
    return H3DGS_OK;
}

// 0 = fine, 1 = some barrier of this rank timed out (a peer never arrived); reads 4 bytes back (synchronises `stream`).
extern "C" int h3dgs_peer_barrier_status(const uint32_t* flags, void* stream)
{
    uint32_t v = 0;
    cudaStream_t s = (cudaStream_t)stream;
    H3_CUDA(cudaMemcpyAsync(&v, flags + H3DGS_MAX_PEERS + 1, sizeof(v), cudaMemcpyDeviceToHost, s));
    H3_CUDA(cudaStreamSynchronize(s));
    return (int)v;
}
