// hierarchy.cu -- the LOD cut: expand_to_size + get_interpolation_weights
// (replaces gaussian_hierarchy._C, call sites train_post.py:91-113,
// render_hierarchy.py:58-80).  Semantics per oracle/oracle.c::oracle_expand_to_size /
// oracle_interpolation_weights (UNPINNED: the gaussian-hierarchy source is absent).
//
// The cut is a flat map over all N nodes (a node is emitted when it is finer than the
// target while its parent is not, or when it is too coarse but holds leaf Gaussians),
// followed by an order-preserving compaction: mark -> inclusive scan -> scatter.
// HBM-bound: 28 B node + 32 B box (+ the parent's box, an L2 hit in BFS order) per node.
#include <cub/cub.cuh>
#include <float.h>
#include "common.cuh"

namespace h3dgs {

struct Node { int depth, parent, start, count_leafs, count_merged, start_children, count_children; };

__device__ __forceinline__ float node_size(const float4* __restrict__ boxes, int id, float vx, float vy, float vz) {
    const float4 mn = __ldg(boxes + 2 * (size_t)id), mx = __ldg(boxes + 2 * (size_t)id + 1);
    const bool inside = vx >= mn.x && vx <= mx.x && vy >= mn.y && vy <= mx.y && vz >= mn.z && vz <= mx.z;
    if (inside) return FLT_MAX;
    const float cx = fmaxf(mn.x, fminf(mx.x, vx)) - vx;
    const float cy = fmaxf(mn.y, fminf(mx.y, vy)) - vy;
    const float cz = fmaxf(mn.z, fminf(mx.z, vz)) - vz;
    const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
    return mn.w / dist;
}

__global__ void __launch_bounds__(256)
mark_nodes_kernel(int N, const int* __restrict__ nodes, const float4* __restrict__ boxes, float target,
                  const float* __restrict__ target_dev, const float* __restrict__ viewpoint, int* __restrict__ counts)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (target_dev) target = *target_dev;
    const float vx = viewpoint[0], vy = viewpoint[1], vz = viewpoint[2];
    const int* nd = nodes + 7 * (size_t)n;
    const int depth = nd[0], parent = nd[1], cl = nd[3], cm = nd[4];
    const float size = node_size(boxes, n, vx, vy, vz);
    int count = 0;
    if (size >= target) count = cl;
    else if (parent != -1) {
        const float psize = node_size(boxes, parent, vx, vy, vz);
        if (psize >= target) { count = cl; if (depth != 0) count += cm; }
    }
    counts[n] = count;
}

__global__ void __launch_bounds__(256)
put_render_indices_kernel(int N, const int* __restrict__ nodes, const int* __restrict__ counts,
                          const int* __restrict__ offsets, int* __restrict__ render_indices,
                          int* __restrict__ parent_indices, int* __restrict__ nodes_of_render)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int count = counts[n];
    if (count == 0) return;
    const int off = offsets[n] - count;          // inclusive scan
    const int* nd = nodes + 7 * (size_t)n;
    const int parent = nd[1], start = nd[2];
    const int pg = parent != -1 ? nodes[7 * (size_t)parent + 2] : -1;
    for (int k = 0; k < count; k++) {
        render_indices[off + k] = start + k;
        parent_indices[off + k] = pg;
        nodes_of_render[off + k] = n;
    }
}

// transition weight of node `id` under `parent` (oracle_interpolation_weights)
__device__ __forceinline__ float transition_weight(const float4* __restrict__ boxes, int id, int parent, float target,
                                                   float vx, float vy, float vz)
{
    if (parent == -1) return 1.0f;
    const float psize = node_size(boxes, parent, vx, vy, vz);
    if (psize > 2.0f * target) return 1.0f;
    const float size = node_size(boxes, id, vx, vy, vz);
    const float start = fmaxf(0.5f * psize, size);
    const float diff = psize - start;
    if (diff <= 0) return 1.0f;
    const float tdiff = fmaxf(0.0f, target - start);
    return fmaxf(1.0f - (tdiff / diff), 0.0f);
}

__global__ void __launch_bounds__(256)
interpolation_weights_kernel(int n, const int* __restrict__ node_indices, float target, const int* __restrict__ nodes,
                             const float4* __restrict__ boxes, float vx, float vy, float vz, float* __restrict__ ts,
                             int* __restrict__ kids)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = node_indices[i];
    const int parent = nodes[7 * (size_t)id + 1];
    ts[i] = transition_weight(boxes, id, parent, target, vx, vy, vz);
    kids[i] = parent == -1 ? 1 : nodes[7 * (size_t)parent + 6];
}

// Device-side cut (h3dgs_lod_cut): put_render_indices + interpolation_weights in one kernel -- the
// emitting thread already holds the node and its parent -- plus the count for the consumer.
__global__ void __launch_bounds__(256)
put_cut_kernel(int N, const int* __restrict__ nodes, const float4* __restrict__ boxes, float target,
               const float* __restrict__ target_dev, const float* __restrict__ viewpoint, const int* __restrict__ counts, const int* __restrict__ offsets,
               int* __restrict__ render_indices, int* __restrict__ parent_indices, int* __restrict__ nodes_of_render,
               float* __restrict__ ts, int* __restrict__ kids, int* __restrict__ total)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (target_dev) target = *target_dev;
    if (n == N - 1) *total = offsets[N - 1];
    const int count = counts[n];
    if (count == 0) return;
    const int off = offsets[n] - count;          // inclusive scan
    const int* nd = nodes + 7 * (size_t)n;
    const int parent = nd[1], start = nd[2];
    const int pg = parent != -1 ? nodes[7 * (size_t)parent + 2] : -1;
    const float t = transition_weight(boxes, n, parent, target, viewpoint[0], viewpoint[1], viewpoint[2]);
    const int k = parent == -1 ? 1 : nodes[7 * (size_t)parent + 6];
    for (int j = 0; j < count; j++) {
        render_indices[off + j] = start + j;
        parent_indices[off + j] = pg;
        nodes_of_render[off + j] = n;
        ts[off + j] = t;
        kids[off + j] = k;
    }
}

static size_t expand_scan_bytes(int N) {
    size_t b = 0;
    cub::DeviceScan::InclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, N > 0 ? N : 1);
    return b;
}

}  // namespace h3dgs

using namespace h3dgs;

extern "C" size_t h3dgs_expand_scratch_bytes(int32_t N) {
    const size_t n = (size_t)(N > 0 ? N : 1);
    return align_up(n * 4) * 2 + align_up(expand_scan_bytes(N)) + 256;
}

extern "C" int h3dgs_expand_to_size(int32_t N, const int32_t* nodes, const float* boxes, float target_size,
                                    const float* viewpoint, float, float, float, int32_t* render_indices,
                                    int32_t* parent_indices, int32_t* nodes_for_render_indices, void* scratch,
                                    void* stream)
{
    if (N <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* base = (uint8_t*)scratch;
    int* counts = (int*)base;
    int* offsets = (int*)(base + align_up((size_t)N * 4));
    void* temp = base + 2 * align_up((size_t)N * 4);
    size_t temp_bytes = expand_scan_bytes(N);
    const int blocks = (N + 255) / 256;
    { ProfScope prof(H3DGS_STAGE_LOD_CUT, s);
    mark_nodes_kernel<<<blocks, 256, 0, s>>>(N, nodes, (const float4*)boxes, target_size, nullptr, viewpoint, counts);
    H3_LAUNCHED("mark_nodes", 0, s);
    H3_CUDA(cub::DeviceScan::InclusiveSum(temp, temp_bytes, counts, offsets, N, s));
    H3_LAUNCHED("expand_scan", 0, s);
    put_render_indices_kernel<<<blocks, 256, 0, s>>>(N, nodes, counts, offsets, render_indices, parent_indices,
                                                     nodes_for_render_indices);
    H3_LAUNCHED("put_render_indices", 0, s); }
    void* pin = nullptr;
    if (int rc = pinned_scratch(&pin)) return rc;
    H3_CUDA(cudaMemcpyAsync(pin, offsets + (N - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
    H3_CUDA(cudaStreamSynchronize(s));
    return *static_cast<const int*>(pin);
}

extern "C" int h3dgs_lod_cut(int32_t N, const int32_t* nodes, const float* boxes, float target_size,
                             const float* target_size_dev, const float* viewpoint, int32_t* render_indices, int32_t* parent_indices,
                             int32_t* nodes_for_render_indices, float* ts, int32_t* num_kids, int32_t* count,
                             void* scratch, void* stream)
{
    if (N <= 0) { set_error("lod_cut: empty hierarchy"); return H3DGS_EINVAL; }
    if (!nodes || !boxes || !viewpoint || !render_indices || !parent_indices || !nodes_for_render_indices || !ts ||
        !num_kids || !count || !scratch) { set_error("lod_cut: NULL argument"); return H3DGS_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* base = (uint8_t*)scratch;
    int* counts = (int*)base;
    int* offsets = (int*)(base + align_up((size_t)N * 4));
    void* temp = base + 2 * align_up((size_t)N * 4);
    size_t temp_bytes = expand_scan_bytes(N);
    const int blocks = (N + 255) / 256;
    ProfScope prof(H3DGS_STAGE_LOD_CUT, s);
    // rows after the cut: index -1 = "skip" for the rasterizer (the head is overwritten below)
    H3_CUDA(cudaMemsetAsync(render_indices, 0xFF, (size_t)N * sizeof(int32_t), s));
    mark_nodes_kernel<<<blocks, 256, 0, s>>>(N, nodes, (const float4*)boxes, target_size, target_size_dev, viewpoint, counts);
    H3_LAUNCHED("mark_nodes", 0, s);
    H3_CUDA(cub::DeviceScan::InclusiveSum(temp, temp_bytes, counts, offsets, N, s));
    H3_LAUNCHED("expand_scan", 0, s);
    put_cut_kernel<<<blocks, 256, 0, s>>>(N, nodes, (const float4*)boxes, target_size, target_size_dev, viewpoint, counts, offsets,
                                          render_indices, parent_indices, nodes_for_render_indices, ts, num_kids, count);
    H3_LAUNCHED("put_cut", 0, s);
    return H3DGS_OK;
}

extern "C" int h3dgs_get_interpolation_weights(int32_t n, const int32_t* node_indices, float target_size,
                                               const int32_t* nodes, const float* boxes, float vx, float vy, float vz,
                                               float, float, float, float* ts, int32_t* num_kids, void* stream)
{
    if (n <= 0) return H3DGS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope prof(H3DGS_STAGE_LOD_WEIGHTS, s);
    interpolation_weights_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, node_indices, target_size, nodes,
                                                                 (const float4*)boxes, vx, vy, vz, ts, num_kids);
    H3_LAUNCHED("interpolation_weights", 0, s);
    return H3DGS_OK;
}
