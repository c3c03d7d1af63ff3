// classic_blend.cu -- MEASUREMENT BASELINE ONLY (not product, not linked into libh3dgs.so).
//
// The reference's CUDA rasterizer (hierarchy-rasterizer @ 63fa2476) is absent from
// /root/reference, so bench.py cannot time it.  This file is a stand-in for its two dominant
// kernels in the formulation the 3D-Gaussian-Splatting paper describes: one CTA per 16x16
// tile, every thread fetches one list entry by INDEX (point_list -> per-Gaussian arrays) into
// shared memory per round, one pixel per thread, and in the backward pass every contributing
// pixel issues its own global atomicAdd per gradient value.  It consumes the SAME binned
// state as our kernels (ranges, point_list, unsorted 48-B records), so the comparison
// isolates the blend kernels.  Flat mode only (no hierarchy weight, no depth).
#include <cuda_runtime.h>
#include <stdint.h>

#define TILE 16
#define BLOCK (TILE * TILE)

struct __align__(16) Rec { float4 a, b, c; };   // same record layout as csrc/common.cuh

__global__ void __launch_bounds__(BLOCK)
classic_forward(int W, int H, int gx, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                const Rec* __restrict__ recs, const float* __restrict__ bg, float* __restrict__ out_color,
                float* __restrict__ final_T, uint32_t* __restrict__ n_contrib)
{
    __shared__ float4 s_a[BLOCK], s_b[BLOCK], s_c[BLOCK];
    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % gx, tile_y = blockIdx.x / gx;
    const int px = tile_x * TILE + (tid & 15), py = tile_y * TILE + (tid >> 4);
    const bool inside = px < W && py < H;
    const uint2 range = ranges[blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BLOCK - 1) / BLOCK;
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    uint32_t contributor = 0, last = 0;
    for (int i = 0; i < rounds; i++, todo -= BLOCK) {
        if (__syncthreads_count(done) == BLOCK) break;
        const int progress = i * BLOCK + tid;
        if ((int)range.x + progress < (int)range.y) {
            const Rec r = recs[point_list[range.x + progress]];
            s_a[tid] = r.a; s_b[tid] = r.b; s_c[tid] = r.c;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK, todo); j++) {
            contributor++;
            const float4 a = s_a[j], b = s_b[j];
            const float dx = a.x - (float)px, dy = a.y - (float)py;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, b.y * __expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 c = s_c[j];
            C0 += c.x * alpha * T; C1 += c.y * alpha * T; C2 += c.z * alpha * T;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px, plane = (size_t)H * W;
        final_T[pix] = T; n_contrib[pix] = last;
        out_color[pix] = C0 + T * bg[0]; out_color[plane + pix] = C1 + T * bg[1]; out_color[2 * plane + pix] = C2 + T * bg[2];
    }
}

__global__ void __launch_bounds__(BLOCK)
classic_backward(int W, int H, int gx, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const Rec* __restrict__ recs, const float* __restrict__ bg, const float* __restrict__ final_T,
                 const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix, float* __restrict__ accum)
{
    __shared__ float4 s_a[BLOCK], s_b[BLOCK], s_c[BLOCK];
    __shared__ uint32_t s_id[BLOCK];
    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x % gx, tile_y = blockIdx.x / gx;
    const int px = tile_x * TILE + (tid & 15), py = tile_y * TILE + (tid >> 4);
    const bool inside = px < W && py < H;
    const size_t pix = (size_t)py * W + px, plane = (size_t)H * W;
    const uint2 range = ranges[blockIdx.x];
    const int rounds = ((int)(range.y - range.x) + BLOCK - 1) / BLOCK;
    int todo = (int)(range.y - range.x);
    bool done = !inside;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    uint32_t contributor = todo;
    const int last_contributor = inside ? (int)n_contrib[pix] : 0;
    float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f, g[3] = {0, 0, 0};
    if (inside) { g[0] = dL_dpix[pix]; g[1] = dL_dpix[plane + pix]; g[2] = dL_dpix[2 * plane + pix]; }
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int i = 0; i < rounds; i++, todo -= BLOCK) {
        __syncthreads();
        const int progress = i * BLOCK + tid;
        if ((int)range.x + progress < (int)range.y) {
            const uint32_t id = point_list[range.y - progress - 1];
            const Rec r = recs[id];
            s_id[tid] = id; s_a[tid] = r.a; s_b[tid] = r.b; s_c[tid] = r.c;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(BLOCK, todo); j++) {
            contributor--;
            if ((int)contributor >= last_contributor) continue;
            const float4 a = s_a[j], b = s_b[j];
            const float dx = a.x - (float)px, dy = a.y - (float)py;
            const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
            if (power > 0.0f) continue;
            const float G = __expf(power);
            const float alpha = fminf(0.99f, b.y * G);
            if (alpha < 1.0f / 255.0f) continue;
            T = T / (1.f - alpha);
            const float dchannel_dcolor = alpha * T;
            const float4 c4 = s_c[j];
            const float c[3] = {c4.x, c4.y, c4.z};
            float dL_dalpha = 0.f;
            float* o = accum + (size_t)s_id[j] * 10;
            for (int ch = 0; ch < 3; ch++) {
                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                last_color[ch] = c[ch];
                dL_dalpha += (c[ch] - accum_rec[ch]) * g[ch];
                atomicAdd(o + 6 + ch, dchannel_dcolor * g[ch]);
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            dL_dalpha += (-T_final / (1.f - alpha)) * (bg[0] * g[0] + bg[1] * g[1] + bg[2] * g[2]);
            const float dL_dG = b.y * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * a.z - gdy * a.w, dG_ddely = -gdy * b.x - gdx * a.w;
            atomicAdd(o + 0, dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(o + 1, dL_dG * dG_ddely * ddely_dy);
            atomicAdd(o + 2, -0.5f * gdx * dx * dL_dG);
            atomicAdd(o + 3, -0.5f * gdx * dy * dL_dG);
            atomicAdd(o + 4, -0.5f * gdy * dy * dL_dG);
            atomicAdd(o + 5, G * dL_dalpha);
        }
    }
}

extern "C" int classic_render_forward(int W, int H, const void* ranges, const void* point_list, const void* recs,
                                      const float* bg, float* out_color, float* final_T, void* n_contrib, void* stream)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    classic_forward<<<gx * gy, BLOCK, 0, (cudaStream_t)stream>>>(W, H, gx, (const uint2*)ranges, (const uint32_t*)point_list,
                                                                 (const Rec*)recs, bg, out_color, final_T, (uint32_t*)n_contrib);
    return (int)cudaGetLastError();
}
extern "C" int classic_render_backward(int W, int H, const void* ranges, const void* point_list, const void* recs,
                                       const float* bg, const float* final_T, const void* n_contrib, const float* dL_dpix,
                                       float* accum, void* stream)
{
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    classic_backward<<<gx * gy, BLOCK, 0, (cudaStream_t)stream>>>(W, H, gx, (const uint2*)ranges, (const uint32_t*)point_list,
                                                                  (const Rec*)recs, bg, final_T, (const uint32_t*)n_contrib,
                                                                  dL_dpix, accum);
    return (int)cudaGetLastError();
}
