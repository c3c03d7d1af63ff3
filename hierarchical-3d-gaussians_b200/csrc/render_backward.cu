// render_backward.cu -- K7: per-tile gradient replay (replaces BACKWARD::render).
// Semantics per oracle/oracle.c::oracle_render_backward.
//
// v1 ("pixel-parallel"): one CTA per tile, records streamed back-to-front with the
// same TMA double buffer as the forward; per-entry partials are reduced over the 32
// pixels of a warp with shuffles and lane 0 issues one red.global.add per value --
// 32x fewer atomics than the classic one-atomic-per-pixel formulation.
#include "common.cuh"
#include "tma.cuh"

namespace h3dgs {

constexpr int kBwdBatch = 256;
constexpr int kBwdStages = 2;

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_down_sync(0xffffffffu, v, 16);
    v += __shfl_down_sync(0xffffffffu, v, 8);
    v += __shfl_down_sync(0xffffffffu, v, 4);
    v += __shfl_down_sync(0xffffffffu, v, 2);
    v += __shfl_down_sync(0xffffffffu, v, 1);
    return v;
}

template <bool HIER, bool DEPTH>
__global__ void __launch_bounds__(256)
render_backward_kernel(int W, int H, int gx, int shard_count, int shard_index, const uint2* __restrict__ ranges,
                       const Record* __restrict__ sorted, const uint32_t* __restrict__ point_list,
                       const float* __restrict__ bg, const float* __restrict__ final_T,
                       const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ tile_max_contrib,
                       const float* __restrict__ dL_dcolor, const float* __restrict__ dL_dinvdepth,
                       float* __restrict__ accum)
{
    __shared__ __align__(128) Record s_rec[kBwdStages][kBwdBatch];
    __shared__ uint32_t s_id[kBwdStages][kBwdBatch];
    __shared__ __align__(8) uint64_t s_full[kBwdStages];

    const int tid = threadIdx.x, lane = tid & 31;
    const int tile_x = blockIdx.x % gx;
    const int tile_y = (blockIdx.x / gx) * shard_count + shard_index;
    const int tile = tile_y * gx + tile_x;
    const uint2 range = ranges[tile];
    const int n = min((int)(range.y - range.x), (int)tile_max_contrib[tile]);   // nothing beyond the last contributor
    const int nb = (n + kBwdBatch - 1) / kBwdBatch;
    if (nb == 0) return;
    const Record* src = sorted + range.x;
    const uint32_t* ids = point_list + range.x;

    // Gaussian ids of a batch are staged with plain loads (their global address is only
    // 4-B aligned, below the 16-B granularity of bulk copies).  iteration it = 0..nb-1
    // handles batch b = nb-1-it (back to front).
    auto stage_ids = [&](int it) {
        const int b = nb - 1 - it, st = it % kBwdStages;
        const int e = b * kBwdBatch + tid;
        if (e < n) s_id[st][tid] = ids[e];
    };
    for (int it = 0; it < kBwdStages && it < nb; it++) stage_ids(it);
    if (tid == 0) {
        for (int s = 0; s < kBwdStages; s++) mbar_init(&s_full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int it) {
        const int b = nb - 1 - it, st = it % kBwdStages;
        const uint32_t bytes = (uint32_t)min(kBwdBatch, n - b * kBwdBatch) * (uint32_t)sizeof(Record);
        mbar_arrive_expect_tx(&s_full[st], bytes);
        tma_load_1d(&s_rec[st][0], src + (size_t)b * kBwdBatch, bytes, &s_full[st]);
    };
    if (tid == 0)
        for (int it = 0; it < kBwdStages && it < nb; it++) issue(it);

    const int px = tile_x * kTile + (tid & 15), py = tile_y * kTile + (tid >> 4);
    const bool inside = px < W && py < H;
    const float fpx = (float)px, fpy = (float)py;
    const size_t pix = (size_t)py * W + px, plane = (size_t)H * W;
    const float T_final = inside ? final_T[pix] : 0.f;
    float T = T_final;
    const int last = inside ? (int)n_contrib[pix] : 0;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f;
    if (inside) {
        g0 = dL_dcolor[pix]; g1 = dL_dcolor[plane + pix]; g2 = dL_dcolor[2 * plane + pix];
        if (DEPTH) gd = dL_dinvdepth[pix];
    }
    const float bg_dot = bg[0] * g0 + bg[1] * g1 + bg[2] * g2;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accd = 0.f;       // accum_rec
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, lcd = 0.f, last_alpha = 0.f;
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

    for (int it = 0; it < nb; it++) {
        const int st = it % kBwdStages, b = nb - 1 - it;
        mbar_wait(&s_full[st], (uint32_t)((it / kBwdStages) & 1));
        const int cnt = min(kBwdBatch, n - b * kBwdBatch);
        const Record* rec = &s_rec[st][0];
        for (int j = cnt - 1; j >= 0; j--) {
            const int e = b * kBwdBatch + j;              // 0-based list position; contributor number e+1
            float d_mx = 0.f, d_my = 0.f, d_cx = 0.f, d_cy = 0.f, d_cz = 0.f, d_op = 0.f;
            float d_r = 0.f, d_g = 0.f, d_b = 0.f, d_iv = 0.f;
            bool active = false;
            if (e < last) {
                const float4 a = rec[j].a;
                const float4 bb = rec[j].b;
                const float dx = a.x - fpx, dy = a.y - fpy;
                const float power = -0.5f * (a.z * dx * dx + bb.x * dy * dy) - a.w * dx * dy;
                if (power <= 0.0f) {
                    const float G = fast_exp(power);
                    const float abase = fminf(kAlphaCap, bb.y * G);
                    float alpha, dadb;
                    hier_alpha_grad<HIER>(abase, bb.z, __float_as_uint(bb.w), alpha, dadb);
                    if (alpha >= kAlphaSkip) {
                        active = true;
                        const float4 c = rec[j].c;
                        T = T / (1.f - alpha);
                        const float dchannel_dcolor = alpha * T;
                        float dL_dalpha = 0.f;
                        acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0; lc0 = c.x; dL_dalpha += (c.x - acc0) * g0;
                        acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1; lc1 = c.y; dL_dalpha += (c.y - acc1) * g1;
                        acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2; lc2 = c.z; dL_dalpha += (c.z - acc2) * g2;
                        d_r = dchannel_dcolor * g0; d_g = dchannel_dcolor * g1; d_b = dchannel_dcolor * g2;
                        if (DEPTH) {
                            accd = last_alpha * lcd + (1.f - last_alpha) * accd; lcd = c.w;
                            dL_dalpha += (c.w - accd) * gd;
                            d_iv = dchannel_dcolor * gd;
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                        const float dL_dab = dL_dalpha * dadb;
                        const float dL_dG = bb.y * dL_dab;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = -gdx * a.z - gdy * a.w;
                        const float dG_ddely = -gdy * bb.x - gdx * a.w;
                        d_mx = dL_dG * dG_ddelx * ddelx_dx;
                        d_my = dL_dG * dG_ddely * ddely_dy;
                        d_cx = -0.5f * gdx * dx * dL_dG;
                        d_cy = -0.5f * gdx * dy * dL_dG;
                        d_cz = -0.5f * gdy * dy * dL_dG;
                        d_op = G * dL_dab;
                    }
                }
            }
            if (__any_sync(0xffffffffu, active)) {
                d_mx = warp_sum(d_mx); d_my = warp_sum(d_my);
                d_cx = warp_sum(d_cx); d_cy = warp_sum(d_cy); d_cz = warp_sum(d_cz);
                d_op = warp_sum(d_op);
                d_r = warp_sum(d_r); d_g = warp_sum(d_g); d_b = warp_sum(d_b);
                if (DEPTH) d_iv = warp_sum(d_iv);
                if (lane == 0) {
                    float* o = accum + (size_t)s_id[st][j] * kAccum;
                    atomicAdd(o + 0, d_mx); atomicAdd(o + 1, d_my);
                    atomicAdd(o + 2, d_cx); atomicAdd(o + 3, d_cy); atomicAdd(o + 4, d_cz);
                    atomicAdd(o + 5, d_op);
                    atomicAdd(o + 6, d_r); atomicAdd(o + 7, d_g); atomicAdd(o + 8, d_b);
                    if (DEPTH) atomicAdd(o + 9, d_iv);
                }
            }
        }
        __syncthreads();                      // every thread is done with stage st (records and ids)
        if (it + kBwdStages < nb) {
            if (tid == 0) issue(it + kBwdStages);
            stage_ids(it + kBwdStages);       // read two iterations later, after the next __syncthreads
        }
    }
}

int launch_render_backward(const h3dgs_raster_args& a, const uint32_t* ranges, const Record* sorted_records,
                           const uint32_t* point_list, const float* final_T, const uint32_t* n_contrib,
                           const uint32_t* tile_max_contrib, const float* dL_dcolor, const float* dL_dinvdepth,
                           float* accum, cudaStream_t s)
{
    const int W = a.image_width, H = a.image_height;
    const int gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
    const int sc = a.shard_count > 0 ? a.shard_count : 1, si = a.shard_count > 0 ? a.shard_index : 0;
    const int rows = (gy + sc - 1 - si) / sc;
    if (rows <= 0 || gx <= 0) return H3DGS_OK;
    const bool hier = a.interpolation_weights != nullptr;
    const bool depth = a.do_depth != 0 && dL_dinvdepth != nullptr;
    const dim3 grid(gx * rows), block(256);
    ProfScope prof(H3DGS_STAGE_RENDER_BWD, s);
#define LAUNCH(HI, DE)                                                                                          \
    render_backward_kernel<HI, DE><<<grid, block, 0, s>>>(W, H, gx, sc, si, (const uint2*)ranges, sorted_records, \
                                                          point_list, a.bg, final_T, n_contrib, tile_max_contrib, \
                                                          dL_dcolor, dL_dinvdepth, accum)
    if (hier) { if (depth) LAUNCH(true, true); else LAUNCH(true, false); }
    else      { if (depth) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    H3_LAUNCHED("render_backward", a.debug, s);
    return H3DGS_OK;
}

}  // namespace h3dgs
