// render_forward.cu -- K6: per-tile front-to-back alpha blend (replaces FORWARD::render).
// Semantics per oracle/oracle.c::oracle_render_forward.
//
// B200 design: one CTA per 16x16 tile.  The tile's depth-sorted records were
// materialised contiguously by binning.cu, so a batch of 256 records is ONE 12 KB
// cp.async.bulk (TMA) transfer into shared memory, double-buffered behind an
// mbarrier; no thread spends registers or LSU issue slots on staging.  Every pixel
// thread then reads records as shared-memory broadcasts.  Early termination: a
// CTA-wide vote (__syncthreads_count) per batch, which doubles as the "stage free"
// signal for the producer thread.
#include "common.cuh"
#include "tma.cuh"

namespace h3dgs {

constexpr int kFwdBatch = 256;
constexpr int kFwdStages = 2;

template <bool HIER, bool DEPTH>
__global__ void __launch_bounds__(256)
render_forward_kernel(int W, int H, int gx, int shard_count, int shard_index, const uint2* __restrict__ ranges,
                      const Record* __restrict__ sorted, const float* __restrict__ bg, float* __restrict__ out_color,
                      float* __restrict__ out_invdepth, float* __restrict__ final_T, uint32_t* __restrict__ n_contrib,
                      uint32_t* __restrict__ tile_max_contrib)
{
    __shared__ __align__(128) Record s_rec[kFwdStages][kFwdBatch];
    __shared__ __align__(8) uint64_t s_full[kFwdStages];
    __shared__ uint32_t s_max;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile_x = blockIdx.x % gx;
    const int tile_y = (blockIdx.x / gx) * shard_count + shard_index;
    const int tile = tile_y * gx + tile_x;
    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + kFwdBatch - 1) / kFwdBatch;
    const Record* src = sorted + range.x;

    if (tid == 0) {
        for (int s = 0; s < kFwdStages; s++) mbar_init(&s_full[s], 1);
        s_max = 0;
        fence_mbar_init();
    }
    __syncthreads();
    int issued = 0;
    if (tid == 0) {
        for (int s = 0; s < kFwdStages && s < nb; s++) {
            const uint32_t bytes = (uint32_t)min(kFwdBatch, n - s * kFwdBatch) * (uint32_t)sizeof(Record);
            mbar_arrive_expect_tx(&s_full[s], bytes);
            tma_load_1d(&s_rec[s][0], src + (size_t)s * kFwdBatch, bytes, &s_full[s]);
        }
    }
    issued = min(kFwdStages, nb);

    const int px = tile_x * kTile + (tid & 15), py = tile_y * kTile + (tid >> 4);
    const bool inside = px < W && py < H;
    const float fpx = (float)px, fpy = (float)py;
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, invd = 0.f;
    uint32_t last = 0;

    int waited = 0;
    for (int b = 0; b < nb; b++) {
        const int st = b % kFwdStages;
        mbar_wait(&s_full[st], (uint32_t)((b / kFwdStages) & 1));
        waited = b + 1;
        const int cnt = min(kFwdBatch, n - b * kFwdBatch);
        // Per group of 32 entries each lane tests ONE entry against this warp's 16x2-pixel
        // strip and a ballot compacts the survivors, so culled entries cost nothing per pixel.
        // The survivor loop is warp-uniform (same mask in every lane) and its body is
        // straight-line + one short reconvergent `if`: a per-thread `continue`/`break` here
        // leaves the warp split into fragments that each re-walk the list (measured: 18x the
        // instructions).  A warp leaves the batch only when all of its 32 pixels are done.
        {
            const Record* rec = &s_rec[st][0];
            const uint32_t base = (uint32_t)(b * kFwdBatch);
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                if (__all_sync(0xffffffffu, done)) break;
                const int jl = j0 + lane;
                const bool hit = jl < cnt && strip_hit(__float_as_uint(rec[jl].b.w), warp);
                uint32_t m = __ballot_sync(0xffffffffu, hit);
                while (m) {
                    const int j = j0 + __ffs(m) - 1;
                    m &= m - 1;
                    const float4 a = rec[j].a;
                    const float4 bb = rec[j].b;
                    const float dx = a.x - fpx, dy = a.y - fpy;
                    const float power = -0.5f * (a.z * dx * dx + bb.x * dy * dy) - a.w * dx * dy;
                    float alpha = fminf(kAlphaCap, bb.y * fast_exp(power));
                    alpha = hier_alpha<HIER>(alpha, bb.z, __float_as_uint(bb.w));
                    const float test_T = T * (1.0f - alpha);
                    bool valid = !done && power <= 0.0f && alpha >= kAlphaSkip;
                    if (valid && test_T < kTStop) { done = true; valid = false; }
                    if (valid) {
                        const float4 c = rec[j].c;
                        const float w = alpha * T;
                        C0 += c.x * w; C1 += c.y * w; C2 += c.z * w;
                        if (DEPTH) invd += c.w * w;
                        T = test_T;
                        last = base + (uint32_t)j + 1u;
                    }
                }
            }
        }
        const int ndone = __syncthreads_count(done);
        if (ndone == 256) break;
        if (b + kFwdStages < nb) {
            if (tid == 0) {
                const int nb2 = b + kFwdStages;
                const uint32_t bytes = (uint32_t)min(kFwdBatch, n - nb2 * kFwdBatch) * (uint32_t)sizeof(Record);
                mbar_arrive_expect_tx(&s_full[st], bytes);
                tma_load_1d(&s_rec[st][0], src + (size_t)nb2 * kFwdBatch, bytes, &s_full[st]);
            }
            issued = b + kFwdStages + 1;
        }
    }
    // a CTA must not retire while a bulk copy into its shared memory is in flight
    if (tid == 0)
        for (int b = waited; b < issued; b++) mbar_wait(&s_full[b % kFwdStages], (uint32_t)((b / kFwdStages) & 1));

    if (inside) {
        const size_t pix = (size_t)py * W + px;
        final_T[pix] = T;
        n_contrib[pix] = last;
        if (shard_count > 1) {
            // packed shard layout [local tile row][channel][16][W]: one contiguous slab per
            // rank, so the image exchange is a single all-gather (h3dgs/dist.py)
            const size_t local_row = blockIdx.x / gx;
            const size_t o = ((local_row * 3) * kTile + (size_t)(py & (kTile - 1))) * W + px;
            const size_t cs = (size_t)kTile * W;
            out_color[o] = C0 + T * bg[0];
            out_color[o + cs] = C1 + T * bg[1];
            out_color[o + 2 * cs] = C2 + T * bg[2];
            if (DEPTH) out_invdepth[(local_row * kTile + (size_t)(py & (kTile - 1))) * W + px] = invd;
        } else {
            const size_t plane = (size_t)H * W;
            out_color[pix] = C0 + T * bg[0];
            out_color[plane + pix] = C1 + T * bg[1];
            out_color[2 * plane + pix] = C2 + T * bg[2];
            if (DEPTH) out_invdepth[pix] = invd;
        }
    }
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, last);
    if ((tid & 31) == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    if (tid == 0) tile_max_contrib[tile] = s_max;
}

int launch_render_forward(const h3dgs_raster_args& a, const uint32_t* ranges, const Record* sorted_records,
                          float* out_color, float* out_invdepth, float* final_T, uint32_t* n_contrib,
                          uint32_t* tile_max_contrib, cudaStream_t s)
{
    const int W = a.image_width, H = a.image_height;
    const int gx = (W + kTile - 1) / kTile, gy = (H + kTile - 1) / kTile;
    const int sc = a.shard_count > 0 ? a.shard_count : 1, si = a.shard_count > 0 ? a.shard_index : 0;
    const int rows = (gy + sc - 1 - si) / sc;
    if (rows <= 0 || gx <= 0) return H3DGS_OK;
    const bool hier = a.interpolation_weights != nullptr;
    const bool depth = a.do_depth != 0;
    const dim3 grid(gx * rows), block(256);
    ProfScope prof(H3DGS_STAGE_RENDER_FWD, s);
#define LAUNCH(HI, DE)                                                                                         \
    render_forward_kernel<HI, DE><<<grid, block, 0, s>>>(W, H, gx, sc, si, (const uint2*)ranges, sorted_records, \
                                                         a.bg, out_color, out_invdepth, final_T, n_contrib,      \
                                                         tile_max_contrib)
    if (hier) { if (depth) LAUNCH(true, true); else LAUNCH(true, false); }
    else      { if (depth) LAUNCH(false, true); else LAUNCH(false, false); }
#undef LAUNCH
    H3_LAUNCHED("render_forward", a.debug, s);
    return H3DGS_OK;
}

}  // namespace h3dgs
