// render_backward.cu -- K7: per-tile gradient replay (replaces BACKWARD::render).
// Semantics per oracle/oracle.c::oracle_render_backward.
//
// One CTA (128 threads, two pixels each) per tile, records streamed back-to-front with the same TMA
// double buffer as the forward; per-entry partials of a warp's 64 pixels are reduced with a
// transpose-reduce (12 shuffles for 10 values) and 10 lanes issue ONE red.global.add for the warp --
// 64x fewer atomics than the classic one-atomic-per-pixel formulation.
#include "common.cuh"
#include "tma.cuh"

namespace h3dgs {

constexpr int kBwdBatch = 256;
constexpr int kBwdStages = 2;

// Reduce NV per-lane values over the 32 lanes of a warp with a transpose-reduce: at every
// butterfly level each lane keeps half of its values and ships the other half, so the whole
// reduction costs 5+3+2+1+1 = 12 shuffles for 10 values instead of 10 x 5 = 50, and the 10
// totals end up on 10 DIFFERENT lanes -- which then issue ONE predicated red.global.add for
// the warp (a contiguous 40-B row) instead of 10 serial atomics from lane 0.
// Returns this lane's total; `slot` (precomputed per lane by reduce_slot) says which value it is.
__device__ __forceinline__ int reduce_slot(int lane) {
    if (lane & 1) return -1;
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1;
    int ai;                                  // index within the 5 values kept after level 16
    if (!b3) { if (b2 && b1) return -1; ai = b2 ? 2 : b1; }
    else     { if (b2) return -1; ai = 3 + b1; }
    return 5 * b4 + ai;
}
__device__ __forceinline__ float xchg_add(float keep, float send, int mask) {
    return keep + __shfl_xor_sync(0xffffffffu, send, mask);
}
__device__ __forceinline__ float transpose_reduce10(const float (&v)[10], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float a[5];
#pragma unroll
    for (int k = 0; k < 5; k++) a[k] = xchg_add(b4 ? v[k + 5] : v[k], b4 ? v[k] : v[k + 5], 16);
    // 5 -> (3 | 2)
    float b[3];
    b[0] = xchg_add(b3 ? a[3] : a[0], b3 ? a[0] : a[3], 8);
    b[1] = xchg_add(b3 ? a[4] : a[1], b3 ? a[1] : a[4], 8);
    b[2] = xchg_add(b3 ? 0.f : a[2], b3 ? a[2] : 0.f, 8);
    // 3 -> (2 | 1)
    float c[2];
    c[0] = xchg_add(b2 ? b[2] : b[0], b2 ? b[0] : b[2], 4);
    c[1] = xchg_add(b2 ? 0.f : b[1], b2 ? b[1] : 0.f, 4);
    // 2 -> (1 | 1)
    float d = xchg_add(b1 ? c[1] : c[0], b1 ? c[0] : c[1], 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

constexpr int kBwdThreads = 128;      // two vertically adjacent pixels per thread (see render_forward.cu)

// per-pixel replay state
struct PixState {
    float T, acc_s, last_cg, last_alpha;   // transmittance, (accum_rec . g), (last colour . g), last alpha
};

// alpha of one pixel for entry (a, bb), with exactly the forward's decisions
template <bool HIER>
__device__ __forceinline__ bool pixel_alpha(const float4& a, const float4& bb, uint32_t kb, float dx, float dy, bool in_list,
                                            float& G, float& alpha, float& dadb)
{
    G = 0.f; alpha = 0.f; dadb = 1.f;
    bool valid = false;
    const float power = -0.5f * (a.z * dx * dx + bb.x * dy * dy) - a.w * dx * dy;
    if (in_list && power <= 0.0f) {
        G = fast_exp(power);
        const float abase = fminf(kAlphaCap, bb.y * G);
        hier_alpha_grad<HIER>(abase, bb.z, kb, alpha, dadb);
        valid = alpha >= kAlphaSkip;
    }
    if (!valid) { G = 0.f; alpha = 0.f; }
    return valid;
}

// One pixel's contribution of entry (a, bb, c) to the 10 per-Gaussian sums; branch-free so that an
// invalid pixel (G = alpha = 0) adds exact zeros.
template <bool DEPTH>
__device__ __forceinline__ void pixel_grad(const float4& a, const float4& bb, float dx, float dy, bool valid, float G,
                                           float alpha, float dadb, float cg, float T_final, float bg_dot, float g0,
                                           float g1, float g2, float gd, PixState& st, float (&v)[10])
{
    const float rcp = fast_rcp(1.f - alpha);                   // one reciprocal serves T and the bg term
    const float Tn = st.T * rcp;
    const float as_n = st.last_alpha * st.last_cg + (1.f - st.last_alpha) * st.acc_s;
    const float w = valid ? alpha * Tn : 0.f;                 // dchannel_dcolor
    const float dL_dalpha = (cg - as_n) * Tn - (T_final * rcp) * bg_dot;
    const float dL_dab = valid ? dL_dalpha * dadb : 0.f;
    if (valid) { st.T = Tn; st.acc_s = as_n; st.last_cg = cg; st.last_alpha = alpha; }
    const float dL_dG = bb.y * dL_dab;
    const float gdx = G * dx, gdy = G * dy;
    // constant factors (0.5 W, 0.5 H, -0.5) are applied once per Gaussian in preprocess_backward
    v[0] += dL_dG * (-gdx * a.z - gdy * a.w);
    v[1] += dL_dG * (-gdy * bb.x - gdx * a.w);
    v[2] += gdx * dx * dL_dG;
    v[3] += gdx * dy * dL_dG;
    v[4] += gdy * dy * dL_dG;
    v[5] += G * dL_dab;
    v[6] += w * g0; v[7] += w * g1; v[8] += w * g2;
    if (DEPTH) v[9] += w * gd;
}

template <bool HIER, bool DEPTH>
__global__ void __launch_bounds__(kBwdThreads)
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

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = reduce_slot(lane);
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
        for (int k = tid; k < kBwdBatch; k += kBwdThreads) {
            const int e = b * kBwdBatch + k;
            if (e < n) s_id[st][k] = ids[e];
        }
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

    int px, py0;
    quad_pixel(tile_x, tile_y, warp, lane, px, py0);
    const int py1 = py0 + 1;
    const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
    const float fpx = (float)px, fpy0 = (float)py0, fpy1 = (float)py1;
    const size_t pix0 = (size_t)py0 * W + px, pix1 = (size_t)py1 * W + px, plane = (size_t)H * W;
    const float Tf0 = in0 ? final_T[pix0] : 0.f, Tf1 = in1 ? final_T[pix1] : 0.f;
    PixState st0 = {Tf0, 0.f, 0.f, 0.f}, st1 = {Tf1, 0.f, 0.f, 0.f};
    const int last0 = in0 ? (int)n_contrib[pix0] : 0, last1 = in1 ? (int)n_contrib[pix1] : 0;
    float ga0 = 0.f, ga1 = 0.f, ga2 = 0.f, gad = 0.f, gb0 = 0.f, gb1 = 0.f, gb2 = 0.f, gbd = 0.f;
    if (in0) { ga0 = dL_dcolor[pix0]; ga1 = dL_dcolor[plane + pix0]; ga2 = dL_dcolor[2 * plane + pix0]; if (DEPTH) gad = dL_dinvdepth[pix0]; }
    if (in1) { gb0 = dL_dcolor[pix1]; gb1 = dL_dcolor[plane + pix1]; gb2 = dL_dcolor[2 * plane + pix1]; if (DEPTH) gbd = dL_dinvdepth[pix1]; }
    const float bgd0 = bg[0] * ga0 + bg[1] * ga1 + bg[2] * ga2, bgd1 = bg[0] * gb0 + bg[1] * gb1 + bg[2] * gb2;
    const int wlast = (int)__reduce_max_sync(0xffffffffu, (unsigned)max(last0, last1));   // nothing in this quadrant beyond it
    const uint32_t qbit = 1u << (kQuadShift + warp);

    for (int it = 0; it < nb; it++) {
        const int st = it % kBwdStages, b = nb - 1 - it;
        mbar_wait(&s_full[st], (uint32_t)((it / kBwdStages) & 1));
        const int cnt = min(kBwdBatch, n - b * kBwdBatch);
        const Record* rec = &s_rec[st][0];
        // back to front; per group of 32 entries a ballot compacts the entries that can reach
        // this warp's quadrant at all (see render_forward.cu)
        for (int j0 = (cnt - 1) & ~31; j0 >= 0; j0 -= 32) {
            const int jl = j0 + lane;
            const bool hit = jl < cnt && (b * kBwdBatch + jl) < wlast && (__float_as_uint(rec[jl].b.w) & qbit) != 0u;
            uint32_t m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int top = 31 - __clz(m);
                m &= ~(1u << top);
                const int j = j0 + top;
                const int e = b * kBwdBatch + j;              // 0-based list position; contributor number e+1
                const float4 a = rec[j].a;
                const float4 bb = rec[j].b;
                const uint32_t kb = __float_as_uint(bb.w);
                const float dx = a.x - fpx, dy0 = a.y - fpy0, dy1 = a.y - fpy1;
                float G0, al0, dd0, G1, al1, dd1;
                const bool v0 = pixel_alpha<HIER>(a, bb, kb, dx, dy0, e < last0, G0, al0, dd0);
                const bool v1 = pixel_alpha<HIER>(a, bb, kb, dx, dy1, e < last1, G1, al1, dd1);
                if (!__any_sync(0xffffffffu, v0 || v1)) continue;            // warp-uniform
                const float4 c = rec[j].c;
                float cg0 = c.x * ga0 + c.y * ga1 + c.z * ga2, cg1 = c.x * gb0 + c.y * gb1 + c.z * gb2;
                if (DEPTH) { cg0 += c.w * gad; cg1 += c.w * gbd; }
                float v[10];
#pragma unroll
                for (int k = 0; k < 10; k++) v[k] = 0.f;
                pixel_grad<DEPTH>(a, bb, dx, dy0, v0, G0, al0, dd0, cg0, Tf0, bgd0, ga0, ga1, ga2, gad, st0, v);
                pixel_grad<DEPTH>(a, bb, dx, dy1, v1, G1, al1, dd1, cg1, Tf1, bgd1, gb0, gb1, gb2, gbd, st1, v);
                const float total = transpose_reduce10(v, lane);
                if (slot >= 0 && (DEPTH || slot < 9))
                    atomicAdd(accum + (size_t)s_id[st][j] * kAccum + slot, total);
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
    const dim3 grid(gx * rows), block(kBwdThreads);
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
