// render_backward.cu -- K7: per-tile gradient replay (replaces BACKWARD::render).
// Semantics per oracle/oracle.c::oracle_render_backward.
//
// One CTA (128 threads, two pixels each) per tile, records streamed back-to-front with the same TMA
// double buffer as the forward; the arithmetic of a thread's two pixels runs on packed FP32x2
// instructions (one FFMA2 / FMUL2 / FADD2 serves both: the kernel is issue-bound); per-entry partials of a warp's 64 pixels are reduced with a
// transpose-reduce (12 shuffles for 10 values) and 10 lanes issue ONE red.global.add for the warp --
// 64x fewer atomics than the classic one-atomic-per-pixel formulation.
#include "common.cuh"
#include "tma.cuh"

namespace h3dgs {

// H3_BLEND_OCC8 (build switch, A/B on hardware): 224-entry batches (26.8 KB of shared memory per CTA) and a 64-register
// cap, so that 8 instead of 7 CTAs share an SM
#ifdef H3_BLEND_OCC8
constexpr int kBwdBatch = 224;
constexpr int kBwdMinBlocks = 8;
#else
constexpr int kBwdBatch = 256;
constexpr int kBwdMinBlocks = 1;
#endif
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
// two exchanges whose additions share one packed FADD2
__device__ __forceinline__ f2 xchg_add2(float keep0, float send0, float keep1, float send1, int mask) {
    return add2(pk(keep0, keep1), pk(__shfl_xor_sync(0xffffffffu, send0, mask), __shfl_xor_sync(0xffffffffu, send1, mask)));
}
__device__ __forceinline__ float transpose_reduce10(const float (&v)[10], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float a[5];
    upk(xchg_add2(b4 ? v[5] : v[0], b4 ? v[0] : v[5], b4 ? v[6] : v[1], b4 ? v[1] : v[6], 16), a[0], a[1]);
    upk(xchg_add2(b4 ? v[7] : v[2], b4 ? v[2] : v[7], b4 ? v[8] : v[3], b4 ? v[3] : v[8], 16), a[2], a[3]);
    a[4] = xchg_add(b4 ? v[9] : v[4], b4 ? v[4] : v[9], 16);
    // 5 -> (3 | 2)
    float b[3];
    upk(xchg_add2(b3 ? a[3] : a[0], b3 ? a[0] : a[3], b3 ? a[4] : a[1], b3 ? a[1] : a[4], 8), b[0], b[1]);
    b[2] = xchg_add(b3 ? 0.f : a[2], b3 ? a[2] : 0.f, 8);
    // 3 -> (2 | 1)
    float c[2];
    upk(xchg_add2(b2 ? b[2] : b[0], b2 ? b[0] : b[2], b2 ? 0.f : b[1], b2 ? b[1] : 0.f, 4), c[0], c[1]);
    // 2 -> (1 | 1)
    float d = xchg_add(b1 ? c[1] : c[0], b1 ? c[0] : c[1], 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// Group walk: the same reduction over the 8 lanes of a group (xor 4, 2, 1): 5 + 3 + 2 = 10 shuffles leave the 10 totals
// on 6 of the 8 lanes -- lanes with bit0 = 0 hold two (r0, r1), lanes with bit0 = 1 and bit1 = 0 hold one (r0).
// group_slots(lane & 7) gives the accum columns of (r0, r1), -1 = none.
__device__ __forceinline__ void group_slots(int k, int& s0, int& s1) {
    const int b2 = (k >> 2) & 1, b1 = (k >> 1) & 1, b0 = k & 1, base = 5 * b2;
    if (!b0) { s0 = base + (b1 ? 3 : 0); s1 = base + (b1 ? 4 : 1); }
    else { s0 = b1 ? -1 : base + 2; s1 = -1; }
}
__device__ __forceinline__ void transpose_reduce10_g8(const float (&v)[10], int lane, float& r0, float& r1) {
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    float a[5];
    upk(xchg_add2(b2 ? v[5] : v[0], b2 ? v[0] : v[5], b2 ? v[6] : v[1], b2 ? v[1] : v[6], 4), a[0], a[1]);
    upk(xchg_add2(b2 ? v[7] : v[2], b2 ? v[2] : v[7], b2 ? v[8] : v[3], b2 ? v[3] : v[8], 4), a[2], a[3]);
    a[4] = xchg_add(b2 ? v[9] : v[4], b2 ? v[4] : v[9], 4);
    float b[3];
    upk(xchg_add2(b1 ? a[3] : a[0], b1 ? a[0] : a[3], b1 ? a[4] : a[1], b1 ? a[1] : a[4], 2), b[0], b[1]);
    b[2] = xchg_add(b1 ? 0.f : a[2], b1 ? a[2] : 0.f, 2);
    upk(xchg_add2(b0 ? b[2] : b[0], b0 ? b[0] : b[2], b0 ? 0.f : b[1], b0 ? b[1] : 0.f, 1), r0, r1);
}

constexpr int kBwdThreads = 128;      // two vertically adjacent pixels per thread (see render_forward.cu)

// Tile-sharded frames (NCCL or peer mode): the sums of a rank's own tiles go into ITS accumulator; the exchange follows
// (peer mode: preprocess_backward.cu::peer_push_kernel stores the finished partial rows into the owners' staging areas).
// A first version added every (tile, Gaussian) row straight into the owner's memory with system-scope red.add: 4-byte
// reductions over NVLink made the replay 3.5x slower (2 GPUs: 1.28 vs 0.36 ms, profiles/r02_m2_*).
template <bool HIER, bool DEPTH, bool GROUPS>
__global__ void __launch_bounds__(kBwdThreads, kBwdMinBlocks)
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
    __shared__ uint8_t s_list[GROUPS ? kBwdThreads / 32 : 1][4][GROUPS ? kBwdBatch : 4];   // group walk: per warp, four lists of entry positions

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
    if (GROUPS) group_pixel(tile_x, tile_y, warp, lane, px, py0);
    else quad_pixel(tile_x, tile_y, warp, lane, px, py0);
    int gs0, gs1;
    group_slots(lane & 7, gs0, gs1);
    const int grp = lane >> 3;
    const int py1 = py0 + 1;
    const bool in0 = px < W && py0 < H, in1 = px < W && py1 < H;
    const float fpx = (float)px;
    const f2 nfpy = pk(-(float)py0, -(float)py1);
    const size_t pix0 = (size_t)py0 * W + px, pix1 = (size_t)py1 * W + px, plane = (size_t)H * W;
    const float Tf0 = in0 ? final_T[pix0] : 0.f, Tf1 = in1 ? final_T[pix1] : 0.f;
    const f2 Tf = pk(Tf0, Tf1);
    PairState ps = {Tf, bc(0.f)};
    const int last0 = in0 ? (int)n_contrib[pix0] : 0, last1 = in1 ? (int)n_contrib[pix1] : 0;
    float ga0 = 0.f, ga1 = 0.f, ga2 = 0.f, gad = 0.f, gb0 = 0.f, gb1 = 0.f, gb2 = 0.f, gbd = 0.f;
    if (in0) { ga0 = dL_dcolor[pix0]; ga1 = dL_dcolor[plane + pix0]; ga2 = dL_dcolor[2 * plane + pix0]; if (DEPTH) gad = dL_dinvdepth[pix0]; }
    if (in1) { gb0 = dL_dcolor[pix1]; gb1 = dL_dcolor[plane + pix1]; gb2 = dL_dcolor[2 * plane + pix1]; if (DEPTH) gbd = dL_dinvdepth[pix1]; }
    const f2 g0 = pk(ga0, gb0), g1 = pk(ga1, gb1), g2 = pk(ga2, gb2), gd = pk(gad, gbd);
    const f2 neg_bgd = pk(-(bg[0] * ga0 + bg[1] * ga1 + bg[2] * ga2), -(bg[0] * gb0 + bg[1] * gb1 + bg[2] * gb2));
    const int wlast = (int)__reduce_max_sync(0xffffffffu, (unsigned)max(last0, last1));   // nothing in this quadrant beyond it
    const int qsel = kBlockShift + 4 * warp;                // this warp's four block bits in the entries' reach mask
    // group walk: nothing of a group's 16 pixels lies beyond its own last contributor
    int gl0, gl1, gl2, gl3;
    {
        int gm = max(last0, last1);
        gm = max(gm, __shfl_xor_sync(0xffffffffu, gm, 4)); gm = max(gm, __shfl_xor_sync(0xffffffffu, gm, 2));
        gm = max(gm, __shfl_xor_sync(0xffffffffu, gm, 1));
        gl0 = __shfl_sync(0xffffffffu, gm, 0); gl1 = __shfl_sync(0xffffffffu, gm, 8);
        gl2 = __shfl_sync(0xffffffffu, gm, 16); gl3 = __shfl_sync(0xffffffffu, gm, 24);
    }

    for (int it = 0; it < nb; it++) {
        const int st = it % kBwdStages, b = nb - 1 - it;
        mbar_wait(&s_full[st], (uint32_t)((it / kBwdStages) & 1));
        const int cnt = min(kBwdBatch, n - b * kBwdBatch);
        const Record* rec = &s_rec[st][0];
        // one entry at the thread's two pixels; has = false: this lane's group has no entry in this iteration
        auto replay_entry = [&](int j, bool has) {
            const int e = b * kBwdBatch + j;              // 0-based list position; contributor number e+1
            const float4 a = rec[j].a;
            const float4 bb = rec[j].b;
            const uint32_t gid = s_id[st][j];             // loaded with the record: same uniform address arithmetic
            const uint32_t kb = __float_as_uint(bb.w);
            const float dx = a.x - fpx;
            // alpha of the two pixels with exactly the forward's arithmetic and decisions
            f2 d, G, al, dadb = bc(1.0f);                // dadb is only read with HIER
            const f2 pw = pair_power(a, bb, dx, nfpy, d);
            pair_gauss(pw, bb.y, G, al);
            float pw0, pw1, al0, al1;
            upk(pw, pw0, pw1); upk(al, al0, al1);
            // the hierarchy weight only lowers alpha (1 - (1-a)^(1/k) <= a), so an entry that no pixel of
            // the warp takes at its base alpha is skipped before that arithmetic
            bool v0 = has && e < last0 && pw0 <= 0.0f && al0 >= kAlphaSkip;
            bool v1 = has && e < last1 && pw1 <= 0.0f && al1 >= kAlphaSkip;
            if (lane == 0) H3_STAT(0, 1);
            if ((lane & 7) == 0 && has) H3_STAT(3, 1);
            if (!__any_sync(0xffffffffu, v0 || v1)) { if (lane == 0) H3_STAT(1, 1); return; }            // warp-uniform
            if (HIER) {
                pair_hier_alpha<HIER, true>(al, bb.z, kb & kSortedKidsMask, al, dadb);
                upk(al, al0, al1);
                v0 = v0 && al0 >= kAlphaSkip;
                v1 = v1 && al1 >= kAlphaSkip;
            }
            H3_STAT(2, (v0 ? 1 : 0) + (v1 ? 1 : 0));
#ifdef H3_SIMT_EMU
            { const uint32_t tk = __ballot_sync(0xffffffffu, v0 || v1);
              if (GROUPS && (lane & 7) == 0 && has && ((tk >> (8 * grp)) & 0xFFu) == 0u) H3_STAT(4, 1); }
#endif
            G = sel2(v0, v1, G, bc(0.f));
            al = sel2(v0, v1, al, bc(0.f));
            const float4 c = rec[j].c;
            f2 cg = fma2(bc(c.z), g2, fma2(bc(c.y), g1, mul2(bc(c.x), g0)));
            if (DEPTH) cg = fma2(bc(c.w), gd, cg);
            float v[10];
            pair_grad<HIER, DEPTH>(a, bb, dx, d, G, al, dadb, cg, Tf, neg_bgd, g0, g1, g2, gd, ps, v);
            if (GROUPS) {
                // every group reduces its own entry over its 8 lanes; groups without a taker stay silent
                const bool taker = ((__ballot_sync(0xffffffffu, v0 || v1) >> (8 * grp)) & 0xFFu) != 0u;
                float r0, r1;
                transpose_reduce10_g8(v, lane, r0, r1);
                float* row = accum + (size_t)gid * kAccum;
                if (taker && gs0 >= 0 && (DEPTH || gs0 < 9)) atomicAdd(row + gs0, r0);
                if (taker && gs1 >= 0 && (DEPTH || gs1 < 9)) atomicAdd(row + gs1, r1);
            } else {
                const float total = transpose_reduce10(v, lane);
                float* row = accum + (size_t)gid * kAccum;
                if (slot >= 0 && (DEPTH || slot < 9)) atomicAdd(row + slot, total);
            }
        };
        if (GROUPS) {
            // Group walk (see render_forward.cu): per batch the warp compacts four lists of entry positions -- an entry is
            // listed for a group when its block bit is set and it lies before the group's last contributor -- and the
            // groups replay their lists back to front in lockstep.
            uint8_t* lst = &s_list[warp][0][0];
            const uint32_t lt = (1u << lane) - 1u;
            int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int jl = j0 + lane, e = b * kBwdBatch + jl;
                uint32_t nib = jl < cnt ? (__float_as_uint(rec[jl].b.w) >> qsel) & 0xFu : 0u;
                nib &= (e < gl0 ? 1u : 0u) | (e < gl1 ? 2u : 0u) | (e < gl2 ? 4u : 0u) | (e < gl3 ? 8u : 0u);
                const uint32_t m0 = __ballot_sync(0xffffffffu, nib & 1u), m1 = __ballot_sync(0xffffffffu, nib & 2u);
                const uint32_t m2 = __ballot_sync(0xffffffffu, nib & 4u), m3 = __ballot_sync(0xffffffffu, nib & 8u);
                if (nib & 1u) lst[c0 + __popc(m0 & lt)] = (uint8_t)jl;
                if (nib & 2u) lst[kBwdBatch + c1 + __popc(m1 & lt)] = (uint8_t)jl;
                if (nib & 4u) lst[2 * kBwdBatch + c2 + __popc(m2 & lt)] = (uint8_t)jl;
                if (nib & 8u) lst[3 * kBwdBatch + c3 + __popc(m3 & lt)] = (uint8_t)jl;
                c0 += __popc(m0); c1 += __popc(m1); c2 += __popc(m2); c3 += __popc(m3);
            }
            __syncwarp();
            const int mylen = grp == 0 ? c0 : grp == 1 ? c1 : grp == 2 ? c2 : c3;
            const int maxlen = max(max(c0, c1), max(c2, c3));
            const uint8_t* my = lst + grp * kBwdBatch;
            for (int i = maxlen - 1; i >= 0; i--) {
                const bool has = i < mylen;
                replay_entry(has ? (int)my[i] : 0, has);
            }
            __syncwarp();                     // the lists are rebuilt for the next batch
        } else {
            // back to front; per round of 32 entries a ballot compacts the entries that can reach this warp's quadrant at all
            for (int j0 = (cnt - 1) & ~31; j0 >= 0; j0 -= 32) {
                const int jl = j0 + lane;
                const uint32_t nib = (jl < cnt && (b * kBwdBatch + jl) < wlast) ? (__float_as_uint(rec[jl].b.w) >> qsel) & 0xFu : 0u;
                uint32_t m = __ballot_sync(0xffffffffu, nib != 0u);
                while (m != 0u) {
                    const int top = 31 - __clz(m);
                    m &= ~(1u << top);
                    replay_entry(j0 + top, true);
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
    const dim3 grid(gx * rows), block(kBwdThreads);
    ProfScope prof(H3DGS_STAGE_RENDER_BWD, s);
    const bool groups = use_group_walk();
#define LAUNCH(HI, DE, GR)                                                                                          \
    render_backward_kernel<HI, DE, GR><<<grid, block, 0, s>>>(W, H, gx, sc, si, (const uint2*)ranges, sorted_records, \
                                                              point_list, a.bg, final_T, n_contrib, tile_max_contrib, \
                                                              dL_dcolor, dL_dinvdepth, accum)
#define LAUNCH2(HI, DE) do { if (groups) LAUNCH(HI, DE, true); else LAUNCH(HI, DE, false); } while (0)
    if (hier) { if (depth) LAUNCH2(true, true); else LAUNCH2(true, false); }
    else      { if (depth) LAUNCH2(false, true); else LAUNCH2(false, false); }
#undef LAUNCH2
#undef LAUNCH
    H3_LAUNCHED("render_backward", a.debug, s);
    return H3DGS_OK;
}

}  // namespace h3dgs
