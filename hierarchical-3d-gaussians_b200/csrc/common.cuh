// common.cuh -- shared definitions for the sm_100a kernels of libh3dgs.so.
// Constants are the published algorithm's (see oracle/oracle.c and DESIGN.md
// "recalled constants"; the hierarchy-rasterizer source is absent from /root/reference).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/h3dgs.h"

namespace h3dgs {

constexpr float kNearPlane = 0.2f;
constexpr float kFovClamp = 1.3f;
constexpr float kDilation = 0.3f;
constexpr float kLambdaFloor = 0.1f;
constexpr int kTile = H3DGS_TILE;
constexpr int kTilePixels = kTile * kTile;
constexpr float kAlphaCap = 0.99f;
constexpr float kAlphaSkip = 1.0f / 255.0f;
constexpr float kTStop = 0.0001f;
constexpr float kWEps = 0.0000001f;
constexpr uint32_t kKidsMask = 0xFFFFFu;
constexpr int kClampShift = 20, kQuadShift = 24;

// Per-Gaussian projected record: 3 x float4 = 48 B, 16-B aligned, so a batch of
// records is one contiguous cp.async.bulk (TMA) transfer.
//   a = {x, y, conic.x, conic.y}
//   b = {conic.z, opacity, t, kbits}     kbits: bits 0..19 num_node_kids, 20..22 SH clamp flags; in the
//                                        per-tile SORTED copy also bits 24..27 = mask of the tile's four 8x8-pixel
//                                        quadrants this entry can reach (quadrant culling, binning.cu)
//   c = {r, g, b, invdepth}
struct __align__(16) Record { float4 a, b, c; };
static_assert(sizeof(Record) == 48, "record must be 48 bytes");

__host__ __device__ inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- state buffer layouts (byte offsets inside the three alloc'd buffers) ----
struct GeomLayout {
    size_t depths, tiles_touched, offsets, records, scan_temp, total;
    size_t scan_temp_bytes;
};
struct BinLayout {
    size_t keys_unsorted, keys_sorted, vals_unsorted, vals_sorted, sort_temp, sorted_records, total;
    size_t sort_temp_bytes;
};
struct ImgLayout {
    size_t final_T, n_contrib, ranges, tile_max_contrib, tile_count, scan_info, total;
};
// written by tile_scan_kernel, read back by the host (the one num_rendered round trip)
struct ScanInfo { uint32_t D, max_count; };
// largest per-tile list the shared-memory sort handles; bigger lists fall back to the global CUB sort
constexpr int kTileSortCap = 8192;

GeomLayout geom_layout(int P);
BinLayout bin_layout(int64_t D);
ImgLayout img_layout(int W, int H);

// per-stage event timing (api.cu); no-ops unless h3dgs_profile_enable(1)
void prof_begin(int stage, cudaStream_t s);
void prof_end(int stage, cudaStream_t s);
struct ProfScope {
    int stage; cudaStream_t s;
    ProfScope(int st, cudaStream_t ss) : stage(st), s(ss) { prof_begin(stage, s); }
    ~ProfScope() { prof_end(stage, s); }
};

// 64-B pinned host scratch per device for the small D2H read-backs (num_rendered, cut size): a
// pageable destination makes cudaMemcpyAsync stage through the driver (api.cu)
int pinned_scratch(void** out);

// error plumbing (api.cu)
void set_error(const char* fmt, ...);
extern int64_t g_launches;

#define H3_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            h3dgs::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,          \
                             cudaGetErrorString(e__));                                     \
            return H3DGS_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

// after a kernel launch: count it, catch launch errors; in debug mode also sync
#define H3_LAUNCHED(name, debug, stream)                                                   \
    do {                                                                                   \
        h3dgs::g_launches++;                                                               \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ == cudaSuccess && (debug)) e__ = cudaStreamSynchronize(stream);            \
        if (e__ != cudaSuccess) {                                                          \
            h3dgs::set_error("kernel %s failed: %s", name, cudaGetErrorString(e__));       \
            return H3DGS_ECUDA;                                                            \
        }                                                                                  \
    } while (0)

// ---- stage entry points (one per .cu file) ----
int launch_preprocess(const h3dgs_raster_args& a, int32_t* radii, float* depths, uint32_t* tiles_touched,
                      Record* records, uint32_t* tile_count, cudaStream_t s);
int launch_tile_scan(const h3dgs_raster_args& a, const uint32_t* tile_count, uint32_t* ranges, ScanInfo* info,
                     cudaStream_t s);
int launch_tile_binning(const h3dgs_raster_args& a, const int32_t* radii, const float* depths, const Record* records,
                        int64_t D, uint32_t max_count, uint8_t* bin, const BinLayout& bl, const uint32_t* ranges,
                        uint32_t* tile_count, cudaStream_t s);
int launch_preprocess_color(const h3dgs_raster_args& a, const int32_t* radii, Record* records, cudaStream_t s);
int launch_sh_backward(const h3dgs_raster_args& a, const int32_t* radii, const Record* records, const float* accum,
                       float* dL_dmeans3D, float* dL_dsh, cudaStream_t s);
int launch_scan(const uint32_t* in, uint32_t* out, int n, void* temp, size_t temp_bytes, cudaStream_t s, bool debug);
size_t scan_temp_bytes(int n);
size_t sort_temp_bytes(int64_t n);
int launch_binning(const h3dgs_raster_args& a, const int32_t* radii, const float* depths, const uint32_t* offsets,
                   const Record* records, int64_t D, uint8_t* bin, const BinLayout& bl, uint32_t* ranges,
                   cudaStream_t s);
int launch_render_forward(const h3dgs_raster_args& a, const uint32_t* ranges, const Record* sorted_records,
                          float* out_color, float* out_invdepth, float* final_T, uint32_t* n_contrib,
                          uint32_t* tile_max_contrib, cudaStream_t s);
int launch_render_backward(const h3dgs_raster_args& a, const uint32_t* ranges, const Record* sorted_records,
                           const uint32_t* point_list, const float* final_T, const uint32_t* n_contrib,
                           const uint32_t* tile_max_contrib, const float* dL_dcolor, const float* dL_dinvdepth,
                           float* accum /*[P][10] zeroed*/, cudaStream_t s);
int launch_preprocess_backward(const h3dgs_raster_args& a, const int32_t* radii, const Record* records,
                               const float* accum, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh,
                               float* dL_dcolors, float* dL_dopacities, float* dL_dscales, float* dL_drots,
                               float* dL_dcov3D, cudaStream_t s);

// Hierarchy transition weight on the per-pixel blending weight (UNPINNED semantics,
// DESIGN.md "hierarchy alpha"): a' = t a + (1-t)(1 - (1-a)^(1/k)); identity for k<=1 or t>=1.
// One definition for forward and backward so both take identical skip decisions.
#ifdef __CUDACC__
// exp(x) for x <= 0 as one FMUL + MUFU.EX2 (ftz: results below 2^-126 flush to 0, far
// below the 1/255 alpha cut).  Shared by forward and backward so both see the same alpha.
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

// 1/x for x in [0.01, 2^20]: MUFU.RCP + one Newton step (no range/denormal handling needed
// here; ~1 ulp), 3 instructions instead of the ~10 of the IEEE-rounded __frcp_rn
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * (2.0f - x * r);
}
// log2(x) for normal x (MUFU.LG2 without the denormal pre-scaling of __log2f)
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Pixel layout of the blend kernels: CTA = 128 threads = 4 warps; warp q owns the 8x8 quadrant
// (q & 1, q >> 1) of the 16x16 tile; lane l owns column (l & 7) and the two rows 2*(l >> 3), +1.
__device__ __forceinline__ void quad_pixel(int tile_x, int tile_y, int warp, int lane, int& px, int& py0) {
    px = tile_x * kTile + 8 * (warp & 1) + (lane & 7);
    py0 = tile_y * kTile + 8 * (warp >> 1) + 2 * (lane >> 3);
}

template <bool HIER>
__device__ __forceinline__ void hier_alpha_grad(float a, float t, uint32_t kbits, float& alpha, float& dadb) {
    alpha = a; dadb = 1.0f;
    if (!HIER) return;
    const uint32_t k = kbits & kKidsMask;
    if (k <= 1u || t >= 1.0f) return;
    // 1 - (1-a)^(1/k) = -expm1(log1p(-a)/k).  Near the 1/255 skip threshold a is small and the
    // direct form cancels catastrophically (abs error ~2e-7 on a value ~4e-3 moves the skip
    // decision for 100x more pixels than in flat mode), so small a uses the two series
    // (relative error < 1e-7); larger a goes through MUFU.LG2 / MUFU.EX2.
    const float ik = fast_rcp((float)k);
    const float l2 = __log2f(1.0f - a);
    const float L = -a * (1.0f + a * (0.5f + a * (0.33333334f + a * (0.25f + a * 0.2f))));
    const float y = L * ik;
    const float omr_series = -y * (1.0f + y * (0.5f + y * (0.16666667f + y * 0.041666668f)));
    const float omr = a < 0.0625f ? omr_series : 1.0f - fast_exp2(l2 * ik);
    alpha = t * a + (1.0f - t) * omr;
    dadb = t + (1.0f - t) * ik * fast_exp2(l2 * (ik - 1.0f));
}
template <bool HIER>
__device__ __forceinline__ float hier_alpha(float a, float t, uint32_t kbits) {
    float alpha, dadb;
    hier_alpha_grad<HIER>(a, t, kbits, alpha, dadb);
    return alpha;
}

// ---- packed FP32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) -------------------------------------------
// A blend thread owns two vertically adjacent pixels; the blend kernels are issue-bound, and every
// per-pixel FP32 operation of the pair is ONE instruction on a packed register pair {lo = pixel 0,
// hi = pixel 1}.  Per-entry scalars enter as broadcast operands (pk(x, x) costs nothing: SASS
// `.F32` operand form), so nothing is spent on packing.  Each lane is an IEEE fma/mul/add.rn.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ float lo(f2 v) { float a, b; upk(v, a, b); return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; upk(v, a, b); return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// acc += a * b with the accumulator tied to its register pair (loop-carried sums: no copies)
__device__ __forceinline__ void fma2_acc(f2 a, f2 b, f2& acc) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ float hsum(f2 v) { float a, b; upk(v, a, b); return a + b; }
__device__ __forceinline__ f2 sel2(bool p0, bool p1, f2 x, f2 y) {      // per-lane p ? x : y
    float x0, x1, y0, y1; upk(x, x0, x1); upk(y, y0, y1);
    return pk(p0 ? x0 : y0, p1 ? x1 : y1);
}
__device__ __forceinline__ f2 ex2_2(f2 v) { float a, b; upk(v, a, b); return pk(fast_exp2(a), fast_exp2(b)); }

// Gaussian exponent of one entry at the thread's two pixels, shared by forward and backward so that
// both take identical decisions:  power_i = -1/2 (cx dx^2 + cz dy_i^2) - cy dx dy_i, evaluated as
// (C d_i + B) d_i + A with A = (cx dx)(-dx/2), B = -cy dx, C = -cz/2.
// nfpy = {-py0, -py1};  d = {a.y - py0, a.y - py1} is returned for the gradients.
__device__ __forceinline__ f2 pair_power(const float4& a, const float4& bb, float dx, f2 nfpy, f2& d) {
    d = add2(bc(a.y), nfpy);
    const float A = (a.z * dx) * (dx * -0.5f), B = -a.w * dx, C = -0.5f * bb.x;
    return fma2(fma2(bc(C), d, bc(B)), d, bc(A));
}
// G = exp(power) and the capped base alpha min(0.99, opacity G) of the pair
__device__ __forceinline__ void pair_gauss(f2 power, float opacity, f2& G, f2& abase) {
    G = ex2_2(mul2(power, bc(1.4426950408889634f)));
    float a0, a1; upk(mul2(bc(opacity), G), a0, a1);
    abase = pk(fminf(kAlphaCap, a0), fminf(kAlphaCap, a1));
}
// hierarchy transition weight on the pair (see hier_alpha_grad): k and t are per-entry, so the early
// out is warp-uniform.  GRAD = false drops the derivative.
template <bool HIER, bool GRAD>
__device__ __forceinline__ void pair_hier_alpha(f2 a, float t, uint32_t kbits, f2& alpha, f2& dadb) {
    alpha = a; dadb = bc(1.0f);
    if (!HIER) return;
    const uint32_t k = kbits & kKidsMask;
    if (k <= 1u || t >= 1.0f) return;
    const float ik = fast_rcp((float)k), u = 1.0f - t;
    float a0, a1; upk(a, a0, a1);
    float o0, o1; upk(sub2(bc(1.0f), a), o0, o1);
    const f2 l2 = pk(fast_log2(o0), fast_log2(o1));            // 1 - a is in [0.01, 1]
    // -log1p(-a) = a (1 + a/2 + a^2/3 + a^3/4 + a^4/5);  yn = -log1p(-a)/k >= 0
    f2 L = fma2(a, bc(0.2f), bc(0.25f));
    L = fma2(a, L, bc(0.33333334f));
    L = fma2(a, L, bc(0.5f));
    L = fma2(a, L, bc(1.0f));
    const f2 yn = mul2(mul2(a, L), bc(ik));
    // -expm1(-yn) = yn (1 - yn/2 + yn^2/6 - yn^3/24)
    f2 S = fma2(yn, bc(-0.041666668f), bc(0.16666667f));
    S = fma2(yn, S, bc(-0.5f));
    S = fma2(yn, S, bc(1.0f));
    const f2 omr_series = mul2(yn, S);
    const f2 omr_mufu = sub2(bc(1.0f), ex2_2(mul2(l2, bc(ik))));
    const f2 omr = sel2(a0 < 0.0625f, a1 < 0.0625f, omr_series, omr_mufu);
    alpha = fma2(bc(u), omr, mul2(bc(t), a));
    if (GRAD) dadb = fma2(bc(u * ik), ex2_2(mul2(l2, bc(ik - 1.0f))), bc(t));
}
#endif

// accum row layout (floats): 0,1 dmean2D.xy | 2,3,4 dconic | 5 dopacity | 6,7,8 dcolor | 9 dinvdepth
constexpr int kAccum = 10;

}  // namespace h3dgs
