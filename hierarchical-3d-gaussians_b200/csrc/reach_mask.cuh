// reach_mask.cuh -- which 4x4-pixel blocks of a tile an entry can reach (used by the record gather of
// binning.cu; the blend kernels skip everything else).  Also compiled for the host by tests/emul/reach_mask_test.cpp,
// which checks it against a brute-force per-pixel evaluation.
#pragma once
#ifdef H3_REACH_HOST_TEST
#include <math.h>
#include <stdint.h>
#include <algorithm>
#define __device__
#define __forceinline__ inline
struct float4 { float x, y, z, w; };
#define H3_FAST_LOGF(x) logf(x)
using std::max; using std::min;
namespace h3dgs { constexpr int kTile = 16; }
#else
#define H3_FAST_LOGF(x) __logf(x)
#endif

namespace h3dgs {

// Which of the sixteen 4x4-pixel blocks of tile (tile_x, tile_y) can entry (a, b) reach at all?
// Bit 4 q + g: quadrant q = (qx, qy) of the tile (the warp of the blend CTAs), block g = (bx, by) inside it
// (an 8-lane group of that warp); the quadrant is reached iff any of its four bits is set.
// alpha >= 1/255  <=>  q(e) = A ex^2 + 2 B ex ey + C ey^2 <= 2 ln(255 o) =: bound, e = pixel - mean.
// Scanline form: inside a band of rows [y0, y0+3] the ellipse {q <= bound} projects onto the x interval
// [lo, hi]; hi is attained at the band's row nearest to the ellipse's rightmost point (ey = -B sqrt(bound/(det C))),
// lo at the row nearest to the leftmost point (-ey): two square roots per band.  The ellipse cut by the band is
// convex, so a block column [4 bx, 4 bx + 3] is reached iff it overlaps [lo, hi] -- the same (exact, over the
// rectangle of pixel centres) criterion as a per-block minimisation of q, at a fraction of the instructions.
// Margins cover fp32 rounding; skipping a block never changes a result.
__device__ __forceinline__ uint32_t block_mask16(const float4& a, const float4& b, int tile_x, int tile_y)
{
    const float A = a.z, B = a.w, C = b.x;
    const float det = A * C - B * B;
    const float o255 = b.y * 255.0f;
    if (!(o255 > 1.0f)) return 0u;                     // can never reach alpha >= 1/255
    const float bound = 2.0f * H3_FAST_LOGF(o255) * 1.002f + 1e-3f;
    const float rx = a.x - (float)(tile_x * kTile), ry = a.y - (float)(tile_y * kTile);
    if (!(det > 0.0f && A > 0.0f && C > 0.0f) || !(bound == bound && rx == rx && ry == ry)) return 0xFFFFu;
    const float invA = 1.0f / A, BoA = B * invA, Ab = A * bound;
    const float er = -B * sqrtf(bound / (det * C));     // row offset (from the mean) of the rightmost point; leftmost: -er
    if (!(er == er) || fabsf(er) > 1e6f || !(Ab < 3.0e37f)) return 0xFFFFu;
    uint32_t rows = 0u;                                  // row-major: bit 4 by + bx
#pragma unroll
    for (int by = 0; by < 4; by++) {
        const float y0 = (float)(4 * by) - ry, y1 = y0 + 3.0f;          // the band's rows relative to the mean
        const float e_hi = fminf(y1, fmaxf(y0, er)), e_lo = fminf(y1, fmaxf(y0, -er));
        const float d_hi = Ab - det * e_hi * e_hi, d_lo = Ab - det * e_lo * e_lo;
        if (fmaxf(d_hi, d_lo) < -1e-3f * Ab) continue;                  // the band misses the ellipse
        const float hi = rx - BoA * e_hi + (sqrtf(fmaxf(d_hi, 0.0f)) * invA * 1.001f + 2e-3f);
        const float lo = rx - BoA * e_lo - (sqrtf(fmaxf(d_lo, 0.0f)) * invA * 1.001f + 2e-3f);
        // block bx covers pixel centres 4 bx .. 4 bx + 3 in tile coordinates
        const int b0 = max(0, (int)ceilf((fmaxf(lo, -8.0f) - 3.0f) * 0.25f)), b1 = min(3, (int)floorf(fminf(hi, 24.0f) * 0.25f));
        if (b0 <= b1) rows |= ((2u << b1) - (1u << b0)) << (4 * by);
    }
    // row-major -> quadrant-major: per byte (two block rows) the bit pairs {0,1},{4,5} | {2,3},{6,7} form the nibbles
    const uint32_t m = (rows & 0x0303u) | ((rows >> 2) & 0x0C0Cu) | ((rows << 2) & 0x3030u) | (rows & 0xC0C0u);
    return m;
}

}  // namespace h3dgs
