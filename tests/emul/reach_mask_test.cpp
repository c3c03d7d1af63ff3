// CPU check of csrc/reach_mask.cuh::block_mask16 (test infrastructure only): for random projected Gaussians and tile
// positions, every pixel of the tile whose alpha = min(0.99, o exp(-q/2)) reaches 1/255 must lie in a block the mask
// marks (skipping an unmarked block must never change a result), and the mask must be tight: marked blocks the
// ellipse does not touch at all (continuous rectangle test on a 16x oversampled grid) stay below a small fraction.
#define H3_REACH_HOST_TEST
#include "../../hierarchical-3d-gaussians_b200/csrc/reach_mask.cuh"
#include <stdio.h>
#include <stdlib.h>
using namespace h3dgs;

static double urand() { return (double)rand() / RAND_MAX; }

int main() {
    srand(7);
    long entries = 0, missed = 0, marked = 0, marked_empty = 0, needed = 0;
    for (int it = 0; it < 200000; it++) {
        // conic from a random covariance (sigma 0.3 .. 12 px, any orientation, +0.3 dilation like K1)
        const double s1 = 0.3 * pow(40.0, urand()), s2 = 0.3 * pow(40.0, urand()), th = urand() * 3.14159265;
        const double c = cos(th), s = sin(th);
        const double cxx = c * c * s1 * s1 + s * s * s2 * s2 + 0.3, cxy = c * s * (s1 * s1 - s2 * s2), cyy = s * s * s1 * s1 + c * c * s2 * s2 + 0.3;
        const double det = cxx * cyy - cxy * cxy;
        float4 a, b;
        a.z = (float)(cyy / det); a.w = (float)(-cxy / det); b.x = (float)(cxx / det);
        b.y = (float)(it % 7 == 0 ? 1.6 * urand() : pow(10.0, -2.6 * urand()));        // opacity (may exceed 1 in hierarchy mode)
        const int tx = rand() % 120, ty = rand() % 68;
        a.x = (float)(tx * 16 + (urand() * 60.0 - 22.0)); a.y = (float)(ty * 16 + (urand() * 60.0 - 22.0));
        b.z = 1.0f; b.w = 0.0f;
        const uint32_t m = block_mask16(a, b, tx, ty);
        entries++;
        uint32_t need = 0, touch = 0;
        for (int py = 0; py < 16; py++)
            for (int px = 0; px < 16; px++) {
                const float dx = a.x - (float)(tx * 16 + px), dy = a.y - (float)(ty * 16 + py);
                const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                const int bx = px >> 2, by = py >> 2;
                const int bit = 4 * ((bx >> 1) | ((by >> 1) << 1)) + ((bx & 1) | ((by & 1) << 1));
                if (power <= 0.0f && fminf(0.99f, b.y * expf(power)) >= 1.0f / 255.0f) need |= 1u << bit;
            }
        // continuous test, oversampled: does {q <= 2 ln(255 o)} touch the block's rectangle of pixel centres?
        const double bound = 2.0 * log(255.0 * (double)b.y);
        if (bound > 0)
            for (int by = 0; by < 4; by++)
                for (int bx = 0; bx < 4; bx++) {
                    bool hit = false;
                    for (int i = 0; i <= 48 && !hit; i++)
                        for (int j = 0; j <= 48 && !hit; j++) {
                            const double ex = tx * 16 + 4 * bx + i / 16.0 - a.x, ey = ty * 16 + 4 * by + j / 16.0 - a.y;
                            if (a.z * ex * ex + 2.0 * a.w * ex * ey + b.x * ey * ey <= bound) hit = true;
                        }
                    if (hit) touch |= 1u << (4 * ((bx >> 1) | ((by >> 1) << 1)) + ((bx & 1) | ((by & 1) << 1)));
                }
        if (need & ~m) { missed++; if (missed < 5) printf("MISSED it=%d need=%04x mask=%04x\n", it, need, m); }
        marked += __builtin_popcount(m); needed += __builtin_popcount(need);
        marked_empty += __builtin_popcount(m & ~touch);
    }
    printf("entries %ld missed %ld marked %ld needed %ld marked_untouched %ld\n", entries, missed, marked, needed, marked_empty);
    if (missed) return 1;
    if ((double)marked_empty > 0.02 * (double)marked) { printf("mask not tight\n"); return 2; }
    return 0;
}
