// Host build of csrc/cov_grad.cuh (the K8 + K9c arithmetic of preprocess_backward.cu) behind a C entry
// point, so the CPU suite can run it on whole scenes and compare with the double-precision oracle.
// Test infrastructure only.
#define H3_HOST_EMU
#include "../../hierarchical-3d-gaussians_b200/csrc/cov_grad.cuh"
using namespace h3dgs;

extern "C" void emu_cov_chain(int P, const float* view, const float* means, float fx, float fy, float tanx, float tany,
                              const float* scales, float scale_mod, const float* rots, const float* cov3D_precomp,
                              const float* dconic, const float* dinvdepth, const int* radii,
                              float* dmean, float* g6, float* dscale, float* dq)
{
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        float cov6[6], R[3][3] = {{0}}, Mm[3][3] = {{0}}, s[3] = {0, 0, 0};
        const bool have_sr = cov3D_precomp == nullptr;
        if (have_sr) {
            for (int k = 0; k < 3; k++) s[k] = scale_mod * scales[3 * i + k];
            cov3d_from_scale_quat(s, rots + 4 * i, R, Mm, cov6);
        } else {
            for (int k = 0; k < 6; k++) cov6[k] = cov3D_precomp[6 * i + k];
        }
        CovGradOut o = {};
        cov_chain_backward(view, means[3 * i], means[3 * i + 1], means[3 * i + 2], fx, fy, tanx, tany, 1.3f, 0.3f, cov6,
                           dconic[3 * i], dconic[3 * i + 1], dconic[3 * i + 2], dinvdepth ? dinvdepth[i] : 0.f,
                           dinvdepth != nullptr, have_sr, s, have_sr ? rots + 4 * i : nullptr, R, Mm, scale_mod, o);
        for (int k = 0; k < 3; k++) dmean[3 * i + k] = o.dmean[k];
        for (int k = 0; k < 6; k++) g6[6 * i + k] = o.g6[k];
        if (have_sr) {
            for (int k = 0; k < 3; k++) dscale[3 * i + k] = o.dscale[k];
            for (int k = 0; k < 4; k++) dq[4 * i + k] = o.dq[k];
        }
    }
}
