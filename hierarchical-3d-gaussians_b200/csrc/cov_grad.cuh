// cov_grad.cuh -- K8 + K9c for one Gaussian: the gradient of the 2D conic back to the 3D covariance,
// the view-space mean (through the projection Jacobian), the scale and the rotation.
//
// Precision.  The published formulas are evaluated in fp32 EXCEPT one block, which is fp64:
//     A V (2x3),  a, b, c = A V A^T + dilation,  denom = a c - b^2,  1 / (denom^2 + 1e-7),  dL/da, dL/db, dL/dc
// i.e. the 2x2 screen covariance and the gradient of its inverse -- sums of terms several orders of
// magnitude larger than the result (~50 fp64 operations).  Measured with the host build of this header on
// the test scenes against the double-precision oracle (norm-wise, tests/test_cov_grad_cpu.py): all-fp32
// misses the 1e-5 parity bar on the rotation gradient (1.5e-5); with only denom .. dL/dc in fp64 the
// outputs are still off by up to 7e-6 (covariance) / 5e-6 (rotation); with the whole block in fp64
// every output is within 2e-6, as good as the whole chain in fp64 -- which is what the kernel did before,
// at 156 registers and ~400 DFMA per Gaussian.
//
// Compiles for the device (included from preprocess_backward.cu) and, with -DH3_HOST_EMU, for the host:
// tests/emul/cov_grad_emul.cpp exposes it to the CPU suite, which checks it against the oracle.
#pragma once
#include <math.h>

#ifdef H3_HOST_EMU
#define H3_CG_FN static inline
#else
#define H3_CG_FN __device__ __forceinline__
#endif

namespace h3dgs {

struct CovGradOut {
    float dmean[3];     // K8 contribution to dL/dmean3D (through the Jacobian and the depth output)
    float g6[6];        // dL/dcov3D, order xx,xy,xz,yy,yz,zz (off-diagonals counted for both uses)
    float dscale[3];    // valid when scale/rotation were given
    float dq[4];        // w, x, y, z
};

// 3D covariance from (scale * mod, quaternion): Sigma = M^T M with M[k][j] = s_k R[j][k]
H3_CG_FN void cov3d_from_scale_quat(const float s[3], const float q[4], float R[3][3], float Mm[3][3], float cov6[6]) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0][0] = 1.f - 2.f * (y * y + z * z); R[0][1] = 2.f * (x * y - r * z); R[0][2] = 2.f * (x * z + r * y);
    R[1][0] = 2.f * (x * y + r * z); R[1][1] = 1.f - 2.f * (x * x + z * z); R[1][2] = 2.f * (y * z - r * x);
    R[2][0] = 2.f * (x * z - r * y); R[2][1] = 2.f * (y * z + r * x); R[2][2] = 1.f - 2.f * (x * x + y * y);
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < 3; j++) Mm[k][j] = s[k] * R[j][k];
    int o = 0;
    for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++) cov6[o++] = Mm[0][a] * Mm[0][b] + Mm[1][a] * Mm[1][b] + Mm[2][a] * Mm[2][b];
}

// v: world->view matrix, transposed storage v[4c + k] (scene/cameras.py:95).  (mx,my,mz): the (lerped) mean.
// dcx,dcy,dcz: dL/dconic with the -1/2 of the quadratic form already applied.  g_iv: dL/d(inverse depth).
// have_sr: scale/rotation inputs (then s = scale * mod, q, R, Mm as produced by cov3d_from_scale_quat).
H3_CG_FN void cov_chain_backward(const float* v, float mx, float my, float mz, float fx, float fy, float tanx, float tany,
                                 float fov_clamp, float dilation, const float cov6[6], float dcx, float dcy, float dcz,
                                 float g_iv, bool use_depth, bool have_sr, const float s[3], const float q[4],
                                 const float R[3][3], const float Mm[3][3], float scale_mod, CovGradOut& out)
{
    float tx = v[0] * mx + v[4] * my + v[8] * mz + v[12];
    float ty = v[1] * mx + v[5] * my + v[9] * mz + v[13];
    const float tz = v[2] * mx + v[6] * my + v[10] * mz + v[14];
    const float limx = fov_clamp * tanx, limy = fov_clamp * tany;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
    const float J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
    float A[2][3];
    for (int c = 0; c < 3; c++) {
        A[0][c] = J00 * v[4 * c + 0] + J02 * v[4 * c + 2];
        A[1][c] = J11 * v[4 * c + 1] + J12 * v[4 * c + 2];
    }
    const float V[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
    // ---- fp64 from here to dL/d(a,b,c): cov2D = A V A^T + dilation, accumulated in double from the fp32
    // factors (a, b, c are sums of terms ~1e3..1e6 whose differences matter below) ----
    double AVd[2][3];
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) AVd[r][c] = (double)A[r][0] * V[0][c] + (double)A[r][1] * V[1][c] + (double)A[r][2] * V[2][c];
    const double ad = (AVd[0][0] * A[0][0] + AVd[0][1] * A[0][1] + AVd[0][2] * A[0][2]) + dilation;
    const double bd = AVd[0][0] * A[1][0] + AVd[0][1] * A[1][1] + AVd[0][2] * A[1][2];
    const double cd = (AVd[1][0] * A[1][0] + AVd[1][1] * A[1][1] + AVd[1][2] * A[1][2]) + dilation;
    float AV[2][3];
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) AV[r][c] = (float)AVd[r][c];

    // conic = (c, -b, a) / denom: gradient w.r.t. (a, b, c)
    float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
    {
        const double denom = ad * cd - bd * bd;
        const double denom2inv = 1.0 / ((denom * denom) + 0.0000001);
        if (denom2inv != 0.) {
            const double x = dcx, y = dcy, z = dcz;
            dL_da = (float)(denom2inv * (-cd * cd * x + 2 * bd * cd * y + (denom - ad * cd) * z));
            dL_dc = (float)(denom2inv * (-ad * ad * z + 2 * ad * bd * y + (denom - ad * cd) * x));
            dL_db = (float)(denom2inv * 2 * (bd * cd * x - (denom + 2 * bd * bd) * y + ad * bd * z));
        }
    }
    // ---- fp32 again.  cov2D = A V A^T : d/dV (off-diagonals appear twice in the symmetric V) ----
    float* g6 = out.g6;
    g6[0] = A[0][0] * A[0][0] * dL_da + A[0][0] * A[1][0] * dL_db + A[1][0] * A[1][0] * dL_dc;
    g6[3] = A[0][1] * A[0][1] * dL_da + A[0][1] * A[1][1] * dL_db + A[1][1] * A[1][1] * dL_dc;
    g6[5] = A[0][2] * A[0][2] * dL_da + A[0][2] * A[1][2] * dL_db + A[1][2] * A[1][2] * dL_dc;
    g6[1] = 2 * A[0][0] * A[0][1] * dL_da + (A[0][0] * A[1][1] + A[0][1] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][1] * dL_dc;
    g6[2] = 2 * A[0][0] * A[0][2] * dL_da + (A[0][0] * A[1][2] + A[0][2] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][2] * dL_dc;
    g6[4] = 2 * A[0][2] * A[0][1] * dL_da + (A[0][1] * A[1][2] + A[0][2] * A[1][1]) * dL_db + 2 * A[1][1] * A[1][2] * dL_dc;

    // d/dA[r][c] = 2 (AV)[r][c] dL_d{a,c} + (AV)[other][c] dL_db ;  A = J Rwv
    float dA[2][3];
    for (int c = 0; c < 3; c++) {
        dA[0][c] = 2 * AV[0][c] * dL_da + AV[1][c] * dL_db;
        dA[1][c] = 2 * AV[1][c] * dL_dc + AV[0][c] * dL_db;
    }
    const float dJ00 = dA[0][0] * v[0] + dA[0][1] * v[4] + dA[0][2] * v[8];
    const float dJ02 = dA[0][0] * v[2] + dA[0][1] * v[6] + dA[0][2] * v[10];
    const float dJ11 = dA[1][0] * v[1] + dA[1][1] * v[5] + dA[1][2] * v[9];
    const float dJ12 = dA[1][0] * v[2] + dA[1][1] * v[6] + dA[1][2] * v[10];
    const float itz = 1.f / tz, tz2 = itz * itz, tz3 = tz2 * itz;
    const float dL_dtx = x_grad_mul * -fx * tz2 * dJ02;
    const float dL_dty = y_grad_mul * -fy * tz2 * dJ12;
    float dL_dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * tx) * tz3 * dJ02 + (2 * fy * ty) * tz3 * dJ12;
    if (use_depth) dL_dtz -= g_iv / (tz * tz);
    out.dmean[0] = v[0] * dL_dtx + v[1] * dL_dty + v[2] * dL_dtz;
    out.dmean[1] = v[4] * dL_dtx + v[5] * dL_dty + v[6] * dL_dtz;
    out.dmean[2] = v[8] * dL_dtx + v[9] * dL_dty + v[10] * dL_dtz;

    // ---- K9c: cov3D -> scale, rotation ----
    if (!have_sr) return;
    const float dS[3][3] = {{g6[0], 0.5f * g6[1], 0.5f * g6[2]}, {0.5f * g6[1], g6[3], 0.5f * g6[4]}, {0.5f * g6[2], 0.5f * g6[4], g6[5]}};
    float dM[3][3];
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < 3; j++) dM[k][j] = 2.0f * (Mm[k][0] * dS[0][j] + Mm[k][1] * dS[1][j] + Mm[k][2] * dS[2][j]);
    for (int k = 0; k < 3; k++) out.dscale[k] = scale_mod * (R[0][k] * dM[k][0] + R[1][k] * dM[k][1] + R[2][k] * dM[k][2]);
    float dR[3][3];
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++) dR[j][k] = s[k] * dM[k][j];
    const float qr = q[0], qx = q[1], qy = q[2], qz = q[3];
    out.dq[0] = 2 * qz * (dR[1][0] - dR[0][1]) + 2 * qy * (dR[0][2] - dR[2][0]) + 2 * qx * (dR[2][1] - dR[1][2]);
    out.dq[1] = 2 * qy * (dR[0][1] + dR[1][0]) + 2 * qz * (dR[0][2] + dR[2][0]) + 2 * qr * (dR[2][1] - dR[1][2]) - 4 * qx * (dR[1][1] + dR[2][2]);
    out.dq[2] = 2 * qx * (dR[0][1] + dR[1][0]) + 2 * qr * (dR[0][2] - dR[2][0]) + 2 * qz * (dR[1][2] + dR[2][1]) - 4 * qy * (dR[0][0] + dR[2][2]);
    out.dq[3] = 2 * qr * (dR[1][0] - dR[0][1]) + 2 * qx * (dR[0][2] + dR[2][0]) + 2 * qy * (dR[1][2] + dR[2][1]) - 4 * qz * (dR[0][0] + dR[1][1]);
}

}  // namespace h3dgs
#undef H3_CG_FN
