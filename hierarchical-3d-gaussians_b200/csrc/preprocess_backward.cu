// preprocess_backward.cu -- K8 + K9 fused: per-Gaussian chain rule from the 2D-space
// accumulators (dmean2D, dconic, dopacity, dcolor, dinvdepth) back to the inputs
// (replaces BACKWARD::computeCov2DCUDA + BACKWARD::preprocessCUDA).  Semantics per
// oracle/oracle.c::oracle_preprocess_backward.  The covariance chain (K8, K9c) is evaluated in
// fp64: the published formulas contain cancellations (denom - a*c = -b^2; the rotation
// gradient of a near-isotropic Gaussian) that cost fp32 one to two digits, and a few hundred
// DFMA per Gaussian are free in an HBM-bound stream.  One thread per Gaussian; a pure HBM
// stream: reads 40 B accum + 44 B params + 192 B SH, writes 248 B of gradients.
// Every output row is written (zeros for culled Gaussians) so the caller never
// pays a separate memset pass over the gradient tensors.
#include "common.cuh"

namespace h3dgs {

__device__ __constant__ float bSH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                           -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float bSH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                           0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                           -0.5900435899266435f};
constexpr float bSH_C0 = 0.28209479177387814f;
constexpr float bSH_C1 = 0.4886025119029199f;

__device__ __forceinline__ void store_zero(float* p, int n) {
    for (int k = 0; k < n; k++) p[k] = 0.f;
}

__global__ void __launch_bounds__(256)
preprocess_backward_kernel(int P, int deg, int M, const float* __restrict__ means3D, const float* __restrict__ scales,
                           float scale_mod, const float* __restrict__ rots, const float* __restrict__ shs,
                           const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
                           const float* __restrict__ ts, const int* __restrict__ ridx, const int* __restrict__ pidx,
                           const float* __restrict__ view, const float* __restrict__ proj,
                           const float* __restrict__ campos, int W, int H, float tanx, float tany, float fx, float fy,
                           int use_depth, const int* __restrict__ radii, const Record* __restrict__ records,
                           const float* __restrict__ accum, float* __restrict__ dL_dmeans3D,
                           float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dsh, float* __restrict__ dL_dcolors,
                           float* __restrict__ dL_dopacities, float* __restrict__ dL_dscales,
                           float* __restrict__ dL_drots, float* __restrict__ dL_dcov3D)
{
    __shared__ float s_view[16], s_proj[16], s_cam[3];
    if (threadIdx.x < 16) { s_view[threadIdx.x] = view[threadIdx.x]; s_proj[threadIdx.x] = proj[threadIdx.x]; }
    if (threadIdx.x < 3) s_cam[threadIdx.x] = campos[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int SH3 = M * 3;

    if (radii[i] <= 0) {
        store_zero(dL_dmeans2D + 3 * i, 3);
        if (ridx) return;                    // scatter mode: full-size gradients are pre-zeroed
        store_zero(dL_dmeans3D + 3 * i, 3);
        dL_dopacities[i] = 0.f;
        if (dL_dsh) {
            float* o = dL_dsh + (size_t)i * SH3;
            if ((SH3 & 3) == 0) { for (int k = 0; k < SH3 / 4; k++) reinterpret_cast<float4*>(o)[k] = make_float4(0, 0, 0, 0); }
            else store_zero(o, SH3);
        }
        if (dL_dcolors) store_zero(dL_dcolors + 3 * i, 3);
        if (dL_dscales) store_zero(dL_dscales + 3 * i, 3);
        if (dL_drots) store_zero(dL_drots + 4 * i, 4);
        if (dL_dcov3D) store_zero(dL_dcov3D + 6 * i, 6);
        return;
    }

    const float* v = s_view;
    const float* ac = accum + (size_t)i * kAccum;
    // the per-tile replay leaves its constant factors to us: d(pixel)/d(ndc) = 0.5 W, 0.5 H and the
    // -1/2 of the quadratic form
    const float g_mx = ac[0] * (0.5f * W), g_my = ac[1] * (0.5f * H);
    const float dcx = -0.5f * ac[2], dcy = -0.5f * ac[3], dcz = -0.5f * ac[4], g_op = ac[5];
    const float g_r = ac[6], g_g = ac[7], g_b = ac[8], g_iv = ac[9];
    // gather + parent lerp exactly as the forward did (preprocess.cu)
    int c = i, p = i;
    float t = 1.0f, u = 0.0f;
    if (ridx) {
        c = ridx[i]; p = pidx[i]; if (p < 0) p = c;
        t = ts[i]; u = 1.0f - t;
    }
    const bool lerp = ridx != nullptr && u != 0.0f;
#define LERP(a, b) (lerp ? __fadd_rn(__fmul_rn(t, (a)), __fmul_rn(u, (b))) : (a))
    const float mx = LERP(means3D[3 * c], means3D[3 * p]);
    const float my = LERP(means3D[3 * c + 1], means3D[3 * p + 1]);
    const float mz = LERP(means3D[3 * c + 2], means3D[3 * p + 2]);
    float qsign = 1.0f;

    // cov3D (recomputed: cheaper than a 24-B round trip through HBM)
    double cov6[6];
    double R[3][3], Mm[3][3], sc[3];
    double qr = 1., qx = 0., qy = 0., qz = 0.;
    if (cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov6[k] = cov3D_precomp[6 * c + k];
    } else {
        float4 qq = *reinterpret_cast<const float4*>(rots + 4 * c);
        if (lerp) {
            float4 qp = *reinterpret_cast<const float4*>(rots + 4 * p);
            const float dot = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qq.x, qp.x), __fmul_rn(qq.y, qp.y)), __fmul_rn(qq.z, qp.z)), __fmul_rn(qq.w, qp.w));
            if (dot < 0.f) qsign = -1.0f;
            qq.x = LERP(qq.x, qsign * qp.x); qq.y = LERP(qq.y, qsign * qp.y);
            qq.z = LERP(qq.z, qsign * qp.z); qq.w = LERP(qq.w, qsign * qp.w);
        }
        qr = qq.x; qx = qq.y; qy = qq.z; qz = qq.w;
        R[0][0] = 1. - 2. * (qy * qy + qz * qz); R[0][1] = 2. * (qx * qy - qr * qz); R[0][2] = 2. * (qx * qz + qr * qy);
        R[1][0] = 2. * (qx * qy + qr * qz); R[1][1] = 1. - 2. * (qx * qx + qz * qz); R[1][2] = 2. * (qy * qz - qr * qx);
        R[2][0] = 2. * (qx * qz - qr * qy); R[2][1] = 2. * (qy * qz + qr * qx); R[2][2] = 1. - 2. * (qx * qx + qy * qy);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            sc[k] = (double)scale_mod * (double)LERP(scales[3 * c + k], scales[3 * p + k]);
#pragma unroll
            for (int j = 0; j < 3; j++) Mm[k][j] = sc[k] * R[j][k];
        }
        int o = 0;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = a; b < 3; b++) cov6[o++] = Mm[0][a] * Mm[0][b] + Mm[1][a] * Mm[1][b] + Mm[2][a] * Mm[2][b];
    }

    double dmean[3] = {0., 0., 0.};
    double g6[6];
    // ---- K8: conic -> cov2D -> cov3D, and mean through the Jacobian ----
    {
        double tx = (double)v[0] * mx + (double)v[4] * my + (double)v[8] * mz + (double)v[12];
        double ty = (double)v[1] * mx + (double)v[5] * my + (double)v[9] * mz + (double)v[13];
        const double tz = (double)v[2] * mx + (double)v[6] * my + (double)v[10] * mz + (double)v[14];
        const double limx = kFovClamp * tanx, limy = kFovClamp * tany;
        const double txtz = tx / tz, tytz = ty / tz;
        tx = fmin(limx, fmax(-limx, txtz)) * tz;
        ty = fmin(limy, fmax(-limy, tytz)) * tz;
        const double x_grad_mul = (txtz < -limx || txtz > limx) ? 0. : 1.;
        const double y_grad_mul = (tytz < -limy || tytz > limy) ? 0. : 1.;
        const double J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
        const double J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
        double A[2][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            A[0][c] = J00 * (double)v[4 * c + 0] + J02 * (double)v[4 * c + 2];
            A[1][c] = J11 * (double)v[4 * c + 1] + J12 * (double)v[4 * c + 2];
        }
        const double V[3][3] = {{cov6[0], cov6[1], cov6[2]}, {cov6[1], cov6[3], cov6[4]}, {cov6[2], cov6[4], cov6[5]}};
        double AV[2][3];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) AV[r][c] = A[r][0] * V[0][c] + A[r][1] * V[1][c] + A[r][2] * V[2][c];
        const double a = (AV[0][0] * A[0][0] + AV[0][1] * A[0][1] + AV[0][2] * A[0][2]) + kDilation;
        const double b = AV[0][0] * A[1][0] + AV[0][1] * A[1][1] + AV[0][2] * A[1][2];
        const double c_ = (AV[1][0] * A[1][0] + AV[1][1] * A[1][1] + AV[1][2] * A[1][2]) + kDilation;
        const double denom = a * c_ - b * b;
        double dL_da = 0., dL_db = 0., dL_dc = 0.;
        const double denom2inv = 1.0 / ((denom * denom) + 0.0000001);
#pragma unroll
        for (int k = 0; k < 6; k++) g6[k] = 0.;
        if (denom2inv != 0.) {
            dL_da = denom2inv * (-c_ * c_ * dcx + 2 * b * c_ * dcy + (denom - a * c_) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c_) * dcx);
            dL_db = denom2inv * 2 * (b * c_ * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            g6[0] = A[0][0] * A[0][0] * dL_da + A[0][0] * A[1][0] * dL_db + A[1][0] * A[1][0] * dL_dc;
            g6[3] = A[0][1] * A[0][1] * dL_da + A[0][1] * A[1][1] * dL_db + A[1][1] * A[1][1] * dL_dc;
            g6[5] = A[0][2] * A[0][2] * dL_da + A[0][2] * A[1][2] * dL_db + A[1][2] * A[1][2] * dL_dc;
            g6[1] = 2 * A[0][0] * A[0][1] * dL_da + (A[0][0] * A[1][1] + A[0][1] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][1] * dL_dc;
            g6[2] = 2 * A[0][0] * A[0][2] * dL_da + (A[0][0] * A[1][2] + A[0][2] * A[1][0]) * dL_db + 2 * A[1][0] * A[1][2] * dL_dc;
            g6[4] = 2 * A[0][2] * A[0][1] * dL_da + (A[0][1] * A[1][2] + A[0][2] * A[1][1]) * dL_db + 2 * A[1][1] * A[1][2] * dL_dc;
        }
        double dA[2][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            dA[0][c] = 2 * AV[0][c] * dL_da + AV[1][c] * dL_db;
            dA[1][c] = 2 * AV[1][c] * dL_dc + AV[0][c] * dL_db;
        }
        const double dJ00 = dA[0][0] * (double)v[0] + dA[0][1] * (double)v[4] + dA[0][2] * (double)v[8];
        const double dJ02 = dA[0][0] * (double)v[2] + dA[0][1] * (double)v[6] + dA[0][2] * (double)v[10];
        const double dJ11 = dA[1][0] * (double)v[1] + dA[1][1] * (double)v[5] + dA[1][2] * (double)v[9];
        const double dJ12 = dA[1][0] * (double)v[2] + dA[1][1] * (double)v[6] + dA[1][2] * (double)v[10];
        const double itz = 1. / tz, tz2 = itz * itz, tz3 = tz2 * itz;
        const double dL_dtx = x_grad_mul * -fx * tz2 * dJ02;
        const double dL_dty = y_grad_mul * -fy * tz2 * dJ12;
        double dL_dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * tx) * tz3 * dJ02 + (2 * fy * ty) * tz3 * dJ12;
        if (use_depth) dL_dtz -= g_iv / (tz * tz);
        dmean[0] += (double)v[0] * dL_dtx + (double)v[1] * dL_dty + (double)v[2] * dL_dtz;
        dmean[1] += (double)v[4] * dL_dtx + (double)v[5] * dL_dty + (double)v[6] * dL_dtz;
        dmean[2] += (double)v[8] * dL_dtx + (double)v[9] * dL_dty + (double)v[10] * dL_dtz;
    }
    // ---- K9a: screen-space mean -> 3D mean ----
    {
        const float* q = s_proj;
        const float hw = q[3] * mx + q[7] * my + q[11] * mz + q[15];
        const float m_w = 1.0f / (hw + kWEps);
        const float mul1 = (q[0] * mx + q[4] * my + q[8] * mz + q[12]) * m_w * m_w;
        const float mul2 = (q[1] * mx + q[5] * my + q[9] * mz + q[13]) * m_w * m_w;
        dmean[0] += (q[0] * m_w - q[3] * mul1) * g_mx + (q[1] * m_w - q[3] * mul2) * g_my;
        dmean[1] += (q[4] * m_w - q[7] * mul1) * g_mx + (q[5] * m_w - q[7] * mul2) * g_my;
        dmean[2] += (q[8] * m_w - q[11] * mul1) * g_mx + (q[9] * m_w - q[11] * mul2) * g_my;
    }
    dL_dmeans2D[3 * i] = g_mx; dL_dmeans2D[3 * i + 1] = g_my; dL_dmeans2D[3 * i + 2] = 0.f;
    // EMIT: direct store (flat / pre-gathered inputs) or t / (1-t) scatter into the full-size
    // gradients (red.global.add; a parent row is shared by its k children)
#define EMIT(ptr, width, k, val)                                                           \
    do {                                                                                   \
        const float v__ = (val);                                                           \
        if (!ridx) (ptr)[(size_t)i * (width) + (k)] = v__;                                 \
        else {                                                                             \
            atomicAdd((ptr) + (size_t)c * (width) + (k), t * v__);                         \
            if (lerp) atomicAdd((ptr) + (size_t)p * (width) + (k), u * v__);               \
        }                                                                                  \
    } while (0)
    EMIT(dL_dopacities, 1, 0, g_op);

    // ---- K9b: colour -> SH and view direction ----
    if (colors_precomp) {
        if (dL_dcolors) { dL_dcolors[3 * c] = g_r; dL_dcolors[3 * c + 1] = g_g; dL_dcolors[3 * c + 2] = g_b; }
    } else {
        const uint32_t kb = __float_as_uint(records[i].b.w);
        const float dRGB[3] = {(kb >> kClampShift) & 1u ? 0.f : g_r, (kb >> (kClampShift + 1)) & 1u ? 0.f : g_g, (kb >> (kClampShift + 2)) & 1u ? 0.f : g_b};
        const float d0x = mx - s_cam[0], d0y = my - s_cam[1], d0z = mz - s_cam[2];
        const float len = sqrtf(d0x * d0x + d0y * d0y + d0z * d0z);
        const float x = d0x / len, y = d0y / len, z = d0z / len;
        const float* sh = shs + (size_t)c * SH3;
        const float* shp = shs + (size_t)p * SH3;
        float* dsh = dL_dsh + (size_t)i * SH3;
        float out[48];
        float ddx = 0.f, ddy = 0.f, ddz = 0.f;
        const int ncoef = (deg + 1) * (deg + 1);
        float c_[48];
        if ((SH3 & 3) == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++)
                if (4 * k < 3 * ncoef) {
                    float4 t4 = __ldg(reinterpret_cast<const float4*>(sh) + k);
                    if (lerp) {
                        const float4 p4 = __ldg(reinterpret_cast<const float4*>(shp) + k);
                        t4.x = LERP(t4.x, p4.x); t4.y = LERP(t4.y, p4.y); t4.z = LERP(t4.z, p4.z); t4.w = LERP(t4.w, p4.w);
                    }
                    c_[4 * k] = t4.x; c_[4 * k + 1] = t4.y; c_[4 * k + 2] = t4.z; c_[4 * k + 3] = t4.w;
                }
        } else {
#pragma unroll
            for (int k = 0; k < 48; k++) if (k < 3 * ncoef) c_[k] = LERP(__ldg(sh + k), __ldg(shp + k));
        }
#pragma unroll
        for (int k = 0; k < 48; k++) out[k] = 0.f;
#define S(k, ch) c_[(k) * 3 + (ch)]
#define DS(k, ch) out[(k) * 3 + (ch)]
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float g = dRGB[ch];
            float dx_ = 0.f, dy_ = 0.f, dz_ = 0.f;
            DS(0, ch) = bSH_C0 * g;
            if (deg > 0) {
                DS(1, ch) = -bSH_C1 * y * g; DS(2, ch) = bSH_C1 * z * g; DS(3, ch) = -bSH_C1 * x * g;
                dx_ = -bSH_C1 * S(3, ch); dy_ = -bSH_C1 * S(1, ch); dz_ = bSH_C1 * S(2, ch);
                if (deg > 1) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                    DS(4, ch) = bSH_C2[0] * xy * g; DS(5, ch) = bSH_C2[1] * yz * g;
                    DS(6, ch) = bSH_C2[2] * (2.f * zz - xx - yy) * g;
                    DS(7, ch) = bSH_C2[3] * xz * g; DS(8, ch) = bSH_C2[4] * (xx - yy) * g;
                    dx_ += bSH_C2[0] * y * S(4, ch) + bSH_C2[2] * 2.f * -x * S(6, ch) + bSH_C2[3] * z * S(7, ch) + bSH_C2[4] * 2.f * x * S(8, ch);
                    dy_ += bSH_C2[0] * x * S(4, ch) + bSH_C2[1] * z * S(5, ch) + bSH_C2[2] * 2.f * -y * S(6, ch) + bSH_C2[4] * 2.f * -y * S(8, ch);
                    dz_ += bSH_C2[1] * y * S(5, ch) + bSH_C2[2] * 2.f * 2.f * z * S(6, ch) + bSH_C2[3] * x * S(7, ch);
                    if (deg > 2) {
                        DS(9, ch) = bSH_C3[0] * y * (3.f * xx - yy) * g;
                        DS(10, ch) = bSH_C3[1] * xy * z * g;
                        DS(11, ch) = bSH_C3[2] * y * (4.f * zz - xx - yy) * g;
                        DS(12, ch) = bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy) * g;
                        DS(13, ch) = bSH_C3[4] * x * (4.f * zz - xx - yy) * g;
                        DS(14, ch) = bSH_C3[5] * z * (xx - yy) * g;
                        DS(15, ch) = bSH_C3[6] * x * (xx - 3.f * yy) * g;
                        dx_ += bSH_C3[0] * S(9, ch) * 3.f * 2.f * xy + bSH_C3[1] * S(10, ch) * yz + bSH_C3[2] * S(11, ch) * -2.f * xy
                             + bSH_C3[3] * S(12, ch) * -3.f * 2.f * xz + bSH_C3[4] * S(13, ch) * (-3.f * xx + 4.f * zz - yy)
                             + bSH_C3[5] * S(14, ch) * 2.f * xz + bSH_C3[6] * S(15, ch) * 3.f * (xx - yy);
                        dy_ += bSH_C3[0] * S(9, ch) * 3.f * (xx - yy) + bSH_C3[1] * S(10, ch) * xz + bSH_C3[2] * S(11, ch) * (-3.f * yy + 4.f * zz - xx)
                             + bSH_C3[3] * S(12, ch) * -3.f * 2.f * yz + bSH_C3[4] * S(13, ch) * -2.f * xy
                             + bSH_C3[5] * S(14, ch) * -2.f * yz + bSH_C3[6] * S(15, ch) * -3.f * 2.f * xy;
                        dz_ += bSH_C3[1] * S(10, ch) * xy + bSH_C3[2] * S(11, ch) * 4.f * 2.f * yz + bSH_C3[3] * S(12, ch) * 3.f * (2.f * zz - xx - yy)
                             + bSH_C3[4] * S(13, ch) * 4.f * 2.f * xz + bSH_C3[5] * S(14, ch) * (xx - yy);
                    }
                }
            }
            ddx += dx_ * g; ddy += dy_ * g; ddz += dz_ * g;
        }
#undef S
#undef DS
        if (ridx && (SH3 & 3) == 0) {
            // 128-bit vector reductions (red.global.add.v4.f32, sm_90+): 12 per 192-B SH row instead
            // of 48 scalar ones; only the active coefficients carry gradient, the rest stays zero
#pragma unroll
            for (int k = 0; k < 12; k++)
                if (4 * k < 3 * ncoef) {
                    atomicAdd(reinterpret_cast<float4*>(dL_dsh + (size_t)c * SH3) + k,
                              make_float4(t * out[4 * k], t * out[4 * k + 1], t * out[4 * k + 2], t * out[4 * k + 3]));
                    if (lerp)
                        atomicAdd(reinterpret_cast<float4*>(dL_dsh + (size_t)p * SH3) + k,
                                  make_float4(u * out[4 * k], u * out[4 * k + 1], u * out[4 * k + 2], u * out[4 * k + 3]));
                }
        } else if (ridx) {
#pragma unroll
            for (int k = 0; k < 48; k++)
                if (k < 3 * ncoef) {
                    atomicAdd(dL_dsh + (size_t)c * SH3 + k, t * out[k]);
                    if (lerp) atomicAdd(dL_dsh + (size_t)p * SH3 + k, u * out[k]);
                }
        } else if ((SH3 & 3) == 0) {
#pragma unroll
            for (int k = 0; k < 12; k++)
                if (4 * k < SH3) reinterpret_cast<float4*>(dsh)[k] = make_float4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
        } else {
#pragma unroll
            for (int k = 0; k < 48; k++) if (k < SH3) dsh[k] = out[k];
        }
        const float sum2 = d0x * d0x + d0y * d0y + d0z * d0z;
        const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        dmean[0] += ((sum2 - d0x * d0x) * ddx - d0y * d0x * ddy - d0z * d0x * ddz) * invsum32;
        dmean[1] += (-d0x * d0y * ddx + (sum2 - d0y * d0y) * ddy - d0z * d0y * ddz) * invsum32;
        dmean[2] += (-d0x * d0z * ddx - d0y * d0z * ddy + (sum2 - d0z * d0z) * ddz) * invsum32;
    }
    EMIT(dL_dmeans3D, 3, 0, (float)dmean[0]); EMIT(dL_dmeans3D, 3, 1, (float)dmean[1]); EMIT(dL_dmeans3D, 3, 2, (float)dmean[2]);

    // ---- K9c: cov3D -> scale, rotation ----
    if (cov3D_precomp) {
        if (dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) dL_dcov3D[6 * c + k] = (float)g6[k];
        }
    } else {
        const double dS[3][3] = {{g6[0], 0.5 * g6[1], 0.5 * g6[2]}, {0.5 * g6[1], g6[3], 0.5 * g6[4]}, {0.5 * g6[2], 0.5 * g6[4], g6[5]}};
        double dM[3][3];
#pragma unroll
        for (int k = 0; k < 3; k++)
#pragma unroll
            for (int j = 0; j < 3; j++) dM[k][j] = 2.0 * (Mm[k][0] * dS[0][j] + Mm[k][1] * dS[1][j] + Mm[k][2] * dS[2][j]);
#pragma unroll
        for (int k = 0; k < 3; k++)
            EMIT(dL_dscales, 3, k, (float)((double)scale_mod * (R[0][k] * dM[k][0] + R[1][k] * dM[k][1] + R[2][k] * dM[k][2])));
        double dR[3][3];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) dR[j][k] = sc[k] * dM[k][j];
        float4 dq;
        dq.x = (float)(2 * qz * (dR[1][0] - dR[0][1]) + 2 * qy * (dR[0][2] - dR[2][0]) + 2 * qx * (dR[2][1] - dR[1][2]));
        dq.y = (float)(2 * qy * (dR[0][1] + dR[1][0]) + 2 * qz * (dR[0][2] + dR[2][0]) + 2 * qr * (dR[2][1] - dR[1][2]) - 4 * qx * (dR[1][1] + dR[2][2]));
        dq.z = (float)(2 * qx * (dR[0][1] + dR[1][0]) + 2 * qr * (dR[0][2] - dR[2][0]) + 2 * qz * (dR[1][2] + dR[2][1]) - 4 * qy * (dR[0][0] + dR[2][2]));
        dq.w = (float)(2 * qr * (dR[1][0] - dR[0][1]) + 2 * qx * (dR[0][2] + dR[2][0]) + 2 * qy * (dR[1][2] + dR[2][1]) - 4 * qz * (dR[0][0] + dR[1][1]));
        if (!ridx) *reinterpret_cast<float4*>(dL_drots + 4 * i) = dq;
        else {
            atomicAdd(reinterpret_cast<float4*>(dL_drots) + c, make_float4(t * dq.x, t * dq.y, t * dq.z, t * dq.w));
            if (lerp) {
                const float us = u * qsign;
                atomicAdd(reinterpret_cast<float4*>(dL_drots) + p, make_float4(us * dq.x, us * dq.y, us * dq.z, us * dq.w));
            }
        }
        if (dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) dL_dcov3D[6 * c + k] = (float)g6[k];
        }
    }
#undef EMIT
#undef LERP
}

int launch_preprocess_backward(const h3dgs_raster_args& a, const int32_t* radii, const Record* records,
                               const float* accum, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dsh,
                               float* dL_dcolors, float* dL_dopacities, float* dL_dscales, float* dL_drots,
                               float* dL_dcov3D, cudaStream_t s)
{
    if (a.P == 0) return H3DGS_OK;
    const float fx = a.image_width / (2.0f * a.tanfovx), fy = a.image_height / (2.0f * a.tanfovy);
    ProfScope prof(H3DGS_STAGE_PREPROCESS_BWD, s);
    preprocess_backward_kernel<<<(a.P + 255) / 256, 256, 0, s>>>(
        a.P, a.sh_degree, a.sh_coeffs, a.means3D, a.scales, a.scale_modifier, a.rotations, a.shs, a.cov3D_precomp,
        a.colors_precomp, a.interpolation_weights, a.render_indices, a.parent_indices, a.viewmatrix, a.projmatrix, a.campos, a.image_width, a.image_height, a.tanfovx, a.tanfovy,
        fx, fy, a.do_depth, radii, records, accum, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacities,
        dL_dscales, dL_drots, dL_dcov3D);
    H3_LAUNCHED("preprocess_backward", a.debug, s);
    return H3DGS_OK;
}

}  // namespace h3dgs
