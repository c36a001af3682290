// ecc_synth.cu -- synthetic benchmark data of libecc_b200: circular trajectory and analytic
// cone-beam projections of an ellipsoid phantom.
//
// WHAT (reference, code/):
//   ProjTable::makeCircularTrajectory      HeaderOnly/Utils/Projtable.hxx:138-165
//   cameraPerspective / cameraLookAt       LibProjectiveGeometry/CameraOpenGL.hxx:11-31
//   makeProjectionMatrix / normalize       LibProjectiveGeometry/ProjectionMatrix.cpp:12-18,133-145
//   apply_weight_cos_principal_ray         LibEpipolarConsistency/Gui/PreProccess.cpp:146-166
// The phantom itself is ours (the reference ships no synthetic data).
#include <cmath>

#include "ecc_geometry.cuh"
#include "ecc_internal.h"

namespace eccb200 {

namespace {

struct ViewRays {
    double Mi[9];  // inverse of the left 3x3 block, col-major
    double C[3];   // source position
    float sdd_px, ppu, ppv;
};

__global__ void synth_kernel(const ViewRays* views, int n_u, int n_v, const double* ell, int n_ell,
                             int cos_weight, int zero_border, float* images)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= n_u || v >= n_v) return;
    const ViewRays& V = views[blockIdx.z];
    double d[3];
    for (int r = 0; r < 3; r++) d[r] = V.Mi[r] * u + V.Mi[r + 3] * v + V.Mi[r + 6];
    const double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int r = 0; r < 3; r++) d[r] /= dn;
    double val = 0.0;
    for (int e = 0; e < n_ell; e++) {
        const double* q = ell + 7 * e;
        const double o0 = (V.C[0] - q[0]) / q[3], o1 = (V.C[1] - q[1]) / q[4], o2 = (V.C[2] - q[2]) / q[5];
        const double w0 = d[0] / q[3], w1 = d[1] / q[4], w2 = d[2] / q[5];
        const double a = w0 * w0 + w1 * w1 + w2 * w2;
        const double b = o0 * w0 + o1 * w1 + o2 * w2;
        const double c = o0 * o0 + o1 * o1 + o2 * o2 - 1.0;
        const double disc = b * b - a * c;
        if (disc > 0) val += q[6] * 2.0 * sqrt(disc) / a;
    }
    float pix = (float)val;
    if (cos_weight) {
        const float pou = (float)u - V.ppu, pov = (float)v - V.ppv;
        pix *= V.sdd_px / sqrtf(pou * pou + pov * pov + V.sdd_px * V.sdd_px);
    }
    if (zero_border && (u == 0 || v == 0 || u == n_u - 1 || v == n_v - 1)) pix = 0.f;
    images[((size_t)blockIdx.z * n_v + v) * n_u + u] = pix;
}

inline void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double norm3(const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

void view_rays(const double* P, ViewRays& V)
{
    const double* M = P;  // first nine entries = left 3x3 block, col-major
    const double det = det3d(M[0], M[3], M[6], M[1], M[4], M[7], M[2], M[5], M[8]);
    double* Mi = V.Mi;
    Mi[0] = (M[4] * M[8] - M[7] * M[5]) / det;
    Mi[3] = -(M[3] * M[8] - M[6] * M[5]) / det;
    Mi[6] = (M[3] * M[7] - M[6] * M[4]) / det;
    Mi[1] = -(M[1] * M[8] - M[7] * M[2]) / det;
    Mi[4] = (M[0] * M[8] - M[6] * M[2]) / det;
    Mi[7] = -(M[0] * M[7] - M[6] * M[1]) / det;
    Mi[2] = (M[1] * M[5] - M[4] * M[2]) / det;
    Mi[5] = -(M[0] * M[5] - M[3] * M[2]) / det;
    Mi[8] = (M[0] * M[4] - M[3] * M[1]) / det;
    for (int r = 0; r < 3; r++) V.C[r] = -(Mi[r] * P[9] + Mi[r + 3] * P[10] + Mi[r + 6] * P[11]);
    // intrinsics of K[R|t]: principal point = (m1.m3, m2.m3)/|m3|^2, focal length = |m1 - ppu m3|/|m3|
    const double m1[3] = {P[0], P[3], P[6]}, m2[3] = {P[1], P[4], P[7]}, m3[3] = {P[2], P[5], P[8]};
    const double n3 = m3[0] * m3[0] + m3[1] * m3[1] + m3[2] * m3[2];
    const double ppu = (m1[0] * m3[0] + m1[1] * m3[1] + m1[2] * m3[2]) / n3;
    const double ppv = (m2[0] * m3[0] + m2[1] * m3[1] + m2[2] * m3[2]) / n3;
    const double t[3] = {m1[0] - ppu * m3[0], m1[1] - ppu * m3[1], m1[2] - ppu * m3[2]};
    V.sdd_px = (float)(norm3(t) / std::sqrt(n3));
    V.ppu = (float)ppu;
    V.ppv = (float)ppv;
}

}  // namespace

int synth_projections(ecc_context* ctx, const double* Ps_h, int n, int n_u, int n_v,
                      const double* ell_h, int n_ell, int cos_weight, int zero_border,
                      float* images_d)
{
    std::vector<ViewRays> views(n);
    for (int i = 0; i < n; i++) view_rays(Ps_h + 12 * i, views[i]);
    ViewRays* views_d = nullptr;
    double* ell_d = nullptr;
    ECC_CUDA(ctx, cudaMalloc(&views_d, sizeof(ViewRays) * n));
    ECC_CUDA(ctx, cudaMalloc(&ell_d, sizeof(double) * 7 * (n_ell > 0 ? n_ell : 1)));
    ECC_CUDA(ctx, cudaMemcpyAsync(views_d, views.data(), sizeof(ViewRays) * n, cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(ell_d, ell_h, sizeof(double) * 7 * n_ell, cudaMemcpyHostToDevice, ctx->stream));
    dim3 block(32, 8);
    const int slot = prof_begin(ctx, FAM_SYNTH);
    for (int first = 0; first < n; first += 32768) {  // grid.z limit
        const int cnt = (n - first < 32768) ? n - first : 32768;
        dim3 grid((n_u + 31) / 32, (n_v + 7) / 8, cnt);
        synth_kernel<<<grid, block, 0, ctx->stream>>>(views_d + first, n_u, n_v, ell_d, n_ell, cos_weight, zero_border,
                                                      images_d + (size_t)first * n_u * n_v);
    }
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(views_d);
    cudaFree(ell_d);
    return ECC_OK;
}

}  // namespace eccb200

extern "C" void ecc_make_circular_trajectory(int n_proj, double sid, double sdd, int n_u, int n_v,
                                             double max_angle_deg, double pixel_spacing, double* Ps)
{
    using namespace eccb200;
    const double kPi = 3.14159265358979323846;
    // cameraPerspective(fovy, n_u, n_v): focal length in pixels from the (reference's) fovy
    const double fovy = std::atan(n_v * pixel_spacing / sdd);
    const double f = n_v / (2.0 * std::tan(0.5 * fovy));
    const double ct = std::cos(0.5 * kPi), st = std::sin(0.5 * kPi);
    for (int i = 0; i < n_proj; i++) {
        const double ang = i * (max_angle_deg / n_proj) / 180.0 * kPi;
        const double eye[3] = {sid * std::cos(ang), 0.0, sid * std::sin(ang)};
        double fwd[3] = {-eye[0], -eye[1], -eye[2]};
        const double nf = norm3(fwd);
        for (int k = 0; k < 3; k++) fwd[k] /= nf;
        const double up0[3] = {0, 1, 0};
        double left[3], up[3];
        cross3(up0, fwd, left);
        const double nl = norm3(left);
        for (int k = 0; k < 3; k++) left[k] /= nl;
        cross3(fwd, left, up);
        const double* rows[3] = {left, up, fwd};
        double R[3][3], t[3];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) R[r][c] = (r == 2 ? -1.0 : 1.0) * rows[r][c];
        for (int r = 0; r < 3; r++) t[r] = -(R[r][0] * eye[0] + R[r][1] * eye[1] + R[r][2] * eye[2]);
        // P = K [R|t] with K = [f 0 n_u/2; 0 f n_v/2; 0 0 1]
        double Pm[3][4];
        for (int c = 0; c < 3; c++) {
            Pm[0][c] = f * R[0][c] + 0.5 * n_u * R[2][c];
            Pm[1][c] = f * R[1][c] + 0.5 * n_v * R[2][c];
            Pm[2][c] = R[2][c];
        }
        Pm[0][3] = f * t[0] + 0.5 * n_u * t[2];
        Pm[1][3] = f * t[1] + 0.5 * n_v * t[2];
        Pm[2][3] = t[2];
        auto normalize = [](double (*M)[4]) {
            double n3 = std::sqrt(M[2][0] * M[2][0] + M[2][1] * M[2][1] + M[2][2] * M[2][2]);
            const double d = det3d(M[0][0], M[0][1], M[0][2], M[1][0], M[1][1], M[1][2], M[2][0], M[2][1], M[2][2]);
            if (d < 0) n3 = -n3;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++) M[r][c] *= (1.0 / n3);
        };
        normalize(Pm);
        double Q[3][4];
        for (int r = 0; r < 3; r++) {  // times the 4x4 rotation about x by 90 degrees
            Q[r][0] = Pm[r][0];
            Q[r][1] = Pm[r][1] * ct + Pm[r][2] * st;
            Q[r][2] = -Pm[r][1] * st + Pm[r][2] * ct;
            Q[r][3] = Pm[r][3];
        }
        normalize(Q);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) Ps[12 * i + r + 3 * c] = Q[r][c];
    }
}
