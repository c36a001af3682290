// ecc_models.cuh -- the perturbation models of the correction loops, host + device with the SAME bits.
//
// What is computed follows the reference (aaichert/EpipolarConsistency, code/LibProjectiveGeometry/Models/):
//   ModelSimilarity2D::getInstance           ModelSimilarity2D.hxx:52-72   x = translation u, v, rotation, scale -> 3x3
//   ModelSimilarity3D::getInstance           ModelSimilarity3D.hxx:64-87   x = translation X, Y, Z, rotation about X, Y, Z
//                                                                          (R = Rx Ry Rz), scale -> 4x4
//   ModelCameraSimilarity2D3D::getInstance   ModelCameraSimilarity2D3D.hxx:89-92   P' = H2D(x[0..3]) * P * T3D(x[4..10])
//   ModelFDCT::applyModel                    tools/FDCTMotionCorrection/ModelFDCT.hxx:26-62   one such model per view
// How it is computed is ours: everything in fp64 through the separately rounded operations of R2 (ecc_geometry.cuh: no
// fused multiply-add on either side) and an own sine / cosine (det_sincos) instead of the C library's on the host and
// libdevice's on the GPU, so that the K x n matrices a correction loop asks for can be expanded ON THE DEVICE from their
// parameter vectors (ecc_evaluate_batch_params: 11 doubles per model instead of a host round trip per matrix set) and
// still be, bit for bit, the matrices the host-side models of the facade hand out (ecc_model_*).
#pragma once
#include "ecc_geometry.cuh"

namespace eccb200 {

// sin and cos of x in fp64 with one IEEE operation sequence for host and device.  Argument reduction by pi/2 in three
// 33-bit pieces (Cody-Waite; exact products for |x| < ~1e5 rad -- model angles are fractions of a radian), then the
// classical minimax polynomials on [-pi/4, pi/4] (the published fdlibm coefficients).  Within one ulp of the correctly
// rounded values (tests/test_library_cpu.py compares with numpy over six decades).
__host__ __device__ inline void det_sincos(double x, double* s_out, double* c_out)
{
    const double two_over_pi = 6.36619772367581382433e-01;
    const double p1 = 1.57079632673412561417e+00, p2 = 6.07710050630396597660e-11, p3 = 2.02226624871116645580e-21;
    const double fn = rint(R2::mul(x, two_over_pi));
    double y = R2::sub(x, R2::mul(fn, p1));
    y = R2::sub(y, R2::mul(fn, p2));
    y = R2::sub(y, R2::mul(fn, p3));
    const double z = R2::mul(y, y);
    // sin(y) = y + y z (S1 + z (S2 + z (S3 + z (S4 + z (S5 + z S6)))))
    double r = R2::add(-2.50507602534068634195e-08, R2::mul(z, 1.58969099521155010221e-10));
    r = R2::add(2.75573137070700676789e-06, R2::mul(z, r));
    r = R2::add(-1.98412698298579493134e-04, R2::mul(z, r));
    r = R2::add(8.33333333332248946124e-03, R2::mul(z, r));
    r = R2::add(-1.66666666666666324348e-01, R2::mul(z, r));
    const double sn = R2::add(y, R2::mul(R2::mul(y, z), r));
    // cos(y) = w + (((1 - w) - z/2) + z z (C1 + z (C2 + ... z C6))),  w = 1 - z/2
    double q = R2::add(2.08757232129817482790e-09, R2::mul(z, -1.13596475577881948265e-11));
    q = R2::add(-2.75573143513906633035e-07, R2::mul(z, q));
    q = R2::add(2.48015872894767294178e-05, R2::mul(z, q));
    q = R2::add(-1.38888888888741095749e-03, R2::mul(z, q));
    q = R2::add(4.16666666666666019037e-02, R2::mul(z, q));
    const double hz = R2::mul(0.5, z);
    const double w = R2::sub(1.0, hz);
    const double cs = R2::add(w, R2::add(R2::sub(R2::sub(1.0, w), hz), R2::mul(R2::mul(z, z), q)));
    const long long k = (long long)fn;
    switch ((int)(k & 3)) {
        case 0: *s_out = sn; *c_out = cs; break;
        case 1: *s_out = cs; *c_out = -sn; break;
        case 2: *s_out = -sn; *c_out = -cs; break;
        default: *s_out = -cs; *c_out = sn; break;
    }
}

// 3x3, column-major (H[r + 3c]).
__host__ __device__ inline void model_similarity_2d(const double* x, double* H)
{
    for (int i = 0; i < 9; i++) H[i] = (i % 4 == 0) ? 1.0 : 0.0;
    if (x[2] != 0) {
        double s, c;
        det_sincos(x[2], &s, &c);
        H[0] = c; H[0 + 3] = -s;
        H[1] = s; H[1 + 3] = c;
    }
    H[0 + 6] = x[0];
    H[1 + 6] = x[1];
    if (x[3] != 0) {
        const double f = R2::add(1.0, x[3]);
        H[0] = R2::mul(H[0], f); H[1] = R2::mul(H[1], f); H[3] = R2::mul(H[3], f); H[4] = R2::mul(H[4], f);
    }
}

// 4x4, column-major (T[r + 4c]).
__host__ __device__ inline void model_similarity_3d(const double* x, double* T)
{
    for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    if (x[3] != 0 || x[4] != 0 || x[5] != 0) {
        double sx, cx, sy, cy, sz, cz;
        det_sincos(x[3], &sx, &cx);
        det_sincos(x[4], &sy, &cy);
        det_sincos(x[5], &sz, &cz);
        const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};  // row-major
        const double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
        const double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
        double A[9], R[9];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                double acc = 0.0;
                for (int k = 0; k < 3; k++) acc = R2::add(acc, R2::mul(Rx[3 * r + k], Ry[3 * k + c]));
                A[3 * r + c] = acc;
            }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                double acc = 0.0;
                for (int k = 0; k < 3; k++) acc = R2::add(acc, R2::mul(A[3 * r + k], Rz[3 * k + c]));
                R[3 * r + c] = acc;
            }
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) T[r + 4 * c] = R[3 * r + c];
    }
    T[0 + 12] = x[0];
    T[1 + 12] = x[1];
    T[2 + 12] = x[2];
    if (x[6] != 0) {
        const double f = R2::add(1.0, x[6]);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) T[r + 4 * c] = R2::mul(T[r + 4 * c], f);
    }
}

// out = H (3x3) * P (3x4) * T (4x4), all column-major; out may alias P.
__host__ __device__ inline void model_transform(const double* H, const double* P, const double* T, double* out)
{
    double HP[12], res[12];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++) {
            double acc = 0.0;
            for (int k = 0; k < 3; k++) acc = R2::add(acc, R2::mul(H[r + 3 * k], P[k + 3 * c]));
            HP[r + 3 * c] = acc;
        }
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++) {
            double acc = 0.0;
            for (int k = 0; k < 4; k++) acc = R2::add(acc, R2::mul(HP[r + 3 * k], T[k + 4 * c]));
            res[r + 3 * c] = acc;
        }
    for (int i = 0; i < 12; i++) out[i] = res[i];
}

// P' = H2D(x[0..3]) * P * T3D(x[4..10])
__host__ __device__ inline void model_camera_similarity_2d3d(const double* P, const double* x, double* out)
{
    double H[9], T[16];
    model_similarity_2d(x, H);
    model_similarity_3d(x + 4, T);
    model_transform(H, P, T, out);
}

// C = A * B, 3x3 column-major, separately rounded operations.
__host__ __device__ inline void mat3_mul(const double* A, const double* B, double* C)
{
    double R[9];
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) {
            double acc = 0.0;
            for (int k = 0; k < 3; k++) acc = R2::add(acc, R2::mul(A[r + 3 * k], B[k + 3 * c]));
            R[r + 3 * c] = acc;
        }
    for (int i = 0; i < 9; i++) C[i] = R[i];
}

// ModelFDCTCalibrationCorrection::getTransforms (Models/ModelFDCTCalibrationCorrection.hxx:150-203): ONE correction for the
// whole trajectory.  geom = mean principal point u, v, mean source-isocentre and source-detector distance (the model's
// pp_u, pp_v, sid, sdd); x = translation u, v, yaw, pitch, roll, delta SID, delta SDD.
//   H = H_shift * Hpp * H_roll * H_scale * Hppinv   (detector shift incl. tan(yaw|pitch) sdd; roll and SDD scale about the
//   principal point), T = diag(s, s, s, 1) with s = (sid + x5) / sid.   H 3x3, T 4x4, column-major.
__host__ __device__ inline void model_calibration_correction(const double* geom, const double* x, double* H, double* T)
{
    const double pp_u = geom[0], pp_v = geom[1], sid = geom[2], sdd = geom[3];
    const double sid_scale = R2::div(R2::add(sid, x[5]), sid), sdd_scale = R2::div(R2::add(sdd, x[6]), sdd);
    double H_roll[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, H_shift[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, H_scale[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double Hpp[9] = {1, 0, 0, 0, 1, 0, pp_u, pp_v, 1.0}, Hppinv[9] = {1, 0, 0, 0, 1, 0, -pp_u, -pp_v, 1.0};
    if (x[4] != 0) {
        double s, c;
        det_sincos(x[4], &s, &c);
        H_roll[0] = c; H_roll[3] = -s;
        H_roll[1] = s; H_roll[4] = c;
    }
    double sy, cy, sp, cp;
    det_sincos(x[2], &sy, &cy);
    det_sincos(x[3], &sp, &cp);
    H_shift[6] = R2::add(R2::add(0.0, x[0]), R2::mul(R2::div(sy, cy), sdd));
    H_shift[7] = R2::add(R2::add(0.0, x[1]), R2::mul(R2::div(sp, cp), sdd));
    H_scale[0] = sdd_scale;  // block<2,2>(0,0) *= sdd_scale on the identity
    H_scale[4] = sdd_scale;
    mat3_mul(H_shift, Hpp, H);
    mat3_mul(H, H_roll, H);
    mat3_mul(H, H_scale, H);
    mat3_mul(H, Hppinv, H);
    for (int i = 0; i < 16; i++) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    T[0] = sid_scale;
    T[5] = sid_scale;
    T[10] = sid_scale;
}

// Geometry::normalizeProjectionMatrix (LibProjectiveGeometry/ProjectionMatrix.cpp:12-18): P times -sign(det M) / |m3|... as
// the reference writes it: norm_m3 negated when det M < 0, then P * (1 / norm_m3).
__host__ __device__ inline void model_normalize(double* P)
{
    double norm_m3 = R2::root(R2::add(R2::add(R2::mul(P[2], P[2]), R2::mul(P[5], P[5])), R2::mul(P[8], P[8])));
    const double det = det3d(P[0], P[3], P[6], P[1], P[4], P[7], P[2], P[5], P[8]);
    if (det < 0) norm_m3 = -norm_m3;
    const double f = R2::div(1.0, norm_m3);
    for (int i = 0; i < 12; i++) P[i] = R2::mul(P[i], f);
}

}  // namespace eccb200
