// ecc_geometry.cuh -- geometry shared by the pair kernels and the host side of libecc_b200.
//
// What is computed follows the reference (aaichert/EpipolarConsistency, code/...):
//   pair enumeration      LibEpipolarConsistency/EpipolarConsistencyCommon.hxx:52-79   (get_ij)
//   K0/K1 record          EpipolarConsistencyCommon.hxx:82-149                         (computeK01)
//   line -> dtr sample    EpipolarConsistencyCommon.hxx:152-171                        (lineToSampleDtr)
//   pinv^T / source pos.  LibUtilsCuda/culaut/xprojectionmatrix.hxx:20-52,93-105
// How it is computed is ours: closed-form pair index, cofactor linear algebra in fp64, everything
// in registers and fused into the consuming kernel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace eccb200 {

// The reference's device code spells pi as this float literal (3.14159274f).
#define ECC_PI_F 3.14159265359f

// Start index of row i in the row-major enumeration of the strict upper triangle of an n x n matrix.
__host__ __device__ inline long long pair_row_start(long long i, long long n)
{
    return i * (2 * n - i - 1) / 2;
}

// k -> (i,j), i<j<n, same order as the reference's get_ij but O(1) instead of O(n).
__host__ __device__ inline void pair_from_index(long long k, int n, int& i, int& j)
{
    const double b = 2.0 * n - 1.0;
    double disc = b * b - 8.0 * (double)k;
    if (disc < 0) disc = 0;
    long long row = (long long)((b - sqrt(disc)) * 0.5);
    if (row < 0) row = 0;
    if (row > n - 2) row = n - 2;
    while (row + 1 <= n - 2 && pair_row_start(row + 1, n) <= k) ++row;
    while (row > 0 && pair_row_start(row, n) > k) --row;
    i = (int)row;
    j = (int)(k - pair_row_start(row, n)) + i + 1;
}

__host__ __device__ inline long long pair_to_index(int i, int j, int n)
{
    return pair_row_start(i, n) + (j - i - 1);
}

// What a pair needs to sample its two Radon intermediates.
struct PairMaps {
    float k0[6];     // 3x2 col-major: (cos k, sin k) -> epipolar line in view 0, image-centre origin
    float k1[6];     // same for view 1
    float baseline;  // distance of the baseline to the origin (K0[6] in the reference)
    float dkappa;    // K1[6]
    float kappa_max; // K1[7]
};

// Builds the PairMaps from the two source positions (C, 4 floats, C[3]==1) and the two
// pseudo-inverse transposes (3x4 col-major floats).
__host__ __device__ inline void make_pair_maps(float half_nu, float half_nv, const float* C0,
                                               const float* C1, const float* A0, const float* A1,
                                               float radius, float num_samples, float dkappa,
                                               bool same_view, PairMaps& pm)
{
    if (same_view) {  // the reference zeroes the record when both views are the same object
        for (int q = 0; q < 6; q++) pm.k0[q] = pm.k1[q] = 0.f;
        pm.baseline = 0.f;
        pm.kappa_max = 0.f;
        pm.dkappa = 0.f;
        return;
    }
    // Pluecker coordinates of the baseline
    const float b01 = C0[0] * C1[1] - C0[1] * C1[0];
    const float b02 = C0[0] * C1[2] - C0[2] * C1[0];
    const float b03 = C0[0] * C1[3] - C0[3] * C1[0];
    const float b12 = C0[1] * C1[2] - C0[2] * C1[1];
    const float b13 = C0[1] * C1[3] - C0[3] * C1[1];
    const float b23 = C0[2] * C1[3] - C0[3] * C1[2];
    const float mom = sqrtf(b12 * b12 + b02 * b02 + b01 * b01);
    const float dir = sqrtf(b03 * b03 + b13 * b13 + b23 * b23);
    // the pencil is spanned by the plane through the origin and the plane farthest from it
    float E[8];
    E[0] = b12 / mom;
    E[1] = -b02 / mom;
    E[2] = b01 / mom;
    E[3] = 0.f;
    E[4] = (-b01 * b13 - b02 * b23) / (mom * dir);
    E[5] = (b01 * b03 - b12 * b23) / (mom * dir);
    E[6] = (b02 * b03 + b12 * b13) / (mom * dir);
    E[7] = -mom / dir;
#pragma unroll
    for (int v = 0; v < 2; v++) {
        const float* A = v ? A1 : A0;
        float* K = v ? pm.k1 : pm.k0;
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int r = 0; r < 3; r++) {
                float s = 0.f;
#pragma unroll
                for (int q = 0; q < 4; q++) s += A[r + 3 * q] * E[q + 4 * c];
                K[r + 3 * c] = s;
            }
        // image-centre origin, unit normal for the kappa=0 line
        K[2] += half_nu * K[0] + half_nv * K[1];
        K[5] += half_nu * K[3] + half_nv * K[4];
        const float len = sqrtf(K[0] * K[0] + K[1] * K[1]);
#pragma unroll
        for (int q = 0; q < 6; q++) K[q] /= len;
    }
    pm.baseline = mom / dir;
    pm.kappa_max = (pm.baseline <= radius) ? 0.5f * ECC_PI_F : asinf(radius / pm.baseline);
    pm.dkappa = (dkappa <= 0.f) ? 2.f * pm.kappa_max / num_samples : dkappa;
}

// kappa of sample m: (m + 1/2) dkappa, in the form nvcc gives the reference's
// "dkappa*0.5f+dkappa*idx_y" (EpipolarConsistencyRadonIntermediate.cu:194,260): one fused multiply-add.
__host__ __device__ inline float kappa_of_sample(float dkappa, int m)
{
    return fmaf(dkappa, (float)m, dkappa * 0.5f);
}

// Number of kappa samples a pair takes: m = 0,1,... while kappa_of_sample(m) < kappa_max and m < cap.
__host__ __device__ inline int pair_num_samples(const PairMaps& pm, int cap)
{
    if (!(pm.dkappa > 0.f)) return 0;
    // first guess, then settle with the exact float test the kernels use
    int m = (int)(pm.kappa_max / pm.dkappa);
    if (m > cap) m = cap;
    if (m < 0) m = 0;
    while (m < cap && kappa_of_sample(pm.dkappa, m) < pm.kappa_max) ++m;
    while (m > 0 && !(kappa_of_sample(pm.dkappa, m - 1) < pm.kappa_max)) --m;
    return m;
}

// ---- fp64 3x4 matrix helpers (host + device) ----------------------------------------------------
__host__ __device__ inline double det3d(double a, double b, double c, double d, double e, double f,
                                        double g, double h, double i)
{
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// A = (P P^T)^-1 P as 3x4 col-major floats; C = null vector of P with C[3]==1.
__host__ __device__ inline void derive_view(const double* P, float* A, float* C)
{
    double G[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += P[r + 3 * k] * P[c + 3 * k];
            G[r + 3 * c] = s;
        }
    const double det = det3d(G[0], G[3], G[6], G[1], G[4], G[7], G[2], G[5], G[8]);
    const double id = 1.0 / det;
    double Gi[9];
    Gi[0] = (G[4] * G[8] - G[7] * G[5]) * id;
    Gi[3] = -(G[3] * G[8] - G[6] * G[5]) * id;
    Gi[6] = (G[3] * G[7] - G[6] * G[4]) * id;
    Gi[1] = -(G[1] * G[8] - G[7] * G[2]) * id;
    Gi[4] = (G[0] * G[8] - G[6] * G[2]) * id;
    Gi[7] = -(G[0] * G[7] - G[6] * G[1]) * id;
    Gi[2] = (G[1] * G[5] - G[4] * G[2]) * id;
    Gi[5] = -(G[0] * G[5] - G[3] * G[2]) * id;
    Gi[8] = (G[0] * G[4] - G[3] * G[1]) * id;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += Gi[r + 3 * k] * P[k + 3 * c];
            A[r + 3 * c] = (float)s;
        }
    double m[4];
    for (int k = 0; k < 4; k++) {
        int c[3], q = 0;
        for (int j = 0; j < 4; j++)
            if (j != k) c[q++] = j;
        m[k] = det3d(P[0 + 3 * c[0]], P[0 + 3 * c[1]], P[0 + 3 * c[2]], P[1 + 3 * c[0]],
                     P[1 + 3 * c[1]], P[1 + 3 * c[2]], P[2 + 3 * c[0]], P[2 + 3 * c[1]],
                     P[2 + 3 * c[2]]);
        if (k & 1) m[k] = -m[k];
    }
    for (int k = 0; k < 4; k++) C[k] = (float)(m[k] / m[3]);
}

}  // namespace eccb200
