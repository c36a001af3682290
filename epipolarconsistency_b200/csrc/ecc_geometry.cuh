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

// K0[7] of the reference's record: the angle under which the two sources see each other from the reference point
// (EpipolarConsistencyCommon.hxx:139-140); unused downstream, kept for the launcher-compatible K01 output.
__host__ __device__ inline float pair_angle(const float* C0, const float* C1)
{
    const float b01 = C0[0] * C1[1] - C0[1] * C1[0];
    const float b02 = C0[0] * C1[2] - C0[2] * C1[0];
    const float b03 = C0[0] * C1[3] - C0[3] * C1[0];
    const float b12 = C0[1] * C1[2] - C0[2] * C1[1];
    const float b13 = C0[1] * C1[3] - C0[3] * C1[1];
    const float b23 = C0[2] * C1[3] - C0[3] * C1[2];
    const float s2 = sqrtf(b12 * b12 + b02 * b02 + b01 * b01);
    const float s3 = sqrtf(b03 * b03 + b13 * b13 + b23 * b23);
    return -2.0f * atan2f(-0.5f * s3, s2 / s3);
}

// kappa of sample m: (m + 1/2) dkappa, with the roundings of the reference's compiled kernel.  Its source reads
// "dkappa*0.5f+dkappa*idx_y" (EpipolarConsistencyRadonIntermediate.cu:194,260); nvcc 12.9 turns that into
//     FMUL t = dkappa * float(idx_y);   FFMA kappa = dkappa * 0.5 + t
// (cuobjdump of the sm_100 build of the reference's own .cu file, kernelEpipolarCosistency at 0x0390/0x03a0): the product with the sample index is
// rounded on its own.  Round 1 had guessed the other fusion, fma(dkappa, m, dkappa/2), which is an ulp off for about a
// third of the samples -- invisible in the sums, but enough to move a texture coordinate across a 1/256 weight step where
// an intermediate is steep (found at BASELINE size: 2 of 1830 pairs off by 0.5 % in ONE sample each, tools/pair_outlier_probe.py).
__host__ __device__ inline float kappa_of_sample(float dkappa, int m)
{
#ifdef __CUDA_ARCH__
    return fmaf(dkappa, 0.5f, __fmul_rn(dkappa, (float)m));
#else
    const float t = dkappa * (float)m;
    return fmaf(dkappa, 0.5f, t);
#endif
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

// ---- pseudo-inverse transpose and source position, bit-faithful to the reference ------------------------
// The reference derives both on the host in fp64 with a Householder QR (culaut::xsqqr / xgeinv,
// LibUtilsCuda/culaut/xgeinv.hxx:40-168, used by xprojectionmatrix.hxx:20-52,93-105) and rounds to fp32.  The
// metric is a sum of squared interpolation residuals, so a one-ulp change of these floats can move a pair value
// at the 1e-3 level (DESIGN.md "conditioning"); we therefore reproduce the same sequence of IEEE operations.
// R2 keeps every product and sum separately rounded (no fused multiply-add), on the device via the _rn intrinsics,
// so host and device give the same bits as the reference's host code.
struct R2 {
    __host__ __device__ static inline double mul(double a, double b)
    {
#ifdef __CUDA_ARCH__
        return __dmul_rn(a, b);
#else
        return a * b;
#endif
    }
    __host__ __device__ static inline double add(double a, double b)
    {
#ifdef __CUDA_ARCH__
        return __dadd_rn(a, b);
#else
        return a + b;
#endif
    }
    __host__ __device__ static inline double sub(double a, double b) { return add(a, -b); }
    __host__ __device__ static inline double div(double a, double b)
    {
#ifdef __CUDA_ARCH__
        return __ddiv_rn(a, b);
#else
        return a / b;
#endif
    }
    __host__ __device__ static inline double root(double a)
    {
#ifdef __CUDA_ARCH__
        return __dsqrt_rn(a);
#else
        return sqrt(a);
#endif
    }
};

// Householder QR of a column-major N x N matrix, A = Q R.  On return the strict upper triangle of A holds R's,
// diag holds R's diagonal, and Q is explicit.  Quirk kept from the reference: the column scale is a running
// maximum over all columns processed so far, and the last column is processed too.
template <int N>
__host__ __device__ inline void householder_qr(double* A, double* Q, double* diag)
{
    double coef[N];
    double scale = 0.0;
    for (int k = 0; k < N; k++) {
        double* col = A + N * k;
        for (int i = k; i < N; i++) {
            const double m = fabs(col[i]);
            if (scale < m) scale = m;
        }
        if (scale == 0.0) {
            coef[k] = diag[k] = 0.0;
            continue;
        }
        for (int i = k; i < N; i++) col[i] = R2::div(col[i], scale);
        double nrm = 0.0;
        for (int i = k; i < N; i++) nrm = R2::add(nrm, R2::mul(col[i], col[i]));
        const double sigma = col[k] > 0.0 ? R2::root(nrm) : -R2::root(nrm);
        col[k] = R2::add(col[k], sigma);
        coef[k] = R2::mul(sigma, col[k]);
        diag[k] = R2::mul(-scale, sigma);
        for (int j = k + 1; j < N; j++) {
            double* other = A + N * j;
            double dot = 0.0;
            for (int i = k; i < N; i++) dot = R2::add(dot, R2::mul(col[i], other[i]));
            const double tau = R2::div(dot, coef[k]);
            for (int i = k; i < N; i++) other[i] = R2::sub(other[i], R2::mul(tau, col[i]));
        }
    }
    diag[N - 1] = -diag[N - 1];
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) Q[i + N * j] = (i == j) ? 1.0 : 0.0;
    for (int k = 0; k < N - 1; k++) {
        if (coef[k] == 0.0) continue;
        const double* col = A + N * k;
        for (int j = 0; j < N; j++) {
            double dot = 0.0;
            for (int i = k; i < N; i++) dot = R2::add(dot, R2::mul(col[i], Q[j + N * i]));
            dot = R2::div(dot, coef[k]);
            for (int i = k; i < N; i++) Q[j + N * i] = R2::sub(Q[j + N * i], R2::mul(dot, col[i]));
        }
    }
}

// A = (P P^T)^-1 P as 3x4 col-major floats; C = null vector of P with C[3]==1.
__host__ __device__ inline void derive_view(const double* P, float* A, float* C)
{
    // Gram matrix P P^T (symmetric 3x3); each entry is a left-to-right sum over the four columns
    double G[9];
    for (int r = 0; r < 3; r++)
        for (int c = r; c < 3; c++) {
            double s = R2::mul(P[r], P[c]);
            for (int k = 1; k < 4; k++) s = R2::add(s, R2::mul(P[r + 3 * k], P[c + 3 * k]));
            G[c + 3 * r] = G[r + 3 * c] = s;
        }
    // inverse through QR: solve R x = Q^T e_i by back substitution, one unit vector at a time
    double Qm[9], dg[3], Gi[9];
    householder_qr<3>(G, Qm, dg);
    for (int i = 0; i < 3; i++) {
        double rhs[3], x[3];
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int q = 0; q < 3; q++) s = R2::add(s, R2::mul(q == i ? 1.0 : 0.0, Qm[q + 3 * j]));
            rhs[j] = s;
        }
        x[2] = R2::div(rhs[2], dg[2]);
        for (int r = 1; r >= 0; r--) {
            double v = rhs[r];
            for (int j = r + 1; j < 3; j++) v = R2::sub(v, R2::mul(G[r + 3 * j], x[j]));
            x[r] = R2::div(v, dg[r]);
        }
        for (int r = 0; r < 3; r++) Gi[r + 3 * i] = x[r];
    }
    // PinvT(r,c) = sum_k P(k,c) * Gi(k,r)
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++) {
            const double* p = P + 3 * c;
            const double* g = Gi + 3 * r;
            A[r + 3 * c] = (float)R2::add(R2::add(R2::mul(p[0], g[0]), R2::mul(p[1], g[1])), R2::mul(p[2], g[2]));
        }
    // source position: last column of Q in the QR decomposition of [P^T | 0] (4x4), de-homogenised
    double M4[16], Q4[16], d4[4];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++) M4[j + 4 * i] = P[i + 3 * j];
    for (int j = 0; j < 4; j++) M4[j + 12] = 0.0;
    householder_qr<4>(M4, Q4, d4);
    for (int i = 0; i < 4; i++) C[i] = (float)R2::div(Q4[i + 12], Q4[15]);
}

// Automatic object radius from one projection matrix (what Metric::getObjectRadius derives from the first
// matrix of the set): focal lengths from the rows of the left 3x3 block, the larger half field of view of the
// n_u x n_v detector, times the source's distance to the origin.
__host__ __device__ inline double object_radius_from_view(const double* P, int n_u, int n_v)
{
    const double m1[3] = {P[0], P[3], P[6]}, m2[3] = {P[1], P[4], P[7]}, m3[3] = {P[2], P[5], P[8]};
    double U[3] = {m3[1] * m2[2] - m3[2] * m2[1], m3[2] * m2[0] - m3[0] * m2[2], m3[0] * m2[1] - m3[1] * m2[0]};
    double V[3] = {m3[1] * m1[2] - m3[2] * m1[1], m3[2] * m1[0] - m3[0] * m1[2], m3[0] * m1[1] - m3[1] * m1[0]};
    const double nU = sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]);
    const double nV = sqrt(V[0] * V[0] + V[1] * V[1] + V[2] * V[2]);
    for (int k = 0; k < 3; k++) { U[k] /= nU; V[k] /= nV; }
    const double t1[3] = {V[1] * m3[2] - V[2] * m3[1], V[2] * m3[0] - V[0] * m3[2], V[0] * m3[1] - V[1] * m3[0]};
    const double t2[3] = {U[1] * m3[2] - U[2] * m3[1], U[2] * m3[0] - U[0] * m3[2], U[0] * m3[1] - U[1] * m3[0]};
    const double fu = m1[0] * t1[0] + m1[1] * t1[1] + m1[2] * t1[2];
    const double fv = m2[0] * t2[0] + m2[1] * t2[1] + m2[2] * t2[2];
    const double fov = fmax(fabs(atan(0.5 * n_u / fu)), fabs(atan(0.5 * n_v / fv)));
    double m[4];
    for (int k = 0; k < 4; k++) {
        int c[3], q = 0;
        for (int j = 0; j < 4; j++)
            if (j != k) c[q++] = j;
        m[k] = det3d(P[0 + 3 * c[0]], P[0 + 3 * c[1]], P[0 + 3 * c[2]], P[1 + 3 * c[0]], P[1 + 3 * c[1]],
                     P[1 + 3 * c[2]], P[2 + 3 * c[0]], P[2 + 3 * c[1]], P[2 + 3 * c[2]]);
        if (k & 1) m[k] = -m[k];
    }
    const double C0 = m[0] / m[3], C1 = m[1] / m[3], C2 = m[2] / m[3];
    return sin(fov) * sqrt(C0 * C0 + C1 * C1 + C2 * C2);
}

}  // namespace eccb200
