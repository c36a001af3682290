// ecc_direct_geometry.cuh -- fp64 geometry of the direct metric (host + device): what the reference derives with Eigen on
// the host per image pair and per epipolar plane, here evaluated in registers by the thread that integrates the line.
//
// Reference (code/): LibEpipolarConsistency/EpipolarConsistencyDirect.cpp:22-62 (computeEpipolarLines), :64-113 (range and
// step of kappa), :128-186 (fan-beam weighting info); LibProjectiveGeometry/ProjectiveGeometry.hxx:188-268 (Pluecker join /
// meet / moment / direction), :333-343 (centralProjectionToPlane); ProjectionMatrix.cpp:21-24,70-76 (pseudo-inverse, camera
// centre); LibEpipolarConsistency/EpipolarConsistency.cpp:49-60 (estimateAngularRange); RectifiedFBCC.h:17-86.
// The reference takes pseudo-inverse and centre from Eigen's JacobiSVD; here (P P^T)^-1 P by cofactors and the centre from
// the 3x3 minors -- the same quantities to a few ulps of fp64 (Eigen is not available to compare bit for bit: this part of
// the direct path is "parity unpinned").
#pragma once
#include <cmath>

namespace eccb200 {

// What a view contributes: A = (P^+)^T (3x4 col-major), centre C (C[3] = 1 for a finite source), P itself.
struct DirectView {
    double A[12];
    double C[4];
    double P[12];
};

// What a pair contributes to every one of its epipolar planes.
struct DirectPair {
    double E0[4], E90[4];  // epipolar planes at 0 and 90 degrees to the origin, Hessian normal form
    double kappa0, dkappa; // first angle and step
    int n_lines;
    int i, j;
    // fan-beam consistency only: rectifying homographies of the two views (3x3 col-major), baseline direction, detector plane
    double H0[9], H1[9];
    double dir[3];
    double E[4];
};

// LinePerspectivity + FBCC_weighting_info (RectifiedFBCC.h:17-24,82-87): 8 floats per line and image, the layout the
// reference's kernel reads.
struct FbccInfo {
    float a, b, c, d;
    float t_prime_ak, d_l_kappa_C_sq;
    float dummy0, dummy1;
};

__host__ __device__ inline void direct_view(const double* P, DirectView& V)
{
    for (int k = 0; k < 12; k++) V.P[k] = P[k];
    double G[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += P[r + 3 * k] * P[c + 3 * k];
            G[r + 3 * c] = s;
        }
    const double c00 = G[4] * G[8] - G[7] * G[5], c01 = G[7] * G[2] - G[1] * G[8], c02 = G[1] * G[5] - G[4] * G[2];
    const double det = G[0] * c00 + G[3] * c01 + G[6] * c02;
    double Gi[9];
    Gi[0] = c00 / det;
    Gi[1] = c01 / det;
    Gi[2] = c02 / det;
    Gi[3] = (G[6] * G[5] - G[3] * G[8]) / det;
    Gi[4] = (G[0] * G[8] - G[6] * G[2]) / det;
    Gi[5] = (G[3] * G[2] - G[0] * G[5]) / det;
    Gi[6] = (G[3] * G[7] - G[6] * G[4]) / det;
    Gi[7] = (G[6] * G[1] - G[0] * G[7]) / det;
    Gi[8] = (G[0] * G[4] - G[3] * G[1]) / det;
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++) V.A[r + 3 * c] = Gi[r] * P[3 * c] + Gi[r + 3] * P[1 + 3 * c] + Gi[r + 6] * P[2 + 3 * c];
    // null vector: signed 3x3 minors
    double m[4];
    for (int k = 0; k < 4; k++) {
        int cc[3], q = 0;
        for (int j = 0; j < 4; j++)
            if (j != k) cc[q++] = j;
        const double* a = P + 3 * cc[0];
        const double* b = P + 3 * cc[1];
        const double* c = P + 3 * cc[2];
        m[k] = a[0] * (b[1] * c[2] - c[1] * b[2]) - b[0] * (a[1] * c[2] - c[1] * a[2]) + c[0] * (a[1] * b[2] - b[1] * a[2]);
        if (k & 1) m[k] = -m[k];
    }
    const double scale = sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2] + m[3] * m[3]);
    for (int k = 0; k < 4; k++) V.C[k] = m[k] / scale;  // unit length like a singular vector
    if (V.C[3] < -1e-12 || V.C[3] > 1e-12) {            // getCameraCenter, ProjectionMatrix.cpp:70-76
        const double w = V.C[3];
        for (int k = 0; k < 4; k++) V.C[k] /= w;
    }
}

__host__ __device__ inline void join_points(const double* A, const double* B, double* L)
{
    L[0] = A[0] * B[1] - A[1] * B[0];
    L[1] = A[0] * B[2] - A[2] * B[0];
    L[2] = A[0] * B[3] - A[3] * B[0];
    L[3] = A[1] * B[2] - A[2] * B[1];
    L[4] = A[1] * B[3] - A[3] * B[1];
    L[5] = A[2] * B[3] - A[3] * B[2];
}
__host__ __device__ inline void meet_planes(const double* A, const double* B, double* L)
{
    L[0] = A[2] * B[3] - A[3] * B[2];
    L[1] = A[3] * B[1] - A[1] * B[3];
    L[2] = A[1] * B[2] - A[2] * B[1];
    L[3] = A[0] * B[3] - A[3] * B[0];
    L[4] = A[2] * B[0] - A[0] * B[2];
    L[5] = A[0] * B[1] - A[1] * B[0];
}
__host__ __device__ inline void join_line_point(const double* L, const double* X, double* E)
{
    E[0] = +X[1] * L[5] - X[2] * L[4] + X[3] * L[3];
    E[1] = -X[0] * L[5] + X[2] * L[2] - X[3] * L[1];
    E[2] = +X[0] * L[4] - X[1] * L[2] + X[3] * L[0];
    E[3] = -X[0] * L[3] + X[1] * L[1] - X[2] * L[0];
}
__host__ __device__ inline void meet_line_plane(const double* L, const double* P, double* X)
{
    X[0] = -P[1] * L[0] - P[2] * L[1] - P[3] * L[2];
    X[1] = +P[0] * L[0] - P[2] * L[3] - P[3] * L[4];
    X[2] = +P[0] * L[1] + P[1] * L[3] - P[3] * L[5];
    X[3] = +P[0] * L[2] + P[1] * L[4] + P[2] * L[5];
}
__host__ __device__ inline bool dehomogenize3(double* X)  // ProjectiveGeometry.hxx:75-91
{
    if (X[3] > 1e-12 || X[3] < -1e-12) {
        const double w = X[3];
        for (int k = 0; k < 4; k++) X[k] /= w;
        return true;
    }
    X[3] = 0;
    const double n = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
    for (int k = 0; k < 4; k++) X[k] /= n;
    return false;
}
__host__ __device__ inline bool dehomogenize2(double* x)  // ProjectiveGeometry.hxx:38-53
{
    if (x[2] > 1e-11 || x[2] < -1e-11) {
        const double w = x[2];
        for (int k = 0; k < 3; k++) x[k] /= w;
        return true;
    }
    x[2] = 0;
    const double n = sqrt(x[0] * x[0] + x[1] * x[1]);
    for (int k = 0; k < 3; k++) x[k] /= n;
    return false;
}

// Range and step of kappa and the number of epipolar planes of a pair (EpipolarConsistencyDirect.cpp:84-113,
// EpipolarConsistency.cpp:49-60); the two planes the pencil is spanned by (:26-40); for fan-beam consistency the rectifying
// homographies H = P_E T(C, E) P^+ of both views (:128-145).
__host__ __device__ inline void direct_pair(const DirectView& V0, const DirectView& V1, double radius, double dkappa, int n_u, int n_v,
                                            bool fbcc, DirectPair& R)
{
    const double kPi = 3.14159265358979323846;
    double B[6];
    join_points(V0.C, V1.C, B);
    const double origin[4] = {0, 0, 0, 1};
    join_line_point(B, origin, R.E0);
    join_line_point(B, R.E0, R.E90);
    const double n0 = sqrt(R.E0[0] * R.E0[0] + R.E0[1] * R.E0[1] + R.E0[2] * R.E0[2]);
    const double n90 = sqrt(R.E90[0] * R.E90[0] + R.E90[1] * R.E90[1] + R.E90[2] * R.E90[2]);
    for (int k = 0; k < 4; k++) { R.E0[k] /= n0; R.E90[k] /= n90; }
    const double dir[3] = {-B[2], -B[4], -B[5]}, mom[3] = {B[3], -B[1], B[0]};
    const double nd = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    const double nm = sqrt(mom[0] * mom[0] + mom[1] * mom[1] + mom[2] * mom[2]);
    const double baseline_dist = nm / nd;
    double lo = -0.5 * kPi, hi = 0.5 * kPi;
    if (!(baseline_dist <= radius)) {
        const double kmax = fabs(asin(radius / baseline_dist));
        lo = -kmax;
        hi = kmax;
    }
    if (dkappa <= 0) {
        const double diag = sqrt((double)(n_u * n_u + n_v * n_v));
        dkappa = 0.5 * (hi - lo) / diag;
    }
    R.kappa0 = lo;
    R.dkappa = dkappa;
    const double planes = (hi - lo) / dkappa;
    R.n_lines = planes < 2147483647.0 ? (int)planes : 2147483647;  // callers refuse absurd counts (kMaxPlanesPerPair)
    for (int k = 0; k < 3; k++) R.dir[k] = dir[k];
    for (int k = 0; k < 9; k++) R.H0[k] = R.H1[k] = 0;
    for (int k = 0; k < 4; k++) R.E[k] = 0;
    if (!fbcc) return;
    // virtual detector plane through the origin, spanned by baseline direction and moment
    const double U[3] = {dir[0] / nd, dir[1] / nd, dir[2] / nd}, Vv[3] = {mom[0] / nm, mom[1] / nm, mom[2] / nm};
    R.E[0] = U[1] * Vv[2] - U[2] * Vv[1];
    R.E[1] = U[2] * Vv[0] - U[0] * Vv[2];
    R.E[2] = U[0] * Vv[1] - U[1] * Vv[0];
    R.E[3] = 0;
    const double PE[12] = {U[0], Vv[0], 0, U[1], Vv[1], 0, U[2], Vv[2], 0, 0, 0, 1.0};  // 3x4 col-major, pixel spacing 1
    for (int view = 0; view < 2; view++) {
        const DirectView& W = view ? V1 : V0;
        const double* C = W.C;
        const double* E = R.E;
        const double T[16] = {  // centralProjectionToPlane(C, E), col-major
            +C[1] * E[1] + C[2] * E[2] + C[3] * E[3], -C[1] * E[0], -C[2] * E[0], -C[3] * E[0],
            -C[0] * E[1], +C[0] * E[0] + C[2] * E[2] + C[3] * E[3], -C[2] * E[1], -C[3] * E[1],
            -C[0] * E[2], -C[1] * E[2], +C[0] * E[0] + C[3] * E[3] + C[1] * E[1], -C[3] * E[2],
            -C[0] * E[3], -C[1] * E[3], -C[2] * E[3], +C[0] * E[0] + C[1] * E[1] + C[2] * E[2]};
        double PT[12];  // P_E * T, 3x4
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 3; r++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += PE[r + 3 * k] * T[k + 4 * c];
                PT[r + 3 * c] = s;
            }
        double* H = view ? R.H1 : R.H0;  // (P_E T) P^+, P^+ = A^T (4x3)
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += PT[r + 3 * k] * W.A[c + 3 * k];
                H[r + 3 * c] = s;
            }
    }
}

// Corresponding epipolar lines of the plane at angle kappa (EpipolarConsistencyDirect.cpp:47-59), Hessian normal form, fp32.
__host__ __device__ inline void direct_lines(const DirectView& V0, const DirectView& V1, const DirectPair& R, double kappa, float* l0, float* l1)
{
    const double ck = cos(kappa), sk = sin(kappa);
    double Ek[4];
    for (int k = 0; k < 4; k++) Ek[k] = ck * R.E0[k] + sk * R.E90[k];
    for (int view = 0; view < 2; view++) {
        const double* A = view ? V1.A : V0.A;
        double l[3];
        for (int r = 0; r < 3; r++) l[r] = A[r] * Ek[0] + A[r + 3] * Ek[1] + A[r + 6] * Ek[2] + A[r + 9] * Ek[3];
        const double n = sqrt(l[0] * l[0] + l[1] * l[1]);
        float* out = view ? l1 : l0;
        for (int r = 0; r < 3; r++) out[r] = (float)(l[r] / n);
    }
}

// fp32 evaluation of the perspectivity exactly as the host code of the reference does it (separately rounded operations).
__host__ __device__ inline float perspectivity_transform(const FbccInfo& f, float t)
{
#ifdef __CUDA_ARCH__
    return __fdiv_rn(__fadd_rn(__fmul_rn(f.a, t), f.b), __fadd_rn(__fmul_rn(f.c, t), f.d));
#else
    return (f.a * t + f.b) / (f.c * t + f.d);
#endif
}

// Fan-beam weighting info of one epipolar line pair (EpipolarConsistencyDirect.cpp:147-186).
__host__ __device__ inline void direct_fbcc(const DirectView& V0, const DirectView& V1, const DirectPair& R, const float* l0f, const float* l1f,
                                            FbccInfo& f0, FbccInfo& f1)
{
    const double l0[3] = {l0f[0], l0f[1], l0f[2]}, l1[3] = {l1f[0], l1f[1], l1f[2]};
    double Ek[4];  // P0^T l0
    for (int c = 0; c < 4; c++) Ek[c] = V0.P[3 * c] * l0[0] + V0.P[1 + 3 * c] * l0[1] + V0.P[2 + 3 * c] * l0[2];
    for (int view = 0; view < 2; view++) {
        const DirectView& W = view ? V1 : V0;
        const double* l = view ? l1 : l0;
        const double* H = view ? R.H1 : R.H0;
        const double EB[4] = {R.dir[0], R.dir[1], R.dir[2], -(R.dir[0] * W.C[0] + R.dir[1] * W.C[1] + R.dir[2] * W.C[2])};
        double Lm[6], Ak[4];
        meet_planes(EB, Ek, Lm);
        meet_line_plane(Lm, R.E, Ak);
        dehomogenize3(Ak);
        double dist = 0;
        for (int k = 0; k < 4; k++) dist += (Ak[k] - W.C[k]) * (Ak[k] - W.C[k]);
        const float d_px = (float)(sqrt(dist) / 1.0);
        double ak[3];
        for (int r = 0; r < 3; r++) ak[r] = W.P[r] * Ak[0] + W.P[r + 3] * Ak[1] + W.P[r + 6] * Ak[2] + W.P[r + 9] * Ak[3];
        dehomogenize2(ak);
        FbccInfo& f = view ? f1 : f0;
        // LinePerspectivity(H, l), RectifiedFBCC.h:41-50 (H(r,c) = H[r + 3c])
        f.a = (float)(H[0] * l[1] - H[1] * l[0]);
        f.b = (float)(H[6] - H[0] * l[0] * l[2]);
        f.c = (float)(H[2] * l[1] - H[5] * l[0]);
        f.d = (float)(H[8] - H[2] * l[0] * l[2] - H[5] * l[1] * l[2]);
#ifdef __CUDA_ARCH__
        const float detp = __fadd_rn(__fmul_rn(f.a, f.d), -__fmul_rn(f.b, f.c));
#else
        const float detp = f.a * f.d - f.b * f.c;
#endif
        if (detp < 0) { f.a *= -1; f.b *= -1; }
        const double t_ak = l[1] * ak[0] / ak[2] - l[0] * ak[1] / ak[2];  // project_to_line
        f.t_prime_ak = perspectivity_transform(f, (float)t_ak);
#ifdef __CUDA_ARCH__
        f.d_l_kappa_C_sq = __fmul_rn(d_px, d_px);
#else
        f.d_l_kappa_C_sq = d_px * d_px;
#endif
        f.dummy0 = f.dummy1 = 0.f;
    }
}

}  // namespace eccb200
