// ecc_oracle.cpp -- CPU oracle (test infrastructure, see ecc_oracle.h for the rules of use).
//
// Every function restates one piece of the reference (aaichert/EpipolarConsistency); the file:line
// it follows is given above each function, relative to the reference's code/ directory.  The code
// is written from the algorithm description, not copied: structure, names and the linear-algebra
// routes (cofactors instead of Householder QR) are ours.  Where the reference's device code would
// be contracted to FMA by nvcc we say so with an explicit fmaf(); everything else is plain fp32
// (build with -ffp-contract=off, see Makefile).
#include "ecc_oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// The reference's device code uses this literal (EpipolarConsistencyCommon.hxx:141,155,
// RadonIntermediate.cu:8); as a float it is 3.14159274f.
const float kPiF = 3.14159265359f;
const double kPiD = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------------
// Texture-unit model.  CUDA linear filtering (LibUtilsCuda/CudaBindlessTexture.cpp:36-40 sets
// linear + clamp): texel centre i sits at coordinate i+0.5, so the fetch position is xb = x-0.5,
// i = floor(xb), weight = frac(xb); indices are clamped to the edge.  TEX8 stores the weight in
// 1.8 fixed point.  Rounding rule calibrated on a B200 (tools/tex_probe.cu, profiles/tex_probe_r01.txt):
// the position is rounded to the nearest 1/256 (half up), which may carry into the texel index; the
// 2-D weights are derived from the two 1-D weights as described in bilinear().
// ---------------------------------------------------------------------------------------------
struct Tap {
    int i0, i1;
    float w;  // weight of i1 (EXACT: full precision; TEX8: a multiple of 1/256)
    int wq;   // TEX8: the same weight as an integer number of 1/256
};

inline Tap make_tap(float x, int n, int interp)
{
    Tap t;
    float xb = x - 0.5f;
    float fl;
    if (interp == ORACLE_INTERP_TEX8) {
        float q = floorf(xb * 256.0f + 0.5f);  // position in 1/256 texels, round to nearest
        fl = floorf(q * (1.0f / 256.0f));
        t.wq = (int)(q - fl * 256.0f);
        t.w = (float)t.wq * (1.0f / 256.0f);
    } else {
        fl = floorf(xb);
        t.w = xb - fl;
        t.wq = 0;
    }
    // Guard the float->int conversion against absurd coordinates (clamp makes them equivalent).
    if (fl < -2.0f) fl = -2.0f;
    if (fl > (float)n) fl = (float)n;
    int i = (int)fl;
    t.i0 = std::min(std::max(i, 0), n - 1);
    t.i1 = std::min(std::max(i + 1, 0), n - 1);
    return t;
}

inline float bilinear(const float* img, int w, int h, float x, float y, int interp)
{
    const Tap tx = make_tap(x, w, interp);
    const Tap ty = make_tap(y, h, interp);
    const float* r0 = img + (size_t)ty.i0 * w;
    const float* r1 = img + (size_t)ty.i1 * w;
    if (interp == ORACLE_INTERP_TEX8) {
        // Measured on B200 (tools/tex_probe.cu, profiles/tex_probe_r01.txt): the four weights are multiples of
        // 1/256 whose row and column sums are the two 1-D weights; the corner product a*b is itself rounded to
        // 1/256 and the other three follow by subtraction (w11 = rn(a*b), w10 = a - w11, w01 = b - w11,
        // w00 = 1 - a - b + w11).  This is NOT the outer product of the quantised 1-D weights.
        const int a = tx.wq, b = ty.wq;
        const int w11 = (a * b + 128) >> 8;
        const int w10 = a - w11, w01 = b - w11, w00 = 256 - a - b + w11;
        return ((float)w00 * r0[tx.i0] + (float)w10 * r0[tx.i1] + (float)w01 * r1[tx.i0] + (float)w11 * r1[tx.i1]) *
               (1.0f / 256.0f);
    }
    const float a = tx.w, b = ty.w;
    // CUDA programming guide, "Linear Filtering":
    // (1-a)(1-b)T[i,j] + a(1-b)T[i+1,j] + (1-a)b T[i,j+1] + ab T[i+1,j+1]
    return (1.0f - a) * (1.0f - b) * r0[tx.i0] + a * (1.0f - b) * r0[tx.i1] +
           (1.0f - a) * b * r1[tx.i0] + a * b * r1[tx.i1];
}

inline void sort_four(float* v)
{
    // any correct sort gives the same middle pair (RadonIntermediate.cu:18-28 uses bubble sort)
    std::sort(v, v + 4);
}

// One Radon bin.  RadonIntermediate.cu:31-143.  count_only: just return the sample count.
inline float radon_bin(const float* img, int n_ui, int n_vi, int ix, int iy, int n_alpha, int n_t,
                       bool derivative, int post, int interp, double* n_samples)
{
    const float n_u = (float)n_ui, n_v = (float)n_vi;
    // :46-52 bin -> (alpha, tau)
    const float x_rel = ix / (float)n_alpha - 0.5f;
    const float y_rel = iy / (float)n_t - 0.5f;
    const float diag = sqrtf(n_u * n_u + n_v * n_v);
    const float alpha = x_rel * kPiF;
    const float tau = y_rel * diag;
    // :54-58 line in Hessian normal form, origin moved from the image centre to the corner
    const float l0 = -sinf(alpha);
    const float l1 = cosf(alpha);
    float l2 = -tau;
    l2 += -0.5f * n_u * l0 - 0.5f * n_v * l1;
    // :61-65 foot point and direction
    float o0 = -l2 * l0;
    float o1 = -l2 * l1;
    const float d0 = l1;
    const float d1 = -l0;
    // :71-84 clip against the box inset by one pixel
    float ts[4] = {(1.0f - o0) / d0, (n_u - 1.0f - o0) / d0, (1.0f - o1) / d1,
                   (n_v - 1.0f - o1) / d1};
    if (d0 * d0 < 1e-12f) { ts[0] = -1e10f; ts[1] = 1e10f; }
    if (d1 * d1 < 1e-12f) { ts[2] = -1e10f; ts[3] = 1e10f; }
    sort_four(ts);
    float t = ts[1];
    const float t_max = ts[2];
    // :89-92 reject lines that miss the image
    const float pu = fmaf(t, d0, o0), pv = fmaf(t, d1, o1);
    const bool inside = (pu <= n_u && pv <= n_v && pu >= 0 && pv >= 0);
    if (!inside || t_max <= t) return 0.0f;
    // :98-99 texel centres
    o0 += 0.5f;
    o1 += 0.5f;
    const float step = 0.66f;  // :102 (double literal .66 stored in a float)
    float sum = 0.0f;
    if (!derivative) {
        // :107-109
        double cnt = 0;
        for (; t <= t_max; t += step) {
            if (img) sum += bilinear(img, n_ui, n_vi, fmaf(t, d0, o0), fmaf(t, d1, o1), interp);
            cnt += 1;
        }
        if (n_samples) *n_samples += cnt;
        return sum * step;
    }
    // :114-123 two parallel lines, half a pixel either side of the bin's line
    o0 -= 0.5f * d1;
    o1 += 0.5f * d0;
    float sumo = 0.0f;
    double cnt = 0;
    for (; t <= t_max; t += step) {
        const float x = fmaf(t, d0, o0), y = fmaf(t, d1, o1);
        if (img) {
            sum += bilinear(img, n_ui, n_vi, x, y, interp);
            sumo += bilinear(img, n_ui, n_vi, x + d1, y - d0, interp);
        }
        cnt += 2;
    }
    if (n_samples) *n_samples += cnt;
    const float result = (sum - sumo) * step;
    // :125-140 post-processing
    if (post == 1) return result < 0 ? -sqrtf(-result) : sqrtf(result);
    if (post == 2) return result < 0 ? -logf(-result + 1.0f) : logf(result + 1.0f);
    return result;
}

// Sample a Radon intermediate along an epipolar line.  EpipolarConsistencyRadonIntermediate.cu:70-84
// with the dtr texture semantics of RadonIntermediate.cpp:192 (normalised, linear, clamp).
inline float redundancy(const float* K, const float* dtr, int n_alpha, int n_t, float range_t,
                        float x0, float x1, bool is_derivative, int interp)
{
    float line[3] = {fmaf(K[3], x1, K[0] * x0), fmaf(K[4], x1, K[1] * x0),
                     fmaf(K[5], x1, K[2] * x0)};
    const int flipped = oracle_line_to_sample(line, range_t);
    const float v = bilinear(dtr, n_alpha, n_t, line[0] * (float)n_alpha, line[1] * (float)n_t,
                             interp);
    return (is_derivative && flipped) ? -v : v;
}

// CUDA's __sincosf is sin.approx/cos.approx (abs. error about 2^-21); we do not model its bits,
// fast_sincos only selects float-rounded double results versus libm sinf/cosf.
inline void sincos_model(float x, int fast, float* s, float* c)
{
    if (fast) {
        *s = (float)sin((double)x);
        *c = (float)cos((double)x);
    } else {
        *s = sinf(x);
        *c = cosf(x);
    }
}

inline void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
inline double det3(double a, double b, double c, double d, double e, double f, double g, double h,
                   double i)
{
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

}  // namespace

extern "C" {

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Sets the OpenMP team size of the oracle's loops (bench.py's CPU legs: torchrun exports OMP_NUM_THREADS=1 to its ranks,
// which would time the reference arm on one core); returns the resulting omp_get_max_threads().
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// EpipolarConsistencyCommon.hxx:52-79 -- row-major enumeration of the strict upper triangle.
void oracle_get_ij(int k, int n, int* i, int* j)
{
    int row = 0;
    int rest = k;  // index inside the current row
    while (rest >= n - row - 1) {
        rest -= n - row - 1;
        row++;
    }
    *i = row;
    *j = row + 1 + rest;
}

// xprojectionmatrix.hxx:20-52: PinvT = (P P^T)^-1 P  (== transpose of P^T (P P^T)^-1).
void oracle_pinv_transpose(const double* P, float* PinvT)
{
    double G[9];  // Gram matrix P P^T, symmetric
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += P[r + 3 * k] * P[c + 3 * k];
            G[r + 3 * c] = s;
        }
    const double det = det3(G[0], G[3], G[6], G[1], G[4], G[7], G[2], G[5], G[8]);
    double Gi[9];  // inverse by cofactors
    Gi[0] = (G[4] * G[8] - G[7] * G[5]) / det;
    Gi[3] = -(G[3] * G[8] - G[6] * G[5]) / det;
    Gi[6] = (G[3] * G[7] - G[6] * G[4]) / det;
    Gi[1] = -(G[1] * G[8] - G[7] * G[2]) / det;
    Gi[4] = (G[0] * G[8] - G[6] * G[2]) / det;
    Gi[7] = -(G[0] * G[7] - G[6] * G[1]) / det;
    Gi[2] = (G[1] * G[5] - G[4] * G[2]) / det;
    Gi[5] = -(G[0] * G[5] - G[3] * G[2]) / det;
    Gi[8] = (G[0] * G[4] - G[3] * G[1]) / det;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += Gi[r + 3 * k] * P[k + 3 * c];
            PinvT[r + 3 * c] = (float)s;
        }
}

// xprojectionmatrix.hxx:93-105: right null vector of P, scaled so that the last entry is one.
void oracle_source_position(const double* P, float* C)
{
    // null vector by 3x3 minors of the 3x4 matrix: C_k = (-1)^k det(P without column k)
    double m[4];
    for (int k = 0; k < 4; k++) {
        int c[3], q = 0;
        for (int j = 0; j < 4; j++)
            if (j != k) c[q++] = j;
        m[k] = det3(P[0 + 3 * c[0]], P[0 + 3 * c[1]], P[0 + 3 * c[2]], P[1 + 3 * c[0]],
                    P[1 + 3 * c[1]], P[1 + 3 * c[2]], P[2 + 3 * c[0]], P[2 + 3 * c[1]],
                    P[2 + 3 * c[2]]);
        if (k & 1) m[k] = -m[k];
    }
    for (int k = 0; k < 4; k++) C[k] = (float)(m[k] / m[3]);
}

// EpipolarConsistency.cpp:35-47 (estimateObjectRadius), ProjectionMatrix.cpp:104-112
// (getCameraFocalLengthPx), :70-76 (camera centre, de-homogenised).
double oracle_object_radius(const double* P, int n_u, int n_v)
{
    const double m1[3] = {P[0], P[3], P[6]}, m2[3] = {P[1], P[4], P[7]}, m3[3] = {P[2], P[5], P[8]};
    double U[3], V[3], t[3];
    cross3(m3, m2, U);
    cross3(m3, m1, V);
    const double nU = norm3(U), nV = norm3(V);
    for (int k = 0; k < 3; k++) { U[k] /= nU; V[k] /= nV; }
    cross3(V, m3, t);
    const double fu = dot3(m1, t);
    cross3(U, m3, t);
    const double fv = dot3(m2, t);
    const double fov = std::max(fabs(atan(0.5 * n_u / fu)), fabs(atan(0.5 * n_v / fv)));
    // camera centre in double (the reference uses an SVD null space here; same point)
    double m[4];
    for (int k = 0; k < 4; k++) {
        int c[3], q = 0;
        for (int j = 0; j < 4; j++)
            if (j != k) c[q++] = j;
        m[k] = det3(P[0 + 3 * c[0]], P[0 + 3 * c[1]], P[0 + 3 * c[2]], P[1 + 3 * c[0]],
                    P[1 + 3 * c[1]], P[1 + 3 * c[2]], P[2 + 3 * c[0]], P[2 + 3 * c[1]],
                    P[2 + 3 * c[2]]);
        if (k & 1) m[k] = -m[k];
    }
    const double C[3] = {m[0] / m[3], m[1] / m[3], m[2] / m[3]};
    return sin(fov) * norm3(C);
}

// EpipolarConsistencyCommon.hxx:82-149.  same_view stands for the reference's pointer test
// C0==C1 (:108-113), which fires when both matrix indices are equal.
void oracle_compute_k01(float half_nu, float half_nv, const float* C0, const float* C1,
                        const float* P0invT, const float* P1invT, float object_radius_mm,
                        float num_samples, float dkappa, int same_view, float* K0, float* K1)
{
    if (same_view) {
        for (int k = 0; k < 8; k++) K0[k] = K1[k] = 0.0f;
        return;
    }
    // :115-120 Pluecker coordinates of the join of the two source positions
    const float b01 = C0[0] * C1[1] - C0[1] * C1[0];
    const float b02 = C0[0] * C1[2] - C0[2] * C1[0];
    const float b03 = C0[0] * C1[3] - C0[3] * C1[0];
    const float b12 = C0[1] * C1[2] - C0[2] * C1[1];
    const float b13 = C0[1] * C1[3] - C0[3] * C1[1];
    const float b23 = C0[2] * C1[3] - C0[3] * C1[2];
    // :122-123 norms of moment and direction
    const float mom = sqrtf(b12 * b12 + b02 * b02 + b01 * b01);
    const float dir = sqrtf(b03 * b03 + b13 * b13 + b23 * b23);
    // :126-129 the two planes spanning the pencil: through the origin, and farthest from it
    const float E[8] = {b12 / mom,
                        -b02 / mom,
                        b01 / mom,
                        0.0f,
                        (-b01 * b13 - b02 * b23) / (mom * dir),
                        (b01 * b03 - b12 * b23) / (mom * dir),
                        (b02 * b03 + b12 * b13) / (mom * dir),
                        -mom / dir};
    // :131-132 K = PinvT(3x4) * E(4x2), column-major
    const float* Pi[2] = {P0invT, P1invT};
    float* Ko[2] = {K0, K1};
    for (int v = 0; v < 2; v++)
        for (int c = 0; c < 2; c++)
            for (int r = 0; r < 3; r++) {
                float s = 0.0f;
                for (int k = 0; k < 4; k++) s += Pi[v][r + 3 * k] * E[k + 4 * c];
                Ko[v][r + 3 * c] = s;
            }
    // :82-89,134-135 lines relative to the image centre; scale by the kappa=0 line's normal
    for (int v = 0; v < 2; v++) {
        float* K = Ko[v];
        K[2] += half_nu * K[0] + half_nv * K[1];
        K[5] += half_nu * K[3] + half_nv * K[4];
        const float len = sqrtf(K[0] * K[0] + K[1] * K[1]);
        for (int k = 0; k < 6; k++) K[k] /= len;
    }
    // :137-148
    K0[6] = mom / dir;
    K0[7] = -2.0f * atan2f(-0.5f * dir, mom / dir);
    K1[7] = (K0[6] <= object_radius_mm) ? 0.5f * kPiF : asinf(object_radius_mm / K0[6]);
    K1[6] = (dkappa <= 0.0f) ? 2.0f * K1[7] / num_samples : dkappa;
}

// EpipolarConsistencyCommon.hxx:152-171
int oracle_line_to_sample(float* line, float range_t)
{
    const float len = sqrtf(line[0] * line[0] + line[1] * line[1]);
    float a = atan2f(line[1], line[0]) / kPiF;
    if (a < 0) a += 2.0f;
    float d = -(line[2] / len) / range_t + 0.5f;
    int flipped = 0;
    if (a > 1.0f) {
        a -= 1.0f;
        d = 1.0f - d;
        flipped = 1;
    }
    line[0] = a;
    line[1] = d;
    return flipped;
}

void oracle_radon(const float* img, int n_u, int n_v, int n_alpha, int n_t, int filter, int post,
                  int interp, float* out)
{
    const bool derivative = (filter == 0);
#pragma omp parallel for schedule(dynamic, 4)
    for (int iy = 0; iy < n_t; iy++)
        for (int ix = 0; ix < n_alpha; ix++)
            out[(size_t)iy * n_alpha + ix] =
                radon_bin(img, n_u, n_v, ix, iy, n_alpha, n_t, derivative, post, interp, nullptr);
}

double oracle_radon_num_samples(int n_u, int n_v, int n_alpha, int n_t, int filter)
{
    double total = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
    for (int iy = 0; iy < n_t; iy++)
        for (int ix = 0; ix < n_alpha; ix++) {
            double c = 0;
            radon_bin(nullptr, n_u, n_v, ix, iy, n_alpha, n_t, filter == 0, 0, 0, &c);
            total += c;
        }
    return total;
}

double oracle_ecc(const double* Ps, int n_views, const float* dtrs, int n_dtrs, int n_alpha,
                  int n_t, float step_alpha, float step_t, int n_u, int n_v, int is_derivative,
                  double object_radius_mm, double dkappa_d, int interp, int fast_sincos,
                  const int* idx4, int n_pairs, float* out, int* ksamples, int use_corr)
{
    (void)step_alpha;
    (void)n_dtrs;
    // EpipolarConsistencyRadonIntermediate.cpp:134-163 host preparation (double -> float)
    std::vector<float> Cs((size_t)n_views * 4), PinvTs((size_t)n_views * 12);
    for (int v = 0; v < n_views; v++) {
        oracle_pinv_transpose(Ps + 12 * v, &PinvTs[12 * v]);
        oracle_source_position(Ps + 12 * v, &Cs[4 * v]);
    }
    // Metric::getObjectRadius (EpipolarConsistency.cpp:76-84); passed on as float (.cpp:189,297)
    if (!(object_radius_mm > 0))
        object_radius_mm = n_views ? oracle_object_radius(Ps, n_u, n_v) : 0.0;
    const float radius = (float)object_radius_mm;
    const float dkappa = (float)dkappa_d;
    // launcher sizing, EpipolarConsistencyRadonIntermediate.cu:320,347-358
    const float image_diagonal = n_t * step_t * 2.0f;
    const float range_t = n_t * step_t;
    int max_num_samples = (dkappa <= 0.0f) ? (int)image_diagonal : (int)(kPiF * 0.5f / dkappa);
    const int sample_cap = ((max_num_samples + 255) / 256) * 256;  // grid.y * block.y
    const bool all_pairs = (idx4 == nullptr);
    if (all_pairs) n_pairs = n_views * (n_views - 1) / 2;

    double total = 0;
    const size_t dtr_len = (size_t)n_alpha * n_t;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
    for (int p = 0; p < n_pairs; p++) {
        int p0, p1, r0, r1;
        if (all_pairs) {
            oracle_get_ij(p, n_views, &p0, &p1);
            r0 = p0;
            r1 = p1;
        } else {
            p0 = idx4[4 * p + 0];
            p1 = idx4[4 * p + 1];
            r0 = idx4[4 * p + 2];
            r1 = idx4[4 * p + 3];
        }
        float K0[8], K1[8];
        oracle_compute_k01(n_u * 0.5f, n_v * 0.5f, &Cs[4 * p0], &Cs[4 * p1], &PinvTs[12 * p0],
                           &PinvTs[12 * p1], radius, image_diagonal, dkappa, p0 == p1, K0, K1);
        const float dk = K1[6], kmax = K1[7];
        const float* d0 = dtrs + dtr_len * r0;
        const float* d1 = dtrs + dtr_len * r1;
        double acc = 0, sxx = 0, syy = 0, sxy = 0;
        int m = 0;
        // .cu:192-197 / :258-263 kappa grid; .cu:86-113 +/- kappa
        for (; m < sample_cap; m++) {
            // "dkappa*0.5f+dkappa*idx_y" as the reference's kernel EXECUTES it (nvcc 12.9, sm_100; cuobjdump of
            // oracle/_ref/ecc_ri.o: FMUL t = dkappa * idx_y, FFMA kappa = dkappa * 0.5 + t): the product with the sample
            // index is rounded on its own
            const float kappa_t = dk * (float)m;
            const float kappa = fmaf(dk, 0.5f, kappa_t);
            if (kappa >= kmax) break;
            float s, c;
            sincos_model(kappa, fast_sincos, &s, &c);
            const float xp = redundancy(K0, d0, n_alpha, n_t, range_t, c, s, is_derivative, interp);
            const float yp = redundancy(K1, d1, n_alpha, n_t, range_t, c, s, is_derivative, interp);
            c = -c;  // .cu:106: minus kappa via the oppositely oriented line
            const float xm = redundancy(K0, d0, n_alpha, n_t, range_t, c, s, is_derivative, interp);
            const float ym = redundancy(K1, d1, n_alpha, n_t, range_t, c, s, is_derivative, interp);
            if (use_corr) {
                // .cu:115-149 with the factor the launcher passes for "1/n": kappa_max/kappa (.cu:209,274)
                const float w = kmax / kappa;
                sxx += (double)(w * (xp * xp + xm * xm));
                syy += (double)(w * (yp * yp + ym * ym));
                sxy += (double)(w * (xp * yp + xm * ym));
            } else {
                const float vp = xp - yp, vm = xm - ym;
                const float consistency = (vp * vp + vm * vm) * K0[6];  // .cu:112
                acc += (double)(consistency * dk);                      // .cu:204,269 (atomicAdd there)
            }
        }
        float value = (float)acc;
        if (use_corr) {
            // cc() and "1 - corr", EpipolarConsistencyRadonIntermediate.cpp:127-131,207-210,304-308 (un-centred)
            const float xx = (float)sxx, yy = (float)syy, xy = (float)sxy;
            value = 1.0f - xy / (sqrtf(xx) * sqrtf(yy));
        }
        if (ksamples) ksamples[p] = m;
        if (out) {
            if (all_pairs) out[p0 + (size_t)p1 * n_views] = value;  // .cu:269
            else out[p] = value;
        }
        total += (double)value;  // .cpp:216-224,312-321 with all weights 1
    }
    return n_pairs ? total / n_pairs : 0.0;
}

// Projtable.hxx:138-165; CameraOpenGL.hxx:11-31; ProjectionMatrix.cpp:12-18,133-145.
void oracle_circular_trajectory(int n_proj, double sid, double sdd, int n_u, int n_v,
                                double max_angle_deg, double pixel_spacing, double* Ps)
{
    const double fovy = atan(n_v * pixel_spacing / sdd);
    const double f = n_v / (2.0 * tan(0.5 * fovy));  // cameraPerspective: height / (2 tan(fovy/2))
    const double K[9] = {f, 0, 0, 0, f, 0, 0.5 * n_u, 0.5 * n_v, 1};  // column-major 3x3
    const double ct = cos(0.5 * kPiD), st = sin(0.5 * kPiD);          // rotation about x by 90 deg
    for (int i = 0; i < n_proj; i++) {
        const double ang = i * (max_angle_deg / n_proj) / 180.0 * kPiD;
        const double eye[3] = {sid * cos(ang), 0.0, sid * sin(ang)};
        // cameraLookAt: rows of R are left, up, -forward
        double fwd[3] = {-eye[0], -eye[1], -eye[2]};
        const double nf = norm3(fwd);
        for (int k = 0; k < 3; k++) fwd[k] /= nf;
        const double up0[3] = {0, 1, 0};
        double left[3], up[3];
        cross3(up0, fwd, left);
        const double nl = norm3(left);
        for (int k = 0; k < 3; k++) left[k] /= nl;
        cross3(fwd, left, up);
        double R[9], t[3];  // R row-major here
        for (int k = 0; k < 3; k++) { R[0 + k] = left[k]; R[3 + k] = up[k]; R[6 + k] = -fwd[k]; }
        for (int r = 0; r < 3; r++) t[r] = -(R[3 * r] * eye[0] + R[3 * r + 1] * eye[1] + R[3 * r + 2] * eye[2]);
        // P = K [R | t]  (3x4, kept as rows for the moment)
        double Pm[3][4];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) {
                double s = 0;
                for (int k = 0; k < 3; k++) s += K[r + 3 * k] * R[3 * k + c];
                Pm[r][c] = s;
            }
            double s = 0;
            for (int k = 0; k < 3; k++) s += K[r + 3 * k] * t[k];
            Pm[r][3] = s;
        }
        auto normalize = [&](double (*M)[4]) {
            double n3 = sqrt(M[2][0] * M[2][0] + M[2][1] * M[2][1] + M[2][2] * M[2][2]);
            const double d = det3(M[0][0], M[0][1], M[0][2], M[1][0], M[1][1], M[1][2], M[2][0],
                                  M[2][1], M[2][2]);
            if (d < 0) n3 = -n3;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++) M[r][c] *= (1.0 / n3);
        };
        normalize(Pm);  // makeProjectionMatrix normalises once
        // right-multiply by T_rot_x: rows/cols 1,2 form [[ct,-st],[st,ct]]
        double Q[3][4];
        for (int r = 0; r < 3; r++) {
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

void oracle_project_ellipsoids(const double* P, int n_u, int n_v, const double* ell, int n_ell,
                               int cos_weight, int zero_border, float* img)
{
    // Back-projection of pixel (u,v): ray direction M^-1 (u,v,1)^T from the source position C.
    const double M[9] = {P[0], P[1], P[2], P[3], P[4], P[5], P[6], P[7], P[8]};  // col-major 3x3
    const double det = det3(M[0], M[3], M[6], M[1], M[4], M[7], M[2], M[5], M[8]);
    double Mi[9];
    Mi[0] = (M[4] * M[8] - M[7] * M[5]) / det;
    Mi[3] = -(M[3] * M[8] - M[6] * M[5]) / det;
    Mi[6] = (M[3] * M[7] - M[6] * M[4]) / det;
    Mi[1] = -(M[1] * M[8] - M[7] * M[2]) / det;
    Mi[4] = (M[0] * M[8] - M[6] * M[2]) / det;
    Mi[7] = -(M[0] * M[7] - M[6] * M[1]) / det;
    Mi[2] = (M[1] * M[5] - M[4] * M[2]) / det;
    Mi[5] = -(M[0] * M[5] - M[3] * M[2]) / det;
    Mi[8] = (M[0] * M[4] - M[3] * M[1]) / det;
    double C[3];
    for (int r = 0; r < 3; r++) C[r] = -(Mi[r] * P[9] + Mi[r + 3] * P[10] + Mi[r + 6] * P[11]);
    // Intrinsics for the cosine weight (PreProccess.cpp:146-166 uses K(0,0), K(0,2), K(1,2) of the
    // RQ decomposition).  For K[R|t]: m3 = r3 (unit after normalisation), pp = (m1.m3, m2.m3),
    // f = |m1 - ppu*m3|.
    const double m1[3] = {P[0], P[3], P[6]}, m2[3] = {P[1], P[4], P[7]}, m3[3] = {P[2], P[5], P[8]};
    const double n3 = dot3(m3, m3);
    const double ppu = dot3(m1, m3) / n3, ppv = dot3(m2, m3) / n3;
    double tmp[3] = {m1[0] - ppu * m3[0], m1[1] - ppu * m3[1], m1[2] - ppu * m3[2]};
    const float sdd_px = (float)(norm3(tmp) / sqrt(n3));
    const float ppuf = (float)ppu, ppvf = (float)ppv;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < n_v; v++)
        for (int u = 0; u < n_u; u++) {
            double d[3];
            for (int r = 0; r < 3; r++) d[r] = Mi[r] * u + Mi[r + 3] * v + Mi[r + 6];
            const double dn = norm3(d);
            for (int r = 0; r < 3; r++) d[r] /= dn;
            double val = 0;
            for (int e = 0; e < n_ell; e++) {
                const double* q = ell + 7 * e;
                // scale space so the ellipsoid becomes the unit sphere
                const double o[3] = {(C[0] - q[0]) / q[3], (C[1] - q[1]) / q[4], (C[2] - q[2]) / q[5]};
                const double w[3] = {d[0] / q[3], d[1] / q[4], d[2] / q[5]};
                const double a = dot3(w, w), b = dot3(o, w), c = dot3(o, o) - 1.0;
                const double disc = b * b - a * c;
                if (disc > 0) val += q[6] * 2.0 * sqrt(disc) / a;  // chord length (|d| = 1)
            }
            float pix = (float)val;
            if (cos_weight) {
                const float pou = (float)u - ppuf, pov = (float)v - ppvf;
                pix *= sdd_px / sqrtf(pou * pou + pov * pov + sdd_px * sdd_px);
            }
            if (zero_border && (u == 0 || v == 0 || u == n_u - 1 || v == n_v - 1)) pix = 0.0f;
            img[(size_t)v * n_u + u] = pix;
        }
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Direct metric (no Radon intermediates).  Restates LibEpipolarConsistency/EpipolarConsistencyDirect.cpp
// (host geometry, Eigen there, plain arrays here), EpipolarConsistencyDirect.cu:31-119 (the line kernel) and
// RectifiedFBCC.h (fan-beam weighting).  PARITY UNPINNED for the host geometry: the reference takes the
// pseudo-inverse and the camera centre from Eigen's JacobiSVD and Eigen is not available here; both are
// restated by closed forms that give the same quantities to fp64 rounding.  The line kernel IS pinned: the
// reference's own EpipolarConsistencyDirect.cu is compiled into oracle/_ref/libecc_ref_cuda.so and the GPU
// tests feed it the same lines.
// ---------------------------------------------------------------------------------------------
namespace {

struct DView {
    double pinvT[12];  // (P^+)^T, 3x4 col-major
    double C[4];
    const double* P;
};

// ProjectionMatrix.cpp:21-24 (pseudoInverse) and :70-76 (getCameraCenter).
void dview(const double* P, DView& V)
{
    V.P = P;
    double G[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += P[r + 3 * k] * P[c + 3 * k];
            G[r + 3 * c] = s;
        }
    const double det = det3(G[0], G[3], G[6], G[1], G[4], G[7], G[2], G[5], G[8]);
    double Gi[9];
    Gi[0] = (G[4] * G[8] - G[7] * G[5]) / det;
    Gi[3] = -(G[3] * G[8] - G[6] * G[5]) / det;
    Gi[6] = (G[3] * G[7] - G[6] * G[4]) / det;
    Gi[1] = -(G[1] * G[8] - G[7] * G[2]) / det;
    Gi[4] = (G[0] * G[8] - G[6] * G[2]) / det;
    Gi[7] = -(G[0] * G[7] - G[6] * G[1]) / det;
    Gi[2] = (G[1] * G[5] - G[4] * G[2]) / det;
    Gi[5] = -(G[0] * G[5] - G[3] * G[2]) / det;
    Gi[8] = (G[0] * G[4] - G[3] * G[1]) / det;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += Gi[r + 3 * k] * P[k + 3 * c];
            V.pinvT[r + 3 * c] = s;
        }
    double m[4], nrm = 0;
    for (int k = 0; k < 4; k++) {
        int c[3], q = 0;
        for (int j = 0; j < 4; j++)
            if (j != k) c[q++] = j;
        m[k] = det3(P[0 + 3 * c[0]], P[0 + 3 * c[1]], P[0 + 3 * c[2]], P[1 + 3 * c[0]], P[1 + 3 * c[1]], P[1 + 3 * c[2]],
                    P[2 + 3 * c[0]], P[2 + 3 * c[1]], P[2 + 3 * c[2]]);
        if (k & 1) m[k] = -m[k];
        nrm += m[k] * m[k];
    }
    nrm = sqrt(nrm);
    for (int k = 0; k < 4; k++) V.C[k] = m[k] / nrm;
    if (V.C[3] < -1e-12 || V.C[3] > 1e-12) {
        const double w = V.C[3];
        for (int k = 0; k < 4; k++) V.C[k] /= w;
    }
}

// ProjectiveGeometry.hxx:188-237.
void pl_join_points(const double* A, const double* B, double* L)
{
    L[0] = A[0] * B[1] - A[1] * B[0]; L[1] = A[0] * B[2] - A[2] * B[0]; L[2] = A[0] * B[3] - A[3] * B[0];
    L[3] = A[1] * B[2] - A[2] * B[1]; L[4] = A[1] * B[3] - A[3] * B[1]; L[5] = A[2] * B[3] - A[3] * B[2];
}
void pl_meet_planes(const double* A, const double* B, double* L)
{
    L[0] = A[2] * B[3] - A[3] * B[2]; L[1] = A[3] * B[1] - A[1] * B[3]; L[2] = A[1] * B[2] - A[2] * B[1];
    L[3] = A[0] * B[3] - A[3] * B[0]; L[4] = A[2] * B[0] - A[0] * B[2]; L[5] = A[0] * B[1] - A[1] * B[0];
}
void pl_join_line_point(const double* L, const double* X, double* E)
{
    E[0] = X[1] * L[5] - X[2] * L[4] + X[3] * L[3];
    E[1] = -X[0] * L[5] + X[2] * L[2] - X[3] * L[1];
    E[2] = X[0] * L[4] - X[1] * L[2] + X[3] * L[0];
    E[3] = -X[0] * L[3] + X[1] * L[1] - X[2] * L[0];
}
void pl_meet_line_plane(const double* L, const double* P, double* X)
{
    X[0] = -P[1] * L[0] - P[2] * L[1] - P[3] * L[2];
    X[1] = P[0] * L[0] - P[2] * L[3] - P[3] * L[4];
    X[2] = P[0] * L[1] + P[1] * L[3] - P[3] * L[5];
    X[3] = P[0] * L[2] + P[1] * L[4] + P[2] * L[5];
}

struct DPair {
    double E0[4], E90[4], lo, hi, dkappa, dir[3], E[4], H0[9], H1[9];
    int n_lines;
};

// EpipolarConsistencyDirect.cpp:26-40 (pencil), :84-105 + EpipolarConsistency.cpp:49-60 (range, step, count), :128-145
// (virtual detector and rectifying homographies).
void dpair(const DView& V0, const DView& V1, double radius, double dkappa, int n_u, int n_v, DPair& R)
{
    double B[6];
    pl_join_points(V0.C, V1.C, B);
    const double origin[4] = {0, 0, 0, 1};
    pl_join_line_point(B, origin, R.E0);
    pl_join_line_point(B, R.E0, R.E90);
    const double n0 = norm3(R.E0), n90 = norm3(R.E90);
    for (int k = 0; k < 4; k++) { R.E0[k] /= n0; R.E90[k] /= n90; }
    const double d[3] = {-B[2], -B[4], -B[5]}, m[3] = {B[3], -B[1], B[0]};  // ProjectiveGeometry.hxx:242-251
    const double dist = norm3(m) / norm3(d);
    if (dist <= radius) { R.lo = -0.5 * kPiD; R.hi = 0.5 * kPiD; }
    else { const double kmax = fabs(asin(radius / dist)); R.lo = -kmax; R.hi = kmax; }
    if (dkappa <= 0) dkappa = 0.5 * (R.hi - R.lo) / sqrt((double)(n_u * n_u + n_v * n_v));
    R.dkappa = dkappa;
    R.n_lines = (int)((R.hi - R.lo) / dkappa);
    for (int k = 0; k < 3; k++) R.dir[k] = d[k];
    double U[3], W[3];
    const double nd = norm3(d), nm = norm3(m);
    for (int k = 0; k < 3; k++) { U[k] = d[k] / nd; W[k] = m[k] / nm; }
    cross3(U, W, R.E);
    R.E[3] = 0;
    for (int view = 0; view < 2; view++) {
        const DView& V = view ? V1 : V0;
        const double* C = V.C;
        const double* E = R.E;
        double T[4][4] = {  // ProjectiveGeometry.hxx:333-343, T[row][col]
            {C[1] * E[1] + C[2] * E[2] + C[3] * E[3], -C[0] * E[1], -C[0] * E[2], -C[0] * E[3]},
            {-C[1] * E[0], C[0] * E[0] + C[2] * E[2] + C[3] * E[3], -C[1] * E[2], -C[1] * E[3]},
            {-C[2] * E[0], -C[2] * E[1], C[0] * E[0] + C[3] * E[3] + C[1] * E[1], -C[2] * E[3]},
            {-C[3] * E[0], -C[3] * E[1], -C[3] * E[2], C[0] * E[0] + C[1] * E[1] + C[2] * E[2]}};
        const double PE[3][4] = {{U[0], U[1], U[2], 0}, {W[0], W[1], W[2], 0}, {0, 0, 0, 1.0}};
        double PT[3][4];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += PE[r][k] * T[k][c];
                PT[r][c] = s;
            }
        double* H = view ? R.H1 : R.H0;  // H(r,c) at H[r + 3c]; P^+(k,c) = pinvT(c,k)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                double s = 0;
                for (int k = 0; k < 4; k++) s += PT[r][k] * V.pinvT[c + 3 * k];
                H[r + 3 * c] = s;
            }
    }
}

// EpipolarConsistencyDirect.cpp:47-59.
void dlines(const DView& V0, const DView& V1, const DPair& R, double kappa, float* l0, float* l1)
{
    double Ek[4];
    for (int k = 0; k < 4; k++) Ek[k] = cos(kappa) * R.E0[k] + sin(kappa) * R.E90[k];
    for (int view = 0; view < 2; view++) {
        const double* A = view ? V1.pinvT : V0.pinvT;
        double l[3];
        for (int r = 0; r < 3; r++) {
            l[r] = 0;
            for (int c = 0; c < 4; c++) l[r] += A[r + 3 * c] * Ek[c];
        }
        const double n = sqrt(l[0] * l[0] + l[1] * l[1]);
        for (int r = 0; r < 3; r++) (view ? l1 : l0)[r] = (float)(l[r] / n);
    }
}

// RectifiedFBCC.h:66-69 in fp32 as the host evaluates it.
float phi_transform(const float* f, float t) { return (f[0] * t + f[1]) / (f[2] * t + f[3]); }

// EpipolarConsistencyDirect.cpp:147-186 + RectifiedFBCC.h:41-50: 8 floats per view (a, b, c, d, t_prime_ak, d_l_kappa_C_sq, 0, 0).
void dfbcc(const DView& V0, const DView& V1, const DPair& R, const float* l0f, const float* l1f, float* f0, float* f1)
{
    const double l0[3] = {l0f[0], l0f[1], l0f[2]}, l1[3] = {l1f[0], l1f[1], l1f[2]};
    double Ek[4];
    for (int c = 0; c < 4; c++) Ek[c] = V0.P[3 * c] * l0[0] + V0.P[1 + 3 * c] * l0[1] + V0.P[2 + 3 * c] * l0[2];
    for (int view = 0; view < 2; view++) {
        const DView& V = view ? V1 : V0;
        const double* l = view ? l1 : l0;
        const double* H = view ? R.H1 : R.H0;
        float* f = view ? f1 : f0;
        const double EB[4] = {R.dir[0], R.dir[1], R.dir[2], -dot3(R.dir, V.C)};
        double L[6], Ak[4];
        pl_meet_planes(EB, Ek, L);
        pl_meet_line_plane(L, R.E, Ak);
        if (Ak[3] > 1e-12 || Ak[3] < -1e-12) { const double w = Ak[3]; for (int k = 0; k < 4; k++) Ak[k] /= w; }
        else { Ak[3] = 0; const double n = norm3(Ak); for (int k = 0; k < 4; k++) Ak[k] /= n; }
        double dd = 0;
        for (int k = 0; k < 4; k++) dd += (Ak[k] - V.C[k]) * (Ak[k] - V.C[k]);
        const float d_px = (float)sqrt(dd);
        double ak[3];
        for (int r = 0; r < 3; r++) ak[r] = V.P[r] * Ak[0] + V.P[r + 3] * Ak[1] + V.P[r + 6] * Ak[2] + V.P[r + 9] * Ak[3];
        if (ak[2] > 1e-11 || ak[2] < -1e-11) { const double w = ak[2]; for (int k = 0; k < 3; k++) ak[k] /= w; }
        else { ak[2] = 0; const double n = sqrt(ak[0] * ak[0] + ak[1] * ak[1]); for (int k = 0; k < 3; k++) ak[k] /= n; }
        f[0] = (float)(H[0] * l[1] - H[1] * l[0]);
        f[1] = (float)(H[6] - H[0] * l[0] * l[2]);
        f[2] = (float)(H[2] * l[1] - H[5] * l[0]);
        f[3] = (float)(H[8] - H[2] * l[0] * l[2] - H[5] * l[1] * l[2]);
        if (f[0] * f[3] - f[1] * f[2] < 0) { f[0] *= -1; f[1] *= -1; }
        f[4] = phi_transform(f, (float)(l[1] * ak[0] / ak[2] - l[0] * ak[1] / ak[2]));
        f[5] = d_px * d_px;
        f[6] = f[7] = 0;
    }
}

// One line integral.  EpipolarConsistencyDirect.cu:44-118; fmaf() where the reference's sm_100 build fuses.  shape 1: the
// loop as that build executes it (blocks of four samples under one test, then two, then one), shape 0: the source loop.
float dline_integral(const float* img, int n_ui, int n_vi, int n_v_clip, const float* l, const float* fbcc, int interp, int shape)
{
    float o0 = -l[2] * l[0], o1 = -l[2] * l[1];
    const float d0 = l[1], d1 = -l[0];
    float ts[4] = {(1.0f - o0) / d0, ((float)(n_ui - 1) - o0) / d0, (1.0f - o1) / d1, ((float)(n_v_clip - 1) - o1) / d1};
    if ((double)(d0 * d0) < 1e-12) { ts[1] = 1e10f; ts[0] = -1e10f; }
    if ((double)(d1 * d1) < 1e-12) { ts[3] = 1e10f; ts[2] = -1e10f; }
    sort_four(ts);
    const float t_min = ts[1], t_max = ts[2];
    const float pu = fmaf(t_min, d0, o0), pv = fmaf(t_min, d1, o1);
    if (!(pu <= (float)n_ui && pv <= (float)n_v_clip && pu >= 0 && pv >= 0)) return 0.0f;
    const float step = 0.4f;
    o0 += 0.5f;
    o1 += 0.5f;
    if (fbcc) {
        const float a = fbcc[0], b = fbcc[1], c = fbcc[2], d = fbcc[3], t_ak = fbcc[4], dsq = fbcc[5];
        const float det = fmaf(a, d, -(b * c)), cc = c * c, dd = d * d, cd2 = d * (c + c);
        float sum = 0;
        for (float t = t_min; t <= t_max; t += step) {
            const float u_prime = fmaf(a, t, b) / fmaf(c, t, d) - t_ak;
            const float den = dd + fmaf(cd2, t, (cc * t) * t);
            const float w = (det / den) / sqrtf(fmaf(u_prime, u_prime, dsq));
            sum = fmaf(bilinear(img, n_ui, n_vi, fmaf(d0, t, o0), fmaf(d1, t, o1), interp) * step, w, sum);
        }
        return sum;
    }
    float sump = 0, summ = 0;
    auto sample = [&](float t) {
        const float u = fmaf(d0, t, o0), v = fmaf(d1, t, o1);
        sump = fmaf(bilinear(img, n_ui, n_vi, fmaf(l[0], 0.5f, u), fmaf(l[1], 0.5f, v), interp), step, sump);
        summ = fmaf(bilinear(img, n_ui, n_vi, fmaf(l[0], -0.5f, u), fmaf(l[1], -0.5f, v), interp), step, summ);
    };
    float t = t_min;
    if (!shape) {
        for (; t <= t_max; t += step) sample(t);
        return sump - summ;
    }
    if (t > t_max) return 0.0f;
    bool none_yet = true;
    if (!(t + 1.2f > t_max)) {
        const float r3 = t_max - 1.2f;
        do {
            const float t1 = t + step, t2 = t1 + step, t3 = t2 + step;
            sample(t); sample(t1); sample(t2); sample(t3);
            t = t3 + step;
        } while (!(t > r3));
        none_yet = false;
    }
    const float t1 = t + step;
    if (!(t1 > t_max)) { sample(t); sample(t1); t = t1 + step; none_yet = false; }
    if (t <= t_max || none_yet) sample(t);
    return sump - summ;
}

}  // namespace

extern "C" {

// The weight one sample of a fan-beam line integral carries (EpipolarConsistencyDirect.cu:86-93, RectifiedFBCC.h:66-79), with
// the roundings of dline_integral.  rec: a, b, c, d, t_prime_ak, d_l_kappa_C_sq.
float oracle_direct_fbcc_weight(const float* rec, float t)
{
    const float a = rec[0], b = rec[1], c = rec[2], d = rec[3];
    const float det = fmaf(a, d, -(b * c)), cc = c * c, dd = d * d, cd2 = d * (c + c);
    const float u_prime = fmaf(a, t, b) / fmaf(c, t, d) - rec[4];
    const float den = dd + fmaf(cd2, t, (cc * t) * t);
    return (det / den) / sqrtf(fmaf(u_prime, u_prime, rec[5]));
}

// cuda_computeLineIntegrals, EpipolarConsistencyDirect.cu:122-142 (n_v_clip: the height the clipping uses; the reference
// passes n_u).  lines: stride floats per line; fbcc nullable: fbcc_stride floats per line.
void oracle_direct_line_integrals(const float* img, int n_u, int n_v, int n_v_clip, const float* lines, int n_lines, int stride,
                                  const float* fbcc, int fbcc_stride, int interp, int shape, float* out)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int k = 0; k < n_lines; k++)
        out[k] = dline_integral(img, n_u, n_v, n_v_clip, lines + (size_t)k * stride, fbcc ? fbcc + (size_t)k * fbcc_stride : nullptr, interp, shape);
}

// Geometry of one pair (computeForImagePair up to its kernel launches, EpipolarConsistencyDirect.cpp:64-186).  All output
// arrays nullable, `capacity` planes each: kappas, lines (3 floats), fbcc records (8 floats).  Returns the number of planes.
int oracle_direct_pair_geometry(const double* P0, const double* P1, double radius, double dkappa, int n_u, int n_v, int capacity,
                                float* kappas, float* lines0, float* lines1, float* fbcc0, float* fbcc1, double* dkappa_out)
{
    DView V0, V1;
    dview(P0, V0);
    dview(P1, V1);
    if (radius <= 0) radius = std::max(oracle_object_radius(P0, n_u, n_v), oracle_object_radius(P1, n_u, n_v));  // :80-82
    DPair R;
    dpair(V0, V1, radius, dkappa, n_u, n_v, R);
    if (dkappa_out) *dkappa_out = R.dkappa;
    for (int q = 0; q < R.n_lines && q < capacity; q++) {
        const float kf = (float)(R.lo + R.dkappa * q);  // :104
        float l0[3], l1[3], f0[8], f1[8];
        dlines(V0, V1, R, (double)kf, l0, l1);
        if (kappas) kappas[q] = kf;
        if (lines0) memcpy(lines0 + 3 * (size_t)q, l0, sizeof(l0));
        if (lines1) memcpy(lines1 + 3 * (size_t)q, l1, sizeof(l1));
        if (fbcc0 || fbcc1) {
            dfbcc(V0, V1, R, l0, l1, f0, f1);
            if (fbcc0) memcpy(fbcc0 + 8 * (size_t)q, f0, sizeof(f0));
            if (fbcc1) memcpy(fbcc1 + 8 * (size_t)q, f1, sizeof(f1));
        }
    }
    return R.n_lines;
}

// computeForImagePair, EpipolarConsistencyDirect.cpp:64-212: the pair's metric and (nullable, `capacity` entries) its
// redundant signals.  n_given > 0: the first n_given kappas are the caller's.  reference_clip: clip against n_u x n_u.
double oracle_direct_pair(const double* P0, const double* P1, const float* img0, const float* img1, int n_u, int n_v, double radius,
                          double dkappa, int fbcc, int interp, int shape, int reference_clip, int n_given, int capacity, float* kappas,
                          float* s0, float* s1, int* n_lines_out)
{
    DView V0, V1;
    dview(P0, V0);
    dview(P1, V1);
    if (radius <= 0) radius = std::max(oracle_object_radius(P0, n_u, n_v), oracle_object_radius(P1, n_u, n_v));
    DPair R;
    dpair(V0, V1, radius, dkappa, n_u, n_v, R);
    const int n_lines = n_given > 0 ? n_given : R.n_lines;
    if (n_lines_out) *n_lines_out = n_lines;
    const int clip = reference_clip ? n_u : n_v;
    std::vector<double> terms(n_lines > 0 ? n_lines : 0);
#pragma omp parallel for schedule(dynamic, 8)
    for (int q = 0; q < n_lines; q++) {
        const float kf = n_given > 0 ? kappas[q] : (float)(R.lo + R.dkappa * q);
        float l0[3], l1[3], f0[8], f1[8];
        dlines(V0, V1, R, (double)kf, l0, l1);
        if (fbcc) dfbcc(V0, V1, R, l0, l1, f0, f1);
        const float v0 = dline_integral(img0, n_u, n_v, clip, l0, fbcc ? f0 : nullptr, interp, shape);
        const float v1 = dline_integral(img1, n_u, n_v, clip, l1, fbcc ? f1 : nullptr, interp, shape);
        terms[q] = (double)((v0 - v1) * (v0 - v1)) * R.dkappa;  // :207-210
        if (q < capacity) {
            if (kappas && n_given == 0) kappas[q] = kf;
            if (s0) s0[q] = v0;
            if (s1) s1[q] = v1;
        }
    }
    double metric = 0;
    for (int q = 0; q < n_lines; q++) metric += terms[q];
    return metric;
}

// MetricDirect::evaluate, EpipolarConsistencyDirect.cpp:236-247: all pairs i < j, cost image entry i + j n (nullable), returns
// the SUM.  radius <= 0: Metric::getObjectRadius, i.e. estimated from the first matrix (EpipolarConsistency.cpp:76-84).
double oracle_direct_evaluate(const double* Ps, int n, const float* images, int n_u, int n_v, double radius, double dkappa, int fbcc,
                              int interp, int shape, int reference_clip, float* cost_image)
{
    if (radius <= 0 && n > 0) radius = oracle_object_radius(Ps, n_u, n_v);
    const size_t px = (size_t)n_u * n_v;
    double cost = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            const double ecc = oracle_direct_pair(Ps + 12 * (size_t)i, Ps + 12 * (size_t)j, images + px * i, images + px * j, n_u, n_v, radius,
                                                  dkappa, fbcc, interp, shape, reference_clip, 0, 0, nullptr, nullptr, nullptr, nullptr);
            cost += ecc;
            if (cost_image) cost_image[i + (size_t)j * n] = (float)ecc;
        }
    return cost;
}

}  // extern "C"
