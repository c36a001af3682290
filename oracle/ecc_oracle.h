/* ecc_oracle.h -- CPU oracle for the Epipolar-Consistency hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference algorithm
 * (aaichert/EpipolarConsistency, paths below are relative to its code/ directory).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (epipolarconsistency_b200/) never links, imports or calls anything in oracle/.
 *
 * Pinning: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 4 / 8c).  The oracle is therefore pinned against outputs of the reference itself:
 *   (a) oracle/_ref/libecc_ref_host.so -- the reference's own host/device-shared headers
 *       (EpipolarConsistencyCommon.hxx, culaut/xprojectionmatrix.hxx) compiled unchanged from
 *       /root/reference; tests/golden/ref_host_vectors.npz holds their outputs on seeded inputs;
 *   (b) oracle/_ref/libecc_ref_cuda.so -- the reference's own CUDA translation units compiled
 *       unchanged for sm_100 and run on a B200; tests/golden/ref_cuda_*.npz holds those outputs.
 *
 * All arithmetic is fp32 unless stated, as in the reference kernels.  Matrices are column-major
 * (Eigen default): a 3x4 projection matrix is P[r + 3*c].
 */
#ifndef ECC_ORACLE_H
#define ECC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Bilinear interpolation flavour (hard part 2). */
enum {
    ORACLE_INTERP_EXACT = 0, /* exact fp32 weights: the "CPU float path"                         */
    ORACLE_INTERP_TEX8  = 1  /* weights quantised to 8 fractional bits like the CUDA texture unit */
};

/* Pair enumeration.  Follows LibEpipolarConsistency/EpipolarConsistencyCommon.hxx:52-79. */
void oracle_get_ij(int k, int n, int* i, int* j);

/* (P^+)^T as 3x4 column-major floats from a 3x4 column-major double matrix.
 * Follows LibUtilsCuda/culaut/xprojectionmatrix.hxx:20-52 (result, not algorithm: we invert
 * P P^T by cofactors instead of Householder QR). */
void oracle_pinv_transpose(const double* P, float* PinvT);

/* Centre of projection (null vector of P, de-homogenised so that C[3]==1).
 * Follows LibUtilsCuda/culaut/xprojectionmatrix.hxx:93-105 (result, not algorithm). */
void oracle_source_position(const double* P, float* C);

/* Automatic object radius.  Follows LibEpipolarConsistency/EpipolarConsistency.cpp:35-47,76-84
 * with getCameraFocalLengthPx from LibProjectiveGeometry/ProjectionMatrix.cpp:104-112. */
double oracle_object_radius(const double* P, int n_u, int n_v);

/* K0/K1 record (8+8 floats).  Follows EpipolarConsistencyCommon.hxx:82-149. */
void oracle_compute_k01(float half_nu, float half_nv, const float* C0, const float* C1,
                        const float* P0invT, const float* P1invT, float object_radius_mm,
                        float num_samples, float dkappa, int same_view, float* K0, float* K1);

/* Line -> (angle, distance) texture coordinates.  Follows EpipolarConsistencyCommon.hxx:152-171.
 * line[3] in/out; returns 1 if the angle was moved by pi. */
int oracle_line_to_sample(float* line, float range_t);

/* Radon intermediate of one image.  Follows LibEpipolarConsistency/RadonIntermediate.cu:31-143
 * (kernel body) with the image sampled like the reference's texture
 * (LibUtilsCuda/CudaBindlessTexture.cpp:36-40: unnormalised, linear, clamp).
 * img: n_v rows of n_u floats.  out: n_t rows of n_alpha floats (alpha fastest).
 * filter: 0 derivative, 2 none (1 = ramp: restated with numpy's FFT in tests/oracle_lib.py::ramp_filter on top of filter 2).  post: 0 identity, 1 sqrt, 2 log. */
void oracle_radon(const float* img, int n_u, int n_v, int n_alpha, int n_t, int filter, int post,
                  int interp, float* out);

/* Number of bilinear samples oracle_radon takes for this geometry (work counter for benches). */
double oracle_radon_num_samples(int n_u, int n_v, int n_alpha, int n_t, int filter);

/* Metric over an explicit pair list (idx4 = P0,P1,dtr0,dtr1 per pair) or, with idx4==NULL, over
 * all n_views*(n_views-1)/2 pairs in oracle_get_ij order.
 * Follows LibEpipolarConsistency/EpipolarConsistencyRadonIntermediate.cu:13-113,151-276 (kernels),
 * :278-409 (launcher sizing) and EpipolarConsistencyRadonIntermediate.cpp:134-163,166-225,267-322
 * (host preparation and reduction).
 * Ps: n_views 3x4 col-major doubles.  dtrs: n_dtrs images of n_t rows x n_alpha floats.
 * out: all-pairs -> n_views*n_views cost image, entry i+j*n_views for i<j, other entries untouched
 *      (may be NULL); pair list -> n_pairs floats (may be NULL).
 * ksamples (may be NULL): number of kappa samples taken per pair.
 * use_corr: the correlation variant (EpipolarConsistencyRadonIntermediate.cu:115-149, .cpp:127-131): pair value =
 *   1 - sum(w x y) / (sqrt(sum(w x x)) sqrt(sum(w y y))) with w = kappa_max/kappa, as the reference computes it.
 * Returns the mean over evaluated pairs. */
double oracle_ecc(const double* Ps, int n_views, const float* dtrs, int n_dtrs, int n_alpha,
                  int n_t, float step_alpha, float step_t, int n_u, int n_v, int is_derivative,
                  double object_radius_mm, double dkappa, int interp, int fast_sincos,
                  const int* idx4, int n_pairs, float* out, int* ksamples, int use_corr);

/* --- synthetic data ------------------------------------------------- */

/* Circular trajectory.  Follows HeaderOnly/Utils/Projtable.hxx:138-165 with
 * LibProjectiveGeometry/CameraOpenGL.hxx:11-31 and ProjectionMatrix.cpp:12-18. */
void oracle_circular_trajectory(int n_proj, double sid, double sdd, int n_u, int n_v,
                                double max_angle_deg, double pixel_spacing, double* Ps);

/* Cone-beam projection of a sum of ellipsoids (centre c, semi-axes r, density rho; 7 doubles
 * each: cx,cy,cz,rx,ry,rz,rho), followed by cosine weighting as in
 * LibEpipolarConsistency/Gui/PreProccess.cpp:146-166 and zeroing of the one-pixel border. */
void oracle_project_ellipsoids(const double* P, int n_u, int n_v, const double* ellipsoids,
                               int n_ell, int cos_weight, int zero_border, float* img);

/* --- direct metric (no Radon intermediates) ------------------------------------------------
 * Restates LibEpipolarConsistency/EpipolarConsistencyDirect.cpp:22-270, EpipolarConsistencyDirect.cu:31-142 and
 * RectifiedFBCC.h.  PARITY UNPINNED for the fp64 host geometry (the reference uses Eigen's JacobiSVD, which is not
 * available here); the line kernel is pinned by the reference's own .cu in oracle/_ref/libecc_ref_cuda.so.
 * shape: 0 = the source loop, 1 = the loop as the reference's sm_100 build executes it (blocks of 4 / 2 / 1 samples). */
void oracle_direct_line_integrals(const float* img, int n_u, int n_v, int n_v_clip, const float* lines, int n_lines, int stride,
                                  const float* fbcc, int fbcc_stride, int interp, int shape, float* out);
float oracle_direct_fbcc_weight(const float* rec, float t);
int oracle_direct_pair_geometry(const double* P0, const double* P1, double radius, double dkappa, int n_u, int n_v, int capacity,
                                float* kappas, float* lines0, float* lines1, float* fbcc0, float* fbcc1, double* dkappa_out);
double oracle_direct_pair(const double* P0, const double* P1, const float* img0, const float* img1, int n_u, int n_v, double radius,
                          double dkappa, int fbcc, int interp, int shape, int reference_clip, int n_given, int capacity, float* kappas,
                          float* s0, float* s1, int* n_lines_out);
double oracle_direct_evaluate(const double* Ps, int n, const float* images, int n_u, int n_v, double radius, double dkappa, int fbcc,
                              int interp, int shape, int reference_clip, float* cost_image);

int oracle_max_threads(void);
int oracle_set_threads(int n);  /* OpenMP threads for the oracle's loops; returns the resulting maximum */

#ifdef __cplusplus
}
#endif
#endif
