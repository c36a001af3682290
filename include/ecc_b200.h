/* ecc_b200.h -- C ABI of the B200-native Epipolar-Consistency hot path (libecc_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/Eigen/torch types.  It replaces
 * the host<->device seam of the reference (aaichert/EpipolarConsistency, paths relative to its
 * code/ directory):
 *
 *   void epipolarConsistency(int n_x,int n_y,int num_dtrs,char* dtrs_d,int n_alpha,int n_t,
 *        float step_alpha,float step_t,int num_Ps,float* Cs_d,float* PinvTs_d,int num_pairs,
 *        int* indices_d,float* K01s_d,float* out_d,float object_radius_mm,float dkappa,
 *        bool isDerivative,bool use_corr,float* out_corr_d)
 *        -- LibEpipolarConsistency/EpipolarConsistencyRadonIntermediate.cpp:16-37 (decl.),
 *           EpipolarConsistencyRadonIntermediate.cu:278-409 (def.)
 *   void computeDerivLineIntegrals(cudaTextureObject_t in,int n_x,int n_y,int n_alpha,int n_t,
 *        int filter,int post_process,float* out_d)
 *        -- LibEpipolarConsistency/RadonIntermediate.cpp:12 (decl.), RadonIntermediate.cu:149-170
 *
 * plus the host logic around them that the reference keeps in C++ classes
 * (MetricRadonIntermediate::setProjectionMatrices / setRadonIntermediates / evaluate*,
 * EpipolarConsistencyRadonIntermediate.cpp:87-106,134-163,166-322; RadonIntermediate::compute,
 * RadonIntermediate.cpp:198-211).  The C++ facade in include/EpipolarConsistency/ re-creates
 * those classes on top of this ABI.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative ECC_ERR_* otherwise (the reference calls
 *     exit() on CUDA errors, LibUtilsCuda/UtilsCuda.hxx:14-28); ecc_last_error() gives the text.
 *   - Projection matrices: 3x4, doubles, COLUMN-major (P[r+3c], Eigen default), n of them packed.
 *   - Images: n_v rows of n_u floats.  Radon intermediates ("dtrs"): n_t rows of n_alpha floats,
 *     alpha fastest, bin (ix,iy) <-> alpha=(ix/n_alpha-1/2)pi, t=(iy/n_t-1/2)diag
 *     (RadonIntermediate.cu:44-52).
 *   - Data pointers marked [h|d] may be host or device memory; the library asks the CUDA runtime
 *     which.  Everything else is host memory.
 *   - One context = one GPU + one stream.  Calls on one context must be serialised by the caller.
 *   - There is no CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef ECC_B200_H
#define ECC_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECC_B200_VERSION 100

enum {
    ECC_OK = 0,
    ECC_ERR_CUDA = -1,        /* a CUDA runtime call failed                       */
    ECC_ERR_INVALID = -2,     /* bad argument                                     */
    ECC_ERR_STATE = -3,       /* matrices / dtrs not set, size mismatch           */
    ECC_ERR_UNSUPPORTED = -4  /* a size the implementation does not cover (e.g. ramp filter with n_t > 3000)   */
};

/* RadonIntermediate::Filter / PostProcess, LibEpipolarConsistency/RadonIntermediate.h:21-28.  ECC_FILTER_RAMP: plain
 * line integrals followed by the ramp filter along t (RadonIntermediate.cu:173-237). */
enum { ECC_FILTER_DERIVATIVE = 0, ECC_FILTER_RAMP = 1, ECC_FILTER_NONE = 2 };
enum { ECC_POST_IDENTITY = 0, ECC_POST_SQRT = 1, ECC_POST_LOG = 2 };

/* Bilinear interpolation flavour.
 * ECC_INTERP_TEXTURE: the GPU's texture filter (1.8 fixed-point weights) exactly as the reference
 *   CUDA path uses it (LibUtilsCuda/CudaBindlessTexture.cpp:36-40) -- the drop-in default.
 * ECC_INTERP_EXACT: full fp32 weights (the "CPU float path" numerics).
 * ECC_INTERP_HYBRID (ecc_radon_compute only): the texture filter's arithmetic -- same sample positions, same 1.8
 *   fixed-point weights -- with part of the bins sampled from shared memory instead of through the texture unit, so
 *   that both sampling pipes of an SM work at once.  Bins differ from ECC_INTERP_TEXTURE by the rounding of the
 *   filter's internal sum only (measured <= 3.2e-5 of the peak, tolerance 1e-4).
 * ECC_INTERP_HYBRID_STATIC (ecc_radon_compute only): the hybrid engine with a FIXED assignment of bins to the two
 *   sampling paths (a function of the bin geometry alone, chosen from the bins' sample counts) instead of the run-time
 *   work queue: results are bit-reproducible from run to run and independent of how a data set is batched or sharded
 *   over GPUs; the same speed at the C3 size, at most 3 % slower at the other sizes measured (every projection goes
 *   through the quad kernel, remainders are padded to four). */
enum { ECC_INTERP_TEXTURE = 0, ECC_INTERP_EXACT = 1, ECC_INTERP_HYBRID = 2, ECC_INTERP_HYBRID_STATIC = 3 };

typedef struct ecc_context ecc_context;

int ecc_version(void);

/* Context life cycle.  device < 0 selects the current CUDA device. */
int ecc_create(int device, ecc_context** ctx);
void ecc_destroy(ecc_context* ctx);
const char* ecc_last_error(const ecc_context* ctx);
/* Run on the caller's CUDA stream (cudaStream_t passed as void*); NULL restores the context's own. */
int ecc_set_stream(ecc_context* ctx, void* cuda_stream);
int ecc_synchronize(ecc_context* ctx);

/* ---- Radon intermediates ----------------------------------------------------------------------
 * Batched replacement of RadonIntermediate(image,...)+compute() -> computeDerivLineIntegrals
 * (RadonIntermediate.cpp:17-31,198-211; kernel RadonIntermediate.cu:31-143).
 * images [h|d]: n_images * n_v * n_u floats.  dtrs_out [h|d]: n_images * n_t * n_alpha floats.
 * Asynchronous on the context's stream when both pointers are device memory. */
int ecc_radon_compute(ecc_context* ctx, const float* images, int n_images, int n_u, int n_v,
                      int n_alpha, int n_t, int filter, int post_process, int interp,
                      float* dtrs_out);
/* Bin sizes exactly as RadonIntermediate::compute sets them (RadonIntermediate.cpp:204-206). */
void ecc_radon_bin_sizes(int n_u, int n_v, int n_alpha, int n_t, double* step_alpha,
                         double* step_t);

/* ECC_INTERP_HYBRID_STATIC assigns a fixed share of every projection's samples to the shared-memory path and the rest to
 * the texture unit.  The built-in share (580 / 605 per mille, by window configuration) is where the two pipes of a B200
 * balance; ecc_radon_calibrate_split measures that balance on the GPU at hand for a geometry (a handful of short launches
 * of the run-time work queue, which balances itself, on one quad of empty images; returns per mille), and
 * ecc_radon_set_split pins a share for this context (0 restores the built-in).  Results are bit-reproducible and
 * independent of batching and sharding for a GIVEN share; ranks of a team must use the same one. */
int ecc_radon_calibrate_split(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int* window_share_permille);
int ecc_radon_set_split(ecc_context* ctx, int window_share_permille);

/* Work counter for benchmarks: bilinear image samples per projection for this geometry (exactly what the
 * kernel takes: clipped lines, step 0.66 px, two lines per bin for the derivative filter). */
int ecc_radon_num_samples(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int filter,
                          double* count);

/* ---- Pre-processing (the step in front of the path) -------------------------------------------------------------
 * PreProccess::process + apply_weight_cos_principal_ray (Gui/PreProccess.cpp:57-166; parameter names as the
 * reference's GetSet keys, Gui/PreProccess.cpp:42-55): per pixel v = v*scale + bias (normalize: scale / max of the
 * image, bias 0), optional -log, negatives / NaN / inf -> 0; zeroed + feathered borders (order: left, right, bottom,
 * top); blanked rectangles (x0,y0,x1,y1); flips; separable Gaussian low-pass when sigma > 0 and half width > 1;
 * cosine weighting about the principal point when cos_weight != 0 and matrices are given. */
typedef struct ecc_preprocess_params {
    double scale, bias;           /* Intensity/Scale, Intensity/Bias                       */
    int normalize, apply_log;     /* Intensity/Normalize, Intensity/Apply Minus Logarithm  */
    int border_zero[4];           /* Border/Zero Border: left, right, bottom, top          */
    int border_feather[4];        /* Border/Feather                                        */
    int n_blanks;                 /* Border/Blanks                                         */
    const int* blanks;            /* n_blanks * 4 ints, host                               */
    int flip_u, flip_v;           /* Geometry/Flip u-Axis, Flip v-Axis                     */
    double gaussian_sigma;        /* Lowpass Filter/Gaussian Sigma                         */
    int half_kernel_width;        /* Lowpass Filter/Half Kernel Width                      */
    int cos_weight;               /* apply_weight_cos_principal_ray                        */
} ecc_preprocess_params;
/* The reference's defaults (Gui/PreProccess.h:17-41): scale 1, bias 0, zero border 1 px, feather 16 px, sigma 1.84, k 5;
 * cos_weight on (the loaders always apply it, Gui/InputDataDirect.cpp:85). */
void ecc_preprocess_defaults(ecc_preprocess_params* params);
/* images [h|d]: n * n_v * n_u floats, processed in place.  Ps (host, nullable): n matrices for the cosine weighting. */
int ecc_preprocess(ecc_context* ctx, float* images, int n, int n_u, int n_v, const ecc_preprocess_params* params,
                   const double* Ps);

/* K(0,0), K(0,2), K(1,2) of Geometry::getCameraIntrinsics (LibProjectiveGeometry/ProjectionMatrix.cpp:27-67; RQ decomposition
 * with positive diagonal, K(2,2) = 1): focal length in pixels and principal point, as the cosine weighting uses them. Host only. */
void ecc_camera_intrinsics(const double* P, double* focal_px, double* principal_u, double* principal_v);

/* ---- Metric state (MetricRadonIntermediate) ------------------------------------------------ */

/* setRadonIntermediates (EpipolarConsistencyRadonIntermediate.cpp:87-106).  dtrs [h|d]:
 * n_dtrs * n_t * n_alpha floats.  Device memory is BORROWED (the reference's rule "do not delete
 * or change dtrs during lifetime of Metric", EpipolarConsistencyRadonIntermediate.h:45) unless its
 * alignment forces a private copy; host memory is uploaded into a buffer the context owns. */
int ecc_set_radon_intermediates(ecc_context* ctx, const float* dtrs, int n_dtrs, int n_alpha,
                                int n_t, double step_alpha, double step_t, int n_u, int n_v,
                                int is_derivative);

/* The same for dtrs that live in separate device allocations, one pointer each (the reference keeps a
 * std::vector<RadonIntermediate*>, EpipolarConsistencyRadonIntermediate.h:23): zero copy, every pointer must be
 * 512-byte aligned device memory (ecc_device_alloc) and n_alpha a multiple of 8. dtrs: host array of n_dtrs pointers. */
int ecc_set_radon_intermediate_pointers(ecc_context* ctx, const float* const* dtrs, int n_dtrs, int n_alpha,
                                        int n_t, double step_alpha, double step_t, int n_u, int n_v,
                                        int is_derivative);

/* Device memory for callers without a CUDA toolchain of their own (the C++ facade): allocate / free / copy in
 * any direction (host<->device, device<->device); ecc_copy synchronises. */
int ecc_device_alloc(ecc_context* ctx, size_t bytes, void** ptr);
int ecc_device_free(ecc_context* ctx, void* ptr);
int ecc_copy(ecc_context* ctx, void* dst, const void* src, size_t bytes);

/* UtilsCuda::BindlessTexture2D<float> for callers without a CUDA toolchain (LibUtilsCuda/CudaBindlessTexture.h:18-33, .cpp:17-67):
 * a CUDA array holding a copy of the w x h image [h|d] and a texture object over it -- clamp addressing, linear or point
 * filter, normalised or pixel coordinates.  tex is a cudaTextureObject_t, array a cudaArray_t.  The metric path does not
 * need them (its textures sit over the linear intermediates); they exist for code that reads `->tex` / `->array`. */
int ecc_texture_create(ecc_context* ctx, const float* image, int w, int h, int normalized, int interpolate,
                       unsigned long long* tex, void** array);
int ecc_texture_destroy(ecc_context* ctx, unsigned long long tex, void* array);
/* BindlessTexture2D::readback: the image held by the array, out [h|d] w*h floats. */
int ecc_texture_readback(ecc_context* ctx, void* array, int w, int h, float* out);

/* setProjectionMatrices (EpipolarConsistencyRadonIntermediate.cpp:134-163): pseudo-inverse
 * transposes and source positions are derived on the device in fp64 and stored as fp32. */
int ecc_set_projection_matrices(ecc_context* ctx, const double* Ps, int n);
/* Replace one matrix of the current set (tracking loops change a single view per step,
 * Gui/SingleImageMotion.h:84-90). */
int ecc_update_projection_matrix(ecc_context* ctx, int index, const double* P);

/* The fp32 pseudo-inverse transposes (n*12, 3x4 col-major) and source positions (n*4, C[3]==1) the metric
 * works with, read back from the device (either pointer may be NULL). */
int ecc_get_derived_views(ecc_context* ctx, float* PinvTs, float* Cs);
/* The same derivation on the host (no GPU needed): bit-identical to the device path and to the reference's
 * culaut routines (LibUtilsCuda/culaut/xprojectionmatrix.hxx:20-52,93-105). */
void ecc_derive_views_host(const double* Ps, int n, float* PinvTs, float* Cs);

/* Metric::setObjectRadius / getObjectRadius (EpipolarConsistency.cpp:70-84), 0 = automatic. */
int ecc_set_object_radius(ecc_context* ctx, double radius_mm);
int ecc_get_object_radius(ecc_context* ctx, double* radius_mm);
/* Metric::setEpipolarPlaneStep / setdKappa (EpipolarConsistency.cpp:86-89), 0 = automatic. */
int ecc_set_epipolar_plane_step(ecc_context* ctx, double dkappa_rad);
int ecc_set_interpolation(ecc_context* ctx, int interp);
/* MetricRadonIntermediate::useCorrelation (EpipolarConsistencyRadonIntermediate.h:43): every evaluate* call then scores a
 * pair by 1 - cc, cc the un-centred correlation of the two redundant signals exactly as the reference forms it (weights
 * kappa_max/kappa per sample, EpipolarConsistencyRadonIntermediate.cu:115-149,209,274; .cpp:127-131) instead of the SSD. */
int ecc_use_correlation(ecc_context* ctx, int on);

/* ---- Metric evaluation ------------------------------------------------------------------------
 * evaluate(float* cost_image) (EpipolarConsistencyRadonIntermediate.cpp:166-225): all n(n-1)/2
 * pairs.  cost_image [h|d], nullable: n*n floats, entry i+j*n (i<j) receives the pair's value,
 * all other entries keep the caller's values.  mean (nullable): arithmetic mean over pairs. */
int ecc_evaluate(ecc_context* ctx, float* cost_image, double* mean);

/* Same, restricted to pairs [pair_begin, pair_end) of the get_ij enumeration
 * (EpipolarConsistencyCommon.hxx:52-79); sum = plain sum of those pairs' values.  This is the
 * unit of multi-GPU partitioning. */
int ecc_evaluate_range(ecc_context* ctx, long long pair_begin, long long pair_end,
                       float* cost_image, double* sum);

/* evaluate(indices,out) (EpipolarConsistencyRadonIntermediate.cpp:267-322): idx4 holds
 * (P0,P1,dtr0,dtr1) per pair.  idx4 [h|d]; out [h|d], nullable: n_pairs floats. */
int ecc_evaluate_indices(ecc_context* ctx, const int* idx4, int n_pairs, float* out, double* mean);

/* One step of a tracking / single-view loop (Gui/SingleImageMotion.h:84-90: one matrix changes, the pair list stays):
 * ecc_update_projection_matrix(index, P) followed by ecc_evaluate_indices(idx4, n_pairs, out, mean), with the same
 * results.  From the third call with the same index, list and settings on, the step is replayed as ONE recorded CUDA graph
 * {upload of the derived view, pair kernel, sums, results into pinned host memory} instead of seven API calls (BASELINE
 * config C5: the step is launch-latency bound; see ecc_track_info).  idx4 [h|d]; out: host memory or NULL (device memory takes the plain path). */
int ecc_update_and_evaluate(ecc_context* ctx, int index, const double* P, const int* idx4, int n_pairs, float* out,
                            double* mean);
/* What ecc_update_and_evaluate replays at the moment: 0 = nothing recorded (plain calls), 1 = the fused recording (lists
 * of fewer than 64 pairs per SM: one 80-byte copy and ONE kernel, whose last CTA adds up, delivers values and sum into
 * pinned host memory and raises the flag the host waits on), 2 = the plain recording (two copies, pair kernel, finalize
 * + sum kernel).  The number is the count of kernel launches per call.  kernels_per_call [h]. */
int ecc_track_info(ecc_context* ctx, int* kernels_per_call, long long* replays);

/* Batched mode (new capability): n_sets complete projection-matrix sets
 * (n_sets * n * 12 doubles, n = number of matrices per set = current n of the context) scored
 * against the same dtrs in one launch.  idx4 nullable (all pairs).  out [h|d], nullable:
 * n_sets * n_pairs floats in pair-list order (all pairs: get_ij order).  means: n_sets doubles. */
int ecc_evaluate_batch(ecc_context* ctx, const double* Ps_sets, int n_sets, const int* idx4,
                       int n_pairs, float* out, double* means);

/* ---- Perturbation models of the correction loops, expanded on the device ------------------------------------------
 * A model instance is 11 doubles: ModelCameraSimilarity2D3D (LibProjectiveGeometry/Models/ModelCameraSimilarity2D3D.hxx:89-92),
 * P' = H2D(x[0..3]) * P * T3D(x[4..10]) with x[0..3] = translation u, v, 2D rotation, 2D scale (ModelSimilarity2D.hxx:52-72)
 * and x[4..10] = translation X, Y, Z, rotation about X, Y, Z (R = Rx Ry Rz), 3D scale (ModelSimilarity3D.hxx:64-87).
 * The four host functions and the device expansion execute the same fp64 operation sequence (own sine / cosine, no fused
 * multiply-add): the matrices of ecc_model_expand / ecc_evaluate_batch_params are bit for bit those of
 * ecc_model_camera_similarity_2d3d.  H: 3x3, T: 4x4, P: 3x4, all column-major. */
void ecc_model_similarity_2d(const double* x4, double* H);
void ecc_model_similarity_3d(const double* x7, double* T);
void ecc_model_transform(const double* H, const double* P, const double* T, double* P_out);
void ecc_model_camera_similarity_2d3d(const double* P, const double* x11, double* P_out);

/* Batched mode fed with PARAMETER VECTORS instead of matrices: n_sets * m model instances (params, [h|d], 11 doubles each)
 * applied to the n base matrices (base_Ps [h|d], n*12 doubles, NULL = the context's current set; n = its size).
 * view_to_param [h|d] (n ints, nullable): which of a set's m instances moves view v, negative = the view keeps its base
 * matrix; NULL = every view has its own instance (m == n: ModelFDCT::applyModel, tools/FDCTMotionCorrection/ModelFDCT.hxx:26-62).
 * m = 1 with a map covers the other loops: one moving view (Gui/SingleImageMotion.h:84-90, Gui/FDCTMotionCorrection.hxx:91-97),
 * one transform for a group of views (tools/Registration/Registration3D3D.hxx).  One kernel expands the matrices in
 * registers, derives pinv^T / source positions / each set's automatic object radius, then the sets are scored as by
 * ecc_evaluate_batch -- with the same results as ecc_evaluate_batch fed with the host-expanded matrices. */
int ecc_evaluate_batch_params(ecc_context* ctx, const double* base_Ps, const double* params, int n_sets, int m,
                              const int* view_to_param, const int* idx4, int n_pairs, float* out, double* means);
/* The expanded matrices themselves (Ps_out [h|d]: n_sets * n * 12 doubles), computed by the same device kernel. */
int ecc_model_expand(ecc_context* ctx, const double* base_Ps, const double* params, int n_sets, int m,
                     const int* view_to_param, double* Ps_out);

/* The same with EXPLICIT homographies: an instance is H (3x3) followed by T (4x4), 25 doubles, column-major; P' = H P T,
 * divided by -sign(det M) |m3| when `normalize` is set (Geometry::normalizeProjectionMatrix, ProjectionMatrix.cpp:12-18).
 * view_to_transform NULL: m == n (one instance per view) or m == 1 (ONE correction for the whole trajectory -- the
 * calibration-correction loop: ModelFDCTCalibrationCorrection::transform, Models/ModelFDCTCalibrationCorrection.hxx:136-148;
 * also one rigid transform for all views, tools/Registration/Registration3D3D.hxx).  transforms [h|d]: n_sets * m * 25. */
int ecc_evaluate_batch_transforms(ecc_context* ctx, const double* base_Ps, const double* transforms, int n_sets, int m,
                                  const int* view_to_transform, int normalize, const int* idx4, int n_pairs, float* out,
                                  double* means);
int ecc_transform_expand(ecc_context* ctx, const double* base_Ps, const double* transforms, int n_sets, int m,
                         const int* view_to_transform, int normalize, double* Ps_out);
/* ModelFDCTCalibrationCorrection::getTransforms (Models/ModelFDCTCalibrationCorrection.hxx:150-203): geom4 = the model's mean
 * principal point u, v, source-isocentre and source-detector distance; x7 = translation u, v, yaw, pitch, roll, delta SID,
 * delta SDD.  H 3x3, T 4x4, column-major.  Host function; the fp64 operation sequence is the library's own (shared with
 * nothing on the device: the homographies are inputs of ecc_evaluate_batch_transforms). */
void ecc_model_calibration_correction(const double* geom4, const double* x7, double* H, double* T);
/* Geometry::normalizeProjectionMatrix on the host, with the device path's bits. */
void ecc_model_normalize(double* P);

/* evaluateForImagePair (EpipolarConsistencyRadonIntermediate.cpp:324-393, "visualization only"): the two redundant
 * signals of ONE pair, sampled on the device with the metric's own lookup (the reference walks them on the CPU with
 * RadonIntermediate::sample, whose texel mapping differs slightly from the metric's).  Entries are in
 * ascending kappa, -kappa_max ... +kappa_max, at kappa = +-(m + 1/2) dkappa; entry q belongs to the epipolar lines
 * K0 (cos k, sin k) and K1 (cos k, sin k), whose (l0, l1) go to lines0 / lines1 (2 floats per entry).  All arrays are
 * host memory of `capacity` entries, each nullable; n_samples receives the number of entries the pair has (also when
 * capacity is smaller).  weight = K0[6] * dkappa and value = the pair's metric, so that
 * value = weight * sum_q (signal0[q] - signal1[q])^2 (SSD variant; the reference's own return value is the LAST term
 * only -- `ecc=` for `ecc+=`, SURVEY.md Appendix B -- which is not reproduced). */
int ecc_pair_signals(ecc_context* ctx, int p0, int p1, int dtr0, int dtr1, int capacity, float* kappas, float* signal0,
                     float* signal1, float* lines0, float* lines1, int* n_samples, double* weight, double* value);

/* The reference's K01 record of every pair (EpipolarConsistencyCommon.hxx:92-149, what kernelEpipolarConsistencyComputeK01
 * leaves in K01s, EpipolarConsistencyRadonIntermediate.cu:13-67): 16 floats per pair = K0[0..5] (3x2 col-major map from
 * (cos kappa, sin kappa) to the epipolar line in view 0, image-centre origin), K0[6] baseline distance, K0[7] angle,
 * K1[0..5], K1[6] dkappa, K1[7] kappa_max -- computed on the device with the current matrices and settings, exactly the
 * values the pair kernels work with.  idx4 [h|d] nullable: all pairs in get_ij order.  K01s [h|d]: n_pairs * 16 floats. */
int ecc_pair_maps(ecc_context* ctx, const int* idx4, int n_pairs, float* K01s);

/* Number of kappa samples each pair takes (the work measure for equal-work partitioning), in
 * get_ij order; counts: n(n-1)/2 ints, host. */
int ecc_pair_sample_counts(ecc_context* ctx, int* counts);
/* Cut the pair enumeration into n_parts contiguous ranges of (nearly) equal total kappa samples;
 * bounds: n_parts+1 entries, bounds[0]=0, bounds[n_parts]=n(n-1)/2. */
int ecc_partition_pairs(ecc_context* ctx, int n_parts, long long* bounds);

/* ---- Direct metric: epipolar consistency straight from the projection images (no Radon intermediates) -----------------
 * EpipolarConsistency::MetricDirect and computeForImagePair (LibEpipolarConsistency/EpipolarConsistencyDirect.h:17-60,
 * .cpp:64-270; kernel_computeLineIntegrals / cuda_computeLineIntegrals, EpipolarConsistencyDirect.cu:31-142; the fan-beam
 * variant's weighting, RectifiedFBCC.h).  The reference derives the epipolar lines of ONE pair on the host, uploads them and
 * launches one thread per line, once per image; here all pairs of the data set are one launch (geometry in fp64 on the
 * device, one CTA per 32 epipolar planes of a pair, fixed-order sums).  Uses the context's projection matrices
 * (ecc_set_projection_matrices), object radius (0: estimated from the FIRST matrix, Metric::getObjectRadius) and epipolar
 * plane step (0: per pair, half the kappa range over the image diagonal, EpipolarConsistencyDirect.cpp:91-96). */

/* MetricDirect::setProjectionImages: n pre-processed projections [h|d] of n_v rows x n_u floats, copied into CUDA arrays
 * behind pixel-coordinate, linear-filter, clamp textures (BindlessTexture2D<float>'s defaults). */
int ecc_direct_set_images(ecc_context* ctx, const float* images, int n, int n_u, int n_v);
/* The same for images that live in separate allocations (the reference keeps a std::vector<BindlessTexture2D<float>*>,
 * EpipolarConsistencyDirect.h:31): images = host array of n pointers, each [h|d]. */
int ecc_direct_set_image_pointers(ecc_context* ctx, const float* const* images, int n, int n_u, int n_v);
/* MetricDirect::setFanBeamConsistency: the rectified fan-beam consistency (weighted integrals) instead of the derivative. */
int ecc_direct_set_fan_beam(ecc_context* ctx, int fbcc);
/* The reference's launcher hands the kernel n_u for BOTH image sizes (EpipolarConsistencyDirect.cu:135), i.e. it clips
 * every line against an n_u x n_u box.  Off (default): lines are clipped against the image.  On: as the reference. */
int ecc_direct_set_reference_clip(ecc_context* ctx, int on);
/* MetricDirect::evaluate (EpipolarConsistencyDirect.cpp:236-247): all pairs i < j.  cost_image [h|d], nullable: n*n floats,
 * entry i + j*n receives the pair's value, other entries keep the caller's.  sum: the SUM over the pairs (the direct
 * metric returns the sum, not the mean). */
int ecc_direct_evaluate(ecc_context* ctx, float* cost_image, double* sum);
/* The same restricted to pairs [pair_begin, pair_end) of that enumeration (i < j, i outer): the unit of multi-GPU sharding --
 * pairs are independent, every GPU holds the images and matrices, the only exchange is the final sum.  sum = over those pairs. */
int ecc_direct_evaluate_range(ecc_context* ctx, long long pair_begin, long long pair_end, float* cost_image, double* sum);
/* Cuts the enumeration into n_parts contiguous ranges of (nearly) equal work = epipolar planes (the range of kappa, and with
 * a fixed plane step the number of planes, differs from pair to pair); bounds: n_parts + 1 entries, host. */
int ecc_direct_partition(ecc_context* ctx, int n_parts, long long* bounds);
/* MetricDirect::evaluateForImagePair / computeForImagePair (.cpp:64-212, 249-258): one pair and its redundant signals.
 * kappas / samples0 / samples1: host arrays of `capacity` floats, each nullable.  n_given > 0: the first n_given entries of
 * kappas are the caller's epipolar plane angles (the reference's "unless provided"), else they receive the angles used.
 * n_lines: how many planes the pair has (also when capacity is smaller).  value = sum_q (s0[q] - s1[q])^2 * dkappa. */
int ecc_direct_evaluate_pair(ecc_context* ctx, int i, int j, int n_given, int capacity, float* kappas, float* samples0,
                             float* samples1, int* n_lines, double* value);
/* The host-side geometry of one pair as the reference's computeForImagePair prepares it for its kernel, from the same
 * code the device kernel runs: kappas (capacity floats), the corresponding epipolar lines of both views in Hessian normal
 * form (3 floats per plane) and the FBCC_weighting_info records (RectifiedFBCC.h:82-87: 8 floats per plane and view). Every
 * array is host memory and nullable.  No GPU needed beyond the context. */
int ecc_direct_pair_geometry(ecc_context* ctx, int i, int j, int capacity, float* kappas, float* lines0, float* lines1,
                             float* fbcc0, float* fbcc1, int* n_lines, double* dkappa);
/* The launcher-level seam, cuda_computeLineIntegrals (EpipolarConsistencyDirect.cu:122-142): integrals of image `image`
 * (index into the set images) along n_lines given lines.  lines [h|d]: line_stride floats per line, the first three =
 * (l0, l1, l2) in Hessian normal form; fbcc [h|d], nullable: fbcc_stride floats per line, the first six = the
 * FBCC_weighting_info record; out [h|d]: n_lines floats. */
int ecc_direct_line_integrals(ecc_context* ctx, int image, const float* lines, int n_lines, int line_stride,
                              const float* fbcc, int fbcc_stride, float* out);

/* ---- Multi-GPU team: the GPUs of one node on one data set, one process (or thread) and one context per GPU ----------
 * The reference is single-GPU.  A team shards the path with NO collective on the data path: every
 * rank owns one device block [flags | pair values | the Radon intermediates of all n_total projections] that is mapped
 * into all other ranks (CUDA IPC over NVLink/NVSwitch); the Radon kernels store every finished bin into all blocks, the
 * pair values are published the same way, and flag barriers in peer memory order the stages.  Results (mean, cost
 * image) are identical on every rank and bit-identical to a single GPU's for the same intermediates.
 *   rank r:  ecc_team_create(...) -> exchange the handles (any transport: MPI, torch.distributed, a pipe)
 *            -> ecc_team_connect(all handles) -> per step: ecc_team_radon_compute(own block of projections)
 *            -> ecc_set_radon_intermediates(dtrs of ecc_team_block) once -> ecc_set_projection_matrices
 *            -> ecc_team_evaluate.
 * Every rank must make the same sequence of ecc_team_radon_compute / ecc_team_evaluate / ecc_team_barrier calls. */
#define ECC_TEAM_HANDLE_BYTES 64
/* Allocates this rank's block for n_total projections of n_t x n_alpha bins; handle_out (ECC_TEAM_HANDLE_BYTES, nullable
 * when world == 1) identifies it to the other processes.  world <= 16.  A team of one needs no connect. */
int ecc_team_create(ecc_context* ctx, int rank, int world, int n_total, int n_alpha, int n_t, unsigned char* handle_out);
/* handles: world * ECC_TEAM_HANDLE_BYTES bytes in rank order (the own entry is ignored). */
int ecc_team_connect(ecc_context* ctx, const unsigned char* handles);
/* The same for ranks that live in ONE process (one context per GPU, or several on one GPU): blocks[r] = rank r's block
 * (ecc_team_block).  Team calls are then made from one thread per rank, or in any order that keeps every rank's stream
 * fed: the barriers wait on the device, not on the host. */
int ecc_team_connect_pointers(ecc_context* ctx, void* const* blocks);
/* block: start of the own block; dtrs: the n_total * n_t * n_alpha floats inside it (either may be NULL). */
int ecc_team_block(ecc_context* ctx, void** block, float** dtrs);
int ecc_team_destroy(ecc_context* ctx);
/* Radon intermediates of projections [first, first + n_local) of the data set from images [h|d] (n_local * n_v * n_u
 * floats), stored into every rank's block, followed by a team barrier on the stream: work queued after this call sees
 * the intermediates of ALL ranks.  Asynchronous for device images. */
int ecc_team_radon_compute(ecc_context* ctx, const float* images, int first, int n_local, int n_u, int n_v, int filter,
                           int post_process, int interp);
/* Sharding in QUADS of projections for the static-split engine (ECC_INTERP_HYBRID_STATIC works on four interleaved
 * projections at a time; a shard that is not a whole number of quads pads its last one -- 496 projections on 8 GPUs are 15.5
 * quads per rank, 3 % of the step).  ecc_team_radon_shard cuts the ceil(n_total / 4) quads into `world` equal intervals; the
 * quad an interval boundary falls into is computed by BOTH neighbours, each taking its share (lo_num / den .. hi_num / den)
 * of the quad's bins -- which bins follows from the geometry alone, every bin is computed by exactly one rank, with the
 * arithmetic a single GPU would use, and stored into all ranks' blocks.  A pure function (no context): rank r supplies
 * projections [first, first + count) -- up to three more than n_total / world -- to ecc_team_radon_compute_part. */
int ecc_team_radon_shard(int n_total, int world, int rank, int* first, int* count, int* lo_num, int* hi_num, int* den);
int ecc_team_radon_compute_part(ecc_context* ctx, const float* images, int first, int n_local, int lo_num, int hi_num, int den, int n_u,
                                int n_v, int filter, int post_process, int interp);
/* All pairs, cut into world ranges of equal kappa-sample count (ecc_partition_pairs); this rank scores range [rank].
 * cost_image [h|d], nullable: as ecc_evaluate, complete on every rank.  mean: over all pairs, the same on every rank. */
int ecc_team_evaluate(ecc_context* ctx, float* cost_image, double* mean);
/* A barrier of its own on the stream (e.g. before a rank overwrites data its peers may still read). */
int ecc_team_barrier(ecc_context* ctx);

/* ---- Host helpers that the reference keeps next to the metric ------------------------------ */

/* ProjTable::makeCircularTrajectory (HeaderOnly/Utils/Projtable.hxx:138-165). Ps: n*12 doubles. */
void ecc_make_circular_trajectory(int n_proj, double sid, double sdd, int n_u, int n_v,
                                  double max_angle_deg, double pixel_spacing, double* Ps);
/* Synthetic cone-beam projections of ellipsoids (7 doubles each: centre, semi-axes, density) with
 * cosine weighting (Gui/PreProccess.cpp:146-166) and zeroed border; images [d]: n*n_v*n_u. */
int ecc_synth_projections(ecc_context* ctx, const double* Ps, int n, int n_u, int n_v,
                          const double* ellipsoids, int n_ellipsoids, int cos_weight,
                          int zero_border, float* images);

/* ---- Instrumentation ------------------------------------------------------------------------
 * With profiling on, every kernel launch is bracketed by CUDA events on the context's stream.
 * ecc_profile_get sums them up per kernel family ("radon", "pairs", "geometry", "reduce",
 * "synth").  Querying synchronises the stream. */
int ecc_profile_enable(ecc_context* ctx, int on);
int ecc_profile_reset(ecc_context* ctx);
int ecc_profile_get(ecc_context* ctx, const char* family, double* total_ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* ECC_B200_H */
