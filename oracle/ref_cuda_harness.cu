// ref_cuda_harness.cu -- drives the REFERENCE's own CUDA launchers (compiled unchanged from
// /root/reference/code/LibEpipolarConsistency/{RadonIntermediate,EpipolarConsistencyRadonIntermediate,EpipolarConsistencyDirect}.cu
// by oracle/Makefile) so that tests and bench.py can run "the reference CUDA path" on the GPU box.
// Ours: this harness only.  It re-creates, without Eigen/GetSet, the ~100 lines of host logic that sit
// between the reference's class API and its two launchers:
//   RadonIntermediate ctor + compute      RadonIntermediate.cpp:17-31,198-211
//   BindlessTexture2D ctor                LibUtilsCuda/CudaBindlessTexture.cpp:17-44 (array, linear, clamp)
//   setProjectionMatrices                 EpipolarConsistencyRadonIntermediate.cpp:134-163 (culaut, included verbatim)
//   evaluate / evaluate(indices)          EpipolarConsistencyRadonIntermediate.cpp:166-225,267-322
// Test infrastructure only; the product never links this.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using std::abs;
#include <LibUtilsCuda/culaut/xprojectionmatrix.hxx>
#include <LibEpipolarConsistency/EpipolarConsistencyCommon.hxx>  // get_ij (pair enumeration), header-only

// the reference's launchers (C++ linkage, defined in the reference .cu files)
void computeDerivLineIntegrals(cudaTextureObject_t in, int n_x, int n_y, int n_alpha, int n_t, int filter,
                               int post_process, float* out_d);
void epipolarConsistency(int n_x, int n_y, int num_dtrs, char* dtrs_d, int n_alpha, int n_t, float step_alpha,
                         float step_t, int num_Ps, float* Cs_d, float* PinvTs_d, int num_pairs, int* indices_d,
                         float* K01s_d, float* out_d, float object_radius_mm, float dkappa, bool isDerivative,
                         bool use_corr, float* out_corr_d);

// EpipolarConsistencyDirect.cu:122-142 (C++ linkage, defined in the reference's .cu)
void cuda_computeLineIntegrals(short n_lines, float* lines_d, short line_stride, float* fbcc_d, short fbcc_stride, cudaTextureObject_t I,
                               short n_u, short n_v, float* integrals_out_d);

namespace {

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            fprintf(stderr, "ref harness CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return -1;                                                                         \
        }                                                                                      \
    } while (0)

struct ArrayTex {
    cudaArray_t arr = nullptr;
    cudaTextureObject_t tex = 0;
};

// Same settings as the reference's BindlessTexture2D(w,h,buffer,device,interpolate=true,normalized)
int make_array_texture(const float* src_d, int w, int h, bool normalized, ArrayTex& out)
{
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    CK(cudaMallocArray(&out.arr, &desc, w, h));
    CK(cudaMemcpy2DToArray(out.arr, 0, 0, src_d, sizeof(float) * w, sizeof(float) * w, h, cudaMemcpyDeviceToDevice));
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray;
    res.res.array.array = out.arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.normalizedCoords = normalized;
    td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = cudaAddressModeClamp;
    td.addressMode[1] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    CK(cudaCreateTextureObject(&out.tex, &res, &td, NULL));
    return 0;
}

void free_array_texture(ArrayTex& t)
{
    if (t.tex) cudaDestroyTextureObject(t.tex);
    if (t.arr) cudaFreeArray(t.arr);
    t = ArrayTex();
}

typedef void (*LauncherFn)(int, int, int, char*, int, int, float, float, int, float*, float*, int, int*, float*, float*, float, float, bool, bool,
                           float*);

struct RefMetric {
    LauncherFn launcher = nullptr;       // diagnostics: another implementation of the launcher's signature (null = the reference's)
    std::vector<float*> linear;          // diagnostics: dtrs kept in linear memory when the textures are pitch2D
    int n_views = 0, n_dtrs = 0, n_alpha = 0, n_t = 0, n_u = 0, n_v = 0;
    float step_alpha = 0, step_t = 0;
    bool is_derivative = true;
    std::vector<ArrayTex> dtrs;
    cudaTextureObject_t* tex_d = nullptr;
    float *Cs_d = nullptr, *PinvTs_d = nullptr, *K01s_d = nullptr, *out_d = nullptr, *corr_d = nullptr;
    int* idx_d = nullptr;
    size_t k01_cap = 0, out_cap = 0, corr_cap = 0, idx_cap = 0;
};

}  // namespace

extern "C" {

// Radon intermediates of n images with the reference kernel.  Host or device pointers (cudaMemcpyDefault).  ms (nullable):
// GPU time of the launcher calls only (texture set-up excluded), via CUDA events.
int ref_cuda_radon(const float* images_h, int n_images, int n_u, int n_v, int n_alpha, int n_t, int filter,
                   int post, float* dtrs_h, float* ms)
{
    const size_t img = (size_t)n_u * n_v, dtr = (size_t)n_alpha * n_t;
    float *img_d = nullptr, *out_d = nullptr;
    CK(cudaMalloc(&img_d, sizeof(float) * img));
    CK(cudaMalloc(&out_d, sizeof(float) * dtr));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float total = 0.f;
    for (int k = 0; k < n_images; k++) {
        CK(cudaMemcpy(img_d, images_h + img * k, sizeof(float) * img, cudaMemcpyDefault));
        ArrayTex t;
        if (make_array_texture(img_d, n_u, n_v, false, t)) return -1;
        CK(cudaEventRecord(e0));
        computeDerivLineIntegrals(t.tex, n_u, n_v, n_alpha, n_t, filter, post, out_d);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float m = 0.f;
        CK(cudaEventElapsedTime(&m, e0, e1));
        total += m;
        CK(cudaMemcpy(dtrs_h + dtr * k, out_d, sizeof(float) * dtr, cudaMemcpyDefault));
        free_array_texture(t);
    }
    if (ms) *ms = total;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(img_d);
    cudaFree(out_d);
    return 0;
}

int ref_cuda_radon_any(const void* images, int n_images, int n_u, int n_v, int n_alpha, int n_t, int filter, int post,
                       void* dtrs, float* ms)
{
    return ref_cuda_radon((const float*)images, n_images, n_u, n_v, n_alpha, n_t, filter, post, (float*)dtrs, ms);
}

// The direct metric's line kernel (the reference's kernel_computeLineIntegrals through its own launcher) on one image.
// All pointers host or device (cudaMemcpyDefault); n_lines < 32768 (the launcher's `short`).  fbcc nullable.  ms (nullable):
// GPU time of the launcher call.
int ref_cuda_direct_line_integrals(const float* image, int n_u, int n_v, const float* lines, int n_lines, int line_stride,
                                   const float* fbcc, int fbcc_stride, float* out, float* ms)
{
    if (n_lines <= 0 || n_lines > 32767) return -2;
    const size_t img = (size_t)n_u * n_v;
    float *img_d = nullptr, *lines_d = nullptr, *fbcc_d = nullptr, *out_d = nullptr;
    CK(cudaMalloc(&img_d, sizeof(float) * img));
    CK(cudaMemcpy(img_d, image, sizeof(float) * img, cudaMemcpyDefault));
    ArrayTex t;
    if (make_array_texture(img_d, n_u, n_v, false, t)) return -1;  // BindlessTexture2D<float>(w, h, buffer): pixel coordinates
    CK(cudaMalloc(&lines_d, sizeof(float) * n_lines * line_stride));
    CK(cudaMemcpy(lines_d, lines, sizeof(float) * n_lines * line_stride, cudaMemcpyDefault));
    if (fbcc) {
        CK(cudaMalloc(&fbcc_d, sizeof(float) * n_lines * fbcc_stride));
        CK(cudaMemcpy(fbcc_d, fbcc, sizeof(float) * n_lines * fbcc_stride, cudaMemcpyDefault));
    }
    CK(cudaMalloc(&out_d, sizeof(float) * n_lines));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    cuda_computeLineIntegrals((short)n_lines, lines_d, (short)line_stride, fbcc_d, (short)fbcc_stride, t.tex, (short)n_u, (short)n_v, out_d);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    if (ms) CK(cudaEventElapsedTime(ms, e0, e1));
    CK(cudaMemcpy(out, out_d, sizeof(float) * n_lines, cudaMemcpyDefault));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    free_array_texture(t);
    cudaFree(img_d);
    cudaFree(lines_d);
    if (fbcc_d) cudaFree(fbcc_d);
    cudaFree(out_d);
    return 0;
}

// Metric object: dtrs become array textures exactly like RadonIntermediate::getTexture() (normalised).
void* ref_cuda_metric_create(const float* dtrs_h, int n_dtrs, int n_alpha, int n_t, float step_alpha, float step_t,
                             int n_u, int n_v, int is_derivative)
{
    RefMetric* M = new RefMetric();
    M->n_dtrs = n_dtrs; M->n_alpha = n_alpha; M->n_t = n_t; M->n_u = n_u; M->n_v = n_v;
    M->step_alpha = step_alpha; M->step_t = step_t; M->is_derivative = is_derivative != 0;
    const size_t dtr = (size_t)n_alpha * n_t;
    float* tmp_d = nullptr;
    if (cudaMalloc(&tmp_d, sizeof(float) * dtr) != cudaSuccess) return nullptr;
    std::vector<cudaTextureObject_t> handles(n_dtrs);
    M->dtrs.resize(n_dtrs);
    for (int k = 0; k < n_dtrs; k++) {
        cudaMemcpy(tmp_d, dtrs_h + dtr * k, sizeof(float) * dtr, cudaMemcpyDefault);  // host or device dtrs
        if (make_array_texture(tmp_d, n_alpha, n_t, true, M->dtrs[k])) return nullptr;
        handles[k] = M->dtrs[k].tex;
    }
    cudaFree(tmp_d);
    cudaMalloc(&M->tex_d, sizeof(cudaTextureObject_t) * n_dtrs);
    cudaMemcpy(M->tex_d, handles.data(), sizeof(cudaTextureObject_t) * n_dtrs, cudaMemcpyHostToDevice);
    return M;
}

void* ref_cuda_metric_create_any(const void* dtrs, int n_dtrs, int n_alpha, int n_t, float step_alpha, float step_t, int n_u,
                                 int n_v, int is_derivative)
{
    return ref_cuda_metric_create((const float*)dtrs, n_dtrs, n_alpha, n_t, step_alpha, step_t, n_u, n_v, is_derivative);
}

// Diagnostics: the same metric object with its dtr textures over pitched LINEAR memory (cudaResourceTypePitch2D, what
// libecc_b200 samples) instead of CUDA arrays -- to tell texture-type effects from arithmetic effects.
void* ref_cuda_metric_create_pitch2d(const void* dtrs, int n_dtrs, int n_alpha, int n_t, float step_alpha, float step_t, int n_u, int n_v,
                                     int is_derivative)
{
    RefMetric* M = new RefMetric();
    M->n_dtrs = n_dtrs; M->n_alpha = n_alpha; M->n_t = n_t; M->n_u = n_u; M->n_v = n_v;
    M->step_alpha = step_alpha; M->step_t = step_t; M->is_derivative = is_derivative != 0;
    const size_t dtr = (size_t)n_alpha * n_t;
    std::vector<cudaTextureObject_t> handles(n_dtrs);
    M->dtrs.resize(n_dtrs);
    for (int k = 0; k < n_dtrs; k++) {
        float* lin = nullptr;
        if (cudaMalloc(&lin, sizeof(float) * dtr) != cudaSuccess) return nullptr;
        cudaMemcpy(lin, (const float*)dtrs + dtr * k, sizeof(float) * dtr, cudaMemcpyDefault);
        M->linear.push_back(lin);
        cudaResourceDesc res;
        memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypePitch2D;
        res.res.pitch2D.devPtr = lin;
        res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
        res.res.pitch2D.width = n_alpha;
        res.res.pitch2D.height = n_t;
        res.res.pitch2D.pitchInBytes = sizeof(float) * n_alpha;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.normalizedCoords = 1;
        td.filterMode = cudaFilterModeLinear;
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        if (cudaCreateTextureObject(&M->dtrs[k].tex, &res, &td, NULL) != cudaSuccess) return nullptr;
        handles[k] = M->dtrs[k].tex;
    }
    cudaMalloc(&M->tex_d, sizeof(cudaTextureObject_t) * n_dtrs);
    cudaMemcpy(M->tex_d, handles.data(), sizeof(cudaTextureObject_t) * n_dtrs, cudaMemcpyHostToDevice);
    return M;
}

// Diagnostics: run evaluate() through ANOTHER function with the launcher's signature (e.g. libecc_b200's own
// epipolarConsistency symbol) on this object's textures and buffers; null restores the reference's launcher.
void ref_cuda_metric_set_launcher(void* h, void* fn) { ((RefMetric*)h)->launcher = (LauncherFn)fn; }

void ref_cuda_metric_destroy(void* h)
{
    RefMetric* M = (RefMetric*)h;
    if (!M) return;
    for (float* p : M->linear) cudaFree(p);
    for (auto& t : M->dtrs) free_array_texture(t);
    cudaFree(M->tex_d); cudaFree(M->Cs_d); cudaFree(M->PinvTs_d); cudaFree(M->K01s_d); cudaFree(M->out_d);
    cudaFree(M->corr_d); cudaFree(M->idx_d);
    delete M;
}

// setProjectionMatrices with the reference's culaut routines (double -> float), then upload.
int ref_cuda_metric_set_matrices(void* h, const double* Ps, int n)
{
    RefMetric* M = (RefMetric*)h;
    std::vector<float> A((size_t)12 * n), Cs((size_t)4 * n);
    for (int i = 0; i < n; i++) {
        culaut::projection_matrix_pseudoinverse_transpose<double, float>(Ps + 12 * i, &A[12 * i]);
        culaut::projection_matrix_source_position<double, float>(Ps + 12 * i, &Cs[4 * i]);
    }
    if (n != M->n_views) {
        cudaFree(M->Cs_d); cudaFree(M->PinvTs_d);
        CK(cudaMalloc(&M->Cs_d, sizeof(float) * 4 * n));
        CK(cudaMalloc(&M->PinvTs_d, sizeof(float) * 12 * n));
        M->n_views = n;
    }
    CK(cudaMemcpy(M->PinvTs_d, A.data(), sizeof(float) * 12 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(M->Cs_d, Cs.data(), sizeof(float) * 4 * n, cudaMemcpyHostToDevice));
    return 0;
}

static int grow(float** p, size_t* cap, size_t n)
{
    if (*cap >= n) return 0;
    cudaFree(*p);
    CK(cudaMalloc(p, sizeof(float) * n));
    *cap = n;
    return 0;
}

// evaluate(): all pairs (idx4 == NULL; out_h = n*n cost image, pre-zeroed by the caller if wanted) or a
// pair list (out_h = n_pairs).  `out` is zeroed on the device before the launch: the reference zeroes it
// inside the accumulating kernel, which races -- the in-kernel store then writes
// zero over zero or over a partial sum; see tests for how the comparison copes.  Returns the mean.
// ms (nullable): GPU time of the launcher call (both kernels + its two device syncs).
double ref_cuda_metric_evaluate(void* h, const int* idx4, int n_pairs, float radius, float dkappa, float* out_h,
                                float* ms)
{
    RefMetric* M = (RefMetric*)h;
    const int n = M->n_views;
    const bool all = (idx4 == nullptr);
    const int pairs = all ? n * (n - 1) / 2 : n_pairs;
    const size_t out_len = all ? (size_t)n * n : (size_t)pairs;
    // One record more than the list has, zeroed: the reference's index-list kernel lets thread idx_x == num_pairs through
    // ("if (idx_x>num_pairs) return;", EpipolarConsistencyRadonIntermediate.cu:165), which reads K01s, the index quad and
    // the texture handle of a pair BEHIND the list and stores to out / out_corr there.  With zeros it finds dkappa = kappa_max
    // = 0 and leaves; with whatever the allocator left behind the buffers it faulted (CUDA error 700) once the library's
    // own allocations changed.
    if (grow(&M->K01s_d, &M->k01_cap, (size_t)(pairs + 1) * 16)) return -1;
    if (grow(&M->out_d, &M->out_cap, out_len + 1)) return -1;
    if (grow(&M->corr_d, &M->corr_cap, (size_t)pairs + 1)) return -1;
    cudaMemset(M->K01s_d, 0, sizeof(float) * 16 * (pairs + 1));
    cudaMemset(M->out_d, 0, sizeof(float) * (out_len + 1));
    cudaMemset(M->corr_d, 0, sizeof(float) * (pairs + 1));
    if (!all) {
        if (M->idx_cap < (size_t)(pairs + 1) * 4) {
            cudaFree(M->idx_d);
            cudaMalloc(&M->idx_d, sizeof(int) * 4 * (pairs + 1));
            M->idx_cap = (size_t)(pairs + 1) * 4;
        }
        cudaMemset(M->idx_d, 0, sizeof(int) * 4 * (pairs + 1));
        cudaMemcpy(M->idx_d, idx4, sizeof(int) * 4 * pairs, cudaMemcpyHostToDevice);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    (M->launcher ? M->launcher : (LauncherFn)epipolarConsistency)(M->n_u, M->n_v, M->n_dtrs, (char*)M->tex_d, M->n_alpha, M->n_t, M->step_alpha,
                                                                   M->step_t, n, M->Cs_d, M->PinvTs_d, all ? 0 : pairs,
                                                                   all ? nullptr : M->idx_d, M->K01s_d, M->out_d, radius, dkappa,
                                                                   M->is_derivative, false, M->corr_d);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float m = 0.f;
    cudaEventElapsedTime(&m, e0, e1);
    if (ms) *ms = m;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::vector<float> out(out_len);
    cudaMemcpy(out.data(), M->out_d, sizeof(float) * out_len, cudaMemcpyDeviceToHost);
    double sum = 0;
    if (all) {
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) {
                const float v = out[(size_t)i + (size_t)j * n];
                sum += v;
                if (out_h) out_h[(size_t)i + (size_t)j * n] = v;
            }
    } else {
        for (int k = 0; k < pairs; k++) {
            sum += out[k];
            if (out_h) out_h[k] = out[k];
        }
    }
    return pairs ? sum / pairs : 0.0;
}

// The K01 records the reference's kernelEpipolarConsistencyComputeK01 left on the device at the last evaluate call
// (16 floats per pair, EpipolarConsistencyCommon.hxx:92-149).  Returns the number of records copied.
int ref_cuda_metric_get_k01(void* h, float* K01s_h, int n_pairs)
{
    RefMetric* M = (RefMetric*)h;
    if (!M || !M->K01s_d || (size_t)n_pairs * 16 > M->k01_cap) return -1;
    CK(cudaMemcpy(K01s_h, M->K01s_d, sizeof(float) * 16 * (size_t)n_pairs, cudaMemcpyDeviceToHost));
    return n_pairs;
}

// evaluate() with useCorrelation(true): the launcher accumulates six floats per pair (five weighted sums + the pair
// weight, EpipolarConsistencyRadonIntermediate.cu:115-149,182-189), the host forms 1 - cc per pair
// (EpipolarConsistencyRadonIntermediate.cpp:127-131,200-211,304-308) and the weighted mean (all weights are 1).
double ref_cuda_metric_evaluate_corr(void* h, const int* idx4, int n_pairs, float radius, float dkappa, float* out_h)
{
    RefMetric* M = (RefMetric*)h;
    const int n = M->n_views;
    const bool all = (idx4 == nullptr);
    const int pairs = all ? n * (n - 1) / 2 : n_pairs;
    const size_t out_len = all ? (size_t)n * n : (size_t)pairs;
    // one zeroed record more than the list has: see ref_cuda_metric_evaluate
    if (grow(&M->K01s_d, &M->k01_cap, (size_t)(pairs + 1) * 16)) return -1;
    if (grow(&M->out_d, &M->out_cap, out_len + 1)) return -1;
    if (grow(&M->corr_d, &M->corr_cap, (size_t)(pairs + 1) * 6)) return -1;
    cudaMemset(M->K01s_d, 0, sizeof(float) * 16 * (pairs + 1));
    cudaMemset(M->out_d, 0, sizeof(float) * (out_len + 1));
    cudaMemset(M->corr_d, 0, sizeof(float) * (pairs + 1) * 6);
    if (!all) {
        if (M->idx_cap < (size_t)(pairs + 1) * 4) {
            cudaFree(M->idx_d);
            cudaMalloc(&M->idx_d, sizeof(int) * 4 * (pairs + 1));
            M->idx_cap = (size_t)(pairs + 1) * 4;
        }
        cudaMemset(M->idx_d, 0, sizeof(int) * 4 * (pairs + 1));
        cudaMemcpy(M->idx_d, idx4, sizeof(int) * 4 * pairs, cudaMemcpyHostToDevice);
    }
    epipolarConsistency(M->n_u, M->n_v, M->n_dtrs, (char*)M->tex_d, M->n_alpha, M->n_t, M->step_alpha, M->step_t, n,
                        M->Cs_d, M->PinvTs_d, all ? 0 : pairs, all ? nullptr : M->idx_d, M->K01s_d, M->out_d, radius,
                        dkappa, M->is_derivative, true, M->corr_d);
    std::vector<float> sums((size_t)pairs * 6);
    cudaMemcpy(sums.data(), M->corr_d, sizeof(float) * pairs * 6, cudaMemcpyDeviceToHost);
    double weighted = 0, weights = 0;
    for (int k = 0; k < pairs; k++) {
        const float* q = &sums[(size_t)6 * k];
        const float corr = q[4] / (sqrt(q[2]) * sqrt(q[3]));
        const float weight = q[5];
        const float v = (1.0f - corr) * weight;
        if (out_h) {
            if (all) {
                short i, j;
                get_ij(k, n, i, j);
                out_h[(size_t)i + (size_t)j * n] = v;
            } else
                out_h[k] = v;
        }
        weighted += v;
        weights += weight;
    }
    return weights > 0 ? weighted / weights : 0.0;
}

}  // extern "C"
