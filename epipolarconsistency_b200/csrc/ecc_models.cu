// ecc_models.cu -- perturbation models on the device: K parameter vectors in, K matrix sets scored, no host round trip.
//
// The reference's correction loops (Gui/FDCTMotionCorrection.hxx:13-98 with tools/FDCTMotionCorrection/ModelFDCT.hxx:26-62,
// Gui/SingleImageMotion.h, Gui/Registration.h, tools/Registration/Registration3D3D.hxx) turn a parameter vector into n
// projection matrices on the host (P' = H2D P T3D, Models/ModelCameraSimilarity2D3D.hxx:89-92), derive pinv^T / source
// position for all of them on the host and upload both per evaluation.  Here a launch takes K x m parameter vectors
// (11 doubles each) and one base matrix set; expand_derive_kernel builds each view's matrix in registers (ecc_models.cuh:
// the host models' bits), derives its pinv^T and source position (derive_view: the reference's culaut bits) and the set's
// automatic object radius, and the pair kernel scores all K sets -- matrices never exist in memory.
#include <cstring>

#include "ecc_internal.h"
#include "ecc_models.cuh"

using namespace eccb200;

namespace {

// thread = (set, view).  view_to_param (nullable = identity): which of the set's m parameter vectors moves this view, < 0 = none.
// Ps_out (nullable): the expanded matrices, [n_sets][n][12].
// kind 0: an instance is the 11 parameters of ModelCameraSimilarity2D3D; kind 1: an instance is an explicit pair of
// homographies, H (9 doubles) then T (16 doubles), column-major -- P' = H P T, normalised (Geometry::normalizeProjectionMatrix)
// when `normalize` is set (ModelFDCTCalibrationCorrection::transform).  map_all0: no view map and m == 1 -> every view takes
// instance 0 (one correction for the whole trajectory).
__global__ void expand_derive_kernel(const double* __restrict__ base, const double* __restrict__ params, const int* __restrict__ view_to_param,
                                     int n, int m, int n_sets, int n_u, int n_v, double fixed_radius, float* __restrict__ PinvTs,
                                     float* __restrict__ Cs, float* __restrict__ radii, double* __restrict__ Ps_out, int kind = 0,
                                     int normalize = 0, int map_all0 = 0)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)n_sets * n) return;
    const int set = (int)(k / n), view = (int)(k - (long long)set * n);
    double P[12];
    for (int q = 0; q < 12; q++) P[q] = base[(size_t)12 * view + q];
    const int pi = view_to_param ? view_to_param[view] : (map_all0 ? 0 : view);
    if (pi >= 0 && pi < m) {
        if (kind == 0) {
            double x[11];
            const double* src = params + ((size_t)set * m + pi) * 11;
            for (int q = 0; q < 11; q++) x[q] = src[q];
            model_camera_similarity_2d3d(P, x, P);
        } else {
            double HT[25];
            const double* src = params + ((size_t)set * m + pi) * 25;
            for (int q = 0; q < 25; q++) HT[q] = src[q];
            model_transform(HT, P, HT + 9, P);
            if (normalize) model_normalize(P);
        }
    }
    if (Ps_out)
        for (int q = 0; q < 12; q++) Ps_out[(size_t)12 * k + q] = P[q];
    if (PinvTs) {
        float A[12], C[4];
        derive_view(P, A, C);
        for (int q = 0; q < 12; q++) PinvTs[(size_t)12 * k + q] = A[q];
        for (int q = 0; q < 4; q++) Cs[(size_t)4 * k + q] = C[q];
        if (radii && view == 0) radii[set] = (float)(fixed_radius > 0 ? fixed_radius : object_radius_from_view(P, n_u, n_v));
    }
}

struct DeviceGuard {
    explicit DeviceGuard(ecc_context* c) { cudaSetDevice(c->device); }
};

// Uploads base matrices (NULL = the context's current set), parameter vectors and the view map into the batch buffers.
int stage_models(ecc_context* ctx, const double* base_Ps, const double* params, int n_sets, int m, const int* view_to_param,
                 const double** base_d, const double** params_d, const int** map_d, int stride = 11)
{
    BatchBuffers& B = ctx->batch;
    const int n = ctx->n_views;
    const size_t p_bytes = sizeof(double) * stride * (size_t)n_sets * m + sizeof(int) * (size_t)n;
    size_t cap = B.params_cap;
    int rc = ensure_bytes(ctx, (void**)&B.params_d, &cap, p_bytes);
    B.params_cap = cap;
    if (rc) return rc;
    cap = B.base_cap;
    rc = ensure_bytes(ctx, (void**)&B.base_d, &cap, sizeof(double) * 12 * (size_t)n);
    B.base_cap = cap;
    if (rc) return rc;
    ECC_CUDA(ctx, cudaMemcpyAsync(B.params_d, params, sizeof(double) * stride * (size_t)n_sets * m, cudaMemcpyDefault, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(B.base_d, base_Ps ? base_Ps : ctx->Ps_h.data(), sizeof(double) * 12 * (size_t)n, cudaMemcpyDefault, ctx->stream));
    *map_d = nullptr;
    if (view_to_param) {
        int* dst = (int*)(B.params_d + stride * (size_t)n_sets * m);
        ECC_CUDA(ctx, cudaMemcpyAsync(dst, view_to_param, sizeof(int) * (size_t)n, cudaMemcpyDefault, ctx->stream));
        *map_d = dst;
    }
    *base_d = B.base_d;
    *params_d = B.params_d;
    return ECC_OK;
}

int check_models(ecc_context* ctx, const double* params, int n_sets, int m, const int* view_to_param, const char* who, bool one_for_all = false)
{
    if (!params || n_sets < 1 || m < 1) return fail(ctx, ECC_ERR_INVALID, std::string(who) + ": bad argument");
    if (ctx->n_views <= 0) return fail(ctx, ECC_ERR_STATE, "projection matrices not set");
    if (!view_to_param && m != ctx->n_views && !(one_for_all && m == 1))
        return fail(ctx, ECC_ERR_INVALID, std::string(who) + ": without a view map every view needs its own parameter vector (m == number of matrices)");
    if (view_to_param && !is_device_pointer(view_to_param))
        for (int v = 0; v < ctx->n_views; v++)
            if (view_to_param[v] >= m) return fail(ctx, ECC_ERR_INVALID, std::string(who) + ": view map refers to a parameter vector that does not exist");
    return ECC_OK;
}

}  // namespace

extern "C" {

void ecc_model_similarity_2d(const double* x, double* H)
{
    if (x && H) model_similarity_2d(x, H);
}

void ecc_model_similarity_3d(const double* x, double* T)
{
    if (x && T) model_similarity_3d(x, T);
}

void ecc_model_transform(const double* H, const double* P, const double* T, double* P_out)
{
    if (H && P && T && P_out) model_transform(H, P, T, P_out);
}

void ecc_model_camera_similarity_2d3d(const double* P, const double* x, double* P_out)
{
    if (P && x && P_out) model_camera_similarity_2d3d(P, x, P_out);
}

int ecc_model_expand(ecc_context* ctx, const double* base_Ps, const double* params, int n_sets, int m, const int* view_to_param,
                     double* Ps_out)
{
    if (!ctx || !Ps_out) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    int rc = check_models(ctx, params, n_sets, m, view_to_param, "ecc_model_expand");
    if (rc) return rc;
    const int n = ctx->n_views;
    const size_t count = (size_t)n_sets * n;
    const double *base_d, *params_d;
    const int* map_d;
    if ((rc = stage_models(ctx, base_Ps, params, n_sets, m, view_to_param, &base_d, &params_d, &map_d))) return rc;
    if ((rc = batch_reserve(ctx, n_sets, true))) return rc;
    const bool out_dev = is_device_pointer(Ps_out);
    double* dst = out_dev ? Ps_out : ctx->batch.Ps_d;
    const int slot = prof_begin(ctx, FAM_GEOMETRY);
    expand_derive_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(base_d, params_d, map_d, n, m, n_sets, ctx->n_u, ctx->n_v,
                                                                                 ctx->object_radius, nullptr, nullptr, nullptr, dst);
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    if (!out_dev) ECC_CUDA(ctx, cudaMemcpyAsync(Ps_out, dst, sizeof(double) * 12 * count, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

int ecc_evaluate_batch_params(ecc_context* ctx, const double* base_Ps, const double* params, int n_sets, int m, const int* view_to_param,
                              const int* idx4, int n_pairs, float* out, double* means)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    int rc = check_models(ctx, params, n_sets, m, view_to_param, "ecc_evaluate_batch_params");
    if (rc) return rc;
    PairLaunch L;
    if ((rc = batch_begin(ctx, idx4, n_pairs, L, "ecc_evaluate_batch_params"))) return rc;
    const int n = ctx->n_views;
    const size_t count = (size_t)n_sets * n;
    const double *base_d, *params_d;
    const int* map_d;
    if ((rc = stage_models(ctx, base_Ps, params, n_sets, m, view_to_param, &base_d, &params_d, &map_d))) return rc;
    if ((rc = batch_reserve(ctx, n_sets, false))) return rc;
    BatchBuffers& B = ctx->batch;
    const int slot = prof_begin(ctx, FAM_GEOMETRY);
    expand_derive_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(base_d, params_d, map_d, n, m, n_sets, ctx->n_u, ctx->n_v,
                                                                                 ctx->object_radius, B.A_d, B.Cs_d, B.radii_d, nullptr);
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    return batch_finish(ctx, L, n_sets, out, means);
}

void ecc_model_calibration_correction(const double* geom4, const double* x7, double* H, double* T)
{
    if (geom4 && x7 && H && T) model_calibration_correction(geom4, x7, H, T);
}

void ecc_model_normalize(double* P)
{
    if (P) model_normalize(P);
}

int ecc_transform_expand(ecc_context* ctx, const double* base_Ps, const double* transforms, int n_sets, int m, const int* view_to_transform,
                         int normalize, double* Ps_out)
{
    if (!ctx || !Ps_out) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    int rc = check_models(ctx, transforms, n_sets, m, view_to_transform, "ecc_transform_expand", true);
    if (rc) return rc;
    const int n = ctx->n_views;
    const size_t count = (size_t)n_sets * n;
    const double *base_d, *params_d;
    const int* map_d;
    if ((rc = stage_models(ctx, base_Ps, transforms, n_sets, m, view_to_transform, &base_d, &params_d, &map_d, 25))) return rc;
    if ((rc = batch_reserve(ctx, n_sets, true))) return rc;
    const bool out_dev = is_device_pointer(Ps_out);
    double* dst = out_dev ? Ps_out : ctx->batch.Ps_d;
    const int slot = prof_begin(ctx, FAM_GEOMETRY);
    expand_derive_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(base_d, params_d, map_d, n, m, n_sets, ctx->n_u, ctx->n_v,
                                                                                 ctx->object_radius, nullptr, nullptr, nullptr, dst, 1,
                                                                                 normalize ? 1 : 0, (!view_to_transform && m == 1 && n != 1) ? 1 : 0);
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    if (!out_dev) ECC_CUDA(ctx, cudaMemcpyAsync(Ps_out, dst, sizeof(double) * 12 * count, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

int ecc_evaluate_batch_transforms(ecc_context* ctx, const double* base_Ps, const double* transforms, int n_sets, int m,
                                  const int* view_to_transform, int normalize, const int* idx4, int n_pairs, float* out, double* means)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    int rc = check_models(ctx, transforms, n_sets, m, view_to_transform, "ecc_evaluate_batch_transforms", true);
    if (rc) return rc;
    PairLaunch L;
    if ((rc = batch_begin(ctx, idx4, n_pairs, L, "ecc_evaluate_batch_transforms"))) return rc;
    const int n = ctx->n_views;
    const size_t count = (size_t)n_sets * n;
    const double *base_d, *params_d;
    const int* map_d;
    if ((rc = stage_models(ctx, base_Ps, transforms, n_sets, m, view_to_transform, &base_d, &params_d, &map_d, 25))) return rc;
    if ((rc = batch_reserve(ctx, n_sets, false))) return rc;
    BatchBuffers& B = ctx->batch;
    const int slot = prof_begin(ctx, FAM_GEOMETRY);
    expand_derive_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(base_d, params_d, map_d, n, m, n_sets, ctx->n_u, ctx->n_v,
                                                                                 ctx->object_radius, B.A_d, B.Cs_d, B.radii_d, nullptr, 1,
                                                                                 normalize ? 1 : 0, (!view_to_transform && m == 1 && n != 1) ? 1 : 0);
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    return batch_finish(ctx, L, n_sets, out, means);
}

}  // extern "C"
