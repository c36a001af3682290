// ecc_compat.cu -- the reference's two launcher symbols with THEIR OWN signatures, backed by the sm_100a kernels.
//
// The reference's host classes reach the device through exactly two free functions with C++ linkage, declared `extern`
// in its .cpp files and defined in its .cu files:
//   void epipolarConsistency(int,int,int,char*,int,int,float,float,int,float*,float*,int,int*,float*,float*,float,float,
//                            bool,bool,float*)      LibEpipolarConsistency/EpipolarConsistencyRadonIntermediate.cpp:16-37 (decl.),
//                                                   EpipolarConsistencyRadonIntermediate.cu:278-409 (def.)
//   void computeDerivLineIntegrals(cudaTextureObject_t,int,int,int,int,int,int,float*)
//                                                   LibEpipolarConsistency/RadonIntermediate.cpp:12 (decl.), RadonIntermediate.cu:149-170 (def.)
// Linking the reference's UNMODIFIED EpipolarConsistencyRadonIntermediate.cpp / RadonIntermediate.cpp against libecc_b200.so
// instead of its own two .cu files resolves these symbols here ("launcher swap", INTEGRATION.md option B;
// tests/cpp/launcher_swap_check.cpp calls them exactly as those files do).  Same arguments, same buffers written:
//   * out_d: all pairs -> entry i + j*n (i<j) of the n x n image, other entries untouched; pair list -> out[p];
//   * K01s_d: the 16-float record of every pair (kernelEpipolarConsistencyComputeK01's output);
//   * out_corr_d: SSD -> one weight 1.0 per pair; correlation -> six sums per pair (x, y, xx, yy, xy, weight 1.0),
//     from which the reference's host code forms 1 - cc (.cpp:127-131,200-224,304-321).
// Work is issued on the legacy default stream and the call returns after it has finished, as the reference's launchers do
// (cudaDeviceSynchronize after each kernel).  Errors print and exit(), the reference's convention (UtilsCuda.hxx:14-28).
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#include "ecc_geometry.cuh"
#include "ecc_internal.h"

using namespace eccb200;

namespace {

constexpr int kMaxDevices = 64;

// One context per device, created on first use, working on the legacy default stream (ordered with the caller's own
// default-stream work, which is all the reference ever uses).
ecc_context* compat_context()
{
    static ecc_context* table[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    if (!table[dev]) {
        ecc_context* c = nullptr;
        if (ecc_create(dev, &c) != ECC_OK) return nullptr;
        ecc_set_stream(c, (void*)cudaStreamLegacy);
        table[dev] = c;
    }
    return table[dev];
}

[[noreturn]] void die(ecc_context* ctx, const char* where, int rc)
{
    std::fprintf(stderr, "libecc_b200 (%s): error %d: %s\n", where, rc, ctx ? ecc_last_error(ctx) : "no CUDA device (there is no CPU fallback)");
    std::exit(rc ? rc : 1);
}

int compat_interp()
{
    // Radon engine of the launcher-compatible entry point (one projection per call, as the reference calls it).
    // Default: the texture engine -- bit-identical to the reference's kernel and the fastest engine for a SINGLE image
    // (1.47 ms at 1240x960 -> 768x768; the quad kernel would pad one image to four).  ECC_COMPAT_RADON=hybrid selects the
    // one-image two-pipe kernel (0.82 ms, bins within 3.2e-5 of the peak), =exact the fp32-weight kernel.
    static const int v = [] {
        const char* e = getenv("ECC_COMPAT_RADON");
        if (!e) return (int)ECC_INTERP_TEXTURE;
        const std::string s(e);
        if (s == "hybrid") return (int)ECC_INTERP_HYBRID;
        if (s == "hybrid-static") return (int)ECC_INTERP_HYBRID_STATIC;
        if (s == "exact") return (int)ECC_INTERP_EXACT;
        return (int)ECC_INTERP_TEXTURE;
    }();
    return v;
}

}  // namespace

// ---- C++ linkage on purpose: these are the reference's symbols ---------------------------------------------------------

void computeDerivLineIntegrals(cudaTextureObject_t in, int n_x, int n_y, int n_alpha, int n_t, int filter, int post_process, float* out_d)
{
    ecc_context* ctx = compat_context();
    if (!ctx) die(nullptr, "computeDerivLineIntegrals", ECC_ERR_CUDA);
    cudaSetDevice(ctx->device);
    // the image behind the texture object: a CUDA array in the reference (BindlessTexture2D), linear / pitched memory accepted too
    cudaResourceDesc res = {};
    if (cudaGetTextureObjectResourceDesc(&res, in) != cudaSuccess) {
        fail(ctx, ECC_ERR_INVALID, "computeDerivLineIntegrals: not a texture object");
        die(ctx, "computeDerivLineIntegrals", ECC_ERR_INVALID);
    }
    const size_t bytes = sizeof(float) * (size_t)n_x * n_y;
    int rc = ensure_bytes(ctx, (void**)&ctx->img_stage_d, &ctx->img_stage_bytes, bytes);
    if (rc) die(ctx, "computeDerivLineIntegrals", rc);
    const float* image = ctx->img_stage_d;
    cudaError_t e = cudaSuccess;
    if (res.resType == cudaResourceTypeArray) {
        e = cudaMemcpy2DFromArrayAsync(ctx->img_stage_d, sizeof(float) * n_x, res.res.array.array, 0, 0, sizeof(float) * n_x, n_y,
                                       cudaMemcpyDeviceToDevice, ctx->stream);
    } else if (res.resType == cudaResourceTypePitch2D) {
        if (res.res.pitch2D.pitchInBytes == sizeof(float) * (size_t)n_x) image = (const float*)res.res.pitch2D.devPtr;
        else e = cudaMemcpy2DAsync(ctx->img_stage_d, sizeof(float) * n_x, res.res.pitch2D.devPtr, res.res.pitch2D.pitchInBytes,
                                   sizeof(float) * n_x, n_y, cudaMemcpyDeviceToDevice, ctx->stream);
    } else if (res.resType == cudaResourceTypeLinear) {
        image = (const float*)res.res.linear.devPtr;
    } else {
        fail(ctx, ECC_ERR_UNSUPPORTED, "computeDerivLineIntegrals: texture over a mip-mapped array");
        die(ctx, "computeDerivLineIntegrals", ECC_ERR_UNSUPPORTED);
    }
    if (e != cudaSuccess) die(ctx, "computeDerivLineIntegrals", cuda_fail(ctx, e, "image copy", __FILE__, __LINE__));
    rc = radon_batch(ctx, image, 1, n_x, n_y, n_alpha, n_t, filter, post_process, compat_interp(), out_d);
    if (rc) die(ctx, "computeDerivLineIntegrals", rc);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) die(ctx, "computeDerivLineIntegrals", ECC_ERR_CUDA);
}

void epipolarConsistency(int n_x, int n_y, int num_dtrs, char* dtrs_d, int n_alpha, int n_t, float step_alpha, float step_t, int num_Ps,
                         float* Cs_d, float* PinvTs_d, int num_pairs, int* indices_d, float* K01s_d, float* out_d, float object_radius_mm,
                         float dkappa, bool isDerivative, bool use_corr, float* out_corr_d)
{
    (void)step_alpha;
    ecc_context* ctx = compat_context();
    if (!ctx) die(nullptr, "epipolarConsistency", ECC_ERR_CUDA);
    cudaSetDevice(ctx->device);
    const bool all = (indices_d == nullptr);
    const long long pairs = all ? (long long)num_dtrs * (num_dtrs - 1) / 2 : (long long)num_pairs;
    if (pairs <= 0) return;
    PairLaunch L;
    L.n_views = all ? num_dtrs : num_Ps;  // all pairs: the reference enumerates i<j<num_dtrs (.cu:229-231)
    L.n_sets = 1;
    L.pair_begin = 0;
    L.n_pairs = pairs;
    L.idx4_d = indices_d;
    L.Cs_d = Cs_d;
    L.PinvTs_d = PinvTs_d;
    L.tex_d = (const cudaTextureObject_t*)dtrs_d;  // normalised, linear, clamp: what RadonIntermediate::getTexture builds
    L.dtr_ptrs_d = nullptr;
    L.dtr_pitch = 0;
    L.n_dtrs = num_dtrs;
    L.n_alpha = n_alpha;
    L.n_t = n_t;
    L.half_nu = n_x * 0.5f;
    L.half_nv = n_y * 0.5f;
    L.range_t = n_t * step_t;
    L.image_diagonal = n_t * step_t * 2.f;
    L.radius = object_radius_mm;
    L.dkappa = dkappa;
    L.radii_d = nullptr;
    const int max_samples = (dkappa <= 0.f) ? (int)L.image_diagonal : (int)(ECC_PI_F * 0.5f / dkappa);
    L.sample_cap = (max_samples + 255) / 256 * 256;
    L.is_derivative = isDerivative ? 1 : 0;
    L.interp = ECC_INTERP_TEXTURE;
    L.use_corr = use_corr ? 1 : 0;
    L.mode_items = 0;
    L.defer_finalize = 0;
    L.splits = 1;
    L.partials_d = nullptr;
    L.corr_sums_d = use_corr ? out_corr_d : nullptr;
    L.image_d = nullptr;
    int rc = ECC_OK;
    if (indices_d && (uintptr_t)indices_d % 16 != 0) {  // the kernels read a pair's four indices as one 16-byte load
        size_t cap = ctx->idx_cap * sizeof(int);
        rc = ensure_bytes(ctx, (void**)&ctx->idx_d, &cap, sizeof(int) * 4 * (size_t)pairs);
        ctx->idx_cap = cap / sizeof(int);
        if (rc) die(ctx, "epipolarConsistency", rc);
        cudaMemcpyAsync(ctx->idx_d, indices_d, sizeof(int) * 4 * (size_t)pairs, cudaMemcpyDeviceToDevice, ctx->stream);
        L.idx4_d = ctx->idx_d;
    }
    if (all) {  // compact values into scratch, scattered into the caller's n x n image by the same kernel
        size_t cap = ctx->vals_cap * sizeof(float);
        rc = ensure_bytes(ctx, (void**)&ctx->vals_d, &cap, sizeof(float) * (size_t)pairs);
        ctx->vals_cap = cap / sizeof(float);
        if (rc) die(ctx, "epipolarConsistency", rc);
        L.vals_d = ctx->vals_d;
        L.image_d = out_d;
    } else {
        L.vals_d = out_d;
    }
    if (K01s_d && (rc = launch_pair_maps(ctx, L, K01s_d))) die(ctx, "epipolarConsistency", rc);
    if ((rc = launch_pairs(ctx, L))) die(ctx, "epipolarConsistency", rc);
    // weights: 1.0 per pair (.cu:188,254); the correlation kernel writes them with its sums
    if (!use_corr && out_corr_d && (rc = launch_fill(ctx, out_corr_d, (size_t)pairs, 1, 1.0f))) die(ctx, "epipolarConsistency", rc);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) die(ctx, "epipolarConsistency", ECC_ERR_CUDA);
}

// ---- the reference's texture wrapper for callers without a CUDA toolchain (the C++ facade's UtilsCuda::BindlessTexture2D) -------
extern "C" {

int ecc_texture_create(ecc_context* ctx, const float* image, int w, int h, int normalized, int interpolate, unsigned long long* tex_out,
                       void** array_out)
{
    if (!ctx || !image || !tex_out || !array_out || w < 1 || h < 1) return ECC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    cudaArray_t arr = nullptr;
    ECC_CUDA(ctx, cudaMallocArray(&arr, &desc, w, h));
    cudaError_t e = cudaMemcpy2DToArrayAsync(arr, 0, 0, image, sizeof(float) * w, sizeof(float) * w, h, cudaMemcpyDefault, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFreeArray(arr);
        return cuda_fail(ctx, e, "cudaMemcpy2DToArrayAsync", __FILE__, __LINE__);
    }
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td = {};
    td.normalizedCoords = normalized ? 1 : 0;
    td.filterMode = interpolate ? cudaFilterModeLinear : cudaFilterModePoint;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    e = cudaCreateTextureObject(&tex, &res, &td, nullptr);
    if (e != cudaSuccess) {
        cudaFreeArray(arr);
        return cuda_fail(ctx, e, "cudaCreateTextureObject", __FILE__, __LINE__);
    }
    *tex_out = (unsigned long long)tex;
    *array_out = (void*)arr;
    return ECC_OK;
}

int ecc_texture_destroy(ecc_context* ctx, unsigned long long tex, void* array)
{
    if (!ctx) return ECC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (tex) ECC_CUDA(ctx, cudaDestroyTextureObject((cudaTextureObject_t)tex));
    if (array) ECC_CUDA(ctx, cudaFreeArray((cudaArray_t)array));
    return ECC_OK;
}

int ecc_texture_readback(ecc_context* ctx, void* array, int w, int h, float* out)
{
    if (!ctx || !array || !out || w < 1 || h < 1) return ECC_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ECC_CUDA(ctx, cudaMemcpy2DFromArrayAsync(out, sizeof(float) * w, (cudaArray_t)array, 0, 0, sizeof(float) * w, h, cudaMemcpyDefault, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

}  // extern "C"
