// ecc_capi.cu -- the C ABI of libecc_b200 (include/ecc_b200.h): context, host logic, launches.
//
// Host logic restated from the reference's C++ classes (code/LibEpipolarConsistency/):
//   MetricRadonIntermediate::setRadonIntermediates   EpipolarConsistencyRadonIntermediate.cpp:87-106
//   MetricRadonIntermediate::setProjectionMatrices   EpipolarConsistencyRadonIntermediate.cpp:134-163
//   MetricRadonIntermediate::evaluate (3 overloads)  EpipolarConsistencyRadonIntermediate.cpp:166-322
//   Metric::getObjectRadius / estimateObjectRadius   EpipolarConsistency.cpp:35-47,76-84
//   RadonIntermediate::compute                       RadonIntermediate.cpp:198-211
// Differences by design: all state lives in one context per GPU, every copy and launch is
// asynchronous on the context's stream with one synchronisation at the end of a call that returns
// a value, errors are returned instead of exit().
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "ecc_geometry.cuh"
#include "ecc_internal.h"

using namespace eccb200;

namespace eccb200 {

int fail(ecc_context* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->last_error = msg;
    return code;
}

int cuda_fail(ecc_context* ctx, cudaError_t e, const char* what, const char* file, int line)
{
    std::string m = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) +
                    ") in " + what + " at " + file + ":" + std::to_string(line);
    return fail(ctx, ECC_ERR_CUDA, m);
}

bool is_device_pointer(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure_bytes(ecc_context* ctx, void** ptr, size_t* cap, size_t bytes)
{
    if (*cap >= bytes && *ptr) return ECC_OK;
    if (*ptr) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ECC_CUDA(ctx, cudaFree(*ptr));
        *ptr = nullptr;
        *cap = 0;
    }
    ECC_CUDA(ctx, cudaMalloc(ptr, bytes ? bytes : 16));
    *cap = bytes ? bytes : 16;
    return ECC_OK;
}

int ensure_pinned(ecc_context* ctx, size_t bytes)
{
    if (ctx->pinned_bytes >= bytes && ctx->pinned_h) return ECC_OK;
    if (ctx->pinned_h) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ECC_CUDA(ctx, cudaFreeHost(ctx->pinned_h));
        ctx->pinned_h = nullptr;
        ctx->pinned_bytes = 0;
    }
    ECC_CUDA(ctx, cudaHostAlloc(&ctx->pinned_h, bytes ? bytes : 16, cudaHostAllocDefault));
    ctx->pinned_bytes = bytes ? bytes : 16;
    return ECC_OK;
}

int prof_begin(ecc_context* ctx, int family)
{
    if (!ctx->profiling) return -1;
    ProfileSlot s;
    s.family = family;
    if (cudaEventCreate(&s.start) != cudaSuccess || cudaEventCreate(&s.stop) != cudaSuccess) return -1;
    cudaEventRecord(s.start, ctx->stream);
    ctx->prof_slots.push_back(s);
    return (int)ctx->prof_slots.size() - 1;
}

void prof_end(ecc_context* ctx, int slot)
{
    if (slot < 0) return;
    cudaEventRecord(ctx->prof_slots[slot].stop, ctx->stream);
}

void prof_collect(ecc_context* ctx)
{
    if (ctx->prof_slots.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->prof_slots) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.start, s.stop) == cudaSuccess) {
            ctx->prof_ms[s.family] += ms;
            ctx->prof_launches[s.family] += 1;
        }
        cudaEventDestroy(s.start);
        cudaEventDestroy(s.stop);
    }
    ctx->prof_slots.clear();
}

}  // namespace eccb200

namespace {

struct Guard {  // make the context's device current for the duration of a call
    explicit Guard(ecc_context* c) { cudaSetDevice(c->device); }
};

size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

void destroy_dtr_textures(ecc_context* ctx)
{
    for (auto t : ctx->dtr_tex_h) cudaDestroyTextureObject(t);
    ctx->dtr_tex_h.clear();
    for (auto a : ctx->dtr_arrays) cudaFreeArray(a);
    ctx->dtr_arrays.clear();
}

// Fills the launch record with everything that depends only on the context state.
int fill_launch(ecc_context* ctx, PairLaunch& L)
{
    if (ctx->n_views <= 0) return fail(ctx, ECC_ERR_STATE, "projection matrices not set");
    if (ctx->n_dtrs <= 0 || !ctx->dtr_ptrs_d) return fail(ctx, ECC_ERR_STATE, "Radon intermediates not set");
    double radius = 0;
    int rc = ecc_get_object_radius(ctx, &radius);
    if (rc) return rc;
    L.n_views = ctx->n_views;
    L.n_sets = 1;
    L.pair_begin = 0;
    L.n_pairs = 0;
    L.idx4_d = nullptr;
    L.Cs_d = ctx->Cs_d;
    L.PinvTs_d = ctx->PinvTs_d;
    L.tex_d = ctx->dtr_tex_d;
    L.dtr_ptrs_d = ctx->dtr_ptrs_d;
    L.dtr_pitch = ctx->dtr_pitch;
    L.n_dtrs = ctx->n_dtrs;
    L.n_alpha = ctx->n_alpha;
    L.n_t = ctx->n_t;
    L.half_nu = ctx->n_u * 0.5f;
    L.half_nv = ctx->n_v * 0.5f;
    // launcher sizing as in the reference (EpipolarConsistencyRadonIntermediate.cu:320,347-358)
    L.range_t = ctx->n_t * ctx->step_t;
    L.image_diagonal = ctx->n_t * ctx->step_t * 2.f;
    L.radius = (float)radius;
    L.dkappa = (float)ctx->dkappa;
    L.radii_d = nullptr;
    const int max_samples = (L.dkappa <= 0.f) ? (int)L.image_diagonal : (int)(ECC_PI_F * 0.5f / L.dkappa);
    L.sample_cap = (max_samples + 255) / 256 * 256;
    L.is_derivative = ctx->is_derivative;
    L.interp = ctx->interp;
    L.use_corr = ctx->use_corr;
    L.vals_d = nullptr;
    L.image_d = nullptr;
    L.splits = 1;
    L.mode_items = 0;
    L.defer_finalize = 0;
    L.partials_d = nullptr;
    L.corr_sums_d = nullptr;
    return ECC_OK;
}

int upload_and_derive(ecc_context* ctx, const double* Ps, size_t count, double** Ps_d, float** Cs_d,
                      float** A_d, size_t* cap, int views_per_set = 0, float* radii_d = nullptr)
{
    if (*cap < count) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (*Ps_d) cudaFree(*Ps_d);
        if (*Cs_d) cudaFree(*Cs_d);
        if (*A_d) cudaFree(*A_d);
        *Ps_d = nullptr; *Cs_d = nullptr; *A_d = nullptr; *cap = 0;
        ECC_CUDA(ctx, cudaMalloc(Ps_d, sizeof(double) * 12 * count));
        ECC_CUDA(ctx, cudaMalloc(Cs_d, sizeof(float) * 4 * count));
        ECC_CUDA(ctx, cudaMalloc(A_d, sizeof(float) * 12 * count));
        *cap = count;
    }
    ECC_CUDA(ctx, cudaMemcpyAsync(*Ps_d, Ps, sizeof(double) * 12 * count, cudaMemcpyDefault, ctx->stream));
    return launch_derive_views(ctx, *Ps_d, (int)count, *A_d, *Cs_d, views_per_set, ctx->n_u, ctx->n_v,
                               ctx->object_radius, radii_d);
}

}  // namespace

extern "C" {

static void track_free(ecc_context* ctx);

int ecc_version(void) { return ECC_B200_VERSION; }

int ecc_create(int device, ecc_context** out)
{
    if (!out) return ECC_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return ECC_ERR_CUDA;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return ECC_ERR_CUDA;
    if (device >= count) return ECC_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return ECC_ERR_CUDA;
    ecc_context* ctx = new ecc_context();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return ECC_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = ctx;
    return ECC_OK;
}

void ecc_destroy(ecc_context* ctx)
{
    if (!ctx) return;
    Guard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    prof_collect(ctx);
    destroy_dtr_textures(ctx);
    free_image_pool(ctx);
    free_hybrid(ctx);
    free_hybrid4(ctx);
    team_free(ctx);
    track_free(ctx);
    free_direct(ctx);
    if (ctx->ramp_g_d) cudaFree(ctx->ramp_g_d);
    if (ctx->pre_work_d) cudaFree(ctx->pre_work_d);
    if (ctx->pre_small_d) cudaFree(ctx->pre_small_d);
    if (ctx->copy_stream) {
        cudaStreamDestroy(ctx->copy_stream);
        for (int b = 0; b < 2; b++) { cudaEventDestroy(ctx->ev_copied[b]); cudaEventDestroy(ctx->ev_consumed[b]); }
    }
    if (ctx->down_stream) {
        cudaStreamDestroy(ctx->down_stream);
        for (int b = 0; b < 2; b++) { cudaEventDestroy(ctx->ev_out_ready[b]); cudaEventDestroy(ctx->ev_out_free[b]); }
    }
    void* bufs[] = {ctx->batch.Ps_d, ctx->batch.Cs_d, ctx->batch.A_d, ctx->batch.radii_d, ctx->batch.params_d, ctx->batch.base_d, ctx->Ps_d, ctx->Cs_d, ctx->PinvTs_d, ctx->dtrs_owned, ctx->dtr_tex_d, (void*)ctx->dtr_ptrs_d, ctx->vals_d, ctx->partials_d, ctx->pair_records_d,
                    ctx->sums_d, ctx->idx_d, ctx->counts_d, ctx->img_stage_d, ctx->out_stage_d, ctx->cost_d};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (ctx->pinned_h) cudaFreeHost(ctx->pinned_h);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* ecc_last_error(const ecc_context* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

int ecc_set_stream(ecc_context* ctx, void* cuda_stream)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    cudaStreamSynchronize(ctx->stream);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return ECC_OK;
}

int ecc_synchronize(ecc_context* ctx)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

void ecc_radon_bin_sizes(int n_u, int n_v, int n_alpha, int n_t, double* step_alpha, double* step_t)
{
    const double diagonal = std::sqrt((double)n_v * n_v + (double)n_u * n_u);
    if (step_t) *step_t = diagonal / n_t;
    if (step_alpha) *step_alpha = 3.1415926535897931 / n_alpha;
}

int ecc_radon_compute(ecc_context* ctx, const float* images, int n_images, int n_u, int n_v, int n_alpha,
                      int n_t, int filter, int post, int interp, float* dtrs_out)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    return eccb200::radon_compute_impl(ctx, images, n_images, n_u, n_v, n_alpha, n_t, filter, post, interp, dtrs_out, true);
}

}  // extern "C"

int eccb200::fill_pair_launch(ecc_context* ctx, PairLaunch& L) { return fill_launch(ctx, L); }

int eccb200::radon_compute_impl(ecc_context* ctx, const float* images, int n_images, int n_u, int n_v, int n_alpha, int n_t,
                                int filter, int post, int interp, float* dtrs_out, bool final_sync, QuadPart part)
{
    if (!images || !dtrs_out || n_images < 0 || n_u < 2 || n_v < 2 || n_alpha < 1 || n_t < 1)
        return fail(ctx, ECC_ERR_INVALID, "ecc_radon_compute: bad argument");
    if (filter != ECC_FILTER_DERIVATIVE && filter != ECC_FILTER_NONE && filter != ECC_FILTER_RAMP) return fail(ctx, ECC_ERR_INVALID, "bad filter");
    if (post < 0 || post > 2) return fail(ctx, ECC_ERR_INVALID, "bad post_process");
    if (interp != ECC_INTERP_TEXTURE && interp != ECC_INTERP_EXACT && interp != ECC_INTERP_HYBRID && interp != ECC_INTERP_HYBRID_STATIC)
        return fail(ctx, ECC_ERR_INVALID, "bad interp");
    if (n_images == 0) return ECC_OK;
    const bool in_dev = is_device_pointer(images), out_dev = is_device_pointer(dtrs_out);
    const size_t img_elems = (size_t)n_u * n_v, dtr_elems = (size_t)n_alpha * n_t;
    if (in_dev && out_dev) return radon_batch(ctx, images, n_images, n_u, n_v, n_alpha, n_t, filter, post, interp, dtrs_out, part);
    // Host memory on either side: stream the batch through device staging in chunks.  Host images go through two
    // staging buffers on a copy stream of their own, so that the upload of chunk i+1 runs under the kernels of chunk i
    // (the first chunk is small: its upload is the only one that is exposed).
    // Chunk schedule for host images: 4, 8, 16, 32, 64, 128, 128, ... -- the first upload (one quad of projections) is the
    // only exposed one, every later upload (at most twice the images of the chunk computing above it) hides as long as the
    // link delivers 17 GB/s, and the persistent kernel gets few, large launches (1.2 GB of staging at the C3 image size).
    // ECC_FIRST_CHUNK: development knob (a multiple of four: chunks are whole quads of projections).
    const int chunk = n_images < 128 ? n_images : 128;
    static const int first_env = getenv("ECC_FIRST_CHUNK") ? atoi(getenv("ECC_FIRST_CHUNK")) : 4;
    const int first_want = first_env >= 4 ? first_env / 4 * 4 : 4;
    static const int growth_env = getenv("ECC_CHUNK_GROWTH") ? atoi(getenv("ECC_CHUNK_GROWTH")) : 0;  // development knob
    const int growth = growth_env >= 2 ? growth_env : 2;
    const int first_chunk = (!in_dev && n_images > first_want) ? first_want : chunk;
    int rc;
    if (!in_dev && (rc = ensure_bytes(ctx, (void**)&ctx->img_stage_d, &ctx->img_stage_bytes, sizeof(float) * img_elems * chunk * 2))) return rc;
    // Host memory for the intermediates: two staging buffers and a download stream, so that the download of chunk i runs under
    // the kernels of chunk i+1 (the Radon kernels leave DRAM and the copy engines idle); only the last chunk's download is
    // exposed.
    if (!out_dev && (rc = ensure_bytes(ctx, (void**)&ctx->out_stage_d, &ctx->out_stage_bytes, sizeof(float) * dtr_elems * chunk * 2))) return rc;
    if (!out_dev && !ctx->down_stream) {
        ECC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            ECC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out_ready[b], cudaEventDisableTiming));
            ECC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out_free[b], cudaEventDisableTiming));
        }
    }
    if (!in_dev && !ctx->copy_stream) {
        ECC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            ECC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
            ECC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_consumed[b], cudaEventDisableTiming));
        }
    }
    if (!in_dev) {
        // the staging buffers may still be read by work queued earlier on the compute stream
        ECC_CUDA(ctx, cudaEventRecord(ctx->ev_consumed[0], ctx->stream));
        ECC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[0], 0));
    }
    // the chunks grow: size the quad kernel's staging once for the largest one instead of at every growth step
    if ((interp == ECC_INTERP_HYBRID || interp == ECC_INTERP_HYBRID_STATIC) && filter == ECC_FILTER_DERIVATIVE && n_images >= 3 &&
        (rc = radon_hybrid4_reserve(ctx, n_u, n_v, chunk)))
        return rc;
    // Host memory for the intermediates: the LAST chunk's download is the exposed one, so large batches end with two short
    // chunks (32, then 16 projections: 38 MB, under a millisecond on the link, against 302 MB behind a chunk of 128).
    const int tail = (!out_dev && n_images >= 256) ? 48 : 0, body = n_images - tail;
    int k = 0, want = first_chunk;
    for (int first = 0; first < n_images; k++) {
        int n;
        if (first >= body) {
            n = (first == body) ? 32 : n_images - first;
        } else {
            n = (body - first < want) ? body - first : want;
            if (body - first - n > 0 && body - first - n < 8 && body - first <= chunk && n >= 8) n = body - first;  // no tiny last launch
        }
        want = (growth * want < chunk) ? growth * want / 4 * 4 : chunk;
        const float* src = images + (size_t)first * img_elems;
        const int b = k & 1;
        if (!in_dev) {
            float* stage = ctx->img_stage_d + (size_t)b * img_elems * chunk;
            if (k >= 2) ECC_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[b], 0));
            ECC_CUDA(ctx, cudaMemcpyAsync(stage, src, sizeof(float) * img_elems * n, cudaMemcpyHostToDevice, ctx->copy_stream));
            ECC_CUDA(ctx, cudaEventRecord(ctx->ev_copied[b], ctx->copy_stream));
            ECC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[b], 0));
            src = stage;
        }
        float* dst = out_dev ? dtrs_out + (size_t)first * dtr_elems : ctx->out_stage_d + (size_t)b * dtr_elems * chunk;
        if (!out_dev && k >= 2) ECC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out_free[b], 0));  // chunk k-2 has left this buffer
        rc = radon_batch(ctx, src, n, n_u, n_v, n_alpha, n_t, filter, post, interp, dst, part.sub(first == 0, first + n == n_images));
        if (rc) return rc;
        if (!in_dev) ECC_CUDA(ctx, cudaEventRecord(ctx->ev_consumed[b], ctx->stream));
        if (!out_dev) {
            ECC_CUDA(ctx, cudaEventRecord(ctx->ev_out_ready[b], ctx->stream));
            ECC_CUDA(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->ev_out_ready[b], 0));
            ECC_CUDA(ctx, cudaMemcpyAsync(dtrs_out + (size_t)first * dtr_elems, dst, sizeof(float) * dtr_elems * n, cudaMemcpyDeviceToHost, ctx->down_stream));
            ECC_CUDA(ctx, cudaEventRecord(ctx->ev_out_free[b], ctx->down_stream));
        }
        first += n;
    }
    if (!out_dev)  // the call's work, downloads included, is ordered on the context's stream as before
        for (int b = 0; b < 2 && b < k; b++) ECC_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out_free[b], 0));
    if (final_sync) ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

extern "C" {

void ecc_camera_intrinsics(const double* P, double* focal_px, double* principal_u, double* principal_v)
{
    double fu = 0, u0 = 0, v0 = 0;
    if (P) camera_intrinsics_host(P, &fu, &u0, &v0);
    if (focal_px) *focal_px = fu;
    if (principal_u) *principal_u = u0;
    if (principal_v) *principal_v = v0;
}

void ecc_preprocess_defaults(ecc_preprocess_params* p)
{
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->scale = 1.0;
    for (int k = 0; k < 4; k++) { p->border_zero[k] = 1; p->border_feather[k] = 16; }
    p->gaussian_sigma = 1.84;
    p->half_kernel_width = 5;
    p->cos_weight = 1;
}

int ecc_preprocess(ecc_context* ctx, float* images, int n, int n_u, int n_v, const ecc_preprocess_params* params, const double* Ps)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!images || !params || n < 0 || n_u < 1 || n_v < 1 || params->n_blanks < 0 || (params->n_blanks > 0 && !params->blanks))
        return fail(ctx, ECC_ERR_INVALID, "ecc_preprocess: bad argument");
    for (int k = 0; k < 4; k++)
        if (params->border_zero[k] < 0 || params->border_feather[k] < 0) return fail(ctx, ECC_ERR_INVALID, "ecc_preprocess: negative border");
    if (n == 0) return ECC_OK;
    if (is_device_pointer(images)) return preprocess_batch(ctx, images, n, n_u, n_v, params, Ps);
    // host images: through the device staging buffer in chunks
    const size_t len = (size_t)n_u * n_v;
    const int chunk = n < 32 ? n : 32;
    int rc = ensure_bytes(ctx, (void**)&ctx->img_stage_d, &ctx->img_stage_bytes, sizeof(float) * len * chunk);
    if (rc) return rc;
    for (int first = 0; first < n; first += chunk) {
        const int m = (n - first < chunk) ? n - first : chunk;
        ECC_CUDA(ctx, cudaMemcpyAsync(ctx->img_stage_d, images + (size_t)first * len, sizeof(float) * len * m, cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = preprocess_batch(ctx, ctx->img_stage_d, m, n_u, n_v, params, Ps ? Ps + (size_t)12 * first : nullptr))) return rc;
        ECC_CUDA(ctx, cudaMemcpyAsync(images + (size_t)first * len, ctx->img_stage_d, sizeof(float) * len * m, cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ECC_OK;
}

int ecc_radon_calibrate_split(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int* window_share_permille)
{
    if (!ctx || !window_share_permille) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (n_u < 2 || n_v < 2 || n_alpha < 1 || n_t < 1) return fail(ctx, ECC_ERR_INVALID, "ecc_radon_calibrate_split: bad argument");
    return radon_hybrid4_calibrate(ctx, n_u, n_v, n_alpha, n_t, 5, window_share_permille);
}

int ecc_radon_set_split(ecc_context* ctx, int window_share_permille)
{
    if (!ctx) return ECC_ERR_INVALID;
    if (window_share_permille < 0 || window_share_permille > 1000) return fail(ctx, ECC_ERR_INVALID, "ecc_radon_set_split: 0 (built-in) .. 1000 per mille");
    ctx->hybrid4.share_override = window_share_permille;
    return ECC_OK;
}

int ecc_radon_num_samples(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int filter, double* count)
{
    if (!ctx || !count) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (n_u < 2 || n_v < 2 || n_alpha < 1 || n_t < 1) return fail(ctx, ECC_ERR_INVALID, "ecc_radon_num_samples: bad argument");
    return radon_num_samples(ctx, n_u, n_v, n_alpha, n_t, filter, count);
}

// Installs n_dtrs device-resident dtrs given by one pointer each (row pitch `pitch` floats): texture objects over
// the caller's memory (zero copy) + the pointer table for the exact-interpolation path.
static int install_dtrs(ecc_context* ctx, const std::vector<const float*>& ptrs, size_t pitch, int n_alpha, int n_t,
                        double step_alpha, double step_t, int n_u, int n_v, int is_derivative)
{
    const int n_dtrs = (int)ptrs.size();
    const bool same = (ptrs == ctx->dtr_ptrs_h && n_alpha == ctx->n_alpha && n_t == ctx->n_t && pitch == ctx->dtr_pitch &&
                       (int)ctx->dtr_tex_h.size() == n_dtrs);
    if (!same) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        destroy_dtr_textures(ctx);
        ctx->dtr_tex_h.resize(n_dtrs, 0);
        static const int dev_arrays = getenv("ECC_DTR_ARRAYS") ? atoi(getenv("ECC_DTR_ARRAYS")) : 0;
        for (int k = 0; k < n_dtrs; k++) {
            cudaResourceDesc res = {};
            res.resType = cudaResourceTypePitch2D;
            res.res.pitch2D.devPtr = const_cast<float*>(ptrs[k]);
            res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
            res.res.pitch2D.width = n_alpha;
            res.res.pitch2D.height = n_t;
            res.res.pitch2D.pitchInBytes = sizeof(float) * pitch;
            if (dev_arrays) {  // development: tiled CUDA arrays as the reference's (measured: same speed, same bits)
                cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
                cudaArray_t arr = nullptr;
                ECC_CUDA(ctx, cudaMallocArray(&arr, &desc, n_alpha, n_t));
                ctx->dtr_arrays.push_back(arr);
                ECC_CUDA(ctx, cudaMemcpy2DToArrayAsync(arr, 0, 0, ptrs[k], sizeof(float) * pitch, sizeof(float) * n_alpha, n_t,
                                                       cudaMemcpyDeviceToDevice, ctx->stream));
                res = cudaResourceDesc();
                res.resType = cudaResourceTypeArray;
                res.res.array.array = arr;
            }
            cudaTextureDesc td = {};
            td.normalizedCoords = 1;  // as the reference's dtr textures (RadonIntermediate.cpp:192)
            td.filterMode = cudaFilterModeLinear;
            td.addressMode[0] = cudaAddressModeClamp;
            td.addressMode[1] = cudaAddressModeClamp;
            td.readMode = cudaReadModeElementType;
            ECC_CUDA(ctx, cudaCreateTextureObject(&ctx->dtr_tex_h[k], &res, &td, nullptr));
        }
        size_t cap_bytes = ctx->dtr_tex_cap * sizeof(cudaTextureObject_t);
        int rc = ensure_bytes(ctx, (void**)&ctx->dtr_tex_d, &cap_bytes, sizeof(cudaTextureObject_t) * n_dtrs);
        ctx->dtr_tex_cap = cap_bytes / sizeof(cudaTextureObject_t);
        if (rc) return rc;
        size_t pcap = ctx->dtr_ptrs_cap * sizeof(float*);
        rc = ensure_bytes(ctx, (void**)&ctx->dtr_ptrs_d, &pcap, sizeof(float*) * n_dtrs);
        ctx->dtr_ptrs_cap = pcap / sizeof(float*);
        if (rc) return rc;
        ECC_CUDA(ctx, cudaMemcpyAsync(ctx->dtr_tex_d, ctx->dtr_tex_h.data(), sizeof(cudaTextureObject_t) * n_dtrs,
                                      cudaMemcpyHostToDevice, ctx->stream));
        ECC_CUDA(ctx, cudaMemcpyAsync(ctx->dtr_ptrs_d, ptrs.data(), sizeof(float*) * n_dtrs, cudaMemcpyHostToDevice, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->dtr_ptrs_h = ptrs;
    }
    ctx->dtr_pitch = pitch;
    ctx->n_dtrs = n_dtrs;
    ctx->n_alpha = n_alpha;
    ctx->n_t = n_t;
    ctx->n_u = n_u;
    ctx->n_v = n_v;
    ctx->step_alpha = (float)step_alpha;
    ctx->step_t = (float)step_t;
    ctx->is_derivative = is_derivative ? 1 : 0;
    return ECC_OK;
}

int ecc_set_radon_intermediates(ecc_context* ctx, const float* dtrs, int n_dtrs, int n_alpha, int n_t,
                                double step_alpha, double step_t, int n_u, int n_v, int is_derivative)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!dtrs || n_dtrs < 1 || n_alpha < 1 || n_t < 1 || n_u < 1 || n_v < 1)
        return fail(ctx, ECC_ERR_INVALID, "ecc_set_radon_intermediates: bad argument");
    const bool dev = is_device_pointer(dtrs);
    const size_t tight_stride = (size_t)n_alpha * n_t;
    // Texture objects over linear memory need a 32-byte aligned pitch and a 512-byte aligned base.
    const bool borrowable = dev && (n_alpha % 8 == 0) && ((tight_stride * sizeof(float)) % 512 == 0) &&
                            ((uintptr_t)dtrs % 512 == 0);
    const float* base;
    size_t pitch, stride;
    if (borrowable) {
        base = dtrs;
        pitch = n_alpha;
        stride = tight_stride;
    } else {
        pitch = round_up(n_alpha, 8);
        stride = round_up(pitch * n_t, 128);
        int rc = ensure_bytes(ctx, (void**)&ctx->dtrs_owned, &ctx->dtrs_owned_bytes, sizeof(float) * stride * n_dtrs);
        if (rc) return rc;
        if (pitch == (size_t)n_alpha && stride == tight_stride) {
            ECC_CUDA(ctx, cudaMemcpyAsync(ctx->dtrs_owned, dtrs, sizeof(float) * stride * n_dtrs, cudaMemcpyDefault, ctx->stream));
        } else {
            ECC_CUDA(ctx, cudaMemsetAsync(ctx->dtrs_owned, 0, sizeof(float) * stride * n_dtrs, ctx->stream));
            for (int k = 0; k < n_dtrs; k++)
                ECC_CUDA(ctx, cudaMemcpy2DAsync(ctx->dtrs_owned + stride * k, sizeof(float) * pitch, dtrs + tight_stride * k,
                                                sizeof(float) * n_alpha, sizeof(float) * n_alpha, n_t, cudaMemcpyDefault, ctx->stream));
        }
        base = ctx->dtrs_owned;
    }
    std::vector<const float*> ptrs(n_dtrs);
    for (int k = 0; k < n_dtrs; k++) ptrs[k] = base + stride * k;
    ctx->dtrs_d = base;
    ctx->dtr_stride = stride;
    return install_dtrs(ctx, ptrs, pitch, n_alpha, n_t, step_alpha, step_t, n_u, n_v, is_derivative);
}

int ecc_set_radon_intermediate_pointers(ecc_context* ctx, const float* const* dtrs, int n_dtrs, int n_alpha, int n_t,
                                        double step_alpha, double step_t, int n_u, int n_v, int is_derivative)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!dtrs || n_dtrs < 1 || n_alpha < 1 || n_t < 1 || n_u < 1 || n_v < 1)
        return fail(ctx, ECC_ERR_INVALID, "ecc_set_radon_intermediate_pointers: bad argument");
    if (n_alpha % 8 != 0) return fail(ctx, ECC_ERR_INVALID, "zero-copy dtrs need n_alpha to be a multiple of 8 (32-byte texture pitch)");
    std::vector<const float*> ptrs(dtrs, dtrs + n_dtrs);
    for (int k = 0; k < n_dtrs; k++)
        if (!is_device_pointer(ptrs[k]) || (uintptr_t)ptrs[k] % 512 != 0)
            return fail(ctx, ECC_ERR_INVALID, "every dtr must be device memory aligned to 512 bytes (use ecc_device_alloc)");
    ctx->dtrs_d = nullptr;
    ctx->dtr_stride = 0;
    return install_dtrs(ctx, ptrs, (size_t)n_alpha, n_alpha, n_t, step_alpha, step_t, n_u, n_v, is_derivative);
}

int ecc_device_alloc(ecc_context* ctx, size_t bytes, void** ptr)
{
    if (!ctx || !ptr) return ECC_ERR_INVALID;
    Guard g(ctx);
    ECC_CUDA(ctx, cudaMalloc(ptr, bytes ? bytes : 16));  // cudaMalloc returns at least 512-byte aligned memory
    return ECC_OK;
}

int ecc_device_free(ecc_context* ctx, void* ptr)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!ptr) return ECC_OK;
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ECC_CUDA(ctx, cudaFree(ptr));
    return ECC_OK;
}

int ecc_copy(ecc_context* ctx, void* dst, const void* src, size_t bytes)
{
    if (!ctx || (bytes && (!dst || !src))) return ECC_ERR_INVALID;
    Guard g(ctx);
    ECC_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

int ecc_set_projection_matrices(ecc_context* ctx, const double* Ps, int n)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (n < 0 || (n > 0 && !Ps)) return fail(ctx, ECC_ERR_INVALID, "ecc_set_projection_matrices: bad argument");
    // optimiser loops hand over the whole set every step (Gui/SingleImageMotion.h:84-90): an unchanged set keeps its
    // derived views on the device
    if (n > 0 && n == ctx->n_views && ctx->Ps_d && !is_device_pointer(Ps) && ctx->Ps_h.size() == (size_t)12 * n &&
        std::memcmp(ctx->Ps_h.data(), Ps, sizeof(double) * 12 * n) == 0)
        return ECC_OK;
    ctx->geometry_version++;
    ctx->n_views = n;
    ctx->Ps_h.resize((size_t)12 * n);
    if (n == 0) return ECC_OK;
    if (is_device_pointer(Ps)) {
        ECC_CUDA(ctx, cudaMemcpyAsync(ctx->Ps_h.data(), Ps, sizeof(double) * 12 * n, cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        std::memcpy(ctx->Ps_h.data(), Ps, sizeof(double) * 12 * n);
    }
    return upload_and_derive(ctx, ctx->Ps_h.data(), n, &ctx->Ps_d, &ctx->Cs_d, &ctx->PinvTs_d, &ctx->Ps_cap);
}

int ecc_update_projection_matrix(ecc_context* ctx, int index, const double* P)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!P || index < 0 || index >= ctx->n_views) return fail(ctx, ECC_ERR_INVALID, "ecc_update_projection_matrix: bad index");
    // the stream may still be reading the previous host copy of this matrix (async H2D from pageable
    // memory completes before the call returns, so overwriting is safe)
    std::memcpy(&ctx->Ps_h[(size_t)12 * index], P, sizeof(double) * 12);
    ctx->geometry_version++;
    // One view: derived on the host (the same fp64 operation sequence as the device kernel, bit-identical:
    // tests/test_gpu_parity.py::test_derived_views_device_equals_host_equals_reference) and uploaded as 64 bytes -- a
    // one-thread fp64 kernel would sit on the critical path of every tracking step for ~8 us.  (Ps_d is only the staging
    // of the whole-set derivation and is not kept current here.)
    float A[12], C[4];
    derive_view(P, A, C);
    ECC_CUDA(ctx, cudaMemcpyAsync(ctx->PinvTs_d + (size_t)12 * index, A, sizeof(A), cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(ctx->Cs_d + (size_t)4 * index, C, sizeof(C), cudaMemcpyHostToDevice, ctx->stream));
    return ECC_OK;
}

int ecc_get_derived_views(ecc_context* ctx, float* PinvTs, float* Cs)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (ctx->n_views == 0) return ECC_OK;
    if (PinvTs) ECC_CUDA(ctx, cudaMemcpyAsync(PinvTs, ctx->PinvTs_d, sizeof(float) * 12 * ctx->n_views, cudaMemcpyDeviceToHost, ctx->stream));
    if (Cs) ECC_CUDA(ctx, cudaMemcpyAsync(Cs, ctx->Cs_d, sizeof(float) * 4 * ctx->n_views, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

void ecc_derive_views_host(const double* Ps, int n, float* PinvTs, float* Cs)
{
    for (int v = 0; v < n; v++) {
        float A[12], C[4];
        derive_view(Ps + (size_t)12 * v, A, C);
        if (PinvTs) std::memcpy(PinvTs + (size_t)12 * v, A, sizeof(A));
        if (Cs) std::memcpy(Cs + (size_t)4 * v, C, sizeof(C));
    }
}

int ecc_set_object_radius(ecc_context* ctx, double r)
{
    if (!ctx) return ECC_ERR_INVALID;
    ctx->object_radius = r;
    return ECC_OK;
}

int ecc_get_object_radius(ecc_context* ctx, double* radius)
{
    if (!ctx || !radius) return ECC_ERR_INVALID;
    if (ctx->object_radius > 0) {
        *radius = ctx->object_radius;
        return ECC_OK;
    }
    if (ctx->n_views == 0) {
        *radius = 0;
        return ECC_OK;
    }
    // estimateObjectRadius on the first matrix (EpipolarConsistency.cpp:35-47)
    *radius = object_radius_from_view(ctx->Ps_h.data(), ctx->n_u, ctx->n_v);
    return ECC_OK;
}

int ecc_set_epipolar_plane_step(ecc_context* ctx, double dkappa)
{
    if (!ctx) return ECC_ERR_INVALID;
    ctx->dkappa = dkappa;
    return ECC_OK;
}

int ecc_use_correlation(ecc_context* ctx, int on)
{
    if (!ctx) return ECC_ERR_INVALID;
    ctx->use_corr = on ? 1 : 0;
    return ECC_OK;
}

int ecc_set_interpolation(ecc_context* ctx, int interp)
{
    if (!ctx) return ECC_ERR_INVALID;
    if (interp != ECC_INTERP_TEXTURE && interp != ECC_INTERP_EXACT) return fail(ctx, ECC_ERR_INVALID, "bad interp");
    ctx->interp = interp;
    return ECC_OK;
}

int ecc_evaluate_range(ecc_context* ctx, long long pair_begin, long long pair_end, float* cost_image, double* sum)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    PairLaunch L;
    int rc = fill_launch(ctx, L);
    if (rc) return rc;
    const long long n = ctx->n_views, total = n * (n - 1) / 2;
    if (ctx->n_dtrs < ctx->n_views) return fail(ctx, ECC_ERR_STATE, "all-pairs evaluation needs one dtr per projection matrix");
    if (pair_begin < 0 || pair_end > total || pair_begin > pair_end) return fail(ctx, ECC_ERR_INVALID, "pair range out of bounds");
    const long long count = pair_end - pair_begin;
    if (sum) *sum = 0.0;
    if (count == 0) return ECC_OK;
    size_t cap = ctx->vals_cap * sizeof(float);
    if ((rc = ensure_bytes(ctx, (void**)&ctx->vals_d, &cap, sizeof(float) * count))) return rc;
    ctx->vals_cap = cap / sizeof(float);
    L.pair_begin = pair_begin;
    L.n_pairs = count;
    L.mode_items = total;  // a range of the all-pairs job: every pair is computed as the whole job would compute it
    L.vals_d = ctx->vals_d;
    const bool img_dev = cost_image && is_device_pointer(cost_image);
    if (cost_image) {
        if (img_dev) {
            L.image_d = cost_image;
        } else {
            L.image_d = nullptr;  // scatter on the host from the compact values instead
        }
    }
    if ((rc = launch_pairs(ctx, L))) return rc;
    if (sum) {
        size_t scap = ctx->sums_cap * sizeof(double);
        if ((rc = ensure_bytes(ctx, (void**)&ctx->sums_d, &scap, sizeof(double)))) return rc;
        ctx->sums_cap = scap / sizeof(double);
        if ((rc = launch_sum_sets(ctx, ctx->vals_d, count, 1, ctx->sums_d))) return rc;
    }
    const bool host_image = cost_image && !img_dev;
    if (sum || host_image) {
        const size_t need = sizeof(double) + (host_image ? sizeof(float) * count : 0);
        if ((rc = ensure_pinned(ctx, need))) return rc;
        double* sum_h = (double*)ctx->pinned_h;
        float* vals_h = (float*)((char*)ctx->pinned_h + sizeof(double));
        if (sum) ECC_CUDA(ctx, cudaMemcpyAsync(sum_h, ctx->sums_d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (host_image) ECC_CUDA(ctx, cudaMemcpyAsync(vals_h, ctx->vals_d, sizeof(float) * count, cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (sum) *sum = *sum_h;
        if (host_image) {
            // entries i + j*n for the evaluated pairs only; everything else keeps the caller's values
            int i, j;
            pair_from_index(pair_begin, (int)n, i, j);
            for (long long k = 0; k < count; k++) {
                cost_image[(size_t)i + (size_t)j * n] = vals_h[k];
                if (++j >= n) { ++i; j = i + 1; }
            }
        }
    }
    return ECC_OK;
}

int ecc_evaluate(ecc_context* ctx, float* cost_image, double* mean)
{
    if (!ctx) return ECC_ERR_INVALID;
    const long long n = ctx->n_views, total = n * (n - 1) / 2;
    double sum = 0.0;
    int rc = ecc_evaluate_range(ctx, 0, total, cost_image, mean ? &sum : nullptr);
    if (rc) return rc;
    if (mean) *mean = total ? sum / (double)total : 0.0;
    return ECC_OK;
}

static int stage_indices(ecc_context* ctx, const int* idx4, int n_pairs, const int** idx_d)
{
    if (is_device_pointer(idx4) && ((uintptr_t)idx4 % 16 == 0)) {
        *idx_d = idx4;
        return ECC_OK;
    }
    size_t cap = ctx->idx_cap * sizeof(int);
    int rc = ensure_bytes(ctx, (void**)&ctx->idx_d, &cap, sizeof(int) * 4 * n_pairs);
    ctx->idx_cap = cap / sizeof(int);
    if (rc) return rc;
    ECC_CUDA(ctx, cudaMemcpyAsync(ctx->idx_d, idx4, sizeof(int) * 4 * n_pairs, cudaMemcpyDefault, ctx->stream));
    *idx_d = ctx->idx_d;
    return ECC_OK;
}

static int check_indices(ecc_context* ctx, const int* idx4, int n_pairs)
{
    // the reference checks only in debug builds (EpipolarConsistencyRadonIntermediate.cpp:248-275);
    // out-of-range indices would fault the GPU, so host-side lists are always checked here.
    if (is_device_pointer(idx4)) return ECC_OK;
    for (int k = 0; k < n_pairs; k++) {
        const int* q = idx4 + 4 * k;
        if (q[0] < 0 || q[0] >= ctx->n_views || q[1] < 0 || q[1] >= ctx->n_views || q[2] < 0 || q[2] >= ctx->n_dtrs ||
            q[3] < 0 || q[3] >= ctx->n_dtrs)
            return fail(ctx, ECC_ERR_INVALID, "index array contains invalid indices");
    }
    return ECC_OK;
}

int ecc_evaluate_indices(ecc_context* ctx, const int* idx4, int n_pairs, float* out, double* mean)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (n_pairs < 0 || (n_pairs > 0 && !idx4)) return fail(ctx, ECC_ERR_INVALID, "ecc_evaluate_indices: bad argument");
    PairLaunch L;
    int rc = fill_launch(ctx, L);
    if (rc) return rc;
    if (mean) *mean = 0.0;
    if (n_pairs == 0) return ECC_OK;
    if ((rc = check_indices(ctx, idx4, n_pairs))) return rc;
    if ((rc = stage_indices(ctx, idx4, n_pairs, &L.idx4_d))) return rc;
    const bool out_dev = out && is_device_pointer(out);
    if (out_dev) {
        L.vals_d = out;
    } else {
        size_t cap = ctx->vals_cap * sizeof(float);
        if ((rc = ensure_bytes(ctx, (void**)&ctx->vals_d, &cap, sizeof(float) * n_pairs))) return rc;
        ctx->vals_cap = cap / sizeof(float);
        L.vals_d = ctx->vals_d;
    }
    L.n_pairs = n_pairs;
    if ((rc = launch_pairs(ctx, L))) return rc;
    if (mean) {
        size_t scap = ctx->sums_cap * sizeof(double);
        if ((rc = ensure_bytes(ctx, (void**)&ctx->sums_d, &scap, sizeof(double)))) return rc;
        ctx->sums_cap = scap / sizeof(double);
        if ((rc = launch_sum_sets(ctx, L.vals_d, n_pairs, 1, ctx->sums_d))) return rc;
    }
    const bool host_out = out && !out_dev;
    if (mean || host_out) {
        if ((rc = ensure_pinned(ctx, sizeof(double) + sizeof(float) * (size_t)n_pairs))) return rc;
        double* sum_h = (double*)ctx->pinned_h;
        float* vals_h = (float*)((char*)ctx->pinned_h + sizeof(double));
        if (mean) ECC_CUDA(ctx, cudaMemcpyAsync(sum_h, ctx->sums_d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (host_out) ECC_CUDA(ctx, cudaMemcpyAsync(vals_h, L.vals_d, sizeof(float) * n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (mean) *mean = *sum_h / (double)n_pairs;
        if (host_out) std::memcpy(out, vals_h, sizeof(float) * n_pairs);
    }
    return ECC_OK;
}

// ---- tracking step as one CUDA graph -------------------------------------------------------------------------------
static void track_free(ecc_context* ctx)
{
    TrackGraph& T = ctx->track;
    if (T.exec) cudaGraphExecDestroy(T.exec);
    if (T.idx_d) cudaFree(T.idx_d);
    if (T.pinned) cudaFreeHost(T.pinned);
    if (T.live_d) cudaFree(T.live_d);
    if (T.done_d) cudaFree(T.done_d);
    T = TrackGraph();
}

// The pinned block of a tracking step: what the host hands over and what the launch hands back.
struct TrackPinned {
    float view[16];     // [12 floats (P+)^T | 4 floats C] of the live view
    unsigned seq;       // sequence number of the call, uploaded with the view
    unsigned pad[3];
    double sum;         // sum of the listed pairs' values
    volatile unsigned flag;  // the sequence number once values and sum are in place (fused recording)
    unsigned pad2;
    float vals[1];      // [n_pairs]
};

// Records a tracking step on the context's own stream.  Fused recording (CTA-per-pair launches, i.e. lists of fewer than 64
// pairs per SM): TWO nodes -- one 80-byte copy {derived view, sequence number} and the pair kernel, whose last CTA adds up,
// writes values and sum into the pinned block, stores the view into the arrays and raises the flag the host waits on.
// Otherwise FOUR nodes: two copies of the derived view into the arrays, pair kernel, finalize + sum into the pinned block.
static int track_capture(ecc_context* ctx, int index, int n_pairs, bool want_out, const PairLaunch& L_in, const int* idx_user_d)
{
    TrackGraph& T = ctx->track;
    if (T.exec) { cudaGraphExecDestroy(T.exec); T.exec = nullptr; }
    TrackPinned* pin = (TrackPinned*)T.pinned;
    static const bool no_fuse = getenv("ECC_TRACK_NO_FUSE") != nullptr;  // development: the four-node recording
    T.fused = false;
    for (int attempt = no_fuse ? 1 : 0; attempt < 2; attempt++) {
        const bool fuse = attempt == 0;
        cudaStream_t saved = ctx->stream, cap = ctx->own_stream;
        if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return ECC_ERR_CUDA; }
        ctx->stream = cap;
        int rc = ECC_OK;
        bool fused_ok = false;
        do {
            PairLaunch L = L_in;
            L.idx4_d = idx_user_d ? idx_user_d : T.idx_d;  // a device-resident list is read in place at every replay
            L.vals_d = ctx->vals_d;
            L.n_pairs = n_pairs;
            L.defer_finalize = 1;
            if (fuse) {
                if (cudaMemcpyAsync(T.live_d, pin, 80, cudaMemcpyHostToDevice, cap) != cudaSuccess) { rc = ECC_ERR_CUDA; break; }
                L.live_d = T.live_d;
                L.live_index = index;
                L.done_d = T.done_d;
                L.fused_sum_out = &pin->sum;
                L.fused_vals_out = want_out ? pin->vals : nullptr;
                L.fused_flag = const_cast<unsigned*>(&pin->flag);
            } else {
                if (cudaMemcpyAsync(ctx->PinvTs_d + (size_t)12 * index, pin->view, sizeof(float) * 12, cudaMemcpyHostToDevice, cap) != cudaSuccess) { rc = ECC_ERR_CUDA; break; }
                if (cudaMemcpyAsync(ctx->Cs_d + (size_t)4 * index, pin->view + 12, sizeof(float) * 4, cudaMemcpyHostToDevice, cap) != cudaSuccess) { rc = ECC_ERR_CUDA; break; }
            }
            PairLaunch R;
            if ((rc = launch_pairs(ctx, L, &R))) break;
            fused_ok = fuse && R.done_d != nullptr;
            if (fuse && !fused_ok) break;  // not a CTA-per-pair launch: record again, the plain way
            // plain: finalize + sum in one launch that writes the results straight into the pinned block (no copy nodes)
            if (!fuse && (rc = launch_finalize_sum(ctx, R, &pin->sum, want_out ? pin->vals : nullptr))) break;
        } while (0);
        ctx->stream = saved;
        cudaGraph_t graph = nullptr;
        const cudaError_t e = cudaStreamEndCapture(cap, &graph);
        if (rc == ECC_OK && (e != cudaSuccess || !graph)) rc = ECC_ERR_CUDA;
        if (rc == ECC_OK && fuse && !fused_ok) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            continue;
        }
        if (rc == ECC_OK && cudaGraphInstantiate(&T.exec, graph, 0) != cudaSuccess) { T.exec = nullptr; rc = ECC_ERR_CUDA; }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc == ECC_OK) T.fused = fused_ok;
        return rc;
    }
    return ECC_ERR_CUDA;
}

int ecc_update_and_evaluate(ecc_context* ctx, int index, const double* P, const int* idx4, int n_pairs, float* out, double* mean)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!P || index < 0 || index >= ctx->n_views) return fail(ctx, ECC_ERR_INVALID, "ecc_update_and_evaluate: bad index");
    if (n_pairs < 0 || (n_pairs > 0 && !idx4)) return fail(ctx, ECC_ERR_INVALID, "ecc_update_and_evaluate: bad argument");
    TrackGraph& T = ctx->track;
    const bool out_dev = out && is_device_pointer(out);
    const bool idx_dev = idx4 && is_device_pointer(idx4);
    static const bool graphs_off = getenv("ECC_NO_GRAPHS") != nullptr;
    if (graphs_off || T.failed || ctx->profiling || out_dev || n_pairs == 0 || (idx_dev && (uintptr_t)idx4 % 16 != 0)) {
        int rc = ecc_update_projection_matrix(ctx, index, P);
        if (rc) return rc;
        return ecc_evaluate_indices(ctx, idx4, n_pairs, out, mean);
    }
    int rc;
    if (!idx_dev && (rc = check_indices(ctx, idx4, n_pairs))) return rc;
    // host copy of the matrix first: the automatic object radius reads it (first view only)
    std::memcpy(&ctx->Ps_h[(size_t)12 * index], P, sizeof(double) * 12);
    ctx->geometry_version++;
    PairLaunch L;
    if ((rc = fill_launch(ctx, L))) return rc;
    const bool same_list = idx_dev ? false : ((int)T.idx_h.size() == 4 * n_pairs && std::memcmp(T.idx_h.data(), idx4, sizeof(int) * 4 * n_pairs) == 0);
    const std::vector<double> key = {(double)index, (double)n_pairs, idx_dev ? (double)(uintptr_t)idx4 : -1.0, (double)(out != nullptr),
                                     (double)L.radius, (double)L.dkappa, (double)L.interp, (double)L.use_corr, (double)L.is_derivative,
                                     (double)L.n_views, (double)L.n_dtrs, (double)L.n_alpha, (double)L.n_t, (double)L.range_t,
                                     (double)L.sample_cap, (double)L.dtr_pitch, (double)L.half_nu, (double)L.half_nv,
                                     (double)(uintptr_t)L.tex_d, (double)(uintptr_t)L.dtr_ptrs_d, (double)(uintptr_t)ctx->Ps_d,
                                     (double)(uintptr_t)ctx->Cs_d, (double)(uintptr_t)ctx->PinvTs_d, (double)(uintptr_t)ctx->vals_d,
                                     (double)(uintptr_t)ctx->sums_d, (double)(uintptr_t)ctx->partials_d, (double)(uintptr_t)ctx->dtr_tex_h.size()};
    const bool replay = T.exec && key == T.key && (idx_dev || same_list);
    if (!replay) {
        const bool seen = key == T.pending_key && (idx_dev || same_list);
        if (!seen) {
            // first call with these parameters: the plain path (it also sizes every buffer the graph will use)
            T.pending_key.clear();
            rc = ecc_update_projection_matrix(ctx, index, P);
            if (!rc) rc = ecc_evaluate_indices(ctx, idx4, n_pairs, out, mean);
            if (rc) return rc;
            if (!idx_dev) T.idx_h.assign(idx4, idx4 + (size_t)4 * n_pairs);
            // the key as the NEXT call will see it (buffers may have been allocated just now)
            std::vector<double> k2 = key;
            PairLaunch L2;
            if (fill_launch(ctx, L2) == ECC_OK) {
                k2[18] = (double)(uintptr_t)L2.tex_d; k2[19] = (double)(uintptr_t)L2.dtr_ptrs_d;
            }
            k2[20] = (double)(uintptr_t)ctx->Ps_d; k2[21] = (double)(uintptr_t)ctx->Cs_d; k2[22] = (double)(uintptr_t)ctx->PinvTs_d;
            k2[23] = (double)(uintptr_t)ctx->vals_d; k2[24] = (double)(uintptr_t)ctx->sums_d; k2[25] = (double)(uintptr_t)ctx->partials_d;
            T.pending_key = k2;
            return ECC_OK;
        }
        // second call: record.  Own copies of the pair list and the pinned block first (outside the capture).
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!idx_dev) {
            if (T.idx_cap < sizeof(int) * 4 * (size_t)n_pairs) {
                if (T.idx_d) cudaFree(T.idx_d);
                T.idx_d = nullptr; T.idx_cap = 0;
                ECC_CUDA(ctx, cudaMalloc(&T.idx_d, sizeof(int) * 4 * (size_t)n_pairs));
                T.idx_cap = sizeof(int) * 4 * (size_t)n_pairs;
            }
            ECC_CUDA(ctx, cudaMemcpy(T.idx_d, idx4, sizeof(int) * 4 * (size_t)n_pairs, cudaMemcpyHostToDevice));
        }
        const size_t need = sizeof(TrackPinned) + sizeof(float) * (size_t)n_pairs;
        if (T.pinned_bytes < need) {
            if (T.pinned) cudaFreeHost(T.pinned);
            T.pinned = nullptr; T.pinned_bytes = 0;
            ECC_CUDA(ctx, cudaMallocHost(&T.pinned, need));
            T.pinned_bytes = need;
            std::memset(T.pinned, 0, need);
        }
        if (!T.live_d) {
            ECC_CUDA(ctx, cudaMalloc(&T.live_d, 80));
            ECC_CUDA(ctx, cudaMalloc(&T.done_d, sizeof(unsigned)));
            ECC_CUDA(ctx, cudaMemset(T.done_d, 0, sizeof(unsigned)));
        }
        if (track_capture(ctx, index, n_pairs, out != nullptr, L, idx_dev ? idx4 : nullptr) != ECC_OK) {
            T.failed = true;  // e.g. a driver without stream capture for one of the nodes: plain path from now on
            rc = ecc_update_projection_matrix(ctx, index, P);
            if (rc) return rc;
            return ecc_evaluate_indices(ctx, idx4, n_pairs, out, mean);
        }
        T.key = key;
        T.pending_key.clear();
    }
    // replay: the matrix goes through the pinned block, everything else is in the recorded nodes
    TrackPinned* pin = (TrackPinned*)T.pinned;
    derive_view(P, pin->view, pin->view + 12);  // on the host, bit-identical to the device derivation
    pin->seq = ++T.seq ? T.seq : ++T.seq;       // never zero
    ECC_CUDA(ctx, cudaGraphLaunch(T.exec, ctx->stream));
    if (T.fused) {
        // the kernel's last CTA raises the flag once values and sum are in the pinned block: the host watches the flag instead
        // of waiting for the stream to drain (the stream is looked at now and then, so that a failed launch cannot hang the call)
        static const bool no_poll = getenv("ECC_TRACK_NO_POLL") != nullptr;  // development
        unsigned spins = 0;
        while (!no_poll && pin->flag != pin->seq) {
            if ((++spins & 0x3fffu) == 0) {
                const cudaError_t q = cudaStreamQuery(ctx->stream);
                if (q == cudaSuccess) break;  // finished: the flag is there, or the launch did not do its work (checked below)
                if (q != cudaErrorNotReady) return fail(ctx, ECC_ERR_CUDA, std::string("ecc_update_and_evaluate: ") + cudaGetErrorString(q));
            }
        }
        if (no_poll || pin->flag != pin->seq) ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pin->flag != pin->seq) return fail(ctx, ECC_ERR_CUDA, "ecc_update_and_evaluate: the fused launch did not complete");
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    T.replays++;
    if (mean) *mean = pin->sum / (double)n_pairs;
    if (out) std::memcpy(out, pin->vals, sizeof(float) * (size_t)n_pairs);
    return ECC_OK;
}

int ecc_track_info(ecc_context* ctx, int* kernels_per_call, long long* replays)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (kernels_per_call) *kernels_per_call = !ctx->track.exec ? 0 : (ctx->track.fused ? 1 : 2);
    if (replays) *replays = ctx->track.replays;
    return ECC_OK;
}

}  // extern "C"

// Batched evaluation, common parts: the pair list of the launch ...
int eccb200::batch_begin(ecc_context* ctx, const int* idx4, int n_pairs, PairLaunch& L, const char* who)
{
    int rc = fill_launch(ctx, L);
    if (rc) return rc;
    const long long n = ctx->n_views;
    if (idx4) {
        if (n_pairs < 1) return fail(ctx, ECC_ERR_INVALID, std::string(who) + ": empty pair list");
        if ((rc = check_indices(ctx, idx4, n_pairs))) return rc;
        if ((rc = stage_indices(ctx, idx4, n_pairs, &L.idx4_d))) return rc;
        L.n_pairs = n_pairs;
    } else {
        if (ctx->n_dtrs < ctx->n_views) return fail(ctx, ECC_ERR_STATE, "all-pairs evaluation needs one dtr per projection matrix");
        L.n_pairs = n * (n - 1) / 2;
    }
    return ECC_OK;
}

// ... room for the derived views of n_sets matrix sets (every set gets the object radius the reference would derive from
// that set's first matrix) ...
int eccb200::batch_reserve(ecc_context* ctx, int n_sets, bool want_matrices)
{
    BatchBuffers& B = ctx->batch;
    const size_t count = (size_t)n_sets * ctx->n_views;
    if (B.radii_cap < (size_t)n_sets) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (B.radii_d) cudaFree(B.radii_d);
        B.radii_d = nullptr;
        B.radii_cap = 0;
        ECC_CUDA(ctx, cudaMalloc(&B.radii_d, sizeof(float) * n_sets));
        B.radii_cap = n_sets;
    }
    if (B.cap < count || (want_matrices && !B.Ps_d)) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (B.Ps_d) cudaFree(B.Ps_d);
        if (B.Cs_d) cudaFree(B.Cs_d);
        if (B.A_d) cudaFree(B.A_d);
        B.Ps_d = nullptr; B.Cs_d = nullptr; B.A_d = nullptr; B.cap = 0;
        ECC_CUDA(ctx, cudaMalloc(&B.Ps_d, sizeof(double) * 12 * count));
        ECC_CUDA(ctx, cudaMalloc(&B.Cs_d, sizeof(float) * 4 * count));
        ECC_CUDA(ctx, cudaMalloc(&B.A_d, sizeof(float) * 12 * count));
        B.cap = count;
    }
    return ECC_OK;
}

// ... and launch, sums, downloads once the batch buffers hold the derived views of all sets.
int eccb200::batch_finish(ecc_context* ctx, PairLaunch& L, int n_sets, float* out, double* means)
{
    BatchBuffers& B = ctx->batch;
    int rc;
    L.radii_d = B.radii_d;
    L.Cs_d = B.Cs_d;
    L.PinvTs_d = B.A_d;
    L.n_sets = n_sets;
    const size_t items = (size_t)n_sets * L.n_pairs;
    const bool out_dev = out && is_device_pointer(out);
    if (out_dev) {
        L.vals_d = out;
    } else {
        size_t cap = ctx->vals_cap * sizeof(float);
        if ((rc = ensure_bytes(ctx, (void**)&ctx->vals_d, &cap, sizeof(float) * items))) return rc;
        ctx->vals_cap = cap / sizeof(float);
        L.vals_d = ctx->vals_d;
    }
    if ((rc = launch_pairs(ctx, L))) return rc;
    if (means) {
        size_t scap = ctx->sums_cap * sizeof(double);
        if ((rc = ensure_bytes(ctx, (void**)&ctx->sums_d, &scap, sizeof(double) * n_sets))) return rc;
        ctx->sums_cap = scap / sizeof(double);
        if ((rc = launch_sum_sets(ctx, L.vals_d, L.n_pairs, n_sets, ctx->sums_d))) return rc;
    }
    const bool host_out = out && !out_dev;
    if (means || host_out) {
        if ((rc = ensure_pinned(ctx, sizeof(double) * n_sets + (host_out ? sizeof(float) * items : 0)))) return rc;
        double* sums_h = (double*)ctx->pinned_h;
        float* vals_h = (float*)((char*)ctx->pinned_h + sizeof(double) * n_sets);
        if (means) ECC_CUDA(ctx, cudaMemcpyAsync(sums_h, ctx->sums_d, sizeof(double) * n_sets, cudaMemcpyDeviceToHost, ctx->stream));
        if (host_out) ECC_CUDA(ctx, cudaMemcpyAsync(vals_h, L.vals_d, sizeof(float) * items, cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (means)
            for (int s = 0; s < n_sets; s++) means[s] = sums_h[s] / (double)L.n_pairs;
        if (host_out) std::memcpy(out, vals_h, sizeof(float) * items);
    }
    return ECC_OK;
}

extern "C" {

int ecc_evaluate_batch(ecc_context* ctx, const double* Ps_sets, int n_sets, const int* idx4, int n_pairs, float* out,
                       double* means)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!Ps_sets || n_sets < 1) return fail(ctx, ECC_ERR_INVALID, "ecc_evaluate_batch: bad argument");
    PairLaunch L;
    int rc = batch_begin(ctx, idx4, n_pairs, L, "ecc_evaluate_batch");
    if (rc) return rc;
    if ((rc = batch_reserve(ctx, n_sets, true))) return rc;
    BatchBuffers& B = ctx->batch;
    const size_t count = (size_t)n_sets * ctx->n_views;
    ECC_CUDA(ctx, cudaMemcpyAsync(B.Ps_d, Ps_sets, sizeof(double) * 12 * count, cudaMemcpyDefault, ctx->stream));
    if ((rc = launch_derive_views(ctx, B.Ps_d, (int)count, B.A_d, B.Cs_d, ctx->n_views, ctx->n_u, ctx->n_v, ctx->object_radius, B.radii_d)))
        return rc;
    return batch_finish(ctx, L, n_sets, out, means);
}

int ecc_pair_signals(ecc_context* ctx, int p0, int p1, int dtr0, int dtr1, int capacity, float* kappas, float* signal0,
                     float* signal1, float* lines0, float* lines1, int* n_samples, double* weight, double* value)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    PairLaunch L;
    int rc = fill_launch(ctx, L);
    if (rc) return rc;
    const int idx_h[4] = {p0, p1, dtr0, dtr1};
    if ((rc = check_indices(ctx, idx_h, 1))) return rc;
    if (capacity < 0) return fail(ctx, ECC_ERR_INVALID, "ecc_pair_signals: negative capacity");
    // scratch: [4 ints idx | 2 ints head | sample_cap * 13 floats], reusing the partial-sum buffer
    const size_t words = 8 + (size_t)L.sample_cap * 13;
    size_t cap = ctx->partials_cap * sizeof(float);
    if ((rc = ensure_bytes(ctx, (void**)&ctx->partials_d, &cap, sizeof(float) * words))) return rc;
    ctx->partials_cap = cap / sizeof(float);
    int* idx_d = (int*)ctx->partials_d;
    int* head_d = idx_d + 4;
    float* rec_d = ctx->partials_d + 8;
    if ((rc = ensure_pinned(ctx, sizeof(float) * words))) return rc;
    std::memcpy(ctx->pinned_h, idx_h, sizeof(idx_h));
    std::memset((char*)ctx->pinned_h + 16, 0, 16);
    ECC_CUDA(ctx, cudaMemcpyAsync(idx_d, ctx->pinned_h, 32, cudaMemcpyHostToDevice, ctx->stream));
    L.idx4_d = idx_d;
    L.n_pairs = 1;
    if ((rc = launch_pair_signals(ctx, L, rec_d, head_d))) return rc;
    ECC_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_h, ctx->partials_d, sizeof(float) * words, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int* head = (const int*)ctx->pinned_h + 4;
    const float* rec = (const float*)ctx->pinned_h + 8;
    const int count = head[0];
    float w;
    std::memcpy(&w, &head[1], sizeof(w));
    if (weight) *weight = (double)w;
    if (n_samples) *n_samples = 2 * count;
    // ascending kappa: -kappa_max ... -dkappa/2, +dkappa/2 ... +kappa_max
    for (int q = 0; q < 2 * count && q < capacity; q++) {
        const bool minus = q < count;
        const int m = minus ? count - 1 - q : q - count;
        const float* r = rec + (size_t)m * 13;
        if (kappas) kappas[q] = minus ? -r[0] : r[0];
        if (signal0) signal0[q] = minus ? r[3] : r[1];
        if (signal1) signal1[q] = minus ? r[4] : r[2];
        if (lines0) { lines0[2 * q] = minus ? r[9] : r[5]; lines0[2 * q + 1] = minus ? r[10] : r[6]; }
        if (lines1) { lines1[2 * q] = minus ? r[11] : r[7]; lines1[2 * q + 1] = minus ? r[12] : r[8]; }
    }
    if (value) {
        float v = 0.f;
        if ((rc = ecc_evaluate_indices(ctx, idx_h, 1, &v, nullptr))) return rc;
        *value = (double)v;
    }
    return ECC_OK;
}

int ecc_pair_maps(ecc_context* ctx, const int* idx4, int n_pairs, float* K01s)
{
    if (!ctx || !K01s) return ECC_ERR_INVALID;
    Guard g(ctx);
    PairLaunch L;
    int rc = fill_launch(ctx, L);
    if (rc) return rc;
    const long long n = ctx->n_views;
    if (idx4) {
        if (n_pairs < 0) return fail(ctx, ECC_ERR_INVALID, "ecc_pair_maps: bad argument");
        if (n_pairs == 0) return ECC_OK;
        if ((rc = check_indices(ctx, idx4, n_pairs))) return rc;
        if ((rc = stage_indices(ctx, idx4, n_pairs, &L.idx4_d))) return rc;
        L.n_pairs = n_pairs;
    } else {
        L.n_pairs = n * (n - 1) / 2;
        if (L.n_pairs == 0) return ECC_OK;
    }
    const size_t bytes = sizeof(float) * 16 * (size_t)L.n_pairs;
    if (is_device_pointer(K01s)) return launch_pair_maps(ctx, L, K01s);
    size_t cap = ctx->partials_cap * sizeof(float);
    if ((rc = ensure_bytes(ctx, (void**)&ctx->partials_d, &cap, bytes))) return rc;
    ctx->partials_cap = cap / sizeof(float);
    if ((rc = launch_pair_maps(ctx, L, ctx->partials_d))) return rc;
    ECC_CUDA(ctx, cudaMemcpyAsync(K01s, ctx->partials_d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

int ecc_pair_sample_counts(ecc_context* ctx, int* counts)
{
    if (!ctx || !counts) return ECC_ERR_INVALID;
    Guard g(ctx);
    PairLaunch L;
    int rc = fill_launch(ctx, L);
    if (rc) return rc;
    const long long n = ctx->n_views, total = n * (n - 1) / 2;
    if (total == 0) return ECC_OK;
    L.n_pairs = total;
    size_t cap = ctx->counts_cap * sizeof(int);
    if ((rc = ensure_bytes(ctx, (void**)&ctx->counts_d, &cap, sizeof(int) * total))) return rc;
    ctx->counts_cap = cap / sizeof(int);
    if ((rc = launch_pair_counts(ctx, L, ctx->counts_d))) return rc;
    ECC_CUDA(ctx, cudaMemcpyAsync(counts, ctx->counts_d, sizeof(int) * total, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

int ecc_partition_pairs(ecc_context* ctx, int n_parts, long long* bounds)
{
    if (!ctx || !bounds || n_parts < 1) return ECC_ERR_INVALID;
    const long long n = ctx->n_views, total = n * (n - 1) / 2;
    // the cut depends on the geometry and the sampling settings only: keep it while those stand
    const std::vector<double> key = {(double)ctx->geometry_version, (double)n_parts, ctx->object_radius, ctx->dkappa,
                                     (double)ctx->n_t, (double)ctx->step_t, (double)ctx->n_u, (double)ctx->n_v, (double)n};
    if (key == ctx->partition_key && (int)ctx->partition_bounds.size() == n_parts + 1) {
        std::memcpy(bounds, ctx->partition_bounds.data(), sizeof(long long) * (n_parts + 1));
        return ECC_OK;
    }
    // counts, running sum and cuts stay on the device: one small download instead of one int per pair (an optimiser changes
    // the matrices every step, so this runs every step on every rank)
    if (n_parts > 1023) return fail(ctx, ECC_ERR_INVALID, "ecc_partition_pairs: too many parts");
    Guard g(ctx);
    bounds[0] = 0;
    for (int p = 1; p <= n_parts; p++) bounds[p] = total;
    if (total > 0) {
        PairLaunch L;
        int rc = fill_launch(ctx, L);
        if (rc) return rc;
        L.n_pairs = total;
        size_t cap = ctx->counts_cap * sizeof(int);
        // [counts | bounds] in one buffer (bounds 8-byte aligned behind the counts)
        const size_t counts_bytes = round_up(sizeof(int) * (size_t)total, 16);
        if ((rc = ensure_bytes(ctx, (void**)&ctx->counts_d, &cap, counts_bytes + sizeof(long long) * (n_parts + 1)))) return rc;
        ctx->counts_cap = cap / sizeof(int);
        long long* bounds_d = (long long*)((char*)ctx->counts_d + counts_bytes);
        if ((rc = launch_pair_counts(ctx, L, ctx->counts_d))) return rc;
        if ((rc = launch_partition(ctx, ctx->counts_d, total, n_parts, bounds_d))) return rc;
        if ((rc = ensure_pinned(ctx, sizeof(long long) * (n_parts + 1)))) return rc;
        ECC_CUDA(ctx, cudaMemcpyAsync(ctx->pinned_h, bounds_d, sizeof(long long) * (n_parts + 1), cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::memcpy(bounds, ctx->pinned_h, sizeof(long long) * (n_parts + 1));
    }
    ctx->partition_key = key;
    ctx->partition_bounds.assign(bounds, bounds + n_parts + 1);
    return ECC_OK;
}

int ecc_synth_projections(ecc_context* ctx, const double* Ps, int n, int n_u, int n_v, const double* ellipsoids,
                          int n_ellipsoids, int cos_weight, int zero_border, float* images)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!Ps || !images || n < 1 || n_u < 1 || n_v < 1 || (n_ellipsoids > 0 && !ellipsoids))
        return fail(ctx, ECC_ERR_INVALID, "ecc_synth_projections: bad argument");
    if (!is_device_pointer(images)) return fail(ctx, ECC_ERR_INVALID, "ecc_synth_projections: images must be device memory");
    return synth_projections(ctx, Ps, n, n_u, n_v, ellipsoids, n_ellipsoids, cos_weight, zero_border, images);
}

int ecc_profile_enable(ecc_context* ctx, int on)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!on) prof_collect(ctx);
    ctx->profiling = on != 0;
    return ECC_OK;
}

int ecc_profile_reset(ecc_context* ctx)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    prof_collect(ctx);
    for (int f = 0; f < FAM_COUNT; f++) {
        ctx->prof_ms[f] = 0;
        ctx->prof_launches[f] = 0;
    }
    return ECC_OK;
}

int ecc_profile_get(ecc_context* ctx, const char* family, double* total_ms, long long* launches)
{
    if (!ctx || !family) return ECC_ERR_INVALID;
    Guard g(ctx);
    static const char* names[FAM_COUNT] = {"radon", "pairs", "geometry", "reduce", "synth", "stage"};
    prof_collect(ctx);
    for (int f = 0; f < FAM_COUNT; f++)
        if (std::strcmp(family, names[f]) == 0) {
            if (total_ms) *total_ms = ctx->prof_ms[f];
            if (launches) *launches = ctx->prof_launches[f];
            return ECC_OK;
        }
    return fail(ctx, ECC_ERR_INVALID, "unknown kernel family");
}

}  // extern "C"
