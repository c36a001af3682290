// ecc_radon.cu -- Radon-intermediate kernels of libecc_b200 (sm_100a).
//
// WHAT (reference, code/LibEpipolarConsistency/RadonIntermediate.cu):
//   radonDerivative<derivative>  :31-143   per (alpha,t) bin: line integral of the image with step
//                                          0.66 px, or the difference of two such integrals one
//                                          pixel apart (t-derivative), optional sqrt/log
//   computeDerivLineIntegrals    :149-170  launcher (block 32x4, lanes = adjacent ALPHA)
// HOW (ours): a whole batch of projections per launch; warp lanes are adjacent t bins of the SAME
// angle, so the 32 lines of a warp are parallel, 2 px apart, equally long and march through the
// image side by side (coherent texture footprints, no divergence), and the 8 warps of a CTA are
// 8 neighbouring angles over the same strip of the image (L1 reuse).  Sample positions follow the
// reference exactly (clipped entry point, t += 0.66f accumulation, +-1/2 px derivative lines).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "ecc_geometry.cuh"
#include "ecc_internal.h"

namespace eccb200 {

namespace {

constexpr int kLanesT = 32;   // adjacent t bins per warp
constexpr int kAngles = 8;    // adjacent angles per CTA
constexpr float kStep = 0.66f;

struct BinLine {
    float o0, o1, d0, d1, t, t_max;
    bool valid;
};

// Bin (ix,iy) -> line, clipped against the image inset by one pixel (RadonIntermediate.cu:44-92).
__device__ __forceinline__ BinLine bin_line(int ix, int iy, int n_alpha, int n_t, float n_u,
                                            float n_v)
{
    BinLine L;
    const float x_rel = ix / (float)n_alpha - 0.5f;
    const float y_rel = iy / (float)n_t - 0.5f;
    const float diag = sqrtf(n_u * n_u + n_v * n_v);
    const float alpha = x_rel * ECC_PI_F;
    const float tau = y_rel * diag;
    const float l0 = -sinf(alpha);
    const float l1 = cosf(alpha);
    float l2 = -tau;
    l2 += -0.5f * n_u * l0 - 0.5f * n_v * l1;
    L.o0 = -l2 * l0;
    L.o1 = -l2 * l1;
    L.d0 = l1;
    L.d1 = -l0;
    float ta = (1.f - L.o0) / L.d0, tb = (n_u - 1.f - L.o0) / L.d0;
    float tc = (1.f - L.o1) / L.d1, td = (n_v - 1.f - L.o1) / L.d1;
    if (L.d0 * L.d0 < 1e-12f) { ta = -1e10f; tb = 1e10f; }
    if (L.d1 * L.d1 < 1e-12f) { tc = -1e10f; td = 1e10f; }
    // middle two of the four intersections = entry and exit of the box
    const float lo1 = fminf(ta, tb), hi1 = fmaxf(ta, tb);
    const float lo2 = fminf(tc, td), hi2 = fmaxf(tc, td);
    L.t = fmaxf(lo1, lo2);
    L.t_max = fminf(hi1, hi2);
    // a line that misses the box has its "middle two" in the other order
    if (fminf(hi1, hi2) < fmaxf(lo1, lo2)) { L.t = fminf(hi1, hi2); L.t_max = fmaxf(lo1, lo2); }
    const float pu = L.o0 + L.t * L.d0, pv = L.o1 + L.t * L.d1;
    const bool inside = (pu <= n_u && pv <= n_v && pu >= 0.f && pv >= 0.f);
    L.valid = inside && !(L.t_max <= L.t);
    return L;
}

__device__ __forceinline__ float post_process(float r, int post)
{
    if (post == ECC_POST_SQRT) return r < 0.f ? -sqrtf(-r) : sqrtf(r);
    if (post == ECC_POST_LOG) return r < 0.f ? -logf(-r + 1.f) : logf(r + 1.f);
    return r;
}

// Exact-fp32 bilinear sample through the gather path: the four texels of the cell are fetched with
// one tld4 aimed at the cell centre, the weights are computed here in full precision.
__device__ __forceinline__ float sample_exact(cudaTextureObject_t tex, float x, float y)
{
    const float xb = x - 0.5f, yb = y - 0.5f;
    const float fx = floorf(xb), fy = floorf(yb);
    const float wx = xb - fx, wy = yb - fy;
    const float4 g = tex2Dgather<float4>(tex, fx + 1.0f, fy + 1.0f, 0);
    // gather order: (i,j+1) (i+1,j+1) (i+1,j) (i,j)
    return (1.f - wx) * (1.f - wy) * g.w + wx * (1.f - wy) * g.z + (1.f - wx) * wy * g.x +
           wx * wy * g.y;
}

template <int INTERP>
__device__ __forceinline__ float sample_img(cudaTextureObject_t tex, float x, float y)
{
    if (INTERP == ECC_INTERP_TEXTURE) return tex2D<float>(tex, x, y);
    return sample_exact(tex, x, y);
}

// Bin tiling of a CTA.  Lanes: a quad of four consecutive lanes covers qa x 4/qa (angle x t) bins, the eight
// quads of a warp are arranged ga along the angle axis x 8/ga along t, so a warp covers
// la = qa*ga angles x lt = 32/la t bins.  The warps of the CTA (blockDim.y) are arranged wa along the angle axis
// x blockDim.y/wa along t.  Examples: qa=1,ga=1: 32 parallel lines 2 px apart; qa=4,ga=8: a fan of 32 lines
// through one t (the reference's arrangement); qa=2,ga=1: quads of 2 angles x 2 t, 16 t bins per warp.
struct Tiling {
    int qa, ga, wa;
};

template <bool DERIV, int INTERP>
__global__ void __launch_bounds__(256)
radon_kernel(const cudaTextureObject_t* __restrict__ images, int n_u_i, int n_v_i, int n_alpha,
             int n_t, int post, Tiling tl, float* __restrict__ out)
{
    const int la = tl.qa * tl.ga, lt = 32 / la, qt = 4 / tl.qa, wt = blockDim.y / tl.wa;
    const int lane = threadIdx.x, warp = threadIdx.y, quad = lane >> 2, ql = lane & 3;
    const int da = (quad % tl.ga) * tl.qa + ql % tl.qa;  // angle offset inside the warp tile
    const int dt = (quad / tl.ga) * qt + ql / tl.qa;     // t offset inside the warp tile
    const int ix = (blockIdx.x * tl.wa + warp % tl.wa) * la + da;  // angle bin
    const int iy = (blockIdx.y * wt + warp / tl.wa) * lt + dt;     // t bin
    if (ix >= n_alpha || iy >= n_t) return;
    const cudaTextureObject_t tex = images[blockIdx.z];
    float* dst = out + (size_t)blockIdx.z * n_t * n_alpha + (size_t)iy * n_alpha + ix;

    BinLine L = bin_line(ix, iy, n_alpha, n_t, (float)n_u_i, (float)n_v_i);
    if (!L.valid) {
        *dst = 0.f;
        return;
    }
    float o0 = L.o0 + 0.5f, o1 = L.o1 + 0.5f;  // texel centres
    const float d0 = L.d0, d1 = L.d1, t_max = L.t_max;
    float t = L.t;
    float sum = 0.f;
    if (!DERIV) {
        for (; t <= t_max; t += kStep) sum += sample_img<INTERP>(tex, o0 + t * d0, o1 + t * d1);
        *dst = sum * kStep;
        return;
    }
    // two parallel lines half a pixel either side of the bin's line
    o0 -= 0.5f * d1;
    o1 += 0.5f * d0;
    float sumo = 0.f;
    for (; t <= t_max; t += kStep) {
        sum += sample_img<INTERP>(tex, o0 + t * d0, o1 + t * d1);
        sumo += sample_img<INTERP>(tex, o0 + t * d0 + d1, o1 + t * d1 - d0);
    }
    *dst = post_process((sum - sumo) * kStep, post);
}

// Derivative filter with the two lines of a bin on neighbouring lanes: a warp is 16 adjacent t bins x 2 lines,
// i.e. 32 parallel lines about one pixel apart.
template <int INTERP>
__global__ void __launch_bounds__(256)
radon_kernel_split(const cudaTextureObject_t* __restrict__ images, int n_u_i, int n_v_i, int n_alpha,
                   int n_t, int post, float* __restrict__ out)
{
    const int lane_global = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = lane_global >> 1;                              // t bin
    const int which = lane_global & 1;                            // 0: line at +1/2, 1: the line one pixel further
    const int ix = blockIdx.y * blockDim.y + threadIdx.y;         // angle bin
    if (ix >= n_alpha) return;
    const bool in_range = iy < n_t;
    const cudaTextureObject_t tex = images[blockIdx.z];
    float sum = 0.f;
    bool valid = false;
    if (in_range) {
        BinLine L = bin_line(ix, iy, n_alpha, n_t, (float)n_u_i, (float)n_v_i);
        valid = L.valid;
        if (valid) {
            float o0 = L.o0 + 0.5f, o1 = L.o1 + 0.5f;
            const float d0 = L.d0, d1 = L.d1, t_max = L.t_max;
            o0 -= 0.5f * d1;
            o1 += 0.5f * d0;
            if (which == 0) {
                for (float t = L.t; t <= t_max; t += kStep) sum += sample_img<INTERP>(tex, o0 + t * d0, o1 + t * d1);
            } else {
                for (float t = L.t; t <= t_max; t += kStep) sum += sample_img<INTERP>(tex, o0 + t * d0 + d1, o1 + t * d1 - d0);
            }
        }
    }
    const float other = __shfl_xor_sync(0xffffffffu, sum, 1);
    if (in_range && which == 0) {
        float* dst = out + (size_t)blockIdx.z * n_t * n_alpha + (size_t)iy * n_alpha + ix;
        *dst = valid ? post_process((sum - other) * kStep, post) : 0.f;
    }
}

// Ramp filter along t (reference RadonIntermediate.cu:173-237: batched cuFFT R2C over the columns, bin k multiplied by
// k * (-0.5f / (n_t * (n_t/2+1))), unnormalised C2R).  That is a circular convolution of every alpha column with the
// real kernel g[j] = sum_k H_k e^{2 pi i jk/n} (H extended Hermitian); g is tabulated once per n_t on the host in
// fp64 and applied directly -- O(n_t^2) per column, 0.45 GFLOP per 768x768 intermediate, not on the hot path.
constexpr int kRampCols = 16;
__global__ void __launch_bounds__(256)
ramp_kernel(float* __restrict__ data, const float* __restrict__ g, int n_alpha, int n_t)
{
    extern __shared__ float ramp_smem[];
    float* gs = ramp_smem;                    // [n_t]
    float* xs = ramp_smem + n_t;              // [n_t][kRampCols]
    float* img = data + (size_t)blockIdx.y * n_t * n_alpha;
    const int c0 = blockIdx.x * kRampCols;
    for (int k = threadIdx.x; k < n_t; k += blockDim.x) gs[k] = g[k];
    for (int k = threadIdx.x; k < n_t * kRampCols; k += blockDim.x) {
        const int m = k / kRampCols, c = k - m * kRampCols;
        xs[k] = (c0 + c < n_alpha) ? img[(size_t)m * n_alpha + c0 + c] : 0.f;
    }
    __syncthreads();
    const int c = threadIdx.x % kRampCols, lane_t = threadIdx.x / kRampCols, lanes = blockDim.x / kRampCols;
    if (c0 + c >= n_alpha) return;
    for (int t = lane_t; t < n_t; t += lanes) {
        float acc = 0.f;
        int j = t;  // (t - m) mod n_t
        for (int m = 0; m < n_t; m++) {
            acc = fmaf(xs[m * kRampCols + c], gs[j], acc);
            j = (j == 0) ? n_t - 1 : j - 1;
        }
        img[(size_t)t * n_alpha + c0 + c] = acc;
    }
}

// Work counter: the number of bilinear samples the Radon kernel takes per projection for this geometry (same
// clipping and the same t += 0.66f walk, no fetches).
__global__ void radon_count_kernel(int n_u_i, int n_v_i, int n_alpha, int n_t, int lines_per_bin,
                                   unsigned long long* total)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    const int iy = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long cnt = 0;
    if (ix < n_alpha && iy < n_t) {
        BinLine L = bin_line(ix, iy, n_alpha, n_t, (float)n_u_i, (float)n_v_i);
        if (L.valid)
            for (float t = L.t; t <= L.t_max; t += kStep) cnt += lines_per_bin;
    }
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && cnt) atomicAdd(total, cnt);
}

int ensure_pool(ecc_context* ctx, int n_u, int n_v, int count)
{
    ImagePool& P = ctx->pool;
    if (P.n_u == n_u && P.n_v == n_v && P.count >= count) return ECC_OK;
    // launches queued earlier on the stream may still sample the arrays that are about to go
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_image_pool(ctx);
    P.n_u = n_u;
    P.n_v = n_v;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    for (int k = 0; k < count; k++) {
        cudaArray_t arr = nullptr;
        ECC_CUDA(ctx, cudaMallocArray(&arr, &desc, n_u, n_v, cudaArrayTextureGather));
        P.arrays.push_back(arr);
        cudaResourceDesc res = {};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr;
        cudaTextureDesc td = {};
        td.normalizedCoords = 0;
        td.filterMode = cudaFilterModeLinear;
        td.addressMode[0] = cudaAddressModeClamp;
        td.addressMode[1] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        ECC_CUDA(ctx, cudaCreateTextureObject(&tex, &res, &td, nullptr));
        P.tex_h.push_back(tex);
        P.count = k + 1;
    }
    ECC_CUDA(ctx, cudaMalloc(&P.tex_d, sizeof(cudaTextureObject_t) * count));
    ECC_CUDA(ctx, cudaMemcpyAsync(P.tex_d, P.tex_h.data(), sizeof(cudaTextureObject_t) * count,
                                  cudaMemcpyHostToDevice, ctx->stream));
    return ECC_OK;
}

}  // namespace

// In-place ramp filter of n intermediates (device memory).
int ramp_filter(ecc_context* ctx, float* dtrs_d, int n, int n_alpha, int n_t)
{
    if (ctx->ramp_n_t != n_t) {
        // g[j] = H_0 + 2 sum_{0<k<n/2} H_k cos(2 pi jk/n) (+ H_{n/2} (-1)^j for even n), H_k = k * scale in fp32 as the
        // reference forms it (RadonIntermediate.cu:181-182,218)
        const int n_theta = n_t / 2 + 1;
        const float scale = -0.5f / (n_t * n_theta);
        std::vector<float> g(n_t);
        const double w = 2.0 * 3.14159265358979323846 / n_t;
        for (int j = 0; j < n_t; j++) {
            double acc = 0.0;
            for (int k = 1; k < n_theta; k++) {
                const double Hk = (double)((float)k * scale);
                const long long jk = ((long long)j * k) % n_t;
                const bool nyquist = (n_t % 2 == 0) && (k == n_t / 2);
                acc += (nyquist ? 1.0 : 2.0) * Hk * cos(w * (double)jk);
            }
            g[j] = (float)acc;
        }
        if (ctx->ramp_g_d) cudaFree(ctx->ramp_g_d);
        ctx->ramp_g_d = nullptr;
        ctx->ramp_n_t = 0;
        ECC_CUDA(ctx, cudaMalloc(&ctx->ramp_g_d, sizeof(float) * n_t));
        ECC_CUDA(ctx, cudaMemcpyAsync(ctx->ramp_g_d, g.data(), sizeof(float) * n_t, cudaMemcpyHostToDevice, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // g lives on this stack frame
        ctx->ramp_n_t = n_t;
    }
    const size_t smem = sizeof(float) * ((size_t)n_t + (size_t)n_t * kRampCols);
    if (smem > 220 * 1024) return fail(ctx, ECC_ERR_UNSUPPORTED, "ramp filter: n_t too large for the shared-memory tile");
    ECC_CUDA(ctx, cudaFuncSetAttribute(ramp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int slot = prof_begin(ctx, FAM_RADON);
    ramp_kernel<<<dim3((n_alpha + kRampCols - 1) / kRampCols, n), 256, smem, ctx->stream>>>(dtrs_d, ctx->ramp_g_d, n_alpha, n_t);
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

int radon_num_samples(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int filter, double* count)
{
    unsigned long long* total_d = nullptr;
    ECC_CUDA(ctx, cudaMalloc(&total_d, sizeof(unsigned long long)));
    ECC_CUDA(ctx, cudaMemsetAsync(total_d, 0, sizeof(unsigned long long), ctx->stream));
    dim3 block(32, 8), grid((n_alpha + 31) / 32, (n_t + 7) / 8);
    radon_count_kernel<<<grid, block, 0, ctx->stream>>>(n_u, n_v, n_alpha, n_t, filter == ECC_FILTER_DERIVATIVE ? 2 : 1, total_d);
    unsigned long long total = 0;
    ECC_CUDA(ctx, cudaMemcpyAsync(&total, total_d, sizeof(total), cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(total_d);
    *count = (double)total;
    return ECC_OK;
}

void free_image_pool(ecc_context* ctx)
{
    ImagePool& P = ctx->pool;
    for (auto t : P.tex_h) cudaDestroyTextureObject(t);
    for (auto a : P.arrays) cudaFreeArray(a);
    if (P.tex_d) cudaFree(P.tex_d);
    P = ImagePool();
}

// images_d / out_d are device pointers.  The batch is processed in chunks of at most kPool images:
// copy the chunk into the array pool (device-to-device, async), one kernel launch per chunk.
int radon_batch(ecc_context* ctx, const float* images_d, int n_images, int n_u, int n_v,
                int n_alpha, int n_t, int filter, int post, int interp, float* out_d, QuadPart part)
{
    constexpr int kPool = 64;  // projections per launch (texture engines: one CUDA array per projection)
    // the quad kernel is persistent: every launch ends with a tail (CTAs run dry one by one) and starts with its image
    // staging, so fewer, larger launches (2.4 GB of staging at the C3 image size)
    constexpr int kPoolQuad = 128;
    const int pool = n_images < kPool ? n_images : kPool;
    const bool deriv = (filter == ECC_FILTER_DERIVATIVE);  // ramp: plain line integrals first (RadonIntermediate.cu:160-168)
    if (!part.whole() && !(interp == ECC_INTERP_HYBRID_STATIC && deriv))
        return fail(ctx, ECC_ERR_INVALID, "parts of a quad of projections: only the static-split engine (ECC_INTERP_HYBRID_STATIC, derivative filter) shares a quad");
    // development knob: ECC_HYBRID_QUADS=0 keeps the one-image-per-item hybrid kernel for everything
    static const int use_quads = getenv("ECC_HYBRID_QUADS") ? atoi(getenv("ECC_HYBRID_QUADS")) : 1;
    if (interp == ECC_INTERP_HYBRID_STATIC && deriv) {
        // reproducible mode: every projection through the quad kernel with its static split (a remainder is padded to a
        // quad), so that a projection's bins do not depend on the batch it arrives in
        for (int first = 0; first < n_images; first += kPoolQuad) {
            const int n = (n_images - first < kPoolQuad) ? n_images - first : kPoolQuad;
            const int rc4 = radon_hybrid4_launch(ctx, images_d + (size_t)first * n_u * n_v, n, n_u, n_v, n_alpha, n_t, post,
                                                 out_d + (size_t)first * n_t * n_alpha, true, part.sub(first == 0, first + n == n_images));
            if (rc4) return rc4;
        }
        return ECC_OK;
    }
    if (interp == ECC_INTERP_HYBRID_STATIC) interp = ECC_INTERP_TEXTURE;  // other filters: the texture engine is reproducible
    if (interp == ECC_INTERP_HYBRID && deriv && use_quads && n_images >= 3) {
        // quads of projections through the quad kernel (it stages its images itself, four interleaved per texel); a
        // remainder of three is padded to a quad, one or two left-over images take the one-image hybrid kernel below
        const int rem = n_images % 4;
        const int n_quad_images = (rem == 3) ? n_images : n_images - rem;
        for (int first = 0; first < n_quad_images; first += kPoolQuad) {
            const int n = (n_quad_images - first < kPoolQuad) ? n_quad_images - first : kPoolQuad;
            const int rc4 = radon_hybrid4_launch(ctx, images_d + (size_t)first * n_u * n_v, n, n_u, n_v, n_alpha, n_t, post,
                                                 out_d + (size_t)first * n_t * n_alpha);
            if (rc4) return rc4;
        }
        if (n_quad_images == n_images) return ECC_OK;
        return radon_batch(ctx, images_d + (size_t)n_quad_images * n_u * n_v, n_images - n_quad_images, n_u, n_v, n_alpha, n_t,
                           filter, post, interp, out_d + (size_t)n_quad_images * n_t * n_alpha);
    }
    int rc = ensure_pool(ctx, n_u, n_v, pool);
    if (rc) return rc;
    for (int first = 0; first < n_images; first += pool) {
        const int n = (n_images - first < pool) ? n_images - first : pool;
        for (int k = 0; k < n; k++) {
            const float* src = images_d + (size_t)(first + k) * n_u * n_v;
            ECC_CUDA(ctx, cudaMemcpy2DToArrayAsync(ctx->pool.arrays[k], 0, 0, src,
                                                   sizeof(float) * n_u, sizeof(float) * n_u, n_v,
                                                   cudaMemcpyDeviceToDevice, ctx->stream));
        }
        // development knobs (environment): ECC_RADON_QA/GA lane tiling (QA=0: split-line kernel), ECC_RADON_WA warps
        // along alpha, ECC_RADON_BY warps per CTA
        static const int qa = getenv("ECC_RADON_QA") ? atoi(getenv("ECC_RADON_QA")) : 2;
        static const int ga = getenv("ECC_RADON_GA") ? atoi(getenv("ECC_RADON_GA")) : 1;
        static const int by = getenv("ECC_RADON_BY") ? atoi(getenv("ECC_RADON_BY")) : 8;
        static const int wa_env = getenv("ECC_RADON_WA") ? atoi(getenv("ECC_RADON_WA")) : 4;
        const int la = qa;
        dim3 block(kLanesT, by);
        float* dst = out_d + (size_t)first * n_t * n_alpha;
        const cudaTextureObject_t* texs = ctx->pool.tex_d;
        if (interp == ECC_INTERP_HYBRID && deriv) {
            const int rc2 = radon_hybrid_launch(ctx, texs, images_d + (size_t)first * n_u * n_v, n, n_u, n_v, n_alpha, n_t, post, dst);
            if (rc2) return rc2;
            continue;
        }
        const int slot = prof_begin(ctx, FAM_RADON);
        if (la == 0 && deriv) {
            dim3 grid((2 * n_t + kLanesT - 1) / kLanesT, (n_alpha + by - 1) / by, n);
            if (interp != ECC_INTERP_EXACT) radon_kernel_split<ECC_INTERP_TEXTURE><<<grid, block, 0, ctx->stream>>>(texs, n_u, n_v, n_alpha, n_t, post, dst);
            else radon_kernel_split<ECC_INTERP_EXACT><<<grid, block, 0, ctx->stream>>>(texs, n_u, n_v, n_alpha, n_t, post, dst);
        } else {
            Tiling tl;
            tl.qa = qa > 0 ? qa : 4;
            tl.ga = ga;
            tl.wa = wa_env;
            const int tile_a = tl.qa * tl.ga * tl.wa, tile_t = (32 / (tl.qa * tl.ga)) * (by / tl.wa);
            dim3 grid((n_alpha + tile_a - 1) / tile_a, (n_t + tile_t - 1) / tile_t, n);
#define ECC_LAUNCH(K) K<<<grid, block, 0, ctx->stream>>>(texs, n_u, n_v, n_alpha, n_t, post, tl, dst)
            if (interp != ECC_INTERP_EXACT) { if (deriv) ECC_LAUNCH((radon_kernel<true, ECC_INTERP_TEXTURE>)); else ECC_LAUNCH((radon_kernel<false, ECC_INTERP_TEXTURE>)); }
            else { if (deriv) ECC_LAUNCH((radon_kernel<true, ECC_INTERP_EXACT>)); else ECC_LAUNCH((radon_kernel<false, ECC_INTERP_EXACT>)); }
#undef ECC_LAUNCH
        }
        prof_end(ctx, slot);
        ECC_CUDA(ctx, cudaGetLastError());
        if (filter == ECC_FILTER_RAMP) {
            const int rc3 = ramp_filter(ctx, dst, n, n_alpha, n_t);
            if (rc3) return rc3;
        }
        // multi-GPU team: these engines do not mirror their stores themselves (the hybrid kernels do)
        if (ctx->team.mirror_radon) {
            const int rc4 = team_publish(ctx, dst, sizeof(float) * (size_t)n * n_t * n_alpha);
            if (rc4) return rc4;
        }
    }
    return ECC_OK;
}

}  // namespace eccb200
