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

template <bool DERIV, int INTERP>
__global__ void __launch_bounds__(kLanesT* kAngles)
radon_kernel(const cudaTextureObject_t* __restrict__ images, int n_u_i, int n_v_i, int n_alpha,
             int n_t, int post, float* __restrict__ out)
{
    const int iy = blockIdx.x * kLanesT + threadIdx.x;  // t bin: lanes
    const int ix = blockIdx.y * kAngles + threadIdx.y;  // angle bin: warps
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

int ensure_pool(ecc_context* ctx, int n_u, int n_v, int count)
{
    ImagePool& P = ctx->pool;
    if (P.n_u == n_u && P.n_v == n_v && P.count >= count) return ECC_OK;
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
                int n_alpha, int n_t, int filter, int post, int interp, float* out_d)
{
    constexpr int kPool = 32;
    const int pool = n_images < kPool ? n_images : kPool;
    int rc = ensure_pool(ctx, n_u, n_v, pool);
    if (rc) return rc;
    const bool deriv = (filter == ECC_FILTER_DERIVATIVE);
    for (int first = 0; first < n_images; first += pool) {
        const int n = (n_images - first < pool) ? n_images - first : pool;
        for (int k = 0; k < n; k++) {
            const float* src = images_d + (size_t)(first + k) * n_u * n_v;
            ECC_CUDA(ctx, cudaMemcpy2DToArrayAsync(ctx->pool.arrays[k], 0, 0, src,
                                                   sizeof(float) * n_u, sizeof(float) * n_u, n_v,
                                                   cudaMemcpyDeviceToDevice, ctx->stream));
        }
        dim3 block(kLanesT, kAngles);
        dim3 grid((n_t + kLanesT - 1) / kLanesT, (n_alpha + kAngles - 1) / kAngles, n);
        float* dst = out_d + (size_t)first * n_t * n_alpha;
        const int slot = prof_begin(ctx, FAM_RADON);
        if (interp == ECC_INTERP_TEXTURE) {
            if (deriv) radon_kernel<true, ECC_INTERP_TEXTURE><<<grid, block, 0, ctx->stream>>>(ctx->pool.tex_d, n_u, n_v, n_alpha, n_t, post, dst);
            else radon_kernel<false, ECC_INTERP_TEXTURE><<<grid, block, 0, ctx->stream>>>(ctx->pool.tex_d, n_u, n_v, n_alpha, n_t, post, dst);
        } else {
            if (deriv) radon_kernel<true, ECC_INTERP_EXACT><<<grid, block, 0, ctx->stream>>>(ctx->pool.tex_d, n_u, n_v, n_alpha, n_t, post, dst);
            else radon_kernel<false, ECC_INTERP_EXACT><<<grid, block, 0, ctx->stream>>>(ctx->pool.tex_d, n_u, n_v, n_alpha, n_t, post, dst);
        }
        prof_end(ctx, slot);
        ECC_CUDA(ctx, cudaGetLastError());
    }
    return ECC_OK;
}

}  // namespace eccb200
