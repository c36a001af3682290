// ecc_preprocess.cu -- projection pre-processing on the device (the step in front of the hot path).
//
// WHAT (reference, code/LibEpipolarConsistency/Gui/PreProccess.cpp):
//   PreProccess::process                          :57-144  per pixel scale / bias / -log / clean-up, zeroed and feathered
//                                                          borders (weighting(), EpipolarConsistencyCommon.hxx:30-35),
//                                                          blanked rectangles, flips, separable Gaussian low-pass
//                                                          (HeaderOnly/NRRD/nrrd_lowpass.hxx:18-33,46-79,184-189)
//   PreProccess::apply_weight_cos_principal_ray   :146-166 cosine weighting about the principal point
//   Geometry::getCameraIntrinsics                 LibProjectiveGeometry/ProjectionMatrix.cpp:27-67 (RQ decomposition;
//                                                          here in closed form from the rows of the left 3x3 block)
// HOW (ours): the reference runs this per image on the CPU (OpenMP); at 496 x 1240 x 960 that would dwarf the 0.4 s of
// the whole GPU pipeline.  Here: a batch of images per call, three small kernels, in place, on the context's stream.
// The reference's quirks are kept because they change pixels: the low-pass sums taps -k .. k-1 of a (2k+1)-tap
// normalised Gaussian (nrrd_lowpass.hxx:58,69) and a border of "zero = z" clears z+1 lines at the left / top.
#include <cmath>
#include <vector>

#include "ecc_internal.h"

namespace eccb200 {

namespace {

struct PreParams {
    int n_u, n_v;
    float scale, bias;
    int normalize, apply_log;
    int zero[4], feather[4];  // left, right, bottom, top
    int flip_u, flip_v;
    int n_blanks;
    int taps;  // low-pass half width k (0: off)
};

__device__ __forceinline__ float feather_weight(float x)
{
    if (x < -1.f || x > 1.f) return 0.f;
    const float xx = x * x;
    return 1.f - 2 * xx + xx * xx;
}

// max of every image (Intensity/Normalize, PreProccess.cpp:66-75)
__global__ void __launch_bounds__(1024) image_max_kernel(const float* __restrict__ img, size_t len, float* __restrict__ out)
{
    __shared__ float part[1024];
    const float* p = img + (size_t)blockIdx.x * len;
    float m = p[0];
    for (size_t k = threadIdx.x; k < len; k += 1024) m = fmaxf(m, p[k]);
    part[threadIdx.x] = m;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (threadIdx.x < w) part[threadIdx.x] = fmaxf(part[threadIdx.x], part[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = part[0];
}

// intensity, borders, blanks (PreProccess.cpp:77-117) and the flips (:119-133) as the index map of the store
__global__ void pointwise_kernel(const float* __restrict__ src, float* __restrict__ dst, PreParams P, const float* __restrict__ maxima,
                                 const int* __restrict__ blanks)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.n_u) return;
    const size_t base = (size_t)blockIdx.z * P.n_u * P.n_v;
    float scale = P.scale, bias = P.bias;
    if (P.normalize) {
        bias = 0.f;
        scale = P.scale / maxima[blockIdx.z];
    }
    float v = __fadd_rn(__fmul_rn(src[base + (size_t)y * P.n_u + x], scale), bias);
    if (P.apply_log) v = -logf(v);
    if (v < 0.f || isnan(v) || isinf(v)) v = 0.f;
    const int w = P.n_u, h = P.n_v;
    {   // left: columns b = 0 .. zero+feather-1
        const int b = x;
        if (b < P.zero[0] + P.feather[0]) v *= (b <= P.zero[0]) ? 0.f : feather_weight(1.f - (float)(b - P.zero[0]) / P.feather[0]);
    }
    {   // right: columns w-b, b = 1 .. zero+feather
        const int b = w - x;
        if (b >= 1 && b <= P.zero[1] + P.feather[1]) v *= (b <= P.zero[1]) ? 0.f : feather_weight(1.f - (float)(b - P.zero[1]) / P.feather[1]);
    }
    {   // bottom: rows h-b, b = 1 .. zero+feather
        const int b = h - y;
        if (b >= 1 && b <= P.zero[2] + P.feather[2]) v *= (b <= P.zero[2]) ? 0.f : feather_weight(1.f - (float)(b - P.zero[2]) / P.feather[2]);
    }
    {   // top: rows b = 0 .. zero+feather-1
        const int b = y;
        if (b < P.zero[3] + P.feather[3]) v *= (b <= P.zero[3]) ? 0.f : feather_weight(1.f - (float)(b - P.zero[3]) / P.feather[3]);
    }
    for (int k = 0; k < P.n_blanks; k++) {
        const int* q = blanks + 4 * k;  // x0, y0, x1, y1
        if (x >= max(0, q[0]) && x < q[2] && y >= max(0, q[1]) && y < q[3]) v = 0.f;
    }
    const int xo = P.flip_u ? w - 1 - x : x, yo = P.flip_v ? h - 1 - y : y;
    dst[base + (size_t)yo * P.n_u + xo] = v;
}

// one pass of the separable low-pass along x (DIR 0) or y (DIR 1): taps o = -k .. k-1, clamped reads, fp64 sum
template <int DIR>
__global__ void lowpass_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_u, int n_v, int k, const double* __restrict__ kernel)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= n_u) return;
    const float* s = src + (size_t)blockIdx.z * n_u * n_v;
    double sum = 0;
    for (int o = -k; o < k; o++) {
        if (DIR == 0) {
            const int xo = min(max(x + o, 0), n_u - 1);
            sum += s[(size_t)y * n_u + xo] * kernel[o + k];
        } else {
            const int yo = min(max(y + o, 0), n_v - 1);
            sum += s[(size_t)yo * n_u + x] * kernel[o + k];
        }
    }
    dst[(size_t)blockIdx.z * n_u * n_v + (size_t)y * n_u + x] = (float)sum;
}

// PreProccess.cpp:146-166; intr: per image sdd_px, ppu, ppv (sdd_px == 0: matrix is zero -> image untouched)
__global__ void cosine_kernel(float* __restrict__ img, int n_u, int n_v, const float* __restrict__ intr)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= n_u) return;
    const float sdd_px = intr[3 * blockIdx.z], ppu = intr[3 * blockIdx.z + 1], ppv = intr[3 * blockIdx.z + 2];
    if (sdd_px == 0.f) return;
    const float pou = (float)u - ppu, pov = (float)v - ppv;
    const float cos_weight = sdd_px / sqrtf(pou * pou + pov * pov + sdd_px * sdd_px);
    img[(size_t)blockIdx.z * n_u * n_v + (size_t)v * n_u + u] *= cos_weight;
}

}  // namespace

// K(0,0), K(0,2), K(1,2) of the RQ decomposition M = K R of the left 3x3 block (K upper triangular, positive
// diagonal, K(2,2) = 1): K K^T = M M^T / |m3|^2.
void camera_intrinsics_host(const double* P, double* fu, double* u0, double* v0)
{
    const double m1[3] = {P[0], P[3], P[6]}, m2[3] = {P[1], P[4], P[7]}, m3[3] = {P[2], P[5], P[8]};
    auto dot = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    const double n3 = dot(m3, m3);
    if (!(n3 > 0)) { *fu = *u0 = *v0 = 0; return; }
    *u0 = dot(m1, m3) / n3;
    *v0 = dot(m2, m3) / n3;
    const double fv = sqrt(fmax(dot(m2, m2) / n3 - *v0 * *v0, 0.0));
    const double s = fv > 0 ? (dot(m1, m2) / n3 - *u0 * *v0) / fv : 0.0;
    *fu = sqrt(fmax(dot(m1, m1) / n3 - s * s - *u0 * *u0, 0.0));
}

int preprocess_batch(ecc_context* ctx, float* images_d, int n, int n_u, int n_v, const ecc_preprocess_params* pp, const double* Ps_h)
{
    PreParams P;
    P.n_u = n_u;
    P.n_v = n_v;
    P.scale = (float)pp->scale;
    P.bias = (float)pp->bias;
    P.normalize = pp->normalize;
    P.apply_log = pp->apply_log;
    for (int k = 0; k < 4; k++) { P.zero[k] = pp->border_zero[k]; P.feather[k] = pp->border_feather[k]; }
    P.flip_u = pp->flip_u;
    P.flip_v = pp->flip_v;
    P.n_blanks = pp->n_blanks;
    P.taps = (pp->gaussian_sigma > 0 && pp->half_kernel_width > 1) ? pp->half_kernel_width : 0;
    const size_t len = (size_t)n_u * n_v;
    // scratch: a second image buffer, maxima, blanks, kernel, intrinsics
    const size_t small_bytes = sizeof(float) * n + sizeof(int) * 4 * (size_t)(P.n_blanks > 0 ? P.n_blanks : 1) + sizeof(double) * (2 * (size_t)P.taps + 1) + sizeof(float) * 3 * n + 64;
    int rc = ensure_bytes(ctx, (void**)&ctx->pre_work_d, &ctx->pre_work_bytes, sizeof(float) * len * n);
    if (rc) return rc;
    rc = ensure_bytes(ctx, (void**)&ctx->pre_small_d, &ctx->pre_small_bytes, small_bytes);
    if (rc) return rc;
    char* small = (char*)ctx->pre_small_d;
    double* kernel_d = (double*)small;  // first: 8-byte alignment
    float* maxima_d = (float*)(kernel_d + 2 * P.taps + 1);
    float* intr_d = maxima_d + n;
    int* blanks_d = (int*)(intr_d + 3 * n);
    std::vector<double> kernel(2 * P.taps + 1, 1.0);
    if (P.taps) {  // NRRD::gaussianKernel, nrrd_lowpass.hxx:18-33
        double sum = 0;
        for (int x = -P.taps; x <= P.taps; x++) sum += (kernel[x + P.taps] = exp(-0.5 * pow(x / pp->gaussian_sigma, 2)));
        for (auto& v : kernel) v /= sum;
    }
    std::vector<float> intr(3 * (size_t)n, 0.f);
    if (pp->cos_weight && Ps_h)
        for (int i = 0; i < n; i++) {
            double fu, u0, v0;
            camera_intrinsics_host(Ps_h + 12 * i, &fu, &u0, &v0);
            intr[3 * i] = (float)fu; intr[3 * i + 1] = (float)u0; intr[3 * i + 2] = (float)v0;
        }
    ECC_CUDA(ctx, cudaMemcpyAsync(kernel_d, kernel.data(), sizeof(double) * kernel.size(), cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(intr_d, intr.data(), sizeof(float) * intr.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (P.n_blanks > 0)
        ECC_CUDA(ctx, cudaMemcpyAsync(blanks_d, pp->blanks, sizeof(int) * 4 * P.n_blanks, cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host vectors above go out of scope

    const dim3 block(128), grid((n_u + 127) / 128, n_v, n);
    const int slot = prof_begin(ctx, FAM_SYNTH);
    if (P.normalize) image_max_kernel<<<n, 1024, 0, ctx->stream>>>(images_d, len, maxima_d);
    pointwise_kernel<<<grid, block, 0, ctx->stream>>>(images_d, ctx->pre_work_d, P, maxima_d, blanks_d);
    if (P.taps) {
        lowpass_kernel<0><<<grid, block, 0, ctx->stream>>>(ctx->pre_work_d, images_d, n_u, n_v, P.taps, kernel_d);
        lowpass_kernel<1><<<grid, block, 0, ctx->stream>>>(images_d, ctx->pre_work_d, n_u, n_v, P.taps, kernel_d);
    }
    ECC_CUDA(ctx, cudaMemcpyAsync(images_d, ctx->pre_work_d, sizeof(float) * len * n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (pp->cos_weight && Ps_h) cosine_kernel<<<grid, block, 0, ctx->stream>>>(images_d, n_u, n_v, intr_d);
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

}  // namespace eccb200
