// ecc_direct.cu -- the direct metric (no Radon intermediates) of libecc_b200, sm_100a.
//
// WHAT (reference, code/LibEpipolarConsistency/):
//   MetricDirect, computeForImagePair      EpipolarConsistencyDirect.h:17-60, .cpp:64-270: per image pair the host derives the
//                                          pencil of epipolar planes, one pair of epipolar lines per plane (Eigen, fp64),
//                                          uploads them, launches
//   kernel_computeLineIntegrals            EpipolarConsistencyDirect.cu:31-119 once per image (one thread per line: derivative
//                                          of the line integral by two parallel lines half a pixel either side, step 0.4 px;
//                                          or the fan-beam weighted integral of RectifiedFBCC.h), downloads both signals
//                                          and sums the squared differences on the host: 4 launches + 5 copies per pair.
// HOW (ours): ONE launch for all pairs of a data set.  The projection images stay resident as textures; a per-view and a
// per-pair kernel derive the fp64 geometry; the line kernel's CTAs are 32 consecutive epipolar planes of one pair -- warp 0
// integrates them in the pair's first image, warp 1 in its second, each thread deriving its own line (and fan-beam record) in
// registers -- and leave one fixed-order partial sum of squared differences; a last kernel adds a pair's partial sums in
// order.  No atomics: results are reproducible.  The line integral is the reference's EXECUTED kernel operation by operation
// (sample positions, fused and unfused roundings, the compiler's four-sample blocks of its loop).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ecc_direct_geometry.cuh"
#include "ecc_geometry.cuh"
#include "ecc_internal.h"

namespace eccb200 {

namespace {

constexpr float kDirectStep = 0.4f;   // EpipolarConsistencyDirect.cu:73
constexpr int kLinesPerCta = 32;
constexpr int kMaxPlanesPerPair = 1 << 22;  // 2 x the image diagonal is the automatic count; a plane step of 1e-6 rad stays below this

// ---- the line integral ----------------------------------------------------------------------------------------------
// line l (Hessian normal form, fp32) through an n_u x n_v image; n_v_clip = the height the clipping uses (the reference's
// launcher passes n_u for it, EpipolarConsistencyDirect.cu:135).
struct ClippedLine {
    float o0, o1, d0, d1, t_min, t_max;
    bool valid;
};
__device__ __forceinline__ ClippedLine clip_line(const float* l, int n_u, int n_v_clip)
{
    ClippedLine L;
    L.o0 = __fmul_rn(-l[2], l[0]);
    L.o1 = __fmul_rn(-l[2], l[1]);
    L.d0 = l[1];
    L.d1 = -l[0];
    float ta = __fdiv_rn(__fadd_rn(1.f, -L.o0), L.d0), tb = __fdiv_rn(__fadd_rn((float)(n_u - 1), -L.o0), L.d0);
    float tc = __fdiv_rn(__fadd_rn(1.f, -L.o1), L.d1), td = __fdiv_rn(__fadd_rn((float)(n_v_clip - 1), -L.o1), L.d1);
    if ((double)__fmul_rn(L.d0, L.d0) < 1e-12) { tb = 1e10f; ta = -1e10f; }
    if ((double)__fmul_rn(L.d1, L.d1) < 1e-12) { td = 1e10f; tc = -1e10f; }
    // the middle two of the four sorted intersections (sort4, :16-27)
    const float lo1 = fminf(ta, tb), hi1 = fmaxf(ta, tb), lo2 = fminf(tc, td), hi2 = fmaxf(tc, td);
    L.t_min = fmaxf(lo1, lo2);
    L.t_max = fminf(hi1, hi2);
    if (fminf(hi1, hi2) < fmaxf(lo1, lo2)) { L.t_min = fminf(hi1, hi2); L.t_max = fmaxf(lo1, lo2); }
    const float pu = fmaf(L.t_min, L.d0, L.o0), pv = fmaf(L.t_min, L.d1, L.o1);
    L.valid = pu <= (float)n_u && pv <= (float)n_v_clip && pu >= 0.f && pv >= 0.f;
    L.o0 = __fadd_rn(L.o0, 0.5f);
    L.o1 = __fadd_rn(L.o1, 0.5f);
    return L;
}

// Derivative of the line integral: two parallel lines half a pixel either side (EpipolarConsistencyDirect.cu:98-117).  The
// source loop `for (t = t_min; t <= t_max; t += step)` is executed by the reference's build as blocks of four samples under
// one test, then two, then one (cuobjdump of its sm_100 object; the same shape nvcc gives RadonIntermediate.cu's loop, see
// ecc_radon_common.cuh: ref_last_sample) -- in fp32 a block's last sample can lie an ulp past t_max.
__device__ __forceinline__ float integrate_ecc(cudaTextureObject_t tex, const float* l, const ClippedLine& L)
{
    const float h0 = l[0], h1 = l[1];
    float t = L.t_min;
    const float t_max = L.t_max;
    if (t > t_max) return 0.f;
    float sump = 0.f, summ = 0.f;
    bool none_yet = true;
#define ECC_DIRECT_SAMPLE(tt)                                                         \
    {                                                                                 \
        const float u = fmaf(L.d0, (tt), L.o0), v = fmaf(L.d1, (tt), L.o1);            \
        const float p = tex2D<float>(tex, fmaf(h0, 0.5f, u), fmaf(h1, 0.5f, v));       \
        const float m = tex2D<float>(tex, fmaf(h0, -0.5f, u), fmaf(h1, -0.5f, v));     \
        sump = fmaf(p, kDirectStep, sump);                                            \
        summ = fmaf(m, kDirectStep, summ);                                            \
    }
    if (!(t + 1.2f > t_max)) {
        const float r3 = t_max - 1.2f;
#pragma unroll 1
        do {
            const float t1 = t + kDirectStep, t2 = t1 + kDirectStep, t3 = t2 + kDirectStep;
            ECC_DIRECT_SAMPLE(t) ECC_DIRECT_SAMPLE(t1) ECC_DIRECT_SAMPLE(t2) ECC_DIRECT_SAMPLE(t3)
            t = t3 + kDirectStep;
        } while (!(t > r3));
        none_yet = false;
    }
    const float t1 = t + kDirectStep;
    if (!(t1 > t_max)) {
        ECC_DIRECT_SAMPLE(t) ECC_DIRECT_SAMPLE(t1)
        t = t1 + kDirectStep;
        none_yet = false;
    }
    if (t <= t_max || none_yet) ECC_DIRECT_SAMPLE(t)
#undef ECC_DIRECT_SAMPLE
    return sump - summ;
}

// Fan-beam consistency: the integral weighted by the derivative of the rectifying perspectivity over the distance of the
// virtual pixel to the source (EpipolarConsistencyDirect.cu:82-96, RectifiedFBCC.h:66-79); roundings as the reference's
// build has them (a plain loop there).
__device__ __forceinline__ float integrate_fbcc(cudaTextureObject_t tex, const ClippedLine& L, const FbccInfo& f)
{
    const float det = fmaf(f.a, f.d, -__fmul_rn(f.b, f.c));
    const float cc = __fmul_rn(f.c, f.c), dd = __fmul_rn(f.d, f.d), cd2 = __fmul_rn(f.d, __fadd_rn(f.c, f.c));
    float sum = 0.f;
#pragma unroll 1
    for (float t = L.t_min; t <= L.t_max; t += kDirectStep) {
        const float u_prime = __fadd_rn(__fdiv_rn(fmaf(f.a, t, f.b), fmaf(f.c, t, f.d)), -f.t_prime_ak);
        const float den = __fadd_rn(dd, fmaf(cd2, t, __fmul_rn(__fmul_rn(cc, t), t)));
        const float w = __fdiv_rn(__fdiv_rn(det, den), __fsqrt_rn(fmaf(u_prime, u_prime, f.d_l_kappa_C_sq)));
        const float s = tex2D<float>(tex, fmaf(L.d0, t, L.o0), fmaf(L.d1, t, L.o1));
        sum = fmaf(__fmul_rn(s, kDirectStep), w, sum);
    }
    return sum;
}

// ---- the launcher-level seam: lines given, one thread per line (cuda_computeLineIntegrals, :122-142) -------------------
__global__ void direct_given_lines_kernel(int n_lines, const float* __restrict__ lines, int line_stride, const float* __restrict__ fbcc,
                                          int fbcc_stride, cudaTextureObject_t tex, int n_u, int n_v_clip, float* __restrict__ out)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_lines) return;
    const float l[3] = {lines[(size_t)idx * line_stride], lines[(size_t)idx * line_stride + 1], lines[(size_t)idx * line_stride + 2]};
    const ClippedLine L = clip_line(l, n_u, n_v_clip);
    float r = 0.f;
    if (L.valid) {
        if (fbcc) {
            FbccInfo f;
            const float* p = fbcc + (size_t)idx * fbcc_stride;
            f.a = p[0]; f.b = p[1]; f.c = p[2]; f.d = p[3]; f.t_prime_ak = p[4]; f.d_l_kappa_C_sq = p[5];
            r = integrate_fbcc(tex, L, f);
        } else {
            r = integrate_ecc(tex, l, L);
        }
    }
    out[idx] = r;
}

// ---- geometry kernels ---------------------------------------------------------------------------------------------
__global__ void direct_views_kernel(const double* __restrict__ Ps, int n, DirectView* __restrict__ views)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    DirectView V;
    direct_view(Ps + 12 * (size_t)k, V);
    views[k] = V;
}

// pair p of the enumeration i < j, i outer (MetricDirect::evaluate, EpipolarConsistencyDirect.cpp:236-247), or a listed pair
__global__ void direct_pairs_kernel(const DirectView* __restrict__ views, int n, const int* __restrict__ pair_ij, int n_pairs, double radius,
                                    double dkappa, int n_u, int n_v, int fbcc, int n_given, DirectPair* __restrict__ pairs,
                                    int* __restrict__ chunks, int* __restrict__ status)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int i = pair_ij[2 * p], j = pair_ij[2 * p + 1];
    DirectPair R;
    direct_pair(views[i], views[j], radius, dkappa, n_u, n_v, fbcc != 0, R);
    R.i = i;
    R.j = j;
    if (n_given > 0) R.n_lines = n_given;  // the caller's kappas (computeForImagePair with a filled `kappas`)
    if (R.n_lines < 0) R.n_lines = 0;
    if (R.n_lines > kMaxPlanesPerPair) {  // refused by the host (status word), never launched
        atomicMax(status, 1);
        R.n_lines = 0;
    }
    pairs[p] = R;
    chunks[p] = (R.n_lines + kLinesPerCta - 1) / kLinesPerCta;
}

// exclusive prefix sum of chunks[0..n) into offsets[0..n], one CTA
__global__ void direct_scan_kernel(const int* __restrict__ chunks, int n, int* __restrict__ offsets, int* __restrict__ status)
{
    __shared__ long long carry_s;
    __shared__ int warp_sums[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int k = base + tid;
        const int v = k < n ? chunks[k] : 0;
        int s = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += o;
        }
        if (lane == 31) warp_sums[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
            for (int off = 1; off < 32; off <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, w, off);
                if (lane >= off) w += o;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const long long before = carry_s + (warp ? warp_sums[warp - 1] : 0) + s - v;
        if (k < n) offsets[k] = (int)before;
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = before + v;
        __syncthreads();
    }
    if (tid == 0) {
        if (carry_s > 2147483647ll) atomicMax(status, 2);  // more CTAs than a grid can have
        offsets[n] = carry_s > 2147483647ll ? 0 : (int)carry_s;
    }
}

struct DirectLaunch {
    const cudaTextureObject_t* tex;
    const DirectView* views;
    const DirectPair* pairs;
    const int* offsets;  // [n_pairs + 1] first CTA of every pair
    int n_pairs;
    int n_u, n_v, n_v_clip;
    int fbcc;
    const float* kappas_in;  // single pair with given kappas, else null
    double* partials;        // one per CTA
    float *sig0, *sig1, *kappas_out;  // single pair: the redundant signals (nullable)
};

#ifndef ECC_DIRECT_MINBLOCKS
#define ECC_DIRECT_MINBLOCKS 16
#endif
__global__ void __launch_bounds__(2 * kLinesPerCta, ECC_DIRECT_MINBLOCKS)
direct_lines_kernel(const __grid_constant__ DirectLaunch p)
{
    __shared__ float v_s[2][kLinesPerCta];
    const int lane = threadIdx.x & 31, img = threadIdx.x >> 5;
    const int cta = blockIdx.x;
    // the pair this CTA belongs to: last p with offsets[p] <= cta
    int lo = 0, hi = p.n_pairs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&p.offsets[mid]) <= cta) lo = mid; else hi = mid - 1;
    }
    const DirectPair& R = p.pairs[lo];
    const int q = (cta - __ldg(&p.offsets[lo])) * kLinesPerCta + lane;
    const bool active = q < R.n_lines;
    float value = 0.f, kappa_f = 0.f;
    if (active) {
        const DirectView& V0 = p.views[R.i];
        const DirectView& V1 = p.views[R.j];
        kappa_f = p.kappas_in ? p.kappas_in[q] : (float)(R.kappa0 + R.dkappa * q);
        float l0[3], l1[3];
        direct_lines(V0, V1, R, (double)kappa_f, l0, l1);
        const float* l = img ? l1 : l0;
        const ClippedLine L = clip_line(l, p.n_u, p.n_v_clip);
        if (L.valid) {
            const cudaTextureObject_t tex = p.tex[img ? R.j : R.i];
            if (p.fbcc) {
                FbccInfo f0, f1;
                direct_fbcc(V0, V1, R, l0, l1, f0, f1);
                value = integrate_fbcc(tex, L, img ? f1 : f0);
            } else {
                value = integrate_ecc(tex, l, L);
            }
        }
    }
    v_s[img][lane] = value;
    __syncthreads();
    if (img == 0) {
        const float diff = __fadd_rn(v_s[0][lane], -v_s[1][lane]);
        double term = active ? (double)__fmul_rn(diff, diff) * R.dkappa : 0.0;
        for (int off = 16; off > 0; off >>= 1) term += __shfl_xor_sync(0xffffffffu, term, off);
        if (lane == 0) p.partials[cta] = term;
        if (active && p.sig0) {
            p.sig0[q] = v_s[0][lane];
            p.sig1[q] = v_s[1][lane];
            p.kappas_out[q] = kappa_f;
        }
    }
}

// a pair's value = its CTAs' partial sums added in order; the cost image entry i + j n (NRRD::ImageView::pixel(i, j))
__global__ void direct_reduce_kernel(const double* __restrict__ partials, const int* __restrict__ offsets, const DirectPair* __restrict__ pairs,
                                     int n_pairs, int n, double* __restrict__ vals, float* __restrict__ image)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double s = 0.0;
    for (int c = offsets[p]; c < offsets[p + 1]; c++) s += partials[c];
    vals[p] = s;
    if (image) image[pairs[p].i + (size_t)pairs[p].j * n] = (float)s;
}
__global__ void direct_total_kernel(const double* __restrict__ vals, int n_pairs, double* __restrict__ total)
{
    // fixed order: 256 strided partial sums, then a tree
    __shared__ double s[256];
    double a = 0.0;
    for (int k = threadIdx.x; k < n_pairs; k += 256) a += vals[k];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s[0];
}

int ensure_direct_images(ecc_context* ctx, int n_u, int n_v, int count)
{
    DirectState& D = ctx->direct;
    if (D.n_u == n_u && D.n_v == n_v && (int)D.arrays.size() == count) return ECC_OK;
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto t : D.tex_h) cudaDestroyTextureObject(t);
    for (auto a : D.arrays) cudaFreeArray(a);
    if (D.tex_d) cudaFree(D.tex_d);
    D.tex_h.clear();
    D.arrays.clear();
    D.tex_d = nullptr;
    D.n_u = n_u;
    D.n_v = n_v;
    D.n_images = 0;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    for (int k = 0; k < count; k++) {
        cudaArray_t arr = nullptr;
        ECC_CUDA(ctx, cudaMallocArray(&arr, &desc, n_u, n_v));
        D.arrays.push_back(arr);
        cudaResourceDesc res = {};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr;
        cudaTextureDesc td = {};  // BindlessTexture2D<float>(w, h, buffer): pixel coordinates, linear, clamp (CudaBindlessTexture.cpp:17-45)
        td.normalizedCoords = 0;
        td.filterMode = cudaFilterModeLinear;
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        ECC_CUDA(ctx, cudaCreateTextureObject(&tex, &res, &td, nullptr));
        D.tex_h.push_back(tex);
    }
    ECC_CUDA(ctx, cudaMalloc(&D.tex_d, sizeof(cudaTextureObject_t) * (count ? count : 1)));
    ECC_CUDA(ctx, cudaMemcpyAsync(D.tex_d, D.tex_h.data(), sizeof(cudaTextureObject_t) * count, cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

// geometry of the listed pairs (host list of 2 * n_pairs view indices) -> D.pairs_d, D.offsets_d; returns the CTA count
int direct_prepare(ecc_context* ctx, const int* pair_ij_h, int n_pairs, int n_given, int* total_ctas)
{
    DirectState& D = ctx->direct;
    const int n = ctx->n_views;
    double radius = ctx->object_radius;
    if (!(radius > 0)) radius = object_radius_from_view(ctx->Ps_h.data(), D.n_u, D.n_v);  // Metric::getObjectRadius: the first matrix
    int rc = ensure_bytes(ctx, (void**)&D.views_d, &D.views_bytes, sizeof(DirectView) * n);
    if (rc) return rc;
    rc = ensure_bytes(ctx, (void**)&D.Ps_d, &D.Ps_bytes, sizeof(double) * 12 * n);
    if (rc) return rc;
    rc = ensure_bytes(ctx, (void**)&D.pairs_d, &D.pairs_bytes, sizeof(DirectPair) * n_pairs);
    if (rc) return rc;
    rc = ensure_bytes(ctx, (void**)&D.ij_d, &D.ij_bytes, sizeof(int) * 2 * n_pairs);
    if (rc) return rc;
    rc = ensure_bytes(ctx, (void**)&D.offsets_d, &D.offsets_bytes, sizeof(int) * (2 * (size_t)n_pairs + 3));
    if (rc) return rc;
    rc = ensure_bytes(ctx, (void**)&D.vals_d, &D.vals_bytes, sizeof(double) * ((size_t)n_pairs + 1));
    if (rc) return rc;
    // layout: [n_pairs] offsets, total, status word, [n_pairs] CTAs per pair
    int* status_d = D.offsets_d + n_pairs + 1;
    int* chunks_d = D.offsets_d + n_pairs + 2;
    ECC_CUDA(ctx, cudaMemsetAsync(status_d, 0, sizeof(int), ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(D.Ps_d, ctx->Ps_h.data(), sizeof(double) * 12 * n, cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(D.ij_d, pair_ij_h, sizeof(int) * 2 * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    const int s = prof_begin(ctx, FAM_GEOMETRY);
    direct_views_kernel<<<(n + 63) / 64, 64, 0, ctx->stream>>>(D.Ps_d, n, (DirectView*)D.views_d);
    direct_pairs_kernel<<<(n_pairs + 63) / 64, 64, 0, ctx->stream>>>((const DirectView*)D.views_d, n, D.ij_d, n_pairs, radius, ctx->dkappa, D.n_u,
                                                                    D.n_v, D.fbcc, n_given, (DirectPair*)D.pairs_d, chunks_d, status_d);
    direct_scan_kernel<<<1, 1024, 0, ctx->stream>>>(chunks_d, n_pairs, D.offsets_d, status_d);
    prof_end(ctx, s);
    ECC_CUDA(ctx, cudaGetLastError());
    int total_and_status[2] = {0, 0};
    ECC_CUDA(ctx, cudaMemcpyAsync(total_and_status, D.offsets_d + n_pairs, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pair_ij_h and the CTA count
    *total_ctas = total_and_status[0];
    if (total_and_status[1] == 1)
        return fail(ctx, ECC_ERR_UNSUPPORTED, "direct metric: the epipolar plane step gives a pair more than 4194304 planes");
    if (total_and_status[1] == 2) return fail(ctx, ECC_ERR_UNSUPPORTED, "direct metric: more than 2^31 CTAs in one evaluation (plane step too fine for this many pairs)");
    return ECC_OK;
}

int direct_launch_lines(ecc_context* ctx, int n_pairs, int total_ctas, const float* kappas_in_d, float* sig0_d, float* sig1_d, float* kappas_out_d,
                        float* image_d, double* total_d)
{
    DirectState& D = ctx->direct;
    int rc = ensure_bytes(ctx, (void**)&D.partials_d, &D.partials_bytes, sizeof(double) * ((size_t)total_ctas + 1));
    if (rc) return rc;
    DirectLaunch L;
    L.tex = D.tex_d;
    L.views = (const DirectView*)D.views_d;
    L.pairs = (const DirectPair*)D.pairs_d;
    L.offsets = D.offsets_d;
    L.n_pairs = n_pairs;
    L.n_u = D.n_u;
    L.n_v = D.n_v;
    L.n_v_clip = D.reference_clip ? D.n_u : D.n_v;
    L.fbcc = D.fbcc;
    L.kappas_in = kappas_in_d;
    L.partials = D.partials_d;
    L.sig0 = sig0_d;
    L.sig1 = sig1_d;
    L.kappas_out = kappas_out_d;
    if (total_ctas > 0) {
        const int s = prof_begin(ctx, FAM_PAIRS);
        direct_lines_kernel<<<total_ctas, 2 * kLinesPerCta, 0, ctx->stream>>>(L);
        prof_end(ctx, s);
    }
    const int s2 = prof_begin(ctx, FAM_REDUCE);
    direct_reduce_kernel<<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(D.partials_d, D.offsets_d, (const DirectPair*)D.pairs_d, n_pairs,
                                                                        ctx->n_views, D.vals_d, image_d);
    direct_total_kernel<<<1, 256, 0, ctx->stream>>>(D.vals_d, n_pairs, total_d);
    prof_end(ctx, s2);
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

struct Guard {
    explicit Guard(ecc_context* c) { cudaSetDevice(c->device); }
};

}  // namespace

void free_direct(ecc_context* ctx)
{
    DirectState& D = ctx->direct;
    for (auto t : D.tex_h) cudaDestroyTextureObject(t);
    for (auto a : D.arrays) cudaFreeArray(a);
    void* bufs[] = {D.tex_d, D.views_d, D.Ps_d, D.pairs_d, D.ij_d, D.offsets_d, D.vals_d, D.partials_d, D.scratch_d};
    for (void* b : bufs)
        if (b) cudaFree(b);
    D = DirectState();
}

}  // namespace eccb200

using namespace eccb200;

// ---- the reference's launcher symbol with ITS OWN signature (C++ linkage) -------------------------------------------------
// EpipolarConsistencyDirect.cpp:10-16 declares it `extern`, EpipolarConsistencyDirect.cu:122-142 defines it: linking the
// reference's unmodified EpipolarConsistencyDirect.cpp against libecc_b200.so instead of its own .cu resolves the symbol here
// (INTEGRATION.md option B).  Same arguments, same buffer written, legacy default stream, returns when the work has finished.
// The reference's definition hands its kernel n_u for both image sizes (:135), i.e. clips against n_u x n_u: so does this
// symbol -- it exists to give the reference's bits (checked: bit-identical integrals) -- unless ECC_COMPAT_DIRECT_CLIP=image.
void cuda_computeLineIntegrals(short n_lines, float* lines_d, short line_stride, float* fbcc_d, short fbcc_stride, cudaTextureObject_t I,
                               short n_u, short n_v, float* integrals_out_d)
{
    if (n_lines <= 0) return;
    static const bool clip_image = [] {
        const char* v = getenv("ECC_COMPAT_DIRECT_CLIP");
        return v && v[0] == 'i';
    }();
    direct_given_lines_kernel<<<(n_lines + 31) / 32, 32, 0, cudaStreamLegacy>>>(n_lines, lines_d, line_stride, fbcc_d, fbcc_stride, I, n_u,
                                                                             clip_image ? n_v : n_u, integrals_out_d);
    const cudaError_t e1 = cudaGetLastError(), e2 = cudaStreamSynchronize(cudaStreamLegacy);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {  // the reference's convention: print and exit (UtilsCuda.hxx:14-28)
        std::fprintf(stderr, "libecc_b200 (cuda_computeLineIntegrals): %s\n", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        std::exit(1);
    }
}

extern "C" {

int ecc_direct_set_images(ecc_context* ctx, const float* images, int n, int n_u, int n_v)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (!images || n < 0 || n_u < 2 || n_v < 2) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_set_images: bad argument");
    int rc = ensure_direct_images(ctx, n_u, n_v, n);
    if (rc) return rc;
    DirectState& D = ctx->direct;
    const size_t px = (size_t)n_u * n_v;
    const cudaMemcpyKind kind = is_device_pointer(images) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    for (int k = 0; k < n; k++)
        ECC_CUDA(ctx, cudaMemcpy2DToArrayAsync(D.arrays[k], 0, 0, images + px * k, sizeof(float) * n_u, sizeof(float) * n_u, n_v, kind, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffer is free again (host images are not retained)
    D.n_images = n;
    // Metric::n_u / n_v (MetricDirect::setProjectionImages, EpipolarConsistencyDirect.cpp:226-234): the size getObjectRadius
    // estimates the radius for, unless Radon intermediates have set it
    if (ctx->n_dtrs == 0) { ctx->n_u = n_u; ctx->n_v = n_v; }
    return ECC_OK;
}

int ecc_direct_set_image_pointers(ecc_context* ctx, const float* const* images, int n, int n_u, int n_v)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    if (n < 0 || (n > 0 && !images) || n_u < 2 || n_v < 2) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_set_image_pointers: bad argument");
    for (int k = 0; k < n; k++)
        if (!images[k]) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_set_image_pointers: null image");
    int rc = ensure_direct_images(ctx, n_u, n_v, n);
    if (rc) return rc;
    DirectState& D = ctx->direct;
    for (int k = 0; k < n; k++)
        ECC_CUDA(ctx, cudaMemcpy2DToArrayAsync(D.arrays[k], 0, 0, images[k], sizeof(float) * n_u, sizeof(float) * n_u, n_v, cudaMemcpyDefault, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    D.n_images = n;
    if (ctx->n_dtrs == 0) { ctx->n_u = n_u; ctx->n_v = n_v; }
    return ECC_OK;
}

int ecc_direct_set_fan_beam(ecc_context* ctx, int fbcc)
{
    if (!ctx) return ECC_ERR_INVALID;
    ctx->direct.fbcc = fbcc ? 1 : 0;
    return ECC_OK;
}

int ecc_direct_set_reference_clip(ecc_context* ctx, int on)
{
    if (!ctx) return ECC_ERR_INVALID;
    ctx->direct.reference_clip = on ? 1 : 0;
    return ECC_OK;
}

// pairs [begin, end) of the enumeration i < j, i outer (MetricDirect::evaluate's loop order): 2 ints per pair
static const int* direct_pair_list(ecc_context* ctx, int n)
{
    DirectState& D = ctx->direct;
    const size_t n_pairs = (size_t)n * (n - 1) / 2;
    if (D.all_pairs_h.size() != 2 * n_pairs || D.all_pairs_n != n) {
        D.all_pairs_h.resize(2 * n_pairs);
        size_t k = 0;
        for (int i = 0; i < n; i++)
            for (int j = i + 1; j < n; j++) { D.all_pairs_h[k++] = i; D.all_pairs_h[k++] = j; }
        D.all_pairs_n = n;
    }
    return D.all_pairs_h.data();
}

static int direct_check_state(ecc_context* ctx, const char* who, long long* n_pairs)
{
    DirectState& D = ctx->direct;
    const int n = ctx->n_views;
    if (n <= 0) return fail(ctx, ECC_ERR_STATE, std::string(who) + ": projection matrices not set");
    if (D.n_images != n) return fail(ctx, ECC_ERR_STATE, std::string(who) + ": number of images and of projection matrices differ");
    *n_pairs = (long long)n * (n - 1) / 2;
    if (*n_pairs > (1ll << 28)) return fail(ctx, ECC_ERR_UNSUPPORTED, std::string(who) + ": too many pairs");
    return ECC_OK;
}

int ecc_direct_evaluate_range(ecc_context* ctx, long long pair_begin, long long pair_end, float* cost_image, double* sum)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    DirectState& D = ctx->direct;
    long long all = 0;
    int rc = direct_check_state(ctx, "ecc_direct_evaluate_range", &all);
    if (rc) return rc;
    if (pair_begin < 0 || pair_end < pair_begin || pair_end > all) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_evaluate_range: bad pair range");
    const int n = ctx->n_views, n_pairs = (int)(pair_end - pair_begin);
    if (n_pairs == 0) {
        if (sum) *sum = 0.0;
        return ECC_OK;
    }
    int total_ctas = 0;
    rc = direct_prepare(ctx, direct_pair_list(ctx, n) + 2 * pair_begin, n_pairs, 0, &total_ctas);
    if (rc) return rc;
    float* image_d = nullptr;
    const bool image_on_host = cost_image && !is_device_pointer(cost_image);
    if (cost_image) {
        if (image_on_host) {
            rc = ensure_bytes(ctx, (void**)&ctx->cost_d, &ctx->cost_cap, sizeof(float) * (size_t)n * n);
            if (rc) return rc;
            ECC_CUDA(ctx, cudaMemcpyAsync(ctx->cost_d, cost_image, sizeof(float) * (size_t)n * n, cudaMemcpyHostToDevice, ctx->stream));
            image_d = ctx->cost_d;
        } else {
            image_d = cost_image;
        }
    }
    rc = direct_launch_lines(ctx, n_pairs, total_ctas, nullptr, nullptr, nullptr, nullptr, image_d, D.vals_d + n_pairs);
    if (rc) return rc;
    double total = 0.0;
    ECC_CUDA(ctx, cudaMemcpyAsync(&total, D.vals_d + n_pairs, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (image_on_host) ECC_CUDA(ctx, cudaMemcpyAsync(cost_image, image_d, sizeof(float) * (size_t)n * n, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (sum) *sum = total;
    return ECC_OK;
}

int ecc_direct_evaluate(ecc_context* ctx, float* cost_image, double* sum)
{
    if (!ctx) return ECC_ERR_INVALID;
    long long all = 0;
    const int rc = direct_check_state(ctx, "ecc_direct_evaluate", &all);
    if (rc) return rc;
    return ecc_direct_evaluate_range(ctx, 0, all, cost_image, sum);
}

int ecc_direct_partition(ecc_context* ctx, int n_parts, long long* bounds)
{
    if (!ctx || !bounds || n_parts < 1) return ECC_ERR_INVALID;
    Guard g(ctx);
    DirectState& D = ctx->direct;
    long long all = 0;
    int rc = direct_check_state(ctx, "ecc_direct_partition", &all);
    if (rc) return rc;
    bounds[0] = 0;
    for (int k = 1; k <= n_parts; k++) bounds[k] = all;
    if (all == 0) return ECC_OK;
    int total_ctas = 0;
    rc = direct_prepare(ctx, direct_pair_list(ctx, ctx->n_views), (int)all, 0, &total_ctas);
    if (rc) return rc;
    // the CTA offset table IS the work prefix (a CTA = 32 planes through both images): cut it into equal parts
    std::vector<int> offsets((size_t)all + 1);
    ECC_CUDA(ctx, cudaMemcpyAsync(offsets.data(), D.offsets_d, sizeof(int) * ((size_t)all + 1), cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    long long p = 0;
    for (int k = 1; k < n_parts; k++) {
        const double target = (double)total_ctas * k / n_parts;
        while (p < all && (double)offsets[p + 1] <= target) p++;
        if (p < all && target - offsets[p] > offsets[p + 1] - target) p++;  // the nearer pair boundary
        bounds[k] = p;
    }
    return ECC_OK;
}

int ecc_direct_evaluate_pair(ecc_context* ctx, int i, int j, int n_given, int capacity, float* kappas, float* samples0, float* samples1,
                             int* n_lines, double* value)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    DirectState& D = ctx->direct;
    const int n = ctx->n_views;
    if (n <= 0) return fail(ctx, ECC_ERR_STATE, "ecc_direct_evaluate_pair: projection matrices not set");
    if (D.n_images != n) return fail(ctx, ECC_ERR_STATE, "ecc_direct_evaluate_pair: number of images and of projection matrices differ");
    if (i < 0 || j < 0 || i >= n || j >= n) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_evaluate_pair: view index out of range");
    if (n_given < 0 || (n_given > 0 && (!kappas || capacity < n_given))) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_evaluate_pair: bad kappas");
    const int ij[2] = {i, j};
    int total_ctas = 0;
    int rc = direct_prepare(ctx, ij, 1, n_given, &total_ctas);
    if (rc) return rc;
    const size_t lines = (size_t)total_ctas * kLinesPerCta;
    rc = ensure_bytes(ctx, (void**)&D.scratch_d, &D.scratch_bytes, sizeof(float) * 4 * (lines + 1));
    if (rc) return rc;
    float *sig0_d = D.scratch_d, *sig1_d = sig0_d + lines, *kap_d = sig1_d + lines, *kin_d = kap_d + lines;
    if (n_given > 0) ECC_CUDA(ctx, cudaMemcpyAsync(kin_d, kappas, sizeof(float) * n_given, cudaMemcpyHostToDevice, ctx->stream));
    rc = direct_launch_lines(ctx, 1, total_ctas, n_given > 0 ? kin_d : nullptr, sig0_d, sig1_d, kap_d, nullptr, D.vals_d + 1);
    if (rc) return rc;
    DirectPair R;
    double total = 0.0;
    ECC_CUDA(ctx, cudaMemcpyAsync(&R, D.pairs_d, sizeof(DirectPair), cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(&total, D.vals_d + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int have = R.n_lines, take = have < capacity ? have : capacity;
    if (take > 0) {
        if (kappas && n_given == 0) ECC_CUDA(ctx, cudaMemcpyAsync(kappas, kap_d, sizeof(float) * take, cudaMemcpyDeviceToHost, ctx->stream));
        if (samples0) ECC_CUDA(ctx, cudaMemcpyAsync(samples0, sig0_d, sizeof(float) * take, cudaMemcpyDeviceToHost, ctx->stream));
        if (samples1) ECC_CUDA(ctx, cudaMemcpyAsync(samples1, sig1_d, sizeof(float) * take, cudaMemcpyDeviceToHost, ctx->stream));
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (n_lines) *n_lines = have;
    if (value) *value = total;
    return ECC_OK;
}

int ecc_direct_pair_geometry(ecc_context* ctx, int i, int j, int capacity, float* kappas, float* lines0, float* lines1, float* fbcc0, float* fbcc1,
                             int* n_lines, double* dkappa)
{
    if (!ctx) return ECC_ERR_INVALID;
    DirectState& D = ctx->direct;
    const int n = ctx->n_views;
    if (n <= 0) return fail(ctx, ECC_ERR_STATE, "ecc_direct_pair_geometry: projection matrices not set");
    if (i < 0 || j < 0 || i >= n || j >= n) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_pair_geometry: view index out of range");
    if (D.n_u < 2 || D.n_v < 2) return fail(ctx, ECC_ERR_STATE, "ecc_direct_pair_geometry: images not set");
    double radius = ctx->object_radius;
    if (!(radius > 0)) radius = object_radius_from_view(ctx->Ps_h.data(), D.n_u, D.n_v);
    DirectView V0, V1;
    direct_view(ctx->Ps_h.data() + 12 * (size_t)i, V0);
    direct_view(ctx->Ps_h.data() + 12 * (size_t)j, V1);
    DirectPair R;
    direct_pair(V0, V1, radius, ctx->dkappa, D.n_u, D.n_v, true, R);
    if (n_lines) *n_lines = R.n_lines;
    if (dkappa) *dkappa = R.dkappa;
    for (int q = 0; q < R.n_lines && q < capacity; q++) {
        const float kf = (float)(R.kappa0 + R.dkappa * q);
        float l0[3], l1[3];
        direct_lines(V0, V1, R, (double)kf, l0, l1);
        if (kappas) kappas[q] = kf;
        if (lines0) memcpy(lines0 + 3 * (size_t)q, l0, sizeof(l0));
        if (lines1) memcpy(lines1 + 3 * (size_t)q, l1, sizeof(l1));
        if (fbcc0 || fbcc1) {
            FbccInfo f0, f1;
            direct_fbcc(V0, V1, R, l0, l1, f0, f1);
            if (fbcc0) memcpy(fbcc0 + 8 * (size_t)q, &f0, sizeof(f0));
            if (fbcc1) memcpy(fbcc1 + 8 * (size_t)q, &f1, sizeof(f1));
        }
    }
    return ECC_OK;
}

int ecc_direct_line_integrals(ecc_context* ctx, int image, const float* lines, int n_lines, int line_stride, const float* fbcc, int fbcc_stride,
                              float* out)
{
    if (!ctx) return ECC_ERR_INVALID;
    Guard g(ctx);
    DirectState& D = ctx->direct;
    if (image < 0 || image >= D.n_images) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_line_integrals: image index out of range");
    if (!lines || !out || n_lines < 0 || line_stride < 3 || (fbcc && fbcc_stride < 6)) return fail(ctx, ECC_ERR_INVALID, "ecc_direct_line_integrals: bad argument");
    if (n_lines == 0) return ECC_OK;
    const size_t lf = (size_t)n_lines * line_stride, ff = fbcc ? (size_t)n_lines * fbcc_stride : 0;
    int rc = ensure_bytes(ctx, (void**)&D.scratch_d, &D.scratch_bytes, sizeof(float) * (lf + ff + n_lines));
    if (rc) return rc;
    float *lines_d = D.scratch_d, *fbcc_d = lines_d + lf, *out_d = fbcc_d + ff;
    ECC_CUDA(ctx, cudaMemcpyAsync(lines_d, lines, sizeof(float) * lf, cudaMemcpyDefault, ctx->stream));
    if (fbcc) ECC_CUDA(ctx, cudaMemcpyAsync(fbcc_d, fbcc, sizeof(float) * ff, cudaMemcpyDefault, ctx->stream));
    const int s = prof_begin(ctx, FAM_PAIRS);
    direct_given_lines_kernel<<<(n_lines + 31) / 32, 32, 0, ctx->stream>>>(n_lines, lines_d, line_stride, fbcc ? fbcc_d : nullptr, fbcc_stride,
                                                                          D.tex_h[image], D.n_u, D.reference_clip ? D.n_u : D.n_v, out_d);
    prof_end(ctx, s);
    ECC_CUDA(ctx, cudaGetLastError());
    ECC_CUDA(ctx, cudaMemcpyAsync(out, out_d, sizeof(float) * n_lines, cudaMemcpyDefault, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ECC_OK;
}

}  // extern "C"
