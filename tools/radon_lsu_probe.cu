// radon_lsu_probe.cu -- go/no-go experiment for a second, shared-memory sampling path in the Radon kernel.
//
// The production Radon kernel is bound by the texture data pipe (profiles/ncu_radon_r01.txt: 98 % of the
// L1TEX tex-wavefront peak, 14 % of the issue slots).  This probe measures, on an image small enough to sit in
// shared memory, (1) whether the texture unit's bilinear filter can be reproduced from shared memory with the
// measured weight model (profiles/tex_probe_r01.txt), bin for bin, and (2) the sample rate of the texture path,
// of the shared-memory path, and of both running side by side in the same CTAs (warp roles).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/radon_lsu_probe tools/radon_lsu_probe.cu
// Run on the GPU box: ./tools/radon_lsu_probe > gpurun_out/radon_lsu_probe.txt
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr float kStep = 0.66f;
constexpr float kPi = 3.14159265359f;

struct BinLine {
    float o0, o1, d0, d1, t, t_max;
    bool valid, swapped;
};

__device__ __forceinline__ BinLine bin_line(int ix, int iy, int n_alpha, int n_t, float n_u, float n_v)
{
    BinLine L;
    const float x_rel = ix / (float)n_alpha - 0.5f;
    const float y_rel = iy / (float)n_t - 0.5f;
    const float diag = sqrtf(n_u * n_u + n_v * n_v);
    const float alpha = x_rel * kPi;
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
    const float lo1 = fminf(ta, tb), hi1 = fmaxf(ta, tb);
    const float lo2 = fminf(tc, td), hi2 = fmaxf(tc, td);
    L.t = fmaxf(lo1, lo2);
    L.t_max = fminf(hi1, hi2);
    L.swapped = fminf(hi1, hi2) < fmaxf(lo1, lo2);  // the line misses the inset box: samples may need the clamp
    if (L.swapped) { L.t = fminf(hi1, hi2); L.t_max = fmaxf(lo1, lo2); }
    const float pu = L.o0 + L.t * L.d0, pv = L.o1 + L.t * L.d1;
    const bool inside = (pu <= n_u && pv <= n_v && pu >= 0.f && pv >= 0.f);
    L.valid = inside && !(L.t_max <= L.t);
    return L;
}

// Bilinear sample from a shared-memory copy of the image (row stride S floats, one replicated column/row at the far
// edges), reproducing the texture unit: position - 1/2 rounded to 1/256 (half up), 2-D weights
// w11 = (a*b + 128) >> 8, w10 = a - w11, w01 = b - w11, w00 = 256 - a - b + w11 (all /256).
__device__ __forceinline__ float sample_smem(const float* __restrict__ img, int S, float x, float y)
{
    const float Xs = fmaf(x, 256.f, -127.5f);          // exact: 256*(x - 1/2) + 1/2
    const float Ys = fmaf(y, 256.f, -127.5f);
    const unsigned Xi = __float_as_uint(__fadd_rd(Xs, 8388608.f));  // 0x4B000000 + floor(Xs)
    const unsigned Yi = __float_as_uint(__fadd_rd(Ys, 8388608.f));
    const unsigned a = Xi & 255u, b = Yi & 255u;
    const unsigned ix = (Xi >> 8) & 0x7FFFu, iy = (Yi >> 8) & 0x7FFFu;
    const float* p = img + iy * S + ix;
    const float v00 = p[0], v10 = p[1], v01 = p[S], v11 = p[S + 1];
    const unsigned w11 = (a * b + 128u) >> 8;
    const unsigned w10 = a - w11, w01 = b - w11, w00 = 256u + w11 - a - b;
    float s = __uint2float_rn(w11) * v11;
    s = fmaf(__uint2float_rn(w01), v01, s);
    s = fmaf(__uint2float_rn(w10), v10, s);
    s = fmaf(__uint2float_rn(w00), v00, s);
    return s * 0.00390625f;
}

// Same with the image stored as float2 {v[x], v[x+1]} per texel: two 8-byte loads per sample.
__device__ __forceinline__ float sample_smem2(const float2* __restrict__ img, int S, float x, float y)
{
    const float Xs = fmaf(x, 256.f, -127.5f);
    const float Ys = fmaf(y, 256.f, -127.5f);
    const unsigned Xi = __float_as_uint(__fadd_rd(Xs, 8388608.f));
    const unsigned Yi = __float_as_uint(__fadd_rd(Ys, 8388608.f));
    const unsigned a = Xi & 255u, b = Yi & 255u;
    const unsigned ix = (Xi >> 8) & 0x7FFFu, iy = (Yi >> 8) & 0x7FFFu;
    const float2* p = img + iy * S + ix;
    const float2 r0 = p[0], r1 = p[S];
    const unsigned w11 = (a * b + 128u) >> 8;
    const unsigned w10 = a - w11, w01 = b - w11, w00 = 256u + w11 - a - b;
    float s = __uint2float_rn(w11) * r1.y;
    s = fmaf(__uint2float_rn(w01), r1.x, s);
    s = fmaf(__uint2float_rn(w10), r0.y, s);
    s = fmaf(__uint2float_rn(w00), r0.x, s);
    return s * 0.00390625f;
}

enum { ROLE_TEX = 0, ROLE_LDS = 1, ROLE_LDS2 = 2 };

// One CTA per image (grid-stride over images); the image sits in shared memory; warps draw 32-bin tiles (2 angles x
// 16 t bins, the production lane tiling) from a CTA-wide counter.  Warps with (warp % 8) < tex_of_8 sample through the
// texture unit, the others from shared memory (layout LAYOUT: 1 plain floats, 2 float2 pairs).
template <int LAYOUT>
__global__ void __launch_bounds__(1024)
probe_kernel(const cudaTextureObject_t* __restrict__ texs, const float* __restrict__ images, int n_img, int n_u, int n_v,
             int n_alpha, int n_t, int S, int tex_of_8, float* __restrict__ out)
{
    extern __shared__ float smem[];
    __shared__ int counter;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int quad = lane >> 2, ql = lane & 3;
    const int da = ql & 1, dt = quad * 2 + (ql >> 1);
    const int tiles_a = n_alpha / 2, tiles = tiles_a * (n_t / 16);
    const bool use_tex = (warp & 7) < tex_of_8;
    for (int img = blockIdx.x; img < n_img; img += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) counter = 0;
        const float* src = images + (size_t)img * n_u * n_v;
        if (tex_of_8 < 8) {
            // rows 0..n_v (row n_v replicates n_v-1), columns 0..n_u (column n_u replicates n_u-1)
            for (int k = threadIdx.x; k < (n_v + 1) * (n_u + 1); k += blockDim.x) {
                const int r = k / (n_u + 1), c = k - r * (n_u + 1);
                const int rr = r < n_v ? r : n_v - 1, cc = c < n_u ? c : n_u - 1;
                const float v = src[rr * n_u + cc];
                if (LAYOUT == 1) smem[r * S + c] = v;
                else {
                    const int c1 = cc + 1 < n_u ? cc + 1 : n_u - 1;
                    reinterpret_cast<float2*>(smem)[r * S + c] = make_float2(v, src[rr * n_u + c1]);
                }
            }
        }
        __syncthreads();
        const cudaTextureObject_t tex = texs[img];
        float* dst_img = out + (size_t)img * n_alpha * n_t;
        for (;;) {
            int tile = 0;
            if (lane == 0) tile = atomicAdd(&counter, 1);
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile >= tiles) break;
            const int ix = (tile % tiles_a) * 2 + da, iy = (tile / tiles_a) * 16 + dt;
            float* dst = dst_img + (size_t)iy * n_alpha + ix;
            BinLine L = bin_line(ix, iy, n_alpha, n_t, (float)n_u, (float)n_v);
            // lines that miss the inset box can leave the image: those tiles go through the texture unit (clamp)
            const bool tile_tex = use_tex || __any_sync(0xffffffffu, L.valid && L.swapped);
            if (!L.valid) { *dst = 0.f; continue; }
            float o0 = L.o0 + 0.5f, o1 = L.o1 + 0.5f;
            const float d0 = L.d0, d1 = L.d1, t_max = L.t_max;
            o0 -= 0.5f * d1;
            o1 += 0.5f * d0;
            float sum = 0.f, sumo = 0.f;
            if (tile_tex) {
                for (float t = L.t; t <= t_max; t += kStep) {
                    sum += tex2D<float>(tex, o0 + t * d0, o1 + t * d1);
                    sumo += tex2D<float>(tex, o0 + t * d0 + d1, o1 + t * d1 - d0);
                }
            } else if (LAYOUT == 1) {
                for (float t = L.t; t <= t_max; t += kStep) {
                    sum += sample_smem(smem, S, o0 + t * d0, o1 + t * d1);
                    sumo += sample_smem(smem, S, o0 + t * d0 + d1, o1 + t * d1 - d0);
                }
            } else {
                const float2* s2 = reinterpret_cast<const float2*>(smem);
                for (float t = L.t; t <= t_max; t += kStep) {
                    sum += sample_smem2(s2, S, o0 + t * d0, o1 + t * d1);
                    sumo += sample_smem2(s2, S, o0 + t * d0 + d1, o1 + t * d1 - d0);
                }
            }
            *dst = (sum - sumo) * kStep;
        }
    }
}

__global__ void count_kernel(int n_u, int n_v, int n_alpha, int n_t, unsigned long long* total)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y;
    unsigned long long cnt = 0;
    if (ix < n_alpha && iy < n_t) {
        BinLine L = bin_line(ix, iy, n_alpha, n_t, (float)n_u, (float)n_v);
        if (L.valid)
            for (float t = L.t; t <= L.t_max; t += kStep) cnt += 2;
    }
    if (cnt) atomicAdd(total, cnt);
}

int main(int argc, char** argv)
{
    const int n_u = argc > 1 ? atoi(argv[1]) : 160, n_v = argc > 2 ? atoi(argv[2]) : 128;
    const int n_alpha = 256, n_t = 96, n_img = 592;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device: %s, SMs %d, max clock %d MHz; image %dx%d, bins %dx%d, %d images\n", prop.name, prop.multiProcessorCount,
           clk_khz / 1000, n_u, n_v, n_alpha, n_t, n_img);

    // smooth + rough synthetic images
    std::vector<float> h((size_t)n_img * n_u * n_v);
    unsigned rng = 12345u;
    for (int k = 0; k < n_img; k++)
        for (int y = 0; y < n_v; y++)
            for (int x = 0; x < n_u; x++) {
                rng = rng * 1664525u + 1013904223u;
                const float noise = (rng >> 8) * (1.f / 16777216.f);
                const float dx = (x - 0.5f * n_u - 7.f * sinf(k)) / (0.33f * n_u), dy = (y - 0.5f * n_v) / (0.3f * n_v);
                const float r2 = dx * dx + dy * dy;
                h[((size_t)k * n_v + y) * n_u + x] = (r2 < 1.f ? 100.f * sqrtf(1.f - r2) : 0.f) + 3.f * noise;
            }
    float *img_d, *out_a, *out_b;
    CK(cudaMalloc(&img_d, h.size() * 4));
    CK(cudaMemcpy(img_d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const size_t out_n = (size_t)n_img * n_alpha * n_t;
    CK(cudaMalloc(&out_a, out_n * 4));
    CK(cudaMalloc(&out_b, out_n * 4));
    std::vector<cudaTextureObject_t> texs(n_img);
    for (int k = 0; k < n_img; k++) {
        cudaResourceDesc res = {};
        res.resType = cudaResourceTypePitch2D;
        res.res.pitch2D.devPtr = img_d + (size_t)k * n_u * n_v;
        res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
        res.res.pitch2D.width = n_u;
        res.res.pitch2D.height = n_v;
        res.res.pitch2D.pitchInBytes = (size_t)n_u * 4;
        cudaTextureDesc td = {};
        td.normalizedCoords = 0;
        td.filterMode = cudaFilterModeLinear;
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&texs[k], &res, &td, nullptr));
    }
    cudaTextureObject_t* texs_d;
    CK(cudaMalloc(&texs_d, sizeof(cudaTextureObject_t) * n_img));
    CK(cudaMemcpy(texs_d, texs.data(), sizeof(cudaTextureObject_t) * n_img, cudaMemcpyHostToDevice));

    unsigned long long* total_d;
    CK(cudaMalloc(&total_d, 8));
    CK(cudaMemset(total_d, 0, 8));
    count_kernel<<<dim3((n_alpha + 63) / 64, n_t), 64>>>(n_u, n_v, n_alpha, n_t, total_d);
    unsigned long long per_img = 0;
    CK(cudaMemcpy(&per_img, total_d, 8, cudaMemcpyDeviceToHost));
    const double samples = (double)per_img * n_img;
    printf("bilinear samples per image %llu, total %.3e\n", per_img, samples);

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount;
    auto run = [&](int layout, int S, int tex_of_8, float* out, const char* name) {
        const size_t smem = (size_t)(n_v + 1) * S * 4 * layout;
        if (layout == 1) CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else CK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            if (layout == 1) probe_kernel<1><<<grid, 1024, smem>>>(texs_d, img_d, n_img, n_u, n_v, n_alpha, n_t, S, tex_of_8, out);
            else probe_kernel<2><<<grid, 1024, smem>>>(texs_d, img_d, n_img, n_u, n_v, n_alpha, n_t, S, tex_of_8, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        printf("%-44s layout %d S %4d smem %6zu B: %8.3f ms  %.3e samples/s  %.2f samples/clk/SM at max clock\n", name, layout, S,
               smem, best, samples / (best * 1e-3), samples / (best * 1e-3) / (prop.multiProcessorCount * (double)clk_khz * 1e3));
        return best;
    };
    auto compare = [&](const char* name) {
        std::vector<float> a(out_n), b(out_n);
        CK(cudaMemcpy(a.data(), out_a, out_n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), out_b, out_n * 4, cudaMemcpyDeviceToHost));
        double peak = 0, maxd = 0;
        size_t differ = 0;
        for (size_t i = 0; i < out_n; i++) {
            peak = fmax(peak, fabs(a[i]));
            const double d = fabs((double)a[i] - b[i]);
            if (d > 0) differ++;
            maxd = fmax(maxd, d);
        }
        printf("%s: bins that differ %zu of %zu, max|diff| %.3e = %.3e of the peak %.3f\n", name, differ, out_n, maxd, maxd / peak, peak);
    };

    const int S1 = n_u + 1;
    run(1, S1, 8, out_a, "texture only (8 of 8 warps)");
    run(1, S1, 0, out_b, "shared memory only, plain floats");
    compare("texture vs shared-memory emulation");
    for (int S : {n_u + 1, n_u + 2, n_u + 4, n_u + 8, n_u + 16, n_u + 17}) run(1, S, 0, out_b, "shared memory only, plain floats");
    run(2, S1, 0, out_b, "shared memory only, float2 pairs");
    compare("texture vs shared-memory emulation (float2 pairs)");
    for (int S : {n_u + 2, n_u + 4, n_u + 5, n_u + 8}) run(2, S, 0, out_b, "shared memory only, float2 pairs");
    for (int k = 1; k < 8; k++) {
        char name[64];
        snprintf(name, sizeof name, "hybrid: %d of 8 warps on the texture path", k);
        run(1, S1, k, out_b, name);
    }
    compare("texture vs hybrid (1 of 8 texture warps)");
    for (int k = 2; k < 7; k++) {
        char name[64];
        snprintf(name, sizeof name, "hybrid float2: %d of 8 warps on the texture path", k);
        run(2, S1, k, out_b, name);
    }
    return 0;
}
