// gather_probe.cu -- what the B200 delivers for RANDOM gathers, the access class of the pair kernel ("Roofline
// peak = measured L2->SM gather bandwidth of the box, random-32B-sector microbenchmark").
//   (1) random 32-byte sectors through the global-load path, working sets from L2-resident (8..96 MB) to the C3 footprint
//       of all Radon intermediates (1.17 GB, HBM),
//   (2) random bilinear fetches through the texture path over pitch-2D textures of 768 x 768 floats (the pair kernel's
//       own lookup, without its locality; one texture per warp and fetch): 16 B of taps per fetch.  A first version drew
//       the texture per LANE: 3.8e9 fetches/s with 496 textures against 1.3e11 with 8 -- a fetch is issued once per
//       distinct handle in the warp, which is why the pair kernel keeps a pair (two handles) per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);    \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ unsigned mix(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// every thread: `iters` rounds of 8 independent 32-byte sector loads (two float4) at hashed sector indices
__global__ void __launch_bounds__(256) sector_gather(const float4* __restrict__ data, unsigned n_sectors, int iters, float* sink)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    unsigned state = mix(tid * 2654435761u + 12345u);
    for (int i = 0; i < iters; i++) {
        float4 a[8], b[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            state = mix(state + 0x9e3779b9u);
            const unsigned s = (unsigned)(((unsigned long long)state * n_sectors) >> 32);
            a[k] = __ldg(&data[(size_t)s * 2]);
            b[k] = __ldg(&data[(size_t)s * 2 + 1]);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += a[k].x + a[k].w + b[k].y + b[k].z;
    }
    if (acc == 1.2345e-30f) sink[0] = acc;
}

__global__ void __launch_bounds__(256) texture_gather(const cudaTextureObject_t* __restrict__ texs, int n_tex, int iters, float* sink)
{
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    unsigned state = mix(tid * 2654435761u + 777u);
    unsigned wstate = mix((tid >> 5) * 40503u + 99u);  // per warp: the texture of a fetch is uniform over the warp, as in the
                                                       // pair kernel (a warp works on one pair = two intermediates)
    for (int i = 0; i < iters; i++) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            state = mix(state + 0x9e3779b9u);
            wstate = mix(wstate + 0x9e3779b9u);
            const unsigned t = (unsigned)(((unsigned long long)(wstate & 0xffffu) * n_tex) >> 16);
            const unsigned s2 = mix(state);
            const float x = (s2 & 0xffffu) * (1.f / 65536.f), y = (s2 >> 16) * (1.f / 65536.f);
            v[k] = tex2D<float>(texs[t], x, y);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k];
    }
    if (acc == 1.2345e-30f) sink[0] = acc;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    printf("device: %s, %d SMs, L2 %.0f MB, SM clock (max) %d MHz\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0, clock_khz / 1000);
    float* sink;
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;

    const size_t max_bytes = 1170ull << 20;
    float4* data;
    CK(cudaMalloc(&data, max_bytes));
    CK(cudaMemset(data, 0, max_bytes));
    printf("\n(1) random 32-byte sectors, global-load path (%d CTAs x %d threads, 8 sectors in flight per thread)\n", blocks, threads);
    const size_t sets_mb[] = {8, 32, 64, 96, 256, 1170};
    for (size_t mb : sets_mb) {
        const unsigned n_sectors = (unsigned)((mb << 20) / 32);
        const int iters = 64;
        sector_gather<<<blocks, threads>>>(data, n_sectors, iters, sink);  // warm (fills L2 where it fits)
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            sector_gather<<<blocks, threads>>>(data, n_sectors, iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double sectors = (double)blocks * threads * iters * 8;
        printf("   working set %5zu MB: %8.1f GB/s  (%.3e sectors/s, %.3f ms)\n", mb, sectors * 32 / (best * 1e-3) / 1e9, sectors / (best * 1e-3), best);
    }

    printf("\n(2) random bilinear fetches, texture path, pitch-2D float textures 768 x 768 (normalised, linear, clamp)\n");
    const int n_a = 768, n_t = 768;
    for (int n_tex : {8, 40, 496}) {
        std::vector<cudaTextureObject_t> texs(n_tex);
        for (int k = 0; k < n_tex; k++) {
            cudaResourceDesc res = {};
            res.resType = cudaResourceTypePitch2D;
            res.res.pitch2D.devPtr = (float*)data + (size_t)k * n_a * n_t;
            res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
            res.res.pitch2D.width = n_a;
            res.res.pitch2D.height = n_t;
            res.res.pitch2D.pitchInBytes = sizeof(float) * n_a;
            cudaTextureDesc td = {};
            td.normalizedCoords = 1;
            td.filterMode = cudaFilterModeLinear;
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.readMode = cudaReadModeElementType;
            CK(cudaCreateTextureObject(&texs[k], &res, &td, nullptr));
        }
        cudaTextureObject_t* texs_d;
        CK(cudaMalloc(&texs_d, sizeof(cudaTextureObject_t) * n_tex));
        CK(cudaMemcpy(texs_d, texs.data(), sizeof(cudaTextureObject_t) * n_tex, cudaMemcpyHostToDevice));
        const int iters = 32;
        texture_gather<<<blocks, threads>>>(texs_d, n_tex, iters, sink);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaEventRecord(e0));
            texture_gather<<<blocks, threads>>>(texs_d, n_tex, iters, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double fetches = (double)blocks * threads * iters * 8;
        printf("   %3d textures (%6.1f MB): %.3e fetches/s = %8.1f GB/s of taps (16 B per fetch), %.3f ms\n", n_tex,
               n_tex * (double)n_a * n_t * 4 / 1048576.0, fetches / (best * 1e-3), fetches * 16 / (best * 1e-3) / 1e9, best);
        for (auto t : texs) cudaDestroyTextureObject(t);
        cudaFree(texs_d);
    }
    return 0;
}
