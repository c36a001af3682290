// tex_locality_probe.cu -- development probe: how much does the float4 bilinear fetch rate of the texture unit depend on WHERE
// the 32 lanes of a warp (and the 4 lanes of a quad) sample?  The Radon kernel's texture path walks lines with a 0.66 px step;
// the probe runs that walk with different lane -> (line, sample) arrangements and reports fetches / clk / SM.
//   A  32 parallel lines 2.04 px apart, one sample per lane and step                      (no footprint shared)
//   B  quads of 2 angles x 2 t (the shipped texture-path tiling), 16 t x 2 angles per warp
//   C  quads = 4 consecutive samples of ONE line (0.66 px apart), 8 parallel lines per warp
//   D  8 consecutive samples of one line per 8 lanes, 4 parallel lines per warp
//   E  32 consecutive samples of one line per warp
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tex_locality_probe tools/tex_locality_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void walk(cudaTextureObject_t tex, int w, int h, float angle, int iters, float* out)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int line, phase, per_step;  // which parallel line of the warp, which sample of a step, samples per step and line
    float dang = 0.f;
    if (MODE == 0) { line = lane; phase = 0; per_step = 1; }
    else if (MODE == 1) { line = (lane >> 2) * 2 + ((lane >> 1) & 1); phase = 0; per_step = 1; dang = (lane & 1) * 0.00409f; }
    else if (MODE == 2) { line = lane >> 2; phase = lane & 3; per_step = 4; }
    else if (MODE == 3) { line = lane >> 3; phase = lane & 7; per_step = 8; }
    else if (MODE == 4) { line = 0; phase = lane; per_step = 32; }
    else { line = (lane >> 2) * 2 + ((lane >> 1) & 1); phase = 0; per_step = 1; dang = (lane & 1) * 0.00409f; }  // F: B with the kernel's skew
    const float a = angle + dang;
    const float dx = cosf(a), dy = sinf(a);
    // warps tile the image: origin of the warp's first line, lines 2.04 px apart along the normal
    const float ox = 40.f + (warp % 37) * 3.1f - dy * 2.04f * line, oy = 40.f + ((warp / 37) % 29) * 3.3f + dx * 2.04f * line;
    float t = 0.66f * phase;
    // F: the lines of the Radon kernel start where they ENTER the image (an axis-aligned edge), not at the foot of a common
    // normal: at the same loop iteration neighbouring lines are 2.04 / cos apart ALONG THE EDGE, i.e. shifted along their own
    // direction by 2.04 tan(angle to the edge normal) per line
    if (MODE == 5) {
        const float th = fabsf(angle) > 0.7853982f ? 1.5707963f - fabsf(angle) : fabsf(angle);
        t += 2.04f * tanf(th) * line;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float span = 700.f;
    for (int k = 0; k < iters; k++) {
        const float4 v = tex2D<float4>(tex, ox + t * dx, oy + t * dy);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        t += 0.66f * per_step;
        if (t > span) t -= span;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int w = 1240, h = 960;
    std::vector<float4> img((size_t)w * h);
    for (size_t i = 0; i < img.size(); i++) img[i] = make_float4((float)(i % 251), (float)(i % 127), (float)(i % 61), (float)(i % 31));
    cudaChannelFormatDesc d = cudaCreateChannelDesc<float4>();
    cudaArray_t arr;
    CK(cudaMallocArray(&arr, &d, w, h));
    CK(cudaMemcpy2DToArray(arr, 0, 0, img.data(), sizeof(float4) * w, sizeof(float4) * w, h, cudaMemcpyHostToDevice));
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = arr;
    cudaTextureDesc td = {};
    td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &res, &td, nullptr));
    const int threads = prop.multiProcessorCount * 1024, iters = 2000;
    float* o_d;
    CK(cudaMalloc(&o_d, (size_t)threads * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const char* names[6] = {"A 32 parallel lines", "B quads 2 angles x 2 t", "C quads = 4 samples of a line", "D 8 samples x 4 lines", "E 32 samples of a line",
                            "F = B, lines skewed as they enter an edge"};
    const float angles[6] = {0.05f, 0.3f, 0.5f, 0.65f, 0.78f, 1.4f};
    for (int ai = 0; ai < 6; ai++)
        for (int mode = 0; mode < 6; mode++) {
            float ms = 0;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) walk<0><<<threads / 256, 256>>>(tex, w, h, angles[ai], iters, o_d);
                else if (mode == 1) walk<1><<<threads / 256, 256>>>(tex, w, h, angles[ai], iters, o_d);
                else if (mode == 2) walk<2><<<threads / 256, 256>>>(tex, w, h, angles[ai], iters, o_d);
                else if (mode == 3) walk<3><<<threads / 256, 256>>>(tex, w, h, angles[ai], iters, o_d);
                else if (mode == 4) walk<4><<<threads / 256, 256>>>(tex, w, h, angles[ai], iters, o_d);
                else walk<5><<<threads / 256, 256>>>(tex, w, h, angles[ai], iters, o_d);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
                cudaEventElapsedTime(&ms, e0, e1);
            }
            const double fetches = (double)threads * iters;
            printf("angle %.2f rad  %-30s: %.3f ms, %.3f float4 fetches/clk/SM\n", angles[ai], names[mode], ms,
                   fetches / (ms * 1e-3) / prop.multiProcessorCount / (clk_khz * 1e3));
        }
    return 0;
}
