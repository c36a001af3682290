// lds_mask_probe.cu -- does a predicated-off lane save shared-memory bandwidth?  LDS.128 of a full warp takes four
// wavefronts (8 lanes x 16 B = 128 B each).  With some lanes predicated off: does the LSU still spend a wavefront per
// quarter-warp that has ANY active lane, or does it pack the active lanes of the warp into fewer wavefronts?
// (The question behind "reload a sample's texels only in the lanes whose cell changed")
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lds_mask_probe tools/lds_mask_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(unsigned mask, int iters, unsigned long long* cycles, float* sink)
{
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < 4096; k += blockDim.x) sm[k] = make_float4(k, k + 1, k + 2, k + 3);
    __syncthreads();
    const bool on = (mask >> lane) & 1u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + (warp * 32 + lane) * 16;  // conflict-free: consecutive 16-byte slots
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            float4 v = make_float4(0, 0, 0, 0);
            const unsigned addr = base + ((i * 16 + u) & 7) * 4096;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n}\n"
                : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                : "r"(addr), "r"((unsigned)on));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc.x + acc.y + acc.z + acc.w == -1.f) sink[0] = acc.x;
}

int main()
{
    unsigned long long* cyc;
    float* sink;
    cudaMalloc(&cyc, sizeof(unsigned long long) * 1024);
    cudaMalloc(&sink, 4);
    const int threads = 512, iters = 2000;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct { const char* name; unsigned mask; } cases[] = {
        {"all 32 lanes", 0xffffffffu},
        {"lower 16 lanes (two whole quarters)", 0x0000ffffu},
        {"lanes 0-7 (one whole quarter)", 0x000000ffu},
        {"every other lane (4 of 8 in every quarter)", 0x55555555u},
        {"one lane in four (2 of 8 in every quarter)", 0x11111111u},
        {"one lane per quarter", 0x01010101u},
        {"lanes 0-3 of every quarter", 0x0f0f0f0fu},
        {"random half", 0x9d2c5680u ^ 0x5a5a1234u},
        {"single lane", 0x00000001u},
    };
    for (auto& c : cases) {
        probe<<<1, threads, 65536>>>(c.mask, 10, cyc, sink);
        probe<<<1, threads, 65536>>>(c.mask, iters, cyc, sink);
        unsigned long long h = 0;
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double per_inst = (double)h / ((double)iters * 16 * (threads / 32));
        printf("%-48s mask %08x  active %2d  %.2f cycles per warp-level LDS.128 (one CTA of %d threads on one SM)\n", c.name, c.mask,
               __builtin_popcount(c.mask), per_inst, threads);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
