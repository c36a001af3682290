// tma_probe.cu -- checks the TMA box load the hybrid Radon kernel relies on: 3-D tensor map over padded images,
// box = BOXW columns x 32 rows x 1 image, unswizzled, zero fill outside, arbitrary (also negative) start coordinates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int BOXW>
__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int boxes, float* out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar), dst = (unsigned)__cvta_generic_to_shared(raw);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(boxes * 32 * BOXW * 4) : "memory");
        for (int q = 0; q < boxes; q++)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst + q * 32 * BOXW * 4),
                         "l"(&map), "r"(c0), "r"(c1 + 32 * q), "r"(c2), "r"(mb)
                         : "memory");
    }
    unsigned done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mb), "r"(0) : "memory");
    } while (!done);
    const float* w = reinterpret_cast<const float*>(raw);
    for (int k = threadIdx.x; k < boxes * 32 * BOXW; k += blockDim.x) out[k] = w[k];
}

template <int BOXW>
int run(int pitch, int rows, int count, int a0, int a1, int a2, int a3)
{
    std::vector<float> h((size_t)pitch * rows * count);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003) + 1.f;
    float *img, *out;
    CK(cudaMalloc(&img, h.size() * 4));
    CK(cudaMemcpy(img, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&out, 6 * 32 * BOXW * 4));
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)count};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * rows * 4};
    const cuuint32_t box[3] = {BOXW, 32, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)sym)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("BOXW %d: encode -> %d\n", BOXW, (int)r);
    if (r) return 1;
    const int tests[][4] = {{a0, a1, a2, a3}};
    for (auto& t : tests) {
        probe<BOXW><<<1, 128, 6 * 32 * BOXW * 4>>>(map, t[0], t[1], t[2], t[3], out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  start (%d,%d,%d) boxes %d: %s\n", t[0], t[1], t[2], t[3], cudaGetErrorString(e)); return 1; }
        std::vector<float> got((size_t)t[3] * 32 * BOXW);
        CK(cudaMemcpy(got.data(), out, got.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int rr = 0; rr < t[3] * 32; rr++)
            for (int c = 0; c < BOXW; c++) {
                const int gc = t[0] + c, gr = t[1] + rr;
                const float want = (gc < 0 || gc >= pitch || gr < 0 || gr >= rows) ? 0.f : h[((size_t)t[2] * rows + gr) * pitch + gc];
                if (got[(size_t)rr * BOXW + c] != want) bad++;
            }
        printf("  start (%d,%d,%d) boxes %d: %d mismatches\n", t[0], t[1], t[2], t[3], bad);
    }
    return 0;
}

int main(int argc, char** argv)
{
    const int w = atoi(argv[1]), a0 = atoi(argv[2]), a1 = atoi(argv[3]), a2 = atoi(argv[4]), a3 = atoi(argv[5]);
    if (w == 32) return run<32>(164, 129, 6, a0, a1, a2, a3);
    return run<36>(164, 129, 6, a0, a1, a2, a3);
}
