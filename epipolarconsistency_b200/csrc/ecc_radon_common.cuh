// ecc_radon_common.cuh -- pieces shared by the hybrid Radon kernels (ecc_radon_hybrid.cu: one image per item,
// ecc_radon_hybrid4.cu: four images per item): the reference's bin -> line rule, the executed reference's sample set,
// mbarrier / TMA / named-barrier helpers and the two-ended work queue.
#pragma once
#include <cuda.h>

#include "ecc_geometry.cuh"
#include "ecc_internal.h"

namespace eccb200 {
namespace {

constexpr float kStep = 0.66f;
constexpr int kWindowWarps = 8;      // warps of a window group = threads / 32 that share one item
constexpr int kItemAngles = 8;       // angles per item
constexpr int kItemT = 32;           // t bins per item
constexpr int kSubTiles = 8;         // 32-bin sub-tiles per item (2 angles x 16 t each)

struct BinLine {
    float o0, o1, d0, d1, t, t_max;
    bool valid, swapped;
};

// Bin (ix,iy) -> line, clipped against the image inset by one pixel (RadonIntermediate.cu:44-92).  "swapped": the
// line misses the inset box (its "middle two" intersections come in the other order) but still passes the
// reference's entry-point test; only such lines can sample outside the image (clamp addressing).
__device__ __forceinline__ BinLine bin_line(int ix, int iy, int n_alpha, int n_t, float n_u, float n_v)
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
    const float lo1 = fminf(ta, tb), hi1 = fmaxf(ta, tb);
    const float lo2 = fminf(tc, td), hi2 = fmaxf(tc, td);
    L.t = fmaxf(lo1, lo2);
    L.t_max = fminf(hi1, hi2);
    L.swapped = fminf(hi1, hi2) < fmaxf(lo1, lo2);
    if (L.swapped) { L.t = fminf(hi1, hi2); L.t_max = fmaxf(lo1, lo2); }
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

// Which samples a bin takes.  The source loop is `for (t = t_min; t <= t_max; t += 0.66f)`, but what the reference's
// kernel EXECUTES (nvcc 12.9 -O3, sm_100; ours in ecc_radon.cu compiles to the same control flow, which is why it is
// bit-identical to it) is that loop unrolled by four with ONE test per block:
//     if (t + 1.98f <= t_max) do { four samples; t += 4 steps } while (t <= t_max - 1.98f);
//     if (t + 0.66f <= t_max) { two samples; t += 2 steps }     if (t <= t_max || nothing taken yet) one sample;
// In exact arithmetic that is the source loop; in fp32 a block's fourth sample can lie an ulp past t_max where the
// source loop would have stopped (about one bin in 10^4 on rough images, one sample of ~2300).  Both paths of this
// kernel take exactly the samples of the executed reference: the texture path by having this shape, the window path
// by walking t ahead once (ref_last_sample) and stopping at the last sample's t.
__device__ __forceinline__ float ref_last_sample(float t, float t_max)
{
    float last = t;
    bool none_yet = true;
    if (!(t + 1.98f > t_max)) {
        const float r3 = t_max - 1.98f;
#pragma unroll 1
        do {
            last = ((t + kStep) + kStep) + kStep;
            t = last + kStep;
        } while (!(t > r3));
        none_yet = false;
    }
    const float t1 = t + kStep;
    if (!(t1 > t_max)) {
        last = t1;
        t = t1 + kStep;
        none_yet = false;
    }
    if (t <= t_max || none_yet) last = t;
    return last;
}

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar)
        : "memory");
}
__device__ __forceinline__ void group_sync()  // the window warps only (named barrier 1)
{
    asm volatile("bar.sync 1, %0;" ::"n"(kWindowWarps * 32) : "memory");
}

// ---- two-ended work queue ----------------------------------------------------------------------------------------
// counters[0]: items handed out from the front (window groups), counters[1]: 32-bin sub-tiles handed out from the back
// (texture warps); both only ever fetch-add.  claim[item]: 0 free, 1 window group, 2 texture warps -- set by the first
// taker with a compare-and-swap on the item's own word.  Fronts claim in increasing, backs in decreasing item order, so
// whoever finds an item claimed by the other side has met it and stops; every item is claimed exactly once.
constexpr unsigned kClaimWindow = 1u, kClaimTexture = 2u;
__device__ __forceinline__ int take_front(unsigned* counters, unsigned* claim, unsigned total_items)
{
    const unsigned f = atomicAdd(&counters[0], 1u);
    if (f >= total_items) return -1;
    return atomicCAS(&claim[f], 0u, kClaimWindow) == 0u ? (int)f : -1;
}
// returns the sub-tile index b (item = total_items - 1 - b / kSubTiles) or -1
__device__ __forceinline__ int take_back(unsigned* counters, unsigned* claim, unsigned total_items)
{
    const unsigned b = atomicAdd(&counters[1], 1u);
    if (b >= total_items * kSubTiles) return -1;
    const unsigned seen = atomicCAS(&claim[total_items - 1u - b / kSubTiles], 0u, kClaimTexture);
    return seen == kClaimWindow ? -1 : (int)b;
}

__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned mbar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(mbar)
        : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && sym &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    }
    return fn;
}

inline int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace
}  // namespace eccb200
