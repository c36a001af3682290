// ecc_radon_hybrid4.cu -- the hybrid Radon kernel over QUADS of projections (sm_100a).
//
// Every projection of a data set has the same bin geometry: the sample positions, cell indices and bilinear weights of
// a bin do not depend on the image.  ecc_radon_hybrid.cu spends ~33 instructions per sample in its shared-memory path,
// ~19 of them on exactly that geometry.  Here four projections are interleaved texel by texel (float4): the window path
// computes position, cell and weights once and applies them to four images with four 16-byte shared-memory loads
// (~14 instructions per sample), and the texture path fetches float4 texels (one fetch = four images, same
// channel-sample rate of the texture unit, a quarter of the instructions).  Queue, chunking, row windows, fall-backs and
// the sample set are those of ecc_radon_hybrid.cu; per image the arithmetic is identical, so are the results.
//   window: [chunk + 4 columns][rows][4 images] floats, fetched by ONE 4-D TMA box per chunk from an edge-replicated
//           interleaved copy (transposed for near-horizontal lines); the innermost box dimension is the image quad
//           (16 bytes), so no start coordinate needs an alignment (cf. the 16-byte rule found in tools/tma_probe.cu).
//           The window's shape is a template parameter (Win4 below): one 64 KB window in general, two 30 KB windows --
//           the next chunk's TMA load under the current chunk's samples -- where the t-bin spacing lets a band of 32
//           bins fit 137 rows; never more shared memory than that, because the texture path lives on the rest as L1.
//   queue:  two-ended at run time (ECC_INTERP_HYBRID), or a static split of every quad's items between the two paths,
//           58-60 % of the samples to the window path (ECC_INTERP_HYBRID_STATIC: reproducible, batch-invariant).
//   several GPUs: every finished bin is also stored into the other ranks' buffers (Mirrors, ecc_team.cu).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ecc_radon_common.cuh"

namespace eccb200 {

namespace {

constexpr int kLead4 = 2;  // window columns in front of the chunk

// Window configuration (compile time): CHUNK pixels of the primary axis per chunk, window = [CHUNK + 4 columns][ROWS rows]
// [4 images]; ROWS = 1 mod 8, so that neighbouring columns start one 16-byte bank group apart; NBUF window buffers
// (2 = the TMA load of the next chunk runs under the current chunk's samples).
//   Win4<16, 201, 1>  one 64 KB window: the general configuration (bands of up to 201 rows: t-bin spacing up to ~4.5 px).
//   Win4<10, 137, 2>  two 30 KB windows in the SAME shared memory (the carve-out stays at 132 KB, the texture path keeps its
//                     124 KB of L1TEX): for t-bin spacings up to 2.1 px (measured at 2.04 and 1.04), where a band of 32 bins fits in 137 rows.
// Measured (C2 size, ms/projection, both pipes): 16/201/1 0.616; with two buffers 12/201 0.859, 8/201 0.785 (the second
// buffer comes out of L1TEX: carve-out 164-228 KB), 8/169 0.605, 8/161..145 0.588, 6/193 0.604, 10/137 0.578.
template <int CHUNK, int ROWS, int NBUF, bool CHUNK_FALLBACK = false>
struct Win4 {
    static constexpr int chunk = CHUNK, rows = ROWS, nbuf = NBUF, boxw = CHUNK + 4, max_chunks = 2048 / CHUNK;
    // a chunk whose band of rows does not fit the window: false = the whole item goes through the texture unit,
    // true = only that chunk's samples do (fetched by the window warps themselves, in sample order)
    static constexpr bool chunk_fallback = CHUNK_FALLBACK;
#ifndef ECC_ANY_ROWS  // development: other residues of the window height (the bank-group offset between neighbouring columns)
    static_assert(ROWS % 8 == 1, "window height: 1 mod 8 keeps neighbouring columns one bank group apart");
#endif
    static_assert(max_chunks <= kWindowWarps * 32, "one window thread per chunk initialises the row tables");
    static constexpr size_t buf_bytes = (size_t)ROWS * boxw * 16;
    static constexpr size_t smem = (buf_bytes + 127) / 128 * 128 * NBUF;
};
typedef Win4<16, 201, 1> Win4General;
#ifndef ECC_FINE_CHUNK  // development knobs: -DECC_FINE_CHUNK=.. -DECC_FINE_ROWS=.. -DECC_FINE_CF=0|1
#define ECC_FINE_CHUNK 10
#define ECC_FINE_ROWS 137
#endif
#ifndef ECC_FINE_CF
#define ECC_FINE_CF 0
#endif
typedef Win4<ECC_FINE_CHUNK, ECC_FINE_ROWS, 2, (ECC_FINE_CF != 0)> Win4Fine;

struct Hybrid4Params {
    const cudaTextureObject_t* texs;  // one float4 texture per quad
    int n_quads, n_img;               // n_img = images that exist (the last quad may be partly empty)
    int n_u, n_v, n_alpha, n_t, post;
    int groups_a, groups_t;
    int mode;
    int tex_map;   // development (ECC_HYBRID4_TEXMAP): bin tiling of a texture warp's sub-tile, see the texture warps' loop
    int lane_map;  // bin tiling of a window warp: 2 = 4 angles x 8 t (default; measured: window path alone 1.10 ms/projection),
                   // 0 = 2 angles x 16 t as the texture warps (1.20), 1 = 1 angle x 32 t (1.42), 3 = 8 angles x 4 t (1.26)
    unsigned* counters;
    unsigned* claim;
    unsigned magic;
    unsigned long long* item_clock;  // development (ECC_ITEM_CLOCK=file): [2][items per quad] cycles spent per item, texture / window path
    int ag_lo, ag_hi;  // development (ECC_ITEM_AG_LO / _HI): only the items of these angle groups are computed (timing per angle)
    int split_items;  // < 0: two-ended run-time queue; >= 0: items order[0 .. split_items) of every quad -> window warps, the
                      // rest -> texture warps (reproducible results: the path of a bin is a function of the geometry alone)
    const int* order; // static split: a permutation of a quad's items (static_split_items)
    // static split: the launch's share of the two item lists (QuadPart: a team shares the quad that straddles two ranks'
    // shards) -- window entries [w_begin, w_end) of n_quads * split_items, texture sub-tiles [x_begin, x_end) of
    // n_quads * (items per quad - split_items) * kSubTiles
    unsigned w_begin, w_end, x_begin, x_end;
    // Window entries from w_last on -- the launch's last quad -- are taken in REVERSE order.  A quad's items are numbered t group
    // by t group from one edge of the t axis, so the window list [0, split) ends with the t groups at the centre: the longest
    // lines, half a millisecond per item, and a launch ended with a few CTAs on those while the rest idled (0.43 ms per launch
    // beyond linear, profiles/radon_launch_size_r02.txt).  Backwards, the launch ends with the short lines at the edge.  The
    // texture list (from the centre to the other edge) already runs that way.  Only the last quad: walking EVERY quad
    // backwards made long launches 1.3 % slower (window and texture warps then work on the centre at the same time).
    unsigned w_last;
    float* out;
    Mirrors mir;  // multi-GPU team: every bin is also stored into the other ranks' buffers (NVLink peer stores)
};

struct Sum4 {
    float x, y, z, w;
};
__device__ __forceinline__ void add4(Sum4& s, const float4& v) { s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }

// The whole bin of four images through the texture unit (float4 texels), in the executed reference's shape
// (ecc_radon_common.cuh: ref_last_sample); two samples per line in flight.
__device__ __forceinline__ void bin_texture4(cudaTextureObject_t tex, const BinLine& L, Sum4& res)
{
    float o0 = L.o0 + 0.5f, o1 = L.o1 + 0.5f;
    const float d0 = L.d0, d1 = L.d1, t_max = L.t_max;
    o0 -= 0.5f * d1;
    o1 += 0.5f * d0;
    Sum4 sum = {0.f, 0.f, 0.f, 0.f}, sumo = {0.f, 0.f, 0.f, 0.f};
    float t = L.t;
    bool none_yet = true;
#define ECC_TEX4_A(tt) tex2D<float4>(tex, fmaf((tt), d0, o0), fmaf((tt), d1, o1))
#define ECC_TEX4_B(tt) tex2D<float4>(tex, fmaf((tt), d0, o0) + d1, fmaf((tt), d1, o1) - d0)
    if (!(t + 1.98f > t_max)) {
        const float r3 = t_max - 1.98f;
#pragma unroll 1
        do {
            const float t1 = t + kStep, t2 = t1 + kStep, t3 = t2 + kStep;
            const float4 a0 = ECC_TEX4_A(t), b0 = ECC_TEX4_B(t), a1 = ECC_TEX4_A(t1), b1 = ECC_TEX4_B(t1);
            const float4 a2 = ECC_TEX4_A(t2), b2 = ECC_TEX4_B(t2), a3 = ECC_TEX4_A(t3), b3 = ECC_TEX4_B(t3);
            add4(sum, a0); add4(sumo, b0);
            add4(sum, a1); add4(sumo, b1);
            add4(sum, a2); add4(sumo, b2);
            add4(sum, a3); add4(sumo, b3);
            t = t3 + kStep;
        } while (!(t > r3));
        none_yet = false;
    }
    const float t1 = t + kStep;
    if (!(t1 > t_max)) {
        const float4 a0 = ECC_TEX4_A(t), b0 = ECC_TEX4_B(t), a1 = ECC_TEX4_A(t1), b1 = ECC_TEX4_B(t1);
        add4(sum, a0); add4(sumo, b0);
        add4(sum, a1); add4(sumo, b1);
        t = t1 + kStep;
        none_yet = false;
    }
    if (t <= t_max || none_yet) {
        add4(sum, ECC_TEX4_A(t));
        add4(sumo, ECC_TEX4_B(t));
    }
#undef ECC_TEX4_A
#undef ECC_TEX4_B
    res.x = (sum.x - sumo.x) * kStep;
    res.y = (sum.y - sumo.y) * kStep;
    res.z = (sum.z - sumo.z) * kStep;
    res.w = (sum.w - sumo.w) * kStep;
}

// ---- packed fp32 pairs (sm_100: add / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2) ----------------------------------------
// The window path's per-image filter arithmetic is the same eight IEEE operations for each of the four interleaved
// images; Blackwell executes them two images per instruction.  Each half of a packed operation is the separately rounded
// fp32 operation (round to nearest), so the results are those of the scalar code bit for bit (checked: same checksums as
// the scalar build for every split) -- what changes is the number of issue slots per sample position.
//   ECC_F32X2 = 0  scalar (round 1): 113 instructions per loop iteration (two positions x four images)
//   ECC_F32X2 = 1  filter AND running sums packed: 90 (ptxas moves the 64-bit running sums back after every sample)
//   ECC_F32X2 = 2  filter packed, running sums scalar: 86 -- the default
// Measured on B200 (C2 size, ms/projection, run-time queue / window warps only; tools/radon_ab.sh, profiles/radon_ab_r02.txt):
//   scalar 0.564 / 0.898,  mode 1 0.575 / 0.927,  mode 2 0.560 / 0.889.
// A quarter fewer instructions buy under 1 %: the window path is bound by shared-memory wavefronts (4 x LDS.128 per
// position, 1.39 wavefronts per conflict-free one), not by issue slots.
#ifndef ECC_F32X2
#define ECC_F32X2 2
#endif
#ifndef ECC_WINDOW_UNROLL  // unroll factor of the window path's sample loop (2 changes nothing measurable: 0.561 / 0.893)
#define ECC_WINDOW_UNROLL 1
#endif
constexpr int kWindowUnroll = ECC_WINDOW_UNROLL;
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ void acc2(f32x2& a, f32x2 b)  // a += b in place (keeps the running sum in its register pair)
{
    asm("add.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Running sums of a line for the four images of a quad, as the window path keeps them.
#if ECC_F32X2 == 1
struct Acc4 {
    f32x2 xy, zw;
};
__device__ __forceinline__ Acc4 acc_zero() { Acc4 a; a.xy = pk(0.f, 0.f); a.zw = a.xy; return a; }
__device__ __forceinline__ void add4(Acc4& s, const float4& v) { s.xy = add2(s.xy, pk(v.x, v.y)); s.zw = add2(s.zw, pk(v.z, v.w)); }
__device__ __forceinline__ Sum4 acc_unpack(const Acc4& a) { Sum4 r; unpk(a.xy, r.x, r.y); unpk(a.zw, r.z, r.w); return r; }
#else
typedef Sum4 Acc4;
__device__ __forceinline__ Acc4 acc_zero() { Acc4 a = {0.f, 0.f, 0.f, 0.f}; return a; }
__device__ __forceinline__ Sum4 acc_unpack(const Acc4& a) { return a; }
#endif

// One sample position applied to the four images of the window: position -> cell and fractions -> weights once
// (ecc_radon_hybrid.cu: sample_window has the derivation), four 16-byte loads, then per image
// v00 + (w10 (v10-v00) + w01 (v01-v00) + w11 (v11-v00)) / 256.
template <int kRows4>
__device__ __forceinline__ void sample_window4(unsigned base, float pri, float sec, Acc4& acc)
{
    const float Ps = fmaf(pri, 256.f, -127.5f);
    const float Ss = fmaf(sec, 256.f, -127.5f);
    const unsigned Pi = __float_as_uint(__fadd_rd(Ps, 8388608.f));
    const unsigned Si = __float_as_uint(__fadd_rd(Ss, 8388608.f));
    const unsigned a = Pi & 255u, b = Si & 255u;
#ifdef ECC_CONFLICT_PROBE  // development: rows forced onto distinct bank groups per 8 lanes (WRONG results; speed bound)
    const unsigned addr = base + (((Pi >> 8) * kRows4 + (((Si >> 8) & ~7u) | ((threadIdx.x - (Pi >> 8)) & 7u))) << 4);
#else
    const unsigned addr = base + (((Pi >> 8) * kRows4 + (Si >> 8)) << 4);
#endif
    const unsigned w11 = (a * b + 128u) >> 8;
    const float f11 = __uint2float_rn(w11), f10 = __uint2float_rn(a - w11), f01 = __uint2float_rn(b - w11);
#if ECC_F32X2
    f32x2 v00a, v00b, v01a, v01b, v10a, v10b, v11a, v11b;  // a = images 0, 1; b = images 2, 3
    asm volatile(
        "ld.shared.v2.b64 {%0, %1}, [%8];\n"
        "ld.shared.v2.b64 {%2, %3}, [%8+16];\n"
        "ld.shared.v2.b64 {%4, %5}, [%8+%9];\n"
        "ld.shared.v2.b64 {%6, %7}, [%8+%10];\n"
        : "=l"(v00a), "=l"(v00b), "=l"(v01a), "=l"(v01b), "=l"(v10a), "=l"(v10b), "=l"(v11a), "=l"(v11b)
        : "r"(addr), "n"(kRows4 * 16), "n"(kRows4 * 16 + 16));  // the next column: one window height further
    const f32x2 p10 = pk(f10, f10), p01 = pk(f01, f01), p11 = pk(f11, f11), scale = pk(0.00390625f, 0.00390625f);
    f32x2 t = mul2(p10, sub2(v10a, v00a));
    t = fma2(p01, sub2(v01a, v00a), t);
    t = fma2(p11, sub2(v11a, v00a), t);
#if ECC_F32X2 == 1
    acc2(acc.xy, fma2(t, scale, v00a));
#else  // packed filter, scalar running sums
    float r0, r1, r2, r3;
    unpk(fma2(t, scale, v00a), r0, r1);
    acc.x += r0;
    acc.y += r1;
#endif
    t = mul2(p10, sub2(v10b, v00b));
    t = fma2(p01, sub2(v01b, v00b), t);
    t = fma2(p11, sub2(v11b, v00b), t);
#if ECC_F32X2 == 1
    acc2(acc.zw, fma2(t, scale, v00b));
#else
    unpk(fma2(t, scale, v00b), r2, r3);
    acc.z += r2;
    acc.w += r3;
#endif
#else
    float4 v00, v01, v10, v11;
    asm volatile(
        "ld.shared.v4.f32 {%0, %1, %2, %3}, [%16];\n"
        "ld.shared.v4.f32 {%4, %5, %6, %7}, [%16+16];\n"
        "ld.shared.v4.f32 {%8, %9, %10, %11}, [%16+%17];\n"
        "ld.shared.v4.f32 {%12, %13, %14, %15}, [%16+%18];\n"
        : "=f"(v00.x), "=f"(v00.y), "=f"(v00.z), "=f"(v00.w), "=f"(v01.x), "=f"(v01.y), "=f"(v01.z), "=f"(v01.w), "=f"(v10.x),
          "=f"(v10.y), "=f"(v10.z), "=f"(v10.w), "=f"(v11.x), "=f"(v11.y), "=f"(v11.z), "=f"(v11.w)
        : "r"(addr), "n"(kRows4 * 16), "n"(kRows4 * 16 + 16));  // the next column: one window height further
#define ECC_ONE(c)                                                              \
    {                                                                           \
        float t = f10 * (v10.c - v00.c);                                        \
        t = fmaf(f01, v01.c - v00.c, t);                                        \
        t = fmaf(f11, v11.c - v00.c, t);                                        \
        acc.c += fmaf(t, 0.00390625f, v00.c);                                   \
    }
    ECC_ONE(x) ECC_ONE(y) ECC_ONE(z) ECC_ONE(w)
#undef ECC_ONE
#endif
}

struct Item4 {
    int quad, ix, iy;
};
__device__ __forceinline__ Item4 item_bins4(int item, int wi, int lane, const Hybrid4Params& p)
{
    Item4 r;
    const int per_quad = p.groups_a * p.groups_t;
    r.quad = item / per_quad;
    const int rem = item - r.quad * per_quad;
    const int tg = rem / p.groups_a, ag = rem - tg * p.groups_a;
    const int q4 = lane >> 2, ql = lane & 3;
    r.ix = ag * kItemAngles + (wi & 3) * 2 + (ql & 1);
    r.iy = tg * kItemT + (wi >> 2) * 16 + q4 * 2 + (ql >> 1);
    return r;
}

__device__ __forceinline__ void store4(const Hybrid4Params& p, const Item4& B, const Sum4& r, bool valid)
{
    const size_t stride = (size_t)p.n_t * p.n_alpha;
    float* dst = p.out + (size_t)B.quad * 4 * stride + (size_t)B.iy * p.n_alpha + B.ix;
    const int left = p.n_img - B.quad * 4;
    const float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (c < left) {
            const float val = valid ? post_process(v[c], p.post) : 0.f;
            float* d = dst + c * stride;
            *d = val;
            for (int r = 0; r < p.mir.n; r++) *(float*)((char*)d + p.mir.delta[r]) = val;
        }
}

template <int MAXTHREADS, int MINBLOCKS, typename W>
__global__ void __launch_bounds__(MAXTHREADS, MINBLOCKS)
radon_hybrid4_kernel(const __grid_constant__ CUtensorMap map_n, const __grid_constant__ CUtensorMap map_t,
                     const __grid_constant__ Hybrid4Params p)
{
    constexpr int kChunk4 = W::chunk, kBoxW4 = W::boxw, kRows4 = W::rows, kMaxChunks4 = W::max_chunks, kNBuf4 = W::nbuf;
    extern __shared__ __align__(128) unsigned char window_raw[];
    __shared__ __align__(8) unsigned long long mbar_store[2];
    __shared__ int s_item, s_jmin, s_jmax, s_fallback;
    __shared__ int win_lo[kMaxChunks4], win_hi[kMaxChunks4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned total_items = (unsigned)(p.n_quads * p.groups_a * p.groups_t);
    const float n_u = (float)p.n_u, n_v = (float)p.n_v;

    if (tid == 0) {
        mbar_init(smem_u32(&mbar_store[0]), 1);
        mbar_init(smem_u32(&mbar_store[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kWindowWarps) {
        if (p.mode == 2) return;
        // ---------------- texture warps ----------------
        for (;;) {
            int sub = 0, item;
            if (p.split_items >= 0) {  // static split: sub-tiles of the items [split_items, per_quad) of every quad, in order
                const int tex_items = p.groups_a * p.groups_t - p.split_items;
                if (lane == 0) sub = (int)(atomicAdd(&p.counters[1], 1u) + p.x_begin);
                sub = __shfl_sync(0xffffffffu, sub, 0);
                if ((unsigned)sub >= p.x_end) break;
                const int x = sub / kSubTiles;
                item = (x / tex_items) * (p.groups_a * p.groups_t) + __ldg(&p.order[p.split_items + x % tex_items]);
            } else {
                if (lane == 0) sub = take_back(p.counters, p.claim, total_items);
                sub = __shfl_sync(0xffffffffu, sub, 0);
                if (sub < 0) break;
                item = (int)total_items - 1 - sub / kSubTiles;
            }
            {
                const int ag_i = (item % (p.groups_a * p.groups_t)) % p.groups_a;
                if (ag_i < p.ag_lo || ag_i > p.ag_hi) continue;
            }
            const long long clock0 = p.item_clock ? clock64() : 0;
            Item4 B = item_bins4(item, sub % kSubTiles, lane, p);
            if (p.tex_map) {  // development: other tilings of the item's 8 angles x 32 t bins into 8 sub-tiles of 32 lanes
                const int wi = sub % kSubTiles, a0 = (B.ix / kItemAngles) * kItemAngles, t0 = (B.iy / kItemT) * kItemT;
                if (p.tex_map == 1) {         // sub-tile 2 angles x 16 t, quad = 1 angle x 4 t
                    B.ix = a0 + (wi & 3) * 2 + (lane >> 4);
                    B.iy = t0 + (wi >> 2) * 16 + (lane & 15);
                } else if (p.tex_map == 2) {  // sub-tile 4 angles x 8 t, quad = 4 angles x 1 t
                    B.ix = a0 + (wi & 1) * 4 + (lane & 3);
                    B.iy = t0 + (wi >> 1) * 8 + (lane >> 2);
                } else if (p.tex_map == 3) {  // sub-tile 8 angles x 4 t, quad = 4 angles x 1 t
                    B.ix = a0 + (lane & 7);
                    B.iy = t0 + wi * 4 + (lane >> 3);
                } else if (p.tex_map == 4) {  // sub-tile 1 angle x 32 t, quad = 1 angle x 4 t
                    B.ix = a0 + wi;
                    B.iy = t0 + lane;
                } else if (p.tex_map == 5) {  // sub-tile 4 angles x 8 t, quad = 2 angles x 2 t
                    B.ix = a0 + (wi & 1) * 4 + ((lane >> 4) & 1) * 2 + (lane & 1);
                    B.iy = t0 + (wi >> 1) * 8 + ((lane >> 2) & 3) * 2 + ((lane >> 1) & 1);
                }
            }
            if (B.ix >= p.n_alpha || B.iy >= p.n_t) continue;
            const BinLine L = bin_line(B.ix, B.iy, p.n_alpha, p.n_t, n_u, n_v);
            Sum4 r = {0.f, 0.f, 0.f, 0.f};
            if (L.valid) bin_texture4(p.texs[B.quad], L, r);
            store4(p, B, r, L.valid);
            if (p.item_clock) {
                __syncwarp();
                if (lane == 0) atomicAdd(&p.item_clock[item % (p.groups_a * p.groups_t)], (unsigned long long)(clock64() - clock0));
            }
        }
        return;
    }

    // ---------------- window warps ----------------
    if (p.mode == 1) return;
    const unsigned window_base = smem_u32(window_raw);
    const unsigned buf_bytes = (unsigned)kRows4 * kBoxW4 * 16u;
    const unsigned buf_stride = (buf_bytes + 127u) & ~127u;  // TMA destinations are 128-byte aligned
    const unsigned mbar0 = smem_u32(&mbar_store[0]);
    unsigned phase0 = 0, phase1 = 0;
    for (;;) {
        if (tid == 0) {
            if (p.split_items >= 0) {  // static split: items [0, split_items) of every quad, in order
                unsigned w = atomicAdd(&p.counters[0], 1u) + p.w_begin;
                // the launch's LAST quad is walked backwards (same entries, other order): see Hybrid4Params::w_last
                if (w >= p.w_last && w < p.w_end) w = p.w_last + (p.w_end - 1u - w);
                s_item = (p.split_items > 0 && w < p.w_end)
                             ? (int)(w / p.split_items) * (p.groups_a * p.groups_t) + __ldg(&p.order[w % p.split_items]) : -1;
            } else {
                s_item = take_front(p.counters, p.claim, total_items);
            }
            s_jmin = INT_MAX;
            s_jmax = INT_MIN;
            s_fallback = 0;
        }
        if (tid < kMaxChunks4) { win_lo[tid] = INT_MAX; win_hi[tid] = INT_MIN; }
        group_sync();
        const int item = s_item;
        if (item < 0) break;
        const long long clock0 = p.item_clock ? clock64() : 0;
        {
            const int ag_i = (item % (p.groups_a * p.groups_t)) % p.groups_a;
            if (ag_i < p.ag_lo || ag_i > p.ag_hi) { group_sync(); continue; }
        }
        Item4 B = item_bins4(item, warp, lane, p);
        if (p.lane_map == 1) {  // warp = one angle, lanes = 32 t bins
            B.ix = (B.ix / kItemAngles) * kItemAngles + warp;
            B.iy = (B.iy / kItemT) * kItemT + lane;
        } else if (p.lane_map == 2) {  // warp = 4 angles x 8 t, a phase of 8 lanes = 2 t x 4 angles
            B.ix = (B.ix / kItemAngles) * kItemAngles + (warp & 1) * 4 + (lane & 3);
            B.iy = (B.iy / kItemT) * kItemT + (warp >> 1) * 8 + (lane >> 2);
        } else if (p.lane_map == 3) {  // warp = 8 angles x 4 t, a phase = 1 t x 8 angles
            B.ix = (B.ix / kItemAngles) * kItemAngles + (lane & 7);
            B.iy = (B.iy / kItemT) * kItemT + warp * 4 + (lane >> 3);
        } else if (p.lane_map == 4) {  // warp = 4 angles x 8 t, a phase = 4 t x 2 angles, angle pairs two apart
            B.ix = (B.ix / kItemAngles) * kItemAngles + (warp & 1) * 4 + ((lane >> 2) & 1) * 2 + (lane & 1);
            B.iy = (B.iy / kItemT) * kItemT + (warp >> 1) * 8 + (lane >> 3) * 2 + ((lane >> 1) & 1);
        }
        const bool in_range = B.ix < p.n_alpha && B.iy < p.n_t;
        BinLine L;
        L.valid = false;
        L.swapped = false;
        if (in_range) L = bin_line(B.ix, B.iy, p.n_alpha, p.n_t, n_u, n_v);
        bool safe = false;
        if (in_range && L.valid) {
            const float ax = L.o0 + L.t * L.d0, ay = L.o1 + L.t * L.d1, bx = L.o0 + L.t_max * L.d0, by = L.o1 + L.t_max * L.d1;
            const float lo = 1.f - 1e-3f, hx = n_u - 1.f + 1e-3f, hy = n_v - 1.f + 1e-3f;
            safe = !L.swapped && fminf(ax, bx) >= lo && fmaxf(ax, bx) <= hx && fminf(ay, by) >= lo && fmaxf(ay, by) <= hy;
        }
        const bool live = safe;

        const int per_quad = p.groups_a * p.groups_t;
        const int ag = (item % per_quad) % p.groups_a;
        const float alpha_mid = ((ag * kItemAngles + 0.5f * (kItemAngles - 1)) / (float)p.n_alpha - 0.5f) * ECC_PI_F;
        const bool vertical = fabsf(sinf(alpha_mid)) > fabsf(cosf(alpha_mid));
        const int dir = (vertical && alpha_mid < 0.f) ? -1 : 1;

        float oA0 = L.o0 + 0.5f, oA1 = L.o1 + 0.5f;
        oA0 -= 0.5f * L.d1;
        oA1 += 0.5f * L.d0;
        const float op = vertical ? oA1 : oA0, dp = vertical ? L.d1 : L.d0;
        const float os = vertical ? oA0 : oA1, ds = vertical ? L.d0 : L.d1;
        const float offp = vertical ? -L.d0 : L.d1, offs = vertical ? L.d1 : -L.d0;
        const float inv_dp = live ? 1.f / dp : 0.f;

        int jmin_l = INT_MAX, jmax_l = INT_MIN;
        if (live) {
            const float pa = fmaf(L.t, dp, op) - 0.5f, pb = fmaf(L.t_max, dp, op) - 0.5f;
            const int ja = (int)floorf(fminf(pa, pb) * (1.f / kChunk4)), jb = (int)floorf(fmaxf(pa, pb) * (1.f / kChunk4));
            jmin_l = ja - 1;
            jmax_l = jb + 1;
        }
        {
            const int wmin = __reduce_min_sync(0xffffffffu, jmin_l), wmax = __reduce_max_sync(0xffffffffu, jmax_l);
            if (lane == 0 && wmin <= wmax) { atomicMin(&s_jmin, wmin + 1); atomicMax(&s_jmax, wmax - 1); }
        }
        group_sync();
        const int jlo = s_jmin, jhi = s_jmax;
        const int n_chunks = jhi - jlo + 1;
        const bool too_long = n_chunks > kMaxChunks4;

        if (n_chunks > 0 && !too_long) {
            for (int k = 0; k < n_chunks; k++) {
                const int j = jlo + k;
                int lo = INT_MAX, hi = INT_MIN;
                if (live && j >= jmin_l && j <= jmax_l) {
                    const float t0 = ((float)(j * kChunk4 - 2) + 0.5f - op) * inv_dp;
                    const float t1 = ((float)(j * kChunk4 + kChunk4 + 2) + 0.5f - op) * inv_dp;
                    const float s0 = fmaf(fmaxf(fminf(t0, t1), L.t - 2.f), ds, os);
                    const float s1 = fmaf(fminf(fmaxf(t0, t1), L.t_max + 2.f), ds, os);
                    lo = (int)floorf(fminf(s0, s1) - 2.0f);
                    hi = (int)floorf(fmaxf(s0, s1) + 1.0f) + 1;
                }
                const int wlo = __reduce_min_sync(0xffffffffu, lo), whi = __reduce_max_sync(0xffffffffu, hi);
                if (lane == 0 && wlo <= whi) { atomicMin(&win_lo[k], wlo); atomicMax(&win_hi[k], whi); }
            }
        }
        group_sync();
        if (n_chunks > 0 && !too_long && tid < n_chunks) {
            if (!W::chunk_fallback && win_lo[tid] <= win_hi[tid] && win_hi[tid] - win_lo[tid] + 1 > kRows4) s_fallback = 1;
        }
        if (tid == 0 && too_long) s_fallback = 1;
        group_sync();
        const bool fallback = s_fallback != 0;

        Sum4 result = {0.f, 0.f, 0.f, 0.f};
        if (n_chunks > 0 && !fallback) {
            const CUtensorMap* map = vertical ? &map_n : &map_t;
            float t = live ? L.t : 3.0e38f;
            const float t_max = live ? ref_last_sample(L.t, L.t_max) : -3.0e38f;
            Acc4 sum = acc_zero(), sumo = acc_zero();
            // issue the load of chunk k (traversal order) into buffer k % kNBuf4
            auto issue = [&](int k) {
                const int kk = dir > 0 ? k : n_chunks - 1 - k;
                const int lo = win_lo[kk], hi = win_hi[kk];
                if (lo > hi) return;
                if (W::chunk_fallback && hi - lo + 1 > kRows4) return;  // sampled through the texture unit instead
                const unsigned b = (kNBuf4 == 2) ? (unsigned)(k & 1) : 0u;
                const unsigned mb = mbar0 + 8u * b;
                mbar_expect_tx(mb, buf_bytes);
                tma_load_4d(window_base + b * buf_stride, map, 0, lo, (jlo + kk) * kChunk4 - kLead4, B.quad, mb);
            };
            if (kNBuf4 == 2 && tid == 0) issue(0);
            for (int k = 0; k < n_chunks; k++) {
                const int kk = dir > 0 ? k : n_chunks - 1 - k;
                const int j = jlo + kk;
                const int lo = win_lo[kk], hi = win_hi[kk];
                group_sync();  // everybody is done with the buffer that is loaded next
                if (kNBuf4 == 2) {
                    if (tid == 0 && k + 1 < n_chunks) issue(k + 1);
                } else {
                    if (tid == 0) issue(k);
                }
                if (lo > hi) continue;
                float lim = t_max;
                if (k + 1 < n_chunks) {
                    const float edge = (float)(dir > 0 ? (j + 1) * kChunk4 : j * kChunk4) + 0.5f;
                    lim = fminf(t_max, (edge - op) * inv_dp);
                }
                if (W::chunk_fallback && hi - lo + 1 > kRows4) {
                    // this chunk's band does not fit the window: its samples through the texture unit, same positions, same
                    // order of additions; two samples (four fetches) in flight
                    const cudaTextureObject_t tex = p.texs[B.quad];
#pragma unroll 1
                    while (t <= lim) {
                        const float t1 = t + kStep;
                        const bool two = t1 <= lim;
                        const float pri = fmaf(t, dp, op), sec = fmaf(t, ds, os);
                        const float pri1 = fmaf(t1, dp, op), sec1 = fmaf(t1, ds, os);
                        const float4 a0 = tex2D<float4>(tex, vertical ? sec : pri, vertical ? pri : sec);
                        const float4 b0 = tex2D<float4>(tex, vertical ? sec + offs : pri + offp, vertical ? pri + offp : sec + offs);
                        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = a1;
                        if (two) {
                            a1 = tex2D<float4>(tex, vertical ? sec1 : pri1, vertical ? pri1 : sec1);
                            b1 = tex2D<float4>(tex, vertical ? sec1 + offs : pri1 + offp, vertical ? pri1 + offp : sec1 + offs);
                        }
                        add4(sum, a0);
                        add4(sumo, b0);
                        if (two) { add4(sum, a1); add4(sumo, b1); }
                        t = two ? t1 + kStep : t1;
                    }
                    continue;
                }
                const unsigned b = (kNBuf4 == 2) ? (unsigned)(k & 1) : 0u;
                if (b == 0) { mbar_wait(mbar0, phase0); phase0 ^= 1u; }
                else { mbar_wait(mbar0 + 8u, phase1); phase1 ^= 1u; }
                const unsigned base = window_base + b * buf_stride - 16u * ((p.magic + (unsigned)(j * kChunk4 - kLead4)) * kRows4 + p.magic + (unsigned)lo);
#pragma unroll kWindowUnroll
                for (; t <= lim; t += kStep) {
                    const float pri = fmaf(t, dp, op), sec = fmaf(t, ds, os);
                    sample_window4<kRows4>(base, pri, sec, sum);
                    sample_window4<kRows4>(base, pri + offp, sec + offs, sumo);
                }
            }
            const Sum4 sa = acc_unpack(sum), so = acc_unpack(sumo);
            result.x = (sa.x - so.x) * kStep;
            result.y = (sa.y - so.y) * kStep;
            result.z = (sa.z - so.z) * kStep;
            result.w = (sa.w - so.w) * kStep;
        }
        if (in_range) {
            if (L.valid && (fallback || !safe)) bin_texture4(p.texs[B.quad], L, result);
            store4(p, B, result, L.valid);
        }
        group_sync();
        if (p.item_clock && tid == 0)
            atomicAdd(&p.item_clock[p.groups_a * p.groups_t + item % (p.groups_a * p.groups_t)], (unsigned long long)(clock64() - clock0));
    }
}

// ---- staging: four images interleaved texel by texel ------------------------------------------------------------------
// the quad's CUDA array             written in place through its surface (texture path); round 2 until the last day: a linear
//                                   copy [q][n_v][n_u] float4 + one cudaMemcpy2DToArrayAsync per quad -- 124 copies of 19 MB
//                                   per C3 step, 5.2 ms of a 286.9 ms step during which no Radon kernel ran
// padn [q][n_v+1][n_u+1] float4        edge-replicated (window path, near-vertical lines)
// padt [q][n_u+1][n_v+1] float4        its transpose (near-horizontal lines)
// Blocks of 32 x 4 texels: a warp writes 512 contiguous bytes of a row, four rows of the array's tiles per block.
__global__ void interleave4_kernel(const float* __restrict__ src, int n_img, int n_u, int n_v,
                                   const cudaSurfaceObject_t* __restrict__ surfs, float4* __restrict__ padn)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 4 + threadIdx.y, q = blockIdx.z;
    if (x > n_u || y > n_v) return;
    const int xs = min(x, n_u - 1), ys = min(y, n_v - 1);
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int img = q * 4 + c;
        v[c] = img < n_img ? src[((size_t)img * n_v + ys) * n_u + xs] : 0.f;
    }
    const float4 t = make_float4(v[0], v[1], v[2], v[3]);
    padn[((size_t)q * (n_v + 1) + y) * (n_u + 1) + x] = t;
    if (x < n_u && y < n_v) surf2Dwrite(t, surfs[q], x * (int)sizeof(float4), y);
}
__global__ void transpose4_kernel(const float4* __restrict__ padn, int n_u, int n_v, float4* __restrict__ padt)
{
    __shared__ float4 tile[16][17];
    const int q = blockIdx.z, x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
    const float4* s = padn + (size_t)q * (n_v + 1) * (n_u + 1);
    {
        const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
        if (x <= n_u && y <= n_v) tile[threadIdx.y][threadIdx.x] = s[(size_t)y * (n_u + 1) + x];
    }
    __syncthreads();
    {
        const int x = x0 + threadIdx.y, y = y0 + threadIdx.x;  // output row = x, column = y
        if (x <= n_u && y <= n_v) padt[((size_t)q * (n_u + 1) + x) * (n_v + 1) + y] = tile[threadIdx.x][threadIdx.y];
    }
}

// Bilinear samples per item of one projection (both lines of every bin): the weights for the static split.
__global__ void item_samples_kernel(int n_u_i, int n_v_i, int n_alpha, int n_t, int groups_a, float* __restrict__ per_item)
{
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y;
    if (ix >= n_alpha || iy >= n_t) return;
    const BinLine L = bin_line(ix, iy, n_alpha, n_t, (float)n_u_i, (float)n_v_i);
    if (!L.valid) return;
    const float count = 2.f * (floorf((L.t_max - L.t) / kStep) + 1.f);
    atomicAdd(&per_item[(iy / kItemT) * groups_a + ix / kItemAngles], count);
}

// Bilinear samples of every item of one projection (host copy).
int item_sample_counts(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int groups_a, int groups_t, std::vector<float>& counts)
{
    const int per_quad = groups_a * groups_t;
    float* counts_d = nullptr;
    ECC_CUDA(ctx, cudaMalloc(&counts_d, sizeof(float) * per_quad));
    ECC_CUDA(ctx, cudaMemsetAsync(counts_d, 0, sizeof(float) * per_quad, ctx->stream));
    item_samples_kernel<<<dim3((n_alpha + 127) / 128, n_t), 128, 0, ctx->stream>>>(n_u, n_v, n_alpha, n_t, groups_a, counts_d);
    counts.assign(per_quad, 0.f);
    ECC_CUDA(ctx, cudaMemcpyAsync(counts.data(), counts_d, sizeof(float) * per_quad, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(counts_d);
    return ECC_OK;
}

// The window path's share of the samples when both pipes run side by side.  Built-in: what the run-time queue settles at on
// B200 at 1965 MHz (window path 0.936, texture path 0.677 projections per ms inside the combined kernel) -- 580 per mille
// with the general window configuration, 605 with the fine one.  The balance is a property of the GPU (ratio of
// shared-memory to texture throughput, clocks, SM count do not enter: both paths scale with them alike, but another
// architecture or a power-capped part may sit elsewhere): ecc_radon_calibrate_split measures it on the GPU at hand,
// ecc_radon_set_split pins it for a context (results then are reproducible for that number).  ECC_HYBRID4_SPLIT
// (per mille) is the development knob.
int static_split_items(ecc_context* ctx, Hybrid4Stage& H, int n_u, int n_v, int n_alpha, int n_t, int groups_a, int groups_t, int cfg,
                       int* split)
{
    if (H.split_items >= 0 && H.split_key[0] == n_u && H.split_key[1] == n_v && H.split_key[2] == n_alpha && H.split_key[3] == n_t &&
        H.split_cfg == cfg && H.split_share == H.share_override) {
        *split = H.split_items;
        return ECC_OK;
    }
    const int per_quad = groups_a * groups_t;
    std::vector<float> counts;
    const int rc = item_sample_counts(ctx, n_u, n_v, n_alpha, n_t, groups_a, groups_t, counts);
    if (rc) return rc;
    double total = 0;
    for (float c : counts) total += c;
    static const int share_env = env_int("ECC_HYBRID4_SPLIT", 0);
    const int permille = H.share_override > 0 ? H.share_override : (share_env > 0 ? share_env : (cfg == 1 ? 605 : 580));
    const double share = permille / 1000.0;
    std::vector<int> order(per_quad);
    for (int k = 0; k < per_quad; k++) order[k] = k;
    int m = -1;
    static const char* order_file = getenv("ECC_HYBRID4_ORDER");  // development: [int split][int order[per_quad]] from a file
    if (order_file) {
        if (FILE* f = fopen(order_file, "rb")) {
            int ms = 0;
            std::vector<int> o(per_quad);
            if (fread(&ms, sizeof(int), 1, f) == 1 && fread(o.data(), sizeof(int), per_quad, f) == (size_t)per_quad) {
                order = o;
                m = ms;
            }
            fclose(f);
        }
    }
    if (m < 0) {
        double run = 0;
        m = 0;
        while (m < per_quad && run + counts[order[m]] * 0.5 < share * total) run += counts[order[m++]];
    }
    if (H.order_len < per_quad) {
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (H.order_d) cudaFree(H.order_d);
        H.order_d = nullptr;
        H.order_len = 0;
        ECC_CUDA(ctx, cudaMalloc(&H.order_d, sizeof(int) * per_quad));
        H.order_len = per_quad;
    }
    ECC_CUDA(ctx, cudaMemcpyAsync(H.order_d, order.data(), sizeof(int) * per_quad, cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `order` is a local
    H.win_prefix.assign(1, 0.0);
    for (int k = 0; k < m; k++) H.win_prefix.push_back(H.win_prefix.back() + counts[order[k]]);
    H.tex_prefix.assign(1, 0.0);
    for (int k = m; k < per_quad; k++)
        for (int sub = 0; sub < kSubTiles; sub++) H.tex_prefix.push_back(H.tex_prefix.back() + counts[order[k]] / kSubTiles);
    H.split_items = m;
    H.split_key[0] = n_u; H.split_key[1] = n_v; H.split_key[2] = n_alpha; H.split_key[3] = n_t;
    H.split_cfg = cfg;
    H.split_share = H.share_override;
    *split = m;
    return ECC_OK;
}

int encode_map4(ecc_context* ctx, CUtensorMap* map, float4* base, int pitch, int rows, int count, int box_rows, int box_w)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(ctx, ECC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[4] = {4, (cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)count};
    const cuuint64_t strides[3] = {16, (cuuint64_t)pitch * 16u, (cuuint64_t)pitch * rows * 16u};
    const cuuint32_t box[4] = {4, (cuuint32_t)box_rows, (cuuint32_t)box_w, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, ECC_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (" + std::to_string((int)r) + ")");
    return ECC_OK;
}

}  // namespace

void free_hybrid4(ecc_context* ctx)
{
    Hybrid4Stage& H = ctx->hybrid4;
    for (auto t : H.tex_h) cudaDestroyTextureObject(t);
    for (auto f : H.surf_h) cudaDestroySurfaceObject(f);
    for (auto a : H.arrays) cudaFreeArray(a);
    if (H.surf_d) cudaFree(H.surf_d);
    if (H.tex_d) cudaFree(H.tex_d);
    if (H.pad_n) cudaFree(H.pad_n);
    if (H.pad_t) cudaFree(H.pad_t);
    if (H.queue) cudaFree(H.queue);
    if (H.order_d) cudaFree(H.order_d);
    const int keep = H.share_override;
    H = Hybrid4Stage();
    H.share_override = keep;
}

template <typename W>
int launch_window_config(ecc_context* ctx, Hybrid4Stage& H, const Hybrid4Params& P, int cfg, int threads, int ctas, int n_u, int n_v)
{
    if (H.map_cfg != cfg) {  // the TMA boxes have the window's shape
        int rc = encode_map4(ctx, (CUtensorMap*)H.map_n, (float4*)H.pad_n, n_u + 1, n_v + 1, H.quads, W::rows, W::boxw);
        if (rc) return rc;
        rc = encode_map4(ctx, (CUtensorMap*)H.map_t, (float4*)H.pad_t, n_v + 1, n_u + 1, H.quads, W::rows, W::boxw);
        if (rc) return rc;
        H.map_cfg = cfg;
    }
#ifdef ECC_CONFLICT_PROBE
    const size_t smem = W::smem + 1024;
#else
    const size_t smem = W::smem;
#endif
    const CUtensorMap& mn = *(const CUtensorMap*)H.map_n;
    const CUtensorMap& mt = *(const CUtensorMap*)H.map_t;
    if (threads <= 512 && ctas <= 2) {
        ECC_CUDA(ctx, cudaFuncSetAttribute(radon_hybrid4_kernel<512, 2, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radon_hybrid4_kernel<512, 2, W><<<ctx->sm_count * ctas, threads, smem, ctx->stream>>>(mn, mt, P);
    } else if (threads <= 384 && ctas == 3) {
        ECC_CUDA(ctx, cudaFuncSetAttribute(radon_hybrid4_kernel<384, 3, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radon_hybrid4_kernel<384, 3, W><<<ctx->sm_count * ctas, threads, smem, ctx->stream>>>(mn, mt, P);
    } else {
        return fail(ctx, ECC_ERR_INVALID, "ECC_HYBRID4_NT / ECC_HYBRID4_CTAS: unsupported combination");
    }
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

// Staging for nq quads of n_u x n_v images (grows, never shrinks; sized once per call chain by radon_hybrid4_reserve).
static int ensure_hybrid4(ecc_context* ctx, int n_u, int n_v, int nq)
{
    Hybrid4Stage& H = ctx->hybrid4;
    if (H.n_u == n_u && H.n_v == n_v && H.quads >= nq) return ECC_OK;
    const int quads = (H.n_u == n_u && H.n_v == n_v && H.quads > nq) ? H.quads : nq;
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the staging that is about to go
    free_hybrid4(ctx);
    H.n_u = n_u;
    H.n_v = n_v;
    const size_t pxp = (size_t)(n_u + 1) * (n_v + 1);
    ECC_CUDA(ctx, cudaMalloc(&H.pad_n, sizeof(float4) * pxp * quads));
    ECC_CUDA(ctx, cudaMalloc(&H.pad_t, sizeof(float4) * pxp * quads));
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    for (int k = 0; k < quads; k++) {
        cudaArray_t arr = nullptr;
        ECC_CUDA(ctx, cudaMallocArray(&arr, &desc, n_u, n_v, cudaArraySurfaceLoadStore));
        H.arrays.push_back(arr);
        cudaResourceDesc res = {};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arr;
        cudaSurfaceObject_t surf = 0;
        ECC_CUDA(ctx, cudaCreateSurfaceObject(&surf, &res));
        H.surf_h.push_back(surf);
        cudaTextureDesc td = {};
        td.normalizedCoords = 0;
        td.filterMode = cudaFilterModeLinear;
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t tex = 0;
        ECC_CUDA(ctx, cudaCreateTextureObject(&tex, &res, &td, nullptr));
        H.tex_h.push_back(tex);
        H.quads = k + 1;
    }
    ECC_CUDA(ctx, cudaMalloc(&H.tex_d, sizeof(cudaTextureObject_t) * quads));
    ECC_CUDA(ctx, cudaMemcpyAsync(H.tex_d, H.tex_h.data(), sizeof(cudaTextureObject_t) * quads, cudaMemcpyHostToDevice, ctx->stream));
    ECC_CUDA(ctx, cudaMalloc(&H.surf_d, sizeof(cudaSurfaceObject_t) * quads));
    ECC_CUDA(ctx, cudaMemcpyAsync(H.surf_d, H.surf_h.data(), sizeof(cudaSurfaceObject_t) * quads, cudaMemcpyHostToDevice, ctx->stream));
    H.map_cfg = -1;  // the tensor maps are encoded at the launch, for the window configuration chosen there
    return ECC_OK;
}

// Sizes the staging once for launches of up to n images (callers that feed a data set in growing chunks).
int radon_hybrid4_reserve(ecc_context* ctx, int n_u, int n_v, int n) { return ensure_hybrid4(ctx, n_u, n_v, (n + 3) / 4); }

// n images (device, dense) -> their Radon intermediates; works on ceil(n/4) quads.
int radon_hybrid4_launch(ecc_context* ctx, const float* images_d, int n, int n_u, int n_v, int n_alpha, int n_t, int post, float* out_d,
                         bool static_split, QuadPart part)
{
    if (!part.whole() && !static_split) return fail(ctx, ECC_ERR_INVALID, "parts of a quad need the static split");
    if (part.den < 1 || part.lo_num < 0 || part.lo_num >= part.den || part.hi_num < 1 || part.hi_num > part.den)
        return fail(ctx, ECC_ERR_INVALID, "radon_hybrid4_launch: bad part of a quad");
    Hybrid4Stage& H = ctx->hybrid4;
    const int nq = (n + 3) / 4;
    {
        const int rce = ensure_hybrid4(ctx, n_u, n_v, nq);
        if (rce) return rce;
    }
    {   // image staging: two kernels, profile family "stage"
        const int s1 = prof_begin(ctx, FAM_STAGE);
        interleave4_kernel<<<dim3((n_u + 1 + 31) / 32, (n_v + 1 + 3) / 4, nq), dim3(32, 4), 0, ctx->stream>>>(images_d, n, n_u, n_v, H.surf_d, (float4*)H.pad_n);
        prof_end(ctx, s1);
        const int s2 = prof_begin(ctx, FAM_STAGE);
        transpose4_kernel<<<dim3((n_u + 1 + 15) / 16, (n_v + 1 + 15) / 16, nq), dim3(16, 16), 0, ctx->stream>>>((const float4*)H.pad_n, n_u, n_v, (float4*)H.pad_t);
        prof_end(ctx, s2);
    }
    Hybrid4Params P;
    P.texs = H.tex_d;
    P.n_quads = nq;
    P.n_img = n;
    P.n_u = n_u;
    P.n_v = n_v;
    P.n_alpha = n_alpha;
    P.n_t = n_t;
    P.post = post;
    P.groups_a = (n_alpha + kItemAngles - 1) / kItemAngles;
    P.groups_t = (n_t + kItemT - 1) / kItemT;
    const size_t queue_words = 2 + (size_t)nq * P.groups_a * P.groups_t;
    if (H.queue_words < queue_words) {
        if (H.queue) cudaFree(H.queue);
        H.queue = nullptr;
        H.queue_words = 0;
        ECC_CUDA(ctx, cudaMalloc(&H.queue, sizeof(unsigned) * queue_words));
        H.queue_words = queue_words;
    }
    ECC_CUDA(ctx, cudaMemsetAsync(H.queue, 0, sizeof(unsigned) * queue_words, ctx->stream));
    P.counters = H.queue;
    P.claim = H.queue + 2;
    P.magic = 0x4B0000u;
    P.out = out_d;
    // window configuration: a band of 32 t bins must fit the window height (plus the spread of the item's 8 angles)
    static const int cfg_env = env_int("ECC_HYBRID4_CFG", -1);
    const double t_spacing = std::sqrt((double)n_u * n_u + (double)n_v * n_v) / n_t;
    const int cfg = cfg_env >= 0 ? cfg_env : (t_spacing <= 2.1 ? 1 : 0);
    P.split_items = -1;
    P.order = nullptr;
    P.w_begin = P.w_end = P.x_begin = P.x_end = 0;
    P.w_last = 0;
    if (static_split) {
        const int rcs = static_split_items(ctx, H, n_u, n_v, n_alpha, n_t, P.groups_a, P.groups_t, cfg, &P.split_items);
        if (rcs) return rcs;
        P.order = H.order_d;
        // the launch's share of the item lists: same floor() on both sides of a cut, so two ranks that share a quad
        // (one with hi = k / den, the other with lo = k / den) take complementary entries
        const long long split = P.split_items, subs = (long long)(P.groups_a * P.groups_t - P.split_items) * kSubTiles;
        if (nq == 1 && part.lo_num >= part.hi_num) return fail(ctx, ECC_ERR_INVALID, "radon_hybrid4_launch: empty part of a quad");
        // A part of a quad takes the same share of the SAMPLES of both lists (the items of a list differ in cost by the length
        // of their lines: cutting by entries gave one rank the short lines of the window list and the long ones of the texture
        // list -- its two pipes out of balance, half of what sharing the quad should have saved was lost).  The cut is a
        // function of the fraction alone, so the two ranks that share a quad take complementary entries.
        auto cut = [&](const std::vector<double>& prefix, int num) -> long long {
            const long long len = (long long)prefix.size() - 1;
            if (num <= 0) return 0;
            if (num >= part.den) return len;
            const double want = prefix.back() * (double)num / (double)part.den;
            return (long long)(std::lower_bound(prefix.begin(), prefix.end(), want) - prefix.begin());
        };
        P.w_begin = (unsigned)cut(H.win_prefix, part.lo_num);
        P.w_end = (unsigned)((long long)(nq - 1) * split + cut(H.win_prefix, part.hi_num));
        static const int keep_order = env_int("ECC_HYBRID4_KEEP_ORDER", 0);  // development: the last quad in list order, too
        const long long last_begin = (long long)(nq - 1) * split;
        P.w_last = keep_order ? P.w_end : (unsigned)std::max<long long>(P.w_begin, last_begin);
        P.x_begin = (unsigned)cut(H.tex_prefix, part.lo_num);
        P.x_end = (unsigned)((long long)(nq - 1) * subs + cut(H.tex_prefix, part.hi_num));
    }
    static const int ag_lo = env_int("ECC_ITEM_AG_LO", 0), ag_hi = env_int("ECC_ITEM_AG_HI", INT_MAX);
    P.ag_lo = ag_lo;
    P.ag_hi = ag_hi;
    P.mir = team_mirrors(ctx, out_d);
    static const char* clock_file = getenv("ECC_ITEM_CLOCK");
    P.item_clock = nullptr;
    const size_t clock_words = 2 * (size_t)P.groups_a * P.groups_t;
    if (clock_file) {
        ECC_CUDA(ctx, cudaMalloc(&P.item_clock, sizeof(unsigned long long) * clock_words));
        ECC_CUDA(ctx, cudaMemsetAsync(P.item_clock, 0, sizeof(unsigned long long) * clock_words, ctx->stream));
    }
    static const int nt = env_int("ECC_HYBRID4_NT", 8);
    static const int ctas = env_int("ECC_HYBRID4_CTAS", 2);
    static const int mode = env_int("ECC_HYBRID_MODE", 0);
    P.mode = mode;
    static const int lane_map = env_int("ECC_HYBRID4_LANEMAP", 2);
    P.lane_map = lane_map;
    static const int tex_map = env_int("ECC_HYBRID4_TEXMAP", 0);
    P.tex_map = tex_map;
    const int threads = (kWindowWarps + nt) * 32;
    const int slot = prof_begin(ctx, FAM_RADON);
    const int rcl = cfg == 1 ? launch_window_config<Win4Fine>(ctx, H, P, cfg, threads, ctas, n_u, n_v)
                             : launch_window_config<Win4General>(ctx, H, P, cfg, threads, ctas, n_u, n_v);
    prof_end(ctx, slot);
    if (clock_file) {  // development: the per-item cycle counts of this launch, overwritten by every launch
        std::vector<unsigned long long> h(clock_words);
        cudaMemcpyAsync(h.data(), P.item_clock, sizeof(unsigned long long) * clock_words, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(P.item_clock);
        if (FILE* f = fopen(clock_file, "wb")) {
            const int hdr[4] = {P.groups_a, P.groups_t, nq, 0};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h.data(), sizeof(unsigned long long), clock_words, f);
            fclose(f);
        }
    }
    return rcl;
}

// Runs the RUN-TIME queue (which balances the two pipes by itself) on one quad of this geometry and reads off where the two
// sides met: the share of the samples the window path took.  Median of `repeats` launches, per mille.
int radon_hybrid4_calibrate(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int repeats, int* permille)
{
    Hybrid4Stage& H = ctx->hybrid4;
    const size_t px = (size_t)n_u * n_v, bins = (size_t)n_alpha * n_t;
    float *img_d = nullptr, *out_d = nullptr;
    ECC_CUDA(ctx, cudaMalloc(&img_d, sizeof(float) * px * 4));
    ECC_CUDA(ctx, cudaMalloc(&out_d, sizeof(float) * bins * 4));
    ECC_CUDA(ctx, cudaMemsetAsync(img_d, 0, sizeof(float) * px * 4, ctx->stream));  // the work does not depend on the pixel values
    const int groups_a = (n_alpha + kItemAngles - 1) / kItemAngles, groups_t = (n_t + kItemT - 1) / kItemT;
    const int per_quad = groups_a * groups_t;
    std::vector<float> counts;
    int rc = item_sample_counts(ctx, n_u, n_v, n_alpha, n_t, groups_a, groups_t, counts);
    double total = 0;
    for (float c : counts) total += c;
    std::vector<int> shares;
    std::vector<unsigned> claim(per_quad);
    for (int r = 0; r < (repeats > 0 ? repeats : 1) + 1 && rc == ECC_OK; r++) {
        rc = radon_hybrid4_launch(ctx, img_d, 4, n_u, n_v, n_alpha, n_t, ECC_POST_IDENTITY, out_d, false);
        if (rc) break;
        if (cudaMemcpyAsync(claim.data(), H.queue + 2, sizeof(unsigned) * per_quad, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = fail(ctx, ECC_ERR_CUDA, "radon_hybrid4_calibrate: reading the queue back failed");
            break;
        }
        if (r == 0) continue;  // warm-up launch (tensor maps, staging)
        double window = 0;
        for (int k = 0; k < per_quad; k++)
            if (claim[k] == kClaimWindow) window += counts[k];
        shares.push_back(total > 0 ? (int)(1000.0 * window / total + 0.5) : 0);
    }
    cudaFree(img_d);
    cudaFree(out_d);
    if (rc) return rc;
    std::sort(shares.begin(), shares.end());
    *permille = shares.empty() ? 0 : shares[shares.size() / 2];
    return ECC_OK;
}

}  // namespace eccb200
