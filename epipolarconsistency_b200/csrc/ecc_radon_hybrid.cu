// ecc_radon_hybrid.cu -- Radon-intermediate kernel that feeds BOTH sampling pipes of an SM (sm_100a).
//
// WHAT (reference, code/LibEpipolarConsistency/RadonIntermediate.cu:31-143): per (alpha,t) bin the difference of two
// line integrals one pixel apart, sample step 0.66 px, bilinear image lookups through the texture unit.
// WHY a second path: the texture-only kernel (ecc_radon.cu) sits at 98 % of the L1TEX texture data pipe with 86 % of
// the issue slots idle (profiles/ncu_radon_r01.txt).  The texture filter's arithmetic is known exactly
// (profiles/tex_probe_r01.txt: position rounded to 1/256, w11 = (a*b+128)>>8, ...), so the idle issue slots can
// take the same samples from shared memory (profiles/radon_lsu_probe_r01.txt: within 7e-6 of the peak of the
// texture result, 1.66x the texture-only rate with both running side by side).
// HOW:
//   * persistent CTAs of 8 "window" warps + NT texture warps; work = items of 8 angles x 32 t bins per image, drawn
//     from a two-ended queue: window groups take items from the front, texture warps take 32-bin sub-tiles from the
//     back, so the split between the two pipes balances itself.
//   * a window group walks its 256 lines (two per bin) in lock step through 28-pixel chunks of the image along the
//     lines' primary axis; the band of rows all lines need inside a chunk is fetched by TMA (cp.async.bulk.tensor,
//     one 200-row x 36-column box, zero fill outside) into shared memory, double buffered behind an mbarrier.  Near-
//     vertical angles read a transposed copy of the image, so the code path is the same.
//   * sample positions are the reference's, bit for bit (clipped entry, t += 0.66f in fp32, +-1/2 px lines).
#include <climits>
#include <cstdlib>

#include "ecc_radon_common.cuh"

namespace eccb200 {

namespace {

constexpr int kChunk = 28;           // pixels of the primary axis per chunk
constexpr int kBoxW = 36;            // window columns [28j-4, 28j+32): chunk + 2 either side, start a multiple of 4
                                     // (TMA: the innermost start coordinate must be 16-byte aligned, tools/tma_probe.cu)
constexpr int kBoxLead = 4;          // columns in front of the chunk
constexpr int kRows = 200;           // window rows (secondary axis) = the contiguous axis of the shared-memory window;
                                     // 200 = 8 mod 32 words keeps neighbouring columns on different banks
constexpr int kMaxChunks = 64;       // primary axis up to ~1790 px; longer -> the item goes through the texture unit

// The whole bin through the texture unit, in the executed reference's shape (see above); sums in sample order.
__device__ __forceinline__ float bin_texture(cudaTextureObject_t tex, const BinLine& L)
{
    float o0 = L.o0 + 0.5f, o1 = L.o1 + 0.5f;  // texel centres
    const float d0 = L.d0, d1 = L.d1, t_max = L.t_max;
    o0 -= 0.5f * d1;
    o1 += 0.5f * d0;
    float sum = 0.f, sumo = 0.f;
    float t = L.t;
    bool none_yet = true;
    // explicit fma: the executed reference contracts o + t*d into one FFMA and adds the line offset afterwards; left to
    // the compiler, a product t*d0 that happens to be at hand from the entry-point test is reused uncontracted
#define ECC_TEX_A(tt) tex2D<float>(tex, fmaf((tt), d0, o0), fmaf((tt), d1, o1))
#define ECC_TEX_B(tt) tex2D<float>(tex, fmaf((tt), d0, o0) + d1, fmaf((tt), d1, o1) - d0)
    if (!(t + 1.98f > t_max)) {
        const float r3 = t_max - 1.98f;
        // software pipeline: the fetches of the next block of four are in flight while this block's results are added
        // (the texture warps are few; what keeps the texture pipe busy is the number of fetches they have in flight)
        float t1 = t + kStep, t2 = t1 + kStep, t3 = t2 + kStep;
        float a0 = ECC_TEX_A(t), b0 = ECC_TEX_B(t), a1 = ECC_TEX_A(t1), b1 = ECC_TEX_B(t1);
        float a2 = ECC_TEX_A(t2), b2 = ECC_TEX_B(t2), a3 = ECC_TEX_A(t3), b3 = ECC_TEX_B(t3);
        t = t3 + kStep;
#pragma unroll 1
        while (!(t > r3)) {
            t1 = t + kStep; t2 = t1 + kStep; t3 = t2 + kStep;
            const float na0 = ECC_TEX_A(t), nb0 = ECC_TEX_B(t), na1 = ECC_TEX_A(t1), nb1 = ECC_TEX_B(t1);
            const float na2 = ECC_TEX_A(t2), nb2 = ECC_TEX_B(t2), na3 = ECC_TEX_A(t3), nb3 = ECC_TEX_B(t3);
            sum += a0; sumo += b0;
            sum += a1; sumo += b1;
            sum += a2; sumo += b2;
            sum += a3; sumo += b3;
            a0 = na0; b0 = nb0; a1 = na1; b1 = nb1; a2 = na2; b2 = nb2; a3 = na3; b3 = nb3;
            t = t3 + kStep;
        }
        sum += a0; sumo += b0;
        sum += a1; sumo += b1;
        sum += a2; sumo += b2;
        sum += a3; sumo += b3;
        none_yet = false;
    }
    const float t1 = t + kStep;
    if (!(t1 > t_max)) {
        const float a0 = ECC_TEX_A(t), b0 = ECC_TEX_B(t), a1 = ECC_TEX_A(t1), b1 = ECC_TEX_B(t1);
        sum += a0; sumo += b0;
        sum += a1; sumo += b1;
        t = t1 + kStep;
        none_yet = false;
    }
    if (t <= t_max || none_yet) {
        sum += ECC_TEX_A(t);
        sumo += ECC_TEX_B(t);
    }
#undef ECC_TEX_A
#undef ECC_TEX_B
    return (sum - sumo) * kStep;
}

// One bilinear sample from the shared-memory window, reproducing the texture unit (profiles/tex_probe_r01.txt):
// (coordinate - 1/2) rounded to 1/256 half up -> cell index and 8-bit fractions a, b; weights
// w11 = (a*b + 128) >> 8, w10 = a - w11, w01 = b - w11, w00 = 256 - a - b + w11, all /256.
// pri/sec: texel-space coordinates along the window's column / row axis.  base: byte address (shared window) such
// that cell (ipri, isec) sits at base + 4*((ipri + M)*kRows + isec + M) with M = 0x4B0000 (the exponent bits of the
// rounding constant are left in the index and folded into base).  The secondary axis is the contiguous one: the lanes
// of a warp are parallel lines 2..3 rows apart at (nearly) the same column, which spreads them over the banks.
__device__ __forceinline__ float sample_window(unsigned base, float pri, float sec)
{
    const float Ps = fmaf(pri, 256.f, -127.5f);  // exact: 256*(pri - 1/2) + 1/2
    const float Ss = fmaf(sec, 256.f, -127.5f);
    const unsigned Pi = __float_as_uint(__fadd_rd(Ps, 8388608.f));  // 0x4B000000 + floor(Ps)
    const unsigned Si = __float_as_uint(__fadd_rd(Ss, 8388608.f));
    const unsigned a = Pi & 255u, b = Si & 255u;
    const unsigned addr = base + (((Pi >> 8) * kRows + (Si >> 8)) << 2);
    float v00, v10, v01, v11;
    asm volatile(
        "ld.shared.f32 %0, [%4];\n"
        "ld.shared.f32 %1, [%4+4];\n"
        "ld.shared.f32 %2, [%4+800];\n"
        "ld.shared.f32 %3, [%4+804];\n"
        : "=f"(v00), "=f"(v01), "=f"(v10), "=f"(v11)
        : "r"(addr));
    const unsigned w11 = (a * b + 128u) >> 8;
    const unsigned w10 = a - w11, w01 = b - w11;  // w00 = 256 - w10 - w01 - w11
    // The texture unit rounds ONCE: its result is the correctly rounded exact sum in 99.7 % of the samples
    // (profiles/tex_round_probe_r01.txt), whereas sum_i w_i v_i in fp32 is up to 1.8 ulp off.  Written about v00,
    //     v00 + (w10 (v10-v00) + w01 (v01-v00) + w11 (v11-v00)) / 256,
    // the differences of neighbouring texels are exact or small, the weighted sum of them carries errors relative to
    // the differences, and the only rounding at the scale of the texels is the last one -- at the same instruction count.
    const float d10 = v10 - v00, d01 = v01 - v00, d11 = v11 - v00;
    float t = __uint2float_rn(w10) * d10;
    t = fmaf(__uint2float_rn(w01), d01, t);
    t = fmaf(__uint2float_rn(w11), d11, t);
    return fmaf(t, 0.00390625f, v00);
}
static_assert(kRows * 4 == 800, "sample_window hard-codes the column pitch");

struct HybridParams {
    const cudaTextureObject_t* texs;
    int n_img, n_u, n_v, n_alpha, n_t, post;
    int groups_a, groups_t;  // items per image = groups_a * groups_t
    int nbuf;                // window buffers (1 or 2)
    int mode;                // development: 0 both paths, 1 texture warps only, 2 window warps only
    unsigned* counters;      // [2], zeroed before the launch
    unsigned* claim;         // [items], zeroed before the launch
    unsigned magic;          // 0x4B0000, kept out of the compiler's sight (see sample_window)
    float* out;
    Mirrors mir;             // multi-GPU team: every bin is also stored into the other ranks' buffers
};

__device__ __forceinline__ void store_bin(const HybridParams& p, size_t index, float val)
{
    float* d = p.out + index;
    *d = val;
    for (int r = 0; r < p.mir.n; r++) *(float*)((char*)d + p.mir.delta[r]) = val;
}

struct ItemBins {
    int img, ix, iy;
};
// item id -> image, and this warp-in-item's bin for the lane (quad tiling: 2 angles x 2 t per quad, 16 t per warp)
__device__ __forceinline__ ItemBins item_bins(int item, int wi, int lane, const HybridParams& p)
{
    ItemBins r;
    const int per_img = p.groups_a * p.groups_t;
    r.img = item / per_img;
    const int rem = item - r.img * per_img;
    const int tg = rem / p.groups_a, ag = rem - tg * p.groups_a;
    const int quad = lane >> 2, ql = lane & 3;
    r.ix = ag * kItemAngles + (wi & 3) * 2 + (ql & 1);
    r.iy = tg * kItemT + (wi >> 2) * 16 + quad * 2 + (ql >> 1);
    return r;
}

template <int MAXTHREADS, int MINBLOCKS>
__global__ void __launch_bounds__(MAXTHREADS, MINBLOCKS)
radon_hybrid_kernel(const __grid_constant__ CUtensorMap map_n, const __grid_constant__ CUtensorMap map_t,
                    const __grid_constant__ HybridParams p)
{
    extern __shared__ __align__(128) unsigned char window_raw[];
    __shared__ __align__(8) unsigned long long mbar_store[2];
    __shared__ int s_item, s_jmin, s_jmax, s_fallback;
    __shared__ int win_lo[kMaxChunks], win_hi[kMaxChunks];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned total_items = (unsigned)(p.n_img * p.groups_a * p.groups_t);
    const float n_u = (float)p.n_u, n_v = (float)p.n_v;
    const size_t img_stride = (size_t)p.n_t * p.n_alpha;

    if (tid == 0) {
        mbar_init(smem_u32(&mbar_store[0]), 1);
        mbar_init(smem_u32(&mbar_store[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kWindowWarps) {
        if (p.mode == 2) return;
        // ---------------- texture warps: 32-bin sub-tiles from the back of the queue ----------------
        for (;;) {
            int sub = 0;
            if (lane == 0) sub = take_back(p.counters, p.claim, total_items);
            sub = __shfl_sync(0xffffffffu, sub, 0);
            if (sub < 0) break;
            const int item = (int)total_items - 1 - sub / kSubTiles;
            const ItemBins B = item_bins(item, sub % kSubTiles, lane, p);
            if (B.ix >= p.n_alpha || B.iy >= p.n_t) continue;
            const BinLine L = bin_line(B.ix, B.iy, p.n_alpha, p.n_t, n_u, n_v);
            float r = 0.f;
            if (L.valid) r = post_process(bin_texture(p.texs[B.img], L), p.post);
            store_bin(p, B.img * img_stride + (size_t)B.iy * p.n_alpha + B.ix, r);
        }
        return;
    }

    // ---------------- window warps: whole items from the front of the queue ----------------
    if (p.mode == 1) return;
    const unsigned window_base = smem_u32(window_raw);
    const unsigned buf_bytes = (unsigned)kRows * kBoxW * 4u;
    const unsigned mbar0 = smem_u32(&mbar_store[0]);
    unsigned phase0 = 0, phase1 = 0;
    for (;;) {
        if (tid == 0) {
            s_item = take_front(p.counters, p.claim, total_items);
            s_jmin = INT_MAX;
            s_jmax = INT_MIN;
            s_fallback = 0;
        }
        if (tid < kMaxChunks) { win_lo[tid] = INT_MAX; win_hi[tid] = INT_MIN; }
        group_sync();
        const int item = s_item;
        if (item < 0) break;
        const ItemBins B = item_bins(item, warp, lane, p);
        const bool in_range = B.ix < p.n_alpha && B.iy < p.n_t;
        BinLine L;
        L.valid = false;
        L.swapped = false;
        if (in_range) L = bin_line(B.ix, B.iy, p.n_alpha, p.n_t, n_u, n_v);
        // the window path has no clamp addressing: it takes the lines that stay inside the inset box (all but a few at
        // the image border: lines that miss the box, and axis-parallel lines, which the reference does not clip)
        bool safe = false;
        if (in_range && L.valid) {
            const float ax = L.o0 + L.t * L.d0, ay = L.o1 + L.t * L.d1, bx = L.o0 + L.t_max * L.d0, by = L.o1 + L.t_max * L.d1;
            const float lo = 1.f - 1e-3f, hx = n_u - 1.f + 1e-3f, hy = n_v - 1.f + 1e-3f;
            safe = !L.swapped && fminf(ax, bx) >= lo && fmaxf(ax, bx) <= hx && fminf(ay, by) >= lo && fmaxf(ay, by) <= hy;
        }
        const bool live = safe;

        // orientation of the item: primary axis = the one the lines advance along fastest (from the item's middle angle)
        const int per_img = p.groups_a * p.groups_t;
        const int ag = (item % per_img) % p.groups_a;
        const float alpha_mid = ((ag * kItemAngles + 0.5f * (kItemAngles - 1)) / (float)p.n_alpha - 0.5f) * ECC_PI_F;
        const bool vertical = fabsf(sinf(alpha_mid)) > fabsf(cosf(alpha_mid));
        const int dir = (vertical && alpha_mid < 0.f) ? -1 : 1;

        // line A of the bin in texel space, exactly as the reference forms it; line B = A + (d1, -d0)
        float oA0 = L.o0 + 0.5f, oA1 = L.o1 + 0.5f;
        oA0 -= 0.5f * L.d1;
        oA1 += 0.5f * L.d0;
        const float op = vertical ? oA1 : oA0, dp = vertical ? L.d1 : L.d0;
        const float os = vertical ? oA0 : oA1, ds = vertical ? L.d0 : L.d1;
        const float offp = vertical ? -L.d0 : L.d1, offs = vertical ? L.d1 : -L.d0;
        const float inv_dp = live ? 1.f / dp : 0.f;

        // chunks this line touches (by line A's primary pixel coordinate), one spare either side
        int jmin_l = INT_MAX, jmax_l = INT_MIN;
        if (live) {
            const float pa = fmaf(L.t, dp, op) - 0.5f, pb = fmaf(L.t_max, dp, op) - 0.5f;
            const int ja = (int)floorf(fminf(pa, pb) * (1.f / kChunk)), jb = (int)floorf(fmaxf(pa, pb) * (1.f / kChunk));
            jmin_l = ja - 1;
            jmax_l = jb + 1;
        }
        {
            const int wmin = __reduce_min_sync(0xffffffffu, jmin_l), wmax = __reduce_max_sync(0xffffffffu, jmax_l);
            if (lane == 0 && wmin <= wmax) { atomicMin(&s_jmin, wmin + 1); atomicMax(&s_jmax, wmax - 1); }
        }
        group_sync();
        const int jlo = s_jmin, jhi = s_jmax;
        const int n_chunks = jhi - jlo + 1;  // <= 0: no live line in this item
        const bool too_long = n_chunks > kMaxChunks;

        // row window of every chunk: rows (secondary axis) any of the group's lines needs there
        if (n_chunks > 0 && !too_long) {
            for (int k = 0; k < n_chunks; k++) {
                const int j = jlo + k;
                int lo = INT_MAX, hi = INT_MIN;
                if (live && j >= jmin_l && j <= jmax_l) {
                    const float t0 = ((float)(j * kChunk - 2) + 0.5f - op) * inv_dp;
                    const float t1 = ((float)(j * kChunk + kChunk + 2) + 0.5f - op) * inv_dp;
                    const float s0 = fmaf(fmaxf(fminf(t0, t1), L.t - 2.f), ds, os);
                    const float s1 = fmaf(fminf(fmaxf(t0, t1), L.t_max + 2.f), ds, os);
                    // line B is within one pixel of line A; a sample at s reads rows floor(s - 1/2) and the next
                    lo = (int)floorf(fminf(s0, s1) - 2.0f);
                    hi = (int)floorf(fmaxf(s0, s1) + 1.0f) + 1;
                }
                const int wlo = __reduce_min_sync(0xffffffffu, lo), whi = __reduce_max_sync(0xffffffffu, hi);
                if (lane == 0 && wlo <= whi) { atomicMin(&win_lo[k], wlo); atomicMax(&win_hi[k], whi); }
            }
        }
        group_sync();
        if (n_chunks > 0 && !too_long && tid < n_chunks) {
            // the window's first row must be a multiple of 4 (TMA: 16-byte aligned start along the contiguous axis)
            if (win_lo[tid] <= win_hi[tid]) {
                win_lo[tid] &= ~3;
                if (win_hi[tid] - win_lo[tid] + 1 > kRows) s_fallback = 1;
            }
        }
        if (tid == 0 && too_long) s_fallback = 1;
        group_sync();
        const bool fallback = s_fallback != 0;

        float result = 0.f;
        if (n_chunks > 0 && !fallback) {
            // the copy whose contiguous axis is the window's secondary axis
            const CUtensorMap* map = vertical ? &map_n : &map_t;
            float t = live ? L.t : 3.0e38f;
            // stop at the last sample the executed reference takes (an ulp past t_max at times, see ref_last_sample)
            const float t_max = live ? ref_last_sample(L.t, L.t_max) : -3.0e38f;
            float sum = 0.f, sumo = 0.f;
            // issue the loads of chunk k (traversal order) into buffer k % nbuf
            auto issue = [&](int k) {
                const int kk = dir > 0 ? k : n_chunks - 1 - k;
                const int lo = win_lo[kk], hi = win_hi[kk];
                if (lo > hi) return;
                const unsigned b = (p.nbuf == 2) ? (unsigned)(k & 1) : 0u;
                const unsigned mb = mbar0 + 8u * b;
                mbar_expect_tx(mb, buf_bytes);
                tma_load_3d(window_base + b * buf_bytes, map, lo, (jlo + kk) * kChunk - kBoxLead, B.img, mb);
            };
            if (p.nbuf == 2 && tid == 0) issue(0);
            for (int k = 0; k < n_chunks; k++) {
                const int kk = dir > 0 ? k : n_chunks - 1 - k;
                const int j = jlo + kk;
                const int lo = win_lo[kk], hi = win_hi[kk];
                if (p.nbuf == 2) {
                    group_sync();  // everybody is done with the buffer chunk k+1 will overwrite
                    if (tid == 0 && k + 1 < n_chunks) issue(k + 1);
                } else {
                    group_sync();
                    if (tid == 0) issue(k);
                }
                if (lo > hi) continue;  // nobody samples in this chunk
                const unsigned b = (p.nbuf == 2) ? (unsigned)(k & 1) : 0u;
                if (b == 0) { mbar_wait(mbar0, phase0); phase0 ^= 1u; }
                else { mbar_wait(mbar0 + 8u, phase1); phase1 ^= 1u; }
                // this chunk's samples: t up to where line A leaves the chunk (the last chunk takes the rest)
                float lim = t_max;
                if (k + 1 < n_chunks) {
                    const float edge = (float)(dir > 0 ? (j + 1) * kChunk : j * kChunk) + 0.5f;
                    lim = fminf(t_max, (edge - op) * inv_dp);
                }
                const unsigned base = window_base + b * buf_bytes -
                                      4u * ((p.magic + (unsigned)(j * kChunk - kBoxLead)) * kRows + p.magic + (unsigned)lo);
#pragma unroll 1  // one test per sample: the compiler's own unrolling of such a loop tests once per block (see ref_last_sample)
                for (; t <= lim; t += kStep) {
                    const float pri = fmaf(t, dp, op), sec = fmaf(t, ds, os);
                    sum += sample_window(base, pri, sec);
                    sumo += sample_window(base, pri + offp, sec + offs);
                }
            }
            result = (sum - sumo) * kStep;
        }
        if (in_range) {
            // the (rare) lines outside the window path: items too long / too tall for the window, and lines that miss
            // the inset box (they may sample outside the image: clamp addressing)
            if (L.valid && (fallback || !safe)) result = bin_texture(p.texs[B.img], L);
            store_bin(p, B.img * img_stride + (size_t)B.iy * p.n_alpha + B.ix, L.valid ? post_process(result, p.post) : 0.f);
        }
        group_sync();  // s_item / tables are rewritten at the top
    }
}

// ---- image staging for the window path ------------------------------------------------------------------------------
// padded copy [n][n_v+1][pitch_n] (pixel (x,y) at [y][x]) and transposed padded copy [n][n_u+1][pitch_t] ([x][y]);
// entries past the last row / column replicate the edge (clamp addressing of the reference's texture).
__global__ void pad_kernel(const float* __restrict__ src, int n_u, int n_v, int pitch, float* __restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= pitch) return;
    const float* s = src + (size_t)blockIdx.z * n_u * n_v;
    dst[((size_t)blockIdx.z * (n_v + 1) + y) * pitch + x] = s[(size_t)min(y, n_v - 1) * n_u + min(x, n_u - 1)];
}
__global__ void pad_transpose_kernel(const float* __restrict__ src, int n_u, int n_v, int pitch, float* __restrict__ dst)
{
    __shared__ float tile[32][33];
    const float* s = src + (size_t)blockIdx.z * n_u * n_v;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y)
        tile[r][threadIdx.x] = s[(size_t)min(y0 + r, n_v - 1) * n_u + min(x0 + threadIdx.x, n_u - 1)];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int x = x0 + r, y = y0 + threadIdx.x;  // output row = x, column = y
        if (x <= n_u && y < pitch) dst[((size_t)blockIdx.z * (n_u + 1) + x) * pitch + y] = tile[threadIdx.x][r];
    }
}

int encode_map(ecc_context* ctx, CUtensorMap* map, float* base, int pitch, int rows, int count)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(ctx, ECC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)count};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4u, (cuuint64_t)pitch * rows * 4u};
    const cuuint32_t box[3] = {kRows, kBoxW, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, ECC_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return ECC_OK;
}

}  // namespace

void free_hybrid(ecc_context* ctx)
{
    HybridStage& H = ctx->hybrid;
    if (H.pad_n) cudaFree(H.pad_n);
    if (H.pad_t) cudaFree(H.pad_t);
    if (H.queue) cudaFree(H.queue);
    H = HybridStage();
}

// One launch over n images whose texture objects are texs_d[0..n) and whose pixels are images_d (dense, device).
int radon_hybrid_launch(ecc_context* ctx, const cudaTextureObject_t* texs_d, const float* images_d, int n, int n_u,
                        int n_v, int n_alpha, int n_t, int post, float* out_d)
{
    HybridStage& H = ctx->hybrid;
    const int pitch_n = (n_u + 1 + 3) & ~3, pitch_t = (n_v + 1 + 3) & ~3;
    if (H.n_u != n_u || H.n_v != n_v || H.count < n) {
        const int count = n > H.count ? n : H.count;
        ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the staging that is about to go
        free_hybrid(ctx);
        ECC_CUDA(ctx, cudaMalloc(&H.pad_n, sizeof(float) * (size_t)count * (n_v + 1) * pitch_n));
        ECC_CUDA(ctx, cudaMalloc(&H.pad_t, sizeof(float) * (size_t)count * (n_u + 1) * pitch_t));
        H.n_u = n_u;
        H.n_v = n_v;
        H.count = count;
        static_assert(sizeof(CUtensorMap) == sizeof(H.map_n), "tensor map storage");
        int rc = encode_map(ctx, (CUtensorMap*)H.map_n, H.pad_n, pitch_n, n_v + 1, count);
        if (rc) return rc;
        rc = encode_map(ctx, (CUtensorMap*)H.map_t, H.pad_t, pitch_t, n_u + 1, count);
        if (rc) return rc;
    }
    pad_kernel<<<dim3((pitch_n + 127) / 128, n_v + 1, n), 128, 0, ctx->stream>>>(images_d, n_u, n_v, pitch_n, H.pad_n);
    pad_transpose_kernel<<<dim3((n_u + 1 + 31) / 32, (pitch_t + 31) / 32, n), dim3(32, 8), 0, ctx->stream>>>(images_d, n_u, n_v, pitch_t, H.pad_t);

    HybridParams P;
    P.texs = texs_d;
    P.n_img = n;
    P.n_u = n_u;
    P.n_v = n_v;
    P.n_alpha = n_alpha;
    P.n_t = n_t;
    P.post = post;
    P.groups_a = (n_alpha + kItemAngles - 1) / kItemAngles;
    P.groups_t = (n_t + kItemT - 1) / kItemT;
    // development knobs (environment): texture warps per CTA, window rows, buffers, CTAs per SM
    static const int nt = env_int("ECC_HYBRID_NT", 8);
    static const int nbuf = env_int("ECC_HYBRID_NBUF", 1) == 2 ? 2 : 1;
    static const int ctas = env_int("ECC_HYBRID_CTAS", 3);
    static const int mode = env_int("ECC_HYBRID_MODE", 0);
    P.mode = mode;
    P.nbuf = nbuf;
    const size_t queue_words = 2 + (size_t)n * P.groups_a * P.groups_t;
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
    P.mir = team_mirrors(ctx, out_d);
    const int threads = (kWindowWarps + nt) * 32;
    const size_t smem = (size_t)kRows * kBoxW * 4 * nbuf;
    const CUtensorMap& mn = *(const CUtensorMap*)H.map_n;
    const CUtensorMap& mt = *(const CUtensorMap*)H.map_t;
    const int slot = prof_begin(ctx, FAM_RADON);
    if (threads <= 384 && ctas >= 4) {
        ECC_CUDA(ctx, cudaFuncSetAttribute(radon_hybrid_kernel<384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radon_hybrid_kernel<384, 4><<<ctx->sm_count * ctas, threads, smem, ctx->stream>>>(mn, mt, P);
    } else if (threads <= 512 && ctas >= 3) {
        ECC_CUDA(ctx, cudaFuncSetAttribute(radon_hybrid_kernel<512, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radon_hybrid_kernel<512, 3><<<ctx->sm_count * ctas, threads, smem, ctx->stream>>>(mn, mt, P);
    } else {
        if (threads > 768) return fail(ctx, ECC_ERR_INVALID, "ECC_HYBRID_NT too large");
        ECC_CUDA(ctx, cudaFuncSetAttribute(radon_hybrid_kernel<768, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        radon_hybrid_kernel<768, 2><<<ctx->sm_count * ctas, threads, smem, ctx->stream>>>(mn, mt, P);
    }
    prof_end(ctx, slot);
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

}  // namespace eccb200
