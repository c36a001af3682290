// ecc_team.cu -- the hot path over the GPUs of one node WITHOUT collectives on the data path (sm_100a, NVLink/NVSwitch).
//
// the path has one real exchange step (every GPU needs all Radon intermediates before it scores its
// share of the pairs) and a final reduction.  Both used to be NCCL calls after the kernels (all-gather of 146 MB per
// rank at C3 / 8 GPUs, two all-reduces).  Here every rank owns one device block
//        [ flags | pair values, 2 x n(n-1)/2 floats (used alternately) | Radon intermediates of ALL projections ]
// mapped into every other rank's address space (CUDA IPC between the per-GPU processes), and
//   * the Radon kernels store each finished bin into ALL blocks (one local + world-1 NVLink peer stores per bin; a bin
//     costs ~2300 samples, so the stores are free and the exchange rides under the line integration tile by tile),
//   * a flag barrier in peer memory (one release store per peer, acquire spins on the own flags) replaces the collective's
//     synchronisation,
//   * each rank scores its equal-work range of pairs and publishes the values the same way; after a second barrier every
//     rank holds all pair values and sums them in the single-GPU order: the multi-GPU mean and cost image are the same
//     bits on every rank, and the same bits a single GPU produces from the same intermediates.
// The reference is single-GPU (EpipolarConsistencyRadonIntermediate.cpp:166-225 runs everything on one device).
#include <cstring>

#include "ecc_geometry.cuh"
#include "ecc_internal.h"

using namespace eccb200;

namespace {

constexpr size_t kFlagBytes = 512;                       // flags[world] at the start of a block
constexpr unsigned long long kBarrierTimeoutNs = 60ull * 1000 * 1000 * 1000;  // a rank that never arrives: give up, report

size_t round512(size_t v) { return (v + 511) / 512 * 512; }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Thread r < world: tell rank r that this rank has arrived at `epoch` (release store into slot [rank] of r's flags: all
// stores this rank issued before, including peer stores of earlier kernels on the stream, are visible to whoever reads
// the flag with acquire), then wait until rank r has arrived as well.
__global__ void team_barrier_kernel(unsigned* const* flag_tables, int rank, int world, unsigned epoch, unsigned* status)
{
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    unsigned* theirs = flag_tables[r] + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    const unsigned* mine = flag_tables[rank] + r;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        unsigned seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int)(seen - epoch) >= 0) break;
        if (global_timer_ns() - t0 > kBarrierTimeoutNs) {
            atomicExch(status, 1u);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

template <typename T>
__global__ void team_publish_kernel(const T* __restrict__ src, size_t count, const Mirrors mir)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
        const T v = src[k];
#pragma unroll
        for (int r = 0; r < kMaxPeers; r++)
            if (r < mir.n) *(T*)((char*)(src + k) + mir.delta[r]) = v;
    }
}

// cost image entries i + j*n (i<j) from the pair values in get_ij order (EpipolarConsistencyCommon.hxx:52-79)
__global__ void scatter_cost_kernel(const float* __restrict__ vals, long long total, int n, float* __restrict__ image)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    int i, j;
    pair_from_index(k, n, i, j);
    image[(size_t)i + (size_t)j * n] = vals[k];
}

struct DeviceGuard {
    explicit DeviceGuard(ecc_context* c) { cudaSetDevice(c->device); }
};

Mirrors all_mirrors(const Team& T)
{
    Mirrors m;
    m.n = 0;
    for (int r = 0; r < T.world; r++)
        if (r != T.rank) m.delta[m.n++] = (long long)((char*)T.bases[r] - (char*)T.bases[T.rank]);
    for (int k = m.n; k < kMaxPeers; k++) m.delta[k] = 0;
    return m;
}

bool in_own_block(const Team& T, const void* p)
{
    return T.block && (const char*)p >= (const char*)T.block && (const char*)p < (const char*)T.block + T.block_bytes;
}

}  // namespace

namespace eccb200 {

Mirrors team_mirrors(const ecc_context* ctx, const void* out)
{
    const Team& T = ctx->team;
    if (T.connected && T.mirror_radon && T.world > 1 && in_own_block(T, out)) return all_mirrors(T);
    Mirrors m;
    std::memset(&m, 0, sizeof(m));
    return m;
}

int team_publish(ecc_context* ctx, const void* ptr, size_t bytes)
{
    Team& T = ctx->team;
    if (!T.connected) return fail(ctx, ECC_ERR_STATE, "team not connected");
    if (T.world < 2 || bytes == 0) return ECC_OK;
    if (!in_own_block(T, ptr)) return fail(ctx, ECC_ERR_INVALID, "team_publish: range outside the team block");
    const Mirrors m = all_mirrors(T);
    if (((size_t)ptr | bytes) % 16 == 0) {
        const size_t count = bytes / 16;
        const int blocks = (int)((count + 255) / 256 < (size_t)ctx->sm_count * 8 ? (count + 255) / 256 : (size_t)ctx->sm_count * 8);
        team_publish_kernel<float4><<<blocks, 256, 0, ctx->stream>>>((const float4*)ptr, count, m);
    } else {
        const size_t count = bytes / 4;
        const int blocks = (int)((count + 255) / 256 < (size_t)ctx->sm_count * 8 ? (count + 255) / 256 : (size_t)ctx->sm_count * 8);
        team_publish_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float*)ptr, count, m);
    }
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

int team_barrier(ecc_context* ctx)
{
    Team& T = ctx->team;
    if (!T.connected) return fail(ctx, ECC_ERR_STATE, "team not connected");
    if (T.world < 2) return ECC_OK;
    T.epoch += 1;
    team_barrier_kernel<<<1, 32, 0, ctx->stream>>>(T.flag_tables_d, T.rank, T.world, T.epoch, T.status_d);
    ECC_CUDA(ctx, cudaGetLastError());
    return ECC_OK;
}

void team_free(ecc_context* ctx)
{
    Team& T = ctx->team;
    if (T.world == 0) return;
    cudaStreamSynchronize(ctx->stream);
    for (void* p : T.opened) cudaIpcCloseMemHandle(p);
    if (T.flag_tables_d) cudaFree(T.flag_tables_d);
    if (T.status_d) cudaFree(T.status_d);
    if (T.block) cudaFree(T.block);
    T = Team();
}

}  // namespace eccb200

namespace {

int finish_connect(ecc_context* ctx)
{
    Team& T = ctx->team;
    std::vector<unsigned*> tables(T.world);
    for (int r = 0; r < T.world; r++) tables[r] = (unsigned*)T.bases[r];
    if (!T.flag_tables_d) ECC_CUDA(ctx, cudaMalloc(&T.flag_tables_d, sizeof(unsigned*) * T.world));
    ECC_CUDA(ctx, cudaMemcpy(T.flag_tables_d, tables.data(), sizeof(unsigned*) * T.world, cudaMemcpyHostToDevice));
    T.connected = true;
    return ECC_OK;
}

}  // namespace

extern "C" {

int ecc_team_create(ecc_context* ctx, int rank, int world, int n_total, int n_alpha, int n_t, unsigned char* handle_out)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    if (world < 1 || world > kMaxPeers + 1 || rank < 0 || rank >= world || n_total < 1 || n_alpha < 1 || n_t < 1)
        return fail(ctx, ECC_ERR_INVALID, "ecc_team_create: bad argument (at most 16 ranks)");
    team_free(ctx);
    Team& T = ctx->team;
    T.rank = rank;
    T.world = world;
    T.n_total = n_total;
    T.n_alpha = n_alpha;
    T.n_t = n_t;
    const size_t pairs = (size_t)n_total * (n_total - 1) / 2;
    T.vals_offset = kFlagBytes;
    T.vals_bytes = round512(sizeof(float) * (pairs ? pairs : 1));
    T.dtrs_offset = T.vals_offset + 2 * T.vals_bytes;
    T.block_bytes = T.dtrs_offset + round512(sizeof(float) * (size_t)n_total * n_t * n_alpha);
    ECC_CUDA(ctx, cudaMalloc(&T.block, T.block_bytes));
    ECC_CUDA(ctx, cudaMemset(T.block, 0, T.dtrs_offset));  // flags and values
    ECC_CUDA(ctx, cudaMalloc(&T.status_d, sizeof(unsigned)));
    ECC_CUDA(ctx, cudaMemset(T.status_d, 0, sizeof(unsigned)));
    ECC_CUDA(ctx, cudaDeviceSynchronize());  // the flags are zero before anybody can learn the handle
    T.bases.assign(world, nullptr);
    T.bases[rank] = T.block;
    if (handle_out) {
        std::memset(handle_out, 0, ECC_TEAM_HANDLE_BYTES);
        if (world > 1) {
            cudaIpcMemHandle_t h;
            static_assert(sizeof(h) <= ECC_TEAM_HANDLE_BYTES, "IPC handle does not fit");
            ECC_CUDA(ctx, cudaIpcGetMemHandle(&h, T.block));
            std::memcpy(handle_out, &h, sizeof(h));
        }
    }
    if (world == 1) return finish_connect(ctx);
    return ECC_OK;
}

int ecc_team_connect(ecc_context* ctx, const unsigned char* handles)
{
    if (!ctx || !handles) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    Team& T = ctx->team;
    if (T.world == 0) return fail(ctx, ECC_ERR_STATE, "ecc_team_create first");
    for (int r = 0; r < T.world; r++) {
        if (r == T.rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * ECC_TEAM_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        ECC_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        T.opened.push_back(p);
        T.bases[r] = p;
    }
    return finish_connect(ctx);
}

int ecc_team_connect_pointers(ecc_context* ctx, void* const* blocks)
{
    if (!ctx || !blocks) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    Team& T = ctx->team;
    if (T.world == 0) return fail(ctx, ECC_ERR_STATE, "ecc_team_create first");
    for (int r = 0; r < T.world; r++) {
        if (r == T.rank) continue;
        if (!blocks[r]) return fail(ctx, ECC_ERR_INVALID, "ecc_team_connect_pointers: null block");
        cudaPointerAttributes a;
        ECC_CUDA(ctx, cudaPointerGetAttributes(&a, blocks[r]));
        if (a.type != cudaMemoryTypeDevice) return fail(ctx, ECC_ERR_INVALID, "ecc_team_connect_pointers: not device memory");
        if (a.device != ctx->device) {
            int can = 0;
            ECC_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, a.device));
            if (!can) return fail(ctx, ECC_ERR_UNSUPPORTED, "no peer access between the team's devices");
            const cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
            cudaGetLastError();
        }
        T.bases[r] = blocks[r];
    }
    return finish_connect(ctx);
}

int ecc_team_block(ecc_context* ctx, void** block, float** dtrs)
{
    if (!ctx || ctx->team.world == 0) return ECC_ERR_INVALID;
    if (block) *block = ctx->team.block;
    if (dtrs) *dtrs = ctx->team.dtrs();
    return ECC_OK;
}

int ecc_team_destroy(ecc_context* ctx)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    team_free(ctx);
    return ECC_OK;
}

int ecc_team_barrier(ecc_context* ctx)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    return team_barrier(ctx);
}

int ecc_team_radon_shard(int n_total, int world, int rank, int* first, int* count, int* lo_num, int* hi_num, int* den)
{
    if (n_total < 0 || world < 1 || rank < 0 || rank >= world || !first || !count || !lo_num || !hi_num || !den) return ECC_ERR_INVALID;
    // Q quads of four projections, cut into `world` equal intervals of Q / world quads each (a rational number of quads):
    // rank r takes [r Q / world, (r + 1) Q / world).  The quad an interval boundary falls into is shared by the two ranks.
    const long long Q = (n_total + 3) / 4, a = (long long)rank * Q, b = (long long)(rank + 1) * Q;
    const long long first_quad = a / world, end_quad = (b + world - 1) / world;  // quads [first_quad, end_quad) are touched
    *den = world;
    *lo_num = (int)(a % world);
    *hi_num = (b % world) ? (int)(b % world) : world;
    *first = (int)(4 * first_quad);
    const long long last = 4 * end_quad < n_total ? 4 * end_quad : n_total;
    *count = (int)(last - 4 * first_quad > 0 ? last - 4 * first_quad : 0);
    if (*count == 0) { *first = 0; *lo_num = 0; *hi_num = world; }
    return ECC_OK;
}

int ecc_team_radon_compute_part(ecc_context* ctx, const float* images, int first, int n_local, int lo_num, int hi_num, int den, int n_u,
                                int n_v, int filter, int post_process, int interp)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    Team& T = ctx->team;
    if (!T.connected) return fail(ctx, ECC_ERR_STATE, "team not connected");
    if (first < 0 || n_local < 0 || first + n_local > T.n_total) return fail(ctx, ECC_ERR_INVALID, "ecc_team_radon_compute: projection range outside the team's data set");
    QuadPart part;
    part.lo_num = lo_num;
    part.hi_num = hi_num;
    part.den = den;
    if (den < 1 || lo_num < 0 || lo_num >= den || hi_num < 1 || hi_num > den) return fail(ctx, ECC_ERR_INVALID, "ecc_team_radon_compute_part: bad part");
    if (!part.whole() && first % 4 != 0) return fail(ctx, ECC_ERR_INVALID, "ecc_team_radon_compute_part: a shared quad starts at a multiple of four projections");
    int rc = ECC_OK;
    if (n_local > 0) {
        if (!images) return fail(ctx, ECC_ERR_INVALID, "ecc_team_radon_compute: null images");
        T.mirror_radon = true;
        rc = radon_compute_impl(ctx, images, n_local, n_u, n_v, T.n_alpha, T.n_t, filter, post_process, interp,
                                T.dtrs() + (size_t)first * T.n_t * T.n_alpha, false, part);
        T.mirror_radon = false;
        if (rc) return rc;
    }
    if ((rc = team_barrier(ctx))) return rc;
    if (n_local > 0 && !is_device_pointer(images)) ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's host buffer is free again
    return ECC_OK;
}

int ecc_team_radon_compute(ecc_context* ctx, const float* images, int first, int n_local, int n_u, int n_v, int filter,
                           int post_process, int interp)
{
    return ecc_team_radon_compute_part(ctx, images, first, n_local, 0, 1, 1, n_u, n_v, filter, post_process, interp);
}

int ecc_team_evaluate(ecc_context* ctx, float* cost_image, double* mean)
{
    if (!ctx) return ECC_ERR_INVALID;
    DeviceGuard g(ctx);
    Team& T = ctx->team;
    if (!T.connected) return fail(ctx, ECC_ERR_STATE, "team not connected");
    PairLaunch L;
    int rc = fill_pair_launch(ctx, L);
    if (rc) return rc;
    const long long n = ctx->n_views, total = n * (n - 1) / 2;
    if (n > T.n_total) return fail(ctx, ECC_ERR_STATE, "more projection matrices than the team was created for");
    if (ctx->n_dtrs < ctx->n_views) return fail(ctx, ECC_ERR_STATE, "all-pairs evaluation needs one dtr per projection matrix");
    if (mean) *mean = 0.0;
    std::vector<long long> bounds(T.world + 1);
    if ((rc = ecc_partition_pairs(ctx, T.world, bounds.data()))) return rc;
    const long long lo = bounds[T.rank], hi = bounds[T.rank + 1];
    float* const vals = T.vals(T.evaluations++);  // this evaluation's value buffer (every rank makes the same sequence of calls)
    if (hi > lo) {
        L.pair_begin = lo;
        L.n_pairs = hi - lo;
        L.mode_items = total;  // as the single-GPU job computes it (ecc_evaluate_range)
        L.vals_d = vals + lo;
        L.image_d = nullptr;
        if ((rc = launch_pairs(ctx, L))) return rc;
        if ((rc = team_publish(ctx, vals + lo, sizeof(float) * (size_t)(hi - lo)))) return rc;
    }
    if ((rc = team_barrier(ctx))) return rc;
    // every rank now holds all values: same fixed-order sum, same cost image everywhere
    size_t scap = ctx->sums_cap * sizeof(double);
    if ((rc = ensure_bytes(ctx, (void**)&ctx->sums_d, &scap, sizeof(double)))) return rc;
    ctx->sums_cap = scap / sizeof(double);
    if (total > 0 && (rc = launch_sum_sets(ctx, vals, total, 1, ctx->sums_d))) return rc;
    const bool img_dev = cost_image && is_device_pointer(cost_image);
    const bool host_image = cost_image && !img_dev;
    if (img_dev && total > 0) {
        scatter_cost_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(vals, total, (int)n, cost_image);
        ECC_CUDA(ctx, cudaGetLastError());
    }
    const size_t need = 16 + (host_image ? sizeof(float) * (size_t)total : 0);
    if ((rc = ensure_pinned(ctx, need))) return rc;
    double* sum_h = (double*)ctx->pinned_h;
    unsigned* status_h = (unsigned*)((char*)ctx->pinned_h + 8);
    float* vals_h = (float*)((char*)ctx->pinned_h + 16);
    *sum_h = 0.0;
    if (total > 0) ECC_CUDA(ctx, cudaMemcpyAsync(sum_h, ctx->sums_d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaMemcpyAsync(status_h, T.status_d, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    if (host_image && total > 0) ECC_CUDA(ctx, cudaMemcpyAsync(vals_h, vals, sizeof(float) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    ECC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*status_h) return fail(ctx, ECC_ERR_STATE, "team barrier timed out: a rank did not arrive");
    if (host_image) {
        long long k = 0;
        for (long long i = 0; i < n; i++)
            for (long long j = i + 1; j < n; j++) cost_image[(size_t)i + (size_t)j * n] = vals_h[k++];
    }
    if (mean) *mean = total ? *sum_h / (double)total : 0.0;
    return ECC_OK;
}

}  // extern "C"
