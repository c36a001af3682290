// ecc_internal.h -- context object and launcher declarations of libecc_b200 (not installed).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ecc_b200.h"

namespace eccb200 {

// Kernel families for the per-launch event profile.
enum Family { FAM_RADON = 0, FAM_PAIRS, FAM_GEOMETRY, FAM_REDUCE, FAM_SYNTH, FAM_STAGE, FAM_COUNT };

struct ProfileSlot {
    cudaEvent_t start, stop;
    int family;
};

// Input-image staging for the Radon kernel: a pool of CUDA arrays (2-D tiled layout for the texture
// units) with one linear-filter texture object each.
struct ImagePool {
    int n_u = 0, n_v = 0, count = 0;
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> tex_h;
    cudaTextureObject_t* tex_d = nullptr;
};

// Staging of the hybrid Radon kernel's shared-memory path: padded and transposed-padded image copies, their TMA
// tensor maps (CUtensorMap, kept as raw bytes so that this header does not need cuda.h) and the work-queue word.
struct HybridStage {
    int n_u = 0, n_v = 0, count = 0;
    float* pad_n = nullptr;
    float* pad_t = nullptr;
    unsigned* queue = nullptr;  // [2 counters][one claim word per item]
    size_t queue_words = 0;
    alignas(64) unsigned char map_n[128] = {};
    alignas(64) unsigned char map_t[128] = {};
};

// Staging of the quad kernel (ecc_radon_hybrid4.cu): four images interleaved per texel.
struct Hybrid4Stage {
    int n_u = 0, n_v = 0, quads = 0;
    std::vector<cudaArray_t> arrays;           // float4 arrays, one per quad
    std::vector<cudaTextureObject_t> tex_h;
    cudaTextureObject_t* tex_d = nullptr;
    std::vector<cudaSurfaceObject_t> surf_h;   // the same arrays as surfaces: the staging kernel writes the texels in place
    cudaSurfaceObject_t* surf_d = nullptr;
    void* pad_n = nullptr;                     // [quads][n_v+1][n_u+1] float4
    void* pad_t = nullptr;                     // [quads][n_u+1][n_v+1] float4
    unsigned* queue = nullptr;
    size_t queue_words = 0;
    alignas(64) unsigned char map_n[128] = {};
    alignas(64) unsigned char map_t[128] = {};
    // static split (ECC_INTERP_HYBRID_STATIC): items [0, split_items) of every quad take the window path; cached per geometry
    int split_key[4] = {0, 0, 0, 0};  // n_u, n_v, n_alpha, n_t
    int split_items = -1;
    int split_cfg = -1;               // window configuration the split was computed for
    int split_share = 0;              // share_override the split was computed for
    int share_override = 0;           // > 0: window path's share of the samples in per mille (ecc_radon_set_split), 0: built-in
    int map_cfg = -1;                 // window configuration the tensor maps are encoded for
    // order of a quad's items for the static split: order[0 .. split_items) -> window path, the rest -> texture path
    int* order_d = nullptr;
    int order_len = 0;
    // running sample counts along the two lists of a quad (window entries; texture sub-tiles): where a PART of a quad is cut
    // (QuadPart), so that both paths get the same share of their samples -- not of their entries
    std::vector<double> win_prefix, tex_prefix;
};

// Peer mirrors of an output buffer (multi-GPU team, ecc_team.cu): a kernel that stores out[k] also stores the same
// element of every other rank's copy, *(float*)((char*)&out[k] + delta[r]) for r < n, over NVLink.
constexpr int kMaxPeers = 15;
struct Mirrors {
    long long delta[kMaxPeers];
    int n;
};

// The GPUs of one node that work on one data set (one process and one context per GPU).  Every rank owns ONE device
// block [flags | pair values | all Radon intermediates] and maps the blocks of all other ranks (CUDA IPC, or plain
// pointers inside one process); kernels write their results into every block directly.
struct Team {
    int rank = 0, world = 0;  // world == 0: no team
    int n_total = 0, n_alpha = 0, n_t = 0;
    size_t block_bytes = 0, vals_offset = 0, vals_bytes = 0, dtrs_offset = 0;
    void* block = nullptr;            // own block (cudaMalloc)
    std::vector<void*> bases;         // [world] block of every rank in this process's address space (own included)
    std::vector<void*> opened;        // the ones mapped with cudaIpcOpenMemHandle
    unsigned** flag_tables_d = nullptr;  // [world] device table of the ranks' flag arrays
    unsigned* status_d = nullptr;     // set by a barrier that timed out
    unsigned epoch = 0;
    unsigned evaluations = 0;         // ecc_team_evaluate calls so far: evaluation k uses value buffer k & 1 (see vals())
    bool connected = false;
    bool mirror_radon = false;        // while ecc_team_radon_compute runs: Radon kernels mirror their stores
    float* dtrs() const { return (float*)((char*)block + dtrs_offset); }
    // Two value buffers, used alternately: a rank that has left evaluation k may already publish the values of evaluation
    // k+1 into its peers while they still sum / scatter / download those of k (nothing orders a peer's reads of k before
    // this rank's next stores otherwise); it cannot reach k+2 -- the buffer of k again -- before every peer has arrived at
    // the barrier of k+1, i.e. has finished reading k.
    float* vals(unsigned evaluation) const { return (float*)((char*)block + vals_offset + (evaluation & 1u) * vals_bytes); }
};

// ecc_update_and_evaluate: the {replace one matrix, evaluate a pair list} step of tracking loops, recorded once as a CUDA
// graph and replayed.  `key` holds everything the recorded nodes depend on (sizes, settings, buffer addresses).
struct TrackGraph {
    cudaGraphExec_t exec = nullptr;
    std::vector<double> key, pending_key;
    std::vector<int> idx_h;      // the host pair list the device copy below was made from
    int* idx_d = nullptr;        // own copy of the pair list (the context's idx_d is scratch of other calls)
    size_t idx_cap = 0;
    void* pinned = nullptr;      // [16 floats view | sequence number, pad | double sum | flag, pad | floats vals]
    size_t pinned_bytes = 0;
    float* live_d = nullptr;     // device copy of the first 80 bytes of the pinned block (one copy node per replay)
    unsigned* done_d = nullptr;  // CTA counter of the fused launch (left at zero by every launch)
    unsigned seq = 0;
    bool fused = false;          // the recording is the two-node one (pair kernel with the fused tail)
    bool failed = false;         // capture did not work here: stay on the plain path
    long long replays = 0;
};

// ecc_evaluate_batch: the K matrix sets of a launch and what is derived from them; separate from the context's current
// matrices, which stay valid across a batched call.
struct BatchBuffers {
    double* Ps_d = nullptr;
    float* Cs_d = nullptr;
    float* A_d = nullptr;
    size_t cap = 0;
    float* radii_d = nullptr;
    size_t radii_cap = 0;
    double* params_d = nullptr;   // ecc_evaluate_batch_params: parameter vectors and base matrices
    size_t params_cap = 0;
    double* base_d = nullptr;
    size_t base_cap = 0;
};

// The direct metric (ecc_direct.cu): the projection images resident as textures, geometry records and scratch.
struct DirectState {
    int n_u = 0, n_v = 0, n_images = 0;
    int fbcc = 0;            // MetricDirect::setFanBeamConsistency
    int reference_clip = 0;  // clip lines against n_u x n_u as the reference's launcher does (EpipolarConsistencyDirect.cu:135)
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> tex_h;
    cudaTextureObject_t* tex_d = nullptr;
    void* views_d = nullptr;   // DirectView per view
    size_t views_bytes = 0;
    double* Ps_d = nullptr;
    size_t Ps_bytes = 0;
    void* pairs_d = nullptr;   // DirectPair per pair
    size_t pairs_bytes = 0;
    int* ij_d = nullptr;
    size_t ij_bytes = 0;
    int* offsets_d = nullptr;  // [n_pairs + 1] first CTA of every pair, then [n_pairs] CTAs per pair
    size_t offsets_bytes = 0;
    double* vals_d = nullptr;  // [n_pairs] pair values, then the total
    size_t vals_bytes = 0;
    double* partials_d = nullptr;
    size_t partials_bytes = 0;
    float* scratch_d = nullptr;
    size_t scratch_bytes = 0;
    std::vector<int> all_pairs_h;  // (i, j) of the enumeration i < j, i outer, for all_pairs_n views
    int all_pairs_n = 0;
};

}  // namespace eccb200

struct ecc_context {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream all work is issued on
    cudaStream_t own_stream = nullptr;  // created by ecc_create
    cudaStream_t copy_stream = nullptr; // uploads of host images under the Radon kernels (created on first use)
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    cudaStream_t down_stream = nullptr; // downloads of finished intermediates to host memory under the next chunk's kernels
    cudaEvent_t ev_out_ready[2] = {nullptr, nullptr}, ev_out_free[2] = {nullptr, nullptr};
    int sm_count = 148;
    std::string last_error;

    // ---- projection matrices ----
    int n_views = 0;
    std::vector<double> Ps_h;   // n_views*12, host copy (getProjectionMatrices, auto radius)
    double* Ps_d = nullptr;     // device staging for derive kernel (capacity Ps_cap matrices)
    float* Cs_d = nullptr;      // n*4
    float* PinvTs_d = nullptr;  // n*12
    size_t Ps_cap = 0;          // capacity in matrices of Ps_d / Cs_d / PinvTs_d
    long long geometry_version = 0;         // bumped whenever a matrix changes
    std::vector<double> partition_key;      // settings the cached pair partition was computed for
    std::vector<long long> partition_bounds;

    // ---- radon intermediates ----
    int n_dtrs = 0, n_alpha = 0, n_t = 0, n_u = 0, n_v = 0, is_derivative = 1;
    float step_alpha = 0.f, step_t = 0.f;
    const float* dtrs_d = nullptr;  // [n_dtrs][n_t][dtr_pitch] when set from one block, else null
    std::vector<const float*> dtr_ptrs_h;  // one device pointer per dtr
    const float** dtr_ptrs_d = nullptr;    // the same table on the device
    size_t dtr_ptrs_cap = 0;
    size_t dtr_pitch = 0;           // floats per row
    size_t dtr_stride = 0;          // floats per dtr
    float* dtrs_owned = nullptr;    // non-null when the context owns the storage
    size_t dtrs_owned_bytes = 0;
    std::vector<cudaTextureObject_t> dtr_tex_h;
    std::vector<cudaArray_t> dtr_arrays;   // development (ECC_DTR_ARRAYS=1): tiled copies of the dtrs instead of textures over the caller's memory
    cudaTextureObject_t* dtr_tex_d = nullptr;
    size_t dtr_tex_cap = 0;

    // ---- settings ----
    double object_radius = 0.0;
    double dkappa = 0.0;
    int interp = ECC_INTERP_TEXTURE;
    int use_corr = 0;

    // ---- scratch ----
    float* vals_d = nullptr;  // one float per evaluated (set,pair)
    float* partials_d = nullptr;  // partial sums of split CTA-per-pair launches
    size_t partials_cap = 0;
    float* pair_records_d = nullptr;  // warp-per-pair launches: the pairs' maps, one 64-byte record each (launch_pairs)
    size_t pair_records_bytes = 0;
    size_t vals_cap = 0;
    double* sums_d = nullptr;  // one double per set
    size_t sums_cap = 0;
    int* idx_d = nullptr;  // uploaded index list
    size_t idx_cap = 0;
    int* counts_d = nullptr;
    size_t counts_cap = 0;
    float* img_stage_d = nullptr;  // device staging for host images / host dtr output
    size_t img_stage_bytes = 0;
    float* out_stage_d = nullptr;
    size_t out_stage_bytes = 0;
    void* pinned_h = nullptr;  // pinned host staging
    size_t pinned_bytes = 0;
    float* cost_d = nullptr;  // n*n device cost image when the caller's is on the host
    size_t cost_cap = 0;

    eccb200::ImagePool pool;
    float* pre_work_d = nullptr;   // pre-processing scratch (ecc_preprocess.cu)
    size_t pre_work_bytes = 0;
    void* pre_small_d = nullptr;
    size_t pre_small_bytes = 0;
    float* ramp_g_d = nullptr;  // ramp-filter kernel g[n_t] (ecc_radon.cu)
    int ramp_n_t = 0;
    eccb200::HybridStage hybrid;
    eccb200::Hybrid4Stage hybrid4;
    eccb200::Team team;
    eccb200::TrackGraph track;
    eccb200::BatchBuffers batch;
    eccb200::DirectState direct;

    // ---- profiling ----
    bool profiling = false;
    std::vector<eccb200::ProfileSlot> prof_slots;
    double prof_ms[eccb200::FAM_COUNT] = {};
    long long prof_launches[eccb200::FAM_COUNT] = {};
};

namespace eccb200 {

int fail(ecc_context* ctx, int code, const std::string& msg);
int cuda_fail(ecc_context* ctx, cudaError_t e, const char* what, const char* file, int line);

#define ECC_CUDA(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) return eccb200::cuda_fail((ctx), e__, #call, __FILE__, __LINE__); \
    } while (0)

// RAII-less profile bracket: begin returns a slot index or -1.
int prof_begin(ecc_context* ctx, int family);
void prof_end(ecc_context* ctx, int slot);
void prof_collect(ecc_context* ctx);

bool is_device_pointer(const void* p);
int ensure_bytes(ecc_context* ctx, void** ptr, size_t* cap, size_t bytes);
int ensure_pinned(ecc_context* ctx, size_t bytes);

// ---- launchers (ecc_pairs.cu) ----
struct PairLaunch {
    // problem
    int n_views;           // matrices per set
    int n_sets;            // 1 unless batched
    long long pair_begin;  // all-pairs mode: first pair of the enumeration
    long long n_pairs;     // pairs per set
    const int* idx4_d;     // pair list (device) or null for all-pairs
    const float* Cs_d;     // n_sets*n_views*4
    const float* PinvTs_d; // n_sets*n_views*12
    // dtrs
    const cudaTextureObject_t* tex_d;
    const float* const* dtr_ptrs_d;  // one device pointer per dtr
    size_t dtr_pitch;
    int n_dtrs, n_alpha, n_t;
    float half_nu, half_nv, range_t, image_diagonal;
    float radius, dkappa;
    const float* radii_d;  // batched mode: one object radius per matrix set (device) or null
    int sample_cap;
    int is_derivative;
    int interp;
    int use_corr;  // correlation variant instead of the SSD
    long long mode_items;  // > 0: choose warp-per-pair / CTA-per-pair and the split count as for a launch of this many pairs
                           // (ranges of a larger job: a pair's fp32 sum must not depend on how the job was cut)
    int defer_finalize;  // split launches: leave the partial sums for launch_finalize_sum
    int splits;          // set by launch_pairs: CTAs per pair (CTA-per-pair launches)
    float* partials_d;   // [items][splits][3] when splits > 1
    // set by launch_pairs, warp-per-pair launches: [items][16] floats -- every pair's maps computed ONCE by one thread of
    // pair_records_kernel instead of by all 32 lanes of the pair's warp (ecc_pairs.cu); null = the warp computes them
    const float* records_d = nullptr;
    float* corr_sums_d;  // correlation variant, nullable: the reference launcher's six sums per pair (forces splits = 1)
    // outputs
    float* vals_d;   // n_sets*n_pairs
    float* image_d;  // all-pairs: n_views*n_views cost image or null (only with n_sets==1)
    // tracking steps (ecc_update_and_evaluate, CTA-per-pair launches of one set): the whole call in ONE launch.  The live view
    // is read from live_d ([12 floats (P+)^T | 4 floats C | sequence number], uploaded by the one copy node in front of the
    // kernel) instead of entry live_index of the arrays; the CTA that finishes last (done_d counts them) adds the splits'
    // partial sums and the pairs' values in launch_finalize_sum's order, writes values and sum to fused_vals_out / fused_sum_out
    // (pinned host memory), stores the live view into the arrays, and publishes the sequence number at fused_flag.
    const float* live_d = nullptr;
    int live_index = -1;
    unsigned* done_d = nullptr;
    double* fused_sum_out = nullptr;
    float* fused_vals_out = nullptr;
    unsigned* fused_flag = nullptr;
};
// resolved (nullable): the launch record as launched (splits, partials_d filled in)
int launch_pairs(ecc_context* ctx, const PairLaunch& L, PairLaunch* resolved = nullptr);
// single matrix set, after launch_pairs(..., &resolved) with defer_finalize: finalize + fixed-order fp64 sum in one launch;
// sum_out / vals_out (nullable) may be pinned host memory
int launch_finalize_sum(ecc_context* ctx, const PairLaunch& resolved, double* sum_out, float* vals_out);
int fill_pair_launch(ecc_context* ctx, PairLaunch& L);  // everything that depends on the context state only (ecc_capi.cu)
int launch_pair_counts(ecc_context* ctx, const PairLaunch& L, int* counts_d);
// K01s_d [n_pairs][16]: the reference's K01 record of every pair of the launch (ecc_pairs.cu: pair_maps_kernel)
int launch_pair_maps(ecc_context* ctx, const PairLaunch& L, float* K01s_d);
int launch_fill(ecc_context* ctx, float* dst_d, size_t count, size_t stride, float value);  // dst[k * stride] = value, k < count
// one pair (L.idx4_d[0..3]): rec_d [sample_cap][13] floats, head_d [2] ints zeroed before the launch (ecc_pairs.cu: pair_signals_kernel)
int launch_pair_signals(ecc_context* ctx, const PairLaunch& L, float* rec_d, int* head_d);
// equal-work cut of counts_d[0..total) (+16 per pair) into n_parts ranges, on the device: bounds_d [n_parts + 1]
int launch_partition(ecc_context* ctx, const int* counts_d, long long total, int n_parts, long long* bounds_d);
int launch_sum_sets(ecc_context* ctx, const float* vals_d, long long n_pairs, int n_sets,
                    double* sums_d);
// radii_d (nullable): entry s receives the automatic object radius of set s (views_per_set matrices per set),
// or fixed_radius when that is > 0.
int launch_derive_views(ecc_context* ctx, const double* Ps_d, int n, float* PinvTs_d, float* Cs_d,
                        int views_per_set = 0, int n_u = 0, int n_v = 0, double fixed_radius = 0.0,
                        float* radii_d = nullptr);

// ---- batched evaluation, the parts ecc_evaluate_batch and ecc_evaluate_batch_params share (ecc_capi.cu) ----
int batch_begin(ecc_context* ctx, const int* idx4, int n_pairs, PairLaunch& L, const char* who);
int batch_reserve(ecc_context* ctx, int n_sets, bool want_matrices);
int batch_finish(ecc_context* ctx, PairLaunch& L, int n_sets, float* out, double* means);

// A call that computes only PART of its first and of its last quad of projections (static-split engine; a team shares the
// quad that straddles two ranks' shards, ecc_team_radon_compute_part): of the first quad the share from lo_num / den on,
// of the last quad the share up to hi_num / den, of both item lists (window path, texture path) alike.  Whole: {0, den, den}.
struct QuadPart {
    int lo_num = 0, hi_num = 1, den = 1;
    bool whole() const { return lo_num == 0 && hi_num == den; }
    // the part that applies to a sub-range of the call's images: its first / last quad is the call's first / last or not
    QuadPart sub(bool has_first, bool has_last) const
    {
        QuadPart q;
        q.den = den;
        q.lo_num = has_first ? lo_num : 0;
        q.hi_num = has_last ? hi_num : den;
        return q;
    }
};

// ---- launchers (ecc_radon.cu) ----
int radon_batch(ecc_context* ctx, const float* images_d, int n_images, int n_u, int n_v,
                int n_alpha, int n_t, int filter, int post, int interp, float* out_d, QuadPart part = QuadPart());
void free_image_pool(ecc_context* ctx);
int ramp_filter(ecc_context* ctx, float* dtrs_d, int n, int n_alpha, int n_t);
// ---- launchers (ecc_radon_hybrid.cu) ----
int radon_hybrid_launch(ecc_context* ctx, const cudaTextureObject_t* texs_d, const float* images_d, int n, int n_u,
                        int n_v, int n_alpha, int n_t, int post, float* out_d);
void free_hybrid(ecc_context* ctx);
// ---- launchers (ecc_radon_hybrid4.cu) ----
int radon_hybrid4_launch(ecc_context* ctx, const float* images_d, int n, int n_u, int n_v, int n_alpha, int n_t, int post, float* out_d,
                         bool static_split = false, QuadPart part = QuadPart());
void free_hybrid4(ecc_context* ctx);
int radon_hybrid4_reserve(ecc_context* ctx, int n_u, int n_v, int n_images);
int radon_hybrid4_calibrate(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int repeats, int* permille);
int radon_num_samples(ecc_context* ctx, int n_u, int n_v, int n_alpha, int n_t, int filter, double* count);

// ---- multi-GPU team (ecc_team.cu) ----
// Mirrors for stores into [out, out + bytes) when that range lies in the team's own block and mirroring is on, else n = 0.
Mirrors team_mirrors(const ecc_context* ctx, const void* out);
// Copies [ptr, ptr + bytes) of the own block to the same place of every other block (for kernels that do not mirror themselves).
int team_publish(ecc_context* ctx, const void* ptr, size_t bytes);
int team_barrier(ecc_context* ctx);
void team_free(ecc_context* ctx);
int radon_compute_impl(ecc_context* ctx, const float* images, int n_images, int n_u, int n_v, int n_alpha, int n_t, int filter,
                       int post, int interp, float* dtrs_out, bool sync_device_path, QuadPart part = QuadPart());

// ---- launchers (ecc_preprocess.cu) ----
int preprocess_batch(ecc_context* ctx, float* images_d, int n, int n_u, int n_v, const ecc_preprocess_params* pp, const double* Ps_h);
void camera_intrinsics_host(const double* P, double* fu, double* u0, double* v0);

// ---- direct metric (ecc_direct.cu) ----
void free_direct(ecc_context* ctx);

// ---- launchers (ecc_synth.cu) ----
int synth_projections(ecc_context* ctx, const double* Ps_h, int n, int n_u, int n_v,
                      const double* ell_h, int n_ell, int cos_weight, int zero_border,
                      float* images_d);

}  // namespace eccb200
