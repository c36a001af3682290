// launcher_swap_check.cpp -- INTEGRATION.md option B, compiled: the reference's launcher symbols (two of the Radon-intermediate path, one of the direct metric), declared exactly as the
// reference's own .cpp files declare them (EpipolarConsistencyRadonIntermediate.cpp:16-37, RadonIntermediate.cpp:12) and
// called the way those files call them (cudaArray textures from BindlessTexture2D, handle table + Cs + PinvTs in device
// memory, K01s / out / out_corr scratch), resolved by libecc_b200.so.  Results are compared with the C ABI's own entry
// points on the same inputs.  Plain g++ + the CUDA runtime, no nvcc: this is host code, as the reference's .cpp files are.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ecc_b200.h"

// ---- the reference's declarations, verbatim in type and order -----------------------------------------------------------
extern void computeDerivLineIntegrals(cudaTextureObject_t in, int n_x, int n_y, int n_alpha, int n_t, int filter, int post_process, float* out_d);
void epipolarConsistency(int n_x, int n_y, int num_dtrs, char* dtrs_d, int n_alpha, int n_t, float step_alpha, float step_t, int num_Ps,
                         float* Cs_d, float* PinvTs_d, int num_pairs, int* indices_d, float* K01s_d, float* out_d, float object_radius_mm,
                         float dkappa, bool isDerivative, bool use_corr, float* out_corr_d);

// EpipolarConsistencyDirect.cpp:10-16, verbatim
extern void cuda_computeLineIntegrals(
	short n_lines,                               // Number of lines
	float* lines_d, short line_stride,           // Lines in Hessian normal form and number of float values to next line
	float *fbcc_d, short fbcc_stride,            // Optional: FBCC_weighting_info for rectification and source-distance-weighting.
	cudaTextureObject_t I, short n_u, short n_v, // The image and its size
	float *integrals_out_d);                     // Output memory: the integrals (size is n_lines)

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); \
            std::exit(2);                                                                              \
        }                                                                                              \
    } while (0)
#define ECC(call)                                                                                  \
    do {                                                                                           \
        int rc__ = (call);                                                                         \
        if (rc__) {                                                                                \
            std::fprintf(stderr, "%s failed: %d %s\n", #call, rc__, ecc_last_error(ctx));          \
            std::exit(3);                                                                          \
        }                                                                                          \
    } while (0)
static int failures = 0;
#define EXPECT(cond, ...)                          \
    do {                                           \
        if (!(cond)) {                             \
            std::printf("FAIL %s: ", #cond);       \
            std::printf(__VA_ARGS__);              \
            std::printf("\n");                     \
            failures++;                            \
        }                                          \
    } while (0)

// UtilsCuda::BindlessTexture2D<float>(w, h, buffer_d, true, true, normalized): array + linear filter + clamp
// (LibUtilsCuda/CudaBindlessTexture.cpp:17-44)
struct ArrayTexture {
    cudaArray_t array = nullptr;
    cudaTextureObject_t tex = 0;
    ArrayTexture(int w, int h, const float* src_d, bool normalized)
    {
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        CK(cudaMallocArray(&array, &desc, w, h));
        CK(cudaMemcpy2DToArray(array, 0, 0, src_d, sizeof(float) * w, sizeof(float) * w, h, cudaMemcpyDeviceToDevice));
        cudaResourceDesc res;
        std::memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;
        res.res.array.array = array;
        cudaTextureDesc td;
        std::memset(&td, 0, sizeof(td));
        td.normalizedCoords = normalized;
        td.filterMode = cudaFilterModeLinear;
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&tex, &res, &td, NULL));
    }
    ~ArrayTexture()
    {
        cudaDestroyTextureObject(tex);
        cudaFreeArray(array);
    }
};

int main()
{
    const int n = 7, n_u = 160, n_v = 128, n_alpha = 96, n_t = 96;
    const size_t img = (size_t)n_u * n_v, dtr = (size_t)n_alpha * n_t;
    ecc_context* ctx = nullptr;
    if (ecc_create(-1, &ctx)) {
        std::fprintf(stderr, "no CUDA device\n");
        return 4;
    }
    // matrices + synthetic projections (device)
    std::vector<double> Ps(12 * n);
    ecc_make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 2.0, Ps.data());
    const double ell[14] = {0, 0, 0, 60, 40, 50, 1.0, 20, -10, 5, 20, 25, 15, 0.5};
    float* images_d = nullptr;
    CK(cudaMalloc(&images_d, sizeof(float) * img * n));
    ECC(ecc_synth_projections(ctx, Ps.data(), n, n_u, n_v, ell, 2, 1, 1, images_d));
    ECC(ecc_synchronize(ctx));

    // ---- RadonIntermediate::compute as the reference runs it: one texture, one launcher call per projection ----------
    float* dtrs_ref_d = nullptr;  // through computeDerivLineIntegrals
    float* dtrs_abi_d = nullptr;  // through ecc_radon_compute
    CK(cudaMalloc(&dtrs_ref_d, sizeof(float) * dtr * n));
    CK(cudaMalloc(&dtrs_abi_d, sizeof(float) * dtr * n));
    for (int k = 0; k < n; k++) {
        ArrayTexture t(n_u, n_v, images_d + img * k, false);
        computeDerivLineIntegrals(t.tex, n_u, n_v, n_alpha, n_t, /*filter Derivative*/ 0, /*Identity*/ 0, dtrs_ref_d + dtr * k);
    }
    ECC(ecc_radon_compute(ctx, images_d, n, n_u, n_v, n_alpha, n_t, ECC_FILTER_DERIVATIVE, ECC_POST_IDENTITY, ECC_INTERP_TEXTURE, dtrs_abi_d));
    ECC(ecc_synchronize(ctx));
    std::vector<float> a(dtr * n), b(dtr * n);
    CK(cudaMemcpy(a.data(), dtrs_ref_d, sizeof(float) * dtr * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), dtrs_abi_d, sizeof(float) * dtr * n, cudaMemcpyDeviceToHost));
    double peak = 0;
    for (float v : b) peak = std::fmax(peak, std::fabs(v));
    EXPECT(peak > 0 && std::memcmp(a.data(), b.data(), sizeof(float) * dtr * n) == 0, "computeDerivLineIntegrals != ecc_radon_compute (texture engine)");
    {   // ramp filter and post-processing go through the same symbol
        ArrayTexture t(n_u, n_v, images_d, false);
        float *r1 = nullptr, *r2 = nullptr;
        CK(cudaMalloc(&r1, sizeof(float) * dtr));
        CK(cudaMalloc(&r2, sizeof(float) * dtr));
        computeDerivLineIntegrals(t.tex, n_u, n_v, n_alpha, n_t, /*Ramp*/ 1, /*SquareRoot*/ 1, r1);
        ECC(ecc_radon_compute(ctx, images_d, 1, n_u, n_v, n_alpha, n_t, ECC_FILTER_RAMP, ECC_POST_SQRT, ECC_INTERP_TEXTURE, r2));
        ECC(ecc_synchronize(ctx));
        std::vector<float> x(dtr), y(dtr);
        CK(cudaMemcpy(x.data(), r1, sizeof(float) * dtr, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(y.data(), r2, sizeof(float) * dtr, cudaMemcpyDeviceToHost));
        EXPECT(std::memcmp(x.data(), y.data(), sizeof(float) * dtr) == 0, "ramp + sqrt through the launcher symbol");
        cudaFree(r1);
        cudaFree(r2);
    }

    // ---- MetricRadonIntermediate as the reference holds its state (.cpp:87-106,134-163) ---------------------------------
    std::vector<ArrayTexture*> tex;
    std::vector<cudaTextureObject_t> handles;
    for (int k = 0; k < n; k++) {
        tex.push_back(new ArrayTexture(n_alpha, n_t, dtrs_abi_d + dtr * k, true));  // RadonIntermediate::getTexture (.cpp:188-196)
        handles.push_back(tex.back()->tex);
    }
    char* tex_dtrs_d = nullptr;
    CK(cudaMalloc(&tex_dtrs_d, sizeof(cudaTextureObject_t) * n));
    CK(cudaMemcpy(tex_dtrs_d, handles.data(), sizeof(cudaTextureObject_t) * n, cudaMemcpyHostToDevice));
    std::vector<float> PinvTs(12 * n), Cs(4 * n);
    ecc_derive_views_host(Ps.data(), n, PinvTs.data(), Cs.data());  // = culaut (bit-identical, tests/test_library_cpu.py)
    float *Cs_d = nullptr, *PinvTs_d = nullptr, *K01s_d = nullptr, *out_d = nullptr, *corr_d = nullptr;
    const int pairs = n * (n - 1) / 2;
    CK(cudaMalloc(&Cs_d, sizeof(float) * 4 * n));
    CK(cudaMalloc(&PinvTs_d, sizeof(float) * 12 * n));
    CK(cudaMalloc(&K01s_d, sizeof(float) * 16 * pairs));
    CK(cudaMalloc(&out_d, sizeof(float) * n * n));
    CK(cudaMalloc(&corr_d, sizeof(float) * 6 * pairs));
    CK(cudaMemcpy(Cs_d, Cs.data(), sizeof(float) * 4 * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(PinvTs_d, PinvTs.data(), sizeof(float) * 12 * n, cudaMemcpyHostToDevice));
    double step_alpha = 0, step_t = 0;
    ecc_radon_bin_sizes(n_u, n_v, n_alpha, n_t, &step_alpha, &step_t);

    // the same state in a context of the C ABI
    ECC(ecc_set_radon_intermediates(ctx, dtrs_abi_d, n, n_alpha, n_t, step_alpha, step_t, n_u, n_v, 1));
    ECC(ecc_set_projection_matrices(ctx, Ps.data(), n));
    ECC(ecc_set_interpolation(ctx, ECC_INTERP_TEXTURE));
    double radius = 0;
    ECC(ecc_get_object_radius(ctx, &radius));
    const float dkappa = 0.002f;
    ECC(ecc_set_epipolar_plane_step(ctx, dkappa));

    // evaluate(float*) (.cpp:166-225): out pre-loaded with the caller's values, all pairs
    std::vector<float> cost_ref((size_t)n * n, -7.f), cost_abi((size_t)n * n, -7.f);
    CK(cudaMemcpy(out_d, cost_ref.data(), sizeof(float) * n * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(corr_d, 0, sizeof(float) * pairs));
    epipolarConsistency(n_u, n_v, n, tex_dtrs_d, n_alpha, n_t, (float)step_alpha, (float)step_t, n, Cs_d, PinvTs_d, 0, 0x0, K01s_d, out_d,
                        (float)radius, dkappa, true, false, corr_d);
    CK(cudaMemcpy(cost_ref.data(), out_d, sizeof(float) * n * n, cudaMemcpyDeviceToHost));
    std::vector<float> weights(pairs), K01(16 * pairs), K01_abi(16 * pairs);
    CK(cudaMemcpy(weights.data(), corr_d, sizeof(float) * pairs, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(K01.data(), K01s_d, sizeof(float) * 16 * pairs, cudaMemcpyDeviceToHost));
    double mean_abi = 0;
    ECC(ecc_evaluate(ctx, cost_abi.data(), &mean_abi));
    ECC(ecc_pair_maps(ctx, nullptr, 0, K01_abi.data()));
    EXPECT(std::memcmp(cost_ref.data(), cost_abi.data(), sizeof(float) * n * n) == 0, "all-pairs cost image through the launcher symbol");
    EXPECT(std::memcmp(K01.data(), K01_abi.data(), sizeof(float) * 16 * pairs) == 0, "K01 records");
    double weighted = 0, wsum = 0;  // the reference's host reduction (.cpp:212-224)
    for (int i = 0, k = 0; i < n; i++)
        for (int j = i + 1; j < n; j++, k++) {
            EXPECT(weights[k] == 1.0f, "weight of pair %d = %g", k, weights[k]);
            weighted += cost_ref[i + (size_t)j * n] * weights[k];
            wsum += weights[k];
        }
    EXPECT(std::fabs(weighted / wsum - mean_abi) <= 1e-12 * std::fabs(mean_abi), "mean %.17g vs %.17g", weighted / wsum, mean_abi);
    EXPECT(cost_ref[0] == -7.f && cost_ref[1] == -7.f && cost_ref[(size_t)1 * n + 0] > 0.f, "untouched entries / sub-diagonal layout");
    EXPECT(K01[6] > 0.f && K01[14] == dkappa && K01[15] > 0.f, "K01 record layout: baseline %g dkappa %g kappa_max %g", K01[6], K01[14], K01[15]);

    // evaluate(indices, out) (.cpp:267-322)
    const int list[] = {0, 5, 0, 5, 3, 1, 3, 1, 2, 6, 2, 6, 4, 4, 4, 4};
    const int n_list = 4;
    int* idx_d = nullptr;
    CK(cudaMalloc(&idx_d, sizeof(list)));
    CK(cudaMemcpy(idx_d, list, sizeof(list), cudaMemcpyHostToDevice));
    CK(cudaMemset(corr_d, 0, sizeof(float) * n_list));
    epipolarConsistency(n_u, n_v, n, tex_dtrs_d, n_alpha, n_t, (float)step_alpha, (float)step_t, n, Cs_d, PinvTs_d, n_list, idx_d, K01s_d, out_d,
                        (float)radius, dkappa, true, false, corr_d);
    std::vector<float> got(n_list), want(n_list);
    CK(cudaMemcpy(got.data(), out_d, sizeof(float) * n_list, cudaMemcpyDeviceToHost));
    ECC(ecc_evaluate_indices(ctx, list, n_list, want.data(), nullptr));
    EXPECT(std::memcmp(got.data(), want.data(), sizeof(float) * n_list) == 0, "pair list through the launcher symbol");
    EXPECT(got[3] == 0.f && got[0] == cost_ref[0 + (size_t)5 * n], "same view twice -> 0; (0,5) = the all-pairs entry");

    // useCorrelation(true): six sums per pair, 1 - cc on the host (.cpp:127-131,200-211)
    CK(cudaMemset(corr_d, 0, sizeof(float) * 6 * pairs));
    epipolarConsistency(n_u, n_v, n, tex_dtrs_d, n_alpha, n_t, (float)step_alpha, (float)step_t, n, Cs_d, PinvTs_d, 0, 0x0, K01s_d, out_d,
                        (float)radius, dkappa, true, true, corr_d);
    std::vector<float> sums(6 * pairs);
    CK(cudaMemcpy(sums.data(), corr_d, sizeof(float) * 6 * pairs, cudaMemcpyDeviceToHost));
    ECC(ecc_use_correlation(ctx, 1));
    std::vector<float> corr_abi((size_t)n * n, 0.f);
    ECC(ecc_evaluate(ctx, corr_abi.data(), nullptr));
    ECC(ecc_use_correlation(ctx, 0));
    for (int i = 0, k = 0; i < n; i++)
        for (int j = i + 1; j < n; j++, k++) {
            const float* q = &sums[6 * k];
            const float cc = q[4] / (std::sqrt(q[2]) * std::sqrt(q[3]));
            EXPECT(q[5] == 1.0f && q[2] > 0.f && q[3] > 0.f, "six sums of pair %d", k);
            EXPECT(std::fabs((1.0f - cc) - corr_abi[i + (size_t)j * n]) <= 2e-6f, "1 - cc of pair %d: %g vs %g", k, 1.0f - cc, corr_abi[i + (size_t)j * n]);
        }

    {   // ---- computeForImagePair's device part as the reference runs it (EpipolarConsistencyDirect.cpp:108-125,188-198): lines
        // interleaved l0 l1 (stride 6), FBCC records interleaved (stride 16), one launcher call per image -----------------------
        ECC(ecc_direct_set_images(ctx, images_d, n, n_u, n_v));
        ECC(ecc_direct_set_reference_clip(ctx, 1));  // the reference's launcher clips against n_u x n_u (EpipolarConsistencyDirect.cu:135)
        ECC(ecc_set_object_radius(ctx, 0.0));
        ECC(ecc_set_epipolar_plane_step(ctx, 0.0));
        int n_lines = 0;
        ECC(ecc_direct_pair_geometry(ctx, 1, 4, 0, 0x0, 0x0, 0x0, 0x0, 0x0, &n_lines, 0x0));
        std::vector<float> l0(3 * n_lines), l1(3 * n_lines), f0(8 * n_lines), f1(8 * n_lines), l01(6 * n_lines), f01(16 * n_lines);
        ECC(ecc_direct_pair_geometry(ctx, 1, 4, n_lines, 0x0, l0.data(), l1.data(), f0.data(), f1.data(), &n_lines, 0x0));
        for (int q = 0; q < n_lines; q++) {
            std::memcpy(&l01[6 * q], &l0[3 * q], 12);
            std::memcpy(&l01[6 * q + 3], &l1[3 * q], 12);
            std::memcpy(&f01[16 * q], &f0[8 * q], 32);
            std::memcpy(&f01[16 * q + 8], &f1[8 * q], 32);
        }
        float *l01_d = nullptr, *f01_d = nullptr, *v_d = nullptr;
        CK(cudaMalloc(&l01_d, sizeof(float) * l01.size()));
        CK(cudaMalloc(&f01_d, sizeof(float) * f01.size()));
        CK(cudaMalloc(&v_d, sizeof(float) * n_lines));
        CK(cudaMemcpy(l01_d, l01.data(), sizeof(float) * l01.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(f01_d, f01.data(), sizeof(float) * f01.size(), cudaMemcpyHostToDevice));
        ArrayTexture I1(n_u, n_v, images_d + img * 1, false), I4(n_u, n_v, images_d + img * 4, false);
        std::vector<float> got(n_lines), want(n_lines);
        for (int fb = 0; fb < 2; fb++)
            for (int view = 0; view < 2; view++) {
                cuda_computeLineIntegrals((short)n_lines, l01_d + 3 * view, 6, fb ? f01_d + 8 * view : 0x0, fb ? 16 : 0, view ? I4.tex : I1.tex,
                                          (short)n_u, (short)n_v, v_d);
                CK(cudaMemcpy(got.data(), v_d, sizeof(float) * n_lines, cudaMemcpyDeviceToHost));
                ECC(ecc_direct_line_integrals(ctx, view ? 4 : 1, view ? l1.data() : l0.data(), n_lines, 3, fb ? (view ? f1.data() : f0.data()) : 0x0, 8,
                                              want.data()));
                double energy = 0;
                for (int q = 0; q < n_lines; q++) energy += (double)want[q] * want[q];
                EXPECT(n_lines > 100 && energy > 0 && std::memcmp(got.data(), want.data(), sizeof(float) * n_lines) == 0,
                       "cuda_computeLineIntegrals (fbcc %d, view %d) != ecc_direct_line_integrals", fb, view);
            }
        cudaFree(l01_d);
        cudaFree(f01_d);
        cudaFree(v_d);
    }

    for (auto* t : tex) delete t;
    ecc_destroy(ctx);
    if (failures) {
        std::printf("%d check(s) failed\n", failures);
        return 1;
    }
    std::printf("OK launcher swap: computeDerivLineIntegrals, epipolarConsistency and cuda_computeLineIntegrals resolved by libecc_b200.so, results identical to the C ABI\n");
    return 0;
}
