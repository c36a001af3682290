// ref_host_wrap.cpp -- C entry points around the REFERENCE's own host/device-shared headers.
// Ours: only this wrapper.  The headers are included from the reference tree where they lie
// (oracle/Makefile passes -I /root/reference/code); nothing of them is copied into this repo.
// The resulting oracle/_ref/libecc_ref_host.so is used by tests/ to pin oracle/ecc_oracle.cpp.
#include <cmath>
#include <cstdlib>
using std::abs;  // culaut/xgeinv.hxx calls unqualified abs() on floating types
#include <LibUtilsCuda/culaut/xprojectionmatrix.hxx>
#include <LibEpipolarConsistency/EpipolarConsistencyCommon.hxx>

extern "C" {

void ref_get_ij(int k, int n, int* i, int* j)
{
    short si, sj;
    get_ij(k, (short)n, si, sj);
    *i = si;
    *j = sj;
}

void ref_pinv_transpose(const double* P, float* PinvT)
{
    culaut::projection_matrix_pseudoinverse_transpose<double, float>(P, PinvT);
}

void ref_source_position(const double* P, float* C)
{
    culaut::projection_matrix_source_position<double, float>(P, C);
}

void ref_compute_k01(float half_nu, float half_nv, float* C0, float* C1, float* P0invT,
                     float* P1invT, float radius, float num_samples, float dkappa, float* K0,
                     float* K1)
{
    computeK01(half_nu, half_nv, C0, C1, P0invT, P1invT, radius, num_samples, dkappa, K0, K1);
}

int ref_line_to_sample(float* line, float range_t)
{
    return lineToSampleDtr(line, range_t) ? 1 : 0;
}

}  // extern "C"

// ---- the reference's own NRRD reader / writer (HeaderOnly/NRRD/nrrd_image.hxx), to cross-check the file layout
#include <NRRD/nrrd_image.hxx>
#include <cstring>

extern "C" {

// Loads `path` with the reference's NRRD::Image<float>.  Returns 0 on failure, else the number of pixels; fills
// sizes and the value of one meta key.
int ref_nrrd_load(const char* path, float* out, int max_len, int* w, int* h, const char* key, char* value, int value_cap)
{
    NRRD::Image<float> img;
    if (!img.load(path)) return 0;
    *w = img.size(0);
    *h = img.size(1);
    const int n = img.length();
    if (n > max_len) return 0;
    for (int i = 0; i < n; i++) out[i] = img[i];
    std::string v = img.meta_info[key];
    std::strncpy(value, v.c_str(), value_cap - 1);
    value[value_cap - 1] = 0;
    return n;
}

// Saves a w x h float image with one meta key using the reference's writer.
int ref_nrrd_save(const char* path, const float* data, int w, int h, const char* key, const char* value)
{
    NRRD::Image<float> img(w, h);
    for (int i = 0; i < w * h; i++) img[i] = data[i];
    img.meta_info[key] = value;
    return img.save(path) ? 1 : 0;
}

}  // extern "C"

// ---- pre-processing pieces the reference keeps in headers: the separable Gaussian low-pass
// (HeaderOnly/NRRD/nrrd_lowpass.hxx: gaussianKernel :18-33, convolve2D :46-79 incl. its tap range -k .. k-1, lowpass2D :186-192)
// and the feathering weight weighting() (LibEpipolarConsistency/EpipolarConsistencyCommon.hxx:30-35).  PreProccess.cpp itself
// needs GetSet and Eigen and cannot be compiled here; its four border loops are restated below around the reference's own
// weighting(), expression for expression (Gui/PreProccess.cpp:88-108).
#include <NRRD/nrrd_lowpass.hxx>

extern "C" {

int ref_gaussian_kernel(double sigma, int k, double* out)
{
    std::vector<double> g = NRRD::gaussianKernel(sigma, k);
    for (size_t i = 0; i < g.size(); i++) out[i] = g[i];
    return (int)g.size();
}

// In place, exactly what PreProccess::process calls (Gui/PreProccess.cpp:139-140).
void ref_lowpass2d(float* data, int w, int h, double sigma, int k)
{
    NRRD::ImageView<float> img(w, h, 1, data);
    NRRD::lowpass2D(img, sigma, k);
}

float ref_weighting(float x) { return weighting(x); }

// zero[4] / feather[4]: left, right, bottom, top (Gui/PreProccess.cpp:88-108)
void ref_border(float* data, int w, int h, const int* zero, const int* feather)
{
    NRRD::ImageView<float> img(w, h, 1, data);
    for (int y = 0; y < img.size(1); y++)
        for (int b = 0; b < zero[0] + feather[0]; b++)
            img.pixel(b, y, 0) *= b <= zero[0] ? 0 : (float)weighting(1 - (float)(b - zero[0]) / feather[0]);
    for (int y = 0; y < img.size(1); y++)
        for (int b = 1; b <= zero[1] + feather[1]; b++)
            img.pixel(img.size(0) - b, y, 0) *= b <= zero[1] ? 0 : (float)weighting(1 - (float)(b - zero[1]) / feather[1]);
    for (int b = 1; b <= zero[2] + feather[2]; b++)
        for (int x = 0; x < img.size(0); x++)
            img.pixel(x, img.size(1) - b, 0) *= b <= zero[2] ? 0 : (float)weighting(1 - (float)(b - zero[2]) / feather[2]);
    for (int b = 0; b < zero[3] + feather[3]; b++)
        for (int x = 0; x < img.size(0); x++)
            img.pixel(x, b, 0) *= b <= zero[3] ? 0 : (float)weighting(1 - (float)(b - zero[3]) / feather[3]);
}

}  // extern "C"

// ---- fan-beam weighting of the direct metric: the reference's own LinePerspectivity (RectifiedFBCC.h, the parts that do not
// need Eigen) and the per-sample weight exactly as kernel_computeLineIntegrals forms it (EpipolarConsistencyDirect.cu:86-93).
#include <LibEpipolarConsistency/RectifiedFBCC.h>

extern "C" {

float ref_fbcc_transform(const float* abcd, float t) { return LinePerspectivity((float*)abcd).transform(t); }
float ref_fbcc_inverse(const float* abcd, float t) { return LinePerspectivity((float*)abcd).inverse(t); }
float ref_fbcc_derivative(const float* abcd, float t) { return LinePerspectivity((float*)abcd).derivative(t); }
float ref_fbcc_weight(const float* rec8, float t)
{
    FBCC_weighting_info fbcc = *((FBCC_weighting_info*)rec8);
    float u_prime = fbcc.phi.transform(t) - fbcc.t_prime_ak;
    float fbcc_weight = fbcc.phi.derivative(t) / sqrtf(u_prime * u_prime + fbcc.d_l_kappa_C_sq);
    return fbcc_weight;
}
int ref_fbcc_record_floats() { return (int)(sizeof(FBCC_weighting_info) / sizeof(float)); }

}  // extern "C"

