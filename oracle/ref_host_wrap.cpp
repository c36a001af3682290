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
