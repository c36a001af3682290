// tex_probe.cu -- measures, on the GPU at hand, (1) how the texture unit quantises bilinear weights
// (needed by the oracle's TEX8 model, SURVEY.md section 7 hard part 2) and (2) the throughput of
// tex2D / tex2Dgather on a CUDA array versus pitched linear memory for coherent access.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tex_probe tools/tex_probe.cu
// Run on the GPU box: ./tools/tex_probe > gpurun_out/tex_probe.txt
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void sample_x(cudaTextureObject_t tex, const float* xs, float y, int n, float* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2D<float>(tex, xs[i], y);
}
__global__ void sample_xy(cudaTextureObject_t tex, const float* xs, const float* ys, int n, float* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex2D<float>(tex, xs[i], ys[i]);
}

// throughput: every thread walks a short line, like the Radon kernel does
template <int MODE>
__global__ void walk(cudaTextureObject_t tex, int w, int h, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float x = 2.0f + (tid % 1024) * 1.03f, y = 2.0f + ((tid / 1024) % 768) * 1.01f;
    float acc = 0.f;
    for (int k = 0; k < iters; k++) {
        if (MODE == 0) acc += tex2D<float>(tex, x, y);
        else { float4 g = tex2Dgather<float4>(tex, x, y, 0); acc += g.x + g.y + g.z + g.w; }
        x += 0.55f; y += 0.36f;
        if (x > w - 2) x -= (w - 4);
        if (y > h - 2) y -= (h - 4);
    }
    out[tid] = acc;
}

template <typename T>
__global__ void walk_vec(cudaTextureObject_t tex, int w, int h, int iters, float* out)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    // lanes paired like the Radon kernel's best tiling: quads of 2x2 neighbours
    float x = 2.0f + (tid % 1024) * 1.03f, y = 2.0f + ((tid / 1024) % 768) * 1.01f;
    float acc = 0.f;
    for (int k = 0; k < iters; k++) {
        T v = tex2D<T>(tex, x, y);
        if constexpr (sizeof(T) == 8) acc += v.x + v.y;
        else acc += v.x + v.y + v.z + v.w;
        x += 0.55f; y += 0.36f;
        if (x > w - 2) x -= (w - 4);
        if (y > h - 2) y -= (h - 4);
    }
    out[tid] = acc;
}

template <typename T>
cudaTextureObject_t make_vec_tex(int w, int h, cudaArray_t* arr_out)
{
    std::vector<T> img((size_t)w * h);
    for (size_t i = 0; i < img.size(); i++) { float* p = (float*)&img[i]; for (size_t c = 0; c < sizeof(T) / 4; c++) p[c] = (float)((i * 7 + c * 13) % 251); }
    cudaChannelFormatDesc d = cudaCreateChannelDesc<T>();
    CK(cudaMallocArray(arr_out, &d, w, h));
    CK(cudaMemcpy2DToArray(*arr_out, 0, 0, img.data(), w * sizeof(T), w * sizeof(T), h, cudaMemcpyHostToDevice));
    cudaResourceDesc res = {};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = *arr_out;
    cudaTextureDesc td = {};
    td.normalizedCoords = 0;
    td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &res, &td, nullptr));
    return tex;
}

struct Tex { cudaArray_t arr = nullptr; float* lin = nullptr; size_t pitch = 0; cudaTextureObject_t tex = 0; };

Tex make_tex(const std::vector<float>& img, int w, int h, bool array, bool normalized)
{
    Tex t;
    cudaResourceDesc res = {};
    if (array) {
        cudaChannelFormatDesc d = cudaCreateChannelDesc<float>();
        CK(cudaMallocArray(&t.arr, &d, w, h, cudaArrayTextureGather));
        CK(cudaMemcpy2DToArray(t.arr, 0, 0, img.data(), w * 4, w * 4, h, cudaMemcpyHostToDevice));
        res.resType = cudaResourceTypeArray;
        res.res.array.array = t.arr;
    } else {
        CK(cudaMallocPitch(&t.lin, &t.pitch, w * 4, h));
        CK(cudaMemcpy2D(t.lin, t.pitch, img.data(), w * 4, w * 4, h, cudaMemcpyHostToDevice));
        res.resType = cudaResourceTypePitch2D;
        res.res.pitch2D.devPtr = t.lin;
        res.res.pitch2D.desc = cudaCreateChannelDesc<float>();
        res.res.pitch2D.width = w;
        res.res.pitch2D.height = h;
        res.res.pitch2D.pitchInBytes = t.pitch;
    }
    cudaTextureDesc td = {};
    td.normalizedCoords = normalized;
    td.filterMode = cudaFilterModeLinear;
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.readMode = cudaReadModeElementType;
    CK(cudaCreateTextureObject(&t.tex, &res, &td, nullptr));
    return t;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, SMs %d, texturePitchAlignment %zu, textureAlignment %zu\n", prop.name, prop.multiProcessorCount,
           prop.texturePitchAlignment, prop.textureAlignment);

    // ---------------- (1) weight quantisation -------------------------------------------------
    const int W = 2048, H = 8;
    std::vector<float> ramp((size_t)W * H);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) ramp[(size_t)y * W + x] = (float)x;  // value == texel index: tex(x) = i + alpha_q
    const int SUB = 8192;  // sub-positions per texel
    float *xs_d, *out_d;
    CK(cudaMalloc(&xs_d, SUB * 4));
    CK(cudaMalloc(&out_d, SUB * 4));
    std::vector<float> xs(SUB), out(SUB);
    for (int array = 1; array >= 0; array--)
        for (int normalized = 0; normalized <= 1; normalized++) {
            Tex t = make_tex(ramp, W, H, array, normalized);
            for (int base : {0, 3, 100, 767, 1500}) {
                for (int k = 0; k < SUB; k++) {
                    float pos = (float)base + 0.5f + (float)k / SUB;  // unnormalised coordinate, xB = base + k/SUB
                    xs[k] = normalized ? pos / (float)W : pos;
                }
                CK(cudaMemcpy(xs_d, xs.data(), SUB * 4, cudaMemcpyHostToDevice));
                sample_x<<<(SUB + 255) / 256, 256>>>(t.tex, xs_d, normalized ? 0.5f / H : 0.5f, SUB, out_d);
                CK(cudaMemcpy(out.data(), out_d, SUB * 4, cudaMemcpyDeviceToHost));
                int bad_rn = 0, bad_tr = 0, bad_rn_pos = 0, levels = 0;
                float prev = -1;
                double first_step = -1;
                for (int k = 0; k < SUB; k++) {
                    float pos = normalized ? xs[k] * (float)W : xs[k];  // what the hardware sees after its own scaling
                    float xb = pos - 0.5f;
                    float frac = xb - floorf(xb);
                    // model A: weight = round(frac*256)/256 ; model B: truncation ; model C: position rounded (same as A here)
                    float a = floorf(frac * 256.f + 0.5f) / 256.f, b = floorf(frac * 256.f) / 256.f;
                    float q = floorf(xb * 256.f + 0.5f) / 256.f;
                    float got = out[k] - floorf(xb);
                    if (fabsf(got - a) > 1e-6f) bad_rn++;
                    if (fabsf(got - b) > 1e-6f) bad_tr++;
                    if (fabsf(out[k] - q) > 1e-4f) bad_rn_pos++;
                    if (out[k] != prev) { levels++; if (levels == 2 && first_step < 0) first_step = (double)k / SUB; prev = out[k]; }
                }
                printf("quant array=%d normalized=%d base=%4d: levels=%d first_step_at_frac=%.6f (rn: %.6f, trunc: %.6f) "
                       "mismatch round-nearest=%d truncate=%d round-position=%d of %d\n",
                       array, normalized, base, levels, first_step, 0.5 / 256, 1.0 / 256, bad_rn, bad_tr, bad_rn_pos, SUB);
            }
            cudaDestroyTextureObject(t.tex);
            if (t.arr) cudaFreeArray(t.arr);
            if (t.lin) cudaFree(t.lin);
        }

    // 2-D check of the filter formula on random data: array vs pitch2D, and vs the fp32 formula with quantised weights
    {
        const int w = 768, h = 768, n = 1 << 16;
        std::vector<float> img((size_t)w * h), px(n), py(n), o1(n), o2(n);
        srand(1);
        for (auto& v : img) v = (float)rand() / RAND_MAX * 200.f - 100.f;
        for (int k = 0; k < n; k++) { px[k] = (float)rand() / RAND_MAX * (w + 2) - 1; py[k] = (float)rand() / RAND_MAX * (h + 2) - 1; }
        float *px_d, *py_d, *o_d;
        CK(cudaMalloc(&px_d, n * 4)); CK(cudaMalloc(&py_d, n * 4)); CK(cudaMalloc(&o_d, n * 4));
        CK(cudaMemcpy(px_d, px.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(py_d, py.data(), n * 4, cudaMemcpyHostToDevice));
        Tex ta = make_tex(img, w, h, true, false), tp = make_tex(img, w, h, false, false);
        sample_xy<<<n / 256, 256>>>(ta.tex, px_d, py_d, n, o_d);
        CK(cudaMemcpy(o1.data(), o_d, n * 4, cudaMemcpyDeviceToHost));
        sample_xy<<<n / 256, 256>>>(tp.tex, px_d, py_d, n, o_d);
        CK(cudaMemcpy(o2.data(), o_d, n * 4, cudaMemcpyDeviceToHost));
        double d_ap = 0, d_model = 0, d_exact = 0;
        auto tap = [&](float x, int nn, int& i0, int& i1, float& wq, float& we) {
            float xb = x - 0.5f;
            float q = floorf(xb * 256.f + 0.5f);
            float fl = floorf(q / 256.f);
            wq = (q - fl * 256.f) / 256.f;
            we = xb - floorf(xb);
            int i = (int)fl;
            i0 = i < 0 ? 0 : (i > nn - 1 ? nn - 1 : i);
            i1 = i + 1 < 0 ? 0 : (i + 1 > nn - 1 ? nn - 1 : i + 1);
        };
        for (int k = 0; k < n; k++) {
            d_ap = fmax(d_ap, fabs(o1[k] - o2[k]));
            int x0, x1, y0, y1; float a, b, ae, be;
            tap(px[k], w, x0, x1, a, ae); tap(py[k], h, y0, y1, b, be);
            float m = (1 - a) * (1 - b) * img[(size_t)y0 * w + x0] + a * (1 - b) * img[(size_t)y0 * w + x1] +
                      (1 - a) * b * img[(size_t)y1 * w + x0] + a * b * img[(size_t)y1 * w + x1];
            d_model = fmax(d_model, fabs(o1[k] - m));
            // exact weights with floor(xb) cells, for scale
            int ex0 = (int)floorf(px[k] - 0.5f), ey0 = (int)floorf(py[k] - 0.5f);
            auto cl = [](int v, int nn) { return v < 0 ? 0 : (v > nn - 1 ? nn - 1 : v); };
            float e = (1 - ae) * (1 - be) * img[(size_t)cl(ey0, h) * w + cl(ex0, w)] + ae * (1 - be) * img[(size_t)cl(ey0, h) * w + cl(ex0 + 1, w)] +
                      (1 - ae) * be * img[(size_t)cl(ey0 + 1, h) * w + cl(ex0, w)] + ae * be * img[(size_t)cl(ey0 + 1, h) * w + cl(ex0 + 1, w)];
            d_exact = fmax(d_exact, fabs(o1[k] - e));
        }
        {   // diagnosis: how many samples disagree with the model, split by region, plus a few examples
            int bad = 0, bad_interior = 0, shown = 0;
            for (int k = 0; k < n; k++) {
                int x0, x1, y0, y1; float a, b, ae, be;
                tap(px[k], w, x0, x1, a, ae); tap(py[k], h, y0, y1, b, be);
                float m = (1 - a) * (1 - b) * img[(size_t)y0 * w + x0] + a * (1 - b) * img[(size_t)y0 * w + x1] +
                          (1 - a) * b * img[(size_t)y1 * w + x0] + a * b * img[(size_t)y1 * w + x1];
                // alternative evaluation orders
                float top = img[(size_t)y0 * w + x0] + a * (img[(size_t)y0 * w + x1] - img[(size_t)y0 * w + x0]);
                float bot = img[(size_t)y1 * w + x0] + a * (img[(size_t)y1 * w + x1] - img[(size_t)y1 * w + x0]);
                float m2 = top + b * (bot - top);
                if (fabsf(o1[k] - m) > 1e-3f) {
                    bad++;
                    bool interior = px[k] > 1 && px[k] < w - 1 && py[k] > 1 && py[k] < h - 1;
                    if (interior) bad_interior++;
                    if (shown < 12) { shown++; printf("  mismatch: x=%.6f y=%.6f got=%.6f model=%.6f lerp-model=%.6f a=%.6f b=%.6f interior=%d\n", px[k], py[k], o1[k], m, m2, a, b, (int)interior); }
                }
            }
            printf("filter2d mismatches > 1e-3: %d of %d (%d interior)\n", bad, n, bad_interior);
        }
        printf("filter2d (values in [-100,100]): max|array - pitch2D| = %.3g, max|array - quantised model| = %.3g, "
               "max|array - exact fp32| = %.3g\n", d_ap, d_model, d_exact);
    }

    // 2-D diagnosis with structured images: T=x, T=y at random (x,y); raw dump for a random image
    {
        const int w = 768, h = 768, n = 4096;
        std::vector<float> ix((size_t)w * h), iy((size_t)w * h), ir((size_t)w * h), px(n), py(n), o(n);
        srand(7);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) { ix[(size_t)y * w + x] = (float)x; iy[(size_t)y * w + x] = (float)y; ir[(size_t)y * w + x] = (float)(rand() % 1000); }
        for (int k = 0; k < n; k++) { px[k] = 2.f + (float)rand() / RAND_MAX * (w - 4); py[k] = 2.f + (float)rand() / RAND_MAX * (h - 4); }
        float *px_d, *py_d, *o_d;
        CK(cudaMalloc(&px_d, n * 4)); CK(cudaMalloc(&py_d, n * 4)); CK(cudaMalloc(&o_d, n * 4));
        CK(cudaMemcpy(px_d, px.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(py_d, py.data(), n * 4, cudaMemcpyHostToDevice));
        auto rn8 = [](float v) { float xb = v - 0.5f; return floorf(xb * 256.f + 0.5f) / 256.f; };
        for (int which = 0; which < 2; which++) {
            Tex t = make_tex(which == 0 ? ix : iy, w, h, true, false);
            sample_xy<<<n / 256, 256>>>(t.tex, px_d, py_d, n, o_d);
            CK(cudaMemcpy(o.data(), o_d, n * 4, cudaMemcpyDeviceToHost));
            int bad = 0, shown = 0;
            double maxerr = 0;
            for (int k = 0; k < n; k++) {
                float want = rn8(which == 0 ? px[k] : py[k]);
                double e = fabs(o[k] - want);
                maxerr = fmax(maxerr, e);
                if (e > 1e-4) { bad++; if (shown++ < 6) printf("  ramp-%c mismatch: x=%.6f y=%.6f got=%.6f rn8=%.6f exact=%.6f\n", which ? 'y' : 'x', px[k], py[k], o[k], want, (which == 0 ? px[k] : py[k]) - 0.5f); }
            }
            printf("ramp-%c at random (x,y): %d of %d differ from round-to-nearest-1/256 position (max %.3g)\n", which ? 'y' : 'x', bad, n, maxerr);
        }
        Tex t = make_tex(ir, w, h, true, false);
        sample_xy<<<n / 256, 256>>>(t.tex, px_d, py_d, n, o_d);
        CK(cudaMemcpy(o.data(), o_d, n * 4, cudaMemcpyDeviceToHost));
        {   // candidate joint-weight models against the random image
            int bad_outer = 0, bad_prod_rn = 0, bad_prod_tr = 0;
            for (int k = 0; k < n; k++) {
                float xb = px[k] - 0.5f, yb = py[k] - 0.5f;
                float qx = floorf(xb * 256.f + 0.5f), qy = floorf(yb * 256.f + 0.5f);
                int i = (int)floorf(qx / 256.f), j = (int)floorf(qy / 256.f);
                int a = (int)(qx - i * 256.f), b = (int)(qy - j * 256.f);
                float T00 = ir[(size_t)j * w + i], T10 = ir[(size_t)j * w + i + 1], T01 = ir[(size_t)(j + 1) * w + i], T11 = ir[(size_t)(j + 1) * w + i + 1];
                float outer = ((256 - a) * (256 - b) * T00 + a * (256 - b) * T10 + (256 - a) * b * T01 + a * b * T11) / 65536.f;
                auto joint = [&](int w11) { return ((256 - a - b + w11) * T00 + (a - w11) * T10 + (b - w11) * T01 + w11 * T11) / 256.f; };
                if (fabsf(o[k] - outer) > 2e-3f) bad_outer++;
                if (fabsf(o[k] - joint((a * b + 128) >> 8)) > 2e-3f) bad_prod_rn++;
                if (fabsf(o[k] - joint((a * b) >> 8)) > 2e-3f) bad_prod_tr++;
            }
            printf("joint weights on a random image (values 0..999), mismatches > 2e-3 of %d: outer product of 8-bit weights %d, "
                   "w11 = round(a*b/256) %d, w11 = trunc(a*b/256) %d\n", n, bad_outer, bad_prod_rn, bad_prod_tr);
        }
        for (int k = 0; k < 10; k++) {
            float xb = px[k] - 0.5f, yb = py[k] - 0.5f;
            int i = (int)floorf(xb), j = (int)floorf(yb);
            printf("  raw: x=%.7f y=%.7f got=%.6f T00=%.0f T10=%.0f T01=%.0f T11=%.0f fracx*256=%.4f fracy*256=%.4f\n", px[k], py[k], o[k],
                   ir[(size_t)j * w + i], ir[(size_t)j * w + i + 1], ir[(size_t)(j + 1) * w + i], ir[(size_t)(j + 1) * w + i + 1],
                   (xb - floorf(xb)) * 256.f, (yb - floorf(yb)) * 256.f);
        }
    }

    // normalised coordinates on non-power-of-two sizes: how does the hardware scale them to texels?
    for (int w : {768, 192, 1000}) {
        const int h = 8, n = 1 << 15;
        std::vector<float> img((size_t)w * h), pa(n), o_arr(n), o_pit(n);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) img[(size_t)y * w + x] = (float)x;
        srand(11);
        for (int k = 0; k < n; k++) pa[k] = (1.0f + (float)rand() / RAND_MAX * (w - 2)) / (float)w;
        float *pa_d, *o_d;
        CK(cudaMalloc(&pa_d, n * 4)); CK(cudaMalloc(&o_d, n * 4));
        CK(cudaMemcpy(pa_d, pa.data(), n * 4, cudaMemcpyHostToDevice));
        Tex ta = make_tex(img, w, h, true, true), tp = make_tex(img, w, h, false, true);
        sample_x<<<n / 256, 256>>>(ta.tex, pa_d, 0.5f / h, n, o_d);
        CK(cudaMemcpy(o_arr.data(), o_d, n * 4, cudaMemcpyDeviceToHost));
        sample_x<<<n / 256, 256>>>(tp.tex, pa_d, 0.5f / h, n, o_d);
        CK(cudaMemcpy(o_pit.data(), o_d, n * 4, cudaMemcpyDeviceToHost));
        int d_ap = 0, bad_f32 = 0, bad_f64 = 0, bad_f32_pit = 0, bad_fma = 0;
        for (int k = 0; k < n; k++) {
            if (o_arr[k] != o_pit[k]) d_ap++;
            float xf = pa[k] * (float)w;                     // fp32 product
            float m1 = floorf((xf - 0.5f) * 256.f + 0.5f) / 256.f;
            double xd = (double)pa[k] * w - 0.5;             // exact product
            float m2 = (float)(floor(xd * 256.0 + 0.5) / 256.0);
            float m3 = floorf(fmaf(pa[k], (float)w * 256.f, -128.f) + 0.5f) / 256.f;  // one rounding
            if (o_arr[k] != m1) bad_f32++;
            if (o_arr[k] != m2) bad_f64++;
            if (o_arr[k] != m3) bad_fma++;
            if (o_pit[k] != m1) bad_f32_pit++;
        }
        printf("normalised coords, width %d, %d random positions: array!=pitch2D %d; array vs fp32(a*N) %d, vs exact(a*N) %d, vs fma %d; pitch2D vs fp32(a*N) %d\n",
               w, n, d_ap, bad_f32, bad_f64, bad_fma, bad_f32_pit);
        cudaFree(pa_d); cudaFree(o_d);
    }

    // ---------------- (2) throughput ------------------------------------------------------------
    {
        const int w = 1240, h = 960;
        std::vector<float> img((size_t)w * h, 1.0f);
        Tex ta = make_tex(img, w, h, true, false), tp = make_tex(img, w, h, false, false);
        const int threads = prop.multiProcessorCount * 2048, iters = 2000;
        float* o_d;
        CK(cudaMalloc(&o_d, (size_t)threads * 4));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&](const char* name, int mode, cudaTextureObject_t tex) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) walk<0><<<threads / 256, 256>>>(tex, w, h, iters, o_d);
                else walk<1><<<threads / 256, 256>>>(tex, w, h, iters, o_d);
                cudaEventRecord(e1);
                CK(cudaEventSynchronize(e1));
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fetches = (double)threads * iters;
            int clk_khz = 0;
            cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
            printf("throughput %-22s: %.3f ms, %.3e fetches/s, %.2f fetches/clk/SM at max clock %d MHz\n", name, ms,
                   fetches / (ms * 1e-3), fetches / (ms * 1e-3) / prop.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1000);
        };
        {
            cudaArray_t a2, a4;
            cudaTextureObject_t t2 = make_vec_tex<float2>(w, h, &a2), t4 = make_vec_tex<float4>(w, h, &a4);
            int clk_khz = 0;
            cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
            for (int ch = 2; ch <= 4; ch += 2) {
                for (int rep = 0; rep < 2; rep++) {
                    cudaEventRecord(e0);
                    if (ch == 2) walk_vec<float2><<<threads / 256, 256>>>(t2, w, h, iters, o_d);
                    else walk_vec<float4><<<threads / 256, 256>>>(t4, w, h, iters, o_d);
                    cudaEventRecord(e1);
                    CK(cudaEventSynchronize(e1));
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                double fetches = (double)threads * iters;
                printf("throughput tex2D float%d array    : %.3f ms, %.3e fetches/s = %.3e channel-samples/s, %.2f fetches/clk/SM at max clock\n", ch, ms,
                       fetches / (ms * 1e-3), ch * fetches / (ms * 1e-3), fetches / (ms * 1e-3) / prop.multiProcessorCount / (clk_khz * 1e3));
            }
        }
        run("tex2D array", 0, ta.tex);
        run("tex2D pitch2D", 0, tp.tex);
        run("tex2Dgather array", 1, ta.tex);
    }
    return 0;
}
