// RadonIntermediate.h -- facade of EpipolarConsistency::RadonIntermediate with the reference's public interface
// (LibEpipolarConsistency/RadonIntermediate.h:18-128, .cpp:17-211) on top of libecc_b200's C ABI.
// Differences by design: the dtr stays in linear device memory (texture objects are created over it inside the
// library, no cudaArray copies), all GPU work goes through ecc_radon_compute, and -- because the reference's loaders build
// one RadonIntermediate per projection in a loop (Gui/InputDataRadonIntermediate.cpp:111-160) while the B200 kernel works on
// quads of projections -- the constructors only STAGE their image on the device: the Radon kernel runs when a result is
// first asked for (readback, getTexture, devicePointer, the metric's setRadonIntermediates, the destructor) or when 32 images
// wait, for all waiting images of the same geometry at once.  A loop written for the reference thereby gets the batched
// engine (ECC_INTERP_HYBRID_STATIC: 0.56 ms per 1240x960 projection instead of 1.47 for one image at a time) without a
// change; results do not depend on how the images were batched.  RadonIntermediate::setDeferredCompute(false) restores
// "compute in the constructor".
#ifndef ECC_FACADE_RADON_INTERMEDIATE_H
#define ECC_FACADE_RADON_INTERMEDIATE_H

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "NRRD.h"
#include "UtilsCuda.h"

namespace EpipolarConsistency {

class RadonIntermediate;

namespace detail {

// Line (l0, l1, l2) relative to the image centre -> sample location in a Radon intermediate: angle of the normal over pi
// in [0, 2), signed distance over the t range + 1/2; angles beyond pi map to the point-mirrored half.  Returns true when the
// mirror was used.  Same arithmetic as the reference's lineToSampleDtr (EpipolarConsistencyCommon.hxx:152-171), float
// throughout, pi as its float literal.
template <typename Line> inline bool lineToSampleDtr(Line& line, float range_t)
{
    const float Pi = 3.14159265359f;
    const float length = std::sqrt((float)(line[0] * line[0] + line[1] * line[1]));
    line[0] = std::atan2((float)line[1], (float)line[0]) / Pi;
    if (line[0] < 0) line[0] += 2;
    line[1] = -(line[2] / length) / range_t + 0.5f;
    if (line[0] > 1) {
        line[0] = line[0] - 1.f;
        line[1] = 1.f - line[1];
        return true;
    }
    return false;
}

// Several intermediates in one device allocation (what a batched launch writes); freed with the last one that refers to it.
struct DtrBlock {
    float* base;
    int refs;
};

// Images of one geometry waiting for their Radon intermediates.
struct RadonBatch {
    int n_x, n_y, n_alpha, n_t, filter, post, interp;
    float* images_d;  // [capacity][n_y][n_x]
    int capacity, count;
    std::vector<RadonIntermediate*> owners;
    RadonBatch() : n_x(0), n_y(0), n_alpha(0), n_t(0), filter(0), post(0), interp(0), images_d(0x0), capacity(0), count(0) {}
};
inline RadonBatch& radon_batch()
{
    static RadonBatch b;
    return b;
}

}  // namespace detail

/// Compute derivative in t-direction of Radon transform of x-ray projection data.
class RadonIntermediate {
public:
    /// Filter applied to Radon transform.
    enum Filter { Derivative = 0, Ramp = 1, None = 2 };
    /// Function applied to each value in the Radon transform.
    enum PostProcess { Identity = 0, SquareRoot = 1, Logarithm = 2 };

    /// Ctor computes the Radon intermediate from a host image (RadonIntermediate.cpp:17-31); see the note on staging above.
    RadonIntermediate(const NRRD::ImageView<float>& projectionData, int size_alpha, int size_t, Filter filter,
                      PostProcess post_process)
        : m_dev(0x0), m_block(0x0), m_tex(0x0), m_pending(false), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(filter), n_x(0),
          n_y(0), n_t(size_t), n_alpha(size_alpha), m_interp(defaultInterpolation())
    {
        compute((const float*)projectionData, projectionData.size(0), projectionData.size(1), size_alpha, size_t, filter, post_process);
    }

    /// Ctor computes the Radon intermediate from a GPU-resident image (RadonIntermediate.cpp:33-45).
    RadonIntermediate(const UtilsCuda::BindlessTexture2D<float>& projectionData, int size_alpha, int size_t, Filter filter,
                      PostProcess post_process)
        : m_dev(0x0), m_block(0x0), m_tex(0x0), m_pending(false), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(filter), n_x(0),
          n_y(0), n_t(size_t), n_alpha(size_alpha), m_interp(defaultInterpolation())
    {
        compute(projectionData.device, projectionData.size[0], projectionData.size[1], size_alpha, size_t, filter, post_process);
    }

    /// Ctor loads a previously saved dtr (RadonIntermediate.cpp:47-67).
    RadonIntermediate(const std::string path)
        : m_dev(0x0), m_block(0x0), m_tex(0x0), m_pending(false), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(Derivative),
          n_x(0), n_y(0), n_t(0), n_alpha(0), m_interp(defaultInterpolation())
    {
        m_raw_cpu.load(path);
        if (!m_raw_cpu) {
            std::cerr << "Failed to load " << path << std::endl;
            return;
        }
        readPropertiesFromMeta(m_raw_cpu.meta_info);
        n_alpha = m_raw_cpu.size(0);
        n_t = m_raw_cpu.size(1);
        upload();
    }

    /// Ctor uses existing CPU memory (RadonIntermediate.cpp:69-80).
    RadonIntermediate(const NRRD::ImageView<float>& radon_intermediate_image)
        : m_dev(0x0), m_block(0x0), m_tex(0x0), m_pending(false), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(Derivative),
          n_x(0), n_y(0), n_t(0), n_alpha(0), m_interp(defaultInterpolation())
    {
        replaceRadonIntermediateData(radon_intermediate_image);
    }

    ~RadonIntermediate()
    {
        if (m_pending) flushPending();  // the batch refers to this object
        delete m_tex;
        release();
    }

    /// Get relevant parameters from meta dictionary (RadonIntermediate.cpp:82-94).
    void readPropertiesFromMeta(std::map<std::string, std::string> dict)
    {
        m_bin_size_angle = stringTo<double>(dict["Bin Size/Angle"]);
        m_bin_size_distance = stringTo<double>(dict["Bin Size/Distance"]);
        n_x = stringTo<int>(dict["Original Image/Width"]);
        n_y = stringTo<int>(dict["Original Image/Height"]);
        if (dict["Filter"] == "Ramp") m_filter = Ramp;
        else if (dict["Filter"] == "Derivative") m_filter = Derivative;
        else m_filter = None;
    }

    /// Store relevant parameters in meta dictionary (RadonIntermediate.cpp:96-103).
    void writePropertiesToMeta(std::map<std::string, std::string>& dict) const
    {
        dict["Bin Size/Angle"] = toString(m_bin_size_angle);
        dict["Bin Size/Distance"] = toString(m_bin_size_distance);
        dict["Original Image/Width"] = toString(n_x);
        dict["Original Image/Height"] = toString(n_y);
        dict["Filter"] = m_filter == Derivative ? "Derivative" : (m_filter == Ramp ? "Ramp" : "None");
    }

    /// Update Radon intermediate data with CPU memory (RadonIntermediate.cpp:105-123).
    void replaceRadonIntermediateData(const NRRD::ImageView<float>& radon_intermediate_image)
    {
        if (m_pending) flushPending();
        m_raw_cpu.clone(radon_intermediate_image);
        n_alpha = radon_intermediate_image.size(0);
        n_t = radon_intermediate_image.size(1);
        readPropertiesFromMeta(radon_intermediate_image.meta_info);
        const double diagonal = std::sqrt((double)n_y * n_y + (double)n_x * n_x);
        if (n_t > 0 && diagonal > 0) m_bin_size_distance = diagonal / n_t;
        upload();
    }

    RadonIntermediate::Filter getFilter() const { return m_filter; }

    /// If true the Radon intermediate is odd: dtr(alpha+Pi,t) = -dtr(alpha,-t).
    bool isDerivative() const { return m_filter == Derivative; }

    /// Read the dtr back to CPU memory and set the meta info (RadonIntermediate.cpp:148-163).
    void readback(bool gpu_memory_only = false)
    {
        if (gpu_memory_only) {
            m_raw_cpu.set(0x0, 0);
            return;
        }
        ensureComputed();
        if (!m_dev) return;
        m_raw_cpu.set(n_alpha, n_t);
        writePropertiesToMeta(m_raw_cpu.meta_info);
        detail::check(ecc_copy(detail::shared_context(), (float*)m_raw_cpu, m_dev, sizeof(float) * (size_t)n_alpha * n_t),
                      detail::shared_context(), "ecc_copy");
    }

    /// Keep only the GPU copy (RadonIntermediate.cpp:141-146).
    void clearRawData() { m_raw_cpu.set(0x0, 0); }

    /// 0: angle 1: distance
    int getRadonBinNumber(int dim) const { return dim ? n_t : n_alpha; }
    /// 0: width 1: height
    int getOriginalImageSize(int dim) const { return dim ? n_y : n_x; }
    /// 0: angle 1: distance
    double getRadonBinSize(int dim = 1) const { return dim ? m_bin_size_distance : m_bin_size_angle; }

    /// The GPU-resident dtr.  (The reference converts to a cudaArray texture here and frees the linear copy; here the linear
    /// copy is what the metric samples, and the handle creates its `tex` / `array` when somebody asks for them.)
    UtilsCuda::BindlessTexture2D<float>* getTexture()
    {
        ensureComputed();
        if (!m_dev) return 0x0;
        if (!m_tex) m_tex = UtilsCuda::BindlessTexture2D<float>::view(n_alpha, n_t, m_dev, true);
        return m_tex;
    }

    /// Access to raw data on CPU (may be an invalid image, see readback).
    NRRD::ImageView<float>& data() { return m_raw_cpu; }
    const NRRD::ImageView<float>& data() const { return m_raw_cpu; }

    /// Sample given a line in the original image, relative to the image centre; the line is overwritten with the sample
    /// location.  Must call readback() before use.  CPU code of the reference for plots (RadonIntermediate.h:85-105): the
    /// location comes from lineToSampleDtr, which already folds angles beyond pi; the second test on the folded angle never
    /// fires and the sign of derivative data is NOT flipped here, as upstream (the metric's own device lookup does flip it).
    template <typename Line> inline float sample(Line& line)
    {
        const float range_t = (float)m_bin_size_distance * getRadonBinNumber(1);
        detail::lineToSampleDtr(line, range_t);
        if (line[0] > 1) {
            line[0] = line[0] - 1.f;
            line[1] = 1.f - line[1];
            return m_filter == Derivative ? -tex2D(line[0], line[1]) : +tex2D(line[0], line[1]);
        }
        return tex2D(line[0], line[1]);
    }

    /// Sample the Radon intermediate in texture coordinates on the CPU copy, with the reference's texel mapping (n-1) s
    /// (RadonIntermediate.h:108; the device lookup of the metric maps n s - 1/2).  Must call readback() before use.
    inline float tex2D(float s, float t) { return (float)m_raw_cpu((double)((n_alpha - 1) * s), (double)((n_t - 1) * t)); }

    /// Engine used by compute(): ECC_INTERP_HYBRID_STATIC (the benchmarked engine: texture-filter arithmetic through both
    /// sampling pipes, reproducible, bins within 5e-5 of the peak of the reference's kernel; the default),
    /// ECC_INTERP_TEXTURE (bit-identical to the reference's kernel), ECC_INTERP_HYBRID (run-time work queue) or
    /// ECC_INTERP_EXACT (fp32 weights).  Affects objects constructed afterwards.
    static int& defaultInterpolation()
    {
        static int interp = [] {
            const char* e = std::getenv("ECC_FACADE_RADON");
            if (!e) return (int)ECC_INTERP_HYBRID_STATIC;
            const std::string s(e);
            if (s == "texture") return (int)ECC_INTERP_TEXTURE;
            if (s == "hybrid") return (int)ECC_INTERP_HYBRID;
            if (s == "exact") return (int)ECC_INTERP_EXACT;
            return (int)ECC_INTERP_HYBRID_STATIC;
        }();
        return interp;
    }
    static void setDefaultInterpolation(int interp) { defaultInterpolation() = interp; }
    /// Engine of THIS object; only meaningful before its computation has run (kept for round-1 callers).
    void setInterpolation(int interp) { m_interp = interp; }

    /// false: every constructor runs its own Radon kernel at once, as the reference does (default: true, see the top of the file).
    static void setDeferredCompute(bool on)
    {
        if (!on) flushPending();
        deferred() = on;
    }
    /// Runs the Radon kernel for all staged images now.
    static void flushPending()
    {
        detail::RadonBatch& B = detail::radon_batch();
        if (B.count == 0) return;
        ecc_context* ctx = detail::shared_context();
        const size_t len = (size_t)B.n_alpha * B.n_t;
        detail::DtrBlock* block = new detail::DtrBlock();
        block->refs = 0;
        void* p = 0x0;
        detail::check(ecc_device_alloc(ctx, sizeof(float) * len * B.count, &p), ctx, "ecc_device_alloc");
        block->base = (float*)p;
        detail::check(ecc_radon_compute(ctx, B.images_d, B.count, B.n_x, B.n_y, B.n_alpha, B.n_t, B.filter, B.post, B.interp, block->base), ctx,
                      "ecc_radon_compute");
        detail::check(ecc_synchronize(ctx), ctx, "ecc_synchronize");
        for (int k = 0; k < B.count; k++) {
            RadonIntermediate* o = B.owners[k];
            o->m_dev = block->base + len * k;
            o->m_block = block;
            o->m_pending = false;
            block->refs++;
        }
        B.count = 0;
        B.owners.clear();
    }

    /// Device pointer of the dtr (n_t rows of n_alpha floats); used by MetricRadonIntermediate.
    const float* devicePointer()
    {
        ensureComputed();
        return m_dev;
    }

protected:
    NRRD::Image<float> m_raw_cpu;                 //< Optional data on CPU.
    float* m_dev;                                 //< Data on GPU (linear, 512-byte aligned).
    detail::DtrBlock* m_block;                    //< Allocation m_dev lives in when it came out of a batch (else m_dev is its own).
    UtilsCuda::BindlessTexture2D<float>* m_tex;   //< Resident handle handed out by getTexture().
    bool m_pending;                               //< The image is staged, the Radon kernel has not run yet.
    double m_bin_size_angle;
    double m_bin_size_distance;
    Filter m_filter;
    int n_x, n_y, n_t, n_alpha;
    int m_interp;

    static bool& deferred()
    {
        static bool on = std::getenv("ECC_FACADE_EAGER") == 0x0;
        return on;
    }
    void ensureComputed()
    {
        if (m_pending) flushPending();
    }
    void release()
    {
        if (m_block) {
            if (--m_block->refs == 0) {
                ecc_device_free(detail::shared_context(), m_block->base);
                delete m_block;
            }
        } else if (m_dev) {
            ecc_device_free(detail::shared_context(), m_dev);
        }
        m_block = 0x0;
        m_dev = 0x0;
    }
    void alloc()
    {
        release();
        delete m_tex;
        m_tex = 0x0;
        void* p = 0x0;
        detail::check(ecc_device_alloc(detail::shared_context(), sizeof(float) * (size_t)n_alpha * n_t, &p), detail::shared_context(),
                      "ecc_device_alloc");
        m_dev = (float*)p;
    }

    void upload()
    {
        if (!m_raw_cpu) return;
        alloc();
        detail::check(ecc_copy(detail::shared_context(), m_dev, (const float*)m_raw_cpu, sizeof(float) * (size_t)n_alpha * n_t),
                      detail::shared_context(), "ecc_copy");
    }

    /// Runs the Radon kernel (RadonIntermediate.cpp:198-211) -- now, or together with the images staged next.
    void compute(const float* image, int w, int h, int size_alpha, int size_dist, Filter filter, PostProcess post_process)
    {
        n_x = w;
        n_y = h;
        n_alpha = size_alpha;
        n_t = size_dist;
        m_filter = filter;
        ecc_radon_bin_sizes(n_x, n_y, n_alpha, n_t, &m_bin_size_angle, &m_bin_size_distance);
        ecc_context* ctx = detail::shared_context();
        const size_t dtr_bytes = sizeof(float) * (size_t)n_alpha * n_t;
        // batched results are handed to the metric one pointer each: every dtr of a block must start 512-byte aligned
        if (!deferred() || dtr_bytes % 512 != 0 || n_alpha % 8 != 0) {
            alloc();
            detail::check(ecc_radon_compute(ctx, image, 1, n_x, n_y, n_alpha, n_t, (int)filter, (int)post_process, m_interp, m_dev), ctx,
                          "ecc_radon_compute");
            detail::check(ecc_synchronize(ctx), ctx, "ecc_synchronize");
            return;
        }
        detail::RadonBatch& B = detail::radon_batch();
        const bool same = B.n_x == w && B.n_y == h && B.n_alpha == size_alpha && B.n_t == size_dist && B.filter == (int)filter &&
                          B.post == (int)post_process && B.interp == m_interp;
        if (!same) {
            flushPending();
            if (B.images_d) ecc_device_free(ctx, B.images_d);
            B.images_d = 0x0;
            B.n_x = w; B.n_y = h; B.n_alpha = size_alpha; B.n_t = size_dist;
            B.filter = (int)filter; B.post = (int)post_process; B.interp = m_interp;
            B.capacity = 32;
            void* p = 0x0;
            detail::check(ecc_device_alloc(ctx, sizeof(float) * (size_t)w * h * B.capacity, &p), ctx, "ecc_device_alloc");
            B.images_d = (float*)p;
        }
        release();
        // the caller's buffer may be reused as soon as the constructor returns: the copy is complete when ecc_copy returns
        detail::check(ecc_copy(ctx, B.images_d + (size_t)w * h * B.count, image, sizeof(float) * (size_t)w * h), ctx, "ecc_copy");
        B.owners.push_back(this);
        B.count++;
        m_pending = true;
        if (B.count == B.capacity) flushPending();
    }

private:
    RadonIntermediate(const RadonIntermediate&);
    RadonIntermediate& operator=(const RadonIntermediate&);
};

/// Computing Radon intermediates with the settings of the reference's GUI section (Gui/ComputeRadonIntermediate.hxx:29-85,
/// without the GetSet plumbing): filter, post-processing, number of bins.
struct RadonIntermediateFunction {
    RadonIntermediate::Filter filter;
    RadonIntermediate::PostProcess post_process;
    struct NumberOfBins {
        int angle;     //< Size of the Radon transform in angle-direction.
        int distance;  //< Size of the Radon transform in distance-direction.
        NumberOfBins() : angle(768), distance(768) {}
    } number_of_bins;

    RadonIntermediateFunction() : filter(RadonIntermediate::Derivative), post_process(RadonIntermediate::Identity) {}

    /// Compute the Radon intermediate of img; as upstream (:70-83) the projection matrix and the pixel spacing, when given,
    /// are recorded in the IMAGE's meta info under "Original Image/Projection Matrix" / "Original Image/Pixel Spacing" (the
    /// loaders copy the image's meta info into the dtr file, Gui/InputDataRadonIntermediate.cpp:158-165).
    template <class PM> RadonIntermediate* compute(NRRD::ImageView<float>& img, PM* P, double* mm_per_px = 0x0)
    {
        RadonIntermediate* dtr = compute(img);
        if (P) img.meta_info["Original Image/Projection Matrix"] = projectionMatrixToString(*P);
        if (mm_per_px) img.meta_info["Original Image/Pixel Spacing"] = toString(*mm_per_px);
        return dtr;
    }
    RadonIntermediate* compute(NRRD::ImageView<float>& img)
    {
        return new RadonIntermediate(img, number_of_bins.angle, number_of_bins.distance, filter, post_process);
    }

    /// "[p00 p01 p02 p03; p10 ...; p20 ...] " with 12 significant digits (LibProjectiveGeometry/EigenToStr.hxx:81-88); the
    /// matrix is any type with operator()(row, col).
    template <class PM> static std::string projectionMatrixToString(const PM& in)
    {
        std::ostringstream strstr;
        strstr << std::setprecision(12) << "[" << in(0, 0) << " " << in(0, 1) << " " << in(0, 2) << " " << in(0, 3) << "; " << in(1, 0) << " "
               << in(1, 1) << " " << in(1, 2) << " " << in(1, 3) << "; " << in(2, 0) << " " << in(2, 1) << " " << in(2, 2) << " " << in(2, 3) << "] ";
        return strstr.str();
    }
};

}  // namespace EpipolarConsistency

#endif
