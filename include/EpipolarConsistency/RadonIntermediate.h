// RadonIntermediate.h -- facade of EpipolarConsistency::RadonIntermediate with the reference's public interface
// (LibEpipolarConsistency/RadonIntermediate.h:18-128, .cpp:17-211) on top of libecc_b200's C ABI.
// Differences by design: the dtr stays in linear device memory (texture objects are created over it inside the
// library, no cudaArray copies), and all GPU work goes through ecc_radon_compute.
#ifndef ECC_FACADE_RADON_INTERMEDIATE_H
#define ECC_FACADE_RADON_INTERMEDIATE_H

#include <cmath>
#include <map>
#include <string>

#include "NRRD.h"
#include "UtilsCuda.h"

namespace EpipolarConsistency {

/// Compute derivative in t-direction of Radon transform of x-ray projection data.
class RadonIntermediate {
public:
    /// Filter applied to Radon transform.
    enum Filter { Derivative = 0, Ramp = 1, None = 2 };
    /// Function applied to each value in the Radon transform.
    enum PostProcess { Identity = 0, SquareRoot = 1, Logarithm = 2 };

    /// Ctor computes the Radon intermediate right away from a host image (RadonIntermediate.cpp:17-31).
    RadonIntermediate(const NRRD::ImageView<float>& projectionData, int size_alpha, int size_t, Filter filter,
                      PostProcess post_process)
        : m_dev(0x0), m_tex(0x0), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(filter), n_x(0), n_y(0),
          n_t(size_t), n_alpha(size_alpha), m_interp(ECC_INTERP_TEXTURE)
    {
        compute((const float*)projectionData, projectionData.size(0), projectionData.size(1), size_alpha, size_t, filter, post_process);
    }

    /// Ctor computes the Radon intermediate right away from a GPU-resident image (RadonIntermediate.cpp:33-45).
    RadonIntermediate(const UtilsCuda::BindlessTexture2D<float>& projectionData, int size_alpha, int size_t, Filter filter,
                      PostProcess post_process)
        : m_dev(0x0), m_tex(0x0), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(filter), n_x(0), n_y(0),
          n_t(size_t), n_alpha(size_alpha), m_interp(ECC_INTERP_TEXTURE)
    {
        compute(projectionData.device, projectionData.size[0], projectionData.size[1], size_alpha, size_t, filter, post_process);
    }

    /// Ctor loads a previously saved dtr (RadonIntermediate.cpp:47-67).
    RadonIntermediate(const std::string path)
        : m_dev(0x0), m_tex(0x0), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(Derivative), n_x(0), n_y(0), n_t(0),
          n_alpha(0), m_interp(ECC_INTERP_TEXTURE)
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
        : m_dev(0x0), m_tex(0x0), m_bin_size_angle(0), m_bin_size_distance(0), m_filter(Derivative), n_x(0), n_y(0), n_t(0),
          n_alpha(0), m_interp(ECC_INTERP_TEXTURE)
    {
        replaceRadonIntermediateData(radon_intermediate_image);
    }

    ~RadonIntermediate()
    {
        delete m_tex;
        if (m_dev) ecc_device_free(detail::shared_context(), m_dev);
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

    /// The GPU-resident dtr.  (The reference converts to a cudaArray texture here and frees the linear copy.)
    UtilsCuda::BindlessTexture2D<float>* getTexture()
    {
        if (!m_dev) return 0x0;
        if (!m_tex) m_tex = UtilsCuda::BindlessTexture2D<float>::view(n_alpha, n_t, m_dev);
        return m_tex;
    }

    /// Access to raw data on CPU (may be an invalid image, see readback).
    NRRD::ImageView<float>& data() { return m_raw_cpu; }
    const NRRD::ImageView<float>& data() const { return m_raw_cpu; }

    /// Engine used by compute(): ECC_INTERP_TEXTURE (bit-identical to the reference's kernel, default),
    /// ECC_INTERP_HYBRID_STATIC (same arithmetic through both sampling pipes, 2.5x faster, within 5e-5 of the peak,
    /// reproducible), ECC_INTERP_HYBRID (run-time work queue) or ECC_INTERP_EXACT (fp32 weights).
    void setInterpolation(int interp) { m_interp = interp; }

    /// Device pointer of the dtr (n_t rows of n_alpha floats); used by MetricRadonIntermediate.
    const float* devicePointer() const { return m_dev; }

protected:
    NRRD::Image<float> m_raw_cpu;                 //< Optional data on CPU.
    float* m_dev;                                 //< Data on GPU (linear, 512-byte aligned).
    UtilsCuda::BindlessTexture2D<float>* m_tex;   //< Resident handle handed out by getTexture().
    double m_bin_size_angle;
    double m_bin_size_distance;
    Filter m_filter;
    int n_x, n_y, n_t, n_alpha;
    int m_interp;

    void alloc()
    {
        if (m_dev) ecc_device_free(detail::shared_context(), m_dev);
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

    /// Runs the Radon kernel (RadonIntermediate.cpp:198-211).
    void compute(const float* image, int w, int h, int size_alpha, int size_t, Filter filter, PostProcess post_process)
    {
        n_x = w;
        n_y = h;
        n_alpha = size_alpha;
        n_t = size_t;
        m_filter = filter;
        ecc_radon_bin_sizes(n_x, n_y, n_alpha, n_t, &m_bin_size_angle, &m_bin_size_distance);
        alloc();
        ecc_context* ctx = detail::shared_context();
        detail::check(ecc_radon_compute(ctx, image, 1, n_x, n_y, n_alpha, n_t, (int)filter, (int)post_process, m_interp, m_dev), ctx,
                      "ecc_radon_compute");
        detail::check(ecc_synchronize(ctx), ctx, "ecc_synchronize");
    }

private:
    RadonIntermediate(const RadonIntermediate&);
    RadonIntermediate& operator=(const RadonIntermediate&);
};

}  // namespace EpipolarConsistency

#endif
