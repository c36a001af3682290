// UtilsCuda.h -- facade stand-ins for the two LibUtilsCuda types the reference's public headers mention
// (LibUtilsCuda/CudaBindlessTexture.h:18-33, CudaMemory.h:24-143).  In the reference a BindlessTexture2D owns a
// cudaArray + texture object; callers of the metric path use it only as "this image / dtr is resident on the GPU"
// (Gui/InputDataRadonIntermediate.cpp:78, EpipolarConsistencyRadonIntermediate.cpp:102).  Here it is a plain
// resident-memory handle: texture objects are created inside libecc_b200 over this memory (zero copy).
#ifndef ECC_FACADE_UTILSCUDA_H
#define ECC_FACADE_UTILSCUDA_H

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../ecc_b200.h"

namespace EpipolarConsistency {
namespace detail {

// The reference prints and exit()s on any CUDA error (LibUtilsCuda/UtilsCuda.hxx:14-28).  Same here, unless
// ECC_FACADE_THROW is defined, in which case a std::runtime_error is thrown.
inline void check(int rc, ecc_context* ctx, const char* what)
{
    if (rc == ECC_OK) return;
    std::string msg = std::string("libecc_b200: ") + what + " failed (" + std::to_string(rc) + "): " +
                      (ctx ? ecc_last_error(ctx) : "no context (no CUDA device? there is no CPU fallback)");
#ifdef ECC_FACADE_THROW
    throw std::runtime_error(msg);
#else
    std::fprintf(stderr, "%s\n", msg.c_str());
    std::exit(rc);
#endif
}

// One process-wide context for image / dtr residency and Radon computation (metrics own their own contexts).
inline ecc_context* shared_context()
{
    static ecc_context* ctx = 0x0;
    if (!ctx) check(ecc_create(-1, &ctx), 0x0, "ecc_create");
    return ctx;
}

}  // namespace detail
}  // namespace EpipolarConsistency

namespace UtilsCuda {

template <typename T> class BindlessTexture2D;

/// A 2-D float image resident in device memory (x fastest).  Owns the memory unless constructed as a view.
template <> class BindlessTexture2D<float> {
    bool owner;

public:
    int size[2];
    const float* device;  // replaces the reference's public members `array` and `tex`

    /// Upload (or adopt) a w x h image.  buffer_is_device: `buffer` is device memory and is copied device-to-device,
    /// as in the reference constructor (LibUtilsCuda/CudaBindlessTexture.cpp:17-44).
    BindlessTexture2D(int w, int h, const float* buffer, bool buffer_is_device = false, bool /*interpolate*/ = true,
                      bool /*normalizedCoords*/ = false)
        : owner(true), device(0x0)
    {
        using namespace EpipolarConsistency::detail;
        size[0] = w;
        size[1] = h;
        (void)buffer_is_device;  // ecc_copy finds out by itself
        void* p = 0x0;
        check(ecc_device_alloc(shared_context(), sizeof(float) * (size_t)w * h, &p), shared_context(), "ecc_device_alloc");
        check(ecc_copy(shared_context(), p, buffer, sizeof(float) * (size_t)w * h), shared_context(), "ecc_copy");
        device = (const float*)p;
    }
    /// Non-owning view of device memory.
    static BindlessTexture2D* view(int w, int h, const float* device_ptr)
    {
        BindlessTexture2D* t = new BindlessTexture2D();
        t->size[0] = w;
        t->size[1] = h;
        t->device = device_ptr;
        return t;
    }
    ~BindlessTexture2D()
    {
        if (owner && device) ecc_device_free(EpipolarConsistency::detail::shared_context(), (void*)device);
    }

private:
    BindlessTexture2D() : owner(false), device(0x0) { size[0] = size[1] = 0; }
    BindlessTexture2D(const BindlessTexture2D&);
    BindlessTexture2D& operator=(const BindlessTexture2D&);
};

}  // namespace UtilsCuda

#endif
