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
#include <vector>

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

/// Device memory of the facade: the reference's MemoryView / MemoryBlock (LibUtilsCuda/CudaMemory.h:24-143) as far as the
/// public headers of the path use them -- implicit cast to T*, allocate() that reallocates only when the size changes,
/// download() = host to device and readback() = device to host (named from the GPU's point of view, as upstream),
/// setZero().  Sizes are element counts; memory comes from ecc_device_alloc (512-byte aligned).
template <typename T> class MemoryBlock {
    T* ptr_d;
    size_t n;
    MemoryBlock(const MemoryBlock&);
    MemoryBlock& operator=(const MemoryBlock&);

public:
    MemoryBlock() : ptr_d(0x0), n(0) {}
    explicit MemoryBlock(size_t count, const T* host_data = 0x0) : ptr_d(0x0), n(0)
    {
        allocate(count);
        if (host_data) download(host_data);
    }
    ~MemoryBlock() { dealloc(); }
    operator T*() { return ptr_d; }
    operator const T*() const { return ptr_d; }
    size_t size() const { return n; }
    void allocate(size_t count)
    {
        using namespace EpipolarConsistency::detail;
        if (count == n) return;
        dealloc();
        if (!count) return;
        void* p = 0x0;
        check(ecc_device_alloc(shared_context(), sizeof(T) * count, &p), shared_context(), "ecc_device_alloc");
        ptr_d = (T*)p;
        n = count;
    }
    void dealloc()
    {
        if (ptr_d) ecc_device_free(EpipolarConsistency::detail::shared_context(), ptr_d);
        ptr_d = 0x0;
        n = 0;
    }
    void download(const T* host_data, size_t count = 0)
    {
        using namespace EpipolarConsistency::detail;
        if (count) allocate(count);
        check(ecc_copy(shared_context(), ptr_d, host_data, sizeof(T) * n), shared_context(), "ecc_copy");
    }
    void readback(T* host_data) const
    {
        using namespace EpipolarConsistency::detail;
        check(ecc_copy(shared_context(), host_data, ptr_d, sizeof(T) * n), shared_context(), "ecc_copy");
    }
    void setZero()
    {
        std::vector<T> zeros(n, T(0));
        if (n) download(zeros.data());
    }
};

template <typename T> class BindlessTexture2D;

/// A 2-D float image resident on the GPU (x fastest): the reference's texture wrapper (LibUtilsCuda/CudaBindlessTexture.h:18-33).
/// `device` is the image in linear device memory -- what libecc_b200 computes on (its own texture objects sit over linear
/// memory, zero copy).  `array` / `tex` -- the reference's public members, a cudaArray_t and a cudaTextureObject_t over a
/// copy of the image -- exist for code that reads them: an object built by the reference's constructor has them at once,
/// a view handed out by RadonIntermediate::getTexture() creates them on the first texture() / cast (so that 496
/// intermediates are not copied into arrays nobody samples).
template <> class BindlessTexture2D<float> {
    bool owner;
    bool interpolate;

public:
    const bool normalizedCoords;
    int size[2];
    const float* device;       // linear device memory (not in the reference)
    void* array;               // cudaArray_t
    unsigned long long tex;    // cudaTextureObject_t

    /// Upload (or adopt) a w x h image.  buffer_is_device: `buffer` is device memory and is copied device-to-device,
    /// as in the reference constructor (LibUtilsCuda/CudaBindlessTexture.cpp:17-44).
    BindlessTexture2D(int w, int h, const float* buffer, bool buffer_is_device = false, bool _interpolate = true,
                      bool _normalizedCoords = false)
        : owner(true), interpolate(_interpolate), normalizedCoords(_normalizedCoords), device(0x0), array(0x0), tex(0)
    {
        using namespace EpipolarConsistency::detail;
        size[0] = w;
        size[1] = h;
        (void)buffer_is_device;  // ecc_copy finds out by itself
        void* p = 0x0;
        check(ecc_device_alloc(shared_context(), sizeof(float) * (size_t)w * h, &p), shared_context(), "ecc_device_alloc");
        check(ecc_copy(shared_context(), p, buffer, sizeof(float) * (size_t)w * h), shared_context(), "ecc_copy");
        device = (const float*)p;
        texture();
    }
    /// Non-owning view of device memory (array and texture object are created on demand).
    static BindlessTexture2D* view(int w, int h, const float* device_ptr, bool _normalizedCoords = true)
    {
        BindlessTexture2D* t = new BindlessTexture2D(_normalizedCoords);
        t->size[0] = w;
        t->size[1] = h;
        t->device = device_ptr;
        return t;
    }
    ~BindlessTexture2D()
    {
        using namespace EpipolarConsistency::detail;
        if (tex || array) ecc_texture_destroy(shared_context(), tex, array);
        if (owner && device) ecc_device_free(shared_context(), (void*)device);
    }

    /// The texture object over a CUDA-array copy of the image (created on first use).
    unsigned long long texture()
    {
        using namespace EpipolarConsistency::detail;
        if (!tex && device)
            check(ecc_texture_create(shared_context(), device, size[0], size[1], normalizedCoords ? 1 : 0, interpolate ? 1 : 0, &tex, &array),
                  shared_context(), "ecc_texture_create");
        return tex;
    }
    inline operator unsigned long long() { return texture(); }  // cudaTextureObject_t

    /// The image back in linear device memory (LibUtilsCuda/CudaBindlessTexture.cpp:46-67).
    void readback(MemoryBlock<float>& buffer)
    {
        using namespace EpipolarConsistency::detail;
        buffer.allocate((size_t)size[0] * size[1]);
        if (array) check(ecc_texture_readback(shared_context(), array, size[0], size[1], (float*)buffer), shared_context(), "ecc_texture_readback");
        else check(ecc_copy(shared_context(), (float*)buffer, device, sizeof(float) * (size_t)size[0] * size[1]), shared_context(), "ecc_copy");
    }

private:
    explicit BindlessTexture2D(bool _normalizedCoords)
        : owner(false), interpolate(true), normalizedCoords(_normalizedCoords), device(0x0), array(0x0), tex(0)
    {
        size[0] = size[1] = 0;
    }
    BindlessTexture2D(const BindlessTexture2D&);
    BindlessTexture2D& operator=(const BindlessTexture2D&);
};

}  // namespace UtilsCuda

#endif
