// Compat.h -- the few third-party types the reference's public headers expose, so that the facade keeps the
// reference's signatures.  With Eigen on the include path the real types are used
// (Geometry::ProjectionMatrix = Eigen::Matrix<double,3,4>, LibProjectiveGeometry/ProjectiveGeometry.hxx:15-26);
// without it (this build container has none) layout-compatible stand-ins are provided: column-major storage,
// data(), operator()(r,c), operator[] -- what the metric path touches.
#ifndef ECC_FACADE_COMPAT_H
#define ECC_FACADE_COMPAT_H

#include <cstring>

#if defined(ECC_FACADE_USE_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Dense>))
#include <Eigen/Dense>
namespace Geometry {
typedef Eigen::Matrix<double, 3, 4> ProjectionMatrix;
}
#else
namespace Eigen {
struct Vector4i {
    int v[4];
    Vector4i() { v[0] = v[1] = v[2] = v[3] = 0; }
    Vector4i(int a, int b, int c, int d) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    int* data() { return v; }
    const int* data() const { return v; }
    int& operator[](int i) { return v[i]; }
    int operator[](int i) const { return v[i]; }
};
}  // namespace Eigen
namespace Geometry {
struct ProjectionMatrix {  // 3x4, column-major like Eigen's default
    double m[12];
    ProjectionMatrix() { std::memset(m, 0, sizeof(m)); }
    double* data() { return m; }
    const double* data() const { return m; }
    double& operator()(int r, int c) { return m[r + 3 * c]; }
    double operator()(int r, int c) const { return m[r + 3 * c]; }
    static ProjectionMatrix Zero() { return ProjectionMatrix(); }
    bool isZero() const
    {
        for (double x : m)
            if (x != 0.0) return false;
        return true;
    }
};
}  // namespace Geometry
#endif

namespace Geometry {
const double Pi = 3.14159265358979323846264338327950288419716939937510582;
}

#endif
