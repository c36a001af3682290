// EpipolarConsistency.h -- facade of the Metric interface (LibEpipolarConsistency/EpipolarConsistency.h:49-94,
// EpipolarConsistency.cpp:70-104).
#ifndef ECC_FACADE_EPIPOLAR_CONSISTENCY_H
#define ECC_FACADE_EPIPOLAR_CONSISTENCY_H

#include <vector>

#include "Compat.h"
#include "UtilsCuda.h"

namespace EpipolarConsistency {

using Geometry::Pi;
using Geometry::ProjectionMatrix;

/// Interface for Epipolar Consistency metric given projection matrices and projection images.
class Metric {
    double object_radius_mm;  //< Radius of object. Zero for automatic.

protected:
    double dkappa;                      //< Angle between epipolar planes during sampling. Zero for automatic.
    std::vector<ProjectionMatrix> Ps;   //< Current projection matrices
    int n_u;                            //< Image size
    int n_v;                            //< Image size

public:
    /// Set radius of object. Zero for automatic.
    Metric& setObjectRadius(double radius_mm = 0)
    {
        object_radius_mm = radius_mm;
        return *this;
    }
    /// Sampling occurs for all planes which intersect the sphere with that radius.
    virtual double getObjectRadius() const { return object_radius_mm; }
    /// Angle between epipolar planes during sampling. Zero for automatic determination per view pair.
    Metric& setEpipolarPlaneStep(double dkappa_rad = 0)
    {
        dkappa = dkappa_rad;
        return *this;
    }
    /// Set projection matrices.
    virtual Metric& setProjectionMatrices(const std::vector<ProjectionMatrix>& _Ps)
    {
        Ps = _Ps;
        return *this;
    }
    /// Get projection matrices.
    const std::vector<ProjectionMatrix>& getProjectionMatrices() const { return Ps; }
    /// Set projections images from single-channel 2D float textures.
    virtual Metric& setProjectionImages(const std::vector<UtilsCuda::BindlessTexture2D<float>*>& Is) = 0;
    /// The number of projections. The number of evaluations will be n*(n-1)/2
    virtual int getNumberOfProjetions() = 0;
    /// Evaluates metric and optionally returns n*n cost image.
    virtual double evaluate(float* cost_image = 0x0) = 0;
    /// Evaluate for just two images i and j and optionally also return redundant values.
    virtual double evaluateForImagePair(int i, int j, std::vector<float>* redundant_samples0 = 0x0,
                                        std::vector<float>* redundant_samples1 = 0x0, std::vector<float>* kappas = 0x0) = 0;

    Metric() : object_radius_mm(0), dkappa(0), n_u(0), n_v(0) {}
    virtual ~Metric() {}

protected:
    double userObjectRadius() const { return object_radius_mm; }
};

}  // namespace EpipolarConsistency

#endif
