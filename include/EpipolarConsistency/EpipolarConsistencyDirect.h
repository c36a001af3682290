// EpipolarConsistencyDirect.h -- facade of EpipolarConsistency::MetricDirect and computeForImagePair with the reference's
// public interface (LibEpipolarConsistency/EpipolarConsistencyDirect.h:17-60, .cpp:64-270) on top of libecc_b200's C ABI.
// The reference scores ONE pair per call sequence (host geometry, two uploads, two kernel launches, two downloads, host
// sum); evaluate() here is one launch for all n(n-1)/2 pairs.
#ifndef ECC_FACADE_METRIC_DIRECT_H
#define ECC_FACADE_METRIC_DIRECT_H

#include <cstring>
#include <vector>

#include "EpipolarConsistency.h"

namespace EpipolarConsistency {

namespace detail {

inline void direct_push_matrices(ecc_context* ctx, const ProjectionMatrix* Ps, int n)
{
    std::vector<double> flat(12 * (size_t)n);
    for (int i = 0; i < n; i++) std::memcpy(&flat[12 * (size_t)i], Ps[i].data(), sizeof(double) * 12);
    check(ecc_set_projection_matrices(ctx, flat.data(), n), ctx, "ecc_set_projection_matrices");
}

inline void direct_push_images(ecc_context* ctx, const UtilsCuda::BindlessTexture2D<float>* const* Is, int n)
{
    std::vector<const float*> ptrs(n);
    for (int i = 0; i < n; i++) ptrs[i] = Is[i]->device;
    check(ecc_direct_set_image_pointers(ctx, ptrs.data(), n, Is[0]->size[0], Is[0]->size[1]), ctx, "ecc_direct_set_image_pointers");
}

// One pair with the reference's conventions for its optional vectors (EpipolarConsistencyDirect.cpp:73-77, 98-106, 200-204):
// a non-empty `kappas` holds the caller's plane angles, an empty one receives the angles used; the sample vectors are sized
// 3 n_lines with the first n_lines filled.
inline double direct_pair(ecc_context* ctx, int i, int j, std::vector<float>* redundant_samples0, std::vector<float>* redundant_samples1,
                          std::vector<float>* kappas)
{
    std::vector<float> tmp0, tmp1, tmp2;
    if (!redundant_samples0) redundant_samples0 = &tmp0;
    if (!redundant_samples1) redundant_samples1 = &tmp1;
    if (!kappas) kappas = &tmp2;
    int n_given = (int)kappas->size(), n_lines = n_given;
    if (kappas->empty()) {
        check(ecc_direct_evaluate_pair(ctx, i, j, 0, 0, 0x0, 0x0, 0x0, &n_lines, 0x0), ctx, "ecc_direct_evaluate_pair");
        kappas->resize(n_lines);
    }
    redundant_samples0->assign((size_t)3 * n_lines, 0.f);
    redundant_samples1->assign((size_t)3 * n_lines, 0.f);
    double value = 0;
    check(ecc_direct_evaluate_pair(ctx, i, j, n_given, n_lines, kappas->data(), redundant_samples0->data(), redundant_samples1->data(),
                                   &n_lines, &value),
          ctx, "ecc_direct_evaluate_pair");
    return value;
}

}  // namespace detail

/// The main algorithm behind epipolar consistency, when not using Radon intermediate functions.
/// object_radius_mm <= 0: the larger of the two views' estimates (EpipolarConsistencyDirect.cpp:80-82); dkappa <= 0: half the
/// kappa range over the image diagonal (:91-96).
inline double computeForImagePair(const ProjectionMatrix& P0, const ProjectionMatrix& P1, const UtilsCuda::BindlessTexture2D<float>& I0,
                                  const UtilsCuda::BindlessTexture2D<float>& I1, double dkappa, double object_radius_mm, bool fbcc = false,
                                  std::vector<float>* redundant_samples0 = 0x0, std::vector<float>* redundant_samples1 = 0x0,
                                  std::vector<float>* kappas = 0x0)
{
    ecc_context* ctx = detail::shared_context();
    const ProjectionMatrix Ps[2] = {P0, P1};
    const UtilsCuda::BindlessTexture2D<float>* Is[2] = {&I0, &I1};
    detail::direct_push_images(ctx, Is, 2);
    if (object_radius_mm <= 0) {  // the context estimates from its first matrix: ask for both and take the larger
        double r[2] = {0, 0};
        for (int k = 0; k < 2; k++) {
            detail::direct_push_matrices(ctx, &Ps[k], 1);
            detail::check(ecc_set_object_radius(ctx, 0.0), ctx, "ecc_set_object_radius");
            detail::check(ecc_get_object_radius(ctx, &r[k]), ctx, "ecc_get_object_radius");
        }
        object_radius_mm = r[0] > r[1] ? r[0] : r[1];
    }
    detail::direct_push_matrices(ctx, Ps, 2);
    detail::check(ecc_set_object_radius(ctx, object_radius_mm), ctx, "ecc_set_object_radius");
    detail::check(ecc_set_epipolar_plane_step(ctx, dkappa > 0 ? dkappa : 0.0), ctx, "ecc_set_epipolar_plane_step");
    detail::check(ecc_direct_set_fan_beam(ctx, fbcc ? 1 : 0), ctx, "ecc_direct_set_fan_beam");
    return detail::direct_pair(ctx, 0, 1, redundant_samples0, redundant_samples1, kappas);
}

/// Compute Epipolar Consistency on the GPU directly from projection images.
class MetricDirect : public Metric {
    /// Projection images as 2D single-channel float textures (not owned).
    std::vector<UtilsCuda::BindlessTexture2D<float>*> Is;
    /// Use standard epipolar consistency with derivative or the rectified version without derivative?
    bool use_fbcc;
    ecc_context* ctx;

    void pushSettings()
    {
        detail::check(ecc_set_object_radius(ctx, userObjectRadius()), ctx, "ecc_set_object_radius");
        detail::check(ecc_set_epipolar_plane_step(ctx, dkappa), ctx, "ecc_set_epipolar_plane_step");
        detail::check(ecc_direct_set_fan_beam(ctx, use_fbcc ? 1 : 0), ctx, "ecc_direct_set_fan_beam");
    }
    MetricDirect(const MetricDirect&);
    MetricDirect& operator=(const MetricDirect&);

public:
    /// Direct evaluation of epipolar consistency metric (for repeated evaluations see also: MetricRadonIntermediate)
    MetricDirect(const std::vector<ProjectionMatrix>& _Ps, const std::vector<UtilsCuda::BindlessTexture2D<float>*>& _Is)
        : Metric(), use_fbcc(false), ctx(0x0)
    {
        detail::check(ecc_create(-1, &ctx), 0x0, "ecc_create");
        setProjectionMatrices(_Ps);
        setProjectionImages(_Is);
    }
    ~MetricDirect() { ecc_destroy(ctx); }

    /// Set projection matrices.
    virtual Metric& setProjectionMatrices(const std::vector<ProjectionMatrix>& _Ps)
    {
        Metric::setProjectionMatrices(_Ps);
        if (!Ps.empty()) detail::direct_push_matrices(ctx, Ps.data(), (int)Ps.size());
        return *this;
    }

    /// Set projections images from single-channel 2D float textures (copied into the metric's own arrays).
    virtual Metric& setProjectionImages(const std::vector<UtilsCuda::BindlessTexture2D<float>*>& _Is)
    {
        if (!_Is.empty()) {
            n_u = _Is.front()->size[0];
            n_v = _Is.front()->size[1];
            detail::direct_push_images(ctx, _Is.data(), (int)_Is.size());
        }
        Is = _Is;
        return *this;
    }

    /// The number of projections. The number of evaluations will be n*(n-1)/2
    virtual int getNumberOfProjetions() { return (int)Is.size(); }

    /// Sampling occurs for all planes which intersect the sphere with that radius (automatic: from the first matrix).
    virtual double getObjectRadius() const
    {
        if (userObjectRadius() > 0) return userObjectRadius();
        if (Ps.empty()) return 0;
        double r = 0;
        detail::check(ecc_set_object_radius(ctx, 0.0), ctx, "ecc_set_object_radius");
        detail::check(ecc_get_object_radius(ctx, &r), ctx, "ecc_get_object_radius");
        return r;
    }

    /// Evaluates metric and optionally returns n*n cost image (entry i + j n for i < j; the SUM over the pairs is returned).
    virtual double evaluate(float* out = 0x0)
    {
        pushSettings();
        double cost = 0;
        detail::check(ecc_direct_evaluate(ctx, out, &cost), ctx, "ecc_direct_evaluate");
        return cost;
    }

    /// Evaluate for just two images i and j and optionally also return redundant values.
    virtual double evaluateForImagePair(int i, int j, std::vector<float>* redundant_samples0 = 0x0, std::vector<float>* redundant_samples1 = 0x0,
                                        std::vector<float>* kappas = 0x0)
    {
        pushSettings();
        return detail::direct_pair(ctx, i, j, redundant_samples0, redundant_samples1, kappas);
    }

    /// Change algorithm to use rectification instead of derivative.
    MetricDirect& setFanBeamConsistency(bool fbcc = true)
    {
        use_fbcc = fbcc;
        return *this;
    }

    /// NEW: clip lines against n_u x n_u as the reference's launcher does (EpipolarConsistencyDirect.cu:135) instead of against
    /// the image -- only for comparisons with the reference on non-square images.
    MetricDirect& setReferenceClip(bool on = true)
    {
        detail::check(ecc_direct_set_reference_clip(ctx, on ? 1 : 0), ctx, "ecc_direct_set_reference_clip");
        return *this;
    }
};

}  // namespace EpipolarConsistency

#endif
