// EpipolarConsistencyRadonIntermediate.h -- facade of EpipolarConsistency::MetricRadonIntermediate with the
// reference's public interface (LibEpipolarConsistency/EpipolarConsistencyRadonIntermediate.h:21-106,
// .cpp:41-322) on top of libecc_b200's C ABI.  New capability beyond the reference: evaluateBatch().
#ifndef ECC_FACADE_METRIC_RADON_INTERMEDIATE_H
#define ECC_FACADE_METRIC_RADON_INTERMEDIATE_H

#include <set>
#include <utility>
#include <vector>

#include "EpipolarConsistency.h"
#include "RadonIntermediate.h"

namespace EpipolarConsistency {

/// Compute Epipolar Consistency on the GPU
class MetricRadonIntermediate : public Metric {
    std::vector<RadonIntermediate*> dtrs;  //< Radon intermediate functions (not owned).
    bool use_corr;
    ecc_context* ctx;

    void chk(int rc, const char* what) const { detail::check(rc, ctx, what); }
    void pushSettings()
    {
        chk(ecc_set_object_radius(ctx, userObjectRadius()), "ecc_set_object_radius");
        chk(ecc_set_epipolar_plane_step(ctx, dkappa), "ecc_set_epipolar_plane_step");
    }

public:
    MetricRadonIntermediate() : Metric(), use_corr(false), ctx(0x0) { detail::check(ecc_create(-1, &ctx), 0x0, "ecc_create"); }

    MetricRadonIntermediate(const std::vector<ProjectionMatrix>& Ps, const std::vector<RadonIntermediate*>& _dtrs)
        : Metric(), use_corr(false), ctx(0x0)
    {
        detail::check(ecc_create(-1, &ctx), 0x0, "ecc_create");
        setProjectionMatrices(Ps);
        setRadonIntermediates(_dtrs);
    }

    ~MetricRadonIntermediate() { ecc_destroy(ctx); }

    /// Deprecated.
    MetricRadonIntermediate& setdKappa(float _dkappa)
    {
        dkappa = _dkappa;
        return *this;
    }

    /// Tell Metric to compute correlation instead of SSD (EpipolarConsistencyRadonIntermediate.h:43): pairs are scored by
    /// 1 - cc, as the reference's use_corr path forms it.
    MetricRadonIntermediate& useCorrelation(bool corr = true)
    {
        detail::check(ecc_use_correlation(ctx, corr ? 1 : 0), ctx, "ecc_use_correlation");
        use_corr = corr;
        return *this;
    }

    /// Interpolation flavour: ECC_INTERP_TEXTURE (reference CUDA numerics, default) or ECC_INTERP_EXACT.
    MetricRadonIntermediate& setInterpolation(int interp)
    {
        chk(ecc_set_interpolation(ctx, interp), "ecc_set_interpolation");
        return *this;
    }

    /// Register the Radon intermediates (zero copy).  DO NOT delete or change _dtrs during lifetime of Metric.
    MetricRadonIntermediate& setRadonIntermediates(const std::vector<RadonIntermediate*>& _dtrs)
    {
        dtrs = _dtrs;
        if (dtrs.empty()) return *this;
        // size and parameters of dtrs (assumed to be identical), EpipolarConsistencyRadonIntermediate.cpp:89-98
        RadonIntermediate* d0 = dtrs[0];
        n_u = d0->getOriginalImageSize(0);
        n_v = d0->getOriginalImageSize(1);
        std::vector<const float*> ptrs(dtrs.size());
        for (size_t i = 0; i < dtrs.size(); i++) {
            dtrs[i]->getTexture();  // "make resident"
            ptrs[i] = dtrs[i]->devicePointer();
        }
        const int n_alpha = d0->getRadonBinNumber(0), n_t = d0->getRadonBinNumber(1);
        if (n_alpha % 8 == 0) {
            chk(ecc_set_radon_intermediate_pointers(ctx, ptrs.data(), (int)ptrs.size(), n_alpha, n_t, d0->getRadonBinSize(0),
                                                    d0->getRadonBinSize(1), n_u, n_v, d0->isDerivative()),
                "ecc_set_radon_intermediate_pointers");
        } else {
            // widths that cannot be a texture pitch: gather into one block, the library repacks it
            const size_t len = (size_t)n_alpha * n_t;
            void* block = 0x0;
            chk(ecc_device_alloc(ctx, sizeof(float) * len * ptrs.size(), &block), "ecc_device_alloc");
            for (size_t i = 0; i < ptrs.size(); i++)
                chk(ecc_copy(ctx, (float*)block + len * i, ptrs[i], sizeof(float) * len), "ecc_copy");
            chk(ecc_set_radon_intermediates(ctx, (const float*)block, (int)ptrs.size(), n_alpha, n_t, d0->getRadonBinSize(0),
                                            d0->getRadonBinSize(1), n_u, n_v, d0->isDerivative()),
                "ecc_set_radon_intermediates");
            chk(ecc_device_free(ctx, block), "ecc_device_free");
        }
        return *this;
    }

    /// Access Radon intermediate functions (for visualization and debugging)
    const std::vector<RadonIntermediate*>& getRadonIntermediates() const { return dtrs; }

    /// Declared but never defined in the reference (EpipolarConsistencyRadonIntermediate.h:52); a no-op here.
    Metric& setRadonIntermediateBinning(int, int) { return *this; }

    /// A TODO no-op in the reference as well (EpipolarConsistencyRadonIntermediate.cpp:121-125).
    virtual Metric& setProjectionImages(const std::vector<UtilsCuda::BindlessTexture2D<float>*>&) { return *this; }

    /// Compute null space and pseudoinverse of projection matrices and convert to float.
    virtual Metric& setProjectionMatrices(const std::vector<ProjectionMatrix>& _Ps)
    {
        Metric::setProjectionMatrices(_Ps);
        if (Ps.empty()) return *this;
        std::vector<double> flat(12 * Ps.size());
        for (size_t i = 0; i < Ps.size(); i++) std::memcpy(&flat[12 * i], Ps[i].data(), sizeof(double) * 12);
        chk(ecc_set_projection_matrices(ctx, flat.data(), (int)Ps.size()), "ecc_set_projection_matrices");
        return *this;
    }

    /// NEW: replace one matrix of the current set; only that view is re-derived on the device (tracking loops change a
    /// single view per step, Gui/SingleImageMotion.h:84-90).
    MetricRadonIntermediate& updateProjectionMatrix(int index, const ProjectionMatrix& P)
    {
        Ps[index] = P;
        chk(ecc_update_projection_matrix(ctx, index, P.data()), "ecc_update_projection_matrix");
        return *this;
    }

    /// NEW: updateProjectionMatrix(index, P) followed by evaluate(indices, out) as ONE call (same results); repeated calls
    /// with the same index and list replay a recorded CUDA graph (ecc_update_and_evaluate): the step of a tracking loop.
    double updateAndEvaluate(int index, const ProjectionMatrix& P, const std::vector<Eigen::Vector4i>& indices, float* out = 0x0)
    {
        Ps[index] = P;
        pushSettings();
        std::vector<int> flat(indices.size() * 4);
        for (size_t i = 0; i < indices.size(); i++)
            for (int k = 0; k < 4; k++) flat[4 * i + k] = indices[i][k];
        double mean = 0;
        chk(ecc_update_and_evaluate(ctx, index, P.data(), flat.data(), (int)indices.size(), out, &mean), "ecc_update_and_evaluate");
        return mean;
    }

    /// Radius of the object: the user's value, or the automatic estimate from the first matrix.
    virtual double getObjectRadius() const
    {
        ecc_set_object_radius(ctx, userObjectRadius());
        double r = 0;
        chk(ecc_get_object_radius(ctx, &r), "ecc_get_object_radius");
        return r;
    }

    /// Evaluates metric without any transformation of the geometry. Out is n*n and mean is returned.
    virtual double evaluate(float* out = 0x0)
    {
        pushSettings();
        double mean = 0;
        chk(ecc_evaluate(ctx, out, &mean), "ecc_evaluate");
        return mean;
    }

    /// The number of projections. The number of evaluations will be n*(n-1)/2
    virtual int getNumberOfProjetions() { return (int)Ps.size(); }

    /// Evaluates metric for just specific views
    double evaluate(const std::set<int>& views, float* _out = 0x0)
    {
        std::vector<Eigen::Vector4i> indices;
        for (auto i = views.begin(); i != views.end(); ++i)
            for (auto j = i; j != views.end(); ++j) {
                if (i == j) continue;
                indices.push_back(Eigen::Vector4i(*i, *j, *i, *j));
            }
        std::vector<float> tmp;
        if (!_out) {
            tmp.resize(indices.size());
            _out = tmp.data();
        }
        return evaluate(indices, _out);
    }

    /// Evaluates metric for explicit (P0,P1,dtr0,dtr1) tuples; writes indices.size() floats to out.
    double evaluate(const std::vector<Eigen::Vector4i>& _indices, float* _out)
    {
        if (_indices.empty()) return 0.0;
        pushSettings();
        std::vector<int> flat(4 * _indices.size());
        for (size_t i = 0; i < _indices.size(); i++)
            for (int k = 0; k < 4; k++) flat[4 * i + k] = _indices[i].data()[k];
        double mean = 0;
        chk(ecc_evaluate_indices(ctx, flat.data(), (int)_indices.size(), _out, &mean), "ecc_evaluate_indices");
        return mean;
    }

    /// NEW: score K complete projection-matrix sets against the same dtrs in one launch; returns the K means.
    std::vector<double> evaluateBatch(const std::vector<std::vector<ProjectionMatrix> >& sets,
                                      const std::vector<Eigen::Vector4i>* _indices = 0x0, float* out = 0x0)
    {
        std::vector<double> means(sets.size(), 0.0);
        if (sets.empty()) return means;
        pushSettings();
        const size_t n = Ps.size();
        std::vector<double> flat(12 * n * sets.size());
        for (size_t s = 0; s < sets.size(); s++)
            for (size_t i = 0; i < n; i++) std::memcpy(&flat[12 * (s * n + i)], sets[s][i].data(), sizeof(double) * 12);
        std::vector<int> idx;
        if (_indices) {
            idx.resize(4 * _indices->size());
            for (size_t i = 0; i < _indices->size(); i++)
                for (int k = 0; k < 4; k++) idx[4 * i + k] = (*_indices)[i].data()[k];
        }
        chk(ecc_evaluate_batch(ctx, flat.data(), (int)sets.size(), _indices ? idx.data() : 0x0, _indices ? (int)_indices->size() : 0, out,
                               means.data()),
            "ecc_evaluate_batch");
        return means;
    }

    /// NEW: the same, fed with PARAMETER VECTORS: K * m instances of ModelCameraSimilarity2D3D (11 doubles each, `params`)
    /// applied to the base matrices on the device (ecc_evaluate_batch_params).  view_to_param (n entries or empty): which of
    /// a set's m instances moves view v, negative = none; empty = one instance per view (m == n, Geometry::ModelFDCT).
    /// Results equal evaluateBatch() of the matrices the host models hand out, without building or uploading them.
    std::vector<double> evaluateBatchParams(const std::vector<ProjectionMatrix>& base, const std::vector<double>& params, int m,
                                            const std::vector<int>& view_to_param = std::vector<int>(),
                                            const std::vector<Eigen::Vector4i>* _indices = 0x0, float* out = 0x0)
    {
        const int K = m > 0 ? (int)(params.size() / ((size_t)11 * m)) : 0;
        std::vector<double> means(K, 0.0);
        if (K == 0) return means;
        pushSettings();
        std::vector<double> flat(12 * base.size());
        for (size_t i = 0; i < base.size(); i++) std::memcpy(&flat[12 * i], base[i].data(), sizeof(double) * 12);
        std::vector<int> idx;
        if (_indices) {
            idx.resize(4 * _indices->size());
            for (size_t i = 0; i < _indices->size(); i++)
                for (int k = 0; k < 4; k++) idx[4 * i + k] = (*_indices)[i].data()[k];
        }
        chk(ecc_evaluate_batch_params(ctx, base.empty() ? 0x0 : flat.data(), params.data(), K, m, view_to_param.empty() ? 0x0 : view_to_param.data(),
                                      _indices ? idx.data() : 0x0, _indices ? (int)_indices->size() : 0, out, means.data()),
            "ecc_evaluate_batch_params");
        return means;
    }

    /// NEW: the same with explicit homographies -- transforms holds K * m instances of 25 doubles (H 3x3 then T 4x4,
    /// column-major), P' = H P T on the device, normalised when asked (ecc_evaluate_batch_transforms).  m == 1: one correction
    /// for the whole trajectory (Geometry::ModelFDCTCalibrationCorrection); m == n: one per view.
    std::vector<double> evaluateBatchTransforms(const std::vector<ProjectionMatrix>& base, const std::vector<double>& transforms, int m,
                                                bool normalize = false, float* out = 0x0)
    {
        const int K = m > 0 ? (int)(transforms.size() / ((size_t)25 * m)) : 0;
        std::vector<double> means(K, 0.0);
        if (K == 0) return means;
        pushSettings();
        std::vector<double> flat(12 * base.size());
        for (size_t i = 0; i < base.size(); i++) std::memcpy(&flat[12 * i], base[i].data(), sizeof(double) * 12);
        chk(ecc_evaluate_batch_transforms(ctx, base.empty() ? 0x0 : flat.data(), transforms.data(), K, m, 0x0, normalize ? 1 : 0, 0x0, 0, out,
                                          means.data()),
            "ecc_evaluate_batch_transforms");
        return means;
    }

    /// Visualisation helper of the reference (EpipolarConsistencyRadonIntermediate.cpp:324-393): that code samples on
    /// the CPU with its own texel mapping.  Here: the pair's metric value from the GPU path; sample vectors are not filled.
    /// The two redundant signals of pair (i, j) for plotting, in ascending kappa (EpipolarConsistencyRadonIntermediate.cpp
    /// :324-393).  Sampled on the device with the metric's own lookup; returns the pair's metric value (the reference
    /// returns the last term of its sum only, SURVEY.md Appendix B -- not reproduced).
    virtual double evaluateForImagePair(int i, int j, std::vector<float>* redundant_samples0 = 0x0,
                                        std::vector<float>* redundant_samples1 = 0x0, std::vector<float>* kappas = 0x0)
    {
        return evaluateForImagePair(i, j, redundant_samples0, redundant_samples1, kappas, 0x0, 0x0);
    }

    /// ... and the sample locations in the two Radon transforms: (l0, l1) of every epipolar line.
    double evaluateForImagePair(int i, int j, std::vector<float>* redundant_samples0, std::vector<float>* redundant_samples1,
                                std::vector<float>* kappas, std::vector<std::pair<float, float> >* radon_samples0,
                                std::vector<std::pair<float, float> >* radon_samples1)
    {
        pushSettings();
        int n = 0;
        double value = 0;
        chk(ecc_pair_signals(ctx, i, j, i, j, 0, 0x0, 0x0, 0x0, 0x0, 0x0, &n, 0x0, &value), "ecc_pair_signals");
        const bool any = redundant_samples0 || redundant_samples1 || kappas || radon_samples0 || radon_samples1;
        if (!any || n == 0) return value;
        std::vector<float> k(n), s0(n), s1(n), l0(2 * (size_t)n), l1(2 * (size_t)n);
        chk(ecc_pair_signals(ctx, i, j, i, j, n, k.data(), s0.data(), s1.data(), l0.data(), l1.data(), &n, 0x0, 0x0), "ecc_pair_signals");
        if (redundant_samples0) redundant_samples0->insert(redundant_samples0->end(), s0.begin(), s0.end());
        if (redundant_samples1) redundant_samples1->insert(redundant_samples1->end(), s1.begin(), s1.end());
        if (kappas) kappas->insert(kappas->end(), k.begin(), k.end());
        for (int q = 0; q < n; q++) {
            if (radon_samples0) radon_samples0->push_back(std::make_pair(l0[2 * q], l0[2 * q + 1]));
            if (radon_samples1) radon_samples1->push_back(std::make_pair(l1[2 * q], l1[2 * q + 1]));
        }
        return value;
    }

protected:
    MetricRadonIntermediate(const MetricRadonIntermediate&);
};

}  // namespace EpipolarConsistency

#endif
