// Adaptors.h -- the callers of the hot path, re-expressed over the facade / C ABI:
//   Geometry::ModelSimilarity2D / ModelSimilarity3D / ModelCameraSimilarity2D3D
//       LibProjectiveGeometry/Models/ModelSimilarity2D.hxx:52-72, ModelSimilarity3D.hxx:64-87,
//       ModelCameraSimilarity2D3D.hxx:89-92 (P' = H2D * P * T3D, 4 + 7 parameters)
//   EpipolarConsistency::SingleImageMotion      LibEpipolarConsistency/Gui/SingleImageMotion.h:13-92
//   EpipolarConsistency::Registration           LibEpipolarConsistency/Gui/Registration.h:13-93
//   EpipolarConsistency::Registration3D3D       tools/Registration/Registration3D3D.hxx:13-115
// Same class names, constructor arguments (minus the LibOpterix parameter-model reference: LibOpterix/NLopt are the
// optimiser shell, out of scope) and evaluate() semantics.  What the B200 path adds is what the reference's loops
// lack: a changed view re-derives ONE matrix instead of all, only the pairs that involve the
// moving view are scored when asked, and K candidate poses are scored in ONE launch (evaluateCandidates) -- the unit
// of work of finite-difference stencils, simplex / population steps and grid scans.
#ifndef ECC_FACADE_ADAPTORS_H
#define ECC_FACADE_ADAPTORS_H

#include <cmath>
#include <vector>

#include "EpipolarConsistencyRadonIntermediate.h"

namespace Geometry {

/// 3x3 and 4x4 homographies, column-major like Eigen's default (Geometry::RP2Homography / RP3Homography).
struct Homography2D {
    double m[9];
    Homography2D() { for (int i = 0; i < 9; i++) m[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    double& operator()(int r, int c) { return m[r + 3 * c]; }
    double operator()(int r, int c) const { return m[r + 3 * c]; }
    const double* data() const { return m; }
};
struct Homography3D {
    double m[16];
    Homography3D() { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    double& operator()(int r, int c) { return m[r + 4 * c]; }
    double operator()(int r, int c) const { return m[r + 4 * c]; }
    const double* data() const { return m; }
};

/// H (3x3) * P (3x4) * T (4x4), all column-major; any type with data() works (Eigen matrices included).
template <class PM>
inline PM transformProjection(const Homography2D& H, const PM& P, const Homography3D& T)
{
    double HP[12];
    const double* p = P.data();
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += H(r, k) * p[k + 3 * c];
            HP[r + 3 * c] = s;
        }
    PM out = P;
    double* o = out.data();
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 3; r++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += HP[r + 3 * k] * T(k, c);
            o[r + 3 * c] = s;
        }
    return out;
}

/// Translation u, v; 2D rotation; 2D scale (ModelSimilarity2D.hxx:14-26,52-72).
struct ModelSimilarity2D {
    std::vector<double> current_values;
    ModelSimilarity2D() : current_values(4, 0.0) {}
    Homography2D getInstance() const
    {
        const std::vector<double>& x = current_values;
        Homography2D H;
        if (x[2] != 0) {
            H(0, 0) = +std::cos(x[2]); H(0, 1) = -std::sin(x[2]);
            H(1, 0) = +std::sin(x[2]); H(1, 1) = +std::cos(x[2]);
        }
        H(0, 2) = x[0];
        H(1, 2) = x[1];
        if (x[3] != 0)
            for (int r = 0; r < 2; r++)
                for (int c = 0; c < 2; c++) H(r, c) *= (1.0 + x[3]);
        return H;
    }
};

/// Translation X, Y, Z; rotation about X, Y, Z (R = Rx * Ry * Rz); 3D scale (ModelSimilarity3D.hxx:16-31,64-87).
struct ModelSimilarity3D {
    std::vector<double> current_values;
    ModelSimilarity3D() : current_values(7, 0.0) {}
    Homography3D getInstance() const
    {
        const std::vector<double>& x = current_values;
        Homography3D T;
        if (x[3] != 0 || x[4] != 0 || x[5] != 0) {
            const double cx = std::cos(x[3]), sx = std::sin(x[3]), cy = std::cos(x[4]), sy = std::sin(x[4]);
            const double cz = std::cos(x[5]), sz = std::sin(x[5]);
            const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};  // row-major
            const double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
            const double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
            double A[9], R[9];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    A[3 * r + c] = 0;
                    for (int k = 0; k < 3; k++) A[3 * r + c] += Rx[3 * r + k] * Ry[3 * k + c];
                }
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) {
                    R[3 * r + c] = 0;
                    for (int k = 0; k < 3; k++) R[3 * r + c] += A[3 * r + k] * Rz[3 * k + c];
                }
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) T(r, c) = R[3 * r + c];
        }
        T(0, 3) = x[0];
        T(1, 3) = x[1];
        T(2, 3) = x[2];
        if (x[6] != 0)
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) T(r, c) *= (1.0 + x[6]);
        return T;
    }
};

/// P' = H2D * P * T3D; parameter vector = the four 2D parameters followed by the seven 3D ones.
class ModelCameraSimilarity2D3D {
    ProjectionMatrix P;

public:
    std::vector<double> current_values;
    explicit ModelCameraSimilarity2D3D(const ProjectionMatrix& _P) : P(_P), current_values(11, 0.0) {}
    void setOriginalProjectionMatrix(const ProjectionMatrix& _P)
    {
        P = _P;
        current_values.assign(11, 0.0);
    }
    Homography2D getTransform2D() const
    {
        ModelSimilarity2D m;
        m.current_values.assign(current_values.begin(), current_values.begin() + 4);
        return m.getInstance();
    }
    Homography3D getTransform3D() const
    {
        ModelSimilarity3D m;
        m.current_values.assign(current_values.begin() + 4, current_values.end());
        return m.getInstance();
    }
    ProjectionMatrix getInstance() const { return transformProjection(getTransform2D(), P, getTransform3D()); }
    /// The matrix for a given parameter vector (does not change current_values).
    ProjectionMatrix getInstance(const std::vector<double>& x) const
    {
        ModelCameraSimilarity2D3D tmp(P);
        tmp.current_values = x;
        return tmp.getInstance();
    }
};

}  // namespace Geometry

namespace EpipolarConsistency {

/// Pre-processing of X-ray projection images (Gui/PreProccess.h:14-52, PreProccess.cpp:57-166) on the device; the
/// GetSet GUI plumbing of the reference (gui_declare_section / gui_retreive_section) stays with the GUI.
struct PreProccess {
    struct Intensity {
        bool normalize;
        double bias, scale;
        bool apply_log;
        Intensity() : normalize(false), bias(0.0), scale(1.0), apply_log(false) {}
    } intensity;
    struct Lowpass {
        double gaussian_sigma;
        int half_kernel_width;
        Lowpass() : gaussian_sigma(1.84), half_kernel_width(5) {}
    } lowpass;
    struct ImageGeometry {
        bool flip_u, flip_v;
        ImageGeometry() : flip_u(false), flip_v(false) {}
    } image_geometry;
    struct Border {
        Eigen::Vector4i zero, feather;  // left, right, bottom, top
        std::vector<Eigen::Vector4i> blanks;
        Border() : zero(1, 1, 1, 1), feather(16, 16, 16, 16) {}
    } border;

    void process(NRRD::ImageView<float>& image) const { run(image, 0x0); }
    void apply_weight_cos_principal_ray(NRRD::ImageView<float>& image, const ProjectionMatrix& P) const
    {
        // only the cosine weighting: identity intensity, no borders, no low-pass
        ecc_preprocess_params p;
        ecc_preprocess_defaults(&p);
        for (int k = 0; k < 4; k++) p.border_zero[k] = p.border_feather[k] = 0;
        p.gaussian_sigma = 0;
        p.cos_weight = 1;
        ecc_context* ctx = detail::shared_context();
        detail::check(ecc_preprocess(ctx, (float*)image, 1, image.size(0), image.size(1), &p, P.data()), ctx, "ecc_preprocess");
    }
    /// process() and the cosine weighting in one pass over the image (what the loaders do back to back,
    /// Gui/InputDataDirect.cpp:75-87).
    void processAndWeight(NRRD::ImageView<float>& image, const ProjectionMatrix& P) const { run(image, &P); }

private:
    void run(NRRD::ImageView<float>& image, const ProjectionMatrix* P) const
    {
        ecc_preprocess_params p;
        ecc_preprocess_defaults(&p);
        p.scale = intensity.scale;
        p.bias = intensity.bias;
        p.normalize = intensity.normalize;
        p.apply_log = intensity.apply_log;
        for (int k = 0; k < 4; k++) { p.border_zero[k] = border.zero[k]; p.border_feather[k] = border.feather[k]; }
        std::vector<int> flat(4 * border.blanks.size());
        for (size_t i = 0; i < border.blanks.size(); i++)
            for (int k = 0; k < 4; k++) flat[4 * i + k] = border.blanks[i][k];
        p.n_blanks = (int)border.blanks.size();
        p.blanks = flat.empty() ? 0x0 : flat.data();
        p.flip_u = image_geometry.flip_u;
        p.flip_v = image_geometry.flip_v;
        p.gaussian_sigma = lowpass.gaussian_sigma;
        p.half_kernel_width = lowpass.half_kernel_width;
        p.cos_weight = P ? 1 : 0;
        ecc_context* ctx = detail::shared_context();
        detail::check(ecc_preprocess(ctx, (float*)image, 1, image.size(0), image.size(1), &p, P ? P->data() : 0x0), ctx, "ecc_preprocess");
    }
};

/// Epipolar consistency metric for changes on one projection matrix (Gui/SingleImageMotion.h).
class SingleImageMotion {
protected:
    std::vector<ProjectionMatrix> Ps;
    std::vector<RadonIntermediate*> dtrs;
    int input_index;
    MetricRadonIntermediate* ecc;
    std::vector<Eigen::Vector4i> indices;
    std::vector<float> tmp_results;
    SingleImageMotion(const SingleImageMotion&);

public:
    SingleImageMotion(std::vector<ProjectionMatrix> _Ps, std::vector<RadonIntermediate*> _dtrs, int _input_index = 0)
        : Ps(_Ps), dtrs(_dtrs), input_index(_input_index), ecc(new MetricRadonIntermediate(Ps, dtrs))
    {
        const int n = (int)dtrs.size();
        for (int i = 0; i < n; i++)
            if (i != input_index) indices.push_back(Eigen::Vector4i(input_index, i, input_index, i));
    }
    virtual ~SingleImageMotion() { delete ecc; }

    MetricRadonIntermediate& getMetricPtr() { return *ecc; }
    const std::vector<float>& getTemporaryResults() { return tmp_results; }
    const std::vector<Eigen::Vector4i>& getIndices() { return indices; }

    /// As the reference (SingleImageMotion.h:62-72): the mean over ALL pairs ("same as the n-1 pairs of the input image
    /// except some constant part is added").
    double evaluate(float* = 0x0) { return ecc->evaluate(); }

    /// Cost for P_input as the matrix of the input image.  One matrix is re-derived on the device
    /// (ecc_update_projection_matrix); the reference re-derives and re-uploads all n (SingleImageMotion.h:84-90).
    virtual double evaluate(const ProjectionMatrix& P_input)
    {
        Ps[input_index] = P_input;
        ecc->updateProjectionMatrix(input_index, P_input);
        return evaluate();
    }

    /// The n-1 pairs that involve the input image only (the code path the reference leaves unreachable after its
    /// early return); out (nullable): n-1 floats in getIndices() order.
    double evaluateMovingPairs(const ProjectionMatrix& P_input, float* out = 0x0)
    {
        Ps[input_index] = P_input;
        return ecc->updateAndEvaluate(input_index, P_input, indices, out);  // one recorded CUDA graph per step
    }

    /// K candidate matrices for the input image scored in one launch (n-1 pairs each); returns the K means.
    std::vector<double> evaluateCandidates(const std::vector<ProjectionMatrix>& candidates)
    {
        std::vector<std::vector<ProjectionMatrix> > sets(candidates.size(), Ps);
        for (size_t k = 0; k < candidates.size(); k++) sets[k][input_index] = candidates[k];
        return ecc->evaluateBatch(sets, &indices);
    }
};

/// Registration of one input image (index 0) against the others (Gui/Registration.h): always the n-1 listed pairs.
class Registration : public SingleImageMotion {
public:
    Registration(std::vector<ProjectionMatrix> _Ps, std::vector<RadonIntermediate*> _dtrs) : SingleImageMotion(_Ps, _dtrs, 0) {}
    double evaluate(float* out = 0x0)
    {
        if (!out) {
            tmp_results.resize(dtrs.size());
            out = tmp_results.data();
        }
        return ecc->evaluate(indices, out);  // mean of the n-1 values (Registration.h:64-75)
    }
    virtual double evaluate(const ProjectionMatrix& P_input)
    {
        Ps[0] = P_input;
        ecc->updateProjectionMatrix(0, P_input);
        return evaluate();
    }
};

/// Registration of two scans in the projection domain (tools/Registration/Registration3D3D.hxx): all n_source x
/// n_target cross pairs; the source matrices are right-multiplied by the candidate transform T.
class Registration3D3D {
    std::vector<ProjectionMatrix> Ps;  // source matrices first
    std::vector<RadonIntermediate*> dtrs;
    int n_source, n_target;
    std::vector<Eigen::Vector4i> indices;  // source index fast
    MetricRadonIntermediate ecc;
    std::vector<float> tmp_results;
    Registration3D3D(const Registration3D3D&);

    std::vector<ProjectionMatrix> transformed(const Geometry::Homography3D& T) const
    {
        std::vector<ProjectionMatrix> out = Ps;
        const Geometry::Homography2D I;
        for (int i = 0; i < n_source; i++) out[i] = Geometry::transformProjection(I, Ps[i], T);
        return out;
    }

public:
    Registration3D3D(bool use_cc, const std::vector<ProjectionMatrix>& Ps_source, const std::vector<RadonIntermediate*>& dtrs_source,
                     const std::vector<ProjectionMatrix>& Ps_target, const std::vector<RadonIntermediate*>& dtrs_target)
        : n_source((int)dtrs_source.size()), n_target((int)dtrs_target.size())
    {
        Ps = Ps_source;
        Ps.insert(Ps.end(), Ps_target.begin(), Ps_target.end());
        dtrs = dtrs_source;
        dtrs.insert(dtrs.end(), dtrs_target.begin(), dtrs_target.end());
        // all pairs of one source and one target projection, source index fast (the reference's own index expression,
        // Registration3D3D.hxx:66, is only a permutation of this when n_source == n_target)
        indices.resize((size_t)n_source * n_target);
        for (int j = 0; j < n_target; j++)
            for (int i = 0; i < n_source; i++) indices[(size_t)j * n_source + i] = Eigen::Vector4i(i, n_source + j, i, n_source + j);
        ecc.setProjectionMatrices(Ps);
        ecc.setRadonIntermediates(dtrs);
        if (use_cc) ecc.useCorrelation();
    }

    MetricRadonIntermediate& getMetricPtr() { return ecc; }
    const std::vector<float>& getTemporaryResults() const { return tmp_results; }
    const std::vector<Eigen::Vector4i>& getIndices() const { return indices; }

    double evaluate(float* out = 0x0)
    {
        if (!out) {
            tmp_results.resize(indices.size());
            out = tmp_results.data();
        }
        return ecc.evaluate(indices, out);
    }
    double evaluate(const Geometry::Homography3D& T_input)
    {
        ecc.setProjectionMatrices(transformed(T_input));
        return evaluate();
    }
    /// K candidate transforms in one launch; returns the K means over the cross pairs.
    std::vector<double> evaluateCandidates(const std::vector<Geometry::Homography3D>& Ts)
    {
        ecc.setProjectionMatrices(Ps);
        std::vector<std::vector<ProjectionMatrix> > sets;
        for (size_t k = 0; k < Ts.size(); k++) sets.push_back(transformed(Ts[k]));
        return ecc.evaluateBatch(sets, &indices);
    }
};

}  // namespace EpipolarConsistency

#endif
