// Adaptors.h -- the callers of the hot path, re-expressed over the facade / C ABI:
//   Geometry::ModelSimilarity2D / ModelSimilarity3D / ModelCameraSimilarity2D3D
//       LibProjectiveGeometry/Models/ModelSimilarity2D.hxx:52-72, ModelSimilarity3D.hxx:64-87,
//       ModelCameraSimilarity2D3D.hxx:89-92 (P' = H2D * P * T3D, 4 + 7 parameters)
//   EpipolarConsistency::SingleImageMotion      LibEpipolarConsistency/Gui/SingleImageMotion.h:13-92
//   EpipolarConsistency::Registration           LibEpipolarConsistency/Gui/Registration.h:13-93
//   EpipolarConsistency::Registration3D3D       tools/Registration/Registration3D3D.hxx:13-115
//   Geometry::ModelFDCT                          tools/FDCTMotionCorrection/ModelFDCT.hxx:26-62 (one 2D/3D similarity per view)
//   EpipolarConsistency::FDCTMoCo                LibEpipolarConsistency/Gui/FDCTMotionCorrection.hxx:13-98
// The models are evaluated by the library (ecc_model_*: one fp64 operation sequence for host and device), so that a matrix
// built here is bit for bit the matrix the device expands from the same parameter vector (evaluateBatchParams).
// Same class names, constructor arguments (minus the LibOpterix parameter-model reference: LibOpterix/NLopt are the
// optimiser shell, out of scope) and evaluate() semantics.  What the B200 path adds is what the reference's loops
// lack: a changed view re-derives ONE matrix instead of all, only the pairs that involve the
// moving view are scored when asked, and K candidate poses are scored in ONE launch (evaluateCandidates) -- the unit
// of work of finite-difference stencils, simplex / population steps and grid scans.
#ifndef ECC_FACADE_ADAPTORS_H
#define ECC_FACADE_ADAPTORS_H

#include <cmath>
#include <cstring>
#include <set>
#include <stdexcept>
#include <utility>
#include <string>
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
    PM out = P;
    ecc_model_transform(H.data(), P.data(), T.data(), out.data());
    return out;
}

/// Translation u, v; 2D rotation; 2D scale (ModelSimilarity2D.hxx:14-26,52-72).
struct ModelSimilarity2D {
    std::vector<double> current_values;
    ModelSimilarity2D() : current_values(4, 0.0) {}
    Homography2D getInstance() const
    {
        Homography2D H;
        ecc_model_similarity_2d(current_values.data(), H.m);
        return H;
    }
};

/// Translation X, Y, Z; rotation about X, Y, Z (R = Rx * Ry * Rz); 3D scale (ModelSimilarity3D.hxx:16-31,64-87).
struct ModelSimilarity3D {
    std::vector<double> current_values;
    ModelSimilarity3D() : current_values(7, 0.0) {}
    Homography3D getInstance() const
    {
        Homography3D T;
        ecc_model_similarity_3d(current_values.data(), T.m);
        return T;
    }
};

/// P' = H2D * P * T3D; parameter vector = the four 2D parameters followed by the seven 3D ones.
class ModelCameraSimilarity2D3D {
    ProjectionMatrix P;
    std::set<int> active_parameters;  // LibOpterix::ParameterModel::active_parameters (LibOpterix/ParameterModel.hxx:60-66)

public:
    std::vector<double> current_values;
    /// All eleven parameters active.
    explicit ModelCameraSimilarity2D3D(const ProjectionMatrix& _P) : P(_P), current_values(11, 0.0)
    {
        for (int i = 0; i < 11; i++) active_parameters.insert(i);
    }
    /// As the reference's constructor (ModelCameraSimilarity2D3D.hxx:60-63): the set of parameters an optimiser may change.
    ModelCameraSimilarity2D3D(const ProjectionMatrix& _P, const std::set<int>& _active) : P(_P), active_parameters(_active), current_values(11, 0.0)
    {
        for (std::set<int>::const_iterator it = _active.begin(); it != _active.end(); ++it)
            if (*it < 0 || *it >= 11) throw std::invalid_argument("Parametrization::Model: Set of active parameters contains invalid indices.");
    }
    int numberOfParameters() const { return 11; }
    int numberOfParametersActive() const { return (int)active_parameters.size(); }
    const std::set<int>& activeParameters() const { return active_parameters; }
    /// Active entries of current_values (ParameterModel::restrict, LibOpterix/ParameterModel.hxx:89-96).
    std::vector<double> restrict() const
    {
        std::vector<double> x;
        for (std::set<int>::const_iterator it = active_parameters.begin(); it != active_parameters.end(); ++it) x.push_back(current_values[*it]);
        return x;
    }
    /// Writes x_active into the active entries (ParameterModel::expand, LibOpterix/ParameterModel.hxx:98-108).
    std::vector<double>& expand(const double* x_active = 0x0)
    {
        if (x_active) {
            int a = 0;
            for (std::set<int>::const_iterator it = active_parameters.begin(); it != active_parameters.end(); ++it, ++a) current_values[*it] = x_active[a];
        }
        return current_values;
    }
    const ProjectionMatrix& getOriginalProjectionMatrix() const { return P; }
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

/// ONE geometric correction for a whole FDCT trajectory (Models/ModelFDCTCalibrationCorrection.hxx:15-206): detector shift
/// u, v; yaw, pitch (detector shifts of tan(angle) sdd); roll and source-detector distance (rotation / scale about the mean
/// principal point); source-isocentre distance (3D scale).  P' = H * H_init * P * T, normalised.  `param`, `active`,
/// getTransforms() and transform() as in the reference (its Model<7> base keeps the parameter vector in `param`).
struct ModelFDCTCalibrationCorrection {
    double pp_u, pp_v, spacing, sid, sdd;
    double param[7];
    bool active[7];
    std::vector<std::string> names;

    static int size() { return 7; }
    static std::vector<std::string> ParameterNames()
    {
        const char* n[7] = {"Translation u", "Translation v", "Yaw", "Pitch", "Roll", "Source Isocenter Distance", "Source Detector Distance"};
        return std::vector<std::string>(n, n + 7);
    }
    enum ParameterSet { Identity, DetectorShifts, DetectorRigid2D, DetectorRotations, SDDandSID, All };

    ModelFDCTCalibrationCorrection(int _nu, int _nv, double _spacing, double _sid, double _sdd, ParameterSet set = Identity)
        : pp_u(0.5 * _nu), pp_v(0.5 * _nv), spacing(_spacing), sid(_sid), sdd(_sdd), names(ParameterNames())
    {
        for (int i = 0; i < 7; i++) param[i] = 0;
        setActiveParameters(set);
    }
    /// Mean principal point, source-isocentre and source-detector distance from the trajectory (computeMeanPPandSIDandSDD,
    /// :53-76; the focal lengths are |m1 x m3| and |m2 x m3| of the normalised matrix -- K(0,0), K(1,1) of the reference's RQ
    /// decomposition for a skew-free detector).
    ModelFDCTCalibrationCorrection(double _spacing, const std::vector<ProjectionMatrix>& Ps, ParameterSet set = Identity)
        : pp_u(0), pp_v(0), spacing(_spacing), sid(0), sdd(0), names(ParameterNames())
    {
        for (int i = 0; i < 7; i++) param[i] = 0;
        setActiveParameters(set);
        computeMeanPPandSIDandSDD(Ps);
    }
    void computeMeanPPandSIDandSDD(const std::vector<ProjectionMatrix>& Ps)
    {
        double sum_sid = 0, sum_sdd = 0, sum_u = 0, sum_v = 0;
        const int n = (int)Ps.size();
        if (!n) return;
        std::vector<double> flat(12 * (size_t)n);
        std::vector<float> A(12 * (size_t)n), C(4 * (size_t)n);
        for (int i = 0; i < n; i++) std::memcpy(&flat[12 * (size_t)i], Ps[i].data(), sizeof(double) * 12);
        ecc_derive_views_host(flat.data(), n, A.data(), C.data());
        for (int i = 0; i < n; i++) {
            double P[12];
            std::memcpy(P, &flat[12 * (size_t)i], sizeof(P));
            ecc_model_normalize(P);
            const double m1[3] = {P[0], P[3], P[6]}, m2[3] = {P[1], P[4], P[7]}, m3[3] = {P[2], P[5], P[8]};
            const double c1[3] = {m1[1] * m3[2] - m1[2] * m3[1], m1[2] * m3[0] - m1[0] * m3[2], m1[0] * m3[1] - m1[1] * m3[0]};
            const double c2[3] = {m2[1] * m3[2] - m2[2] * m3[1], m2[2] * m3[0] - m2[0] * m3[2], m2[0] * m3[1] - m2[1] * m3[0]};
            sum_sdd += spacing * 0.5 * (std::sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]) + std::sqrt(c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2]));
            const float* c = &C[4 * (size_t)i];
            sum_sid += std::sqrt((double)c[0] * c[0] + (double)c[1] * c[1] + (double)c[2] * c[2]);
            const double w = m3[0] * m3[0] + m3[1] * m3[1] + m3[2] * m3[2];  // pp = M m3^T, dehomogenised
            sum_u += (m1[0] * m3[0] + m1[1] * m3[1] + m1[2] * m3[2]) / w;
            sum_v += (m2[0] * m3[0] + m2[1] * m3[1] + m2[2] * m3[2]) / w;
        }
        sid = sum_sid / n;
        sdd = sum_sdd / n;
        pp_u = sum_u / n;
        pp_v = sum_v / n;
    }
    /// Several common sets of active parameters (:110-133)
    void setActiveParameters(ParameterSet set)
    {
        for (int i = 0; i < 7; i++) active[i] = (set == All);
        switch (set) {
            default:
            case Identity:
            case All: break;
            case DetectorShifts: active[0] = active[1] = true; break;
            case SDDandSID: active[5] = active[6] = true; break;
            case DetectorRotations: active[2] = active[3] = active[4] = true; break;
            case DetectorRigid2D: active[0] = active[1] = active[4] = true; break;
        }
    }
    int numberOfParametersActive() const
    {
        int k = 0;
        for (int i = 0; i < 7; i++) k += active[i] ? 1 : 0;
        return k;
    }
    /// Writes x_active into the active entries of `param` (Model::expand).
    void expand(const double* x_active)
    {
        for (int i = 0, a = 0; i < 7; i++)
            if (active[i]) param[i] = x_active[a++];
    }
    /// The homographies H and T of the current parameter vector (:150-203).
    std::pair<Homography2D, Homography3D> getTransforms() const
    {
        std::pair<Homography2D, Homography3D> HT;
        const double geom[4] = {pp_u, pp_v, sid, sdd};
        ecc_model_calibration_correction(geom, param, HT.first.m, HT.second.m);
        return HT;
    }
    /// Ps[i] = normalise(H * H_init * Ps[i] * T) (:136-148)
    void transform(std::vector<ProjectionMatrix>& Ps, const Homography2D& H_init = Homography2D()) const
    {
        double inst[25];
        instance(H_init, inst);
        Homography2D H;
        Homography3D T;
        std::memcpy(H.m, inst, sizeof(H.m));
        std::memcpy(T.m, inst + 9, sizeof(T.m));
        for (size_t i = 0; i < Ps.size(); i++) {
            Ps[i] = transformProjection(H, Ps[i], T);
            ecc_model_normalize(Ps[i].data());
        }
    }
    /// H * H_init (9 doubles) followed by T (16 doubles): one instance of MetricRadonIntermediate::evaluateBatchTransforms.
    void instance(const Homography2D& H_init, double* out25) const
    {
        const std::pair<Homography2D, Homography3D> HT = getTransforms();
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) {
                double acc = 0;
                for (int k = 0; k < 3; k++) acc += HT.first(r, k) * H_init(k, c);
                out25[r + 3 * c] = acc;
            }
        std::memcpy(out25 + 9, HT.second.m, sizeof(double) * 16);
    }
};

/// Parametrization of motion correction for FDCT (tools/FDCTMotionCorrection/ModelFDCT.hxx:9-64): every projection gets
/// its own instance of the stencil model; the raw parameter vector stacks the ACTIVE parameters view by view.
struct ModelFDCT {
    ModelCameraSimilarity2D3D stencil;  //< All projection matrices will abide to this model
    int n_active;                       //< Number of active parameters in stencil
    int n_proj;                         //< Number of projections
    std::vector<double> params;         //< Raw parameter vector (remains constant)
    std::vector<double> params_delta;   //< Raw parameter vector plus those currently being changed
    std::vector<ProjectionMatrix> trajectory;  //< current estimate of the trajectory

    explicit ModelFDCT(const ModelCameraSimilarity2D3D& _stencil) : stencil(_stencil), n_active(_stencil.numberOfParametersActive()), n_proj(0) {}

    /// Apply parameter vector for view i.  If view < 0, all parameters will be updated (ModelFDCT.hxx:26-62).
    std::vector<ProjectionMatrix>& applyModel(int view, const double* delta, const std::vector<ProjectionMatrix>& Ps)
    {
        stack(view, delta, (int)Ps.size());
        if ((int)trajectory.size() != n_proj) trajectory.resize(n_proj);
        for (int i = 0; i < n_proj; i++) {
            ModelCameraSimilarity2D3D HTi = stencil;
            HTi.setOriginalProjectionMatrix(Ps[i]);
            HTi.expand(&params_delta[(size_t)i * n_active]);
            trajectory[i] = HTi.getInstance();
        }
        return trajectory;
    }

    /// The same parameter stacking, expanded to the eleven model parameters per view: the input of the device-side
    /// expansion (MetricRadonIntermediate::evaluateBatchParams, FDCTMoCo::evaluateTrajectories).  Appends n_proj * 11 doubles.
    void appendExpanded(int view, const double* delta, int n_views, std::vector<double>& out)
    {
        stack(view, delta, n_views);
        for (int i = 0; i < n_proj; i++) {
            ModelCameraSimilarity2D3D HTi = stencil;
            HTi.current_values.assign(11, 0.0);
            const std::vector<double>& x = HTi.expand(&params_delta[(size_t)i * n_active]);
            out.insert(out.end(), x.begin(), x.end());
        }
    }

private:
    void stack(int view, const double* delta, int n_views)
    {
        n_proj = n_views;
        const size_t len = (size_t)n_proj * n_active;
        if (params.size() != len) params.assign(len, 0.0);
        params_delta.assign(len, 0.0);
        if (delta) {
            if (view < 0)
                for (size_t i = 0; i < len; i++) params_delta[i] = delta[i];
            else
                for (int i = 0; i < n_active; i++) params_delta[(size_t)view * n_active + i] = delta[i];
        }
        for (size_t i = 0; i < len; i++) params_delta[i] += params[i];
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

/// Motion correction of an FDCT trajectory (Gui/FDCTMotionCorrection.hxx:13-98): as SingleImageMotion, the cost of a
/// candidate matrix for the input image is the mean over ALL pairs (the reference's evaluate() returns ecc->evaluate()).
/// What the B200 path adds: K candidate parameter vectors of a ModelFDCT -- each n x n_active doubles, the whole
/// trajectory perturbed -- scored in ONE launch, expanded to matrices on the device (BASELINE config C4).
class FDCTMoCo : public SingleImageMotion {
public:
    FDCTMoCo(std::vector<ProjectionMatrix> _Ps, std::vector<RadonIntermediate*> _dtrs, int _input_index = 0)
        : SingleImageMotion(_Ps, _dtrs, _input_index)
    {}

    /// Means over all pairs for K candidate parameter vectors (candidates[k]: n * model.n_active doubles, view by view, as
    /// ModelFDCT::applyModel(-1, ...) takes them), applied to the initial matrices.  One launch; no matrix is built on the host.
    std::vector<double> evaluateTrajectories(Geometry::ModelFDCT& model, const std::vector<std::vector<double> >& candidates)
    {
        std::vector<double> params;
        const int n = (int)Ps.size();
        for (size_t k = 0; k < candidates.size(); k++) model.appendExpanded(-1, candidates[k].data(), n, params);
        return ecc->evaluateBatchParams(Ps, params, n);
    }

    /// The same for candidates of ONE view's parameters (view = input index unless given): the n-1 pairs with that view.
    std::vector<double> evaluateViewCandidates(Geometry::ModelFDCT& model, const std::vector<std::vector<double> >& candidates, int view = -1)
    {
        if (view < 0) view = input_index;
        const int n = (int)Ps.size();
        std::vector<double> params;
        for (size_t k = 0; k < candidates.size(); k++) {
            Geometry::ModelCameraSimilarity2D3D one = model.stencil;
            one.current_values.assign(11, 0.0);
            const std::vector<double>& x = one.expand(candidates[k].data());
            params.insert(params.end(), x.begin(), x.end());
        }
        std::vector<int> map(n, -1);
        map[view] = 0;
        std::vector<Eigen::Vector4i> idx;
        for (int i = 0; i < n; i++)
            if (i != view) idx.push_back(Eigen::Vector4i(view, i, view, i));
        return ecc->evaluateBatchParams(Ps, params, 1, map, &idx);
    }
};

/// What the B200 path adds for the calibration-correction loop (tools/FDCTCalibrationCorrection with
/// Models/ModelFDCTCalibrationCorrection.hxx: one parameter vector corrects the WHOLE trajectory, every cost evaluation
/// re-derives and uploads all n matrices): K candidate vectors of the model's ACTIVE parameters scored in ONE launch -- the
/// host builds K pairs of homographies (25 doubles each), the device applies them to all views, normalises, derives and
/// scores.  Returns the K means over all pairs.
inline std::vector<double> evaluateCalibrationCandidates(MetricRadonIntermediate& ecc, const std::vector<ProjectionMatrix>& Ps,
                                                         Geometry::ModelFDCTCalibrationCorrection model,
                                                         const std::vector<std::vector<double> >& candidates,
                                                         const Geometry::Homography2D& H_init = Geometry::Homography2D())
{
    std::vector<double> transforms(25 * candidates.size());
    for (size_t k = 0; k < candidates.size(); k++) {
        model.expand(candidates[k].data());
        model.instance(H_init, &transforms[25 * k]);
    }
    return ecc.evaluateBatchTransforms(Ps, transforms, 1, true);
}

}  // namespace EpipolarConsistency

#endif
