// facade_check.cpp -- exercises the C++ facade (include/EpipolarConsistency/*.h) the way the reference's tools use
// the reference classes.  "cpu": file-format and API-surface checks without a GPU.  "gpu": Radon intermediates +
// metric through the facade on a small synthetic scene; prints numbers that tests/test_facade.py compares with
// the Python mirror of the same ABI.
#define ECC_FACADE_THROW
#include <EpipolarConsistency/EpipolarConsistencyRadonIntermediate.h>
#include <EpipolarConsistency/Adaptors.h>
#include <EpipolarConsistency/EpipolarConsistencyDirect.h>
#include <EpipolarConsistency/Projtable.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

using namespace EpipolarConsistency;

static int fail(const char* what)
{
    std::printf("FAIL %s\n", what);
    return 1;
}

static int run_cpu(const char* tmpdir)
{
    // NRRD round trip with the dtr meta keys (RadonIntermediate.cpp:96-103)
    NRRD::Image<float> img(7, 5);
    for (int i = 0; i < img.length(); i++) ((float*)img)[i] = 0.25f * i - 3.f;
    img.meta_info["Bin Size/Angle"] = toString(3.14159265358979 / 7);
    img.meta_info["Bin Size/Distance"] = toString(12.5);
    img.meta_info["Original Image/Width"] = toString(160);
    img.meta_info["Original Image/Height"] = toString(128);
    img.meta_info["Filter"] = "Derivative";
    const std::string path = std::string(tmpdir) + "/facade_dtr.nrrd";
    if (!img.save(path)) return fail("save");
    // header layout: magic, fields in lexical order, keys, offset comment, blank line
    std::ifstream f(path.c_str(), std::ios::binary);
    std::string line;
    const char* expect[] = {"NRRD0004", "dimension: 2", "encoding: raw", "endian: little", "sizes: 7 5", "spacings: 1 1", "type: float",
                            "Bin Size/Angle:=0.448799", "Bin Size/Distance:=12.5", "Filter:=Derivative", "Original Image/Height:=128",
                            "Original Image/Width:=160"};
    for (const char* e : expect) {
        std::getline(f, line);
        if (line != e) {
            std::printf("header line '%s' != '%s'\n", line.c_str(), e);
            return fail("header");
        }
    }
    std::getline(f, line);
    if (line.compare(0, 22, "# Offset to raw data: ") != 0) return fail("offset comment");
    const int offset = std::atoi(line.c_str() + 22);
    std::getline(f, line);
    if (!line.empty()) return fail("blank line");
    if ((int)f.tellg() != offset) return fail("offset value");
    f.close();
    NRRD::Image<float> back(path);
    if (!back || back.size(0) != 7 || back.size(1) != 5) return fail("load sizes");
    for (int i = 0; i < img.length(); i++)
        if (((float*)back)[i] != ((float*)img)[i]) return fail("load data");
    if (back.meta_info["Filter"] != "Derivative" || stringTo<int>(back.meta_info["Original Image/Width"]) != 160) return fail("load meta");
    // API surface of the compat types
    Geometry::ProjectionMatrix P;
    P(2, 3) = 5.0;
    if (P.data()[2 + 3 * 3] != 5.0) return fail("column-major");
    Eigen::Vector4i v(1, 2, 3, 4);
    if (v.data()[2] != 3) return fail("Vector4i");
    // a file with a single key, for the byte-for-byte comparison with the reference's writer
    NRRD::Image<float> one(7, 5);
    for (int i = 0; i < one.length(); i++) ((float*)one)[i] = 0.25f * i - 3.f;
    one.meta_info["Filter"] = "Derivative";
    if (!one.save(std::string(tmpdir) + "/facade_same.nrrd")) return fail("save same");
    // .ompl: one matrix per line, meta line, comment (Projtable.hxx:168-220)
    {
        std::vector<Geometry::ProjectionMatrix> Ps = ProjTable::makeCircularTrajectory(3, 750.0, 1200.0, 160, 128, 200.0, 2.0);
        Ps[1](0, 3) = 1.0 / 3.0;
        Ps[2](2, 0) = -1234567.890123456;
        const std::string ompl = std::string(tmpdir) + "/facade.ompl";
        if (!ProjTable::saveProjectionsOneMatrixPerLine(Ps, ompl, " three views", 0.308, 1240, 960)) return fail("ompl save");
        std::map<std::string, std::string> meta;
        std::vector<Geometry::ProjectionMatrix> back_ps = ProjTable::loadProjectionsOneMatrixPerLine(ompl, &meta);
        if (back_ps.size() != 3) return fail("ompl count");
        if (meta["comment"] != " three views" || meta["spacing"] != "0.308" || meta["detector_size_px"] != "1240 960") return fail("ompl meta");
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 12; k++) {
                const double a = Ps[i].data()[k], b = back_ps[i].data()[k];
                if (std::fabs(a - b) > 1e-11 * std::fabs(a)) return fail("ompl values (12 significant digits)");
            }
        if (ProjTable::stringToProjectionMatrix("1 2 3")(1, 1) != 1.0) return fail("ompl malformed line -> [I|0]");
    }
    std::printf("OK cpu\n");
    return 0;
}

static int run_load(const char* path)
{
    NRRD::Image<float> img(path);
    if (!img) return fail("load");
    double sum = 0;
    for (int i = 0; i < img.length(); i++) sum += ((float*)img)[i];
    std::printf("loaded %d %d %.9g %s\n", img.size(0), img.size(1), sum, img.meta_info["Filter"].c_str());
    return 0;
}

static int run_gpu(const char* tmpdir)
{
    const int n = 6, n_u = 160, n_v = 128, n_a = 128, n_t = 128;
    std::vector<double> flat(12 * n);
    ecc_make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 2.0, flat.data());
    std::vector<ProjectionMatrix> Ps(n);
    for (int i = 0; i < n; i++) std::memcpy(Ps[i].data(), &flat[12 * i], sizeof(double) * 12);
    // synthetic projections through the C ABI, copied to host images
    ecc_context* ctx = detail::shared_context();
    const double ell[14] = {0, 0, 0, 60, 40, 50, 1.0, 20, -10, 5, 20, 25, 15, 0.5};
    void* dev = 0x0;
    detail::check(ecc_device_alloc(ctx, sizeof(float) * (size_t)n * n_u * n_v, &dev), ctx, "alloc");
    detail::check(ecc_synth_projections(ctx, flat.data(), n, n_u, n_v, ell, 2, 1, 1, (float*)dev), ctx, "synth");
    std::vector<NRRD::Image<float> > images(n);
    std::vector<RadonIntermediate*> dtrs(n);
    for (int i = 0; i < n; i++) {
        images[i].set(n_u, n_v);
        detail::check(ecc_copy(ctx, (float*)images[i], (float*)dev + (size_t)i * n_u * n_v, sizeof(float) * n_u * n_v), ctx, "copy");
        dtrs[i] = new RadonIntermediate(images[i], n_a, n_t, RadonIntermediate::Derivative, RadonIntermediate::Identity);
    }
    ecc_device_free(ctx, dev);
    if (!dtrs[0]->isDerivative() || dtrs[0]->getRadonBinNumber(1) != n_t || dtrs[0]->getOriginalImageSize(0) != n_u) return fail("dtr props");
    {   // the constructors above only staged their images (one batched launch when the first result is asked for); an
        // object computed on its own, the reference's way, has the same bits (static-split engine: batch invariant)
        RadonIntermediate::setDeferredCompute(false);  // runs the staged batch
        RadonIntermediate alone(images[3], n_a, n_t, RadonIntermediate::Derivative, RadonIntermediate::Identity);
        RadonIntermediate::setDeferredCompute(true);
        alone.readback();
        dtrs[3]->readback();
        if (std::memcmp((const float*)alone.data(), (const float*)dtrs[3]->data(), sizeof(float) * n_a * n_t) != 0) return fail("deferred batch vs single computation");
        dtrs[3]->clearRawData();
        // RadonIntermediateFunction::compute (Gui/ComputeRadonIntermediate.hxx:70-83): same dtr, meta side effects on the IMAGE
        RadonIntermediateFunction fn;
        fn.number_of_bins.angle = n_a;
        fn.number_of_bins.distance = n_t;
        double mm = 0.308;
        RadonIntermediate* via_fn = fn.compute(images[3], &Ps[3], &mm);
        via_fn->readback();
        if (std::memcmp((const float*)alone.data(), (const float*)via_fn->data(), sizeof(float) * n_a * n_t) != 0) return fail("RadonIntermediateFunction::compute data");
        delete via_fn;
        if (images[3].meta_info["Original Image/Pixel Spacing"] != "0.308") return fail("meta: pixel spacing");
        const std::string pm = images[3].meta_info["Original Image/Projection Matrix"];
        if (pm.empty() || pm[0] != '[' || pm.substr(pm.size() - 2) != "] " || pm != ProjTable::toString(Ps[3])) return fail("meta: projection matrix");
        if (ProjTable::stringToProjectionMatrix(pm)(1, 2) == 0.0) return fail("meta: projection matrix round trip");
        std::printf("metapm %s\n", pm.c_str());
        // BindlessTexture2D: `tex` / `array` appear on demand for a view, at once for an uploaded image; readback returns the data
        UtilsCuda::BindlessTexture2D<float>* t = dtrs[1]->getTexture();
        if (t->tex != 0 || t->array != 0x0 || !t->normalizedCoords || t->size[0] != n_a || t->size[1] != n_t) return fail("texture view before use");
        const unsigned long long handle = *t;
        if (handle == 0 || t->tex != handle || t->array == 0x0) return fail("texture view after use");
        UtilsCuda::MemoryBlock<float> back;
        t->readback(back);
        std::vector<float> host(n_a * n_t);
        back.readback(host.data());
        dtrs[1]->readback();
        if (std::memcmp(host.data(), (const float*)dtrs[1]->data(), sizeof(float) * n_a * n_t) != 0) return fail("texture readback");
        UtilsCuda::BindlessTexture2D<float> up(n_u, n_v, (const float*)images[0]);
        if (up.tex == 0 || up.array == 0x0 || up.normalizedCoords) return fail("uploaded texture");
        RadonIntermediate from_tex(up, n_a, n_t, RadonIntermediate::Derivative, RadonIntermediate::Identity);
        from_tex.readback();
        dtrs[0]->readback();
        if (std::memcmp((const float*)from_tex.data(), (const float*)dtrs[0]->data(), sizeof(float) * n_a * n_t) != 0) return fail("dtr from a GPU-resident image");
        // CPU sampling helpers of the reference (RadonIntermediate.h:85-108): texel mapping (n-1) s, clamped bilinear
        NRRD::Image<float> ramp(9, 5);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 9; x++) ramp.pixel(x, y) = (float)(x + 10 * y);
        ramp.meta_info["Bin Size/Angle"] = "0.3490658503988659";
        ramp.meta_info["Bin Size/Distance"] = "40.98";
        ramp.meta_info["Original Image/Width"] = "160";
        ramp.meta_info["Original Image/Height"] = "128";
        ramp.meta_info["Filter"] = "Derivative";
        RadonIntermediate small(ramp);
        if (std::fabs(small.tex2D(0.5f, 0.25f) - (4.0f + 10.0f)) > 1e-5f || small.tex2D(0.f, 0.f) != 0.f || small.tex2D(1.f, 1.f) != 48.f) return fail("tex2D");
        if (std::fabs(small.tex2D(0.3f, 0.6f) - (8 * 0.3f + 10 * 4 * 0.6f)) > 1e-4f) return fail("tex2D bilinear");
        const float lines[4][3] = {{0.3f, 0.9f, 20.f}, {-0.7f, 0.2f, -55.f}, {0.1f, -1.3f, 3.f}, {-0.4f, -0.6f, 80.f}};
        for (int q = 0; q < 4; q++) {
            float l[3] = {lines[q][0], lines[q][1], lines[q][2]};
            const float v = small.sample(l);
            std::printf("l2s %.9g %.9g %.9g\n", l[0], l[1], v);
        }
    }
    // save / reload a dtr: identical data and properties
    dtrs[2]->readback();
    const std::string path = std::string(tmpdir) + "/facade_dtr2.nrrd";
    if (!dtrs[2]->data().save(path)) return fail("dtr save");
    RadonIntermediate reloaded(path);
    reloaded.readback();
    for (int k = 0; k < n_a * n_t; k++)
        if (((const float*)reloaded.data())[k] != ((const float*)dtrs[2]->data())[k]) return fail("dtr reload data");
    if (std::fabs(reloaded.getRadonBinSize(1) - dtrs[2]->getRadonBinSize(1)) > 1e-5 * dtrs[2]->getRadonBinSize(1)) return fail("dtr reload props");

    MetricRadonIntermediate ecc(Ps, dtrs);
    std::vector<float> cost(n * n, -1.f);
    const double mean = ecc.evaluate(cost.data());
    std::printf("radius %.9g\n", ecc.getObjectRadius());
    std::printf("mean %.9g\n", mean);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) std::printf("pair %d %d %.9g\n", i, j, cost[i + j * n]);
    if (cost[0] != -1.f) return fail("untouched entries");
    std::set<int> views;
    views.insert(1); views.insert(3); views.insert(4);
    std::printf("subset %.9g\n", ecc.evaluate(views));
    std::vector<Eigen::Vector4i> idx;
    idx.push_back(Eigen::Vector4i(0, 5, 0, 5));
    idx.push_back(Eigen::Vector4i(2, 1, 2, 1));
    float out2[2];
    std::printf("list %.9g %.9g %.9g\n", ecc.evaluate(idx, out2), out2[0], out2[1]);
    std::vector<std::vector<ProjectionMatrix> > sets(2, Ps);
    sets[1][3](0, 3) += 3.0 * sets[1][3](2, 3);  // detector shift of view 3 by 3 px in u
    for (int c = 0; c < 3; c++) sets[1][3](0, c) += 3.0 * sets[1][3](2, c);
    std::vector<double> means = ecc.evaluateBatch(sets);
    std::printf("batch %.9g %.9g\n", means[0], means[1]);
    ecc.setObjectRadius(50.0).setEpipolarPlaneStep(0.002);
    std::printf("fixed %.9g\n", ecc.evaluate());
    {   // evaluateForImagePair: the redundant signals of one pair (visualisation interface of the reference)
        std::vector<float> s0, s1, ks;
        std::vector<std::pair<float, float> > r0, r1;
        const double v = ecc.evaluateForImagePair(1, 4, &s0, &s1, &ks, &r0, &r1);
        double ssd = 0;
        for (size_t q = 0; q < s0.size(); q++) ssd += (double)(s0[q] - s1[q]) * (s0[q] - s1[q]);
        std::printf("signals %.9g %d %.9g %.9g %.9g\n", v, (int)s0.size(), ssd, (double)ks.front(), (double)ks.back());
        if (s1.size() != s0.size() || ks.size() != s0.size() || r0.size() != s0.size() || r1.size() != s0.size()) return fail("signal sizes");
    }
    // ---- adaptors: SingleImageMotion, Registration, Registration3D3D, the similarity models
    {
        Geometry::ModelCameraSimilarity2D3D model(Ps[2]);
        const double x[11] = {1.5, -0.75, 0.004, 0.0, 0.8, -0.4, 0.3, 0.002, -0.003, 0.001, 0.0};
        model.current_values.assign(x, x + 11);
        const ProjectionMatrix P2 = model.getInstance();
        std::printf("model");
        for (int k = 0; k < 12; k++) std::printf(" %.12g", P2.data()[k]);
        std::printf("\n");
        SingleImageMotion sim(Ps, dtrs, 2);
        sim.getMetricPtr().setObjectRadius(0).setEpipolarPlaneStep(0);
        const double all0 = sim.evaluate(Ps[2]);
        const double all1 = sim.evaluate(P2);
        std::vector<float> per_pair(n - 1);
        const double moving1 = sim.evaluateMovingPairs(P2, per_pair.data());
        // the same step again and again: plain path, recorded, replayed -- always the same bits
        for (int rep = 0; rep < 4; rep++) {
            std::vector<float> again(n - 1);
            if (sim.evaluateMovingPairs(P2, again.data()) != moving1) return fail("SingleImageMotion graph replay mean");
            for (int k = 0; k < n - 1; k++)
                if (again[k] != per_pair[k]) return fail("SingleImageMotion graph replay values");
        }
        std::vector<ProjectionMatrix> cands;
        cands.push_back(Ps[2]);
        cands.push_back(P2);
        const std::vector<double> cm = sim.evaluateCandidates(cands);
        std::printf("sim %.9g %.9g %.9g %.9g %.9g\n", all0, all1, moving1, cm[0], cm[1]);
        double s = 0;
        for (int k = 0; k < n - 1; k++) s += per_pair[k];
        if (std::fabs(s / (n - 1) - moving1) > 1e-6 * moving1) return fail("SingleImageMotion per-pair values");
        if (std::fabs(cm[1] - moving1) > 1e-6 * moving1) return fail("SingleImageMotion candidates vs sequential");
        Registration reg(Ps, dtrs);
        const double r0 = reg.evaluate(Ps[0]);
        std::printf("reg %.9g\n", r0);
        std::vector<ProjectionMatrix> Psrc(Ps.begin(), Ps.begin() + 2), Ptgt(Ps.begin() + 2, Ps.end());
        std::vector<RadonIntermediate*> dsrc(dtrs.begin(), dtrs.begin() + 2), dtgt(dtrs.begin() + 2, dtrs.end());
        Registration3D3D r33(false, Psrc, dsrc, Ptgt, dtgt);
        Geometry::ModelSimilarity3D m3;
        const double id = r33.evaluate(m3.getInstance());
        m3.current_values[0] = 2.0;  // 2 mm along X
        m3.current_values[4] = 0.01;
        const double moved = r33.evaluate(m3.getInstance());
        std::vector<Geometry::Homography3D> Ts;
        Ts.push_back(Geometry::Homography3D());
        Ts.push_back(m3.getInstance());
        const std::vector<double> rm = r33.evaluateCandidates(Ts);
        std::printf("reg33 %.9g %.9g %.9g %.9g\n", id, moved, rm[0], rm[1]);
        if (std::fabs(rm[0] - id) > 1e-6 * id || std::fabs(rm[1] - moved) > 1e-6 * moved) return fail("Registration3D3D candidates vs sequential");
        if (!(moved > id)) return fail("Registration3D3D: a displaced source must score worse");
    }
    {   // FDCTMoCo + ModelFDCT: per-view parameter stacking, K trajectories per launch expanded on the device
        std::set<int> active;
        active.insert(0);   // translation u
        active.insert(1);   // translation v
        active.insert(4);   // translation X
        active.insert(8);   // rotation about Y
        Geometry::ModelCameraSimilarity2D3D stencil(Ps[0], active);
        Geometry::ModelFDCT model(stencil);
        if (model.n_active != 4) return fail("ModelFDCT n_active");
        std::vector<std::vector<double> > cands(3, std::vector<double>((size_t)n * 4, 0.0));
        for (int i = 0; i < n; i++) {  // candidate 1: every view moved a little, differently; candidate 2: view 3 only
            cands[1][4 * i + 0] = 0.3 * (i % 3) - 0.2;
            cands[1][4 * i + 1] = 0.1 * i;
            cands[1][4 * i + 2] = 0.5 - 0.2 * i;
            cands[1][4 * i + 3] = 0.001 * (i + 1);
        }
        cands[2][4 * 3 + 0] = 2.5;
        cands[2][4 * 3 + 3] = -0.004;
        FDCTMoCo moco(Ps, dtrs, 3);
        moco.getMetricPtr().setObjectRadius(0).setEpipolarPlaneStep(0);
        const std::vector<double> batch = moco.evaluateTrajectories(model, cands);
        // the reference's way: matrices on the host (applyModel), one full evaluation per candidate
        MetricRadonIntermediate seq(Ps, dtrs);
        double one_by_one[3];
        for (int k = 0; k < 3; k++) {
            const std::vector<ProjectionMatrix>& traj = model.applyModel(-1, cands[k].data(), Ps);
            seq.setProjectionMatrices(traj);
            one_by_one[k] = seq.evaluate();
        }
        std::printf("fdct %.12g %.12g %.12g %.12g %.12g %.12g\n", batch[0], batch[1], batch[2], one_by_one[0], one_by_one[1], one_by_one[2]);
        for (int k = 0; k < 3; k++)
            if (std::fabs(batch[k] - one_by_one[k]) > 1e-12 * std::fabs(one_by_one[k])) return fail("FDCTMoCo: device-expanded trajectories vs host models, one by one");
        if (!(batch[1] > batch[0] && batch[2] > batch[0])) return fail("FDCTMoCo: a perturbed trajectory must score worse");
        // applyModel(view, delta): only that view changes, by exactly the stencil model
        const double d3[4] = {2.5, 0.0, 0.0, -0.004};
        const std::vector<ProjectionMatrix>& t3 = model.applyModel(3, d3, Ps);
        Geometry::ModelCameraSimilarity2D3D single(Ps[3], active);
        single.expand(d3);
        const ProjectionMatrix want3 = single.getInstance();
        for (int q = 0; q < 12; q++)
            if (t3[3].data()[q] != want3.data()[q] || t3[2].data()[q] != Ps[2].data()[q]) return fail("ModelFDCT::applyModel(view)");
        // candidates for one view: the n-1 pairs with it, = SingleImageMotion::evaluateCandidates of the host-built matrices
        std::vector<std::vector<double> > vc(2, std::vector<double>(4, 0.0));
        vc[1][0] = 2.5;
        vc[1][3] = -0.004;
        const std::vector<double> by_params = moco.evaluateViewCandidates(model, vc);
        std::vector<ProjectionMatrix> mats;
        mats.push_back(Ps[3]);
        mats.push_back(want3);
        const std::vector<double> by_mats = moco.evaluateCandidates(mats);
        std::printf("fdctview %.12g %.12g %.12g %.12g\n", by_params[0], by_params[1], by_mats[0], by_mats[1]);
        if (std::fabs(by_params[0] - by_mats[0]) > 1e-12 * by_mats[0] || std::fabs(by_params[1] - by_mats[1]) > 1e-12 * by_mats[1]) return fail("FDCTMoCo::evaluateViewCandidates vs evaluateCandidates");
    }
    {   // ModelFDCTCalibrationCorrection (Models/ModelFDCTCalibrationCorrection.hxx): one correction for the whole trajectory;
        // K candidates expanded on the device in one launch = the host model's transform() + one evaluation per candidate
        Geometry::ModelFDCTCalibrationCorrection calib(2.0, Ps, Geometry::ModelFDCTCalibrationCorrection::All);
        if (std::fabs(calib.pp_u - 0.5 * n_u) > 1e-6 || std::fabs(calib.pp_v - 0.5 * n_v) > 1e-6) return fail("calibration model: mean principal point");
        double f_px = 0, u0 = 0, v0 = 0;  // makeCircularTrajectory's focal length is n_v / (2 tan(atan(n_v px / sdd) / 2)), not sdd / px
        ecc_camera_intrinsics(Ps[0].data(), &f_px, &u0, &v0);
        if (std::fabs(calib.sid - 750.0) > 1e-3 || std::fabs(calib.sdd - 2.0 * f_px) > 1e-6 * calib.sdd) return fail("calibration model: mean SID / SDD");
        if (calib.numberOfParametersActive() != 7 || calib.ParameterNames()[6] != "Source Detector Distance") return fail("calibration model: parameters");
        std::vector<std::vector<double> > cc(4, std::vector<double>(7, 0.0));
        cc[1][0] = 1.5;                       // detector shift u
        cc[2][4] = 0.01; cc[2][6] = 12.0;     // roll and SDD
        cc[3][2] = 0.002; cc[3][5] = -8.0;    // yaw and SID
        MetricRadonIntermediate ecc(Ps, dtrs);
        ecc.setObjectRadius(0).setEpipolarPlaneStep(0);
        const std::vector<double> batch = evaluateCalibrationCandidates(ecc, Ps, calib, cc);
        double one_by_one[4];
        for (int k = 0; k < 4; k++) {
            std::vector<ProjectionMatrix> moved = Ps;
            calib.expand(cc[k].data());
            calib.transform(moved);
            ecc.setProjectionMatrices(moved);
            one_by_one[k] = ecc.evaluate();
        }
        std::printf("calib %.12g %.12g %.12g %.12g %.12g %.12g %.12g %.12g\n", batch[0], batch[1], batch[2], batch[3], one_by_one[0], one_by_one[1],
                    one_by_one[2], one_by_one[3]);
        for (int k = 0; k < 4; k++)
            if (std::fabs(batch[k] - one_by_one[k]) > 1e-12 * std::fabs(one_by_one[k])) return fail("calibration candidates: device-expanded vs host transform(), one by one");
        // detector shift and yaw move the image content: they must score worse.  (Roll + SDD, candidate 2, need not: on this coarse
        // scene a one per cent magnification lowers the interpolation residuals the consistent geometry is left with.)
        if (!(batch[1] > batch[0] && batch[3] > batch[0])) return fail("calibration candidates: a shifted detector must score worse");
        // a detector-shift-only parameter set: two active parameters
        Geometry::ModelFDCTCalibrationCorrection shifts(n_u, n_v, 2.0, 750.0, 1200.0, Geometry::ModelFDCTCalibrationCorrection::DetectorShifts);
        if (shifts.numberOfParametersActive() != 2 || !shifts.active[0] || !shifts.active[1] || shifts.active[4]) return fail("calibration model: DetectorShifts");
    }
    {   // PreProccess facade: defaults clear the border and feather 16 px; cosine weight 1 at the principal point
        NRRD::Image<float> im(n_u, n_v);
        for (int i = 0; i < im.length(); i++) ((float*)im)[i] = 5.f;
        PreProccess pre;
        pre.lowpass.gaussian_sigma = 0;
        pre.processAndWeight(im, Ps[0]);
        const float* q = (const float*)im;
        if (q[0] != 0.f || q[n_u * (n_v / 2)] != 0.f) return fail("PreProccess border");
        if (std::fabs(q[n_u * (n_v / 2) + n_u / 2] - 5.f) > 1e-5f) return fail("PreProccess centre");
        if (!(q[n_u * (n_v / 2) + 8] > 0.f && q[n_u * (n_v / 2) + 8] < 5.f)) return fail("PreProccess feather");
        std::printf("pre %.9g %.9g\n", q[n_u * (n_v / 2) + 8], q[n_u * 30 + 40]);
    }
    {   // MetricDirect / computeForImagePair (EpipolarConsistencyDirect.h): the metric straight from the images
        std::vector<UtilsCuda::BindlessTexture2D<float>*> Is(n);
        for (int i = 0; i < n; i++) Is[i] = new UtilsCuda::BindlessTexture2D<float>(n_u, n_v, (const float*)images[i]);
        MetricDirect direct(Ps, Is);
        if (direct.getNumberOfProjetions() != n) return fail("MetricDirect::getNumberOfProjetions");
        std::vector<float> cost((size_t)n * n, -1.f);
        const double total = direct.evaluate(cost.data());
        std::vector<float> s0, s1, kap;
        const double v13 = direct.evaluateForImagePair(1, 3, &s0, &s1, &kap);
        if (kap.empty() || s0.size() != 3 * kap.size() || s1.size() != 3 * kap.size()) return fail("MetricDirect sample vector sizes");
        if (cost[1 + 3 * n] != (float)v13 || cost[3 + 1 * n] != -1.f || cost[0] != -1.f) return fail("MetricDirect cost image layout");
        double ssd = 0;
        for (size_t q = 0; q < kap.size(); q++) ssd += (double)((s0[q] - s1[q]) * (s0[q] - s1[q]));
        // the caller's kappas: the same planes give the same signals
        std::vector<float> t0, t1, given(kap.begin() + 5, kap.begin() + 40);
        direct.evaluateForImagePair(1, 3, &t0, &t1, &given);
        if (given.size() != 35 || std::memcmp(t0.data(), s0.data() + 5, sizeof(float) * 35) != 0) return fail("MetricDirect with given kappas");
        // the free function on the same two views with the metric's radius and step: the same value
        const double dk = (double)(kap[1] - kap[0]);
        const double free_v = computeForImagePair(Ps[1], Ps[3], *Is[1], *Is[3], 0.0, direct.getObjectRadius());
        direct.setFanBeamConsistency(true);
        const double fb = direct.evaluateForImagePair(1, 3);
        const double free_fb = computeForImagePair(Ps[1], Ps[3], *Is[1], *Is[3], 0.0, direct.getObjectRadius(), true);
        std::printf("direct %.12g %.12g %d %.12g %.9g %.12g %.12g %.12g %.12g\n", total, v13, (int)kap.size(), ssd, dk, free_v, fb, free_fb,
                    direct.getObjectRadius());
        if (free_v != v13 || free_fb != fb) return fail("computeForImagePair vs MetricDirect::evaluateForImagePair");
        for (auto t : Is) delete t;
    }
    for (auto d : dtrs) delete d;
    std::printf("OK gpu\n");
    return 0;
}

int main(int argc, char** argv)
{
    const char* mode = argc > 1 ? argv[1] : "cpu";
    const char* tmpdir = argc > 2 ? argv[2] : "/tmp";
    try {
        if (std::strcmp(mode, "load") == 0) return run_load(tmpdir);
        return std::strcmp(mode, "gpu") == 0 ? run_gpu(tmpdir) : run_cpu(tmpdir);
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
}
