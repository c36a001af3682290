"""The C++ facade (include/EpipolarConsistency/*.h): builds with plain g++ against libecc_b200.so, keeps the
reference's file layout (checked against the reference's own NRRD reader/writer when oracle/_ref is present), and on
the GPU gives the same numbers as the Python mirror of the same ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "facade_check")


def build(target="facade_check"):
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), target], check=True, stdout=subprocess.DEVNULL)


def test_launcher_symbols_have_the_reference_signatures():
    """The reference's unmodified .cpp files link against two C++ symbols (EpipolarConsistencyRadonIntermediate.cpp:16-37,
    RadonIntermediate.cpp:12); libecc_b200.so must export them with exactly those (mangled) signatures."""
    lib = os.path.join(ROOT, "epipolarconsistency_b200", "lib", "libecc_b200.so")
    syms = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    # _Z19epipolarConsistencyiiiPciiffiPfS0_iPiS0_S0_ffbbS0_ : (int,int,int,char*,int,int,float,float,int,float*,float*,int,int*,float*,float*,float,float,bool,bool,float*)
    assert " T _Z19epipolarConsistencyiiiPciiffiPfS0_iPiS0_S0_ffbbS0_" in syms
    # _Z25computeDerivLineIntegralsyiiiiiiPf : (cudaTextureObject_t = unsigned long long,int,int,int,int,int,int,float*)
    assert " T _Z25computeDerivLineIntegralsyiiiiiiPf" in syms
    # _Z25cuda_computeLineIntegralssPfsS_syssS_ : (short,float*,short,float*,short,cudaTextureObject_t,short,short,float*) --
    # EpipolarConsistencyDirect.cpp:10-16, the direct metric's launcher
    assert " T _Z25cuda_computeLineIntegralssPfsS_syssS_" in syms
    build("launcher_swap_check")  # the reference's declarations, verbatim, link against the library


@pytest.mark.gpu
def test_launcher_swap_gives_the_abi_results(tmp_path):
    """INTEGRATION.md option B, compiled and run: the reference's call sequence through its own two launcher symbols
    (cudaArray textures, device handle table, K01s / out / out_corr buffers) gives the C ABI's bits."""
    build("launcher_swap_check")
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "launcher_swap_check")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK launcher swap" in r.stdout, r.stdout + r.stderr


def test_facade_builds_and_cpu_checks(tmp_path):
    build()
    r = subprocess.run([EXE, "cpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "OK cpu" in r.stdout, r.stdout + r.stderr


def test_ompl_layout(tmp_path):
    """One-matrix-per-line projection files (Projtable.hxx:168-220, EigenToStr.hxx:135-142): the facade's writer against an
    independent restatement of the format -- '[a b c d; e f g h; i j k l] ' with 12 significant digits -- and its reader."""
    build()
    subprocess.run([EXE, "cpu", str(tmp_path)], check=True, stdout=subprocess.DEVNULL)
    lines = open(tmp_path / "facade.ompl").read().split("\n")
    assert lines[0] == "# three views"
    assert lines[1] == '#> spacing="0.308" detector_size_px="1240 960"'
    Ps = ol.circular_trajectory(3, 750, 1200, 160, 128, 200, 2.0).reshape(3, 4, 3).transpose(0, 2, 1).copy()  # (n, row, col)
    Ps[1][0, 3] = 1.0 / 3.0
    Ps[2][2, 0] = -1234567.890123456
    for k in range(3):
        want = "[" + "; ".join(" ".join("%.12g" % v for v in row) for row in Ps[k]) + "] "
        assert lines[2 + k] == want
    assert lines[5] == ""


def test_nrrd_layout_against_reference_reader_and_writer(tmp_path):
    R = ol.ref_nrrd()
    if R is None:
        pytest.skip("oracle/_ref/libecc_ref_host.so (with the reference's NRRD code) not built")
    build()
    subprocess.run([EXE, "cpu", str(tmp_path)], check=True, stdout=subprocess.DEVNULL)
    # (1) a file written by the facade loads in the reference's reader
    out = np.zeros(64, np.float32)
    w, h = C.c_int(), C.c_int()
    val = C.create_string_buffer(64)
    n = R.ref_nrrd_load(str(tmp_path / "facade_dtr.nrrd").encode(), out, 64, C.byref(w), C.byref(h), b"Filter", val, 64)
    assert n == 35 and (w.value, h.value) == (7, 5) and val.value == b"Derivative"
    assert np.array_equal(out[:35], 0.25 * np.arange(35, dtype=np.float32) - 3)
    # (2) a file written by the reference's writer is byte-identical to the facade's for the same content
    data = (0.25 * np.arange(35, dtype=np.float32) - 3).astype(np.float32)
    assert R.ref_nrrd_save(str(tmp_path / "ref.nrrd").encode(), data, 7, 5, b"Filter", b"Derivative") == 1
    ref_bytes = open(tmp_path / "ref.nrrd", "rb").read()
    head, raw = ref_bytes.split(b"\n\n", 1)
    assert raw == data.tobytes()
    assert head.startswith(b"NRRD0004\ndimension: 2\nencoding: raw\nendian: little\nsizes: 7 5\n")
    assert ref_bytes == open(tmp_path / "facade_same.nrrd", "rb").read()
    # (3) and it loads in the facade's reader
    r = subprocess.run([EXE, "load", str(tmp_path / "ref.nrrd")], capture_output=True, text=True)
    assert r.stdout.split() == ["loaded", "7", "5", "%.9g" % float(data.astype(np.float64).sum()), "Derivative"], r.stdout


@pytest.mark.gpu
def test_facade_gpu_matches_python_mirror(tmp_path):
    from epipolarconsistency_b200 import api
    build()
    r = subprocess.run([EXE, "gpu", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "OK gpu" in r.stdout, r.stdout + r.stderr
    vals = {}
    pairs = {}
    for line in r.stdout.splitlines():
        t = line.split()
        if t[0] == "pair":
            pairs[(int(t[1]), int(t[2]))] = float(t[3])
        elif t[0] in ("radius", "mean", "subset", "fixed"):
            vals[t[0]] = float(t[1])
        elif t[0] in ("list", "batch", "model", "sim", "reg", "reg33", "pre", "signals", "fdct", "fdctview", "direct", "calib"):
            vals[t[0]] = [float(x) for x in t[1:]]
        elif t[0] == "l2s":
            vals.setdefault("l2s", []).append([float(x) for x in t[1:]])
    # the same scene through the Python mirror
    import torch
    n, n_u, n_v, n_a, n_t = 6, 160, 128, 128, 128
    ell = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5]])
    ctx = api.Context()
    Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 2.0)
    imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
    ctx.synth_projections(Ps, n_u, n_v, ell, imgs)
    dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_HYBRID_STATIC)  # the facade's default engine
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    cost = np.zeros((n, n), np.float32)
    mean = ctx.evaluate(cost)
    assert abs(vals["radius"] - ctx.get_object_radius()) < 1e-6
    assert abs(vals["mean"] - mean) <= 1e-7 * mean
    for (i, j), v in pairs.items():
        assert abs(v - cost[j, i]) <= 1e-6 * abs(cost[j, i]) + 1e-12
    idx = np.array([(1, 3, 1, 3), (1, 4, 1, 4), (3, 4, 3, 4)], np.int32)
    assert abs(vals["subset"] - ctx.evaluate_indices(idx)) <= 1e-7 * vals["subset"]
    # the same two matrix sets through the mirror: set 1 = view 3 shifted by 3 px in u (row0 += 3*row2; Ps are
    # column-major 3x4: entry (r,c) at r + 3c)
    sets = np.stack([Ps, Ps]).reshape(2, n, 12).copy()
    for c in range(4):
        sets[1, 3, 0 + 3 * c] += 3.0 * sets[1, 3, 2 + 3 * c]
    want_batch = ctx.evaluate_batch(sets)
    assert abs(vals["batch"][0] - mean) <= 1e-7 * mean
    assert abs(vals["batch"][0] - want_batch[0]) <= 1e-7 * mean
    assert abs(vals["batch"][1] - want_batch[1]) <= 1e-7 * want_batch[1]
    assert vals["batch"][1] > 1.1 * vals["batch"][0]
    # ---- evaluateForImagePair: value, number of samples, sum of squared differences, first and last kappa
    ctx.set_object_radius(50.0)
    ctx.set_epipolar_plane_step(0.002)
    sig = ctx.pair_signals(1, 4)
    v, cnt, ssd, k_first, k_last = vals["signals"]
    assert int(cnt) == len(sig["kappas"]) and cnt > 100
    assert abs(v - sig["value"]) <= 1e-7 * v
    assert abs(ssd - float(((sig["signal0"].astype(np.float64) - sig["signal1"]) ** 2).sum())) <= 1e-6 * ssd
    assert k_first == sig["kappas"][0] and k_last == sig["kappas"][-1] and k_first == -k_last
    # ---- adaptors: the same quantities through the mirror
    x = [1.5, -0.75, 0.004, 0.0, 0.8, -0.4, 0.3, 0.002, -0.003, 0.001, 0.0]
    P2 = api.camera_similarity_2d3d(Ps[2], x)
    assert np.allclose(vals["model"], P2, rtol=1e-11, atol=1e-11)
    ctx.set_object_radius(0.0)
    ctx.set_epipolar_plane_step(0.0)
    ctx.set_projection_matrices(Ps)
    all0 = ctx.evaluate(None)
    ctx.update_projection_matrix(2, P2)
    all1 = ctx.evaluate(None)
    idx2 = np.array([(2, i, 2, i) for i in range(n) if i != 2], np.int32)
    moving1 = ctx.evaluate_indices(idx2)
    assert abs(vals["sim"][0] - all0) <= 1e-7 * all0 and abs(vals["sim"][1] - all1) <= 1e-7 * all1
    assert abs(vals["sim"][2] - moving1) <= 1e-7 * moving1 and abs(vals["sim"][4] - moving1) <= 1e-6 * moving1
    assert all1 > all0
    ctx.set_projection_matrices(Ps)
    idx0 = np.array([(0, i, 0, i) for i in range(1, n)], np.int32)
    assert abs(vals["reg"][0] - ctx.evaluate_indices(idx0)) <= 1e-7 * vals["reg"][0]
    cross = np.array([(i, j, i, j) for j in range(2, n) for i in range(2)], np.int32)
    assert abs(vals["reg33"][0] - ctx.evaluate_indices(cross)) <= 1e-7 * vals["reg33"][0]
    T = api.similarity_3d([2.0, 0, 0, 0, 0.01, 0, 0])
    moved = Ps.copy()
    for i in range(2):
        moved[i] = (Ps[i].reshape(4, 3).T @ T).T.reshape(12)
    ctx.set_projection_matrices(moved)
    assert abs(vals["reg33"][1] - ctx.evaluate_indices(cross)) <= 1e-7 * vals["reg33"][1]
    # RadonIntermediate::sample / tex2D (CPU helpers of the reference): location = the reference's lineToSampleDtr (oracle, pinned
    # by golden vectors), value = clamped bilinear lookup at ((n_alpha-1) a, (n_t-1) d) in the 9 x 5 ramp x + 10 y
    lines = [(0.3, 0.9, 20.0), (-0.7, 0.2, -55.0), (0.1, -1.3, 3.0), (-0.4, -0.6, 80.0)]
    assert len(vals["l2s"]) == 4
    for line, (a, d, v) in zip(lines, vals["l2s"]):
        loc, flipped = ol.line_to_sample(np.array(line, np.float32), np.float32(40.98) * 5)
        assert abs(loc[0] - a) < 1e-6 and abs(loc[1] - d) < 1e-5
        x, y = 8 * np.clip(a, 0, 1), 4 * np.clip(d, 0, 1)
        assert abs(v - (x + 10 * y)) < 1e-3, (line, a, d, v)  # no sign flip in this CPU helper, as upstream
    # FDCTMoCo: K trajectories expanded on the device = the host models' matrices evaluated one by one (checked in C++);
    # here: the unperturbed candidate is the plain mean
    assert abs(vals["fdct"][0] - all0) <= 1e-7 * all0 and vals["fdct"][1] > all0 and vals["fdct"][2] > all0
    # PreProccess facade vs the oracle restatement (defaults: zero 1, feather 16; no low-pass in this check)
    want = ol.preprocess(np.full((n_v, n_u), 5.0, np.float32), sigma=0.0, feather=(16, 16, 16, 16), P=Ps[0])
    assert abs(vals["pre"][0] - want[n_v // 2, 8]) <= 2e-6 * 5 and abs(vals["pre"][1] - want[30, 40]) <= 2e-6 * 5
    ctx.set_projection_matrices(Ps)
    ctx.set_object_radius(50.0)
    ctx.set_epipolar_plane_step(0.002)
    assert abs(vals["fixed"] - ctx.evaluate(None)) <= 1e-7 * vals["fixed"]
    # MetricDirect / computeForImagePair through the facade = the Python mirror of the same ABI on the same images
    total, v13, n_lines, ssd, dk, free_v, fb, free_fb, radius = vals["direct"]
    ctx.set_object_radius(0.0)
    ctx.set_epipolar_plane_step(0.0)
    ctx.direct_set_images(imgs)
    ctx.direct_set_fan_beam(False)
    same = lambda a, b: abs(a - b) <= 2e-11 * abs(b)  # the facade check prints 12 digits
    assert same(total, ctx.direct_evaluate(None))
    r13 = ctx.direct_evaluate_pair(1, 3)
    assert same(v13, r13["value"]) and int(n_lines) == len(r13["kappas"]) and free_v == v13
    assert abs(ssd * ctx.direct_pair_geometry(1, 3)["dkappa"] - v13) <= 1e-9 * v13
    assert abs(radius - ol.object_radius(Ps[0], n_u, n_v)) <= 1e-9 * radius
    ctx.direct_set_fan_beam(True)
    assert same(fb, ctx.direct_evaluate_pair(1, 3)["value"]) and free_fb == fb
    ctx.direct_set_fan_beam(False)
    # calibration-correction candidates (ModelFDCTCalibrationCorrection): the facade's batch = the Python mirror's batch of the
    # same homographies, and its first candidate (all parameters zero) is the unperturbed mean
    ctx.set_projection_matrices(Ps)
    geom = [0.5 * n_u, 0.5 * n_v, 750.0, 2.0 * api.camera_intrinsics(Ps[0])[0]]  # the trajectory's own focal length
    xs = np.zeros((4, 7))
    xs[1, 0] = 1.5
    xs[2, 4], xs[2, 6] = 0.01, 12.0
    xs[3, 2], xs[3, 5] = 0.002, -8.0
    T = np.stack([api.model_calibration_correction(geom, x) for x in xs]).reshape(4, 1, 25)
    want = ctx.evaluate_batch_transforms(Ps, T, normalize=True)
    for k in range(4):
        assert abs(vals["calib"][k] - want[k]) <= 1e-6 * want[k]  # the facade estimates pp / SID / SDD from the matrices
    assert abs(vals["calib"][0] - all0) <= 1e-5 * all0  # identity correction, matrices re-normalised
    moved = ctx.transform_expand(Ps, T, normalize=True)
    for k in range(4):
        H, T3 = T[k, 0, :9].reshape(3, 3).T, T[k, 0, 9:].reshape(4, 4).T
        for v in range(n):
            Pm = H @ Ps[v].reshape(4, 3).T @ T3
            Pm = Pm / (np.linalg.norm(Pm[2, :3]) * np.sign(np.linalg.det(Pm[:, :3])))
            assert np.allclose(moved[k, v].reshape(4, 3).T, Pm, rtol=1e-12, atol=1e-9)
