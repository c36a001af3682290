"""CPU tests: pin the oracle against the reference's own code (golden vectors generated from the
reference headers, tests/golden/make_ref_host_vectors.py) and against the invariants the reference
states in code and comments."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_host_vectors.npz")
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4]])


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def small_scene():
    n, n_u, n_v = 8, 160, 128
    Ps = ol.circular_trajectory(n, 750, 1200, n_u, n_v, 200, 2.0)
    imgs = np.stack([ol.project_ellipsoids(P, n_u, n_v, ELL) for P in Ps])
    dtrs = np.stack([ol.radon(im, 192, 192) for im in imgs])
    return dict(n=n, n_u=n_u, n_v=n_v, Ps=Ps, imgs=imgs, dtrs=dtrs)


# ---- golden vectors from the reference headers ---------------------------------------------------
@pytest.mark.parametrize("n", [2, 3, 7, 100, 496])
def test_get_ij_matches_reference(gold, n):
    tab = gold[f"get_ij_{n}"]
    mine = np.array([ol.get_ij(k, n) for k in range(tab.shape[0])], np.int32)
    assert np.array_equal(mine, tab)


def test_get_ij_enumerates_every_pair_once():
    n = 37
    seen = {ol.get_ij(k, n) for k in range(n * (n - 1) // 2)}
    assert seen == {(i, j) for i in range(n) for j in range(i + 1, n)}


def test_pinv_and_source_match_reference(gold):
    for P, A_ref, C_ref in zip(gold["Ps"], gold["pinvT"], gold["Cs"]):
        A = ol.pinv_transpose(P)
        Cc = ol.source_position(P)
        assert np.max(np.abs(A - A_ref)) <= 2e-6 * np.max(np.abs(A_ref))
        assert np.max(np.abs(Cc[:3] - C_ref[:3])) <= 2e-6 * np.max(np.abs(C_ref[:3])) + 1e-9
        assert Cc[3] == 1.0


def test_culaut_residuals(gold):
    """P*C ~ 0 and P*(PinvT)^T ~ I (the checks of the reference's TestCudaUtils.cpp:39-57)."""
    for P in gold["Ps"][:8]:
        Pm = P.reshape(4, 3).T
        A = ol.pinv_transpose(P).astype(np.float64).reshape(4, 3).T
        Cc = ol.source_position(P).astype(np.float64)
        assert np.max(np.abs(Pm @ Cc)) < 1e-3 * np.max(np.abs(Pm))
        assert np.max(np.abs(Pm @ A.T - np.eye(3))) < 1e-5


def test_compute_k01_matches_reference(gold):
    Cs, A, pairs = gold["Cs"], gold["pinvT"], gold["pairs"]
    for s, (radius, dk) in enumerate(gold["settings"]):
        for q, (a, b) in enumerate(pairs):
            K0, K1 = ol.compute_k01(160.0, 128.0, Cs[a], Cs[b], A[a], A[b], radius, 819.6, dk)
            ref = gold["K01"][s, q]
            got = np.concatenate([K0, K1])
            scale = np.maximum(np.abs(ref), 1e-3)
            assert np.max(np.abs(got - ref) / scale) < 2e-5, (s, q, got, ref)


def test_line_to_sample_matches_reference(gold):
    for line, ref in zip(gold["lines"], gold["samples"]):
        l, flipped = ol.line_to_sample(line, 409.8)
        assert flipped == int(ref[2])
        assert abs(l[0] - ref[0]) < 1e-6 and abs(l[1] - ref[1]) < 1e-5 * max(1.0, abs(ref[1]))


def test_live_reference_headers_if_built(gold):
    """When oracle/_ref is present, call the reference's compiled headers directly on fresh inputs."""
    R = ol.ref_host()
    if R is None:
        pytest.skip("oracle/_ref/libecc_ref_host.so not built")
    rng = np.random.default_rng(7)
    Ps = ol.circular_trajectory(5, 600, 1100, 200, 180, 360, 1.5) * (1 + 1e-3 * rng.standard_normal((5, 12)))
    for P in Ps:
        a = np.zeros(12, np.float32)
        R.ref_pinv_transpose(P, a)
        assert np.allclose(ol.pinv_transpose(P), a, rtol=2e-6, atol=1e-12)
    for k in range(10):
        i, j = C.c_int(), C.c_int()
        R.ref_get_ij(k, 5, C.byref(i), C.byref(j))
        assert (i.value, j.value) == ol.get_ij(k, 5)


# ---- invariants of the Radon intermediate ----------------------------------------------------------
def test_radon_layout_and_zero_rows(small_scene):
    d = small_scene["dtrs"]
    assert d.shape == (8, 192, 192)
    # lines at |t| = diag/2 miss the image: exactly zero (RadonIntermediate.cu:89-92)
    assert np.all(d[:, 0, :] == 0)
    assert np.abs(d).max() > 1.0


def test_radon_constant_image_line_length():
    """No-filter Radon of a constant image = chord length of the one-pixel-inset box (within a step)."""
    n_u, n_v, n_a, n_t = 96, 64, 64, 64
    r = ol.radon(np.ones((n_v, n_u), np.float32), n_a, n_t, filter=2)
    iy, ix = n_t // 2, n_a // 2  # alpha = 0, tau = 0: horizontal line through the centre
    assert abs(r[iy, ix] - (n_u - 2)) <= 0.67
    # alpha = -pi/2: vertical line through the centre
    assert abs(r[iy, 0] - (n_v - 2)) <= 0.67


def test_radon_derivative_is_difference_of_neighbouring_lines():
    """Derivative filter of a ramp image I(u,v)=v: horizontal lines one pixel apart differ by -length."""
    n_u, n_v = 80, 60
    img = np.tile(np.arange(n_v, dtype=np.float32)[:, None], (1, n_u))
    r = ol.radon(img, 64, 64, filter=0)
    val = r[32, 32]  # alpha=0: normal (0,1); line at +1/2 minus line at -1/2 -> +1 per unit length
    n_samples = np.floor((n_u - 2) / np.float32(0.66)) + 1
    assert abs(val - n_samples * 0.66) < 1e-2 * abs(val)


def test_radon_tex8_close_to_exact(small_scene):
    im = small_scene["imgs"][0]
    a = small_scene["dtrs"][0]
    b = ol.radon(im, 192, 192, interp=ol.INTERP_TEX8)
    rel = np.abs(a - b).max() / np.abs(a).max()
    assert 0 < rel < 1e-2  # quantised weights differ, but only at the 1e-3 level


def test_radon_sample_count_formula():
    n_u, n_v, n_a, n_t = 160, 128, 96, 96
    cnt = ol.radon_num_samples(n_u, n_v, n_a, n_t)
    # mean chord = area of the inset box / diagonal ... per bin: 2 lines * chord / 0.66
    est = 2.0 * (n_u - 2) * (n_v - 2) / (np.hypot(n_u, n_v) * 0.66) * n_a * n_t
    assert abs(cnt - est) / est < 0.05


# ---- invariants of the metric --------------------------------------------------------------------
def test_grangeat_consistency(small_scene):
    s = small_scene
    mean, out, ks = ol.ecc(s["Ps"], s["dtrs"], s["n_u"], s["n_v"], want_ksamples=True)
    assert mean > 0
    Pp = s["Ps"].copy()
    H = np.eye(3)
    H[0, 2], H[1, 2] = 3.0, 2.0  # detector shift of view 2 by (3,2) px
    Pp[2] = (H @ s["Ps"][2].reshape(4, 3).T).T.reshape(12)
    mean_p, _, _ = ol.ecc(Pp, s["dtrs"], s["n_u"], s["n_v"])
    assert mean_p > 10 * mean
    # kappa sample count: auto dkappa = kappa_max/diag -> about diag samples per pair
    assert np.all(np.abs(ks - np.hypot(s["n_u"], s["n_v"])) <= 1)


def test_cost_image_layout_and_mean(small_scene):
    s = small_scene
    n = s["n"]
    mean, out, _ = ol.ecc(s["Ps"], s["dtrs"], s["n_u"], s["n_v"])
    vals = [out[j, i] for i in range(n) for j in range(i + 1, n)]  # entry i + j*n, i<j
    assert all(v > 0 for v in vals)
    assert np.all(np.triu(out) == 0)  # nothing on or above the diagonal of the x-fastest image
    assert abs(np.mean(np.array(vals, np.float64)) - mean) < 1e-9 * mean


def test_index_list_equals_all_pairs(small_scene):
    s = small_scene
    n = s["n"]
    _, out, _ = ol.ecc(s["Ps"], s["dtrs"], s["n_u"], s["n_v"])
    idx = np.array([(i, j, i, j) for i in range(n) for j in range(i + 1, n)], np.int32)
    mean2, out2, _ = ol.ecc(s["Ps"], s["dtrs"], s["n_u"], s["n_v"], idx4=idx)
    ref = np.array([out[j, i] for i, j, _, _ in idx])
    assert np.array_equal(out2, ref)
    # same view twice: the reference zeroes the record -> no samples, value 0
    _, out3, ks3 = ol.ecc(s["Ps"], s["dtrs"], s["n_u"], s["n_v"], idx4=np.array([[1, 1, 1, 1]], np.int32),
                          want_ksamples=True)
    assert out3[0] == 0 and ks3[0] == 0


def test_fixed_dkappa_sample_counts(small_scene):
    s = small_scene
    dk = np.deg2rad(0.05)
    _, _, ks = ol.ecc(s["Ps"], s["dtrs"], s["n_u"], s["n_v"], dkappa=dk, want_ksamples=True)
    r = ol.object_radius(s["Ps"][0], s["n_u"], s["n_v"])
    assert ks.max() <= int(np.pi / 2 / dk) + 1
    assert ks.min() >= 1 and r > 0


def test_trajectory_geometry():
    Ps = ol.circular_trajectory(12, 750, 1200, 320, 256, 360, 1.0)
    for k, P in enumerate(Ps):
        Cc = ol.source_position(P).astype(np.float64)
        assert abs(np.linalg.norm(Cc[:3]) - 750) < 1e-3  # source on the circle of radius sid
        Pm = P.reshape(4, 3).T
        x = Pm @ np.array([0, 0, 0, 1.0])
        assert np.allclose(x[:2] / x[2], [160, 128], atol=1e-6)  # origin projects to the principal point
        assert abs(np.linalg.norm(Pm[2, :3]) - 1) < 1e-12  # normalised


def test_ramp_filter_restatement():
    """The numpy restatement of the reference's cuFFT ramp filter against the definition written out as a circular
    convolution (what the CUDA kernel does), even and odd n_t; constants have no response (H_0 = 0)."""
    rng = np.random.default_rng(2)
    for n_t in (16, 15, 48):
        x = rng.standard_normal((n_t, 5)).astype(np.float32)
        got = ol.ramp_filter(x)
        n_theta = n_t // 2 + 1
        scale = np.float32(-0.5) / np.float32(n_t * n_theta)
        g = np.zeros(n_t)
        for j in range(n_t):
            for k in range(1, n_theta):
                nyq = (n_t % 2 == 0) and (k == n_t // 2)
                g[j] += (1.0 if nyq else 2.0) * float(np.float32(k) * scale) * np.cos(2 * np.pi * ((j * k) % n_t) / n_t)
        want = np.stack([sum(x[m].astype(np.float64) * g[(t - m) % n_t] for m in range(n_t)) for t in range(n_t)])
        assert np.abs(got - want).max() < 1e-5 * np.abs(want).max()
        assert np.abs(ol.ramp_filter(np.ones((n_t, 3), np.float32))).max() < 1e-6


def test_correlation_variant_of_the_oracle():
    """use_corr: 1 - cc (un-centred, weights kappa_max/kappa): ~0 for consistent geometry, larger after a detector
    shift, within [0, 2]; SSD and correlation rank the two situations the same way."""
    n, n_u, n_v, n_a, n_t = 5, 96, 80, 96, 96
    ell = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5]])
    Ps = ol.circular_trajectory(n, 750, 1200, n_u, n_v, 200, 3.0)
    dtr = np.stack([ol.radon(ol.project_ellipsoids(P, n_u, n_v, ell), n_a, n_t) for P in Ps])
    m_ok, out_ok, _ = ol.ecc(Ps, dtr, n_u, n_v, use_corr=True)
    bad = Ps.copy()
    for c in range(4):
        bad[2, 0 + 3 * c] += 4.0 * bad[2, 2 + 3 * c]   # view 2: detector shift of 4 px in u
    m_bad, out_bad, _ = ol.ecc(bad, dtr, n_u, n_v, use_corr=True)
    vals = np.array([out_ok[j, i] for i in range(n) for j in range(i + 1, n)])
    assert (vals > -1e-6).all() and (vals < 2.0).all()
    assert 0 <= m_ok < 0.05 and m_bad > 1.2 * m_ok
    s_ok, _, _ = ol.ecc(Ps, dtr, n_u, n_v)
    s_bad, _, _ = ol.ecc(bad, dtr, n_u, n_v)
    assert s_bad > s_ok


def test_preprocess_restatement_invariants():
    """Oracle of the pre-processing step: default parameters clear exactly the first two columns / rows and the last
    column / row (a border of "zero = 1" runs b = 0..zero inclusive at the left / top, Gui/PreProccess.cpp:91-108); the
    low-pass keeps constants up to the missing last tap; the intrinsics of a synthetic matrix are the ones it was built
    from; cosine weights are 1 at the principal point and < 1 elsewhere."""
    img = np.full((40, 50), 7.0, np.float32)
    out = ol.preprocess(img, sigma=0.0)
    assert (out[:, :1] == 0).all() and (out[:1, :] == 0).all() and (out[:, -1] == 0).all() and (out[-1, :] == 0).all()
    assert (out[2:-1, 2:-1] == 7.0).all()
    k, sigma = 5, 1.84
    kern = np.exp(-0.5 * (np.arange(-k, k + 1) / sigma) ** 2)
    kern /= kern.sum()
    smooth = ol.preprocess(img, sigma=sigma, k=k, zero=(0, 0, 0, 0))
    assert abs(smooth[20, 25] - 7.0 * kern[:-1].sum() ** 2) < 1e-5
    n_u, n_v = 320, 240
    P = ol.circular_trajectory(4, 750, 1200, n_u, n_v, 200, 1.5)[1]
    fu, u0, v0 = ol.camera_intrinsics(P)
    f_traj = n_v / (2.0 * np.tan(0.5 * np.arctan(n_v * 1.5 / 1200)))  # cameraPerspective as makeCircularTrajectory calls it
    assert abs(u0 - 0.5 * n_u) < 1e-6 and abs(v0 - 0.5 * n_v) < 1e-6 and abs(fu - f_traj) < 1e-9 * fu
    w = ol.preprocess(np.ones((n_v, n_u), np.float32), sigma=0.0, zero=(0, 0, 0, 0), P=P)
    assert abs(w[n_v // 2, n_u // 2] - 1.0) < 1e-6 and w[0, 0] < w[n_v // 2, n_u // 2]


# ---- pre-processing (row N3) pinned by the reference's own headers ----------------------------------------------------------
PRE_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_preprocess_vectors.npz")


def test_preprocess_oracle_pinned_by_reference_headers():
    """tests/golden/ref_preprocess_vectors.npz holds outputs of the reference's NRRD::gaussianKernel / lowpass2D and of its
    weighting() inside the border loops of PreProccess::process (generator: tests/golden/make_ref_preprocess_vectors.py).
    The numpy restatement in oracle_lib reproduces them BIT FOR BIT: same taps (-k .. k-1: the last tap is missing in the
    reference), same clamping, fp64 sums stored as float, fp32 feathering weights."""
    g = np.load(PRE_GOLDEN)
    for q, (sigma, k) in enumerate(g["kernel_cases"]):
        k = int(k)
        kern = np.exp(-0.5 * (np.arange(-k, k + 1) / sigma) ** 2)
        kern /= kern.sum()
        assert np.allclose(kern, g[f"kernel_{q}"], rtol=1e-15, atol=0)
    for x, y in zip(g["weighting_x"], g["weighting_y"]):
        assert ol._feather(x) == y
    assert g["weighting_y"][0] == 0 and g["weighting_y"][50] == 1 and g["weighting_y"][-1] == 0
    img = g["lowpass_in"]
    for q, (sigma, k) in enumerate(g["lowpass_cases"]):
        got = ol.preprocess(img, zero=(0, 0, 0, 0), feather=(0, 0, 0, 0), sigma=float(sigma), k=int(k))
        want = g[f"lowpass_out_{q}"]
        # the border "zero = 0, feather = 0" still clears nothing; what is left is the low-pass alone
        assert np.array_equal(got, want), float(np.abs(got - want).max())
    bimg = g["border_in"]
    for q, (zero, feather) in enumerate(g["border_cases"]):
        got = ol.preprocess(bimg, zero=tuple(int(z) for z in zero), feather=tuple(int(f) for f in feather), sigma=0.0)
        assert np.array_equal(got, g[f"border_out_{q}"]), q


def test_preprocess_golden_matches_live_reference_build():
    """Where oracle/_ref is built (this container), the committed vectors are what the reference's headers give today."""
    R = ol.ref_host()
    if R is None or not hasattr(R, "ref_lowpass2d"):
        import pytest
        pytest.skip("oracle/_ref/libecc_ref_host.so (with the pre-processing entry points) not built")
    import ctypes as C
    g = np.load(PRE_GOLDEN)
    f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    R.ref_lowpass2d.argtypes = [f32p, C.c_int, C.c_int, C.c_double, C.c_int]
    work = g["lowpass_in"].copy()
    R.ref_lowpass2d(work, work.shape[1], work.shape[0], 1.84, 5)
    assert np.array_equal(work, g["lowpass_out_0"])


# ---- direct metric (row N4: MetricDirect / FBCC) ---------------------------------------------------------------------------
FBCC_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_fbcc_vectors.npz")
DIRECT_ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4]])


def _centre(P):
    """Camera centre by numpy's SVD (what the reference does with Eigen's: Geometry::getCameraCenter)."""
    c = np.linalg.svd(P.reshape(4, 3).T)[2][-1]
    return c / c[3]


def test_direct_fbcc_weight_pinned_by_reference_header():
    """tests/golden/ref_fbcc_vectors.npz: the reference's own LinePerspectivity and the per-sample fan-beam weight of its
    kernel (host build of RectifiedFBCC.h).  The oracle follows the roundings of the reference's GPU build (fused where nvcc
    fuses), the golden values are the host build's: they agree to fp32 rounding."""
    g = np.load(FBCC_GOLDEN)
    O = ol.oracle()
    worst, checked = 0.0, 0
    for rec, ts, out in zip(g["recs"], g["ts"], g["out"]):
        rec = np.ascontiguousarray(rec)
        for t, (tr, inv, der, w) in zip(ts, out):
            a, b, c, d = (np.float32(v) for v in rec[:4])
            t = np.float32(t)
            assert (a * t + b) / (c * t + d) == tr  # transform, plain fp32
            if abs(c * t + d) < 0.05 * (abs(c * t) + abs(d)):
                continue  # a position next to the pole of the perspectivity: every fp32 evaluation is noise there
            got = O.oracle_direct_fbcc_weight(rec, float(t))
            worst = max(worst, abs(got - w) / abs(w))
            checked += 1
    assert checked > 2500 and worst < 2e-5, (checked, worst)  # up to 20 x conditioning x a few fp32 roundings


def test_direct_fbcc_golden_matches_live_reference_build():
    R = ol.ref_host()
    if R is None or not hasattr(R, "ref_fbcc_weight"):
        import pytest
        pytest.skip("oracle/_ref/libecc_ref_host.so (with the fan-beam entry points) not built")
    g = np.load(FBCC_GOLDEN)
    rec = np.ascontiguousarray(g["recs"][7])
    for t, row in zip(g["ts"][7], g["out"][7]):
        assert R.ref_fbcc_weight(rec, float(t)) == row[3] and R.ref_fbcc_derivative(rec, float(t)) == row[2]


def test_direct_geometry_projective_properties():
    """The host geometry of the direct metric is restated without Eigen (parity unpinned); these properties hold for ANY
    correct implementation: every epipolar line passes through the epipole, the two lines of a plane are projections of
    that plane (points of it project onto both), kappa = 0 is the plane through the origin, and the planes span the object."""
    n_u, n_v = 320, 256
    Ps = ol.circular_trajectory(12, 750.0, 1200.0, n_u, n_v, 200.0, 1.2)
    for (i, j) in [(0, 4), (3, 10), (6, 7)]:
        P0, P1 = Ps[i].reshape(4, 3).T, Ps[j].reshape(4, 3).T
        C0, C1 = _centre(Ps[i]), _centre(Ps[j])
        g = ol.direct_pair_geometry(Ps[i], Ps[j], n_u, n_v)
        m = len(g["kappas"])
        assert m == len(g["lines0"]) and abs(m - 2 * np.hypot(n_u, n_v)) <= 1  # as many planes as twice the diagonal
        assert np.allclose(np.hypot(g["lines0"][:, 0], g["lines0"][:, 1]), 1, atol=1e-6)
        e0, e1 = P0 @ C1, P1 @ C0
        e0, e1 = e0 / e0[2], e1 / e1[2]
        scale0, scale1 = np.abs(e0).max(), np.abs(e1).max()
        assert np.abs(g["lines0"].astype(np.float64) @ e0).max() < 1e-5 * scale0
        assert np.abs(g["lines1"].astype(np.float64) @ e1).max() < 1e-5 * scale1
        # a third point of the plane: back-project a point of line 0, project it into view 1
        for q in (0, m // 3, m // 2, m - 1):
            l0, l1 = g["lines0"][q].astype(np.float64), g["lines1"][q].astype(np.float64)
            x0 = np.array([-l0[2] * l0[0] + 40 * l0[1], -l0[2] * l0[1] - 40 * l0[0], 1.0])  # on l0
            X = np.linalg.pinv(P0) @ x0 + 0.3 * C0
            x1 = P1 @ X
            assert abs(l1 @ (x1 / x1[2])) < 2e-3  # fp32 line coefficients, epipole ~1e3 px away
        # kappa = 0: the plane contains the origin -> its lines pass through the origin's projections
        q0 = int(np.argmin(np.abs(g["kappas"])))
        o0 = P0 @ np.array([0, 0, 0, 1.0])
        assert abs(g["lines0"][q0].astype(np.float64) @ (o0 / o0[2])) < 1.5 * g["dkappa"] * 1200 / 1.2
        # symmetric range, first angle = -kappa_max, step as computeForImagePair derives it
        assert g["kappas"][0] == np.float32(-0.5 * g["dkappa"] * 2 * np.hypot(n_u, n_v)) or abs(g["kappas"][0] + g["kappas"][-1]) < 2 * g["dkappa"]
        # fan-beam records: the perspectivity is orientation preserving (a d - b c >= 0 after the sign fix), distances positive
        f = g["fbcc0"].astype(np.float64)
        assert (f[:, 0] * f[:, 3] - f[:, 1] * f[:, 2] >= 0).all() and (f[:, 5] > 0).all()
        # squared source-to-line distance in pixels: between (source-detector distance)^2 and that plus the image diagonal^2
        assert (f[:, 5] > (0.5 * 750) ** 2).all()


def test_direct_line_integral_invariants():
    n_u, n_v = 96, 80
    rng = np.random.default_rng(5)
    ang = rng.uniform(-np.pi, np.pi, 200)
    pts = np.stack([rng.uniform(10, n_u - 10, 200), rng.uniform(10, n_v - 10, 200)], 1)
    lines = np.stack([np.cos(ang), np.sin(ang), -(np.cos(ang) * pts[:, 0] + np.sin(ang) * pts[:, 1])], 1).astype(np.float32)
    const = np.full((n_v, n_u), 3.0, np.float32)
    for interp in (ol.INTERP_EXACT, ol.INTERP_TEX8):
        # the derivative of a constant image vanishes (both parallel lines see the same samples)
        d = ol.direct_line_integrals(const, lines, interp=interp)
        assert np.abs(d).max() < 1e-3
        # a ramp along u: the derivative across the line is length * l0 (1 px between the two lines)
        ramp = np.tile(np.arange(n_u, dtype=np.float32), (n_v, 1))
        d = ol.direct_line_integrals(ramp, lines, interp=interp, shape=0)
        o = -lines[:, 2:3] * lines[:, :2]
        dvec = np.stack([lines[:, 1], -lines[:, 0]], 1)
        with np.errstate(divide="ignore"):
            ts = np.stack([(1 - o[:, 0]) / dvec[:, 0], (n_u - 1 - o[:, 0]) / dvec[:, 0], (1 - o[:, 1]) / dvec[:, 1], (n_v - 1 - o[:, 1]) / dvec[:, 1]], 1)
        ts.sort(axis=1)
        n_samples = np.floor((ts[:, 2] - ts[:, 1]) / 0.4) + 1
        # (up to one clamped sample at either end: the outer line may leave the last pixel centre by half a pixel)
        assert np.abs(d - lines[:, 0] * n_samples * 0.4).max() < 0.5
    # source loop and executed shape differ by at most the last sample; a line outside the image gives zero
    img = rng.standard_normal((n_v, n_u)).astype(np.float32)
    a = ol.direct_line_integrals(img, lines, shape=0)
    b = ol.direct_line_integrals(img, lines, shape=1)
    assert (a != b).mean() < 0.05
    outside = np.array([[1.0, 0.0, 50.0], [0.0, 1.0, -500.0]], np.float32)
    assert (ol.direct_line_integrals(img, outside) == 0).all()
    # the reference's launcher clips against n_u x n_u: on a wide image only lines through the bottom rows can change
    wide = rng.standard_normal((40, 96)).astype(np.float32)
    l2 = np.array([[0.0, 1.0, -20.0], [1.0, 0.0, -48.0]], np.float32)  # horizontal at v = 20, vertical at u = 48
    r_img = ol.direct_line_integrals(wide, l2, n_v_clip=40)
    r_ref = ol.direct_line_integrals(wide, l2, n_v_clip=96)
    assert r_img[0] == r_ref[0] and r_img[1] != r_ref[1]


def test_direct_pair_consistency_on_the_cpu():
    """Small scene on the CPU: true matrices score lower than a shifted detector; the pair value is the sum its signals give."""
    n_u, n_v = 96, 80
    Ps = ol.circular_trajectory(6, 750.0, 1200.0, n_u, n_v, 200.0, 3.0)
    imgs = [ol.project_ellipsoids(P, n_u, n_v, DIRECT_ELL) for P in Ps]
    for fbcc in (False, True):
        r = ol.direct_pair(Ps[0], Ps[2], imgs[0], imgs[2], fbcc=fbcc)
        d = r["samples0"] - r["samples1"]
        dk = ol.direct_pair_geometry(Ps[0], Ps[2], n_u, n_v)["dkappa"]
        assert r["value"] == np.sum((d * d).astype(np.float64) * dk) or abs(r["value"] - np.sum((d * d).astype(np.float64)) * dk) < 1e-12 * r["value"]
        H = np.array([[1, 0, 2.0], [0, 1, 2.0], [0, 0, 1]])
        Pm = (H @ Ps[2].reshape(4, 3).T).T.reshape(12)
        moved = ol.direct_pair(Ps[0], Pm, imgs[0], imgs[2], fbcc=fbcc)
        assert moved["value"] > 1.5 * r["value"], (fbcc, r["value"], moved["value"])
    total = ol.direct_evaluate(Ps[:3], np.stack(imgs[:3]))
    radius = ol.object_radius(Ps[0], n_u, n_v)
    pairs = [ol.direct_pair(Ps[i], Ps[j], imgs[i], imgs[j], radius=radius)["value"] for i in range(3) for j in range(i + 1, 3)]
    assert total == sum(pairs)
