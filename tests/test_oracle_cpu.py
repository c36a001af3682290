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
