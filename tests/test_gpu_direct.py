"""GPU parity tests of the direct metric (MetricDirect / computeForImagePair, EpipolarConsistencyDirect.{h,cpp,cu},
RectifiedFBCC.h), called through the C ABI.

Pins: the line kernel against the reference's OWN kernel_computeLineIntegrals (compiled unchanged into
oracle/_ref/libecc_ref_cuda.so) on the same lines -- bit for bit; the whole pair against the CPU oracle (fp64 geometry
restated without Eigen: that part is "parity unpinned", so it is additionally checked by projective properties that do not
depend on any restatement)."""
import numpy as np
import pytest

import oracle_lib as ol
from epipolarconsistency_b200 import api

pytestmark = pytest.mark.gpu

ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4],
                [5, 30, 20, 12, 18, 9, 0.8]])
# oracle TEX8 mode = our model of the texture unit (3 % of positions differ by one 1/256 weight step from the hardware's
# own coordinate rounding): signals agree to this fraction of their peak
ORACLE_TEX_SIGNAL_TOL = 2e-3
PAIR_TOL_REF = 1e-3


@pytest.fixture(scope="module")
def ctx():
    c = api.Context()
    yield c
    c.close()


def make_scene(n, n_u, n_v, px, noise=0.0, seed=3, arc=200):
    Ps = ol.circular_trajectory(n, 750, 1200, n_u, n_v, arc, px)
    imgs = np.stack([ol.project_ellipsoids(P, n_u, n_v, ELL) for P in Ps]).astype(np.float32)
    if noise:
        rng = np.random.default_rng(seed)
        imgs = imgs + noise * rng.standard_normal(imgs.shape).astype(np.float32)
    return Ps, np.ascontiguousarray(imgs)


@pytest.fixture(scope="module")
def scene():
    Ps, imgs = make_scene(6, 160, 128, 2.0)
    return dict(n=6, n_u=160, n_v=128, Ps=Ps, imgs=imgs)


def setup(ctx, Ps, imgs, fbcc=False, clip=False, radius=0.0, dkappa=0.0):
    ctx.set_projection_matrices(Ps)
    ctx.direct_set_images(imgs)
    ctx.direct_set_fan_beam(fbcc)
    ctx.direct_set_reference_clip(clip)
    ctx.set_object_radius(radius)
    ctx.set_epipolar_plane_step(dkappa)


# ---------------------------------------------------------------------------------------------------
# the line kernel against the reference's own kernel
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(256, 256), (200, 160), (160, 200)])
@pytest.mark.parametrize("fbcc", [False, True])
def test_line_integrals_bit_identical_to_reference_kernel(ctx, shape, fbcc):
    if ol.ref_cuda() is None or not hasattr(ol.ref_cuda(), "ref_cuda_direct_line_integrals"):
        pytest.skip("oracle/_ref/libecc_ref_cuda.so without the direct kernel")
    n_u, n_v = shape
    Ps, imgs = make_scene(4, n_u, n_v, 1.5, noise=0.05)  # rough images: one sample more or less shows
    setup(ctx, Ps, imgs, clip=True)  # the reference's launcher clips against n_u x n_u
    g = ctx.direct_pair_geometry(0, 2)
    rng = np.random.default_rng(11)
    # the pair's own epipolar lines plus random lines through the image (all angles, vertical and horizontal included)
    m = 3000
    ang = rng.uniform(-np.pi, np.pi, m)
    ang[:4] = [0.0, np.pi / 2, -np.pi / 2, np.pi]
    pts = np.stack([rng.uniform(0, n_u, m), rng.uniform(0, n_v, m)], 1)
    rnd = np.stack([np.cos(ang), np.sin(ang), -(np.cos(ang) * pts[:, 0] + np.sin(ang) * pts[:, 1])], 1).astype(np.float32)
    rnd[0] = [1.0, 0.0, -n_u / 3.0]
    rnd[1] = [0.0, 1.0, -n_v / 3.0]
    lines = np.concatenate([g["lines0"], rnd])
    rec = None
    if fbcc:
        rec = np.concatenate([g["fbcc0"], g["fbcc0"][rng.integers(0, len(g["fbcc0"]), m)]])
    for image in (0, 2):
        want, _ = ol.ref_cuda_direct_line_integrals(imgs[image], lines, rec)
        got = ctx.direct_line_integrals(image, lines, rec)
        assert np.isfinite(want).all()
        assert (want != 0).sum() > len(lines) // 2
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, (bad[:8], got[bad[:8]], want[bad[:8]])


def test_line_integrals_plain_clip_differs_only_where_the_box_differs(ctx):
    """reference_clip off: lines are clipped against the image, not against n_u x n_u."""
    n_u, n_v = 200, 120
    Ps, imgs = make_scene(2, n_u, n_v, 1.5)
    setup(ctx, Ps, imgs, clip=False)
    lines = ctx.direct_pair_geometry(0, 1)["lines0"]
    got = ctx.direct_line_integrals(0, lines)
    want = ol.direct_line_integrals(imgs[0], lines, interp=ol.INTERP_TEX8, shape=1)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= ORACLE_TEX_SIGNAL_TOL * scale


# ---------------------------------------------------------------------------------------------------
# pairs against the oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fbcc", [False, True])
@pytest.mark.parametrize("dkappa", [0.0, 2e-3])
def test_pair_vs_oracle(ctx, scene, fbcc, dkappa):
    setup(ctx, scene["Ps"], scene["imgs"], fbcc=fbcc, dkappa=dkappa)
    radius = ol.object_radius(scene["Ps"][0], scene["n_u"], scene["n_v"])
    for (i, j) in [(0, 1), (0, 3), (2, 5)]:
        got = ctx.direct_evaluate_pair(i, j)
        want = ol.direct_pair(scene["Ps"][i], scene["Ps"][j], scene["imgs"][i], scene["imgs"][j], radius=radius, dkappa=dkappa, fbcc=fbcc)
        assert len(got["kappas"]) == len(want["kappas"]) and len(got["kappas"]) > 100
        assert np.array_equal(got["kappas"], want["kappas"])
        for k in ("samples0", "samples1"):
            scale = np.abs(want[k]).max()
            assert scale > 0
            assert np.abs(got[k] - want[k]).max() <= ORACLE_TEX_SIGNAL_TOL * scale, (i, j, k)
        # the value is the fp64 sum of the squared fp32 differences times dkappa
        dk = dkappa if dkappa > 0 else ctx.direct_pair_geometry(i, j)["dkappa"]
        d = got["samples0"] - got["samples1"]
        assert got["value"] == pytest.approx(float(np.sum((d * d).astype(np.float64)) * dk), rel=1e-12)


def test_pair_geometry_device_equals_host_and_oracle(ctx, scene):
    """The kernel's own fp64 geometry (kappas reported by evaluate_pair) is the host twin's and the oracle's."""
    setup(ctx, scene["Ps"], scene["imgs"])
    radius = ol.object_radius(scene["Ps"][0], scene["n_u"], scene["n_v"])
    g = ctx.direct_pair_geometry(1, 4)
    o = ol.direct_pair_geometry(scene["Ps"][1], scene["Ps"][4], scene["n_u"], scene["n_v"], radius=radius)
    r = ctx.direct_evaluate_pair(1, 4)
    assert np.array_equal(g["kappas"], o["kappas"]) and np.array_equal(g["kappas"], r["kappas"])
    assert g["dkappa"] == pytest.approx(o["dkappa"], rel=1e-14)
    for k in ("lines0", "lines1"):
        assert np.abs(g[k] - o[k]).max() <= 2e-6 * np.abs(o[k]).max()
    for k in ("fbcc0", "fbcc1"):
        assert np.abs(g[k] - o[k]).max() <= 1e-5 * np.abs(o[k]).max()


def test_given_kappas_equal_the_automatic_ones(ctx, scene):
    setup(ctx, scene["Ps"], scene["imgs"])
    auto = ctx.direct_evaluate_pair(0, 2)
    given = ctx.direct_evaluate_pair(0, 2, kappas=auto["kappas"])
    assert np.array_equal(auto["samples0"], given["samples0"]) and np.array_equal(auto["samples1"], given["samples1"])
    assert auto["value"] == given["value"]
    # a subset of the planes: the matching subset of the signals
    sub = ctx.direct_evaluate_pair(0, 2, kappas=auto["kappas"][10:75])
    assert np.array_equal(sub["samples0"], auto["samples0"][10:75])


# ---------------------------------------------------------------------------------------------------
# all pairs
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fbcc", [False, True])
def test_evaluate_all_pairs_equals_pairwise_and_oracle(ctx, scene, fbcc):
    import torch
    n = scene["n"]
    if fbcc:
        # the rectifying perspectivity has its pole at the epipole: views more than ~150 degrees apart (epipole inside the
        # image) have no finite fan-beam value, in the reference as here (test_fbcc_pole_inside_the_image_is_not_finite)
        Ps, imgs = make_scene(n, scene["n_u"], scene["n_v"], 2.0, arc=120)
        scene = dict(scene, Ps=Ps, imgs=imgs)
    setup(ctx, scene["Ps"], scene["imgs"], fbcc=fbcc)
    cost = np.full((n, n), -7.0, np.float32)
    total = ctx.direct_evaluate(cost)
    vals = {}
    for i in range(n):
        for j in range(i + 1, n):
            vals[(i, j)] = ctx.direct_evaluate_pair(i, j)["value"]
            assert cost[j, i] == np.float32(vals[(i, j)])  # entry i + j n
    # untouched entries keep the caller's values (diagonal and the other triangle)
    assert all(cost[i, j] == -7.0 for i in range(n) for j in range(i, n))
    assert total == pytest.approx(sum(vals.values()), rel=1e-13)
    # device cost image, device images: the same bits
    ctx.direct_set_images(torch.from_numpy(scene["imgs"]).cuda())
    cost_d = torch.full((n, n), -7.0, device="cuda")
    total_d = ctx.direct_evaluate(cost_d)
    assert total_d == total and np.array_equal(cost_d.cpu().numpy(), cost)
    # run to run: no atomics
    assert ctx.direct_evaluate(None) == total
    # oracle (texture model)
    want_cost = np.zeros((n, n), np.float32)
    want = ol.direct_evaluate(scene["Ps"], scene["imgs"], fbcc=fbcc, cost_image=want_cost)
    got_pairs = np.array([cost[j, i] for i in range(n) for j in range(i + 1, n)], np.float64)
    want_pairs = np.array([want_cost[j, i] for i in range(n) for j in range(i + 1, n)], np.float64)
    rel = np.abs(got_pairs - want_pairs) / np.maximum(want_pairs, 1e-3 * want_pairs.max())
    print("direct all pairs vs oracle: worst pair", rel.max(), "sum", abs(total - want) / want)
    # measured on B200: worst pair 2.3e-4 (derivative) / 3.2e-4 (fan-beam), sum 2.7e-5 / 2.6e-7 -- what is left is the texture
    # model of the oracle (the hardware's own coordinate rounding), not the kernel: against the reference's kernel the
    # integrals are bit-identical (test_line_integrals_bit_identical_to_reference_kernel)
    assert rel.max() < 2e-3
    assert abs(total - want) / want < 2e-4


def test_pair_ranges_compose_to_the_whole_and_partition_balances_the_planes(ctx, scene):
    """ecc_direct_evaluate_range / ecc_direct_partition (multi-GPU sharding of the direct metric): the ranges of any cut give
    the whole's pair values bit for bit and its sum to fp64 rounding; with a fixed plane step the pairs differ in work
    (kappa range) and the cut follows the planes, not the pair count."""
    n = scene["n"]
    pairs = n * (n - 1) // 2
    setup(ctx, scene["Ps"], scene["imgs"], dkappa=1.5e-3, radius=40.0)
    whole = np.zeros((n, n), np.float32)
    total = ctx.direct_evaluate(whole)
    planes = np.array([len(ctx.direct_pair_geometry(i, j)["kappas"]) for i in range(n) for j in range(i + 1, n)])
    assert planes.max() > 2 * planes.min()  # unequal work
    for parts in (1, 2, 3, 8, 40):
        b = ctx.direct_partition(parts)
        assert b[0] == 0 and b[-1] == pairs and np.all(np.diff(b) >= 0)
        cost = np.zeros((n, n), np.float32)
        sums = [ctx.direct_evaluate_range(int(b[k]), int(b[k + 1]), cost) for k in range(parts)]
        assert np.array_equal(cost, whole)
        assert abs(sum(sums) - total) <= 1e-13 * total
        if parts in (2, 3):
            work = np.array([np.ceil(planes[int(b[k]):int(b[k + 1])] / 32).sum() for k in range(parts)])
            assert work.max() - work.min() <= np.ceil(planes.max() / 32)  # within one pair of equal
    assert ctx.direct_evaluate_range(4, 4, None) == 0.0
    with pytest.raises(api.EccError):
        ctx.direct_evaluate_range(3, pairs + 1, None)


def test_fbcc_pole_inside_the_image_is_not_finite(ctx, scene):
    """Views 166 degrees apart: the baseline passes through the object, the epipole lies inside the image and with it the pole
    of the rectifying perspectivity -- weights overflow on the lines through it.  The reference's kernel has no guard
    (EpipolarConsistencyDirect.cu:86-93), neither has the oracle, neither have we: all three agree that the value is not
    finite, and the derivative variant of the same pair is."""
    setup(ctx, scene["Ps"], scene["imgs"], fbcc=True)
    radius = ol.object_radius(scene["Ps"][0], scene["n_u"], scene["n_v"])
    got = ctx.direct_evaluate_pair(0, 5)
    want = ol.direct_pair(scene["Ps"][0], scene["Ps"][5], scene["imgs"][0], scene["imgs"][5], radius=radius, fbcc=True)
    assert not np.isfinite(got["value"]) and not np.isfinite(want["value"])
    for k in ("samples0", "samples1"):  # which lines hit the pole exactly is a matter of the last bit
        bad_got, bad_want = (~np.isfinite(got[k])).sum(), (~np.isfinite(want[k])).sum()
        assert 0 < bad_got < 0.2 * len(got[k]) and 0 < bad_want < 0.2 * len(want[k])
    ctx.direct_set_fan_beam(False)
    assert np.isfinite(ctx.direct_evaluate_pair(0, 5)["value"])


def test_consistent_geometry_scores_lower_than_a_shifted_detector(ctx):
    """The domain property the metric exists for: with the true matrices the redundant signals agree; moving one view's
    detector by three pixels raises every pair that contains the view, and only those."""
    n, n_u, n_v = 5, 192, 160
    Ps, imgs = make_scene(n, n_u, n_v, 1.6, arc=130)
    for fbcc in (False, True):
        setup(ctx, Ps, imgs, fbcc=fbcc)
        cost0 = np.zeros((n, n), np.float32)
        ctx.direct_evaluate(cost0)
        H = np.array([[1, 0, 3.0], [0, 1, 3.0], [0, 0, 1]])
        Pm = Ps.copy()
        Pm[2] = (H @ Ps[2].reshape(4, 3).T).T.reshape(12)
        ctx.set_projection_matrices(Pm)
        cost1 = np.zeros((n, n), np.float32)
        ctx.direct_evaluate(cost1)
        for i in range(n):
            for j in range(i + 1, n):
                if 2 in (i, j):
                    assert cost1[j, i] > 1.5 * cost0[j, i], (fbcc, i, j, cost0[j, i], cost1[j, i])
                else:
                    assert cost1[j, i] == cost0[j, i]


def test_metric_direct_class_and_errors(ctx, scene):
    m = api.MetricDirect(scene["Ps"], scene["imgs"], ctx=ctx)
    m.setObjectRadius(0.0).setEpipolarPlaneStep(0.0).setFanBeamConsistency(False)
    assert m.getNumberOfProjetions() == scene["n"]
    total = m.evaluate()
    v, s0, s1, k = m.evaluateForImagePair(0, 1)
    assert total > v > 0 and len(s0) == len(s1) == len(k)
    assert m.getObjectRadius() == pytest.approx(ol.object_radius(scene["Ps"][0], scene["n_u"], scene["n_v"]), rel=1e-12)
    # fewer matrices than images
    ctx.set_projection_matrices(scene["Ps"][:3])
    with pytest.raises(api.EccError):
        ctx.direct_evaluate(None)
    with pytest.raises(api.EccError):
        ctx.direct_evaluate_pair(0, 1)
    ctx.set_projection_matrices(scene["Ps"])
    with pytest.raises(api.EccError):
        ctx.direct_evaluate_pair(0, scene["n"])
    # an absurd plane step (a billion planes per pair) is refused, not attempted
    ctx.set_epipolar_plane_step(1e-10)
    with pytest.raises(api.EccError):
        ctx.direct_evaluate(None)
    ctx.set_epipolar_plane_step(0.0)
    assert ctx.direct_evaluate(None) == total
    # an empty set and a single view
    ctx.direct_set_images(scene["imgs"][:1])
    ctx.set_projection_matrices(scene["Ps"][:1])
    assert ctx.direct_evaluate(None) == 0.0


def test_ragged_and_degenerate_sizes(ctx):
    """Tiny odd-sized images, plane counts that are no multiple of the 32 planes of a CTA, a single plane, no plane at all."""
    n_u, n_v = 37, 23
    Ps, imgs = make_scene(3, n_u, n_v, 9.0, arc=90)
    for fbcc in (False, True):
        for dkappa, expect in ((0.0, None), (0.05, None), (0.2, None), (10.0, 0)):
            setup(ctx, Ps, imgs, fbcc=fbcc, dkappa=dkappa)
            radius = ol.object_radius(Ps[0], n_u, n_v)
            cost = np.zeros((3, 3), np.float32)
            total = ctx.direct_evaluate(cost)
            want_cost = np.zeros((3, 3), np.float32)
            want = ol.direct_evaluate(Ps, imgs, dkappa=dkappa, fbcc=fbcc, cost_image=want_cost)
            for (i, j) in ((0, 1), (0, 2), (1, 2)):
                got = ctx.direct_evaluate_pair(i, j)
                ref = ol.direct_pair(Ps[i], Ps[j], imgs[i], imgs[j], radius=radius, dkappa=dkappa, fbcc=fbcc)
                assert len(got["kappas"]) == len(ref["kappas"])
                if expect is not None:
                    assert len(got["kappas"]) == expect and got["value"] == 0.0
                if len(ref["kappas"]):
                    scale = max(np.abs(ref["samples0"]).max(), 1e-6)
                    assert np.abs(got["samples0"] - ref["samples0"]).max() <= 5e-3 * scale
            if want > 0:
                assert abs(total - want) <= 2e-2 * want
            else:
                assert total == 0.0 and not cost.any()
