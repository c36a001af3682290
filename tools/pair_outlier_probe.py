"""Development aid (GPU box): where do single pairs differ between our pair kernel and the reference's CUDA kernel?
Both work on the same K0/K1 maps (bit-identical) and the same intermediates, so a pair beyond 1e-3 must differ in single
kappa samples.  The kappa grid (m + 1/2) dkappa is fixed while the object radius only moves kappa_max = asin(r / K0[6]):
evaluating both implementations at a ladder of radii gives partial sums, and the first radius at which they part
brackets the offending sample; our own per-sample signals around it are then printed (ecc_pair_signals)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from epipolarconsistency_b200 import api  # noqa: E402

ELL = np.array([[0.0, 0.0, 0.0, 80.0, 60.0, 70.0, 1.0], [20.0, -10.0, 5.0, 25.0, 30.0, 20.0, 0.6], [-25.0, 15.0, -10.0, 20.0, 22.0, 28.0, -0.5],
                [5.0, 30.0, 20.0, 22.0, 20.0, 24.0, 0.8], [-10.0, -30.0, -25.0, 30.0, 21.0, 20.0, -0.7]])
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
sel = np.unique(np.array(list(range(200, 232)) + list(range(3, 496, 16))))
n = len(sel)
Ps = api.make_circular_trajectory(496, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)[sel]
ctx = api.Context()
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_TEXTURE)
dk = float(np.deg2rad(0.01))
ctx.set_interpolation(api.INTERP_TEXTURE)
ctx.set_object_radius(0.0)
ctx.set_epipolar_plane_step(dk)
ctx.set_projection_matrices(Ps)
ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
radius = ctx.get_object_radius()
cost = np.zeros((n, n), np.float32)
ctx.evaluate(cost)
ref = ol.RefCudaMetric(Ps, dtrs, n_u, n_v)
_, rcost, _ = ref.evaluate(radius, dk)
pv = lambda c: np.array([c[j, i] for i in range(n) for j in range(i + 1, n)], np.float64)
ij = [(i, j) for i in range(n) for j in range(i + 1, n)]
g, r = pv(cost), pv(rcost)
rel = np.abs(g - r) / np.maximum(np.abs(r), 1e-6 * np.abs(r).max())
order = np.argsort(-rel)
print(f"{n} views, {len(g)} pairs, radius {radius:.3f}; worst pairs:")
for k in order[:8]:
    i, j = ij[k]
    print(f"  pair {k} = views {sel[i]},{sel[j]} (separation {sel[j] - sel[i]}): ours {g[k]:.9g} ref {r[k]:.9g} rel {rel[k]:.3g}")
K = ctx.pair_maps(n_views=n)
for k in order[:2]:
    i, j = ij[k]
    idx = np.array([(i, j, i, j)], np.int32)
    base = float(K[k, 6])
    kmax_full = float(K[k, 15])
    print(f"--- pair {k} views {sel[i]},{sel[j]}: baseline distance {base:.3f}, kappa_max {kmax_full:.6f} = {kmax_full / dk:.1f} samples")
    # ladder of radii -> partial sums over the first M samples
    M_full = int(np.ceil((kmax_full - 0.5 * dk) / dk))
    lo, hi = 0, M_full

    def partial(M):
        rr = base * np.sin(min((M + 0.25) * dk, np.pi / 2 - 1e-6)) if M < M_full else radius
        ctx.set_object_radius(float(rr))
        a = np.zeros(1, np.float32)
        ctx.evaluate_indices(idx, a)
        b = ref.evaluate(float(np.float32(rr)), dk, idx)[1]
        cnt = int(np.ceil((np.arcsin(min(rr / base, 1.0)) - 0.5 * dk) / dk))
        return float(a[0]), float(b[0]), cnt

    a, b, cnt = partial(M_full)
    print(f"    full: ours {a:.9g} ref {b:.9g} samples ~{cnt}")
    while hi - lo > 1:
        mid = (lo + hi) // 2
        a, b, cnt = partial(mid)
        differ = abs(a - b) > 2e-5 * max(abs(b), 1e-12)
        print(f"    first {mid:5d} samples (count {cnt}): ours {a:.9g} ref {b:.9g} rel {abs(a - b) / max(abs(b), 1e-30):.3g} {'DIFFER' if differ else 'agree'}")
        if differ:
            hi = mid
        else:
            lo = mid
    ctx.set_object_radius(0.0)
    sig = ctx.pair_signals(i, j)
    count = len(sig["kappas"]) // 2
    print(f"    first differing sample index ~{hi - 1} of {count}; our signals around it (+kappa and -kappa halves):")
    for m in range(max(0, hi - 4), min(count, hi + 3)):
        qp, qm = count + m, count - 1 - m
        print(f"      m={m} kappa={sig['kappas'][qp]:.6f}: +k s0={sig['signal0'][qp]:.6g} s1={sig['signal1'][qp]:.6g} line0=({sig['lines0'][qp][0]:.6g},{sig['lines0'][qp][1]:.6g}) line1=({sig['lines1'][qp][0]:.6g},{sig['lines1'][qp][1]:.6g})"
              f" | -k s0={sig['signal0'][qm]:.6g} s1={sig['signal1'][qm]:.6g} line0=({sig['lines0'][qm][0]:.6g},{sig['lines0'][qm][1]:.6g}) line1=({sig['lines1'][qm][0]:.6g},{sig['lines1'][qm][1]:.6g})")
# ---- which side of the seam makes the difference?  (1) OUR kernel through the reference launcher's signature on the
# reference's own CUDA-array textures and buffers; (2) the REFERENCE kernel on pitch2D (linear-memory) textures
import ctypes as C
from epipolarconsistency_b200 import _lib
ours_launcher = C.cast(getattr(_lib.load(), "_Z19epipolarConsistencyiiiPciiffiPfS0_iPiS0_S0_ffbbS0_"), C.c_void_p).value
ref.set_launcher(ours_launcher)
_, xcost, _ = ref.evaluate(radius, dk)
ref.set_launcher(None)
ref2 = ol.RefCudaMetric(Ps, dtrs, n_u, n_v, pitch2d=True)
_, pcost, _ = ref2.evaluate(radius, dk)
ref2.close()
x, p2 = pv(xcost), pv(pcost)
print("ours (pitch2D textures) | ours on the reference's array textures | reference (array textures) | reference on pitch2D textures")
for k in order[:8]:
    print(f"  pair {k}: {g[k]:.9g} | {x[k]:.9g} | {r[k]:.9g} | {p2[k]:.9g}")
print("max rel: ours-pitch2D vs ours-array %.3g; ref-array vs ref-pitch2D %.3g; ours-array vs ref-array %.3g; ours-pitch2D vs ref-pitch2D %.3g" % (
    np.max(np.abs(g - x) / np.maximum(np.abs(g), 1e-9)), np.max(np.abs(r - p2) / np.maximum(np.abs(r), 1e-9)),
    np.max(np.abs(x - r) / np.maximum(np.abs(r), 1e-9)), np.max(np.abs(g - p2) / np.maximum(np.abs(p2), 1e-9))))
ref.close()
