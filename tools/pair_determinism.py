"""Development aid: C3 on one GPU, texture engines -- are the pair values the same bits (a) run to run, (b) evaluated in
ranges as 8 ranks would, (c) with the dtrs in another allocation?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
import bench
W = bench.WORKLOADS["c3"]
n, n_u, n_v, n_a, n_t = W["n"], W["n_u"], W["n_v"], W["n_alpha"], W["n_t"]
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, bench.GEO["sid"], bench.GEO["sdd"], n_u, n_v, W["arc"], W["px"])
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, bench.ELLIPSOIDS, imgs)
dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_TEXTURE)
dtrs2 = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_TEXTURE)
print("radon run to run equal:", bool(torch.equal(dtrs, dtrs2)))
del imgs
ctx.set_interpolation(api.INTERP_TEXTURE)
ctx.set_object_radius(0.0)
ctx.set_epipolar_plane_step(float(np.deg2rad(W["dkappa_deg"])))
ctx.set_projection_matrices(Ps)
ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
total = n * (n - 1) // 2
def run_all():
    c = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    m = ctx.evaluate(c)
    return m, c
m1, c1 = run_all()
m2, c2 = run_all()
print("all pairs run to run: mean", repr(m1), repr(m2), "cost equal:", bool(torch.equal(c1, c2)), "differing:", int((c1 != c2).sum()))
bounds = ctx.partition_pairs(8)
c3 = torch.zeros((n, n), dtype=torch.float32, device="cuda")
s = sum(ctx.evaluate_range(int(a), int(b), c3) for a, b in zip(bounds[:-1], bounds[1:]))
print("ranges: mean", repr(s / total), "cost equal to whole:", bool(torch.equal(c1, c3)), "differing:", int((c1 != c3).sum()))
if int((c1 != c3).sum()):
    d = (c1 - c3).abs()
    idx = torch.nonzero(d > 0)[:5].tolist()
    for j, i in idx:
        print("   pair", i, j, float(c1[j, i]), float(c3[j, i]))
ctx.set_radon_intermediates(dtrs2, n_u, n_v, True)
m4, c4 = run_all()
print("other allocation: mean", repr(m4), "cost equal:", bool(torch.equal(c1, c4)), "differing:", int((c1 != c4).sum()))
