"""Development aid: are synthetic projections / Radon intermediates the same bits when produced in shards?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
sys.path.insert(0, ROOT)
import bench
W = bench.WORKLOADS["c3"]
n, n_u, n_v, n_a, n_t = 496, W["n_u"], W["n_v"], W["n_alpha"], W["n_t"]
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, W["sid"], W["sdd"], n_u, n_v, W["arc"], W["px"])
sel = list(range(120, 136))  # 16 views that straddle the 8-GPU shard boundary at 124
full = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, bench.ELLIPSOIDS, full)
a = full[sel].clone()
del full
b = torch.empty((len(sel), n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps[sel[0]:sel[4]], n_u, n_v, bench.ELLIPSOIDS, b[:4])
ctx.synth_projections(Ps[sel[4]:sel[-1] + 1], n_u, n_v, bench.ELLIPSOIDS, b[4:])
print("synthetic images equal:", bool(torch.equal(a, b)), "max diff", float((a - b).abs().max()))
da = ctx.radon_compute(a, n_a, n_t, interp=api.INTERP_TEXTURE)
db = torch.cat([ctx.radon_compute(a[:4], n_a, n_t, interp=api.INTERP_TEXTURE), ctx.radon_compute(a[4:], n_a, n_t, interp=api.INTERP_TEXTURE)])
print("radon (texture engine) equal in shards:", bool(torch.equal(da, db)))
