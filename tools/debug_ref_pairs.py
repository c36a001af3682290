"""Development aid: where do our per-pair values and the reference CUDA kernel's differ, and who is right?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from epipolarconsistency_b200 import api
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
n, n_u, n_v, n_a, n_t = 10, 160, 128, 192, 192
Ps = ol.circular_trajectory(n, 750, 1200, n_u, n_v, 200, 2.0)
imgs = np.stack([ol.project_ellipsoids(P, n_u, n_v, ELL) for P in Ps])
ctx = api.Context()
dtr = ctx.radon_compute(imgs, n_a, n_t)
dtr_o = np.stack([ol.radon(im, n_a, n_t, interp=ol.INTERP_TEX8) for im in imgs])
print("radon: max|gpu texture - oracle tex8|/peak", np.abs(dtr - dtr_o).max() / np.abs(dtr_o).max())
ctx.set_projection_matrices(Ps); ctx.set_radon_intermediates(dtr, n_u, n_v, True)
cost = np.zeros((n, n), np.float32); ctx.evaluate(cost)
R = ol.RefCudaMetric(Ps, dtr, n_u, n_v)
runs = np.stack([R.evaluate(ctx.get_object_radius(), 0.0)[1] for _ in range(8)])
_, orc, ks = ol.ecc(Ps, dtr, n_u, n_v, interp=ol.INTERP_TEX8, fast_sincos=True, want_ksamples=True)
k = 0
for i in range(n):
    for j in range(i + 1, n):
        mine, rmin, rmax, o = cost[j, i], runs[:, j, i].min(), runs[:, j, i].max(), orc[j, i]
        flag = "" if abs(mine - rmax) / rmax < 1e-4 else "  <-- differs"
        print(f"pair ({i},{j}) samples {ks[k]}: ours {mine:.4f} ref[min {rmin:.4f} max {rmax:.4f}] oracle-tex8 {o:.4f} ours/oracle-1 {mine/o-1:+.2e} refmax/oracle-1 {rmax/o-1:+.2e}{flag}")
        k += 1
