"""Where a C3 step's time goes outside the kernels (one GPU): every call of the step timed on the host with a device
synchronize behind it, next to the kernel families' event times of the same step (ecc_profile_*)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402

n = int(os.environ.get("N", 496))
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
full = torch.empty((n, n_t, n_a), dtype=torch.float32, device="cuda")
cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")
ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
total = n * (n - 1) // 2


def sync():
    torch.cuda.synchronize()


def step(k, timing=None):
    P = Ps.copy()
    P[0, 0] = np.nextafter(P[0, 0], np.inf if k % 2 else -np.inf)
    calls = [("radon_compute", lambda: ctx.radon_compute(imgs, n_a, n_t, out=full, interp=api.INTERP_HYBRID_STATIC)),
             ("set_radon_intermediates", lambda: ctx.set_radon_intermediates(full, n_u, n_v, True)),
             ("set_projection_matrices", lambda: ctx.set_projection_matrices(P)),
             ("cost.zero_", lambda: cost.zero_()),
             ("evaluate_range", lambda: ctx.evaluate_range(0, total, cost))]
    for name, f in calls:
        if timing is not None:
            sync()
            t0 = time.perf_counter()
        f()
        if timing is not None:
            t1 = time.perf_counter()  # the call returned
            sync()
            t2 = time.perf_counter()  # the device is idle
            timing.setdefault(name, []).append((t1 - t0, t2 - t0))


for k in range(3):
    step(k)
sync()
t0 = time.perf_counter()
K = 5
for k in range(K):
    step(k)
sync()
whole = (time.perf_counter() - t0) / K
timing = {}
ctx.profile_reset()
ctx.profile_enable(True)
for k in range(K):
    step(k, timing)
prof = {fam: ctx.profile_get(fam) for fam in ("radon", "pairs", "geometry", "reduce", "stage")}
ctx.profile_enable(False)
print(f"whole step, no syncs inside: {whole * 1e3:.3f} ms")
tot = 0.0
for name, v in timing.items():
    a = np.array(v) * 1e3
    print(f"  {name:26s} returns after {a[:, 0].mean():8.3f} ms, device idle after {a[:, 1].mean():8.3f} ms")
    tot += a[:, 1].mean()
print(f"  sum of the calls with a sync behind each: {tot:.3f} ms")
print("kernel families (events), per step:", {k: (round(v[0] / K, 3), v[1] // K) for k, v in prof.items()})
