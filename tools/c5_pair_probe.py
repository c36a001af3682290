"""The pair kernel at the C5 launch shape (1 live view vs 400 reference views, 400 listed pairs, CTA per pair with splits) on
the plain path, for `ncu -k regex:pairs_kernel` (GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402

n, n_u, n_v, n_a, n_t = 401, 1240, 960, 768, 768
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_HYBRID_STATIC)
del imgs
ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
if os.environ.get("RADIUS"):  # development: a tiny object radius leaves a handful of samples per pair -> the kernel's fixed costs
    ctx.set_object_radius(float(os.environ["RADIUS"]))
ctx.set_projection_matrices(Ps)
ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
live = n - 1
idx = np.array([(live, i, live, i) for i in range(live)], np.int32)
if os.environ.get("SKIP_LONG"):  # development: without the pairs whose baseline passes through the object (kappa_max = pi / 2)
    c0 = ctx.pair_sample_counts(n)
    keep = [i for i in range(live) if c0[i * n - i * (i + 1) // 2 + (live - i - 1)] < int(os.environ["SKIP_LONG"])]
    idx = np.array([(live, i, live, i) for i in keep], np.int32)
out = np.zeros(len(idx), np.float32)
for k in range(3):  # warm-up (first launch, staging)
    ctx.evaluate_indices(idx, out)
ctx.profile_enable(True)
for k in range(int(os.environ.get("CALLS", 6))):
    P = Ps[live].copy()
    P[9] += 1e-3 * k
    ctx.update_projection_matrix(live, P)
    m = ctx.evaluate_indices(idx, out)
ms, cnt = ctx.profile_get("pairs")
counts = ctx.pair_sample_counts(n)
listed = np.array([counts[i * n - i * (i + 1) // 2 + (live - i - 1)] for i in idx[:, 1]])
print(f"splits {os.environ.get('ECC_PAIR_SPLITS', 'auto')}: mean {m:.6f}; pair kernel (+ finalize) {1e3 * ms / cnt :.2f} us per call, "
      f"{cnt} launches; kappa samples of the {len(idx)} listed pairs: min {listed.min()} median {int(np.median(listed))} max {listed.max()} total {listed.sum()}")
