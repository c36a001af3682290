"""Radon quad kernel against the launch size (one GPU, images resident): ms per projection for batches of n projections."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402

n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
N = 128
Ps = api.make_circular_trajectory(N, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((N, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
out = torch.empty((N, n_t, n_a), dtype=torch.float32, device="cuda")
for n in (128, 64, 62, 32, 16, 8, 4):
    for _ in range(2):
        ctx.radon_compute(imgs[:n], n_a, n_t, out=out[:n], interp=api.INTERP_HYBRID_STATIC)
    torch.cuda.synchronize()
    ctx.profile_reset()
    ctx.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for _ in range(K):
        ctx.radon_compute(imgs[:n], n_a, n_t, out=out[:n], interp=api.INTERP_HYBRID_STATIC)
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = ctx.profile_get("radon")
    st, _ = ctx.profile_get("stage")
    ctx.profile_enable(False)
    print(f"n = {n:3d}: call {e0.elapsed_time(e1) / K:8.3f} ms = {e0.elapsed_time(e1) / K / n:.4f} ms/projection; "
          f"Radon kernel {ms / K:8.3f} ms = {ms / K / n:.4f} ms/projection; staging {st / K:.3f} ms")
