"""Development aid: Radon intermediates of one rank's share of C3 at 8 GPUs (62 projections) from device-resident and from
pinned host images: where do the milliseconds between `value` and `e2e` go?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api

ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
n = int(os.environ.get("N_PROJ", 62))
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
host = torch.empty((n, n_v, n_u), dtype=torch.float32, pin_memory=True)
host.copy_(imgs)
out = torch.empty((n, n_t, n_a), dtype=torch.float32, device="cuda")
for name, src in (("device", imgs), ("host", host)):
    for _ in range(2):
        ctx.radon_compute(src, n_a, n_t, interp=api.INTERP_HYBRID, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.radon_compute(src, n_a, n_t, interp=api.INTERP_HYBRID, out=out)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:7s} images, {n} projections: {min(ts):.2f} ms (median {sorted(ts)[len(ts)//2]:.2f})", flush=True)
# plain upload rate
t0 = time.perf_counter()
for _ in range(3):
    imgs.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"H2D of {host.numel() * 4 / 1e6:.0f} MB: {dt * 1e3:.2f} ms = {host.numel() * 4 / dt / 1e9:.1f} GB/s")
