"""Development aid: hybrid Radon kernel versus the texture-only kernel -- difference and time, several sizes."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api

ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
cases = [(6, 160, 128, 128, 128, 2.0), (5, 203, 301, 100, 90, 1.5), (3, 96, 64, 50, 70, 3.0), (int(os.environ.get("N_PROJ", 32)), 1240, 960, 768, 768, 0.308)]
if os.environ.get("ONLY_BIG"):
    cases = cases[-1:]
for n, n_u, n_v, n_a, n_t, px in cases:
    Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, px)
    imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
    ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
    imgs += 0.05 * torch.rand_like(imgs)  # texture everywhere, also at the borders
    outs = {}
    for name, interp in (("texture", api.INTERP_TEXTURE), ("hybrid", api.INTERP_HYBRID)):
        out = torch.full((n, n_t, n_a), float("nan"), dtype=torch.float32, device="cuda")
        ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=out)
        torch.cuda.synchronize()
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=out)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        outs[name] = out
        print(f"{n}x {n_u}x{n_v}->{n_a}x{n_t} {name:8s}: {1e3 * dt / n:.3f} ms/projection, finite {bool(torch.isfinite(out).all())}", flush=True)
    a, b = outs["texture"], outs["hybrid"]
    peak = float(a.abs().max())
    d = (a - b).abs()
    print(f"   max|hybrid-texture| = {float(d.max()):.3e} = {float(d.max()) / peak:.3e} of peak {peak:.3f}; bins differing {int((d > 0).sum())} of {d.numel()};"
          f" bins > 1e-4 peak: {int((d > 1e-4 * peak).sum())}", flush=True)
    bad = torch.nonzero(d > 1e-4 * peak)
    for row in bad[:12].tolist():
        k, iy, ix = row
        print(f"      bad bin img {k} iy {iy} ix {ix}: texture {float(a[k, iy, ix]):.4f} hybrid {float(b[k, iy, ix]):.4f}")
