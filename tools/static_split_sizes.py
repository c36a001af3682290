"""Development aid: dynamic vs static-split hybrid engine at several geometries."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
for n, n_u, n_v, n_a, n_t, px in [(100, 512, 512, 256, 256, 0.616), (64, 1024, 1024, 512, 512, 0.3), (64, 1920, 1536, 1024, 1024, 0.2), (64, 640, 480, 768, 768, 0.6),
                                  (64, 1240, 960, 360, 512, 0.308), (40, 203, 301, 100, 90, 1.5)]:
    Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, px)
    imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
    ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
    imgs += 0.05 * torch.rand_like(imgs)
    res = {}
    for name, interp in (("dynamic", api.INTERP_HYBRID), ("static", api.INTERP_HYBRID_STATIC), ("texture", api.INTERP_TEXTURE)):
        out = ctx.radon_compute(imgs, n_a, n_t, interp=interp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=out)
        torch.cuda.synchronize()
        res[name] = ((time.perf_counter() - t0) / 3 / n * 1e3, out)
    peak = float(res["texture"][1].abs().max())
    err = float((res["static"][1] - res["texture"][1]).abs().max()) / peak
    print(f"{n}x {n_u}x{n_v}->{n_a}x{n_t}: dynamic {res['dynamic'][0]:.4f}  static {res['static'][0]:.4f}  texture {res['texture'][0]:.4f} ms/projection; static vs texture {err:.2e} of peak", flush=True)
