"""How far do per-pair ECC values move when the Radon intermediates come from the hybrid engine instead of the texture
engine (which is bit-identical to the reference kernel)?  C3 size, all 122 760 pairs, same pair kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
n = int(os.environ.get("N_PROJ", 496))
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ELL = np.array([[0.0, 0.0, 0.0, 80.0, 60.0, 70.0, 1.0], [20.0, -10.0, 5.0, 25.0, 30.0, 20.0, 0.6], [-25.0, 15.0, -10.0, 20.0, 22.0, 28.0, -0.5],
                [5.0, 30.0, 20.0, 22.0, 20.0, 24.0, 0.8], [-10.0, -30.0, -25.0, 30.0, 21.0, 20.0, -0.7]])
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
res = {}
for name, interp in (("texture", api.INTERP_TEXTURE), ("hybrid", api.INTERP_HYBRID), ("hybrid2", api.INTERP_HYBRID)):
    dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=interp)
    ctx.set_interpolation(api.INTERP_TEXTURE)
    ctx.set_object_radius(0)
    ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    mean = ctx.evaluate(cost)
    c = cost.cpu().numpy()
    iu = np.tril_indices(n, -1)
    res[name] = (mean, c[iu].astype(np.float64), dtrs)
a = res["texture"]
for other in ("hybrid", "hybrid2"):
    b = res[other]
    d = float((a[2] - b[2]).abs().max() / a[2].abs().max())
    floor = 1e-6 * a[1].max()
    rel = np.abs(b[1] - a[1]) / np.maximum(a[1], floor)
    q = np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 0.9999])
    print(f"{other} vs texture: dtr max diff {d:.3e} of peak; mean {b[0]!r} vs {a[0]!r} (rel {abs(b[0]-a[0])/a[0]:.2e}); per-pair rel diff "
          f"quantiles 50/90/99/99.9/99.99% {q}, max {rel.max():.3e}, fraction > 1e-3: {(rel > 1e-3).mean():.2e}, > 1e-4: {(rel > 1e-4).mean():.2e}")
