"""Development aid: time the all-pairs evaluation at C3 size and print a checksum of the per-pair values."""
import hashlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
n = int(os.environ.get("N_PROJ", 496))
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_TEXTURE)
del imgs
for interp, name in ((api.INTERP_TEXTURE, "texture"), (api.INTERP_EXACT, "exact")):
    ctx.set_interpolation(interp)
    ctx.set_object_radius(0)
    ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    ctx.evaluate(cost)
    ctx.profile_reset(); ctx.profile_enable(True)
    for _ in range(5):
        mean = ctx.evaluate(cost)
    ms, k = ctx.profile_get("pairs")
    ctx.profile_enable(False)
    h = hashlib.sha1(cost.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"MINB={os.environ.get('ECC_PAIRS_MINB','4')} {name}: pairs kernel {ms / k:.3f} ms, mean {mean!r}, sha1(cost) {h}", flush=True)
