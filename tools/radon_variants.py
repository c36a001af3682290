"""Development aid: time one Radon mapping variant (ECC_RADON_MAP / ECC_RADON_BY env) on a few full-size images."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
n = int(os.environ.get("N_PROJ", 16))
interp = int(os.environ.get("INTERP", 0))
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
out = torch.empty((n, n_t, n_a), dtype=torch.float32, device="cuda")
ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=out)
torch.cuda.synchronize()
reps = int(os.environ.get("REPS", 3))
ctx.profile_enable(True)
t0 = time.perf_counter()
for _ in range(reps):
    ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=out)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
kms, kn = ctx.profile_get('radon')
print(f"QA={os.environ.get("ECC_RADON_QA","2")} GA={os.environ.get("ECC_RADON_GA","1")} WA={os.environ.get("ECC_RADON_WA","4")} BY={os.environ.get("ECC_RADON_BY","8")} interp={interp}: {1e3*dt/n:.3f} ms/projection "
      f"(kernel only {kms/reps/n:.3f} ms/proj, {kn} launches) {1.352e9*n/dt:.3e} samples/s checksum {float(out.double().sum()):.6f} absmax {float(out.abs().max()):.4f}")
