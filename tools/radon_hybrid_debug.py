"""Development aid: worst bins of hybrid vs texture engine on rough random images of a given shape."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
n_u, n_v, n_a, n_t = [int(x) for x in sys.argv[1:5]]
ctx = api.Context(0)
rng = np.random.default_rng(17)
img = rng.random((5, n_v, n_u), dtype=np.float32) * 10
tex = ctx.radon_compute(img, n_a, n_t, interp=api.INTERP_TEXTURE)
hyb = ctx.radon_compute(img, n_a, n_t, interp=api.INTERP_HYBRID)
d = np.abs(tex - hyb)
peak = np.abs(tex).max()
print("peak", peak, "max err", d.max() / peak, "bad bins", int((d > 1e-4 * peak).sum()))
bad = np.argwhere(d > 1e-4 * peak)
seen = set()
for k, iy, ix in bad[:400]:
    if (iy, ix) in seen: continue
    seen.add((iy, ix))
    alpha = (ix / n_a - 0.5) * 180
    tau = (iy / n_t - 0.5) * np.hypot(n_u, n_v)
    print(f"img {k} iy {iy} ix {ix} alpha {alpha:.2f} deg tau {tau:.1f}: texture {tex[k, iy, ix]:.4f} hybrid {hyb[k, iy, ix]:.4f}")
    if len(seen) > 40: break
