"""Development aid: the static-split hybrid engine -- speed for several split shares, reproducibility, shard invariance."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
import bench
W = bench.WORKLOADS["c3"]
n, n_u, n_v, n_a, n_t = int(os.environ.get("N_PROJ", 64)), W["n_u"], W["n_v"], W["n_alpha"], W["n_t"]
ctx = api.Context(0)
Ps = api.make_circular_trajectory(496, W["sid"], W["sdd"], n_u, n_v, W["arc"], W["px"])[:n]
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, bench.ELLIPSOIDS, imgs)
def timed(interp):
    out = ctx.radon_compute(imgs, n_a, n_t, interp=interp)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=out)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 3 / n * 1e3, out
ms_d, dyn = timed(api.INTERP_HYBRID)
ms_s, st = timed(api.INTERP_HYBRID_STATIC)
tex = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_TEXTURE)
peak = float(tex.abs().max())
print(f"split {os.environ.get('ECC_HYBRID4_SPLIT', '580')}: dynamic {ms_d:.4f} ms/projection, static {ms_s:.4f}; static vs texture engine {float((st - tex).abs().max()) / peak:.2e} of peak")
if os.environ.get("CHECK"):
    st2 = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_HYBRID_STATIC)
    print("static run to run equal:", bool(torch.equal(st, st2)), "| dynamic run to run equal:", bool(torch.equal(dyn, ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_HYBRID))))
    parts = torch.cat([ctx.radon_compute(imgs[a:b], n_a, n_t, interp=api.INTERP_HYBRID_STATIC) for a, b in ((0, 5), (5, 7), (7, 30), (30, n))])
    print("static in shards of 5, 2, 23, rest equal:", bool(torch.equal(st, parts)))
