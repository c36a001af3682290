"""Quick timing survey on the GPU box (development aid, not the bench): C3-sized run of both stages in both
interpolation modes, next to the reference's own CUDA kernels (oracle/_ref) on the same data."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
from epipolarconsistency_b200 import api  # noqa: E402

n = int(os.environ.get("N_PROJ", 496))
n_u, n_v, n_a, n_t = 1240, 960, 768, 768
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4],
                [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
ctx.set_stream(torch.cuda.current_stream())
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
torch.cuda.synchronize()
print("images", imgs.shape, float(imgs.max()))
res = {}


def timed(fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


for name, interp in (("texture", api.INTERP_TEXTURE), ("exact", api.INTERP_EXACT)):
    dtrs = torch.empty((n, n_t, n_a), dtype=torch.float32, device="cuda")
    ctx.radon_compute(imgs[:8], n_a, n_t, interp=interp, out=dtrs[:8])  # warm-up
    t, _ = timed(lambda: ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=dtrs))
    res[f"radon_{name}_ms_per_projection"] = 1e3 * t / n
    print(f"radon {name}: {1e3*t:.1f} ms total, {1e3*t/n:.3f} ms/projection, {1.352e9*n/t:.3e} samples/s")
    ctx.set_interpolation(interp)
    ctx.set_object_radius(0)
    ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    ctx.evaluate(None)
    t, mean = timed(lambda: ctx.evaluate(None), reps=3)
    counts = ctx.pair_sample_counts(n)
    res[f"pairs_{name}_ms"] = 1e3 * t
    print(f"pairs {name}: {1e3*t:.2f} ms for {n*(n-1)//2} pairs, kappa samples {counts.sum():.3e} "
          f"(min {counts.min()} med {int(np.median(counts))} max {counts.max()}), mean {mean:.6g}, "
          f"{counts.sum()*64/t/1e9:.1f} GB/s algorithmic")
    if name == "texture":
        dtr_tex = dtrs
        mean_tex = mean

if ol.ref_cuda() is not None:
    k = 4
    ref, ms = ol.ref_cuda_radon(imgs[:k].cpu().numpy(), n_a, n_t)
    err = np.abs(ref - dtr_tex[:k].cpu().numpy()).max() / np.abs(ref).max()
    res["ref_radon_ms_per_projection"] = ms / k
    print(f"reference CUDA radon: {ms/k:.3f} ms/projection; max|ours-ref|/peak = {err:.3g}")
    m = min(n, int(os.environ.get("N_REF", n)))
    R = ol.RefCudaMetric(Ps[:m], dtr_tex[:m].cpu().numpy(), n_u, n_v)
    radius = ctx.get_object_radius()
    R.evaluate(radius, float(np.deg2rad(0.01)))
    ref_mean, ref_out, ms = R.evaluate(radius, float(np.deg2rad(0.01)))
    res["ref_pairs_ms"] = ms
    ctx.set_interpolation(api.INTERP_TEXTURE)
    ctx.set_projection_matrices(Ps[:m])
    ctx.set_radon_intermediates(dtr_tex[:m], n_u, n_v, True)
    cost = np.zeros((m, m), np.float32)
    mean = ctx.evaluate(cost)
    iu = np.tril_indices(m, -1)
    rel = np.abs(cost[iu] - ref_out[iu]) / ref_out[iu]
    q = np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 0.9999])
    print(f"reference CUDA pairs (n={m}): {ms:.2f} ms; mean ref {ref_mean:.6g} ours {mean:.6g}; per-pair rel err max {rel.max():.3g} "
          f"quantiles 50/90/99/99.9/99.99%: {q}; fraction > 1e-3: {(rel > 1e-3).mean():.2e}; ours>ref among those: {(cost[iu] > ref_out[iu])[rel > 1e-3].mean() if (rel > 1e-3).any() else 0:.2f}")
    R.close()
print(json.dumps(res))
