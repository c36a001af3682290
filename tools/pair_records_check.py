"""Development aid (GPU box): SHA-1 of every pair value of C3 (122 760 pairs) and of a C4 batch (8 perturbed sets x 30 628 pairs,
values and means), and the pair-stage time -- run once with ECC_PAIR_RECORDS=0 (maps computed by every warp) and once
without (maps from pair_records_kernel): the hashes must be equal.  Coarse intermediates (the maps do not depend on them)."""
import hashlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
import bench
W = bench.WORKLOADS["c3"]
n, n_u, n_v = W["n"], W["n_u"], W["n_v"]
n_a = n_t = int(os.environ.get("BINS", 256))
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, bench.GEO["sid"], bench.GEO["sdd"], n_u, n_v, W["arc"], W["px"])
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, bench.ELLIPSOIDS, imgs)
dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_HYBRID_STATIC)
del imgs
ctx.set_interpolation(api.INTERP_TEXTURE)
ctx.set_object_radius(0.0)
ctx.set_epipolar_plane_step(float(np.deg2rad(W["dkappa_deg"])))
ctx.set_projection_matrices(Ps)
ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")
mean = ctx.evaluate(cost)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    ctx.evaluate(cost)
torch.cuda.synchronize()
print("c3 mean", repr(mean), "sha1", hashlib.sha1(cost.cpu().numpy().tobytes()).hexdigest(), "ms/evaluate %.3f" % ((time.perf_counter() - t0) / 5 * 1e3))
# C4 shape: every second view, K perturbed sets
n2 = n // 2
Ps2 = np.ascontiguousarray(Ps[::2])
ctx.set_projection_matrices(Ps2)
ctx.set_radon_intermediates(dtrs[::2].contiguous(), n_u, n_v, True)
K = int(os.environ.get("SETS", 16))
sets, _ = bench.perturbed_sets(api, Ps2, K, np.random.default_rng(42))
sets_d = torch.from_numpy(np.ascontiguousarray(sets)).cuda()
vals = torch.zeros((K, n2 * (n2 - 1) // 2), dtype=torch.float32, device="cuda")
means = ctx.evaluate_batch(sets_d, out=vals)
print("c4 values sha1", hashlib.sha1(vals.cpu().numpy().tobytes()).hexdigest())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    means = ctx.evaluate_batch(sets_d)
torch.cuda.synchronize()
print("c4 means sha1", hashlib.sha1(np.asarray(means, np.float64).tobytes()).hexdigest(), "first", repr(float(means[0])), repr(float(means[1])),
      "ms/launch of %d sets %.3f" % (K, (time.perf_counter() - t0) / 3 * 1e3))
