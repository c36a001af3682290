"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck, one tool per gpurun call):
Radon engines (texture, exact, hybrid, hybrid-static incl. a ragged size and a remainder quad), pre-processing, all-pairs /
list / batched / parameter-batched metric, tracking graph, team of one.  Sizes are tiny: the tools slow kernels 10-100x."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402

ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5]])
ctx = api.Context(0)
for (n, n_u, n_v, n_a, n_t) in ((6, 96, 80, 64, 64), (5, 75, 52, 40, 37)):
    Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 3.0)
    imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
    ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
    p = api.PreprocessParams.defaults()
    ctx.preprocess(imgs.clone(), p, Ps=Ps)
    outs = {}
    for name, interp in (("texture", api.INTERP_TEXTURE), ("exact", api.INTERP_EXACT), ("hybrid", api.INTERP_HYBRID), ("static", api.INTERP_HYBRID_STATIC)):
        outs[name] = ctx.radon_compute(imgs, n_a, n_t, interp=interp)
    torch.cuda.synchronize()
    err = float((outs["static"] - outs["texture"]).abs().max() / outs["texture"].abs().max())
    assert err < 1e-4, err
    ctx.radon_compute(imgs[:2], n_a, n_t, filter=api.FILTER_RAMP, interp=api.INTERP_TEXTURE)
    dtrs = outs["static"]
    ctx.set_interpolation(api.INTERP_TEXTURE)
    ctx.set_object_radius(0.0)
    ctx.set_epipolar_plane_step(float(np.deg2rad(0.5)))
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    cost = np.zeros((n, n), np.float32)
    mean = ctx.evaluate(cost)
    idx = np.array([(n - 1, i, n - 1, i) for i in range(n - 1)], np.int32)
    out = np.zeros(n - 1, np.float32)
    ctx.evaluate_indices(idx, out)
    for k in range(4):
        ctx.update_and_evaluate(n - 1, Ps[n - 1] * (1 + 1e-6 * k), idx, out)
    x = np.zeros((3, n, 11))
    x[1:, :, 0] = 0.5
    x[2, :, 8] = 0.002
    means = ctx.evaluate_batch_params(Ps, x)
    sets = ctx.model_expand(Ps, x)
    means2 = ctx.evaluate_batch(sets)
    assert np.array_equal(means, means2), (means, means2)
    ctx.use_correlation(True)
    ctx.evaluate(cost)
    ctx.use_correlation(False)
    ctx.pair_signals(0, n - 1)
    ctx.pair_maps(n_views=n)
    ctx.partition_pairs(3)
    print(f"ok {n_u}x{n_v}: mean {mean:.6g}, batch means {means}")
team = api.Context(0)
team.team_create(0, 1, 6, 64, 64)
team.close()
ctx.close()
print("sanitize target finished")
