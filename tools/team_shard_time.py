"""The Radon shards of an 8-rank team (C3 geometry) computed one after the other by a team of ONE on a single GPU: what a
rank's launch costs WITHOUT peer stores, next to whole launches of 60 and 64 projections."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402

n, n_u, n_v, n_a, n_t, world = 496, 1240, 960, 768, 768, int(os.environ.get("WORLD", 8))
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(0)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.308)
imgs = torch.empty((128, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps[:128], n_u, n_v, ELL, imgs)
ctx.team_create(0, 1, n, n_a, n_t)
for r in range(min(world, 4)):
    first, count, part = ctx.team_radon_shard(n, world, r)
    src = imgs[:count]  # the pixel values do not matter for the time
    for _ in range(2):
        ctx.team_radon_compute_part(src, first, part, n_u, n_v, interp=api.INTERP_HYBRID_STATIC)
    torch.cuda.synchronize()
    ctx.profile_reset()
    ctx.profile_enable(True)
    K = 5
    for _ in range(K):
        ctx.team_radon_compute_part(src, first, part, n_u, n_v, interp=api.INTERP_HYBRID_STATIC)
    torch.cuda.synchronize()
    ms, cnt = ctx.profile_get("radon")
    st, _ = ctx.profile_get("stage")
    ctx.profile_enable(False)
    quads = count / 4 - (part[0] / part[2]) - (1 - part[1] / part[2])
    print(f"rank {r} of {world}: projections {first}..{first + count - 1}, part {part} = {quads:.2f} quads: Radon kernels {ms / K:.3f} ms "
          f"({cnt // K} launches) = {ms / K / (4 * quads):.4f} ms/projection; staging {st / K:.3f} ms")
