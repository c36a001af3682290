"""Direct metric (MetricDirect) at the size of BASELINE config C1 (100 views of 512 x 512, all 4950 pairs, automatic plane step):
one launch for all pairs against the reference's own kernel (oracle/_ref, same GPU) on the lines of single pairs.
Usage (GPU box): python tools/direct_bench.py [n_views] > profiles/direct_bench_r02.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402  (test infrastructure: the reference kernel and the work counter only)
from epipolarconsistency_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_u = n_v = 512
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context()
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 200.0, 0.6)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
ctx.set_projection_matrices(Ps)
ctx.direct_set_images(imgs)
imgs_h = imgs.cpu().numpy()


def samples_of(lines):
    """Texture samples the kernel takes on these lines (derivative: two per position)."""
    l = lines.astype(np.float64)
    o = -l[:, 2:3] * l[:, :2]
    d = np.stack([l[:, 1], -l[:, 0]], 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        ts = np.stack([(1 - o[:, 0]) / d[:, 0], (n_u - 1 - o[:, 0]) / d[:, 0], (1 - o[:, 1]) / d[:, 1], (n_v - 1 - o[:, 1]) / d[:, 1]], 1)
    ts.sort(axis=1)
    p = o + ts[:, 1:2] * d
    ok = (p[:, 0] <= n_u) & (p[:, 1] <= n_v) & (p[:, 0] >= 0) & (p[:, 1] >= 0) & (ts[:, 2] >= ts[:, 1])
    return float((np.floor((ts[:, 2] - ts[:, 1]) / 0.4) + 1)[ok].sum())


for fbcc in (False, True):
    ctx.direct_set_fan_beam(fbcc)
    ctx.direct_evaluate(None)  # warm-up
    torch.cuda.synchronize()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        total = ctx.direct_evaluate(None)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    # work: positions along the lines of a sample of pairs, scaled
    rng = np.random.default_rng(1)
    pick = [(int(i), int(j)) for i, j in zip(rng.integers(0, n, 40), rng.integers(0, n, 40)) if i < j][:12]
    pos, ref_ms, ours_lines = 0.0, 0.0, 0
    for (i, j) in pick:
        g = ctx.direct_pair_geometry(i, j)
        pos += samples_of(g["lines0"]) + samples_of(g["lines1"])
        ours_lines += len(g["kappas"])
        if ol.ref_cuda() is not None:
            _, m0 = ol.ref_cuda_direct_line_integrals(imgs_h[i], g["lines0"], g["fbcc0"] if fbcc else None)
            _, m1 = ol.ref_cuda_direct_line_integrals(imgs_h[j], g["lines1"], g["fbcc1"] if fbcc else None)
            ref_ms += m0 + m1
    n_pairs = n * (n - 1) // 2
    pos_total = pos / len(pick) * n_pairs
    fetches = pos_total * (1 if fbcc else 2)
    print(f"{'fan-beam' if fbcc else 'derivative'}: {n} views {n_u}x{n_v}, {n_pairs} pairs, {ours_lines / len(pick):.0f} planes per pair: "
          f"{ms:.2f} ms per evaluate() (host call to host result), {n_pairs / ms * 1e3:.0f} pairs/s, sum {total:.6g}; "
          f"{fetches:.3e} texture fetches -> {fetches / ms * 1e-6:.0f} G fetches/s "
          f"(texture unit alone, measured: 1.09e12/s at 1965 MHz, profiles/tex_probe_r01.txt -> {fetches / ms * 1e3 / 1.09e12:.2f})")
    if ref_ms:
        ref_total = ref_ms / len(pick) * n_pairs
        print(f"    reference kernel (oracle/_ref, two launches per pair, kernel time only, {len(pick)} pairs scaled): {ref_total:.1f} ms for all pairs "
              f"-> {ref_total / ms:.2f}x; the reference's host geometry, 5 copies and 2 syncs per pair come on top")
