"""Measures the BASELINE.json configurations that are not the headline bench line:
  C1  100 projections 512x512 -> 256x256, all pairs, dkappa auto, 360 deg
  C2  Radon intermediates only, 1240x960 -> 768x768 (a block of projections; per-projection rate)
  C4  batched correction loop: 64 perturbed matrix sets x 248 projections scored in ONE launch against shared dtrs
  C5  tracking: 1 live view vs 400 reference views, repeated {replace one matrix, evaluate(400 listed pairs)}
One JSON object per configuration on stdout.  Usage: python tools/bench_configs.py [c1 c2 c4 c5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402

ELL = np.array([
    [0.0, 0.0, 0.0, 80.0, 60.0, 70.0, 1.0],
    [20.0, -10.0, 5.0, 25.0, 30.0, 20.0, 0.6],
    [-25.0, 15.0, -10.0, 20.0, 22.0, 28.0, -0.5],
    [5.0, 30.0, 20.0, 22.0, 20.0, 24.0, 0.8],
    [-10.0, -30.0, -25.0, 30.0, 21.0, 20.0, -0.7],
])


def cuda_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    R = np.eye(4)
    i, j = [(1, 2), (0, 2), (0, 1)][axis]
    R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
    return R


def perturb(P_colmajor, rng, sigma_px=0.5, sigma_mm=0.5, sigma_deg=0.2):
    """P' = H2D * P * T3D (ModelCameraSimilarity2D3D, LibProjectiveGeometry/Models/ModelCameraSimilarity2D3D.hxx:89-92):
    a 2-D similarity of the detector (shift u,v in px, in-plane rotation) and a 3-D rigid motion (3 shifts in mm,
    3 rotations), all i.i.d. normal."""
    P = P_colmajor.reshape(4, 3).T  # 3x4
    du, dv = rng.normal(0, sigma_px, 2)
    a = np.deg2rad(rng.normal(0, sigma_deg))
    H = np.array([[np.cos(a), -np.sin(a), du], [np.sin(a), np.cos(a), dv], [0, 0, 1.0]])
    T = np.eye(4)
    T[:3, 3] = rng.normal(0, sigma_mm, 3)
    for ax in range(3):
        T = T @ rot(ax, np.deg2rad(rng.normal(0, sigma_deg)))
    return (H @ P @ T).T.reshape(12)


def scene(ctx, n, n_u, n_v, n_a, n_t, px, arc, interp=api.INTERP_HYBRID):
    Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, arc, px)
    imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
    ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
    dtrs = ctx.radon_compute(imgs, n_a, n_t, interp=interp)
    return Ps, imgs, dtrs


def c1(ctx):
    n, n_u, n_v, n_a, n_t = 100, 512, 512, 256, 256
    Ps, imgs, dtrs = scene(ctx, n, n_u, n_v, n_a, n_t, 0.616, 360.0)
    ms_radon, _ = cuda_time(lambda: ctx.radon_compute(imgs, n_a, n_t, interp=api.INTERP_HYBRID, out=dtrs), 5)
    ctx.set_object_radius(0.0)
    ctx.set_epipolar_plane_step(0.0)
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    ms_pairs, mean = cuda_time(lambda: ctx.evaluate(cost), 10)
    pairs = n * (n - 1) // 2
    return {"config": "C1: 100 proj 512x512 -> 256x256, all pairs, dkappa auto, 360 deg", "radon_ms": ms_radon,
            "radon_intermediates_per_s": n / (ms_radon * 1e-3), "pairs_ms": ms_pairs, "pairs_per_s_metric_only": pairs / (ms_pairs * 1e-3),
            "pairs_per_s_incl_radon": pairs / ((ms_radon + ms_pairs) * 1e-3), "mean_ecc": mean}


def c2(ctx):
    n, n_u, n_v, n_a, n_t = 64, 1240, 960, 768, 768
    Ps, imgs, dtrs = scene(ctx, n, n_u, n_v, n_a, n_t, 0.308, 200.0)
    out = {"config": "C2: Radon intermediates only, 1240x960 -> 768x768, block of 64 projections resident in HBM"}
    samples = ctx.radon_num_samples(n_u, n_v, n_a, n_t)
    for name, interp in (("hybrid", api.INTERP_HYBRID), ("texture", api.INTERP_TEXTURE), ("exact", api.INTERP_EXACT)):
        ms, _ = cuda_time(lambda: ctx.radon_compute(imgs, n_a, n_t, interp=interp, out=dtrs), 3)
        out[name] = {"ms_per_projection": ms / n, "radon_intermediates_per_s": n / (ms * 1e-3), "bilinear_samples_per_s": samples * n / (ms * 1e-3)}
    return out


def c4(ctx):
    n, n_u, n_v, n_a, n_t, K = 248, 1240, 960, 768, 768, 64
    Ps, imgs, dtrs = scene(ctx, n, n_u, n_v, n_a, n_t, 0.308, 200.0)
    del imgs
    ctx.set_object_radius(0.0)
    ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    rng = np.random.default_rng(42)
    sets = np.stack([np.stack([perturb(P, rng) for P in Ps]) for _ in range(K)])
    sets[0] = Ps  # set 0 = unperturbed
    t0 = time.perf_counter()
    means = ctx.evaluate_batch(sets)
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        means = ctx.evaluate_batch(sets)
    wall = (time.perf_counter() - t0) / reps
    single = ctx.evaluate(None)
    pairs = n * (n - 1) // 2
    return {"config": "C4: 64 perturbed matrix sets x 248 projections per launch (dtrs shared), dkappa 0.01 deg",
            "ms_per_launch_host_clock": wall * 1e3, "first_call_ms": first * 1e3, "sets_per_s": K / wall, "pairs_per_s": K * pairs / wall,
            "mean_unperturbed": float(means[0]), "mean_unperturbed_single_call": single, "mean_perturbed_min": float(means[1:].min()),
            "mean_perturbed_max": float(means[1:].max()),
            "note": "host clock around ecc_evaluate_batch: upload of 64x248 matrices, on-device derivation, one pair launch, 64 sums, download"}


def c5(ctx):
    n_ref, n_u, n_v, n_a, n_t = 400, 1240, 960, 768, 768
    n = n_ref + 1
    Ps, imgs, dtrs = scene(ctx, n, n_u, n_v, n_a, n_t, 0.308, 200.0)
    del imgs
    ctx.set_object_radius(0.0)
    ctx.set_epipolar_plane_step(float(np.deg2rad(0.01)))
    ctx.set_projection_matrices(Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    idx = np.array([(n_ref, i, n_ref, i) for i in range(n_ref)], np.int32)  # Gui/SingleImageMotion.h:39-41
    idx_d = torch.from_numpy(idx).cuda()
    rng = np.random.default_rng(7)
    live = [perturb(Ps[n_ref], rng) for _ in range(64)]
    out = {"config": "C5: 1 live view vs 400 reference views, {replace one matrix, evaluate(400 listed pairs)} per call"}
    for name, ix in (("index_list_on_host", idx), ("index_list_resident", idx_d)):
        ctx.profile_reset()
        ctx.profile_enable(True)
        for k in range(50):
            ctx.update_projection_matrix(n_ref, live[k % 64])
            ctx.evaluate_indices(ix)
        pk_ms, pk_n = ctx.profile_get("pairs")
        ctx.profile_enable(False)
        lat = []
        for k in range(2200):
            t0 = time.perf_counter()
            ctx.update_projection_matrix(n_ref, live[k % 64])
            ctx.evaluate_indices(ix)
            lat.append(time.perf_counter() - t0)
        lat = np.array(lat[200:]) * 1e6
        out[name] = {"calls_per_s": 1e6 / lat.mean(), "p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)),
                     "pairs_per_s": n_ref * 1e6 / lat.mean(), "pair_kernel_us": 1e3 * pk_ms / max(pk_n, 1)}
        # the same step through ecc_update_and_evaluate: one recorded CUDA graph per call
        want = []
        for k in range(4):
            ctx.update_projection_matrix(n_ref, live[k])
            want.append(ctx.evaluate_indices(ix))
        got = [ctx.update_and_evaluate(n_ref, live[k % 64], ix) for k in range(8)][4:]
        got = [ctx.update_and_evaluate(n_ref, live[k], ix) for k in range(4)]
        lat = []
        for k in range(2200):
            t0 = time.perf_counter()
            ctx.update_and_evaluate(n_ref, live[k % 64], ix)
            lat.append(time.perf_counter() - t0)
        lat = np.array(lat[200:]) * 1e6
        out[name + "_graph"] = {"calls_per_s": 1e6 / lat.mean(), "p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)),
                                "pairs_per_s": n_ref * 1e6 / lat.mean(), "identical_to_plain_calls": bool(got == want)}
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
    ctx = api.Context(0)
    for w in which:
        r = {"c1": c1, "c2": c2, "c4": c4, "c5": c5}[w](ctx)
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
