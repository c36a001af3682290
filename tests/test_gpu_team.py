"""Multi-GPU team (include/ecc_b200.h "Multi-GPU team", csrc/ecc_team.cu) on the GPU box: every rank's results must be the
single-GPU results BIT FOR BIT (texture engine: deterministic), whichever way the blocks are mapped --
  * ranks as threads of one process, blocks connected by pointer (ecc_team_connect_pointers),
  * ranks as processes, blocks connected through CUDA IPC handles carried by torch.distributed (gloo) -- the bench's way.
The ranks' flag barriers are kernels that wait on one another: every rank needs a GPU of its own (nothing guarantees that
two such kernels are resident at the same time on ONE GPU -- B200_PROFILING.md; sharing a GPU can end in a context-switch
time-out).  The multi-rank tests therefore run with `gpurun --gpus 2` (logs in profiles/) and skip on a one-GPU box, where
the team of ONE and the CPU tests of the host logic (tests/test_distributed_cpu.py, gloo) remain."""
import os
import socket
import subprocess
import sys
import threading

import numpy as np
import pytest

from epipolarconsistency_b200 import api
from team_scene import make_scene

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def need_gpus(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"{world} ranks need {world} GPUs (flag barriers wait on one another); this box has {torch.cuda.device_count()}")


def test_team_of_one_is_the_plain_path():
    """world = 1: the team block, the alternating value buffers and the fixed-order sum without any peer -- same bits as the
    plain calls, evaluation after evaluation."""
    S = make_scene()
    n, n_u, n_v, n_a, n_t = S["n"], S["n_u"], S["n_v"], S["n_a"], S["n_t"]
    want_dtrs, want_cost, want_mean = single_gpu(S, api.INTERP_HYBRID_STATIC)
    c = api.Context()
    try:
        c.team_create(0, 1, n, n_a, n_t)
        c.set_interpolation(api.INTERP_TEXTURE)
        c.set_epipolar_plane_step(S["dkappa"])
        c.team_radon_compute(S["imgs"], 0, n_u, n_v, interp=api.INTERP_HYBRID_STATIC)
        c.team_set_radon_intermediates(n_u, n_v, True)
        for step in range(3):
            c.set_projection_matrices(S["Ps"])
            cost = np.zeros((n, n), np.float32)
            mean = c.team_evaluate(cost)
            assert mean == want_mean and np.array_equal(cost, want_cost), step
        assert np.array_equal(c.team_dtrs().cpu().numpy(), want_dtrs)
        sets = loop_sets(S, 3)
        c.set_projection_matrices(sets[2])
        moved = c.team_evaluate(None)
        assert moved > want_mean
    finally:
        c.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_quad_shards_of_a_team_compose_to_the_single_gpu_result(world):
    """Sharding in quads of projections (ecc_team_radon_shard / ecc_team_radon_compute_part): the shards of `world` ranks,
    computed one after the other by a team of one into the same block, give the single-GPU intermediates bit for bit -- every
    bin of a shared quad is computed by exactly one of the two ranks that share it, with the same arithmetic."""
    S = make_scene()
    n, n_u, n_v, n_a, n_t = S["n"], S["n_u"], S["n_v"], S["n_a"], S["n_t"]
    want_dtrs, _, _ = single_gpu(S, api.INTERP_HYBRID_STATIC)
    c = api.Context()
    try:
        c.team_create(0, 1, n, n_a, n_t)
        c.team_dtrs().fill_(float("nan"))
        covered = np.zeros(n, np.int32)
        shared = 0
        for r in range(world):
            first, count, part = c.team_radon_shard(n, world, r)
            assert first % 4 == 0 and 0 <= first and first + count <= n
            covered[first:first + count] += 1
            shared += part[0] != 0
            if count:
                c.team_radon_compute_part(S["imgs"][first:first + count], first, part, n_u, n_v, interp=api.INTERP_HYBRID_STATIC)
        assert covered.min() >= 1
        assert (shared > 0) == (((n + 3) // 4) % world != 0)  # 11 projections are 3 quads: shared among 2 and among 8 ranks
        assert np.array_equal(c.team_dtrs().cpu().numpy(), want_dtrs)
        # engines without a static split refuse a part of a quad instead of computing something else
        first, count, part = c.team_radon_shard(n, 2, 0)
        with pytest.raises(api.EccError):
            c.team_radon_compute_part(S["imgs"][first:first + count], first, part, n_u, n_v, interp=api.INTERP_TEXTURE)
    finally:
        c.close()


def single_gpu(S, interp):
    c = api.Context()
    try:
        dtrs = c.radon_compute(S["imgs"], S["n_a"], S["n_t"], interp=interp)
        c.set_interpolation(api.INTERP_TEXTURE)
        c.set_epipolar_plane_step(S["dkappa"])
        c.set_radon_intermediates(dtrs, S["n_u"], S["n_v"], True)
        c.set_projection_matrices(S["Ps"])
        cost = np.zeros((S["n"], S["n"]), np.float32)
        mean = c.evaluate(cost)
        return dtrs, cost, mean
    finally:
        c.close()


@pytest.mark.parametrize("world", [2, 3])
def test_team_threads_bit_identical_to_single_gpu(world):
    import torch
    need_gpus(world)
    from epipolarconsistency_b200.distributed import shard_bounds
    S = make_scene()
    n, n_u, n_v, n_a, n_t = S["n"], S["n_u"], S["n_v"], S["n_a"], S["n_t"]
    want_dtrs, want_cost, want_mean = single_gpu(S, api.INTERP_TEXTURE)
    n_dev = torch.cuda.device_count()
    ctxs = [api.Context(r % n_dev, stream=None) for r in range(world)]  # own streams: the ranks run concurrently
    try:
        for r, c in enumerate(ctxs):
            c.team_create(r, world, n, n_a, n_t)
        blocks = [c.team_block()[0] for c in ctxs]
        for c in ctxs:
            c.team_connect_pointers(blocks)
        bounds = shard_bounds(n, world)
        out = [None] * world
        errors = []

        def run(r):
            try:
                c = ctxs[r]
                torch.cuda.set_device(r % n_dev)
                c.set_interpolation(api.INTERP_TEXTURE)
                c.set_epipolar_plane_step(S["dkappa"])
                for step in range(2):
                    lo, hi = bounds[r], bounds[r + 1]
                    c.team_radon_compute(S["imgs"][lo:hi] if hi > lo else None, lo, n_u, n_v, interp=api.INTERP_TEXTURE)
                    c.team_set_radon_intermediates(n_u, n_v, True)
                    c.set_projection_matrices(S["Ps"])
                    cost = np.zeros((n, n), np.float32)
                    mean = c.team_evaluate(cost)
                c.synchronize()
                dtrs = c.team_dtrs().cpu().numpy()
                out[r] = (dtrs, cost, mean)
            except Exception as e:  # noqa: BLE001
                errors.append((r, repr(e)))

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=180)
        assert not errors, errors
        for r in range(world):
            dtrs, cost, mean = out[r]
            assert np.array_equal(dtrs, want_dtrs), f"rank {r}: dtrs differ"
            assert np.array_equal(cost, want_cost), f"rank {r}: cost image differs"
            assert mean == want_mean, f"rank {r}: mean {mean} vs {want_mean}"
    finally:
        for c in ctxs:
            c.close()


def loop_sets(S, steps=7):
    """Matrix sets of an optimiser loop over the team scene: step k = view k shifted by k pixels in u (row0 += k * row2,
    column-major 3x4)."""
    sets = []
    for k in range(steps):
        P = S["Ps"].copy()
        for c in range(4):
            P[k, 0 + 3 * c] += float(k) * P[k, 2 + 3 * c]
        sets.append(P)
    return sets


def test_team_back_to_back_evaluations_with_changing_matrices(tmp_path):
    """An optimiser loop: the intermediates stay, the matrices change, ecc_team_evaluate is called again and again with no
    Radon barrier in between.  A rank that is ahead publishes the values of evaluation k+1 into its peers while they may
    still be summing those of evaluation k: the two alternating value buffers keep them apart (round-1 advisor finding).
    Every evaluation on every rank must be the single-GPU result bit for bit, also when the ranks drift apart (the workers
    sleep at different steps).  Ranks are processes, as in the bench (tests/team_worker.py, mode "loop")."""
    world = 2
    need_gpus(world)
    S = make_scene()
    n, n_u, n_v, n_a, n_t = S["n"], S["n_u"], S["n_v"], S["n_a"], S["n_t"]
    sets = loop_sets(S)
    one = api.Context()
    try:
        dtrs = one.radon_compute(S["imgs"], n_a, n_t, interp=api.INTERP_TEXTURE)
        one.set_interpolation(api.INTERP_TEXTURE)
        one.set_epipolar_plane_step(S["dkappa"])
        one.set_radon_intermediates(dtrs, n_u, n_v, True)
        want = []
        for P in sets:
            one.set_projection_matrices(P)
            cost = np.zeros((n, n), np.float32)
            want.append((one.evaluate(cost), cost))
    finally:
        one.close()
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "team_worker.py"), str(r), str(world), port, str(tmp_path), "texture", "loop"],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
            out += "\n[timeout]"
        logs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n----\n".join(logs)
    for r in range(world):
        res = np.load(os.path.join(tmp_path, f"loop{r}.npz"))
        for k in range(len(sets)):
            assert res["means"][k] == want[k][0], f"rank {r} step {k}: mean {res['means'][k]} vs {want[k][0]}"
            assert np.array_equal(res["costs"][k], want[k][1]), f"rank {r} step {k}: cost image differs"
    assert want[3][0] > want[0][0]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("engine", ["texture", "hybrid", "hybrid-static"])
def test_team_processes_over_cuda_ipc(tmp_path, engine):
    world = 2
    need_gpus(world)
    S = make_scene()
    port = str(_free_port())
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "team_worker.py"), str(r), str(world), port, str(tmp_path), engine],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    logs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
            out += "\n[timeout]"
        logs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n----\n".join(logs)
    interp = {"texture": api.INTERP_TEXTURE, "hybrid": api.INTERP_HYBRID, "hybrid-static": api.INTERP_HYBRID_STATIC}[engine]
    want_dtrs, want_cost, want_mean = single_gpu(S, interp)
    res = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in range(world):
        if engine != "hybrid":  # deterministic engines (texture unit only, or hybrid with the static split): bit for bit
            assert np.array_equal(res[r]["dtrs"], want_dtrs)
            assert np.array_equal(res[r]["cost"], want_cost)
            assert res[r]["means"][1] == want_mean
        else:  # which bins take which pipe depends on scheduling: the engine's own tolerance
            peak = np.abs(want_dtrs).max()
            assert np.abs(res[r]["dtrs"] - want_dtrs).max() < 1e-4 * peak
            assert abs(res[r]["means"][1] - want_mean) < 1e-3 * abs(want_mean)
        assert res[r]["means"][0] == pytest.approx(res[r]["means"][1], rel=1e-3)
        # host and device cost images of the same step are the same values
        assert np.array_equal(res[r]["cost_host"], res[r]["cost"])
        assert res[r]["mean_host"] == res[r]["means"][1]
    # batched mode over the ranks = one batched launch on one GPU over the same dtrs
    c = api.Context()
    try:
        c.set_interpolation(api.INTERP_TEXTURE)
        c.set_epipolar_plane_step(S["dkappa"])
        c.set_radon_intermediates(res[0]["dtrs"], S["n_u"], S["n_v"], True)
        c.set_projection_matrices(S["Ps"])
        want_batch = c.evaluate_batch(res[0]["batch_sets"])
    finally:
        c.close()
    for r in range(world):
        assert np.array_equal(res[r]["batch_means"], want_batch)
    assert want_batch[1] > want_batch[0]
    # every rank holds the same bits
    assert np.array_equal(res[0]["dtrs"], res[1]["dtrs"])
    assert np.array_equal(res[0]["cost"], res[1]["cost"])
    assert res[0]["means"][1] == res[1]["means"][1]
