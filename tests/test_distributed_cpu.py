"""Host-side logic of the multi-GPU path on CPU: world_size 2 over gloo, with a stub compute object in place of
the CUDA context (the product kernels need a GPU; what is tested here is sharding, exchange and reduction)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from epipolarconsistency_b200.distributed import ShardedPipeline, shard_bounds


class StubCompute:
    """Deterministic stand-in for api.Context: 'Radon' of an image is its mean broadcast over the bins; pair (i,j)
    costs 1000*i + j."""

    def __init__(self, n_views, world):
        self.n, self.world = n_views, world

    def radon_compute(self, images, n_alpha, n_t, out=None, **kw):
        out.copy_(images.mean(dim=(1, 2))[:, None, None].expand(-1, n_t, n_alpha))
        return out

    def partition_pairs(self, n_parts):
        total = self.n * (self.n - 1) // 2
        return np.array(shard_bounds(total, n_parts), np.int64)[::1]

    def evaluate_batch(self, Ps_sets, idx4=None):
        return np.array([float(P[0, 0]) * 2.0 + 1.0 for P in Ps_sets], np.float64)  # a set's "mean" = a function of the set

    def evaluate_batch_params(self, base_Ps, params, view_to_param=None, idx4=None):
        return np.array([float(x[0, 0]) * 3.0 - 1.0 for x in params], np.float64)  # a set's "mean" = a function of its parameters

    def direct_partition(self, n_parts):
        total = self.n * (self.n - 1) // 2
        return np.array(shard_bounds(total, n_parts), np.int64)

    def direct_evaluate_range(self, lo, hi, cost_image=None):
        return 0.5 * self.evaluate_range(lo, hi, cost_image)  # the "direct" value of a pair = half its intermediate value

    def evaluate_range(self, lo, hi, cost_image=None, want_sum=True):
        pairs = [(i, j) for i in range(self.n) for j in range(i + 1, self.n)][lo:hi]
        s = 0.0
        for i, j in pairs:
            v = 1000.0 * i + j
            s += v
            if cost_image is not None:
                cost_image[j, i] = v
        return s


class StubTeam(StubCompute):
    """The team entry points of api.Context without a GPU: the "peer stores" are emulated with a gloo all-gather inside
    the stub, so that what is tested is the pipeline's set-up logic (handle exchange in rank order, the common decision,
    the fall-back) and its routing of the two stages."""

    def __init__(self, n_views, world, rank, fail_create=False):
        super().__init__(n_views, world)
        self.rank, self.fail_create = rank, fail_create
        self.connected_with = None
        self.destroyed = False

    def team_create(self, rank, world, n_total, n_alpha, n_t):
        if self.fail_create:
            raise RuntimeError("no peer access")
        self._full = torch.zeros((n_total, n_t, n_alpha))
        return f"handle-of-rank-{rank}".encode()

    def team_connect(self, handles):
        self.connected_with = list(handles)

    def team_destroy(self):
        self.destroyed = True

    def team_dtrs(self):
        return self._full

    def team_radon_compute(self, images, first, n_u, n_v, **kw):
        n_total, n_t, n_alpha = self._full.shape
        bounds = shard_bounds(n_total, self.world)
        mine = torch.zeros((max(b - a for a, b in zip(bounds, bounds[1:])), n_t, n_alpha))
        if images is not None:
            mine[:images.shape[0]] = images.mean(dim=(1, 2))[:, None, None]
        parts = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine)
        for r in range(self.world):
            self._full[bounds[r]:bounds[r + 1]] = parts[r][:bounds[r + 1] - bounds[r]]

    def team_evaluate(self, cost_image=None, want_mean=True):
        total = self.n * (self.n - 1) // 2
        return self.evaluate_range(0, total, cost_image) / total

    # sharding in quads of projections (static-split engine): the library's own shard function (host only, no GPU needed)
    def team_radon_shard(self, n_total, world, rank):
        import ctypes as C
        from epipolarconsistency_b200 import _lib
        v = [C.c_int() for _ in range(5)]
        assert _lib.load().ecc_team_radon_shard(n_total, world, rank, *[C.byref(x) for x in v]) == 0
        return v[0].value, v[1].value, (v[2].value, v[3].value, v[4].value)

    def team_radon_compute_part(self, images, first, part, n_u, n_v, **kw):
        """Emulation: a rank "computes" the bins of its whole quads and, of a shared quad, the bins of its share of the
        flattened bin list; the peer stores are an all-gather of (value, written) and every bin must be written ONCE."""
        n_total, n_t, n_alpha = self._full.shape
        lo_num, hi_num, den = part
        val = torch.zeros((n_total, n_t * n_alpha))
        hit = torch.zeros((n_total, n_t * n_alpha))
        count = 0 if images is None else images.shape[0]
        bins = n_t * n_alpha
        for k in range(count):
            quad = k // 4
            b0 = bins * lo_num // den if quad == 0 else 0
            b1 = bins * hi_num // den if quad == (count + 3) // 4 - 1 else bins
            val[first + k, b0:b1] = images[k].mean()
            hit[first + k, b0:b1] = 1
        vals = [torch.zeros_like(val) for _ in range(self.world)]
        hits = [torch.zeros_like(hit) for _ in range(self.world)]
        dist.all_gather(vals, val)
        dist.all_gather(hits, hit)
        assert torch.equal(sum(hits), torch.ones_like(hit)), "a bin was computed twice or not at all"
        self._full.copy_(sum(vals).reshape(n_total, n_t, n_alpha))


def _worker(rank, world, port, n_total, results, mode="nccl"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode in ("nccl", "batch", "direct"):
            compute, transport = StubCompute(n_total, world), "nccl"
        else:
            compute, transport = StubTeam(n_total, world, rank, fail_create=(mode == "team-fails" and rank == 1)), "team"
        pipe = ShardedPipeline(compute, rank, world, device="cpu", transport=transport)
        kw = {"interp": 3} if mode == "team-quads" else {}  # the static-split engine: sharded in quads of projections
        lo, hi, part = pipe.radon_shard(n_total, 5, 4, **kw)
        if mode == "team-quads":
            assert part is not None and lo % 4 == 0
        else:
            bounds = shard_bounds(n_total, world)
            assert (lo, hi, part) == (bounds[rank], bounds[rank + 1], None)
        local = torch.stack([torch.full((6, 8), float(k)) for k in range(lo, hi)]) if hi > lo else torch.zeros((0, 6, 8))
        full = pipe.radon_allgather(local, n_total, 5, 4, **kw)
        cost = torch.zeros((n_total, n_total))
        mean = pipe.evaluate_all_pairs(n_total, cost)
        extra = None
        if mode == "batch":
            sets = np.arange(7 * n_total * 12, dtype=np.float64).reshape(7, n_total, 12)  # K = 7 sets: ragged over 2 ranks
            extra = pipe.evaluate_batch(sets).tolist()
            params = np.arange(7 * n_total * 11, dtype=np.float64).reshape(7, n_total, 11)  # the same with parameter vectors
            extra = (extra, pipe.evaluate_batch_params(None, params).tolist())
        elif mode == "direct":
            dcost = torch.zeros((n_total, n_total))
            extra = (pipe.direct_evaluate(n_total, dcost), dcost.numpy().copy())
        elif mode != "nccl":
            extra = (pipe.transport, pipe.team_error, compute.connected_with, compute.destroyed)
        results[rank] = (full[:, 0, 0].tolist(), mean, cost.numpy().copy(), extra)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(n_total, world=2, mode="nccl"):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_total, results, mode), nprocs=world, join=True)
    return dict(results)


def test_shard_bounds():
    assert shard_bounds(496, 8) == [0, 62, 124, 186, 248, 310, 372, 434, 496]
    assert shard_bounds(5, 2) == [0, 3, 5]
    assert shard_bounds(1, 4) == [0, 1, 1, 1, 1]


def _check(n_total, mode="nccl"):
    res = _run(n_total, mode=mode)
    pairs = [(i, j) for i in range(n_total) for j in range(i + 1, n_total)]
    want_mean = np.mean([1000.0 * i + j for i, j in pairs])
    for rank in (0, 1):
        dtr_vals, mean, cost, extra = res[rank]
        if mode in ("team", "team-quads"):  # both ranks connected with the handles in rank order
            assert extra[0] == "team" and extra[1] is None
            assert extra[2] == [b"handle-of-rank-0", b"handle-of-rank-1"]
        if mode == "team-fails":  # one rank could not create its block: BOTH fall back to the collectives and say why
            assert extra[0] == "nccl" and "rank 1: no peer access" in extra[1]
            assert extra[2] is None and extra[3]
        assert dtr_vals == [float(k) for k in range(n_total)]  # every rank ends with all dtrs, in order
        assert abs(mean - want_mean) < 1e-9
        for i, j in pairs:
            assert cost[j, i] == 1000.0 * i + j  # every pair written exactly once across the ranks


def test_world2_even_shards():
    _check(6)


def test_world2_ragged_shards():
    _check(5)


def test_world2_team_transport():
    _check(5, mode="team")


def test_world2_team_shards_in_quads_for_the_static_split_engine():
    """11 projections = 3 quads over 2 ranks: 1.5 quads each, the middle quad shared; every bin written exactly once."""
    _check(11, mode="team-quads")
    _check(8, mode="team-quads")  # 2 quads: nothing shared


def test_world2_team_falls_back_together():
    _check(6, mode="team-fails")


def test_world2_batched_sets_are_sharded_and_gathered():
    n_total = 4
    res = _run(n_total, mode="batch")
    want = [float(k * n_total * 12) * 2.0 + 1.0 for k in range(7)]
    want_params = [float(k * n_total * 11) * 3.0 - 1.0 for k in range(7)]
    for rank in (0, 1):
        assert res[rank][3][0] == want  # every rank ends with all K means, in set order
        assert res[rank][3][1] == want_params  # and so for sets given as parameter vectors


def test_world2_direct_metric_is_sharded_by_pairs_and_summed():
    """ShardedPipeline.direct_evaluate: the pair enumeration cut over the ranks, one all-reduce of the sum and of the cost
    image with its disjoint entries; every rank ends with the whole."""
    n_total = 7
    res = _run(n_total, mode="direct")
    pairs = [(i, j) for i in range(n_total) for j in range(i + 1, n_total)]
    want = sum(0.5 * (1000.0 * i + j) for i, j in pairs)
    for rank in (0, 1):
        total, dcost = res[rank][3]
        assert abs(total - want) < 1e-9
        for i, j in pairs:
            assert dcost[j, i] == 1000.0 * i + j  # the stub writes a pair's intermediate value; written exactly once

