"""Sharding of the hot path over the GPUs of one node: one process per GPU, torch.distributed for the plumbing.

SURVEY.md section 8e: the reference is single-GPU; the path shards naturally with ONE real exchange step.
  stage 1  Radon intermediates: projections are independent -> contiguous block of projections per rank, each
           rank needs only its own images; the kernel writes straight into the rank's slice of the full buffer.
  exchange all-gather of the fp32 dtr blocks (NCCL over NVLink/NVSwitch): every rank ends with all dtrs
           (C3: 1.17 GB total, 146 MB per rank at 8 GPUs).
  stage 2  pairs: the get_ij enumeration is cut into `world` contiguous ranges of equal kappa-sample count
           (ecc_partition_pairs; the cost of a pair varies 891..9000 samples at C3), K0/K1 maps are recomputed
           locally, no exchange.
  final    all-reduce(sum) of the n*n cost image (disjoint entries, 984 KB at C3) and of the scalar sum.

Two transports for the exchange and the final reduction:
  "team"   (default when it can be set up) no collective on the data path: every rank's block of device memory is mapped
           into all other ranks (CUDA IPC over NVLink/NVSwitch, include/ecc_b200.h "Multi-GPU team"); the Radon kernels
           store each bin into all blocks, pair values are published the same way, flag barriers in peer memory order
           the stages.  torch.distributed only carries the 64-byte handles at set-up.  Mean and cost image are the same
           bits on every rank.
  "nccl"   all-gather + all-reduce after the kernels (the first version; kept as the comparison and as the transport
           for ranks without peer access).

The class takes the compute object as a parameter (an `api.Context`); the CPU tests drive the same host logic
with a stub compute object over the gloo backend.
"""
import numpy as np


def shard_bounds(n, world):
    """Contiguous block partition of n items over `world` ranks: bounds[r] .. bounds[r+1]."""
    base, rem = divmod(n, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


class ShardedPipeline:
    def __init__(self, compute, rank=0, world=1, device=None, group=None, transport="nccl"):
        self.c = compute
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        self._full = None
        self.transport = transport if world > 1 else "nccl"
        self._team_key = None
        self.team_error = None

    # -- team set-up: allocate the own block, exchange the handles, map the peers ---------------------------------
    def _ensure_team(self, n_total, n_alpha, n_t):
        """True when the team transport is ready for this data-set shape.  Every rank takes the same decision: a rank
        that fails reports it in the handle exchange and all ranks fall back to the collectives."""
        key = (n_total, n_alpha, n_t)
        if self._team_key == key:
            return True
        if self.transport != "team":
            return False
        import torch.distributed as dist
        handle, err = None, None
        try:
            handle = self.c.team_create(self.rank, self.world, n_total, n_alpha, n_t)
        except Exception as e:  # noqa: BLE001 -- reported to every rank below
            err = f"rank {self.rank}: {e}"
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (handle, err), group=self.group)
        errors = [g[1] for g in gathered if g[1]]
        if not errors:
            try:
                self.c.team_connect([g[0] for g in gathered])
            except Exception as e:  # noqa: BLE001
                err = f"rank {self.rank}: {e}"
            gathered2 = [None] * self.world
            dist.all_gather_object(gathered2, err, group=self.group)
            errors = [g for g in gathered2 if g]
        if errors:
            self.team_error = "; ".join(errors)
            self.transport = "nccl"
            try:
                self.c.team_destroy()
            except Exception:  # noqa: BLE001
                pass
            return False
        self._team_key = key
        self._full = self.c.team_dtrs()
        return True

    # -- stage 1 + exchange -------------------------------------------------------------------------------
    def radon_shard(self, n_total, n_alpha, n_t, interp=None, filter=0):
        """(lo, hi, part): the projections [lo, hi) this rank supplies to radon_allgather.  Contiguous blocks of n / world
        projections -- except with the team transport and the static-split engine (interp 3, derivative filter), which
        works on quads of projections: there the quads are cut into `world` equal intervals and the quad a boundary falls
        into is shared by the two neighbours (`part`, include/ecc_b200.h ecc_team_radon_shard): no rank pads a quad."""
        if self.world > 1 and interp == 3 and filter == 0 and hasattr(self.c, "team_radon_shard") and self._ensure_team(n_total, n_alpha, n_t):
            first, count, part = self.c.team_radon_shard(n_total, self.world, self.rank)
            return first, first + count, part
        bounds = shard_bounds(n_total, self.world)
        return bounds[self.rank], bounds[self.rank + 1], None

    def radon_allgather(self, local_images, n_total, n_alpha, n_t, **radon_kwargs):
        """local_images: this rank's projections [lo, hi) of radon_shard (torch tensor on self.device, or pinned host tensor /
        numpy array for the end-to-end path).  Returns the full (n_total, n_t, n_alpha) dtr tensor."""
        import torch
        lo, hi, part = self.radon_shard(n_total, n_alpha, n_t, radon_kwargs.get("interp"), radon_kwargs.get("filter", 0))
        assert local_images.shape[0] == hi - lo, (local_images.shape, lo, hi)
        if self.world > 1 and self._ensure_team(n_total, n_alpha, n_t):
            n_v, n_u = local_images.shape[1], local_images.shape[2]
            if part is not None:
                self.c.team_radon_compute_part(local_images if hi > lo else None, lo, part, n_u, n_v, **radon_kwargs)
            else:
                self.c.team_radon_compute(local_images if hi > lo else None, lo, n_u, n_v, **radon_kwargs)
            return self._full
        bounds = shard_bounds(n_total, self.world)
        if self._full is None or tuple(self._full.shape) != (n_total, n_t, n_alpha):
            self._full = torch.empty((n_total, n_t, n_alpha), dtype=torch.float32, device=self.device)
        full = self._full
        if hi > lo:
            self.c.radon_compute(local_images, n_alpha, n_t, out=full[lo:hi], **radon_kwargs)
        if self.world > 1:
            import torch.distributed as dist
            if n_total % self.world == 0:
                dist.all_gather_into_tensor(full, full[lo:hi], group=self.group)  # in place: slice r of the output
            else:
                for r in range(self.world):  # ragged shards: one broadcast per owner
                    if bounds[r + 1] > bounds[r]:
                        dist.broadcast(full[bounds[r]:bounds[r + 1]], src=r, group=self.group)
        return full

    # -- stage 2 + final reduce ----------------------------------------------------------------------------
    def evaluate_all_pairs(self, n_views, cost_image=None):
        """Metric over all pairs, partitioned by equal work.  cost_image: (n, n) float32 tensor on self.device or
        None; it must be zero on entry wherever pairs are written (every rank adds its disjoint entries).
        Returns the mean over all pairs (same value on every rank)."""
        import torch
        total = n_views * (n_views - 1) // 2
        if self.world > 1 and self._team_key is not None:
            return self.c.team_evaluate(cost_image)  # complete cost image and the mean, the same bits on every rank
        bounds = self.c.partition_pairs(self.world) if self.world > 1 else np.array([0, total])
        lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
        s = self.c.evaluate_range(lo, hi, cost_image)
        if self.world > 1:
            import torch.distributed as dist
            acc = torch.tensor([s], dtype=torch.float64, device=self.device)
            dist.all_reduce(acc, group=self.group)
            if cost_image is not None:
                dist.all_reduce(cost_image, group=self.group)
            s = float(acc.item())
        return s / total if total else 0.0

    # -- direct metric (MetricDirect): pairs are independent, no exchange but the final sum ---------------------------------
    def direct_evaluate(self, n_views, cost_image=None):
        """MetricDirect::evaluate over the GPUs of the node: every rank holds the images and matrices (direct_set_images,
        set_projection_matrices), the pair enumeration is cut into `world` ranges of equal work (epipolar planes,
        ecc_direct_partition), each rank scores its range; one all-reduce of the fp64 sum (and of the cost image, whose
        entries are disjoint and must be zero on entry) is the only collective.  Returns the SUM over all pairs."""
        total = n_views * (n_views - 1) // 2
        bounds = self.c.direct_partition(self.world) if self.world > 1 else np.array([0, total])
        lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
        s = self.c.direct_evaluate_range(lo, hi, cost_image)
        if self.world > 1:
            import torch
            import torch.distributed as dist
            acc = torch.tensor([s], dtype=torch.float64, device=self.device)
            dist.all_reduce(acc, group=self.group)
            if cost_image is not None:
                dist.all_reduce(cost_image, group=self.group)
            s = float(acc.item())
        return s

    # -- batched mode (BASELINE config C4): K matrix sets against the same dtrs ----------------------------------------
    def evaluate_batch(self, Ps_sets, idx4=None):
        """Scores K complete projection-matrix sets (K, n, 12) against the dtrs every rank already holds: the sets are
        block-sharded over the ranks (a set is the unit of work of the optimiser loops, SURVEY.md section 8e "Batched
        mode"), each rank runs one batched launch on its sets, and the K means are exchanged (8 bytes per set -- the
        only traffic).  Returns all K means on every rank."""
        K = Ps_sets.shape[0]
        bounds = shard_bounds(K, self.world)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        mine = self.c.evaluate_batch(Ps_sets[lo:hi], idx4=idx4) if hi > lo else np.zeros(0, np.float64)
        return self._gather_means(mine, bounds)

    def evaluate_batch_params(self, base_Ps, params, view_to_param=None, idx4=None):
        """The same with PARAMETER VECTORS (K, m, 11) instead of matrices: each rank expands its block of sets to matrices on
        its own device (ecc_evaluate_batch_params); only the K means travel."""
        K = params.shape[0]
        bounds = shard_bounds(K, self.world)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        mine = (self.c.evaluate_batch_params(base_Ps, params[lo:hi], view_to_param=view_to_param, idx4=idx4)
                if hi > lo else np.zeros(0, np.float64))
        return self._gather_means(mine, bounds)

    def _gather_means(self, mine, bounds):
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        if self.world == 1:
            return np.asarray(mine, np.float64)
        import torch
        import torch.distributed as dist
        width = max(b - a for a, b in zip(bounds, bounds[1:]))
        buf = torch.zeros(width, dtype=torch.float64, device=self.device)
        buf[:hi - lo] = torch.as_tensor(np.asarray(mine, np.float64), device=self.device)
        parts = [torch.zeros_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        return np.concatenate([parts[r][:bounds[r + 1] - bounds[r]].cpu().numpy() for r in range(self.world)])
