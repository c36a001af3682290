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
    def __init__(self, compute, rank=0, world=1, device=None, group=None):
        self.c = compute
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        self._full = None

    # -- stage 1 + exchange -------------------------------------------------------------------------------
    def radon_allgather(self, local_images, n_total, n_alpha, n_t, **radon_kwargs):
        """local_images: this rank's block of projections (torch tensor on self.device, or pinned host tensor /
        numpy array for the end-to-end path).  Returns the full (n_total, n_t, n_alpha) dtr tensor."""
        import torch
        bounds = shard_bounds(n_total, self.world)
        lo, hi = bounds[self.rank], bounds[self.rank + 1]
        assert local_images.shape[0] == hi - lo, (local_images.shape, lo, hi)
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
