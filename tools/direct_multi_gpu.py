"""Direct metric (MetricDirect::evaluate) sharded over the GPUs of one node: pairs are independent, every rank holds the images,
the only collective is the all-reduce of the fp64 sum (ShardedPipeline.direct_evaluate).  Checks the sharded result against
one GPU's and times both.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/direct_multi_gpu.py [n_views]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api  # noqa: E402
from epipolarconsistency_b200.distributed import ShardedPipeline  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_u = n_v = 512
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4], [5, 30, 20, 12, 18, 9, 0.8]])
ctx = api.Context(local)
Ps = api.make_circular_trajectory(n, 750.0, 1200.0, n_u, n_v, 150.0, 0.6)
imgs = torch.empty((n, n_v, n_u), dtype=torch.float32, device="cuda")
ctx.synth_projections(Ps, n_u, n_v, ELL, imgs)
ctx.set_projection_matrices(Ps)
ctx.direct_set_images(imgs)
ctx.set_epipolar_plane_step(2.5e-4)  # a fixed step: the pairs differ in work, the partition has to follow the planes
pipe = ShardedPipeline(ctx, rank, world, device=torch.device("cuda", local))


def timed(fn, reps=3):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return out, float(ms.item())


cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")


def sharded():
    cost.zero_()
    return pipe.direct_evaluate(n, cost)


total, ms = timed(sharded)
bounds = ctx.direct_partition(world)
single_cost = np.zeros((n, n), np.float32)
single, ms1 = timed(lambda: ctx.direct_evaluate(single_cost))
same_image = bool(np.array_equal(cost.cpu().numpy(), single_cost))
if rank == 0:
    print(f"direct metric, {n} views {n_u}x{n_v}, {n * (n - 1) // 2} pairs, plane step 2.5e-4 rad, {world} GPU(s): "
          f"{ms:.2f} ms sharded (max over ranks, all-reduce of sum and cost image inside) vs {ms1:.2f} ms on one GPU -> {ms1 / ms:.2f}x; "
          f"sum {total:.10g} vs {single:.10g} (rel {abs(total - single) / single:.1e}), cost image identical: {same_image}; "
          f"pair ranges {bounds.tolist()}")
    assert same_image and abs(total - single) <= 1e-12 * single
if world > 1:
    dist.destroy_process_group()
