"""Development aid (torchrun, >= 2 GPUs): latency of the flag barrier in peer memory, and of publishing a rank's pair values."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from epipolarconsistency_b200 import api
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = api.Context(local)
h = ctx.team_create(rank, world, 16, 64, 64)
hs = [None] * world
dist.all_gather_object(hs, h)
ctx.team_connect(hs)
dist.barrier(); torch.cuda.synchronize()
for _ in range(20):
    ctx.team_barrier()
ctx.synchronize()
dist.barrier(); torch.cuda.synchronize()
n = 2000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ctx.team_barrier()
e1.record()
torch.cuda.synchronize()
us_stream = e0.elapsed_time(e1) * 1e3 / n
t0 = time.perf_counter()
for _ in range(200):
    ctx.team_barrier()
    ctx.synchronize()
us_host = (time.perf_counter() - t0) * 1e6 / 200
x = torch.zeros(8, device="cuda")
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(200):
    dist.all_reduce(x)
e1.record()
torch.cuda.synchronize()
us_nccl = e0.elapsed_time(e1) * 1e3 / 200
if rank == 0:
    print(f"world {world}: team barrier {us_stream:.2f} us back to back on the stream, {us_host:.1f} us launch + wait from the host; NCCL all_reduce of 32 bytes {us_nccl:.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
