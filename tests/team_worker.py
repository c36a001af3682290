"""Worker of tests/test_gpu_team.py: one rank of a multi-process team (one process per rank, all on the GPUs the box has;
with a single GPU the ranks share it and the blocks are still mapped through CUDA IPC).  Writes its results as .npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, port, out_dir, engine = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    mode = sys.argv[6] if len(sys.argv) > 6 else "steps"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port
    import torch
    import torch.distributed as dist
    from epipolarconsistency_b200 import api
    from epipolarconsistency_b200.distributed import ShardedPipeline, shard_bounds
    from team_scene import make_scene

    dev_index = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev_index)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = make_scene()
        n, n_u, n_v, n_a, n_t = S["n"], S["n_u"], S["n_v"], S["n_a"], S["n_t"]
        ctx = api.Context(dev_index)
        pipe = ShardedPipeline(ctx, rank, world, device=torch.device("cuda", dev_index), transport="team")
        interp = {"texture": api.INTERP_TEXTURE, "hybrid": api.INTERP_HYBRID, "hybrid-static": api.INTERP_HYBRID_STATIC}[engine]
        lo, hi, _ = pipe.radon_shard(n, n_a, n_t, interp)  # the static-split engine shards in quads of projections
        local = torch.from_numpy(S["imgs"][lo:hi]).cuda()
        ctx.set_interpolation(api.INTERP_TEXTURE)
        ctx.set_epipolar_plane_step(S["dkappa"])
        if mode == "loop":
            # optimiser loop: intermediates once, then evaluation after evaluation with other matrices and NO Radon barrier in
            # between; the ranks drift apart on purpose
            import time
            from test_gpu_team import loop_sets
            full = pipe.radon_allgather(local, n, n_a, n_t, interp=interp)
            ctx.set_radon_intermediates(full, n_u, n_v, True)
            sets = loop_sets(S)
            costs = [torch.zeros((n, n), dtype=torch.float32, device="cuda") for _ in sets]
            loop_means = []
            for k, P in enumerate(sets):
                if (k + rank) % 3 == 0:
                    time.sleep(0.05)
                ctx.set_projection_matrices(P)
                loop_means.append(pipe.evaluate_all_pairs(n, costs[k]))  # device cost image: only the stream orders its reads
            torch.cuda.synchronize()
            assert pipe._team_key is not None, f"team transport not used: {pipe.team_error}"
            np.savez(os.path.join(out_dir, f"loop{rank}.npz"), means=np.array(loop_means), costs=np.stack([c.cpu().numpy() for c in costs]))
            dist.barrier()
            return
        means = []
        for step in range(2):  # twice: the second step overwrites buffers the peers have read
            src = local if step == 0 else S["imgs"][lo:hi]  # device images, then host images
            full = pipe.radon_allgather(src, n, n_a, n_t, interp=interp)
            ctx.set_radon_intermediates(full, n_u, n_v, True)
            ctx.set_projection_matrices(S["Ps"])
            cost = torch.zeros((n, n), dtype=torch.float32, device="cuda")
            means.append(pipe.evaluate_all_pairs(n, cost))
        assert pipe._team_key is not None, f"team transport not used: {pipe.team_error}"
        cost_host = np.zeros((n, n), np.float32)
        mean_host = ctx.team_evaluate(cost_host)
        # batched mode: K matrix sets sharded over the ranks, the K means gathered (dtrs already on every rank)
        K = 5
        sets = np.stack([S["Ps"]] * K).reshape(K, n, 12).copy()
        for k in range(1, K):
            for c in range(4):  # set k: view k shifted by k pixels in u (row0 += k * row2, column-major 3x4)
                sets[k, k, 0 + 3 * c] += float(k) * sets[k, k, 2 + 3 * c]
        batch_means = pipe.evaluate_batch(sets)
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), dtrs=full.cpu().numpy(), cost=cost.cpu().numpy(), means=np.array(means),
                 cost_host=cost_host, mean_host=mean_host, batch_means=batch_means, batch_sets=sets)
        dist.barrier()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
