#!/usr/bin/env python
"""bench.py -- benchmarks of the ECC hot path on the BASELINE.json configurations.

Default (what the driver runs) is the headline configuration C3: full all-pairs ECC for 496 projections of 1240x960
(Radon intermediates 768x768, dkappa 0.01 deg, 200 deg short scan, 122 760 pairs).  `--workload` selects the others:
    c1  100 projections 512x512 -> 256x256, all pairs, dkappa auto, 360 deg        (pairs/s, Radon stage included)
    c2  Radon intermediates only, 496 x 1240x960 -> 768x768, sharded by projection  (Radon intermediates/s)
    c3  the headline                                                              (pairs/s, Radon stage included)
    c4  batched correction loop: 64 perturbed matrix sets x 248 projections/launch  (matrix sets/s)
    c5  tracking: 1 live view vs 400 reference views, {replace matrix, evaluate}     (calls/s, p50/p99 latency)

One STEP = one pass of the hot path over one batch of synthetic input (c1/c3: Radon intermediates of all projections ->
set matrices -> all-pairs ECC -> mean; c2: the Radon stage; c4: one batched launch of 64 sets; c5: 1000 tracking calls).
Every timed step gets projection matrices that differ from the previous step's (one double moved by one ulp), so that no
step is shortened by the library's caches for unchanged geometry (derived views, pair partition).
`value`   whole-job throughput with the inputs already resident in HBM.
`e2e`     the same through the public C ABI with HOST buffers: host->device copies of the step's inputs and the
          device->host read of its results are inside the timed region.
Usage:  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl b200|reference]
        torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)
--impl reference times the CPU restatement of the reference (oracle/, OpenMP over ALL host cores, whatever
OMP_NUM_THREADS says) on a bounded sample of the same workload; the reference ships no CPU implementation of this path
(BASELINE.md section 4).  At N=1 the own arm also reports `cpu_baseline` (the same oracle) and `ref_cuda`: the
reference's own CUDA kernels (oracle/_ref/libecc_ref_cuda.so, compiled unchanged for sm_100) timed on the same GPU in the
same run on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workloads (BASELINE.json configs) ---------------------------------------------
GEO = dict(sid=750.0, sdd=1200.0)
WORKLOADS = {
    "c1": dict(kind="pipeline", n=100, n_u=512, n_v=512, n_alpha=256, n_t=256, px=0.616, arc=360.0, dkappa_deg=0.0,
               name="C1: all-pairs ECC, 100 proj 512x512 -> 256x256 dtr, dkappa auto, 360 deg"),
    "c2": dict(kind="radon", n=496, n_u=1240, n_v=960, n_alpha=768, n_t=768, px=0.308, arc=200.0, dkappa_deg=0.01,
               name="C2: Radon intermediates only, 496 proj 1240x960 -> 768x768 (alpha,t) bins, sharded by projection"),
    "c3": dict(kind="pipeline", n=496, n_u=1240, n_v=960, n_alpha=768, n_t=768, px=0.308, arc=200.0, dkappa_deg=0.01,
               name="C3: all-pairs ECC, 496 proj 1240x960 -> 768x768 dtr, dkappa 0.01 deg, 200 deg arc"),
    "c4": dict(kind="batch", n=248, n_u=1240, n_v=960, n_alpha=768, n_t=768, px=0.308, arc=200.0, dkappa_deg=0.01, sets=64,
               name="C4: batched correction loop, 64 perturbed projection-matrix sets x 248 projections per launch, dkappa 0.01 deg"),
    "c5": dict(kind="tracking", n=401, n_u=1240, n_v=960, n_alpha=768, n_t=768, px=0.308, arc=200.0, dkappa_deg=0.01, calls=1000,
               name="C5: tracking, 1 live frame vs 400 reference projections, {replace one matrix, evaluate(400 listed pairs)} per call"),
    "tiny": dict(kind="pipeline", n=16, n_u=320, n_v=256, n_alpha=192, n_t=192, px=1.2, arc=200.0, dkappa_deg=0.05, name="tiny (debug)"),
    "tiny-radon": dict(kind="radon", n=16, n_u=320, n_v=256, n_alpha=192, n_t=192, px=1.2, arc=200.0, dkappa_deg=0.05, name="tiny Radon only (debug)"),
    "tiny-batch": dict(kind="batch", n=12, n_u=320, n_v=256, n_alpha=192, n_t=192, px=1.2, arc=200.0, dkappa_deg=0.05, sets=6, name="tiny batched (debug)"),
    "tiny-tracking": dict(kind="tracking", n=13, n_u=320, n_v=256, n_alpha=192, n_t=192, px=1.2, arc=200.0, dkappa_deg=0.05, calls=50, name="tiny tracking (debug)"),
}
METRICS = {  # kind -> (metric text, unit)
    "pipeline": ("all-pairs ECC image-pairs/s, end of Radon intermediates included (%s)", "pairs/s"),
    "radon": ("Radon intermediates/s (%s)", "intermediates/s"),
    "batch": ("perturbed projection-matrix sets scored per second, batched all-pairs ECC (%s)", "sets/s"),
    "tracking": ("tracking calls/s: replace one projection matrix + evaluate the listed pairs (%s)", "calls/s"),
}
# phantom: 5 ellipsoids, seed-1234 style fixed list (centre xyz, semi-axes xyz, density)
ELLIPSOIDS = np.array([
    [0.0, 0.0, 0.0, 80.0, 60.0, 70.0, 1.0],
    [20.0, -10.0, 5.0, 25.0, 30.0, 20.0, 0.6],
    [-25.0, 15.0, -10.0, 20.0, 22.0, 28.0, -0.5],
    [5.0, 30.0, 20.0, 22.0, 20.0, 24.0, 0.8],
    [-10.0, -30.0, -25.0, 30.0, 21.0, 20.0, -0.7],
])
DATA = "synthetic: analytic 5-ellipsoid phantom, circular cone-beam trajectory, cosine weighted"
MEASURED_PEAKS = os.path.join(ROOT, "MEASURED_PEAKS.json")
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
SM_COUNT = 148
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of `python bench.py` under ncu --set full, scaled to 128
# projections, by Radon engine (hybrid-static: profiles/ncu_radon_hybrid4_r02.txt, a 112-projection launch: 1728.5 + 287.3 MB;
# hybrid: profiles/ncu_radon_hybrid4_bench_r01b.txt); known for the C3 image size only
RADON_TRAFFIC_128 = {"hybrid-static": 2.3038e9, "hybrid": 1.4524e9}
# the same for one pair-kernel launch of a workload at its full size on one GPU (profiles/ncu_pairs_c3_r02.txt: 320.3 MB read;
# profiles/ncu_pairs_c4_r02_setsinner.txt: 399.4 MB read + 13.1 MB written); None = not captured
PAIRS_TRAFFIC = {"c3": 3.26e8, "c4": 4.125e8}


def hbm_peak():
    try:
        with open(MEASURED_PEAKS) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled every 20 ms through
    NVML (the nvidia-smi binary takes longer per call than an 8-GPU step lasts); nvidia-smi is the fall-back."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]  # nvmlClocksThrottleReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            if uuid is not None:
                for name in ("GPU-" + str(uuid), str(uuid)):
                    try:
                        handle = pynvml.nvmlDeviceGetHandleByUUID(name.encode())
                        break
                    except Exception:
                        handle = None
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = (pynvml, handle)
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            nv, h = self.nvml
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            return [sm, mx, power] + [bool(mask & b) for b in self.BITS]
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return None
        r = [x.strip() for x in out.split(",")]
        return [float(r[0]), float(r[1]), float(r[2])] + [x.lower().startswith("active") for x in r[3:7]]

    def run(self):
        while not self.stop_flag.is_set():
            try:
                row = self.sample()
                if row:
                    self.rows.append(row)
            except Exception:
                if self.nvml is not None:
                    self.nvml = None  # NVML call failed: try the binary from now on
            self.stop_flag.wait(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for k, n in enumerate(self.NAMES) if any(r[3 + k] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def ulp_variants(Ps, count):
    """`count` copies of the matrix set, copy k with ONE double of matrix (k mod n) moved up by k+1 ulps: byte-different from
    every other copy (the library's unchanged-geometry caches do not apply) and the same geometry to fp32."""
    out = []
    for k in range(count):
        P = Ps.copy()
        v = P[k % len(P), 11]
        for _ in range(k + 1):
            v = np.nextafter(v, np.inf)
        P[k % len(P), 11] = v
        out.append(P)
    return out


def perturbed_sets(api, Ps, K, rng, sigma_px=0.5, sigma_mm=0.5, sigma_deg=0.2):
    """K matrix sets P' = H2D P T3D (ModelCameraSimilarity2D3D, LibProjectiveGeometry/Models/ModelCameraSimilarity2D3D.hxx:89-92)
    with i.i.d. normal parameters per view: 0.5 px detector shifts, 0.2 deg rotations, 0.5 mm
    translations, no scaling.  Returns (sets (K, n, 12), parameters (K, n, 11)); set 0 is the unperturbed one."""
    n = len(Ps)
    x = np.zeros((K, n, 11))
    x[:, :, 0:2] = rng.normal(0, sigma_px, (K, n, 2))
    x[:, :, 2] = np.deg2rad(rng.normal(0, sigma_deg, (K, n)))
    x[:, :, 4:7] = rng.normal(0, sigma_mm, (K, n, 3))
    x[:, :, 7:10] = np.deg2rad(rng.normal(0, sigma_deg, (K, n, 3)))
    x[0] = 0.0
    sets = np.stack([np.stack([api.model_camera_similarity_2d3d(Ps[i], x[k, i]) for i in range(n)]) for k in range(K)])
    return sets, x


# ===========================================================================================================
# own arm
# ===========================================================================================================
class Bench:
    """Set-up shared by the workloads: one rank per GPU, synthetic data generated on the device, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from epipolarconsistency_b200 import api
        from epipolarconsistency_b200.distributed import ShardedPipeline, shard_bounds
        self.torch, self.dist, self.api, self.args = torch, dist, api, args
        self.W = WORKLOADS[args.workload]
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        W = self.W
        self.n, self.n_u, self.n_v, self.n_a, self.n_t = W["n"], W["n_u"], W["n_v"], W["n_alpha"], W["n_t"]
        self.ctx = api.Context(self.local_rank)  # bound to torch's current stream
        self.Ps = api.make_circular_trajectory(self.n, GEO["sid"], GEO["sdd"], self.n_u, self.n_v, W["arc"], W["px"])
        self.shard_bounds = shard_bounds
        self.pipe = ShardedPipeline(self.ctx, self.rank, self.world, device=self.dev, transport=args.exchange)
        self.ctx.set_interpolation(api.INTERP_TEXTURE)
        self.ctx.set_object_radius(0.0)
        self.ctx.set_epipolar_plane_step(float(np.deg2rad(W["dkappa_deg"])))
        self.radon_interp = {"hybrid": api.INTERP_HYBRID, "hybrid-static": api.INTERP_HYBRID_STATIC, "texture": api.INTERP_TEXTURE,
                             "exact": api.INTERP_EXACT}[args.radon]
        self.warmup = max(args.warmup, 3)  # the timing rules ask for at least three warm-up steps

    def synth(self, lo, hi):
        images = self.torch.empty((hi - lo, self.n_v, self.n_u), dtype=self.torch.float32, device=self.dev)
        if hi > lo:
            self.ctx.synth_projections(self.Ps[lo:hi], self.n_u, self.n_v, ELLIPSOIDS, images)
        return images

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps bracketed by barrier + synchronize, timed with CUDA events on the launching stream; max over ranks.
        fn(i) runs step i."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for i in range(steps):
            out = fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), out

    def profiled(self, fn, steps):
        """timed() with the per-kernel event profile and the clock sampler around it."""
        try:
            gpu_uuid = self.torch.cuda.get_device_properties(self.local_rank).uuid
        except Exception:
            gpu_uuid = None
        sampler = ClockSampler(self.local_rank, gpu_uuid)
        sampler.start()
        self.ctx.profile_reset()
        self.ctx.profile_enable(True)
        ms_total, out = self.timed(fn, steps)
        prof = {fam: self.ctx.profile_get(fam) for fam in ("radon", "pairs", "geometry", "reduce", "stage")}
        self.ctx.profile_enable(False)
        return ms_total, out, prof, sampler.summary()

    def line(self, value, unit_ms_per_step, scaling):
        metric, unit = METRICS[self.W["kind"]]
        return {"metric": metric % self.args.workload.upper(), "value": value, "unit": unit, "n_gpus": self.world,
                "steps": self.args.steps, "warmup": self.warmup, "ms_per_step": unit_ms_per_step, "higher_is_better": True,
                "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": DATA + "; generated on device"}

    def radon_roofline(self, prof, clocks, projections_per_step_this_rank):
        """Dominant kernel of c1/c2/c3.  Bound: the SM's two on-chip data pipes together -- 64 B/clk through the texture unit
        + 128 B/clk of shared memory per SM -- neither HBM (DRAM moves ~18 MB per projection) nor tensor cores (no
        contraction).  achieved = algorithmic tap bytes per launch (16 B per bilinear sample x samples per projection x
        projections per launch) / the kernel's mean launch time (CUDA events on the launching stream)."""
        args, W = self.args, self.W
        samples = self.ctx.radon_num_samples(self.n_u, self.n_v, self.n_a, self.n_t)
        ms, launches = prof["radon"]
        per_launch = projections_per_step_this_rank * args.steps / max(launches, 1)
        gbs = 16.0 * samples * per_launch / (ms / max(launches, 1) * 1e-3) / 1e9 if ms > 0 else 0.0
        hbm, hbm_src = hbm_peak()
        mhz = clocks.get("sm_mhz") or 1965.0
        hybrid = args.radon.startswith("hybrid")
        onchip = (192.0 if hybrid else 64.0) * SM_COUNT * mhz * 1e6 / 1e9
        big = (self.n_u, self.n_v, self.n_a, self.n_t) == (1240, 960, 768, 768)
        return {"bound": "onchip(tex+lsu)" if hybrid else "onchip(tex)",
                "kernel": "radon_hybrid4_kernel" if hybrid else "radon_kernel",
                "achieved": gbs, "peak": onchip, "unit": "GB/s", "frac": gbs / onchip,
                "peak_source": "%d B/clk/SM (%s) x %d SMs x the SM clock sampled during this run (%.0f MHz)"
                               % (192 if hybrid else 64, "texture data pipe 64 + shared-memory pipe 128" if hybrid else "texture data pipe", SM_COUNT, mhz),
                "hbm_peak": hbm, "hbm_peak_source": hbm_src, "hbm_frac": gbs / hbm,
                "traffic": (RADON_TRAFFIC_128[args.radon] / 128.0 * per_launch if (args.radon in RADON_TRAFFIC_128 and big) else None),
                "algorithmic_bytes_per_launch": 16.0 * samples * per_launch, "projections_per_launch": per_launch,
                "samples_per_s": gbs * 1e9 / 16.0, "tex_rate_frac": (gbs * 1e9 / 16.0) / 1.09e12,
                "note": "16 B of taps per bilinear sample x %.4g samples per projection, all of it on-chip traffic: hbm_frac > 1 "
                        "is expected, `traffic` is what DRAM really moves per launch; tex_rate_frac = samples/s over the texture "
                        "unit's measured rate alone (1.09e12/s at 1965 MHz, profiles/tex_probe_r01.txt)" % samples}

    def pairs_roofline(self, prof, kappa_samples_this_rank_per_launch, clocks=None):
        """Pair kernel: 64 B of taps per kappa sample (4 lookups x 4 taps x 4 B).  The taps are served by the
        texture unit out of L1TEX/L2 (ncu: two thirds of the sectors hit L1TEX), so the bound that applies is the texture data pipe,
        64 B/clk/SM -- the same on-chip roofline as the Radon kernel's texture path; the HBM copy peak and the measured
        random-sector rates of this GPU are side fields (DRAM moves a few hundred MB per launch)."""
        ms, launches = prof["pairs"]
        gbs = 64.0 * kappa_samples_this_rank_per_launch / (ms / max(launches, 1) * 1e-3) / 1e9 if ms > 0 else 0.0
        hbm, hbm_src = hbm_peak()
        mhz = (clocks or {}).get("sm_mhz") or 1965.0
        onchip = 64.0 * SM_COUNT * mhz * 1e6 / 1e9
        return {"bound": "onchip(tex)", "kernel": "pairs_kernel", "achieved": gbs, "peak": onchip, "unit": "GB/s", "frac": gbs / onchip,
                "peak_source": "64 B/clk/SM (texture data pipe) x %d SMs x the SM clock sampled during this run (%.0f MHz)" % (SM_COUNT, mhz),
                "hbm_peak": hbm, "hbm_peak_source": hbm_src, "hbm_frac": gbs / hbm,
                "kappa_samples_per_launch": float(kappa_samples_this_rank_per_launch),
                # random-gather rates of this pool's B200 (tools/gather_probe.cu, profiles/gather_probe_r01.txt): 32-byte
                # sectors from an L2-resident set / from a 1.17 GB set, bilinear texture fetches from 1.1 GB
                "gather_peaks_gbs": {"l2_random_sectors": 4380.0, "hbm_random_sectors": 1021.0, "texture_random_1GB": 427.0},
                "frac_of_l2_gather": gbs / 4380.0, "traffic": PAIRS_TRAFFIC.get(self.args.workload) if self.world == 1 else None,
                "note": "64 B of taps per kappa sample, served by the texture unit from L1TEX/L2: hbm_frac is not a bound (neighbouring "
                        "kappa samples share their sectors on chip), `traffic` is what DRAM moves per launch (ncu, profiles/INDEX.md); "
                        "the kernel waits on its dependent coordinate arithmetic and on first-touch misses, not on the data pipe "
                        "(DESIGN.md section 3.2)"}


def run_pipeline(B):
    """c1 / c3: Radon intermediates of all projections -> matrices -> all pairs -> mean."""
    torch, api, ctx, pipe, args, W = B.torch, B.api, B.ctx, B.pipe, B.args, B.W
    n, n_u, n_v, n_a, n_t, world, rank = B.n, B.n_u, B.n_v, B.n_a, B.n_t, B.world, B.rank
    lo, hi, part = pipe.radon_shard(n, n_a, n_t, interp=B.radon_interp)  # the static-split engine shards in quads of projections
    images = B.synth(lo, hi)  # this rank's projections, resident
    work = (hi - lo) if part is None else n / world           # projections' worth of Radon work on this rank
    uploaded = n if part is None else sum(ctx.team_radon_shard(n, world, r)[1] for r in range(world))  # e2e: images over all ranks
    images_host = torch.empty((hi - lo, n_v, n_u), dtype=torch.float32, pin_memory=True)
    images_host.copy_(images)
    cost_dev = torch.zeros((n, n), dtype=torch.float32, device=B.dev)
    cost_host = torch.zeros((n, n), dtype=torch.float32, pin_memory=True)
    variants = ulp_variants(B.Ps, B.warmup + 2 * args.steps + 2)
    counter = [0]

    def step(src_images, want_cost_on_host):
        Ps = variants[counter[0] % len(variants)]  # never the previous step's bytes
        counter[0] += 1
        full = pipe.radon_allgather(src_images, n, n_a, n_t, interp=B.radon_interp)
        ctx.set_radon_intermediates(full, n_u, n_v, True)
        ctx.set_projection_matrices(Ps)
        cost_dev.zero_()
        mean = pipe.evaluate_all_pairs(n, cost_dev)
        if want_cost_on_host:
            cost_host.copy_(cost_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return mean

    for _ in range(B.warmup):
        step(images, False)
    ms_total, mean, prof, clocks = B.profiled(lambda i: step(images, False), args.steps)
    step(images_host, True)
    ms_e2e, mean_e2e = B.timed(lambda i: step(images_host, True), args.steps)

    n_pairs = n * (n - 1) // 2
    ms_step, ms_step_e2e = ms_total / args.steps, ms_e2e / args.steps
    counts = ctx.pair_sample_counts(n)
    my_lo, my_hi = (pipe.c.partition_pairs(world)[rank:rank + 2] if world > 1 else (0, n_pairs))
    radon_ms, pairs_ms = prof["radon"][0], prof["pairs"][0]
    if rank == 0:
        line = B.line(n_pairs / (ms_step * 1e-3), ms_step, "strong")
        line["config"] = {
            "workload": W["name"], "projections": n, "pairs": n_pairs,
            "interpolation": {"hybrid": "texture-filter arithmetic (reference CUDA numerics); Radon samples split between the texture unit and a shared-memory path with the same 1.8 fixed-point weights",
                              "hybrid-static": "as hybrid, with a fixed (geometry-only) assignment of bins to the two paths: bit-reproducible, independent of batching and sharding",
                              "texture": "texture unit (reference CUDA numerics, bit-identical Radon bins)",
                              "exact": "Radon with exact fp32 weights; metric through the texture unit"}[args.radon],
            "sharding": (f"projections block-sharded over {world} GPU(s)" if part is None else
                         f"quads of four projections cut into {world} equal intervals, the quad on a boundary shared by its two ranks "
                         f"(rank 0: projections {lo}..{hi - 1}, share {part[0]}/{part[2]}..{part[1]}/{part[2]} of its first / last quad)")
                        + ", pairs partitioned by equal kappa samples",
            "exchange": ("none (1 GPU)" if world == 1 else
                         "peer stores: the Radon kernels write every bin into all ranks' buffers over NVLink, pair values published "
                         "the same way, flag barriers in peer memory; no collective on the data path" if pipe._team_key is not None else
                         "NCCL all-gather of the dtr blocks + all-reduce of cost image and sum"
                         + (f" (team transport unavailable: {pipe.team_error})" if pipe.team_error else "")),
            "matrices": "every step gets a matrix set that differs from the previous step's by one ulp in one entry (no cached derivation / partition)",
            "l2": "inputs larger than L2 (%.2f GB images + %.2f GB dtrs per step)" % (n * n_u * n_v * 4 / 1e9, n * n_a * n_t * 4 / 1e9)}
        line["stages"] = {"radon_intermediates_per_s": world * work / ((radon_ms / args.steps) * 1e-3) if radon_ms > 0 else None,
                          "radon_kernel_ms_per_step_rank0": radon_ms / args.steps,
                          "pairs_per_s_metric_only_rank0": float(my_hi - my_lo) / ((pairs_ms / args.steps) * 1e-3) if pairs_ms > 0 else None,
                          "pair_kernel_ms_per_step_rank0": pairs_ms / args.steps,
                          "kernel_share_of_step": {k: v[0] / ms_total for k, v in prof.items()}, "mean_ecc": mean}
        line["e2e"] = {"value": n_pairs / (ms_step_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_step_e2e,
                       "h2d_bytes_per_step": int(uploaded) * n_u * n_v * 4 + n * 96, "d2h_bytes_per_step": n * n * 4 + 8 * world,
                       "mean_ecc": mean_e2e}
        line["gpu_launches"] = int(sum(v[1] for v in prof.values()))
        line["clocks"] = clocks
        line["roofline"] = B.radon_roofline(prof, clocks, work)
        line["roofline_pairs"] = B.pairs_roofline(prof, float(counts[int(my_lo):int(my_hi)].sum()), clocks)
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(W, B.Ps, full_dtrs=pipe._full)
            line["ref_cuda"] = ref_cuda_leg(B, images, pipe._full, radon_ms / args.steps / max(work, 1), pairs_ms / args.steps, ms_step)
        print(json.dumps(line))


def run_radon(B):
    """c2: the Radon stage alone, sharded by projection; on several GPUs every rank ends with all intermediates."""
    torch, api, ctx, pipe, args, W = B.torch, B.api, B.ctx, B.pipe, B.args, B.W
    n, n_u, n_v, n_a, n_t, world, rank = B.n, B.n_u, B.n_v, B.n_a, B.n_t, B.world, B.rank
    lo, hi, part = pipe.radon_shard(n, n_a, n_t, interp=B.radon_interp)
    work = (hi - lo) if part is None else n / world
    uploaded = n if part is None else sum(ctx.team_radon_shard(n, world, r)[1] for r in range(world))
    images = B.synth(lo, hi)
    images_host = torch.empty((hi - lo, n_v, n_u), dtype=torch.float32, pin_memory=True)
    images_host.copy_(images)
    dtrs_host = torch.empty((hi - lo, n_t, n_a), dtype=torch.float32, pin_memory=True)

    def step(src, to_host):
        if to_host and world == 1:
            # host images in, host intermediates out (what the ComputeRadonIntermediate tool does): one call; uploads and
            # downloads run under the kernels of the neighbouring chunks
            ctx.radon_compute(src, n_a, n_t, out=dtrs_host, interp=B.radon_interp)
            return pipe._full
        full = pipe.radon_allgather(src, n, n_a, n_t, interp=B.radon_interp)
        if to_host:  # the step's result: this rank's intermediates, read back
            dtrs_host.copy_(full[lo:hi], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return full

    for _ in range(B.warmup):
        step(images, False)
    ms_total, full, prof, clocks = B.profiled(lambda i: step(images, False), args.steps)
    step(images_host, True)
    ms_e2e, _ = B.timed(lambda i: step(images_host, True), args.steps)
    ms_step, ms_step_e2e = ms_total / args.steps, ms_e2e / args.steps
    if rank == 0:
        line = B.line(n / (ms_step * 1e-3), ms_step, "strong")
        line["config"] = {"workload": W["name"], "projections": n, "engine": args.radon,
                          "sharding": f"projections block-sharded over {world} GPU(s); every rank ends with all intermediates"
                                      + ("" if world == 1 else " (peer stores from inside the kernel)" if pipe._team_key is not None else " (NCCL all-gather)"),
                          "l2": "inputs larger than L2 (%.2f GB images, %.2f GB dtrs per step)" % (n * n_u * n_v * 4 / 1e9, n * n_a * n_t * 4 / 1e9)}
        line["stages"] = {"radon_kernel_ms_per_step_rank0": prof["radon"][0] / args.steps,
                          "ms_per_projection_rank0": prof["radon"][0] / args.steps / max(work, 1),
                          "kernel_share_of_step": {k: v[0] / ms_total for k, v in prof.items()},
                          "checksum": float(full[lo:hi].double().abs().sum().item())}
        line["e2e"] = {"value": n / (ms_step_e2e * 1e-3), "unit": "intermediates/s", "ms_per_step": ms_step_e2e,
                       "h2d_bytes_per_step": int(uploaded) * n_u * n_v * 4, "d2h_bytes_per_step": int(uploaded) * n_a * n_t * 4}
        line["gpu_launches"] = int(sum(v[1] for v in prof.values()))
        line["clocks"] = clocks
        line["roofline"] = B.radon_roofline(prof, clocks, work)
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(W, B.Ps, full_dtrs=None, radon_only=True)
            line["ref_cuda"] = ref_cuda_leg(B, images, None, prof["radon"][0] / args.steps / max(work, 1), None, ms_step)
        print(json.dumps(line))


def run_batch(B):
    """c4: K perturbed matrix sets per launch against intermediates that stay resident (the NLopt correction loops compute
    them once); sets are block-sharded over the ranks, the K means are exchanged."""
    torch, api, ctx, pipe, args, W = B.torch, B.api, B.ctx, B.pipe, B.args, B.W
    n, n_u, n_v, n_a, n_t, world, rank, K = B.n, B.n_u, B.n_v, B.n_a, B.n_t, B.world, B.rank, B.W["sets"]
    images = B.synth(0, n)
    dtrs = ctx.radon_compute(images, n_a, n_t, interp=B.radon_interp)
    del images
    ctx.set_projection_matrices(B.Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    rng = np.random.default_rng(42)
    batches = [perturbed_sets(api, B.Ps, K, rng) for _ in range(3)]  # a new batch every step, as an optimiser would send
    sb = B.shard_bounds(K, world)
    lo, hi = sb[rank], sb[rank + 1]
    resident = [torch.from_numpy(np.ascontiguousarray(b[0][lo:hi])).to(B.dev) for b in batches]  # this rank's sets, in HBM
    params = [np.ascontiguousarray(b[1]) for b in batches]
    use_params = hasattr(ctx, "evaluate_batch_params") and not args.batch_matrices

    def step_resident(i):
        return ctx.evaluate_batch(resident[i % 3]) if hi > lo else None

    def step_host(i):
        if use_params:  # K x n parameter vectors in, expanded to matrices on the device
            return pipe.evaluate_batch_params(B.Ps, params[i % 3])
        return pipe.evaluate_batch(batches[i % 3][0])

    for i in range(B.warmup):
        step_resident(i)
    ms_total, _, prof, clocks = B.profiled(step_resident, args.steps)
    means = step_host(0)
    ms_e2e, means_e2e = B.timed(step_host, args.steps)
    ms_step, ms_step_e2e = ms_total / args.steps, ms_e2e / args.steps
    pairs = n * (n - 1) // 2
    ctx.set_projection_matrices(B.Ps)
    counts = ctx.pair_sample_counts(n)
    if rank == 0:
        line = B.line(K / (ms_step * 1e-3), ms_step, "strong")
        line["config"] = {"workload": W["name"], "projections": n, "sets_per_launch": K, "pairs_per_set": pairs,
                          "perturbation": "ModelCameraSimilarity2D3D per view, i.i.d. N(0, 0.5 px / 0.5 mm / 0.2 deg), seed 42; set 0 unperturbed; a new batch every step",
                          "sharding": f"matrix sets block-sharded over {world} GPU(s), intermediates replicated, means exchanged",
                          "e2e_input": "parameter vectors (11 doubles per view), expanded to matrices on the device" if use_params else "matrices (12 doubles per view) in host memory",
                          "l2": "intermediates larger than L2 (%.2f GB)" % (n * n_a * n_t * 4 / 1e9)}
        line["stages"] = {"pairs_per_s": K * pairs / (ms_step * 1e-3), "pair_kernel_ms_per_launch_rank0": prof["pairs"][0] / max(prof["pairs"][1], 1),
                          "kernel_share_of_step": {k: v[0] / ms_total for k, v in prof.items()},
                          "mean_unperturbed": float(means[0]), "mean_perturbed_min": float(np.min(means[1:])), "mean_perturbed_max": float(np.max(means[1:]))}
        line["e2e"] = {"value": K / (ms_step_e2e * 1e-3), "unit": "sets/s", "ms_per_step": ms_step_e2e,
                       "h2d_bytes_per_step": K * n * (88 if use_params else 96) + (n * 96 if use_params else 0), "d2h_bytes_per_step": 8 * K,
                       "mean_unperturbed": float(means_e2e[0])}
        line["gpu_launches"] = int(sum(v[1] for v in prof.values()))
        line["clocks"] = clocks
        # kappa samples of a set vary little with the perturbation: counted on the unperturbed set
        line["roofline"] = B.pairs_roofline(prof, float(counts.sum()) * (hi - lo), clocks)
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(W, B.Ps, full_dtrs=dtrs, pairs_only=True, per="set")
        print(json.dumps(line))


def run_tracking(B):
    """c5: one live view against n-1 reference views; a call = {replace the live view's matrix, evaluate the listed pairs}
    through ecc_update_and_evaluate (host matrix in, mean out).  Too small to shard: N ranks run N independent replicas."""
    torch, api, ctx, args, W = B.torch, B.api, B.ctx, B.args, B.W
    n, n_u, n_v, n_a, n_t, world, rank, calls = B.n, B.n_u, B.n_v, B.n_a, B.n_t, B.world, B.rank, B.W["calls"]
    images = B.synth(0, n)
    dtrs = ctx.radon_compute(images, n_a, n_t, interp=B.radon_interp)
    del images
    ctx.set_projection_matrices(B.Ps)
    ctx.set_radon_intermediates(dtrs, n_u, n_v, True)
    live = n - 1
    idx = np.array([(live, i, live, i) for i in range(live)], np.int32)  # Gui/SingleImageMotion.h:39-41
    rng = np.random.default_rng(7)
    x = np.zeros((256, 11))
    x[:, 0:2] = rng.normal(0, 0.5, (256, 2))
    x[:, 2] = np.deg2rad(rng.normal(0, 0.2, 256))
    x[:, 4:7] = rng.normal(0, 0.5, (256, 3))
    x[:, 7:10] = np.deg2rad(rng.normal(0, 0.2, (256, 3)))
    poses = [api.camera_similarity_2d3d(B.Ps[live], xi) for xi in x]
    out = np.zeros(live, np.float32)

    def step(i):
        m = 0.0
        for k in range(calls):
            m = ctx.update_and_evaluate(live, poses[(i * calls + k) % 256], idx, out)
        return m

    for i in range(B.warmup):
        step(i)
    # device clock around K steps; the per-kernel event profile is off here: it would force the plain (non-graph) path
    try:
        gpu_uuid = torch.cuda.get_device_properties(B.local_rank).uuid
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(B.local_rank, gpu_uuid)
    sampler.start()
    ms_total, mean = B.timed(step, args.steps)
    clocks = sampler.summary()
    # host clock per call: latency distribution (the call synchronises: mean and values are on the host when it returns)
    lat = []
    B.barrier()
    t_all = time.perf_counter()
    for k in range(args.steps * calls):
        t0 = time.perf_counter()
        ctx.update_and_evaluate(live, poses[k % 256], idx, out)
        lat.append(time.perf_counter() - t0)
    wall = time.perf_counter() - t_all
    lat = np.array(lat) * 1e6
    kernels_per_call, _ = ctx.track_info()
    # the pair kernel alone, from the plain path under the event profile (same kernel, same launch shape)
    ctx.profile_reset()
    ctx.profile_enable(True)
    for k in range(50):
        ctx.update_projection_matrix(live, poses[k])
        ctx.evaluate_indices(idx)
    prof = {fam: ctx.profile_get(fam) for fam in ("radon", "pairs", "geometry", "reduce", "stage")}
    ctx.profile_enable(False)
    ctx.set_projection_matrices(B.Ps)
    counts = ctx.pair_sample_counts(n)
    live_samples = float(sum(counts[api_pair_index(i, live, n)] for i in range(live)))
    ms_step = ms_total / args.steps
    if rank == 0:
        line = B.line(world * calls / (ms_step * 1e-3), ms_step, "weak")
        line["config"] = {"workload": W["name"], "reference_views": live, "pairs_per_call": live, "calls_per_step": calls,
                          "path": "ecc_update_and_evaluate: one recorded CUDA graph per call, " + (
                              "{80-byte view upload, ONE kernel: pairs, sums by its last CTA, values + sum + flag into pinned host memory}" if kernels_per_call == 1
                              else "{64-byte view upload, pair kernel, finalize + sum into pinned host memory}"),
                          "sharding": "replicas only (400 pairs are too few to shard): %d independent replica(s)" % world,
                          "l2": "intermediates larger than L2 (%.2f GB); a call touches the live view's and all reference views' intermediates" % (n * n_a * n_t * 4 / 1e9)}
        line["stages"] = {"pair_kernel_us_plain_path": 1e3 * prof["pairs"][0] / max(prof["pairs"][1], 1), "mean_ecc_last_call": mean,
                          "latency_us": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)), "mean": float(lat.mean()), "calls": len(lat)}}
        line["e2e"] = {"value": world * len(lat) / wall, "unit": "calls/s", "ms_per_step": wall * 1e3 / args.steps,
                       "h2d_bytes_per_step": calls * 64, "d2h_bytes_per_step": calls * (8 + 4 * live),
                       "note": "host clock around the same calls: a call takes its matrix from host memory and returns mean and values to the host"}
        line["gpu_launches"] = int(kernels_per_call * calls * args.steps)  # kernel nodes of the replayed graph (ecc_track_info)
        line["clocks"] = clocks
        line["roofline"] = B.pairs_roofline(prof, live_samples, clocks)
        if args.cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(W, B.Ps, full_dtrs=dtrs, pairs_only=True, per="call", idx=idx)
        print(json.dumps(line))


def api_pair_index(i, j, n):
    return i * (2 * n - i - 1) // 2 + (j - i - 1)


def run_b200(args):
    B = Bench(args)
    {"pipeline": run_pipeline, "radon": run_radon, "batch": run_batch, "tracking": run_tracking}[B.W["kind"]](B)
    if B.world > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


# ===========================================================================================================
# the reference's own CUDA kernels on the same GPU, same run (oracle/_ref/libecc_ref_cuda.so: the reference's .cu files
# compiled unchanged for sm_100, driven by oracle/ref_cuda_harness.cu) -- SURVEY.md section 2b's bar
# ===========================================================================================================
def ref_cuda_leg(B, images, full_dtrs, our_ms_per_projection, our_pairs_ms, our_ms_step, n_radon=8):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import oracle_lib as ol
        if ol.ref_cuda() is None or not hasattr(ol.ref_cuda(), "ref_cuda_radon_any"):
            return {"unavailable": "oracle/_ref/libecc_ref_cuda.so not present (built from /root/reference by oracle/Makefile)"}
        torch = B.torch
        n_radon = min(n_radon, images.shape[0])
        ol.ref_cuda_radon(images[:2].contiguous(), B.n_a, B.n_t)  # warm-up
        _, ms = ol.ref_cuda_radon(images[:n_radon].contiguous(), B.n_a, B.n_t)
        radon_ms = ms / n_radon
        out = {"kind": "reference CUDA kernels (RadonIntermediate.cu:31-170, EpipolarConsistencyRadonIntermediate.cu:13-409), sm_100 build, same GPU, same run",
               "radon_ms_per_projection": radon_ms, "radon_sample": f"{n_radon} of {B.n} projections, CUDA events around computeDerivLineIntegrals",
               "vs_ref_cuda": {"radon": radon_ms / our_ms_per_projection if our_ms_per_projection else None}}
        if full_dtrs is not None:
            M = ol.RefCudaMetric(B.Ps, full_dtrs, B.n_u, B.n_v)
            radius = B.ctx.get_object_radius()
            dk = float(np.deg2rad(B.W["dkappa_deg"]))
            runs = [M.evaluate(radius, dk)[2] for _ in range(4)]
            M.close()
            torch.cuda.synchronize()
            pairs_ms = float(min(runs[1:]))
            est = B.n * radon_ms + pairs_ms
            out.update({"pairs_ms": pairs_ms, "pairs_sample": f"all {B.n * (B.n - 1) // 2} pairs, CUDA events around epipolarConsistency(), best of 3",
                        "ms_per_step_estimate": est, "estimate": f"{B.n} x radon_ms_per_projection + pairs_ms (the reference computes one projection per launch)"})
            out["vs_ref_cuda"].update({"pairs": pairs_ms / our_pairs_ms if our_pairs_ms else None, "step": est / our_ms_step})
        return out
    except Exception as e:  # noqa: BLE001 -- a reported leg must not take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"}


# ===========================================================================================================
# CPU baseline / reference arm: the oracle (CPU restatement), timed on a bounded sample of the same workload
# ===========================================================================================================
def cpu_sample(W, Ps, dtrs_host, n_radon=1, n_pairs_sample=1500, seed=5, idx=None):
    """Times the oracle on n_radon full-size projections and n_pairs_sample random pairs of the enumeration (or of `idx`).
    Returns (seconds radon per projection, seconds per pair, cores, pairs timed)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    n, n_u, n_v, n_a, n_t = W["n"], W["n_u"], W["n_v"], W["n_alpha"], W["n_t"]
    cores = ol.use_all_host_cores()  # torchrun exports OMP_NUM_THREADS=1 to its ranks: do not time the CPU arm on one core
    imgs = [ol.project_ellipsoids(Ps[k], n_u, n_v, ELLIPSOIDS) for k in range(max(n_radon, 1))]
    t_radon = 0.0
    if n_radon > 0:
        t0 = time.perf_counter()
        for im in imgs:
            ol.radon(im, n_a, n_t, interp=ol.INTERP_EXACT)
        t_radon = (time.perf_counter() - t0) / n_radon
    if n_pairs_sample <= 0:
        return t_radon, 0.0, cores, 0
    rng = np.random.default_rng(seed)
    if idx is None:
        total = n * (n - 1) // 2
        ks = rng.choice(total, size=min(n_pairs_sample, total), replace=False)
        idx = np.array([ol.get_ij(int(k), n) * 2 for k in ks], np.int32)  # (P0, P1, dtr0, dtr1)
    else:
        idx = np.array(idx[rng.choice(len(idx), size=min(n_pairs_sample, len(idx)), replace=False)], np.int32)
    if dtrs_host is None:
        # no GPU-computed dtrs at hand (reference arm): 64 stand-in dtrs (scaled copies of one oracle dtr, 151 MB at
        # C3) addressed modulo 64, so that the memory footprint of the lookups is realistic; timing is data independent
        base = ol.radon(imgs[0], n_a, n_t, interp=ol.INTERP_EXACT)
        m = min(64, n)
        dtrs_host = np.stack([base * np.float32(1.0 + 0.01 * k) for k in range(m)])
        idx[:, 2] %= m
        idx[:, 3] %= m
    t0 = time.perf_counter()
    ol.ecc(Ps, dtrs_host, n_u, n_v, dkappa=float(np.deg2rad(W["dkappa_deg"])), interp=ol.INTERP_EXACT, idx4=idx, want_out=False)
    t_pair = (time.perf_counter() - t0) / len(idx)
    return t_radon, t_pair, cores, len(idx)


def cpu_extrapolate(W, t_radon, t_pair, idx_len=None):
    """(value, unit, seconds per unit of the metric, text) for the workload's metric from the two sampled rates."""
    n = W["n"]
    pairs = n * (n - 1) // 2
    kind = W["kind"]
    if kind == "radon":
        return 1.0 / t_radon, "intermediates/s", n * t_radon, f"{n} x Radon"
    if kind == "batch":
        per_set = pairs * t_pair
        return 1.0 / per_set, "sets/s", W["sets"] * per_set, f"{pairs} pairs per set (intermediates resident, as in the GPU arm)"
    if kind == "tracking":
        per_call = (n - 1) * t_pair
        return 1.0 / per_call, "calls/s", W["calls"] * per_call, f"{n - 1} pairs per call"
    est = n * t_radon + pairs * t_pair
    return pairs / est, "pairs/s", est, f"{n} x Radon + {pairs} pairs"


def cpu_baseline(W, Ps, full_dtrs=None, radon_only=False, pairs_only=False, per=None, idx=None):
    dtrs_host = full_dtrs.cpu().numpy() if full_dtrs is not None else None
    t_radon, t_pair, cores, k = cpu_sample(W, Ps, dtrs_host, n_radon=0 if pairs_only else 1, n_pairs_sample=0 if radon_only else 1500, idx=idx)
    value, unit, est, what = cpu_extrapolate(W, t_radon, t_pair)
    return {"value": value, "unit": unit, "cores": cores, "kind": "port",
            "sample": f"oracle (CPU restatement, OpenMP, {cores} threads): "
                      + ("" if pairs_only else f"Radon of 1 of {W['n']} projections ({t_radon:.2f} s)")
                      + ("" if radon_only or pairs_only else " + ")
                      + ("" if radon_only else f"{k} pairs ({t_pair * 1e3:.3f} ms/pair)")
                      + f", scaled linearly to one step = {what} ({est:.1f} s)",
            "radon_intermediates_per_s": (1.0 / t_radon) if t_radon > 0 else None,
            "pairs_per_s_metric_only": (1.0 / t_pair) if t_pair > 0 else None}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    n = W["n"]
    kind = W["kind"]
    Ps = ol.circular_trajectory(n, GEO["sid"], GEO["sdd"], W["n_u"], W["n_v"], W["arc"], W["px"])
    idx = np.array([(n - 1, i, n - 1, i) for i in range(n - 1)], np.int32) if kind == "tracking" else None
    n_radon = 0 if kind in ("batch", "tracking") else 1  # their GPU arms keep the intermediates resident, too
    n_pairs_sample = 0 if kind == "radon" else 1500
    for _ in range(min(args.warmup, 1)):
        cpu_sample(W, Ps, None, n_radon=n_radon, n_pairs_sample=min(200, n_pairs_sample), idx=idx)
    t0 = time.perf_counter()
    acc = []
    for _ in range(args.steps):
        acc.append(cpu_sample(W, Ps, None, n_radon=n_radon, n_pairs_sample=n_pairs_sample, idx=idx))
    wall = time.perf_counter() - t0
    t_radon = float(np.mean([a[0] for a in acc]))
    t_pair = float(np.mean([a[1] for a in acc]))
    cores, k = acc[0][2], acc[0][3]
    value, unit, est, what = cpu_extrapolate(W, t_radon, t_pair)
    metric, _ = METRICS[kind]
    line = {
        "impl": "reference", "metric": metric % args.workload.upper(),
        "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": est * 1e3, "higher_is_better": True, "scaling": "weak" if kind == "tracking" else "strong", "vs_baseline": None, "dtype": "f32",
        "data": DATA,
        "config": {"workload": W["name"], "projections": n, "pairs": n * (n - 1) // 2, "interpolation": "exact fp32 (CPU float path)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": f"per step: Radon of {n_radon} of {n} projections + {k} pairs on {cores} host threads (all cores this "
                                   f"process may use, whatever OMP_NUM_THREADS says), scaled linearly to one step = {what}; measured "
                                   f"{t_radon:.2f} s/projection, {t_pair * 1e3:.3f} ms/pair; wall time of the {args.steps} sampled steps {wall:.1f} s"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference ships no CPU implementation of this path (its Radon transform exists only as a CUDA kernel); "
                "this arm times oracle/ = the CPU restatement pinned against the reference's own headers and CUDA kernels",
    }
    print(json.dumps(line))


def main():
    # stdout carries exactly one JSON line: anything native libraries write to file descriptor 1 (NCCL prints its
    # version there) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false",
                    help="skip the cpu_baseline and ref_cuda legs (N=1 only anyway)")
    ap.add_argument("--exchange", default="team", choices=["team", "nccl"],
                    help="multi-GPU transport: team (peer stores from inside the kernels + flag barriers, default) or nccl (collectives after the kernels)")
    ap.add_argument("--radon", default="hybrid-static", choices=["hybrid", "hybrid-static", "texture", "exact"],
                    help="Radon engine: hybrid-static (texture unit + shared-memory path, fixed split: bit-reproducible and independent of the "
                         "number of GPUs; default), hybrid (run-time work queue between the two paths), texture (bit-identical to the reference "
                         "kernel), exact")
    ap.add_argument("--batch-matrices", action="store_true", help="c4 e2e: send matrices instead of parameter vectors")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
