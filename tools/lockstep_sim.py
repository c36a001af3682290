"""Development aid (CPU only): how many iterations of the quad Radon kernel's window loop are spent with lanes switched off?

A window warp walks its 32 bins in lock step, chunk by chunk along the primary axis; a lane stops at the chunk edge (the
window buffer ends there), so a chunk costs the warp as many iterations as its SLOWEST lane needs, and LDS.128 costs its
four wavefront cycles whatever the active mask (profiles/lds_mask_probe_r02.txt).  Uses tools/bank_sim.py's replay of the
sample positions (same clipping, stepping and chunking; shipped lane tiling 4 angles x 8 t) on random work items and prints
    lock/mean    iterations of the lock-step loop / mean iterations of a lane           (what the kernel pays)
    merged/mean  the same if a lane could run on into the next chunk's buffer (bound)  (what a restructured loop could reach)
Result at the C2/C3 geometry (150 items): chunk 10 (shipped) 1.040 / 1.017, chunk 16 1.032 / 1.017, chunk 20 1.028 / 1.017:
4 % of the window path's shared-memory wavefronts go to switched-off lanes, at most 2.3 % of them are recoverable (the lanes
of a warp are four neighbouring angles whose lines progress at slightly different rates along the primary axis).
Usage: python tools/lockstep_sim.py [items]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bank_sim as bs

n_u, n_v, n_alpha, n_t = 1240, 960, 768, 768
items = int(sys.argv[1]) if len(sys.argv) > 1 else 150
fn = list(bs.LANE_MAPS.values())[0]  # the shipped tiling
for chunk in (10, 16, 20):
    rng = np.random.default_rng(1)
    lane_mean = lock = merged = 0.0
    for _ in range(items):
        ag, tg = int(rng.integers(n_alpha // 8)), int(rng.integers(n_t // 32))
        cells = bs.item_cells(n_u, n_v, n_alpha, n_t, ag, tg, chunk, fn)
        for w in range(8):
            per_chunk, totals = {}, []
            for lane in range(32):
                c = cells[w * 32 + lane]
                if c is None:
                    totals.append(0)
                    continue
                j = c[0]
                u, cnt = np.unique(j, return_counts=True)
                totals.append(len(j))
                for a, b in zip(u, cnt):
                    per_chunk.setdefault(int(a), []).append(int(b))
            lock += sum(max(v) for v in per_chunk.values())
            lane_mean += sum(totals) / 32.0
            merged += max(totals)
    print(f"chunk {chunk:2d}: lock/mean {lock / lane_mean:.4f}   merged/mean {merged / lane_mean:.4f}   ({items} items)")
