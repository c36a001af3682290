"""Development aid (CPU only): shared-memory bank-conflict model of the quad Radon kernel's window path
(epipolarconsistency_b200/csrc/ecc_radon_hybrid4.cu), used to choose the lane -> bin tiling and the window layout before
spending GPU time.

For a sample of work items (8 angles x 32 t bins) of a bin geometry it replays which window cell every lane reads at every
step of the lock-step sample loop (same clipping, same 0.66 px step, same chunking along the primary axis; fp64 instead of
the kernel's fp32, which moves no cell that matters for statistics) and counts, per LDS.128 and quarter-warp, the
wavefronts the access takes: the number of distinct 16-byte cells that fall into the same bank group (cell index mod 8).
ncu of the shipped configuration (lane map 2, 137 rows, profiles/ncu_radon_hybrid4_fine_bench_r01.txt) reports 1.43
wavefronts per conflict-free one; the model's number for the same configuration validates it.

Usage: python tools/bank_sim.py [--items 120] [--rows 137] [--chunk 10]
"""
import argparse
import itertools

import numpy as np

STEP = np.float32(0.66)


def bin_lines(ix, iy, n_alpha, n_t, n_u, n_v):
    """Vectorised bin_line (ecc_radon_common.cuh): arrays o0, o1, d0, d1, t0, t1, valid."""
    x_rel = ix / n_alpha - 0.5
    y_rel = iy / n_t - 0.5
    diag = np.sqrt(n_u * n_u + n_v * n_v)
    alpha = x_rel * np.pi
    tau = y_rel * diag
    l0, l1 = -np.sin(alpha), np.cos(alpha)
    l2 = -tau - 0.5 * n_u * l0 - 0.5 * n_v * l1
    o0, o1 = -l2 * l0, -l2 * l1
    d0, d1 = l1, -l0
    with np.errstate(divide="ignore", invalid="ignore"):
        ta, tb = (1.0 - o0) / d0, (n_u - 1.0 - o0) / d0
        tc, td = (1.0 - o1) / d1, (n_v - 1.0 - o1) / d1
    small0, small1 = d0 * d0 < 1e-12, d1 * d1 < 1e-12
    ta = np.where(small0, -1e10, ta); tb = np.where(small0, 1e10, tb)
    tc = np.where(small1, -1e10, tc); td = np.where(small1, 1e10, td)
    lo1, hi1 = np.minimum(ta, tb), np.maximum(ta, tb)
    lo2, hi2 = np.minimum(tc, td), np.maximum(tc, td)
    t0, t1 = np.maximum(lo1, lo2), np.minimum(hi1, hi2)
    swapped = t1 < t0
    pu, pv = o0 + t0 * d0, o1 + t0 * d1
    inside = (pu <= n_u) & (pv <= n_v) & (pu >= 0) & (pv >= 0)
    valid = inside & ~(t1 <= t0) & ~swapped
    return o0, o1, d0, d1, t0, t1, valid


LANE_MAPS = {
    # name -> function (warp, lane) -> (angle offset 0..7, t offset 0..31) inside the item
    "map2 (4a x 8t; quarter = 4a x 2t) [shipped]": lambda w, l: ((w & 1) * 4 + (l & 3), (w >> 1) * 8 + (l >> 2)),
    "map0 (2a x 16t quads)": lambda w, l: ((w & 3) * 2 + ((l & 3) & 1), (w >> 2) * 16 + (l >> 2) * 2 + ((l & 3) >> 1)),
    "map1 (1a x 32t)": lambda w, l: (w, l),
    "map3 (8a x 4t; quarter = 8a x 1t)": lambda w, l: (l & 7, w * 4 + (l >> 3)),
    "map5 (2a x 16t; quarter = 2a x 4t)": lambda w, l: ((w & 3) * 2 + (l & 1), (w >> 2) * 16 + (l >> 1)),
    "map6 (1a x 32t; quarter = 8 t strided by 4)": lambda w, l: (w, (l & 7) * 4 + (l >> 3)),
    "map7 (4a x 8t; quarter = 2a x 4t)": lambda w, l: ((w & 1) * 4 + (l >> 4) * 2 + (l & 1), (w >> 1) * 8 + ((l >> 1) & 7)),
}


def item_cells(n_u, n_v, n_alpha, n_t, ag, tg, chunk, lane_fn):
    """For item (angle group ag, t group tg): per lane the list of (chunk, rank-in-chunk, col, row) for lines A and B."""
    lanes = [(w, l) for w in range(8) for l in range(32)]
    ia = np.array([ag * 8 + lane_fn(w, l)[0] for w, l in lanes])
    it = np.array([tg * 32 + lane_fn(w, l)[1] for w, l in lanes])
    o0, o1, d0, d1, t0, t1, valid = bin_lines(ia.astype(np.float64), it.astype(np.float64), n_alpha, n_t, float(n_u), float(n_v))
    alpha_mid = ((ag * 8 + 3.5) / n_alpha - 0.5) * np.pi
    vertical = abs(np.sin(alpha_mid)) > abs(np.cos(alpha_mid))
    oA0 = o0 + 0.5 - 0.5 * d1
    oA1 = o1 + 0.5 + 0.5 * d0
    op, dp = (oA1, d1) if vertical else (oA0, d0)
    os_, ds = (oA0, d0) if vertical else (oA1, d1)
    offp, offs = (-d0, d1) if vertical else (d1, -d0)
    out = []
    for k in range(256):
        if not valid[k] or ia[k] >= n_alpha or it[k] >= n_t:
            out.append(None)
            continue
        n_s = int(np.floor((t1[k] - t0[k]) / 0.66)) + 1
        t = t0[k] + 0.66 * np.arange(n_s)
        pri, sec = t * dp[k] + op[k], t * ds[k] + os_[k]
        j = np.floor((pri - 0.5) / chunk).astype(np.int64)  # chunk of the sample (edge at j*chunk + 0.5)
        # rank inside the chunk in the order the lane takes its samples (t ascending)
        order = np.zeros(n_s, np.int64)
        change = np.r_[True, j[1:] != j[:-1]]
        start = np.maximum.accumulate(np.where(change, np.arange(n_s), 0))
        order = np.arange(n_s) - start
        cells = []
        for dpri, dsec in ((0.0, 0.0), (offp[k], offs[k])):
            col = np.floor(((pri + dpri) * 256.0 - 127.5)).astype(np.int64) >> 8
            row = np.floor(((sec + dsec) * 256.0 - 127.5)).astype(np.int64) >> 8
            cells.append((col, row))
        out.append((j, order, cells))
    return out


def wavefronts(addr):
    """addr: (K, 32, 8) int64 cell indices, -1 = inactive lane.  Returns (actual wavefronts, ideal wavefronts) summed."""
    act = addr >= 0
    bank = addr & 7
    same_addr = (addr[..., :, None] == addr[..., None, :]) & act[..., :, None] & act[..., None, :]
    earlier = np.tril(np.ones((8, 8), bool), -1)
    first = act & ~np.any(same_addr & earlier, axis=-1)           # first occurrence of its address
    same_bank = (bank[..., :, None] == bank[..., None, :]) & act[..., :, None]
    degree = np.sum(same_bank & first[..., None, :], axis=-1)     # distinct addresses in lane l's bank group
    degree = np.where(act, degree, 0)
    w = degree.max(axis=-1)
    return int(w.sum()), int((w > 0).sum())


def simulate(n_u, n_v, n_alpha, n_t, items, chunk, rows_mod, lane_fn, layout, seed=1):
    """layout: function (col, row) -> cell index (the address / 16).  Returns wavefronts per conflict-free wavefront."""
    rng = np.random.default_rng(seed)
    groups_a, groups_t = (n_alpha + 7) // 8, (n_t + 31) // 32
    tot, ideal = 0, 0
    for _ in range(items):
        ag, tg = int(rng.integers(groups_a)), int(rng.integers(groups_t))
        cells = item_cells(n_u, n_v, n_alpha, n_t, ag, tg, chunk, lane_fn)
        ks, ls, as_ = [], [], []
        for lane, c in enumerate(cells):
            if c is None:
                continue
            j, order, ab = c
            key = (j + 1024) * 4096 + order
            for line, (col, row) in enumerate(ab):
                for tap, (dc, dr) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                    ks.append(key * 8 + line * 4 + tap)
                    ls.append(np.full(len(key), lane))
                    as_.append(layout(col + dc, row + dr))
        if not ks:
            continue
        ks, ls, as_ = np.concatenate(ks), np.concatenate(ls), np.concatenate(as_)
        # chunks are global along the primary axis, so equal keys of different lanes are the same step of the lock-step loop
        uniq, inv = np.unique(ks, return_inverse=True)
        table = np.full((len(uniq), 256), -1, np.int64)
        table[inv, ls] = as_
        table = table.reshape(-1, 32, 8)
        w, i = wavefronts(table)
        tot += w
        ideal += i
    return tot / max(ideal, 1), ideal


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=40)
    ap.add_argument("--chunk", type=int, default=10)
    ap.add_argument("--size", default="1240,960,768,768")
    args = ap.parse_args()
    n_u, n_v, n_alpha, n_t = (int(x) for x in args.size.split(","))
    layouts = {}
    for rows in (137, 135, 139, 141, 136, 140):
        layouts[f"rows {rows} (= {rows % 8} mod 8)"] = (lambda r: (lambda col, row: (col & 1023) * r + (row & 4095)))(rows)
    layouts["rows 136 + xor swizzle (cell ^ (cell >> 3 & 7))"] = lambda col, row: ((col & 1023) * 136 + (row & 4095)) ^ ((((col & 1023) * 136 + (row & 4095)) >> 3) & 7)
    layouts["rows 137 + xor swizzle"] = lambda col, row: ((col & 1023) * 137 + (row & 4095)) ^ ((((col & 1023) * 137 + (row & 4095)) >> 3) & 7)
    print(f"geometry {n_u}x{n_v} -> {n_alpha}x{n_t}, chunk {args.chunk}, {args.items} random items per configuration")
    for (mname, fn), (lname, lay) in itertools.product(LANE_MAPS.items(), layouts.items()):
        r, ideal = simulate(n_u, n_v, n_alpha, n_t, args.items, args.chunk, 0, fn, lay)
        print(f"{r:6.3f} wavefronts per conflict-free one   {mname:48s} {lname}   ({ideal} quarter-warp accesses)", flush=True)


if __name__ == "__main__":
    main()
