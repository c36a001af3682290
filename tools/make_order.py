"""Development aid: from the per-item cycle counts of the quad Radon kernel (ECC_ITEM_CLOCK files of a texture-only and a
window-only run) build item orders for the static split (ECC_HYBRID4_ORDER files): items sorted by how much cheaper the
window path is than the texture path, cut where the window path's share of the samples reaches the given per mille.
Usage: python tools/make_order.py clock_mode1.bin clock_mode2.bin outdir 580 600 620 ..."""
import sys
import numpy as np
sys.path.insert(0, "tools")
import bank_sim as b


def load(f):
    raw = open(f, "rb").read()
    ga, gt, nq = np.frombuffer(raw[:16], np.int32)[:3]
    return np.frombuffer(raw[16:], np.uint64).astype(np.float64).reshape(2, gt, ga), ga, gt, nq


def samples(n_u, n_v, n_a, n_t, ga, gt):
    ix, iy = np.meshgrid(np.arange(n_a), np.arange(n_t))
    o0, o1, d0, d1, t0, t1, valid = b.bin_lines(ix.astype(float), iy.astype(float), n_a, n_t, float(n_u), float(n_v))
    cnt = np.where(valid, 2 * (np.floor((t1 - t0) / 0.66) + 1), 0)
    pad = np.zeros((gt * 32, ga * 8))
    pad[:n_t, :n_a] = cnt
    return pad.reshape(gt, 32, ga, 8).sum(axis=(1, 3))


if __name__ == "__main__":
    T, ga, gt, nq = load(sys.argv[1])
    W = load(sys.argv[2])[0][1]
    T = T[0]
    S = samples(1240, 960, 768, 768, ga, gt)
    ratio = (T / np.maximum(W, 1.0)).ravel()
    order = np.argsort(-ratio, kind="stable").astype(np.int32)
    cum = np.cumsum(S.ravel()[order]) / S.sum()
    for pm in sys.argv[4:]:
        m = int(np.searchsorted(cum, int(pm) / 1000.0))
        with open(f"{sys.argv[3]}/order_{pm}.bin", "wb") as f:
            f.write(np.int32(m).tobytes())
            f.write(order.tobytes())
        print(pm, "split", m, "window time share", W.ravel()[order[:m]].sum() / W.sum(), "texture", T.ravel()[order[m:]].sum() / T.sum())
