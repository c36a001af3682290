"""Generates tests/golden/ref_fbcc_vectors.npz from the REFERENCE's own RectifiedFBCC.h (LinePerspectivity::transform /
inverse / derivative and the per-sample weight of kernel_computeLineIntegrals, EpipolarConsistencyDirect.cu:86-93, compiled
unchanged into oracle/_ref/libecc_ref_host.so by oracle/Makefile).  The records are those of a real pair (produced by OUR
restatement of the host geometry -- they are inputs here, stored with the outputs).  Run in the build container:
    make -C oracle ref && python tests/golden/make_ref_fbcc_vectors.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

R = ol.ref_host()
assert R is not None and hasattr(R, "ref_fbcc_weight"), "build oracle/_ref first (needs /root/reference)"
assert R.ref_fbcc_record_floats() == 8
rng = np.random.default_rng(20261019)
Ps = ol.circular_trajectory(12, 750.0, 1200.0, 320, 256, 200.0, 1.2)
recs = []
for (i, j) in [(0, 3), (2, 9), (5, 6), (1, 11)]:
    g = ol.direct_pair_geometry(Ps[i], Ps[j], 320, 256)
    pick = rng.integers(0, len(g["kappas"]), 24)
    recs += [g["fbcc0"][pick], g["fbcc1"][pick]]
recs = np.ascontiguousarray(np.concatenate(recs), np.float32)
ts = rng.uniform(-250.0, 250.0, size=(recs.shape[0], 16)).astype(np.float32)
out = np.zeros(ts.shape + (4,), np.float32)
for q, rec in enumerate(recs):
    for k, t in enumerate(ts[q]):
        out[q, k] = (R.ref_fbcc_transform(rec, t), R.ref_fbcc_inverse(rec, t), R.ref_fbcc_derivative(rec, t), R.ref_fbcc_weight(rec, t))
np.savez_compressed(os.path.join(HERE, "ref_fbcc_vectors.npz"), recs=recs, ts=ts, out=out)
print("wrote", recs.shape[0], "records x", ts.shape[1], "positions")
