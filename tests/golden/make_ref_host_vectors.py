"""Generates tests/golden/ref_host_vectors.npz from the REFERENCE's own headers
(oracle/_ref/libecc_ref_host.so = EpipolarConsistencyCommon.hxx + culaut/xprojectionmatrix.hxx compiled
unchanged from /root/reference by oracle/Makefile).  Run in the build container:
    make -C oracle ref && python tests/golden/make_ref_host_vectors.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

R = ol.ref_host()
assert R is not None, "build oracle/_ref first (needs /root/reference)"
rng = np.random.default_rng(20261018)

# --- get_ij over several n
ij = {}
for n in (2, 3, 7, 100, 496):
    tab = np.zeros((n * (n - 1) // 2, 2), np.int32)
    for k in range(tab.shape[0]):
        i, j = C.c_int(), C.c_int()
        R.ref_get_ij(k, n, C.byref(i), C.byref(j))
        tab[k] = (i.value, j.value)
    ij[f"get_ij_{n}"] = tab

# --- projection matrices: a circular trajectory (our generator; inputs are stored) plus random perturbations
Ps = ol.circular_trajectory(24, 750.0, 1200.0, 320, 256, 200.0, 1.2)
Ps = np.concatenate([Ps, Ps * (1 + 1e-3 * rng.standard_normal(Ps.shape))])
pinvT = np.zeros((Ps.shape[0], 12), np.float32)
Cs = np.zeros((Ps.shape[0], 4), np.float32)
for k, P in enumerate(Ps):
    R.ref_pinv_transpose(P, pinvT[k])
    R.ref_source_position(P, Cs[k])

# --- computeK01 for random view pairs, three (radius, dkappa) settings
pairs = rng.integers(0, Ps.shape[0], size=(64, 2)).astype(np.int32)
pairs = pairs[pairs[:, 0] != pairs[:, 1]]
settings = np.array([[100.0, 0.0], [40.0, 0.0], [1000.0, 1.7453292e-4]], np.float32)
K01 = np.zeros((len(settings), pairs.shape[0], 16), np.float32)
for s, (radius, dk) in enumerate(settings):
    for q, (a, b) in enumerate(pairs):
        K0 = np.zeros(8, np.float32)
        K1 = np.zeros(8, np.float32)
        R.ref_compute_k01(160.0, 128.0, Cs[a].copy(), Cs[b].copy(), pinvT[a].copy(), pinvT[b].copy(), radius, 819.6,
                          dk, K0, K1)
        K01[s, q, :8], K01[s, q, 8:] = K0, K1

# --- lineToSampleDtr on random lines
lines = rng.standard_normal((512, 3)).astype(np.float32) * np.array([1, 1, 200], np.float32)
samples = np.zeros((512, 3), np.float32)
for k in range(512):
    l = lines[k, :3].copy()
    flipped = R.ref_line_to_sample(l, 409.8)
    samples[k] = (l[0], l[1], flipped)

np.savez_compressed(os.path.join(HERE, "ref_host_vectors.npz"), Ps=Ps, pinvT=pinvT, Cs=Cs, pairs=pairs,
                    settings=settings, K01=K01, lines=lines, samples=samples, **ij)
print("written", os.path.join(HERE, "ref_host_vectors.npz"))
