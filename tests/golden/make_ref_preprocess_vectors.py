"""Generates tests/golden/ref_preprocess_vectors.npz from the REFERENCE's own headers: NRRD::gaussianKernel / lowpass2D
(HeaderOnly/NRRD/nrrd_lowpass.hxx) and weighting() (EpipolarConsistencyCommon.hxx:30-35) compiled unchanged into
oracle/_ref/libecc_ref_host.so, plus the four border loops of PreProccess::process around that weighting()
(oracle/ref_host_wrap.cpp).  Pins the pre-processing row.  Run in the build container:
    make -C oracle ref && python tests/golden/make_ref_preprocess_vectors.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

R = ol.ref_host()
assert R is not None and hasattr(R, "ref_lowpass2d"), "build oracle/_ref first (needs /root/reference)"
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
R.ref_gaussian_kernel.argtypes = [C.c_double, C.c_int, f64p]
R.ref_lowpass2d.argtypes = [f32p, C.c_int, C.c_int, C.c_double, C.c_int]
R.ref_weighting.argtypes = [C.c_float]
R.ref_weighting.restype = C.c_float
R.ref_border.argtypes = [f32p, C.c_int, C.c_int, i32p, i32p]
rng = np.random.default_rng(20261019)
out = {}

# Gaussian kernels
kernel_cases = np.array([[1.84, 5], [2.5, 3], [0.7, 2]])
for q, (sigma, k) in enumerate(kernel_cases):
    g = np.zeros(2 * int(k) + 1)
    assert R.ref_gaussian_kernel(float(sigma), int(k), g) == 2 * int(k) + 1
    out[f"kernel_{q}"] = g
out["kernel_cases"] = kernel_cases

# weighting()
xs = np.linspace(-1.25, 1.25, 101).astype(np.float32)
out["weighting_x"] = xs
out["weighting_y"] = np.array([R.ref_weighting(float(x)) for x in xs], np.float32)

# low-pass (in place) of rough images, incl. the reference's tap range -k .. k-1 and its edge clamping
lowpass_cases = np.array([[1.84, 5], [2.5, 3]])
img = (rng.random((33, 41), dtype=np.float32) * 100).astype(np.float32)
out["lowpass_in"] = img
out["lowpass_cases"] = lowpass_cases
for q, (sigma, k) in enumerate(lowpass_cases):
    work = img.copy()
    R.ref_lowpass2d(work, img.shape[1], img.shape[0], float(sigma), int(k))
    out[f"lowpass_out_{q}"] = work

# borders: zero + feather, order left, right, bottom, top
border_cases = np.array([[[1, 1, 1, 1], [16, 16, 16, 16]], [[2, 0, 3, 1], [6, 9, 0, 5]], [[0, 0, 0, 0], [4, 4, 4, 4]]], np.int32)
bimg = (rng.random((50, 60), dtype=np.float32) * 10 + 1).astype(np.float32)
out["border_in"] = bimg
out["border_cases"] = border_cases
for q, (zero, feather) in enumerate(border_cases):
    work = bimg.copy()
    R.ref_border(work, bimg.shape[1], bimg.shape[0], np.ascontiguousarray(zero), np.ascontiguousarray(feather))
    out[f"border_out_{q}"] = work

np.savez_compressed(os.path.join(HERE, "ref_preprocess_vectors.npz"), **out)
print("written", os.path.join(HERE, "ref_preprocess_vectors.npz"))
