"""ctypes bindings for the CPU oracle (oracle/libecc_oracle.so) and, when present, the reference's
own code built into oracle/_ref/ (libecc_ref_host.so, libecc_ref_cuda.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under epipolarconsistency_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

INTERP_EXACT = 0
INTERP_TEX8 = 1

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build_oracle(ref=True):
    """Compile oracle/libecc_oracle.so (and oracle/_ref/* when /root/reference exists)."""
    targets = ["libecc_oracle.so"] + (["ref"] if ref else [])
    subprocess.run(["make", "-C", ORACLE_DIR] + targets, check=True, stdout=subprocess.DEVNULL)


_oracle = None


def oracle():
    global _oracle
    if _oracle is not None:
        return _oracle
    path = os.path.join(ORACLE_DIR, "libecc_oracle.so")
    if not os.path.exists(path):
        build_oracle(ref=False)
    L = C.CDLL(path)
    L.oracle_get_ij.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.oracle_pinv_transpose.argtypes = [_f64p, _f32p]
    L.oracle_source_position.argtypes = [_f64p, _f32p]
    L.oracle_object_radius.argtypes = [_f64p, C.c_int, C.c_int]
    L.oracle_object_radius.restype = C.c_double
    L.oracle_compute_k01.argtypes = [C.c_float, C.c_float, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                     C.c_float, C.c_float, C.c_int, _f32p, _f32p]
    L.oracle_line_to_sample.argtypes = [_f32p, C.c_float]
    L.oracle_line_to_sample.restype = C.c_int
    L.oracle_radon.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p]
    L.oracle_radon_num_samples.argtypes = [C.c_int] * 5
    L.oracle_radon_num_samples.restype = C.c_double
    L.oracle_ecc.argtypes = [_f64p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                             C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                             C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.oracle_ecc.restype = C.c_double
    L.oracle_circular_trajectory.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                             C.c_double, C.c_double, _f64p]
    L.oracle_project_ellipsoids.argtypes = [_f64p, C.c_int, C.c_int, _f64p, C.c_int, C.c_int, C.c_int, _f32p]
    L.oracle_direct_line_integrals.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                               C.c_int, _f32p]
    L.oracle_direct_pair_geometry.argtypes = [_f64p, _f64p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    L.oracle_direct_pair_geometry.restype = C.c_int
    L.oracle_direct_pair.argtypes = [_f64p, _f64p, _f32p, _f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.oracle_direct_pair.restype = C.c_double
    L.oracle_direct_evaluate.argtypes = [_f64p, C.c_int, _f32p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p]
    L.oracle_direct_evaluate.restype = C.c_double
    L.oracle_direct_fbcc_weight.argtypes = [_f32p, C.c_float]
    L.oracle_direct_fbcc_weight.restype = C.c_float
    L.oracle_max_threads.restype = C.c_int
    L.oracle_set_threads.argtypes = [C.c_int]
    L.oracle_set_threads.restype = C.c_int
    _oracle = L
    return L


def use_all_host_cores():
    """All host cores this process may run on for the oracle's OpenMP loops, whatever OMP_NUM_THREADS says (torchrun sets
    it to 1 for its ranks).  Returns the thread count in effect."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return oracle().oracle_set_threads(cores)


# ------------------------------------------------------------------------------------------------
# numpy-level helpers
# ------------------------------------------------------------------------------------------------
def get_ij(k, n):
    i, j = C.c_int(), C.c_int()
    oracle().oracle_get_ij(k, n, C.byref(i), C.byref(j))
    return i.value, j.value


def pinv_transpose(P):
    out = np.zeros(12, np.float32)
    oracle().oracle_pinv_transpose(np.ascontiguousarray(P, np.float64).reshape(12), out)
    return out


def source_position(P):
    out = np.zeros(4, np.float32)
    oracle().oracle_source_position(np.ascontiguousarray(P, np.float64).reshape(12), out)
    return out


def object_radius(P, n_u, n_v):
    return oracle().oracle_object_radius(np.ascontiguousarray(P, np.float64).reshape(12), n_u, n_v)


def compute_k01(half_nu, half_nv, C0, C1, P0invT, P1invT, radius, num_samples, dkappa, same_view=False):
    K0 = np.zeros(8, np.float32)
    K1 = np.zeros(8, np.float32)
    f = lambda a: np.ascontiguousarray(a, np.float32)
    oracle().oracle_compute_k01(half_nu, half_nv, f(C0), f(C1), f(P0invT), f(P1invT), radius,
                                num_samples, dkappa, int(same_view), K0, K1)
    return K0, K1


def line_to_sample(line, range_t):
    l = np.ascontiguousarray(line, np.float32).copy()
    flipped = oracle().oracle_line_to_sample(l, range_t)
    return l, flipped


def radon(img, n_alpha, n_t, filter=0, post=0, interp=INTERP_EXACT):
    img = np.ascontiguousarray(img, np.float32)
    n_v, n_u = img.shape
    out = np.zeros((n_t, n_alpha), np.float32)
    if filter == 1:  # ramp = plain line integrals (with post-processing) + ramp filter along t
        oracle().oracle_radon(img, n_u, n_v, n_alpha, n_t, 2, post, interp, out)
        return ramp_filter(out)
    oracle().oracle_radon(img, n_u, n_v, n_alpha, n_t, filter, post, interp, out)
    return out


def ramp_filter(dtr):
    """apply1DRampFilter (reference RadonIntermediate.cu:186-237) restated with numpy's FFT: R2C along t for every
    alpha column, bin k times k * (-0.5f / (n_t * (n_t/2+1))) in fp32 (:181-182,218), unnormalised C2R."""
    n_t = dtr.shape[0]
    n_theta = n_t // 2 + 1
    scale = np.float32(-0.5) / np.float32(n_t * n_theta)
    H = (np.arange(n_theta, dtype=np.float32) * scale).astype(np.float64)
    F = np.fft.rfft(dtr.astype(np.float64), axis=0) * H[:, None]
    return (np.fft.irfft(F, n=n_t, axis=0) * n_t).astype(np.float32)


def radon_num_samples(n_u, n_v, n_alpha, n_t, filter=0):
    return oracle().oracle_radon_num_samples(n_u, n_v, n_alpha, n_t, filter)


def ecc(Ps, dtrs, n_u, n_v, is_derivative=True, object_radius_mm=0.0, dkappa=0.0,
        interp=INTERP_EXACT, fast_sincos=False, idx4=None, want_out=True, want_ksamples=False, use_corr=False):
    """Returns (mean, out, ksamples).  Ps: (n,12) col-major doubles; dtrs: (m, n_t, n_alpha)."""
    Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
    dtrs = np.ascontiguousarray(dtrs, np.float32)
    n = Ps.shape[0]
    m, n_t, n_alpha = dtrs.shape
    diag = np.sqrt(float(n_u) ** 2 + float(n_v) ** 2)
    step_alpha = np.float32(np.pi / n_alpha)
    step_t = np.float32(diag / n_t)
    if idx4 is None:
        n_pairs = n * (n - 1) // 2
        out = np.zeros((n, n), np.float32) if want_out else None
        idx_ptr = None
    else:
        idx4 = np.ascontiguousarray(idx4, np.int32).reshape(-1, 4)
        n_pairs = idx4.shape[0]
        out = np.zeros(n_pairs, np.float32) if want_out else None
        idx_ptr = idx4.ctypes.data_as(C.c_void_p)
    ks = np.zeros(n_pairs, np.int32) if want_ksamples else None
    mean = oracle().oracle_ecc(Ps, n, dtrs, m, n_alpha, n_t, float(step_alpha), float(step_t), n_u, n_v,
                               int(is_derivative), float(object_radius_mm), float(dkappa), interp,
                               int(fast_sincos), idx_ptr, n_pairs,
                               out.ctypes.data_as(C.c_void_p) if out is not None else None,
                               ks.ctypes.data_as(C.c_void_p) if ks is not None else None, int(use_corr))
    return mean, out, ks


# ---- direct metric (EpipolarConsistencyDirect.{cpp,cu}, RectifiedFBCC.h) ----------------------------------------------------
def direct_line_integrals(img, lines, fbcc=None, interp=INTERP_TEX8, shape=1, n_v_clip=None):
    """kernel_computeLineIntegrals on the CPU.  lines [m, >= 3], fbcc [m, >= 6] or None.  shape 1: the loop as the reference's
    sm_100 build executes it.  n_v_clip: height the clipping uses (the reference's launcher passes n_u)."""
    img = np.ascontiguousarray(img, np.float32)
    lines = np.ascontiguousarray(lines, np.float32)
    n_v, n_u = img.shape
    out = np.zeros(lines.shape[0], np.float32)
    if fbcc is not None:
        fbcc = np.ascontiguousarray(fbcc, np.float32)
    oracle().oracle_direct_line_integrals(img, n_u, n_v, n_v if n_v_clip is None else n_v_clip, lines, lines.shape[0], lines.shape[1],
                                          None if fbcc is None else fbcc.ctypes.data, 0 if fbcc is None else fbcc.shape[1], interp,
                                          shape, out)
    return out


def direct_pair_geometry(P0, P1, n_u, n_v, radius=0.0, dkappa=0.0):
    """What computeForImagePair prepares for its kernel: dict(kappas, lines0, lines1, fbcc0, fbcc1, dkappa)."""
    P0 = np.ascontiguousarray(P0, np.float64).reshape(12)
    P1 = np.ascontiguousarray(P1, np.float64).reshape(12)
    dk = C.c_double()
    m = oracle().oracle_direct_pair_geometry(P0, P1, radius, dkappa, n_u, n_v, 0, None, None, None, None, None, C.byref(dk))
    out = dict(kappas=np.zeros(m, np.float32), lines0=np.zeros((m, 3), np.float32), lines1=np.zeros((m, 3), np.float32),
               fbcc0=np.zeros((m, 8), np.float32), fbcc1=np.zeros((m, 8), np.float32))
    oracle().oracle_direct_pair_geometry(P0, P1, radius, dkappa, n_u, n_v, m, out["kappas"].ctypes.data, out["lines0"].ctypes.data,
                                         out["lines1"].ctypes.data, out["fbcc0"].ctypes.data, out["fbcc1"].ctypes.data, C.byref(dk))
    out["dkappa"] = dk.value
    return out


def direct_pair(P0, P1, img0, img1, radius=0.0, dkappa=0.0, fbcc=False, interp=INTERP_TEX8, shape=1, reference_clip=False, kappas=None):
    """computeForImagePair on the CPU: dict(value, kappas, samples0, samples1)."""
    P0 = np.ascontiguousarray(P0, np.float64).reshape(12)
    P1 = np.ascontiguousarray(P1, np.float64).reshape(12)
    img0 = np.ascontiguousarray(img0, np.float32)
    img1 = np.ascontiguousarray(img1, np.float32)
    n_v, n_u = img0.shape
    n = C.c_int()
    args = (P0, P1, img0, img1, n_u, n_v, radius, dkappa, int(fbcc), interp, shape, int(reference_clip))
    if kappas is not None:
        k = np.ascontiguousarray(kappas, np.float32).copy()
        m = k.shape[0]
    else:
        m = oracle().oracle_direct_pair_geometry(P0, P1, radius, dkappa, n_u, n_v, 0, None, None, None, None, None, None)
        k = np.zeros(m, np.float32)
    s0, s1 = np.zeros(m, np.float32), np.zeros(m, np.float32)
    v = oracle().oracle_direct_pair(*args, m if kappas is not None else 0, m, k.ctypes.data, s0.ctypes.data, s1.ctypes.data, C.byref(n))
    return dict(value=v, kappas=k, samples0=s0, samples1=s1)


def direct_evaluate(Ps, images, radius=0.0, dkappa=0.0, fbcc=False, interp=INTERP_TEX8, shape=1, reference_clip=False, cost_image=None):
    """MetricDirect::evaluate on the CPU: the SUM over all pairs (radius 0: from the first matrix)."""
    Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
    images = np.ascontiguousarray(images, np.float32)
    n, n_v, n_u = images.shape
    return oracle().oracle_direct_evaluate(Ps, n, images, n_u, n_v, radius, dkappa, int(fbcc), interp, shape, int(reference_clip),
                                           None if cost_image is None else cost_image.ctypes.data)


def circular_trajectory(n, sid, sdd, n_u, n_v, max_angle_deg, pixel_spacing):
    Ps = np.zeros((n, 12), np.float64)
    oracle().oracle_circular_trajectory(n, sid, sdd, n_u, n_v, max_angle_deg, pixel_spacing, Ps)
    return Ps


def project_ellipsoids(P, n_u, n_v, ellipsoids, cos_weight=True, zero_border=True):
    ell = np.ascontiguousarray(ellipsoids, np.float64).reshape(-1, 7)
    img = np.zeros((n_v, n_u), np.float32)
    oracle().oracle_project_ellipsoids(np.ascontiguousarray(P, np.float64).reshape(12), n_u, n_v, ell,
                                       ell.shape[0], int(cos_weight), int(zero_border), img)
    return img


# ------------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own code
# ------------------------------------------------------------------------------------------------
_ref_host = None


def ref_host():
    """The reference's host-callable headers compiled unchanged; None if not built."""
    global _ref_host
    if _ref_host is not None:
        return _ref_host
    path = os.path.join(ORACLE_DIR, "_ref", "libecc_ref_host.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_get_ij.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ref_pinv_transpose.argtypes = [_f64p, _f32p]
    L.ref_source_position.argtypes = [_f64p, _f32p]
    L.ref_compute_k01.argtypes = [C.c_float, C.c_float, _f32p, _f32p, _f32p, _f32p, C.c_float,
                                  C.c_float, C.c_float, _f32p, _f32p]
    L.ref_line_to_sample.argtypes = [_f32p, C.c_float]
    L.ref_line_to_sample.restype = C.c_int
    if hasattr(L, "ref_fbcc_weight"):
        for name in ("ref_fbcc_transform", "ref_fbcc_inverse", "ref_fbcc_derivative", "ref_fbcc_weight"):
            getattr(L, name).argtypes = [_f32p, C.c_float]
            getattr(L, name).restype = C.c_float
        L.ref_fbcc_record_floats.restype = C.c_int
    _ref_host = L
    return L


_ref_cuda = None


def ref_cuda():
    """The reference's own CUDA kernels (sm_100 build) behind oracle/ref_cuda_harness.cu; None if not built.
    Needs a GPU to call."""
    global _ref_cuda
    if _ref_cuda is not None:
        return _ref_cuda
    path = os.path.join(ORACLE_DIR, "_ref", "libecc_ref_cuda.so")
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    L.ref_cuda_radon.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p,
                                 C.POINTER(C.c_float)]
    L.ref_cuda_metric_create.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ref_cuda_metric_create.restype = C.c_void_p
    L.ref_cuda_metric_destroy.argtypes = [C.c_void_p]
    L.ref_cuda_metric_set_matrices.argtypes = [C.c_void_p, _f64p, C.c_int]
    L.ref_cuda_metric_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                           C.POINTER(C.c_float)]
    L.ref_cuda_metric_evaluate.restype = C.c_double
    if hasattr(L, "ref_cuda_radon_any"):  # host or device pointers (torch tensors)
        L.ref_cuda_radon_any.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.POINTER(C.c_float)]
        L.ref_cuda_metric_create_any.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_cuda_metric_create_any.restype = C.c_void_p
    if hasattr(L, "ref_cuda_metric_set_launcher"):
        L.ref_cuda_metric_create_pitch2d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_cuda_metric_create_pitch2d.restype = C.c_void_p
        L.ref_cuda_metric_set_launcher.argtypes = [C.c_void_p, C.c_void_p]
    if hasattr(L, "ref_cuda_direct_line_integrals"):
        L.ref_cuda_direct_line_integrals.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                     C.c_void_p, C.POINTER(C.c_float)]
    if hasattr(L, "ref_cuda_metric_get_k01"):
        L.ref_cuda_metric_get_k01.argtypes = [C.c_void_p, _f32p, C.c_int]
    if hasattr(L, "ref_cuda_metric_evaluate_corr"):
        L.ref_cuda_metric_evaluate_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.ref_cuda_metric_evaluate_corr.restype = C.c_double
    _ref_cuda = L
    return L


def ref_cuda_direct_line_integrals(img, lines, fbcc=None):
    """The reference's kernel_computeLineIntegrals through its own launcher (which clips against n_u x n_u).  Returns
    (integrals, gpu_ms)."""
    L = ref_cuda()
    img = np.ascontiguousarray(img, np.float32)
    lines = np.ascontiguousarray(lines, np.float32)
    n_v, n_u = img.shape
    out = np.zeros(lines.shape[0], np.float32)
    if fbcc is not None:
        fbcc = np.ascontiguousarray(fbcc, np.float32)
    ms = C.c_float()
    rc = L.ref_cuda_direct_line_integrals(img.ctypes.data, n_u, n_v, lines.ctypes.data, lines.shape[0], lines.shape[1],
                                          None if fbcc is None else fbcc.ctypes.data, 0 if fbcc is None else fbcc.shape[1],
                                          out.ctypes.data, C.byref(ms))
    if rc:
        raise RuntimeError(f"ref_cuda_direct_line_integrals failed ({rc})")
    return out, ms.value


def ref_cuda_radon(images, n_alpha, n_t, filter=0, post=0):
    """Radon intermediates by the reference CUDA kernel.  Returns (dtrs, gpu_ms).  images: numpy array (host) or a
    contiguous float32 torch cuda tensor (then the dtrs come back as a cuda tensor)."""
    L = ref_cuda()
    if hasattr(images, "data_ptr"):
        import torch
        assert images.is_cuda and images.is_contiguous() and images.dtype == torch.float32
        n, n_v, n_u = images.shape
        out = torch.empty((n, n_t, n_alpha), dtype=torch.float32, device=images.device)
        ms = C.c_float()
        torch.cuda.synchronize()
        rc = L.ref_cuda_radon_any(images.data_ptr(), n, n_u, n_v, n_alpha, n_t, filter, post, out.data_ptr(), C.byref(ms))
        assert rc == 0
        return out, ms.value
    images = np.ascontiguousarray(images, np.float32)
    n, n_v, n_u = images.shape
    out = np.zeros((n, n_t, n_alpha), np.float32)
    ms = C.c_float()
    rc = L.ref_cuda_radon(images, n, n_u, n_v, n_alpha, n_t, filter, post, out, C.byref(ms))
    assert rc == 0
    return out, ms.value


class RefCudaMetric:
    """The reference MetricRadonIntermediate's device path (its launcher + kernels)."""

    def __init__(self, Ps, dtrs, n_u, n_v, is_derivative=True, pitch2d=False):
        """pitch2d (diagnostics): dtr textures over pitched linear memory instead of CUDA arrays."""
        self.L = ref_cuda()
        diag = np.sqrt(float(n_u) ** 2 + float(n_v) ** 2)
        if pitch2d:
            import torch
            keep = dtrs if hasattr(dtrs, "data_ptr") else torch.from_numpy(np.ascontiguousarray(dtrs, np.float32)).cuda()
            m, n_t, n_alpha = keep.shape
            torch.cuda.synchronize()
            self.h = self.L.ref_cuda_metric_create_pitch2d(keep.data_ptr(), m, n_alpha, n_t, np.float32(np.pi / n_alpha), np.float32(diag / n_t),
                                                           n_u, n_v, int(is_derivative))
        elif hasattr(dtrs, "data_ptr"):  # contiguous float32 torch cuda tensor
            import torch
            assert dtrs.is_cuda and dtrs.is_contiguous() and dtrs.dtype == torch.float32
            m, n_t, n_alpha = dtrs.shape
            torch.cuda.synchronize()
            self.h = self.L.ref_cuda_metric_create_any(dtrs.data_ptr(), m, n_alpha, n_t, np.float32(np.pi / n_alpha),
                                                       np.float32(diag / n_t), n_u, n_v, int(is_derivative))
        else:
            dtrs = np.ascontiguousarray(dtrs, np.float32)
            m, n_t, n_alpha = dtrs.shape
            self.h = self.L.ref_cuda_metric_create(dtrs, m, n_alpha, n_t, np.float32(np.pi / n_alpha),
                                                   np.float32(diag / n_t), n_u, n_v, int(is_derivative))
        assert self.h
        self.set_matrices(Ps)

    def set_matrices(self, Ps):
        Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
        self.n = Ps.shape[0]
        assert self.L.ref_cuda_metric_set_matrices(self.h, Ps, self.n) == 0

    def evaluate(self, radius, dkappa=0.0, idx4=None):
        """Returns (mean, out, gpu_ms); out is the n*n cost image or the per-pair list."""
        ms = C.c_float()
        if idx4 is None:
            out = np.zeros((self.n, self.n), np.float32)
            mean = self.L.ref_cuda_metric_evaluate(self.h, None, 0, radius, dkappa, out.ctypes.data_as(C.c_void_p),
                                                   C.byref(ms))
        else:
            idx4 = np.ascontiguousarray(idx4, np.int32).reshape(-1, 4)
            out = np.zeros(idx4.shape[0], np.float32)
            mean = self.L.ref_cuda_metric_evaluate(self.h, idx4.ctypes.data_as(C.c_void_p), idx4.shape[0], radius,
                                                   dkappa, out.ctypes.data_as(C.c_void_p), C.byref(ms))
        return mean, out, ms.value

    def set_launcher(self, address):
        """Diagnostics: evaluate() through another function with the reference launcher's signature (None = the reference's)."""
        self.L.ref_cuda_metric_set_launcher(self.h, address)

    def k01(self, n_pairs):
        """The K01 records (n_pairs, 16) the reference's own kernel computed in the last evaluate call."""
        K = np.zeros((n_pairs, 16), np.float32)
        assert self.L.ref_cuda_metric_get_k01(self.h, K, n_pairs) == n_pairs
        return K

    def evaluate_corr(self, radius, dkappa, idx4=None):
        """The reference's correlation variant (useCorrelation(true)).  Returns (mean, out)."""
        if idx4 is None:
            out = np.zeros((self.n, self.n), np.float32)
            mean = self.L.ref_cuda_metric_evaluate_corr(self.h, None, 0, radius, dkappa, out.ctypes.data_as(C.c_void_p))
        else:
            idx4 = np.ascontiguousarray(idx4, np.int32).reshape(-1, 4)
            out = np.zeros(idx4.shape[0], np.float32)
            mean = self.L.ref_cuda_metric_evaluate_corr(self.h, idx4.ctypes.data_as(C.c_void_p), idx4.shape[0], radius,
                                                        dkappa, out.ctypes.data_as(C.c_void_p))
        return mean, out

    def close(self):
        if self.h:
            self.L.ref_cuda_metric_destroy(self.h)
            self.h = None


def ref_nrrd():
    """The reference's own NRRD reader/writer (via oracle/_ref/libecc_ref_host.so); None if not built."""
    L = ref_host()
    if L is None or not hasattr(L, "ref_nrrd_load"):
        return None
    L.ref_nrrd_load.argtypes = [C.c_char_p, _f32p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p,
                                C.c_char_p, C.c_int]
    L.ref_nrrd_save.argtypes = [C.c_char_p, _f32p, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    return L


# ---------------------------------------------------------------------------------------------------------------
# Pre-processing: numpy restatement of Gui/PreProccess.cpp:57-166
# ---------------------------------------------------------------------------------------------------------------
def _feather(x):
    """weighting(), EpipolarConsistencyCommon.hxx:30-35"""
    x = np.float32(x)
    if x < -1 or x > 1:
        return np.float32(0)
    xx = np.float32(x * x)
    return np.float32(np.float32(1) - np.float32(2) * xx + xx * xx)


def camera_intrinsics(P):
    """K(0,0), K(0,2), K(1,2) of Geometry::getCameraIntrinsics (ProjectionMatrix.cpp:27-67): RQ decomposition of the left
    3x3 block with positive diagonal and K(2,2) = 1, by numpy's QR of the row-reversed transpose."""
    M = np.asarray(P, np.float64).reshape(4, 3).T[:, :3]
    Q, R = np.linalg.qr(M[::-1].T)
    K = R.T[::-1, ::-1]
    S = np.diag(np.sign(np.diag(K)))
    K = K @ S
    K = K / K[2, 2]
    return K[0, 0], K[0, 2], K[1, 2]


def preprocess(img, scale=1.0, bias=0.0, normalize=False, apply_log=False, zero=(1, 1, 1, 1), feather=(0, 0, 0, 0),
               blanks=(), flip_u=False, flip_v=False, sigma=1.84, k=5, P=None):
    img = np.array(img, np.float32)
    h, w = img.shape
    scale, bias = np.float32(scale), np.float32(bias)
    if normalize:  # :66-75
        scale, bias = np.float32(scale / img.max()), np.float32(0)
    img = (img * scale).astype(np.float32) + bias  # :81
    if apply_log:
        with np.errstate(all="ignore"):
            img = (-np.log(img)).astype(np.float32)
    img[(img < 0) | ~np.isfinite(img)] = 0  # :85-86
    for b in range(zero[0] + feather[0]):  # left :91-93
        img[:, b] *= np.float32(0) if b <= zero[0] else _feather(1 - np.float32(b - zero[0]) / feather[0])
    for b in range(1, zero[1] + feather[1] + 1):  # right :96-98
        img[:, w - b] *= np.float32(0) if b <= zero[1] else _feather(1 - np.float32(b - zero[1]) / feather[1])
    for b in range(1, zero[2] + feather[2] + 1):  # bottom :101-103
        img[h - b, :] *= np.float32(0) if b <= zero[2] else _feather(1 - np.float32(b - zero[2]) / feather[2])
    for b in range(zero[3] + feather[3]):  # top :106-108
        img[b, :] *= np.float32(0) if b <= zero[3] else _feather(1 - np.float32(b - zero[3]) / feather[3])
    for (x0, y0, x1, y1) in blanks:  # :111-114 (bounds clamped to the image)
        img[max(0, y0):max(0, y1), max(0, x0):max(0, x1)] = 0
    if flip_u:
        img = img[:, ::-1]
    if flip_v:
        img = img[::-1, :]
    img = np.ascontiguousarray(img)
    if sigma > 0 and k > 1:  # nrrd_lowpass.hxx:18-33 kernel, :46-79 taps -k .. k-1 (sic), clamped, fp64 sums, float stores
        kern = np.exp(-0.5 * (np.arange(-k, k + 1) / sigma) ** 2)
        kern /= kern.sum()
        work = np.zeros((h, w), np.float64)
        for o in range(-k, k):
            work += img[:, np.clip(np.arange(w) + o, 0, w - 1)].astype(np.float64) * kern[o + k]
        work = work.astype(np.float32)
        out = np.zeros((h, w), np.float64)
        for o in range(-k, k):
            out += work[np.clip(np.arange(h) + o, 0, h - 1), :].astype(np.float64) * kern[o + k]
        img = out.astype(np.float32)
    if P is not None and np.any(np.asarray(P) != 0):  # apply_weight_cos_principal_ray :146-166
        fu, u0, v0 = (np.float32(x) for x in camera_intrinsics(P))
        pou = np.arange(w, dtype=np.float32)[None, :] - u0
        pov = np.arange(h, dtype=np.float32)[:, None] - v0
        img = img * (fu / np.sqrt(pou * pou + pov * pov + fu * fu, dtype=np.float32))
    return img.astype(np.float32)
