"""Host-side mirror of the reference's operator interface for the hot path, on top of the C ABI.

The reference's public interface for this path is three C++ classes (code/LibEpipolarConsistency/):
`EpipolarConsistency::Metric` (EpipolarConsistency.h:49-94), `MetricRadonIntermediate`
(EpipolarConsistencyRadonIntermediate.h:21-106) and `RadonIntermediate` (RadonIntermediate.h:18-128).
The C++ facade with those exact names lives in include/EpipolarConsistency/; this module is the same
interface for Python drivers (tests, bench, torch.distributed runs): same method names, argument
meaning and return values.  All compute goes through libecc_b200.so -- nothing here computes on the
CPU, and the oracle is never imported.

Device memory is held in torch tensors (plumbing only); host data are numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib

FILTER_DERIVATIVE, FILTER_RAMP, FILTER_NONE = 0, 1, 2
POST_IDENTITY, POST_SQRT, POST_LOG = 0, 1, 2
INTERP_TEXTURE, INTERP_EXACT, INTERP_HYBRID, INTERP_HYBRID_STATIC = 0, 1, 2, 3


class EccError(RuntimeError):
    pass


_F32, _F64, _I32, _I64, _U8 = "float32", "float64", "int32", "int64", "uint8"


def _ptr(x, dtype=None):
    """Raw address of a numpy array / torch tensor (host or device), or None.  dtype: the element type the C ABI reads or
    writes through this pointer ("float32" images / dtrs / outputs, "float64" matrices, "int32" pair lists ...); a buffer
    of another type would be reinterpreted silently, so it is refused."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise TypeError(f"buffer of dtype {x.dtype} passed where the C ABI expects {dtype}")
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        if not x.is_contiguous():
            raise ValueError("tensor must be contiguous")
        if dtype is not None and str(x.dtype).replace("torch.", "") != dtype:
            raise TypeError(f"tensor of dtype {x.dtype} passed where the C ABI expects {dtype}")
        return x.data_ptr()
    raise TypeError(f"unsupported buffer type {type(x)}")


class PreprocessParams(C.Structure):
    """ecc_preprocess_params (include/ecc_b200.h); field names follow the reference's GetSet keys
    (Gui/PreProccess.cpp:42-55)."""
    _fields_ = [("scale", C.c_double), ("bias", C.c_double), ("normalize", C.c_int), ("apply_log", C.c_int),
                ("border_zero", C.c_int * 4), ("border_feather", C.c_int * 4), ("n_blanks", C.c_int),
                ("blanks", C.POINTER(C.c_int)), ("flip_u", C.c_int), ("flip_v", C.c_int), ("gaussian_sigma", C.c_double),
                ("half_kernel_width", C.c_int), ("cos_weight", C.c_int)]

    @classmethod
    def defaults(cls):
        p = cls()
        _lib.load().ecc_preprocess_defaults(C.byref(p))
        return p


class Context:
    """One GPU + one stream (ecc_context).  Thin, explicit wrapper of the C ABI."""

    def __init__(self, device=-1, stream="torch"):
        """stream: "torch" (default) issues all work on torch's current CUDA stream of the device, so that
        tensors produced or consumed by torch ops are ordered with the library's kernels; None keeps the
        context's own non-blocking stream; or pass a torch.cuda.Stream / raw cudaStream_t."""
        self.lib = _lib.load()
        h = _lib.c_ctx()
        rc = self.lib.ecc_create(int(device), C.byref(h))
        if rc != 0:
            raise EccError(f"ecc_create(device={device}) failed with {rc}: no usable CUDA device? "
                           "(this library has no CPU fallback)")
        self.h = h
        if isinstance(stream, str) and stream == "torch":
            import torch
            self.set_stream(torch.cuda.current_stream(device if device >= 0 else None))
        elif stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.ecc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise EccError(f"libecc_b200 error {rc}: {self.lib.ecc_last_error(self.h).decode()}")

    # -- plumbing
    def set_stream(self, stream):
        """stream: torch.cuda.Stream, raw cudaStream_t int, or None for the context's own."""
        raw = None if stream is None else int(getattr(stream, "cuda_stream", stream))
        if raw == 0:
            raw = 1  # cudaStreamLegacy: torch's default stream, spelled so that it is not "NULL = own stream"
        self._check(self.lib.ecc_set_stream(self.h, raw))

    def synchronize(self):
        self._check(self.lib.ecc_synchronize(self.h))

    # -- pre-processing (in place)
    def preprocess(self, images, params, Ps=None, blanks=None):
        """images: (n, n_v, n_u) float32 numpy array or torch cuda tensor; params: PreprocessParams; Ps: (n,12) doubles for
        the cosine weighting; blanks: (k,4) ints (x0,y0,x1,y1)."""
        n, n_v, n_u = images.shape
        keep = None
        if blanks is not None and len(blanks):
            keep = np.ascontiguousarray(blanks, np.int32).reshape(-1, 4)
            params.n_blanks = keep.shape[0]
            params.blanks = keep.ctypes.data_as(C.POINTER(C.c_int))
        else:
            params.n_blanks = 0
        if Ps is not None:
            Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
            assert Ps.shape[0] == n
        self._check(self.lib.ecc_preprocess(self.h, _ptr(images, _F32), n, n_u, n_v, C.addressof(params), _ptr(Ps, _F64)))
        return images

    # -- Radon
    def radon_compute(self, images, n_alpha, n_t, filter=FILTER_DERIVATIVE, post=POST_IDENTITY,
                      interp=INTERP_TEXTURE, out=None):
        """images: (n, n_v, n_u) float32, numpy (host) or torch cuda tensor.  Returns the dtrs
        (n, n_t, n_alpha) in the same kind of memory (or writes into `out`)."""
        n, n_v, n_u = images.shape
        if out is None:
            if isinstance(images, np.ndarray):
                out = np.empty((n, n_t, n_alpha), np.float32)
            else:
                import torch
                out = torch.empty((n, n_t, n_alpha), dtype=torch.float32, device=images.device)
        self._check(self.lib.ecc_radon_compute(self.h, _ptr(images, _F32), n, n_u, n_v, n_alpha, n_t, filter, post,
                                               interp, _ptr(out, _F32)))
        return out

    def radon_calibrate_split(self, n_u, n_v, n_alpha, n_t):
        """Window path's share of the samples (per mille) at which the two sampling pipes of THIS GPU balance for this geometry."""
        v = C.c_int()
        self._check(self.lib.ecc_radon_calibrate_split(self.h, n_u, n_v, n_alpha, n_t, C.byref(v)))
        return v.value

    def radon_set_split(self, permille):
        """Pins the static split of INTERP_HYBRID_STATIC for this context (0 = built-in 580 / 605 per mille)."""
        self._check(self.lib.ecc_radon_set_split(self.h, int(permille)))

    def radon_num_samples(self, n_u, n_v, n_alpha, n_t, filter=FILTER_DERIVATIVE):
        c = C.c_double()
        self._check(self.lib.ecc_radon_num_samples(self.h, n_u, n_v, n_alpha, n_t, filter, C.byref(c)))
        return c.value

    @staticmethod
    def radon_bin_sizes(n_u, n_v, n_alpha, n_t):
        a, t = C.c_double(), C.c_double()
        _lib.load().ecc_radon_bin_sizes(n_u, n_v, n_alpha, n_t, C.byref(a), C.byref(t))
        return a.value, t.value

    # -- metric state
    def set_radon_intermediates(self, dtrs, n_u, n_v, is_derivative=True, step_alpha=None, step_t=None):
        n, n_t, n_alpha = dtrs.shape
        sa, st = self.radon_bin_sizes(n_u, n_v, n_alpha, n_t)
        self._dtrs_keepalive = dtrs  # borrowed by the library when on the device
        self._check(self.lib.ecc_set_radon_intermediates(
            self.h, _ptr(dtrs, _F32), n, n_alpha, n_t, sa if step_alpha is None else step_alpha,
            st if step_t is None else step_t, n_u, n_v, int(is_derivative)))

    def set_projection_matrices(self, Ps):
        Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
        self._check(self.lib.ecc_set_projection_matrices(self.h, _ptr(Ps, _F64), Ps.shape[0]))

    def update_projection_matrix(self, index, P):
        P = np.ascontiguousarray(P, np.float64).reshape(12)
        self._check(self.lib.ecc_update_projection_matrix(self.h, int(index), _ptr(P, _F64)))

    def get_derived_views(self, n_views):
        A = np.zeros((n_views, 12), np.float32)
        Cs = np.zeros((n_views, 4), np.float32)
        self._check(self.lib.ecc_get_derived_views(self.h, _ptr(A, _F32), _ptr(Cs, _F32)))
        return A, Cs

    def set_object_radius(self, r):
        self._check(self.lib.ecc_set_object_radius(self.h, float(r)))

    def get_object_radius(self):
        r = C.c_double()
        self._check(self.lib.ecc_get_object_radius(self.h, C.byref(r)))
        return r.value

    def set_epipolar_plane_step(self, dkappa):
        self._check(self.lib.ecc_set_epipolar_plane_step(self.h, float(dkappa)))

    def use_correlation(self, on=True):
        self._check(self.lib.ecc_use_correlation(self.h, int(bool(on))))

    def set_interpolation(self, interp):
        self._check(self.lib.ecc_set_interpolation(self.h, int(interp)))

    # -- evaluation
    def evaluate(self, cost_image=None, want_mean=True):
        m = C.c_double()
        self._check(self.lib.ecc_evaluate(self.h, _ptr(cost_image, _F32), C.byref(m) if want_mean else None))
        return m.value if want_mean else None

    def evaluate_range(self, begin, end, cost_image=None, want_sum=True):
        s = C.c_double()
        self._check(self.lib.ecc_evaluate_range(self.h, int(begin), int(end), _ptr(cost_image, _F32),
                                                C.byref(s) if want_sum else None))
        return s.value if want_sum else None

    def evaluate_indices(self, idx4, out=None, want_mean=True):
        n_pairs = idx4.shape[0]
        m = C.c_double()
        self._check(self.lib.ecc_evaluate_indices(self.h, _ptr(idx4, _I32), n_pairs, _ptr(out, _F32),
                                                  C.byref(m) if want_mean else None))
        return m.value if want_mean else None

    def update_and_evaluate(self, index, P, idx4, out=None):
        """One tracking step: replace matrix `index`, evaluate the listed pairs; replayed as one CUDA graph from the third
        call with the same index / list / settings on.  Returns the mean."""
        P = np.ascontiguousarray(P, np.float64).reshape(12)
        m = C.c_double()
        self._check(self.lib.ecc_update_and_evaluate(self.h, int(index), _ptr(P, _F64), _ptr(idx4, _I32), idx4.shape[0], _ptr(out, _F32), C.byref(m)))
        return m.value

    def track_info(self):
        """(kernel launches per replayed tracking call: 0 nothing recorded / 1 fused / 2 plain recording, replays so far)"""
        k, r = C.c_int(), C.c_longlong()
        self._check(self.lib.ecc_track_info(self.h, C.byref(k), C.byref(r)))
        return k.value, r.value

    def evaluate_batch(self, Ps_sets, idx4=None, out=None, want_means=True):
        if isinstance(Ps_sets, np.ndarray):
            Ps_sets = np.ascontiguousarray(Ps_sets, np.float64)
        n_sets = Ps_sets.shape[0]
        n_pairs = 0 if idx4 is None else idx4.shape[0]
        means = np.zeros(n_sets, np.float64) if want_means else None
        self._check(self.lib.ecc_evaluate_batch(self.h, _ptr(Ps_sets, _F64), n_sets, _ptr(idx4, _I32), n_pairs, _ptr(out, _F32),
                                                _ptr(means, _F64)))
        return means

    def evaluate_batch_params(self, base_Ps, params, view_to_param=None, idx4=None, out=None, want_means=True):
        """Batched mode fed with parameter vectors: params (K, m, 11) model instances (ModelCameraSimilarity2D3D) applied to
        the base matrices (n, 12) -- None = the current set -- on the device.  view_to_param (n,) int32: instance of a set
        that moves view v, negative = none; None = one instance per view (m == n, ModelFDCT).  Returns the K means."""
        if isinstance(params, np.ndarray):
            params = np.ascontiguousarray(params, np.float64)
        K, m = params.shape[0], params.shape[1]
        if base_Ps is not None and isinstance(base_Ps, np.ndarray):
            base_Ps = np.ascontiguousarray(base_Ps, np.float64).reshape(-1, 12)
        if view_to_param is not None and isinstance(view_to_param, np.ndarray):
            view_to_param = np.ascontiguousarray(view_to_param, np.int32)
        n_pairs = 0 if idx4 is None else idx4.shape[0]
        means = np.zeros(K, np.float64) if want_means else None
        self._check(self.lib.ecc_evaluate_batch_params(self.h, _ptr(base_Ps, _F64), _ptr(params, _F64), K, m, _ptr(view_to_param, _I32),
                                                       _ptr(idx4, _I32), n_pairs, _ptr(out, _F32), _ptr(means, _F64)))
        return means

    def model_expand(self, base_Ps, params, view_to_param=None, n_views=None):
        """The matrices (K, n, 12) the device expands from params (K, m, 11): see evaluate_batch_params."""
        params = np.ascontiguousarray(params, np.float64)
        K, m = params.shape[0], params.shape[1]
        if base_Ps is not None:
            base_Ps = np.ascontiguousarray(base_Ps, np.float64).reshape(-1, 12)
            n_views = base_Ps.shape[0]
        if view_to_param is not None:
            view_to_param = np.ascontiguousarray(view_to_param, np.int32)
        out = np.zeros((K, n_views, 12), np.float64)
        self._check(self.lib.ecc_model_expand(self.h, _ptr(base_Ps, _F64), _ptr(params, _F64), K, m, _ptr(view_to_param, _I32), _ptr(out, _F64)))
        return out

    def evaluate_batch_transforms(self, base_Ps, transforms, view_to_transform=None, normalize=False, idx4=None, out=None):
        """Batched mode fed with explicit homographies: transforms (K, m, 25) = H (3x3) then T (4x4), column-major; P' = H P T
        (normalised when asked) on the device.  view_to_transform None: m == n, or m == 1 = one correction for the whole
        trajectory (ModelFDCTCalibrationCorrection).  Returns the K means."""
        transforms = np.ascontiguousarray(transforms, np.float64)
        K, m = transforms.shape[0], transforms.shape[1]
        if base_Ps is not None:
            base_Ps = np.ascontiguousarray(base_Ps, np.float64).reshape(-1, 12)
        if view_to_transform is not None:
            view_to_transform = np.ascontiguousarray(view_to_transform, np.int32)
        n_pairs = 0 if idx4 is None else idx4.shape[0]
        means = np.zeros(K, np.float64)
        self._check(self.lib.ecc_evaluate_batch_transforms(self.h, _ptr(base_Ps, _F64), _ptr(transforms, _F64), K, m,
                                                           _ptr(view_to_transform, _I32), int(bool(normalize)), _ptr(idx4, _I32), n_pairs,
                                                           _ptr(out, _F32), _ptr(means, _F64)))
        return means

    def transform_expand(self, base_Ps, transforms, view_to_transform=None, normalize=False, n_views=None):
        """The matrices (K, n, 12) the device builds from transforms (K, m, 25): see evaluate_batch_transforms."""
        transforms = np.ascontiguousarray(transforms, np.float64)
        K, m = transforms.shape[0], transforms.shape[1]
        if base_Ps is not None:
            base_Ps = np.ascontiguousarray(base_Ps, np.float64).reshape(-1, 12)
            n_views = base_Ps.shape[0]
        if view_to_transform is not None:
            view_to_transform = np.ascontiguousarray(view_to_transform, np.int32)
        out = np.zeros((K, n_views, 12), np.float64)
        self._check(self.lib.ecc_transform_expand(self.h, _ptr(base_Ps, _F64), _ptr(transforms, _F64), K, m, _ptr(view_to_transform, _I32),
                                                  int(bool(normalize)), _ptr(out, _F64)))
        return out

    def pair_signals(self, i, j, dtr_i=None, dtr_j=None):
        """evaluateForImagePair (EpipolarConsistencyRadonIntermediate.cpp:324-393): the redundant signals of one pair in
        ascending kappa.  Returns dict(kappas, signal0, signal1, lines0, lines1, weight, value)."""
        dtr_i = i if dtr_i is None else dtr_i
        dtr_j = j if dtr_j is None else dtr_j
        n, w, v = C.c_int(), C.c_double(), C.c_double()
        self._check(self.lib.ecc_pair_signals(self.h, i, j, dtr_i, dtr_j, 0, None, None, None, None, None, C.byref(n), None, None))
        m = n.value
        out = dict(kappas=np.zeros(m, np.float32), signal0=np.zeros(m, np.float32), signal1=np.zeros(m, np.float32),
                   lines0=np.zeros((m, 2), np.float32), lines1=np.zeros((m, 2), np.float32))
        self._check(self.lib.ecc_pair_signals(self.h, i, j, dtr_i, dtr_j, m, _ptr(out["kappas"], _F32), _ptr(out["signal0"], _F32),
                                              _ptr(out["signal1"], _F32), _ptr(out["lines0"], _F32), _ptr(out["lines1"], _F32), C.byref(n),
                                              C.byref(w), C.byref(v)))
        out["weight"], out["value"] = w.value, v.value
        return out

    # -- direct metric (include/ecc_b200.h "Direct metric"): no Radon intermediates
    def direct_set_images(self, images, n_u=None, n_v=None):
        """MetricDirect::setProjectionImages: images [n, n_v, n_u] float32, host or device."""
        n = images.shape[0]
        n_v = images.shape[1] if n_v is None else n_v
        n_u = images.shape[2] if n_u is None else n_u
        self._check(self.lib.ecc_direct_set_images(self.h, _ptr(images, _F32), int(n), int(n_u), int(n_v)))

    def direct_set_fan_beam(self, fbcc=True):
        self._check(self.lib.ecc_direct_set_fan_beam(self.h, int(bool(fbcc))))

    def direct_set_reference_clip(self, on=True):
        self._check(self.lib.ecc_direct_set_reference_clip(self.h, int(bool(on))))

    def direct_evaluate(self, cost_image=None):
        """MetricDirect::evaluate: the SUM over all pairs; cost_image [n, n] float32 gets entry [j, i] = pair (i < j)."""
        total = C.c_double()
        self._check(self.lib.ecc_direct_evaluate(self.h, _ptr(cost_image, _F32), C.byref(total)))
        return total.value

    def direct_evaluate_range(self, begin, end, cost_image=None):
        """Pairs [begin, end) of the enumeration i < j (i outer): their sum; the unit of multi-GPU sharding."""
        total = C.c_double()
        self._check(self.lib.ecc_direct_evaluate_range(self.h, int(begin), int(end), _ptr(cost_image, _F32), C.byref(total)))
        return total.value

    def direct_partition(self, n_parts):
        """Bounds of n_parts contiguous pair ranges of equal work (epipolar planes)."""
        bounds = np.zeros(n_parts + 1, np.int64)
        self._check(self.lib.ecc_direct_partition(self.h, int(n_parts), _ptr(bounds, _I64)))
        return bounds

    def direct_evaluate_pair(self, i, j, kappas=None):
        """computeForImagePair: dict(value, kappas, samples0, samples1); kappas given = the caller's plane angles."""
        n, v = C.c_int(), C.c_double()
        if kappas is not None:
            k = np.ascontiguousarray(kappas, np.float32).copy()
            m = k.shape[0]
            s0, s1 = np.zeros(m, np.float32), np.zeros(m, np.float32)
            self._check(self.lib.ecc_direct_evaluate_pair(self.h, int(i), int(j), m, m, _ptr(k, _F32), _ptr(s0, _F32), _ptr(s1, _F32),
                                                          C.byref(n), C.byref(v)))
            return dict(value=v.value, kappas=k, samples0=s0, samples1=s1)
        self._check(self.lib.ecc_direct_evaluate_pair(self.h, int(i), int(j), 0, 0, None, None, None, C.byref(n), None))
        m = n.value
        k, s0, s1 = np.zeros(m, np.float32), np.zeros(m, np.float32), np.zeros(m, np.float32)
        self._check(self.lib.ecc_direct_evaluate_pair(self.h, int(i), int(j), 0, m, _ptr(k, _F32), _ptr(s0, _F32), _ptr(s1, _F32),
                                                      C.byref(n), C.byref(v)))
        return dict(value=v.value, kappas=k, samples0=s0, samples1=s1)

    def direct_pair_geometry(self, i, j):
        """What computeForImagePair prepares for its kernel: dict(kappas, lines0, lines1, fbcc0, fbcc1, dkappa)."""
        n, dk = C.c_int(), C.c_double()
        self._check(self.lib.ecc_direct_pair_geometry(self.h, int(i), int(j), 0, None, None, None, None, None, C.byref(n), C.byref(dk)))
        m = n.value
        out = dict(kappas=np.zeros(m, np.float32), lines0=np.zeros((m, 3), np.float32), lines1=np.zeros((m, 3), np.float32),
                   fbcc0=np.zeros((m, 8), np.float32), fbcc1=np.zeros((m, 8), np.float32))
        self._check(self.lib.ecc_direct_pair_geometry(self.h, int(i), int(j), m, _ptr(out["kappas"], _F32), _ptr(out["lines0"], _F32),
                                                      _ptr(out["lines1"], _F32), _ptr(out["fbcc0"], _F32), _ptr(out["fbcc1"], _F32),
                                                      C.byref(n), C.byref(dk)))
        out["dkappa"] = dk.value
        return out

    def direct_line_integrals(self, image, lines, fbcc=None):
        """cuda_computeLineIntegrals: lines [m, >= 3] float32, fbcc [m, >= 6] float32 or None -> [m] float32 (host)."""
        lines = np.ascontiguousarray(lines, np.float32)
        m = lines.shape[0]
        out = np.zeros(m, np.float32)
        if fbcc is not None:
            fbcc = np.ascontiguousarray(fbcc, np.float32)
        self._check(self.lib.ecc_direct_line_integrals(self.h, int(image), _ptr(lines, _F32), m, lines.shape[1], _ptr(fbcc, _F32),
                                                       0 if fbcc is None else fbcc.shape[1], _ptr(out, _F32)))
        return out

    def pair_maps(self, idx4=None, n_views=None):
        """The reference's K01 records (16 floats per pair) as the pair kernels compute them; all pairs when idx4 is None."""
        n_pairs = n_views * (n_views - 1) // 2 if idx4 is None else idx4.shape[0]
        K = np.zeros((n_pairs, 16), np.float32)
        self._check(self.lib.ecc_pair_maps(self.h, _ptr(idx4, _I32), 0 if idx4 is None else n_pairs, _ptr(K, _F32)))
        return K

    def pair_sample_counts(self, n_views):
        counts = np.zeros(n_views * (n_views - 1) // 2, np.int32)
        self._check(self.lib.ecc_pair_sample_counts(self.h, _ptr(counts, _I32)))
        return counts

    def partition_pairs(self, n_parts):
        bounds = np.zeros(n_parts + 1, np.int64)
        self._check(self.lib.ecc_partition_pairs(self.h, int(n_parts), _ptr(bounds, _I64)))
        return bounds

    # -- multi-GPU team (include/ecc_b200.h "Multi-GPU team"): peer-mapped blocks, no collective on the data path
    TEAM_HANDLE_BYTES = 64

    def team_create(self, rank, world, n_total, n_alpha, n_t):
        """Allocates this rank's block; returns its handle (bytes) for the other processes."""
        handle = np.zeros(self.TEAM_HANDLE_BYTES, np.uint8)
        self._check(self.lib.ecc_team_create(self.h, int(rank), int(world), int(n_total), int(n_alpha), int(n_t), _ptr(handle, _U8)))
        self._team_shape = (int(n_total), int(n_t), int(n_alpha))
        return handle.tobytes()

    def team_connect(self, handles):
        """handles: the world handles in rank order (bytes each)."""
        blob = np.frombuffer(b"".join(handles), np.uint8).copy()
        self._check(self.lib.ecc_team_connect(self.h, _ptr(blob, _U8)))

    def team_connect_pointers(self, blocks):
        """blocks: the ranks' block addresses (team_block()[0]) when all ranks live in this process."""
        arr = (C.c_void_p * len(blocks))(*[int(b) for b in blocks])
        self._check(self.lib.ecc_team_connect_pointers(self.h, C.addressof(arr)))

    def team_block(self):
        """(address of the own block, address of the dtrs inside it)."""
        b, d = C.c_void_p(), C.c_void_p()
        self._check(self.lib.ecc_team_block(self.h, C.byref(b), C.byref(d)))
        return b.value, d.value

    def team_dtrs(self):
        """The block's Radon intermediates as a torch tensor view (n_total, n_t, n_alpha); memory owned by the context."""
        import torch
        _, d = self.team_block()
        n, n_t, n_a = self._team_shape

        class _Mem:  # __cuda_array_interface__ carrier
            pass
        m = _Mem()
        m.__cuda_array_interface__ = {"shape": (n, n_t, n_a), "typestr": "<f4", "data": (d, False), "version": 2}
        m._owner = self
        return torch.as_tensor(m, device=torch.device("cuda", torch.cuda.current_device()))

    def team_destroy(self):
        self._check(self.lib.ecc_team_destroy(self.h))

    def team_radon_compute(self, images, first, n_u, n_v, filter=FILTER_DERIVATIVE, post=POST_IDENTITY, interp=INTERP_TEXTURE):
        n_local = 0 if images is None else images.shape[0]
        self._check(self.lib.ecc_team_radon_compute(self.h, _ptr(images, _F32) if n_local else None, int(first), n_local, n_u, n_v,
                                                    filter, post, interp))

    def team_radon_shard(self, n_total, world, rank):
        """Sharding in quads of projections for the static-split engine: (first, count, (lo_num, hi_num, den))."""
        v = [C.c_int() for _ in range(5)]
        self._check(self.lib.ecc_team_radon_shard(int(n_total), int(world), int(rank), *[C.byref(x) for x in v]))
        return v[0].value, v[1].value, (v[2].value, v[3].value, v[4].value)

    def team_radon_compute_part(self, images, first, part, n_u, n_v, filter=FILTER_DERIVATIVE, post=POST_IDENTITY, interp=INTERP_HYBRID_STATIC):
        """images: projections [first, first + count) of team_radon_shard; part: its (lo_num, hi_num, den)."""
        n_local = 0 if images is None else images.shape[0]
        self._check(self.lib.ecc_team_radon_compute_part(self.h, _ptr(images, _F32) if n_local else None, int(first), n_local, int(part[0]),
                                                         int(part[1]), int(part[2]), n_u, n_v, filter, post, interp))

    def team_set_radon_intermediates(self, n_u, n_v, is_derivative=True):
        """setRadonIntermediates with the team block's dtrs (borrowed, zero copy)."""
        _, d = self.team_block()
        n, n_t, n_alpha = self._team_shape
        sa, st = self.radon_bin_sizes(n_u, n_v, n_alpha, n_t)
        self._check(self.lib.ecc_set_radon_intermediates(self.h, d, n, n_alpha, n_t, sa, st, n_u, n_v, int(is_derivative)))

    def team_evaluate(self, cost_image=None, want_mean=True):
        m = C.c_double()
        self._check(self.lib.ecc_team_evaluate(self.h, _ptr(cost_image, _F32), C.byref(m) if want_mean else None))
        return m.value if want_mean else None

    def team_barrier(self):
        self._check(self.lib.ecc_team_barrier(self.h))

    # -- synthetic data
    def synth_projections(self, Ps, n_u, n_v, ellipsoids, images, cos_weight=True, zero_border=True):
        Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
        ell = np.ascontiguousarray(ellipsoids, np.float64).reshape(-1, 7)
        self._check(self.lib.ecc_synth_projections(self.h, _ptr(Ps, _F64), Ps.shape[0], n_u, n_v, _ptr(ell, _F64), ell.shape[0],
                                                   int(cos_weight), int(zero_border), _ptr(images, _F32)))
        return images

    # -- instrumentation
    def profile_enable(self, on=True):
        self._check(self.lib.ecc_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.lib.ecc_profile_reset(self.h))

    def profile_get(self, family):
        ms, n = C.c_double(), C.c_longlong()
        self._check(self.lib.ecc_profile_get(self.h, family.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value


def make_circular_trajectory(n_proj, sid, sdd, n_u, n_v, max_angle_deg, pixel_spacing):
    """ProjTable::makeCircularTrajectory (HeaderOnly/Utils/Projtable.hxx:138-165); (n,12) col-major."""
    Ps = np.zeros((n_proj, 12), np.float64)
    _lib.load().ecc_make_circular_trajectory(n_proj, sid, sdd, n_u, n_v, max_angle_deg, pixel_spacing, _ptr(Ps))
    return Ps


def camera_intrinsics(P):
    """(focal length in px, principal point u, v) of a projection matrix: K(0,0), K(0,2), K(1,2) of
    Geometry::getCameraIntrinsics (ProjectionMatrix.cpp:27-67)."""
    P = np.ascontiguousarray(P, np.float64).reshape(12)
    f, u, v = C.c_double(), C.c_double(), C.c_double()
    _lib.load().ecc_camera_intrinsics(_ptr(P), C.addressof(f), C.addressof(u), C.addressof(v))
    return f.value, u.value, v.value


def similarity_2d(x):
    """ModelSimilarity2D::getInstance (LibProjectiveGeometry/Models/ModelSimilarity2D.hxx:52-72): x = translation u, v,
    rotation, scale -> 3x3."""
    x = np.asarray(x, np.float64)
    H = np.eye(3)
    if x[2] != 0:
        H[:2, :2] = [[np.cos(x[2]), -np.sin(x[2])], [np.sin(x[2]), np.cos(x[2])]]
    H[0, 2], H[1, 2] = x[0], x[1]
    if x[3] != 0:
        H[:2, :2] *= 1.0 + x[3]
    return H


def similarity_3d(x):
    """ModelSimilarity3D::getInstance (ModelSimilarity3D.hxx:64-87): x = translation X, Y, Z, rotation about X, Y, Z
    (R = Rx Ry Rz), scale -> 4x4."""
    x = np.asarray(x, np.float64)
    T = np.eye(4)
    if x[3] != 0 or x[4] != 0 or x[5] != 0:
        cx, sx, cy, sy, cz, sz = np.cos(x[3]), np.sin(x[3]), np.cos(x[4]), np.sin(x[4]), np.cos(x[5]), np.sin(x[5])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
        T[:3, :3] = Rx @ Ry @ Rz
    T[:3, 3] = x[:3]
    if x[6] != 0:
        T[:3, :3] *= 1.0 + x[6]
    return T


def camera_similarity_2d3d(P, x):
    """ModelCameraSimilarity2D3D::getInstance (ModelCameraSimilarity2D3D.hxx:89-92): P' = H2D(x[:4]) P T3D(x[4:]);
    P and the result are 12 doubles, column-major 3x4."""
    P = np.asarray(P, np.float64).reshape(4, 3).T
    return (similarity_2d(x[:4]) @ P @ similarity_3d(x[4:])).T.reshape(12)


def model_similarity_2d(x):
    """ModelSimilarity2D::getInstance as the LIBRARY computes it on host and device (own sine / cosine, no fused
    multiply-add; similarity_2d above is the independent numpy restatement).  Returns 3x3."""
    x = np.ascontiguousarray(x, np.float64).reshape(4)
    H = np.zeros(9, np.float64)
    _lib.load().ecc_model_similarity_2d(_ptr(x), _ptr(H))
    return H.reshape(3, 3).T.copy()


def model_similarity_3d(x):
    """ModelSimilarity3D::getInstance as the library computes it.  Returns 4x4."""
    x = np.ascontiguousarray(x, np.float64).reshape(7)
    T = np.zeros(16, np.float64)
    _lib.load().ecc_model_similarity_3d(_ptr(x), _ptr(T))
    return T.reshape(4, 4).T.copy()


def model_camera_similarity_2d3d(P, x):
    """ModelCameraSimilarity2D3D::getInstance as the library computes it on host and device: 12 doubles, column-major."""
    P = np.ascontiguousarray(P, np.float64).reshape(12)
    x = np.ascontiguousarray(x, np.float64).reshape(11)
    out = np.zeros(12, np.float64)
    _lib.load().ecc_model_camera_similarity_2d3d(_ptr(P), _ptr(x), _ptr(out))
    return out


def model_calibration_correction(geom4, x7):
    """ModelFDCTCalibrationCorrection::getTransforms as the library computes it: geom4 = mean principal point u, v, source-
    isocentre and source-detector distance; x7 = translation u, v, yaw, pitch, roll, delta SID, delta SDD.
    Returns 25 doubles: H (3x3) then T (4x4), column-major -- one instance of evaluate_batch_transforms."""
    g = np.ascontiguousarray(geom4, np.float64).reshape(4)
    x = np.ascontiguousarray(x7, np.float64).reshape(7)
    out = np.zeros(25, np.float64)
    H, T = out[:9], out[9:]
    _lib.load().ecc_model_calibration_correction(_ptr(g), _ptr(x), H.ctypes.data, T.ctypes.data)
    return out


def model_normalize(P):
    """Geometry::normalizeProjectionMatrix with the library's bits; returns a new 12-vector."""
    out = np.ascontiguousarray(P, np.float64).reshape(12).copy()
    _lib.load().ecc_model_normalize(_ptr(out))
    return out


def derive_views_host(Ps):
    """(P^+)^T and source positions as the metric uses them (fp32), computed on the host."""
    Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12)
    A = np.zeros((Ps.shape[0], 12), np.float32)
    Cs = np.zeros((Ps.shape[0], 4), np.float32)
    _lib.load().ecc_derive_views_host(_ptr(Ps), Ps.shape[0], _ptr(A), _ptr(Cs))
    return A, Cs


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class RadonIntermediate:
    """Mirror of EpipolarConsistency::RadonIntermediate (RadonIntermediate.h:18-128).

    Constructed from a projection image (computes right away, RadonIntermediate.cpp:17-31) or from an
    existing dtr image plus its geometry (RadonIntermediate.cpp:69-80,105-123)."""

    Derivative, Ramp, NoFilter = FILTER_DERIVATIVE, FILTER_RAMP, FILTER_NONE
    Identity, SquareRoot, Logarithm = POST_IDENTITY, POST_SQRT, POST_LOG

    def __init__(self, projection=None, size_alpha=0, size_t=0, filter=FILTER_DERIVATIVE, post_process=POST_IDENTITY,
                 *, dtr=None, original_size=None, interp=INTERP_TEXTURE, ctx=None):
        self.ctx = ctx or default_context()
        self.m_filter = filter
        self._cpu = None
        if projection is not None:
            import torch
            img = projection
            if isinstance(img, np.ndarray):
                img = torch.from_numpy(np.ascontiguousarray(img, np.float32)).cuda()
            self.n_y, self.n_x = img.shape
            self.n_alpha, self.n_t = int(size_alpha), int(size_t)
            self._gpu = self.ctx.radon_compute(img[None], self.n_alpha, self.n_t, filter, post_process, interp)[0]
        else:
            self.replaceRadonIntermediateData(dtr, original_size, filter)
        self.m_bin_size_angle, self.m_bin_size_distance = Context.radon_bin_sizes(self.n_x, self.n_y, self.n_alpha,
                                                                                  self.n_t)

    @classmethod
    def _from_device(cls, tensor, n_x, n_y, filter, ctx):
        self = cls.__new__(cls)
        self.ctx, self.m_filter, self._cpu, self._gpu = ctx, filter, None, tensor
        self.n_t, self.n_alpha = tensor.shape
        self.n_x, self.n_y = n_x, n_y
        self.m_bin_size_angle, self.m_bin_size_distance = Context.radon_bin_sizes(n_x, n_y, self.n_alpha, self.n_t)
        return self

    def replaceRadonIntermediateData(self, dtr, original_size, filter=FILTER_DERIVATIVE):
        import torch
        self._cpu = np.ascontiguousarray(dtr, np.float32)
        self._gpu = torch.from_numpy(self._cpu).cuda()
        self.n_t, self.n_alpha = self._cpu.shape
        self.n_x, self.n_y = original_size
        self.m_filter = filter

    def getFilter(self):
        return self.m_filter

    def isDerivative(self):
        return self.m_filter == FILTER_DERIVATIVE

    def readback(self, gpu_memory_only=False):
        if not gpu_memory_only:
            self._cpu = self._gpu.cpu().numpy()

    def data(self):
        """CPU image (n_t rows x n_alpha); valid after readback()."""
        return self._cpu

    def getTexture(self):
        """The reference returns its texture wrapper; callers use it to make the dtr GPU resident."""
        return self._gpu

    def getRadonBinNumber(self, dim):
        return self.n_t if dim else self.n_alpha

    def getOriginalImageSize(self, dim):
        return self.n_y if dim else self.n_x

    def getRadonBinSize(self, dim=1):
        return self.m_bin_size_distance if dim else self.m_bin_size_angle


def compute_radon_intermediates(images, size_alpha, size_t, filter=FILTER_DERIVATIVE, post_process=POST_IDENTITY,
                                interp=INTERP_TEXTURE, ctx=None):
    """Batched form of the RadonIntermediate image constructor: one launch chain for all projections.
    images: (n, n_v, n_u) numpy or torch-cuda.  Returns a list of RadonIntermediate views into one
    contiguous device tensor (which MetricRadonIntermediate then borrows without copying)."""
    import torch
    ctx = ctx or default_context()
    if isinstance(images, np.ndarray):
        images = torch.from_numpy(np.ascontiguousarray(images, np.float32)).cuda()
    n, n_v, n_u = images.shape
    block = ctx.radon_compute(images, size_alpha, size_t, filter, post_process, interp)
    return [RadonIntermediate._from_device(block[k], n_u, n_v, filter, ctx) for k in range(n)]


class MetricRadonIntermediate:
    """Mirror of EpipolarConsistency::MetricRadonIntermediate
    (EpipolarConsistencyRadonIntermediate.h:21-106, .cpp:41-322) incl. the Metric base interface."""

    def __init__(self, Ps=None, dtrs=None, ctx=None):
        self.ctx = ctx or Context()
        self.Ps = np.zeros((0, 12))
        self.dtrs = []
        if Ps is not None:
            self.setProjectionMatrices(Ps)
        if dtrs is not None:
            self.setRadonIntermediates(dtrs)

    # -- Metric base
    def setObjectRadius(self, radius_mm=0.0):
        self.ctx.set_object_radius(radius_mm)
        return self

    def getObjectRadius(self):
        return self.ctx.get_object_radius()

    def setEpipolarPlaneStep(self, dkappa_rad=0.0):
        self.ctx.set_epipolar_plane_step(dkappa_rad)
        return self

    setdKappa = setEpipolarPlaneStep

    def setProjectionMatrices(self, Ps):
        self.Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12).copy()
        self.ctx.set_projection_matrices(self.Ps)
        return self

    def getProjectionMatrices(self):
        return self.Ps

    def getNumberOfProjetions(self):  # [sic], as in the reference
        return self.Ps.shape[0]

    def useCorrelation(self, corr=True):
        self.ctx.use_correlation(corr)
        return self

    def setInterpolation(self, interp):
        self.ctx.set_interpolation(interp)
        return self

    # -- Radon intermediates
    def setRadonIntermediates(self, dtrs):
        import torch
        self.dtrs = list(dtrs)
        d0 = self.dtrs[0]
        tensors = [d._gpu for d in self.dtrs]
        # zero-copy when the dtrs are consecutive views of one block (compute_radon_intermediates)
        stride = d0.n_t * d0.n_alpha * 4
        base = tensors[0].data_ptr()
        contiguous = all(t.data_ptr() == base + k * stride and t.is_contiguous() for k, t in enumerate(tensors))
        if contiguous and tensors[0]._base is not None and tensors[0]._base.dim() == 3:
            block = tensors[0]._base[:len(tensors)]
        else:
            block = torch.stack(tensors).contiguous()
        self._block = block
        self.ctx.set_radon_intermediates(block, d0.n_x, d0.n_y, d0.isDerivative(),
                                         step_alpha=d0.getRadonBinSize(0), step_t=d0.getRadonBinSize(1))
        return self

    def getRadonIntermediates(self):
        return self.dtrs

    # -- evaluation
    def evaluate(self, arg=None, out=None):
        """evaluate()                 -> mean over all pairs
        evaluate(cost_image)          -> same, n*n float32 image gets entry [j, i] = pair (i<j)
        evaluate(views: set)          -> all pairs among the listed views
        evaluate(indices (k,4), out)  -> explicit (P0,P1,dtr0,dtr1) list"""
        if arg is None:
            return self.ctx.evaluate(None)
        if isinstance(arg, (set, frozenset)):
            v = sorted(arg)
            idx = np.array([(a, b, a, b) for ia, a in enumerate(v) for b in v[ia + 1:]], np.int32).reshape(-1, 4)
            return self.ctx.evaluate_indices(idx, out)
        arr = arg
        if isinstance(arr, np.ndarray) and arr.dtype == np.float32 and arr.ndim == 2 and arr.shape[0] == arr.shape[1] \
                and arr.shape[0] == self.Ps.shape[0] and out is None:
            return self.ctx.evaluate(arr)
        idx = np.ascontiguousarray(arr, np.int32).reshape(-1, 4)
        return self.ctx.evaluate_indices(idx, out)

    def evaluateBatch(self, Ps_sets, indices=None, out=None):
        """New capability: score K projection-matrix sets in one launch (means per set)."""
        idx = None if indices is None else np.ascontiguousarray(indices, np.int32).reshape(-1, 4)
        return self.ctx.evaluate_batch(np.ascontiguousarray(Ps_sets, np.float64), idx, out)


class MetricDirect:
    """Mirror of EpipolarConsistency::MetricDirect (EpipolarConsistencyDirect.h:27-60, .cpp:214-270): epipolar consistency
    straight from the projection images.  evaluate() returns the SUM over the pairs, as the reference's does."""

    def __init__(self, Ps=None, Is=None, ctx=None):
        self.ctx = ctx or Context()
        self.Ps = np.zeros((0, 12))
        self.n_images = 0
        if Ps is not None:
            self.setProjectionMatrices(Ps)
        if Is is not None:
            self.setProjectionImages(Is)

    def setObjectRadius(self, radius_mm=0.0):
        self.ctx.set_object_radius(radius_mm)
        return self

    def getObjectRadius(self):
        return self.ctx.get_object_radius()

    def setEpipolarPlaneStep(self, dkappa_rad=0.0):
        self.ctx.set_epipolar_plane_step(dkappa_rad)
        return self

    def setProjectionMatrices(self, Ps):
        self.Ps = np.ascontiguousarray(Ps, np.float64).reshape(-1, 12).copy()
        self.ctx.set_projection_matrices(self.Ps)
        return self

    def getProjectionMatrices(self):
        return self.Ps

    def setProjectionImages(self, Is):
        """Is: [n, n_v, n_u] float32 array / tensor (host or device)."""
        self.ctx.direct_set_images(Is)
        self.n_images = Is.shape[0]
        return self

    def getNumberOfProjetions(self):  # [sic]
        return self.n_images

    def setFanBeamConsistency(self, fbcc=True):
        self.ctx.direct_set_fan_beam(fbcc)
        return self

    def evaluate(self, out=None):
        return self.ctx.direct_evaluate(out)

    def evaluateForImagePair(self, i, j, kappas=None):
        """Returns (value, redundant_samples0, redundant_samples1, kappas)."""
        r = self.ctx.direct_evaluate_pair(i, j, kappas)
        return r["value"], r["samples0"], r["samples1"], r["kappas"]
