"""epipolarconsistency_b200 -- B200-native (sm_100a) Epipolar-Consistency hot path.

Radon-intermediate computation and the all-pairs / subset / batched ECC metric of
aaichert/EpipolarConsistency, as hand-written CUDA kernels behind a C ABI (include/ecc_b200.h,
lib/libecc_b200.so).  `api` mirrors the reference's MetricRadonIntermediate / RadonIntermediate
interface; `distributed` shards the path over the GPUs of one node with torch.distributed (NCCL).
"""
from ._lib import LIB_PATH, EccLibraryMissing, load  # noqa: F401
from .api import (  # noqa: F401
    FILTER_DERIVATIVE, FILTER_NONE, FILTER_RAMP, INTERP_EXACT, INTERP_TEXTURE, POST_IDENTITY, POST_LOG, POST_SQRT,
    Context, EccError, MetricDirect, MetricRadonIntermediate, RadonIntermediate, compute_radon_intermediates,
    make_circular_trajectory,
)
