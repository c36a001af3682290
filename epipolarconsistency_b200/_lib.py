"""Loader for libecc_b200.so (the C ABI declared in include/ecc_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing, importing a compute entry
point raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C epipolarconsistency_b200/csrc`.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# ECC_B200_LIB: development override (a build of the same sources with other compile-time knobs)
LIB_PATH = os.environ.get("ECC_B200_LIB") or os.path.join(PKG_DIR, "lib", "libecc_b200.so")

c_ctx = C.c_void_p
c_vp = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/ecc_b200.h declares
SIGNATURES = {
    "ecc_version": (C.c_int, []),
    "ecc_create": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "ecc_destroy": (None, [c_ctx]),
    "ecc_last_error": (C.c_char_p, [c_ctx]),
    "ecc_set_stream": (C.c_int, [c_ctx, c_vp]),
    "ecc_synchronize": (C.c_int, [c_ctx]),
    "ecc_radon_compute": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "ecc_radon_bin_sizes": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ecc_radon_calibrate_split": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "ecc_radon_set_split": (C.c_int, [c_ctx, C.c_int]),
    "ecc_radon_num_samples": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "ecc_set_radon_intermediates": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]),
    "ecc_set_radon_intermediate_pointers": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]),
    "ecc_device_alloc": (C.c_int, [c_ctx, C.c_size_t, C.POINTER(c_vp)]),
    "ecc_device_free": (C.c_int, [c_ctx, c_vp]),
    "ecc_copy": (C.c_int, [c_ctx, c_vp, c_vp, C.c_size_t]),
    "ecc_texture_create": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(c_vp)]),
    "ecc_texture_destroy": (C.c_int, [c_ctx, C.c_ulonglong, c_vp]),
    "ecc_texture_readback": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, c_vp]),
    "ecc_set_projection_matrices": (C.c_int, [c_ctx, c_vp, C.c_int]),
    "ecc_update_projection_matrix": (C.c_int, [c_ctx, C.c_int, c_vp]),
    "ecc_get_derived_views": (C.c_int, [c_ctx, c_vp, c_vp]),
    "ecc_derive_views_host": (None, [c_vp, C.c_int, c_vp, c_vp]),
    "ecc_set_object_radius": (C.c_int, [c_ctx, C.c_double]),
    "ecc_get_object_radius": (C.c_int, [c_ctx, C.POINTER(C.c_double)]),
    "ecc_set_epipolar_plane_step": (C.c_int, [c_ctx, C.c_double]),
    "ecc_set_interpolation": (C.c_int, [c_ctx, C.c_int]),
    "ecc_use_correlation": (C.c_int, [c_ctx, C.c_int]),
    "ecc_evaluate": (C.c_int, [c_ctx, c_vp, C.POINTER(C.c_double)]),
    "ecc_evaluate_range": (C.c_int, [c_ctx, C.c_longlong, C.c_longlong, c_vp, C.POINTER(C.c_double)]),
    "ecc_evaluate_indices": (C.c_int, [c_ctx, c_vp, C.c_int, c_vp, C.POINTER(C.c_double)]),
    "ecc_update_and_evaluate": (C.c_int, [c_ctx, C.c_int, c_vp, c_vp, C.c_int, c_vp, C.POINTER(C.c_double)]),
    "ecc_track_info": (C.c_int, [c_ctx, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "ecc_evaluate_batch": (C.c_int, [c_ctx, c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp]),
    "ecc_pair_signals": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                  C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ecc_pair_maps": (C.c_int, [c_ctx, c_vp, C.c_int, c_vp]),
    "ecc_model_similarity_2d": (None, [c_vp, c_vp]),
    "ecc_model_similarity_3d": (None, [c_vp, c_vp]),
    "ecc_model_transform": (None, [c_vp, c_vp, c_vp, c_vp]),
    "ecc_model_camera_similarity_2d3d": (None, [c_vp, c_vp, c_vp]),
    "ecc_evaluate_batch_params": (C.c_int, [c_ctx, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "ecc_evaluate_batch_transforms": (C.c_int, [c_ctx, c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp, C.c_int, c_vp, c_vp]),
    "ecc_transform_expand": (C.c_int, [c_ctx, c_vp, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp]),
    "ecc_model_calibration_correction": (None, [c_vp, c_vp, c_vp, c_vp]),
    "ecc_model_normalize": (None, [c_vp]),
    "ecc_model_expand": (C.c_int, [c_ctx, c_vp, c_vp, C.c_int, C.c_int, c_vp, c_vp]),
    "ecc_direct_set_images": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int]),
    "ecc_direct_set_image_pointers": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int]),
    "ecc_direct_set_fan_beam": (C.c_int, [c_ctx, C.c_int]),
    "ecc_direct_set_reference_clip": (C.c_int, [c_ctx, C.c_int]),
    "ecc_direct_evaluate": (C.c_int, [c_ctx, c_vp, C.POINTER(C.c_double)]),
    "ecc_direct_evaluate_range": (C.c_int, [c_ctx, C.c_longlong, C.c_longlong, c_vp, C.POINTER(C.c_double)]),
    "ecc_direct_partition": (C.c_int, [c_ctx, C.c_int, c_vp]),
    "ecc_direct_evaluate_pair": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, C.POINTER(C.c_int),
                                          C.POINTER(C.c_double)]),
    "ecc_direct_pair_geometry": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, C.POINTER(C.c_int),
                                          C.POINTER(C.c_double)]),
    "ecc_direct_line_integrals": (C.c_int, [c_ctx, C.c_int, c_vp, C.c_int, C.c_int, c_vp, C.c_int, c_vp]),
    "ecc_pair_sample_counts": (C.c_int, [c_ctx, c_vp]),
    "ecc_partition_pairs": (C.c_int, [c_ctx, C.c_int, c_vp]),
    "ecc_team_create": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp]),
    "ecc_team_connect": (C.c_int, [c_ctx, c_vp]),
    "ecc_team_connect_pointers": (C.c_int, [c_ctx, c_vp]),
    "ecc_team_block": (C.c_int, [c_ctx, C.POINTER(c_vp), C.POINTER(c_vp)]),
    "ecc_team_destroy": (C.c_int, [c_ctx]),
    "ecc_team_radon_compute": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ecc_team_radon_shard": (C.c_int, [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 5),
    "ecc_team_radon_compute_part": (C.c_int, [c_ctx, c_vp] + [C.c_int] * 10),
    "ecc_team_evaluate": (C.c_int, [c_ctx, c_vp, C.POINTER(C.c_double)]),
    "ecc_team_barrier": (C.c_int, [c_ctx]),
    "ecc_make_circular_trajectory": (None, [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, c_vp]),
    "ecc_camera_intrinsics": (None, [c_vp, c_vp, c_vp, c_vp]),
    "ecc_preprocess_defaults": (None, [c_vp]),
    "ecc_preprocess": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
    "ecc_synth_projections": (C.c_int, [c_ctx, c_vp, C.c_int, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, c_vp]),
    "ecc_profile_enable": (C.c_int, [c_ctx, C.c_int]),
    "ecc_profile_reset": (C.c_int, [c_ctx]),
    "ecc_profile_get": (C.c_int, [c_ctx, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
}

_lib = None


class EccLibraryMissing(RuntimeError):
    pass


def load():
    """Returns the ctypes handle of libecc_b200.so with all signatures set.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EccLibraryMissing(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `make -C epipolarconsistency_b200/csrc`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
