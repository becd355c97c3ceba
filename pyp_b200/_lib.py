"""ctypes binding of libcspb200.so (include/cspb200.h).

The library is the product: there is no Python / CPU fallback.  Importing this module fails
loudly when the shared object has not been built (run ``python -c 'import __graft_entry__ as g;
g.build()'`` or ``pyp_b200/csrc/build.sh``), and every call raises :class:`CspbError` when the
CUDA side reports an error (for instance no sm_100 GPU).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSPB_LIB") or os.path.join(_HERE, "libcspb200.so")  # CSPB_LIB: alternative build (kernel tuning experiments)

HOST, DEVICE = 0, 1

# .cistem projection row — src/pyp/inout/metadata/cistem_star_file.py:596-628 (order),
# :127-185 (types); 128 bytes, little endian.
ROW_DTYPE = np.dtype(
    [
        ("position_in_stack", "<u4"),
        ("psi", "<f4"), ("theta", "<f4"), ("phi", "<f4"),
        ("x_shift", "<f4"), ("y_shift", "<f4"),
        ("defocus_1", "<f4"), ("defocus_2", "<f4"), ("defocus_angle", "<f4"), ("phase_shift", "<f4"),
        ("image_is_active", "<i4"),
        ("occupancy", "<f4"), ("logp", "<f4"), ("sigma", "<f4"), ("score", "<f4"),
        ("pixel_size", "<f4"), ("voltage_kv", "<f4"), ("cs_mm", "<f4"), ("amplitude_contrast", "<f4"),
        ("beam_tilt_x", "<f4"), ("beam_tilt_y", "<f4"), ("image_shift_x", "<f4"), ("image_shift_y", "<f4"),
        ("original_x", "<f4"), ("original_y", "<f4"),
        ("imind", "<i4"), ("pind", "<i4"), ("tind", "<i4"), ("rind", "<i4"), ("find", "<i4"),
        ("fshift_x", "<f4"), ("fshift_y", "<f4"),
    ]
)
assert ROW_DTYPE.itemsize == 128


class RefineCfg(C.Structure):
    _fields_ = [
        ("box", C.c_int32), ("pad", C.c_int32),
        ("pixel_size", C.c_float), ("mask_radius", C.c_float),
        ("low_res_limit", C.c_float), ("high_res_limit", C.c_float), ("signed_cc_limit", C.c_float),
        ("search_mask_radius", C.c_float), ("search_high_res", C.c_float), ("angular_step", C.c_float),
        ("best_matches", C.c_int32),
        ("search_range_x", C.c_float), ("search_range_y", C.c_float),
        ("defocus_range", C.c_float), ("defocus_step", C.c_float),
        ("global_search", C.c_int32), ("local_refine", C.c_int32),
        ("refine_psi", C.c_int32), ("refine_theta", C.c_int32), ("refine_phi", C.c_int32),
        ("refine_x", C.c_int32), ("refine_y", C.c_int32), ("refine_defocus", C.c_int32),
        ("apply_mask", C.c_int32), ("normalize", C.c_int32), ("invert_contrast", C.c_int32),
        ("whiten", C.c_int32), ("symmetry_order", C.c_int32), ("local_iterations", C.c_int32),
        ("use_priors", C.c_int32), ("prior_mean_x", C.c_float), ("prior_mean_y", C.c_float),
        ("prior_var_x", C.c_float), ("prior_var_y", C.c_float),
        ("optimizer", C.c_int32),
        ("reserved", C.c_int32 * 1),
    ]


class ReconCfg(C.Structure):
    _fields_ = [
        ("box", C.c_int32), ("pad", C.c_int32),
        ("pixel_size", C.c_float), ("mask_radius", C.c_float), ("resolution_limit", C.c_float),
        ("score_bfactor", C.c_float), ("score_weighting", C.c_int32), ("score_threshold", C.c_float),
        ("normalize", C.c_int32), ("invert_contrast", C.c_int32), ("per_particle_split", C.c_int32),
        ("average_score", C.c_float),
        ("reserved", C.c_int32 * 8),
    ]


# `_extended.cistem` blocks — cistem_star_file.py:247-248
PARTICLE_DTYPE = np.dtype([("pind", "<i4"), ("shift_x", "<f4"), ("shift_y", "<f4"), ("shift_z", "<f4"), ("psi", "<f4"), ("theta", "<f4"),
                           ("phi", "<f4"), ("x_position_3d", "<f4"), ("y_position_3d", "<f4"), ("z_position_3d", "<f4"), ("score", "<f4"), ("occ", "<f4")])
TILT_DTYPE = np.dtype([("tind", "<i4"), ("rind", "<i4"), ("shift_x", "<f4"), ("shift_y", "<f4"), ("angle", "<f4"), ("axis", "<f4")])
assert PARTICLE_DTYPE.itemsize == 48 and TILT_DTYPE.itemsize == 24


class CspCfg(C.Structure):
    """csp_* keys of .pyp_config.toml (include/cspb200.h cspb_csp_cfg)."""
    _fields_ = [
        ("mode", C.c_int32), ("window_min", C.c_int32), ("window_max", C.c_int32), ("iterations", C.c_int32),
        ("random_evals", C.c_int32), ("grid_search", C.c_int32),
        ("angle_step", C.c_float), ("shift_step", C.c_float),
        ("tol_particle_psi", C.c_float), ("tol_particle_theta", C.c_float), ("tol_particle_phi", C.c_float),
        ("tol_particle_shift", C.c_float),
        ("tol_tilt_angle", C.c_float), ("tol_tilt_axis", C.c_float), ("tol_tilt_shift", C.c_float), ("tol_defocus", C.c_float),
        ("seed", C.c_uint32), ("min_projections", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class SelectCfg(C.Structure):
    """Score shaping between refine3d and reconstruct3d (include/cspb200.h cspb_select_cfg; scores.py:300-761)."""
    _fields_ = [
        ("cutoff", C.c_float), ("mindef", C.c_float), ("maxdef", C.c_float), ("firstframe", C.c_int32), ("lastframe", C.c_int32),
        ("mintilt", C.c_float), ("maxtilt", C.c_float), ("minazh", C.c_float), ("maxazh", C.c_float),
        ("minscore", C.c_float), ("maxscore", C.c_float), ("renumber", C.c_int32), ("threshold_override", C.c_float),
        ("reserved", C.c_int32 * 4),
    ]


class CspbError(RuntimeError):
    pass


_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
_SIGNATURES = {
    "cspb_abi_version": (C.c_int, []),
    "cspb_device_count": (C.c_int, []),
    "cspb_create": (_i, [_i, C.POINTER(_vp)]),
    "cspb_destroy": (_i, [_vp]),
    "cspb_last_error": (C.c_char_p, [_vp]),
    "cspb_sync": (_i, [_vp]),
    "cspb_stream": (_i, [_vp, C.POINTER(_vp)]),
    "cspb_launch_count": (_i64, [_vp]),
    "cspb_profile_enable": (_i, [_vp, _i]),
    "cspb_profile_get": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(_i64)]),
    "cspb_profile_count_loads": (_i, [_vp, _i]),
    "cspb_profile_get_loads": (_i, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "cspb_refine_cfg_default": (_i, [C.POINTER(RefineCfg), _i, _f]),
    "cspb_refine_configure": (_i, [_vp, C.POINTER(RefineCfg)]),
    "cspb_refine_reset_images": (_i, [_vp]),
    "cspb_refine_set_ring_weights": (_i, [_vp, _vp, _i]),
    "cspb_refine_phase_sum": (_i, [_vp, _vp, _i, _vp]),
    "cspb_refine_set_focus_mask": (_i, [_vp, C.c_float, C.c_float, C.c_float, C.c_float]),
    "cspb_set_reference": (_i, [_vp, _vp, _i, _i]),
    "cspb_set_symmetry": (_i, [_vp, _vp, _i]),
    "cspb_refine_load_images": (_i, [_vp, _vp, _i, _i, _i]),
    "cspb_refine_keep_spectra": (_i, [_vp, _i]),
    "cspb_pipeline_batches": (_i, [_i, _i, _i, _vp, _i]),
    "cspb_refine_num_images": (_i, [_vp]),
    "cspb_refine_score": (_i, [_vp, _vp, _i, _vp]),
    "cspb_refine_score_poses": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp]),
    "cspb_refine_score_grad": (_i, [_vp, _vp, _i, _i, _vp]),
    "cspb_refine_matching_projections": (_i, [_vp, _vp, _i, _vp]),
    "cspb_refine_set_search_grid": (_i, [_vp, _vp, _i]),
    "cspb_refine_run": (_i, [_vp, _vp, _i, _vp, C.POINTER(_i64)]),
    "cspb_refine_run_device": (_i, [_vp, _vp, _i, C.POINTER(_i64)]),
    "cspb_refine_get_noise_curve": (_i, [_vp, _vp, _i]),
    "cspb_refine_set_noise_curve": (_i, [_vp, _vp, _i]),
    "cspb_csp_cfg_default": (_i, [C.POINTER(CspCfg)]),
    "cspb_csp_run": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, C.POINTER(CspCfg), _i, _i, C.POINTER(_i64)]),
    "cspb_csp_compose": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _f, _f, _vp]),
    "cspb_csp_extract": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _i]),
    "cspb_spa_extract": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _f, _vp, _i, _i]),
    "cspb_refine_reconstruct": (_i, [_vp, _vp, _vp, _i, _i, C.POINTER(_i64)]),
    "cspb_select_cfg_default": (_i, [C.POINTER(SelectCfg)]),
    "cspb_select_scores": (_i, [_vp, _vp, _i, _vp, C.POINTER(SelectCfg), _i, C.POINTER(C.c_double)]),
    "cspb_class_occupancies": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _i]),
    "cspb_global_weights": (_i, [_vp, _vp, _i, _i, _vp, _i, C.POINTER(_i)]),
    "cspb_refine_select_reconstruct": (_i, [_vp, _vp, _vp, _i, C.POINTER(SelectCfg), C.POINTER(_i64), C.POINTER(C.c_double)]),
    "cspb_recon_cfg_default": (_i, [C.POINTER(ReconCfg), _i, _f]),
    "cspb_recon_begin": (_i, [_vp, C.POINTER(ReconCfg)]),
    "cspb_recon_insert": (_i, [_vp, _vp, _vp, _i, _i]),
    "cspb_recon_insert_weighted": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cspb_recon_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i64)]),
    "cspb_recon_device_ptr": (_i, [_vp, _i, C.POINTER(_vp)]),
    "cspb_recon_get_dump": (_i, [_vp, _i, _vp, _i]),
    "cspb_recon_add_dump": (_i, [_vp, _i, _vp, _i]),
    "cspb_recon_finalize": (_i, [_vp, _f, _f, _vp, _vp, _vp, _vp, _i, _i]),
    "cspb_recon_end": (_i, [_vp]),
    "cspb_fft2_r2c": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "cspb_fft2_c2r": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "cspb_cufft2_r2c": (_i, [_vp, _vp, _vp, _i, _i, _i]),
    "cspb_ctf_image": (_i, [_vp, _vp, _i, _vp]),
    "cspb_project": (_i, [_vp, _f, _f, _f, _vp]),
    "cspb_band_counts": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "cspb_gather_peak": (_i, [_vp, C.c_size_t, _i, C.POINTER(_f)]),
    "cspb_wave_units": (_i, [_vp]),
}

# every symbol include/cspb200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))


def load():
    """Load libcspb200.so and attach the prototypes.  Raises if the library is not built."""
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA engine has not been built. "
            "Run pyp_b200/csrc/build.sh (nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def ptr(a):
    """Raw pointer of a numpy array (must be C-contiguous) or an int device pointer."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return C.c_void_p(a.ctypes.data)
