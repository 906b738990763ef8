"""ctypes binding of libirgs_b200.so -- the C ABI declared in include/irgs_b200.h.

The product path has no fallback: if the CUDA library is missing or fails to load, every entry point raises.
"""
import ctypes
import os

from . import build as _build

_LIB = None

_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float

# name -> (restype, argtypes); mirrors include/irgs_b200.h one to one (tests/test_abi.py checks the symbol list)
PROTOTYPES = {
    "irgs_last_error": (ctypes.c_char_p, []),
    "irgs_version": (_i32, []),
    "irgs_tracer_create": (_i32, [ctypes.POINTER(_vp), _i32]),
    "irgs_tracer_destroy": (_i32, [_vp]),
    "irgs_build_from_proxy": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "irgs_refit_from_proxy": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "irgs_build_from_surfels": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp]),
    "irgs_refit_from_surfels": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp]),
    "irgs_num_surfels": (_i64, [_vp]),
    "irgs_get_bounds": (_i32, [_vp, _vp, _vp, _vp]),
    "irgs_intersection_test": (_i32, [_vp, _i64] + [_vp] * 7 + [_f32, _vp, _vp]),
    "irgs_trace_forward": (_i32, [_vp, _i64, _i32, _i32, _i32] + [_vp] * 9 + [_vp] * 5 + [_vp, _vp, _i32, _f32, _f32, _i32, _vp]),
    "irgs_trace_backward": (_i32, [_vp, _i64, _i32, _i32, _i32] + [_vp] * 9 + [_vp] * 5 + [_vp, _vp, _i32] + [_vp] * 5
                            + [_vp] * 4 + [_f32, _f32, _i32, _vp]),
    "irgs_incident_rays": (_i32, [_vp, _vp, _vp, _vp]),
    "irgs_trace_forward_incident": (_i32, [_vp, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_vp] * 5 + [_vp, _vp, _i32, _f32, _f32, _i32, _vp]),
    "irgs_trace_backward_incident": (_i32, [_vp, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_vp] * 5 + [_vp, _vp, _i32] + [_vp] * 5
                                     + [_vp] * 6 + [_f32, _f32, _i32, _vp]),
    "irgs_shade_forward": (_i32, [_vp] * 8 + [_f32, _vp, _vp]),
    "irgs_shade_backward": (_i32, [_vp] * 8 + [_f32] + [_vp] * 6),
    "irgs_env_lookup_forward": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "irgs_env_lookup_backward": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "irgs_unpack_grads": (_i32, [_vp, _i64, _i32] + [_vp] * 6 + [_vp]),
    "irgs_trace_forward_host": (_i32, [_vp, _i64, _i32, _i32, _i32] + [_vp] * 9 + [_vp] * 5 + [_f32, _f32, _i32, _i64]),
    "irgs_trace_fwd_bwd_host": (_i32, [_vp, _i64, _i32, _i32, _i32] + [_vp] * 9 + [_vp] * 5 + [_i64] + [_vp] * 3
                                + [_vp] * 2 + [_f32, _f32, _i32, _i64]),
    "irgs_trace_fwd_bwd_incident_host": (_i32, [_vp, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_vp] * 5 + [_i64] + [_vp] * 3
                                         + [_vp] * 2 + [_f32, _f32, _i32, _i64, _vp]),
    "irgs_camera_rays": (_i32, [_vp, _vp, _vp, _vp]),
    "irgs_trace_forward_camera": (_i32, [_vp, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_vp] * 5 + [_vp, _vp, _i32, _f32, _f32, _i32, _vp]),
    "irgs_trace_backward_camera": (_i32, [_vp, _vp, _i32, _i32, _i32] + [_vp] * 7 + [_vp] * 5 + [_vp, _vp, _i32] + [_vp] * 5
                                   + [_vp] * 4 + [_f32, _f32, _i32, _vp]),
    "irgs_relight_hit": (_i32, [_i64, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp]),
    "irgs_relight_combine": (_i32, [_i64, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _vp, _vp, _vp]),
    "irgs_surfel_frames": (_i32, [_vp, _vp, _vp, ctypes.POINTER(_f32), _i64, _vp, _vp, _vp, _vp]),
    "irgs_unpack_grads_params": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, ctypes.POINTER(_f32)] + [_vp] * 5 + [_vp]),
    "irgs_normalize_outputs": (_i32, [_i64, _i32, _f32] + [_vp] * 10 + [_vp]),
    "irgs_normalize_outputs_backward": (_i32, [_i64, _i32, _f32] + [_vp] * 10 + [_vp]),
    "irgs_stride_multiplier": (_i64, [_i64]),
    "irgs_launch_count": (_i64, []),
    "irgs_reset_launch_count": (None, []),
    "irgs_set_option": (_i32, [_vp, ctypes.c_char_p, _i64]),
    "irgs_get_info": (_i64, [_vp, ctypes.c_char_p]),
    "irgs_set_stats": (_i32, [_vp, _i32]),
    "irgs_get_stats": (_i32, [_vp, ctypes.POINTER(_i64)]),
}


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the sources are newer) the CUDA library.  Raises RuntimeError when unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    override = os.environ.get("IRGS_LIB")   # tuning experiments only: an alternative build of the same sources
    if override:
        path = override
    try:
        if not override and _build.needs_build():
            _build.build()
    except Exception as e:  # nvcc missing on a box that only has the prebuilt .so is fine; a missing .so is not
        if not os.path.exists(path):
            raise RuntimeError(f"irgs_b200: CUDA library {path} is missing and could not be built ({e}); "
                               "run `python -m irgs_b200.build`. There is no CPU fallback.") from e
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise RuntimeError(f"irgs_b200: cannot load {path}: {e}. There is no CPU fallback.") from e
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("irgs_b200: " + load().irgs_last_error().decode("utf-8", "replace"))
