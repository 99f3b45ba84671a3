"""ctypes binding of libreni_b200.so (the C ABI declared in include/reni_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is
called without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# RENI_B200_LIB overrides the path (tools/variant_time.py times alternative builds of the same ABI)
LIB_PATH = os.environ.get("RENI_B200_LIB") or os.path.join(_HERE, "lib", "libreni_b200.so")

FLAG_SAVE_FOR_BACKWARD = 1
FLAG_NEED_DW = 2
FLAG_LOSS = 4
FLAG_FILM = 8
FLAG_FILM_PERMAP = 16
FLAG_PREPARE_WEIGHTS = 32
FLAG_TILE_MAJOR_BWD = 64
FLAG_LAYER_MAJOR_BWD = 128
FLAG_FWD_SINGLE_TERM = 256
FLAG_FWD_TWO_TERM = 512
FLAG_GRID_DIRECTIONS = 1024
FLAG_GRID_SINEWEIGHT = 2048

EQUIVARIANCE = {"None": 0, "SO2": 1, "SO3": 2}


class RENIConfig(C.Structure):
    """reni_config_t -- mirrors RENIAutoDecoder's constructor arguments (RENI.py:91-104)."""

    _fields_ = [
        ("ndims", C.c_int32),
        ("equivariance", C.c_int32),
        ("hidden_features", C.c_int32),
        ("hidden_layers", C.c_int32),
        ("out_features", C.c_int32),
        ("last_layer_linear", C.c_int32),
        ("output_activation", C.c_int32),
        ("first_omega_0", C.c_float),
        ("hidden_omega_0", C.c_float),
    ]


class AdamSegment(C.Structure):
    """reni_adam_segment_t."""

    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64)]


class RENILibraryError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); every symbol declared in include/reni_b200.h
_vp, _i32, _i64, _u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
_cfgp = C.POINTER(RENIConfig)
SIGNATURES = {
    "reni_abi_version": (_i32, []),
    "reni_strerror": (C.c_char_p, [_i32]),
    "reni_in_features": (_i64, [_cfgp]),
    "reni_workspace_bytes": (_i64, [_cfgp, _i64, _i64, _i32]),
    "reni_prepare_weights": (_i32, [_cfgp, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _vp]),
    "reni_forward": (_i32, [_cfgp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _vp]),
    "reni_backward": (_i32, [_cfgp, _vp, _vp, _i64, C.POINTER(_vp), _i64, _i64, _vp, _vp, _vp, C.POINTER(_vp),
                             C.POINTER(_vp), _vp, _i64, _i32, _vp]),
    "reni_loss_forward_backward": (_i32, [_cfgp, _vp, _vp, _i64, C.POINTER(_vp), C.POINTER(_vp), _i64, _i64, _vp, _vp,
                                          _i64, C.c_float, C.c_float, _i32, _vp, _vp, _vp, C.POINTER(_vp),
                                          C.POINTER(_vp), _vp, _i64, _i32, _vp]),
    "reni_film_forward": (_i32, [_cfgp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _i64, _i32, _vp]),
    "reni_film_backward": (_i32, [_cfgp, _vp, _vp, _i64, C.POINTER(_vp), C.POINTER(_vp), _i64, _i64, _vp, _vp, _vp, _vp,
                                  C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _i32, _vp]),
    "reni_film_prepare_maps": (_i32, [_cfgp, _vp, C.POINTER(_vp), C.POINTER(_vp), _i64, _i64, _vp, _i64, _i32, _vp]),
    "reni_film_map_scratch_bytes": (_i64, [C.POINTER(C.c_int32), _i32, _i64]),
    "reni_film_map_forward": (_i32, [_cfgp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int32), _i32, _i64,
                                     _vp, _vp, _vp, _i64, _vp]),
    "reni_film_map_acts_bytes": (_i64, [C.POINTER(C.c_int32), _i32, _i64]),
    "reni_film_map_forward_train": (_i32, [_cfgp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_int32),
                                           _i32, _i64, _vp, _vp, _vp, _i64, _vp]),
    "reni_film_map_backward": (_i32, [_cfgp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(C.c_int32), _i32, _i64, _vp, _vp,
                                      _vp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _i64, _vp]),
    "reni_film_loss_forward_backward": (_i32, [_cfgp, _vp, _vp, _vp, _i64, C.POINTER(_vp), C.POINTER(_vp), _i64, _i64, _vp,
                                               _vp, _i64, C.c_float, _i32, _vp, _vp, _vp, _vp, C.POINTER(_vp),
                                               C.POINTER(_vp), _vp, _i64, _i32, _vp]),
    "reni_adam_step": (_i32, [C.POINTER(AdamSegment), _i32, _vp, C.c_double, C.c_double, C.c_double, C.c_double, _vp]),
    "reni_vad_sample": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "reni_vad_backward": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, C.c_float, C.c_float, _vp, _vp, _vp, _vp]),
    "reni_envmap_shade_forward": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, C.c_float, C.c_float, C.c_float, _vp, _vp]),
    "reni_envmap_shade_backward": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, C.c_float, C.c_float, C.c_float, _vp, _vp]),
    "reni_allreduce_flag_bytes": (_i64, []),
    "reni_allreduce": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, C.c_float, _vp, _vp, _vp]),
    "reni_debug_set_phase_events": (_i32, [C.POINTER(_vp), _i32]),
    "reni_debug_last_cuda_error": (C.c_char_p, []),
    "reni_phase_bits": (_i32, []),
    "reni_debug_set_trace": (_i32, [_vp]),
    "reni_debug_set_overlap": (_i32, [_i32, _i32]),
    "reni_probe_remote_tx": (_i32, [_vp, _u32, _vp, _i32, _vp]),
    "reni_selftest_umma": (_i32, [_vp, _u32, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp]),
    "reni_selftest_umma2": (_i32, [_vp, _u32, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp]),
}


def load() -> C.CDLL:
    """Load libreni_b200.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RENILibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  reni_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name):
            continue  # reported by missing_symbols(); compute wrappers check before calling
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def missing_symbols():
    lib = load()
    return [n for n in SIGNATURES if not hasattr(lib, n)]


def check(code: int, what: str = "reni_b200") -> None:
    if code != 0:
        msg = load().reni_strerror(int(code)).decode()
        if int(code) == -4:  # RENI_ERR_CUDA: say which runtime error
            msg += ": " + load().reni_debug_last_cuda_error().decode()
        raise RENILibraryError(f"{what} failed: {msg} (code {code})")
