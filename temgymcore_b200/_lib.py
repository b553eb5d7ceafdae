"""ctypes binding of ``libtemgym_b200.so`` (the C ABI of ``include/temgym_b200.h``).

The library is built in-tree by ``temgymcore_b200/csrc/build.sh`` (nvcc, sm_100a only).
Loading fails loudly when the library is missing; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TG_LIB_PATH") or os.path.join(_HERE, "libtemgym_b200.so")   # TG_LIB_PATH: experiment builds

TG_MAX_PEERS = 8
TG_MAX_COMPS = 24
TG_NPARAM = 48

TG_OP_PLANE, TG_OP_LENS, TG_OP_DEFLECTOR, TG_OP_BIPRISM = 0, 1, 2, 3
TG_OP_KRIVANEK, TG_OP_OFFSET, TG_OP_THICKLENS, TG_OP_ROTATOR = 4, 5, 6, 7
TG_F_NOPROP = 1
TG_F_DIST = 2
TG_JAC_NONE, TG_JAC_ABCD5, TG_JAC_FULL7 = 0, 1, 2
TG_METHOD = {"auto": 0, "sfu": 1, "tensor": 2, "tensor_tf32": 3, "tensor_4m": 4, "tensor_3m": 5,
             "tensor_binned": 6}
TG_OK, TG_EINVAL, TG_ECUDA, TG_ENOTSEPARABLE, TG_EUNSUPPORTED = 0, -1, -2, -3, -4


class tg_comp(C.Structure):
    _fields_ = [("op", C.c_int32), ("flags", C.c_int32), ("z", C.c_double),
                ("p", C.c_double * TG_NPARAM)]


class tg_model(C.Structure):
    _fields_ = [("n_comp", C.c_int32), ("reserved", C.c_int32),
                ("comp", tg_comp * TG_MAX_COMPS)]


class tg_ray_in(C.Structure):
    _fields_ = [("ptr", C.c_void_p * 7), ("value", C.c_double * 7)]


TG_MAX_PERRAY = 4


class tg_perray(C.Structure):
    _fields_ = [("n", C.c_int32), ("comp", C.c_int32 * TG_MAX_PERRAY), ("ptr", (C.c_void_p * 4) * TG_MAX_PERRAY)]


class tg_seed(C.Structure):
    _fields_ = [("comp", C.c_int32), ("slot", C.c_int32), ("lane", C.c_int32), ("reserved", C.c_int32),
                ("weight", C.c_double)]


TG_GRAD_LANES = 8
TG_MAX_SEEDS = 32


class TemGymError(RuntimeError):
    pass


_vp, _i64, _i32, _dp = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_double)

# every symbol include/temgym_b200.h declares, with its ctypes signature
SIGNATURES = {
    "tg_last_error": (C.c_char_p, []),
    "tg_abi_version": (_i32, []),
    "tg_device_count": (_i32, []),
    "tg_trace_f64": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(tg_ray_in), C.POINTER(_vp), _vp, _i32, _vp]),
    "tg_trace_perray_f64": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(tg_ray_in), C.POINTER(tg_perray), C.POINTER(_vp),
                                   _vp, _i32, _vp]),
    "tg_trace_f64_host": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(tg_ray_in), C.POINTER(_vp), _vp, _i32, _i32]),
    "tg_trace_grad_f64": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(tg_ray_in), C.POINTER(C.c_int32),
                                 C.POINTER(tg_seed), _i32, C.POINTER(_vp), _vp, _vp]),
    "tg_trace_jets_f64": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(tg_ray_in), _i32, C.POINTER(_vp), _vp, _vp,
                                 _vp, _vp]),
    "tg_krivanek_f64": (_i32, [_i64, _vp, _vp, _dp, _vp, _vp, _vp, _vp]),
    "tg_fibonacci_spiral_f64": (_i32, [_i64, C.c_double, C.c_double, _vp, _vp, _vp]),
    "tg_concentric_rings_count": (C.c_int64, [_i64, C.c_double]),
    "tg_concentric_rings_f64": (_i32, [_i64, C.c_double, _i64, _vp, _vp, _vp]),
    "tg_decompose_qinv_f64": (_i32, [_i64, _vp, _vp, _i32, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tg_transfer_rays_f64": (_i32, [_i64, _vp, _i32, _dp, _vp, _vp]),
    "tg_stem4d_backproject": (_i32, [C.POINTER(C.c_int), _dp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "tg_stem4d_indices": (_i32, [C.POINTER(C.c_int), _dp, _i32, _i32, _vp, _vp]),
    "tg_metres_to_pixels": (_i32, [_i64, _vp, _vp, _dp, _vp, _vp, _i32, _vp]),
    "tg_metres_to_pixels_host": (_i32, [_i64, _vp, _vp, _dp, _vp, _vp, _i32, _i32]),
    "tg_into_image_i64": (_i32, [_i64, _vp, _vp, _i32, _i32, _vp, _vp]),
    "tg_beamlet_coeffs_f64": (_i32, [_i64] + [_vp] * 13 + [_vp]),
    "tg_beamlet_coeffs_abcd_f64": (_i32, [_i64] + [_vp] * 10 + [_vp]),
    "tg_input_coeffs_f64": (_i32, [_i64] + [_vp] * 7 + [_vp]),
    "tg_gaussian_qinv_f64": (_i32, [_i64] + [_vp] * 5 + [_vp]),
    "tg_wave_numbers_f64": (_i32, [_i64] + [_vp] * 4 + [_vp]),
    "tg_field_sum_grid": (_i32, [_i64, _vp, _dp, _i32, _i32, _i32, _i32, _vp, _i32, _i32,
                                 C.POINTER(C.c_longlong), _vp]),
    "tg_field_sum_points": (_i32, [_i64, _vp, _i64, _vp, _vp, _i32, _vp]),
    "tg_field_sum_separable": (_i32, [_i64, _vp, _dp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "tg_field_sum": (_i32, [_i64, _vp, _dp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp]),
    "tg_make_gaussian_image_f64": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(_vp), _vp, _vp, _vp,
                                          _vp, _vp, _dp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp]),
    "tg_field_sum_verdict": (_i32, [_i64, _vp, _dp, _i32, _i32, _i32, C.POINTER(C.c_int), _vp]),
    "tg_gemm_tf32x3": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, C.c_longlong, _vp, C.c_longlong, _i32, _vp]),
    "tg_gemm_f16x3": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, C.c_longlong, _vp, C.c_longlong, _i32, _vp]),
    "tg_gemm_chunk_k": (_i32, []),
    "tg_cgemm3_f16x3": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, C.c_longlong, _vp, C.c_longlong, _i32, _vp]),
    "tg_gemm_schedule": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "tg_gemm_schedule_ragged": (_i32, [_i32, _vp, _i32, _vp, _i32, _vp, _i32]),
    "tg_binned_last_chunks": (_i32, []),
    "tg_peer_alloc": (_i32, [C.c_uint64, C.POINTER(_vp), _vp]),
    "tg_peer_open": (_i32, [_vp, C.POINTER(_vp)]),
    "tg_peer_close": (_i32, [_vp]),
    "tg_peer_free": (_i32, [_vp]),
    "tg_field_sum_peers": (_i32, [_i64, _vp, _dp, _i32, _i32, _i32, _i32, C.POINTER(_vp), _i32, _i32, _i32, _i32,
                                  _i32, _vp]),
    "tg_peer_barrier": (_i32, [C.POINTER(_vp), _i32, _i32, C.c_uint64, _vp]),
    "tg_peer_barrier_auto": (_i32, [C.POINTER(_vp), _i32, _i32, _vp, C.c_double, _vp]),
    "tg_make_gaussian_image_peers": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(_vp), _vp, _vp, _vp, _vp, _vp, _dp,
                                            _i32, _i32, _i32, _i32, C.POINTER(_vp), _i32, _i32, _i32, _i32, _i32,
                                            _vp]),
    "tg_make_gaussian_image_host": (_i32, [C.POINTER(tg_model), _i64, C.POINTER(_vp), _vp, _vp, _vp,
                                           _vp, _vp, _dp, _i32, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _i32]),
}

_lib = None


def load() -> C.CDLL:
    """Load the native library (once).  Raises TemGymError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TemGymError(
            f"{LIB_PATH} not found: build it with temgymcore_b200/csrc/build.sh "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.tg_abi_version() != 1:
        raise TemGymError("libtemgym_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != TG_OK:
        msg = load().tg_last_error()
        raise TemGymError(f"{what or 'temgym_b200'} failed (rc={rc}): "
                          f"{msg.decode(errors='replace') if msg else ''}")


def ptr_array(ptrs: Sequence[int | None]):
    arr = (_vp * len(ptrs))()
    for i, p in enumerate(ptrs):
        arr[i] = p if p else None
    return arr


def dbl_array(vals: Sequence[float]):
    return (C.c_double * len(vals))(*[float(v) for v in vals])
