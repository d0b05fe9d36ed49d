"""ctypes binding of libmgb.so (include/mgb.h).  No torch types cross this boundary: only raw device or
host pointers, sizes and the current CUDA stream handle.

The library has no CPU fallback; if it is missing it is built (nvcc must exist), and if that fails the
import error propagates - the product path fails loudly rather than degrading.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

MGB_OK = 0
MGB_VERSION = 202      # must equal mgb_version() (csrc/runtime.cu): bumped whenever a signature changes
BASIS_MONOMIAL = 0
BASIS_CHEBYSHEV = 1
MAX_FACTORS = 16
MAX_TERMS = 256


class SeriesDesc(C.Structure):
    _fields_ = [("n_factors", C.c_int), ("factor_width", C.c_int), ("n_terms", C.c_int), ("basis", C.c_int),
                ("scale", C.c_double), ("shift", C.c_double), ("lo", C.c_double), ("hi", C.c_double),
                ("pair_square", C.c_int)]


_P = C.c_void_p
_LL = C.c_longlong
_D = C.c_double
_I = C.c_int
_DESC = C.POINTER(SeriesDesc)

# name -> (restype, argtypes); must list every symbol include/mgb.h declares (tests/test_abi.py checks)
PROTOTYPES = {
    "mgb_last_error": (C.c_char_p, []),
    "mgb_version": (_I, []),
    "mgb_launch_count": (_LL, []),
    "mgb_series_gram_fwd": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _LL, _P]),
    "mgb_series_gram_fwd_batched": (_I, [_DESC, _P, _LL, _LL, _LL, _P, _LL, _LL, _LL, _P, _P, _LL, _LL, _LL, _P]),
    "mgb_series_gram_diag": (_I, [_DESC, _P, _LL, _P, _LL, _LL, _P, _P, _P]),
    "mgb_series_gram_bwd": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _LL, _LL, _P, _P, _P, _P]),
    "mgb_so3_to_quaternion": (_I, [_P, _LL, _LL, _P, _P, _P]),
    "mgb_series_gram_dx": (_I, [_DESC, _I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _LL, _LL, _P, _P, _P]),
    "mgb_son_basis_size": (_LL, [_I, _I]),
    "mgb_son_gram": (_I, [_I, _I, _P, _LL, _LL, _P, _LL, _LL, _P, _I, _P, _LL, _P]),
    "mgb_hyperbolic_gram_fwd": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _I, _D, _D, _I, _P, _LL, _P]),
    "mgb_hyperbolic_gram_bwd_coef": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _I, _D, _D, _I, _P, _LL, _P, _P]),
    "mgb_hyperbolic_gram_bwd_tval": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _I, _D, _D, _I, _P, _LL, _P, _P]),
    "mgb_hyperbolic_heat_nodes": (_I, [_I, _I, _P, _LL, _P, _I, _D, _D, _I, _P, _P]),
    "mgb_euclidean_gram_fwd": (_I, [_P, _LL, _LL, _P, _LL, _LL, _I, _P, _P, _I, _P, _LL, _P]),
    "mgb_euclidean_gram_bwd_coef": (_I, [_P, _LL, _LL, _P, _LL, _LL, _I, _P, _I, _P, _LL, _P, _P]),
    "mgb_spd2_gram_fwd": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P, _I, _P, _P, _I, _P, _LL, _P]),
    "mgb_spd2_gram_fwd_lattice": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P, _I, _P, _P, _I, _I, _I, _D, _D, _P, _LL, _P]),
    "mgb_spd2_gram_bwd_scales": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P, _I, _P, _P, _I, _P, _LL, _P, _P, _P]),
    "mgb_spd2_gram_bwd_coef": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _I, _P, _P, _I, _P, _LL, _P, _P]),
    "mgb_spd_decompose": (_I, [_I, _I, _P, _LL, _LL, _P, _LL, _P]),
    "mgb_radial_profile": (_I, [_I, _I, _P, _LL, _P, _P, _I, _D, _D, _I, _P, _P]),
    "mgb_radial_profile_bwd_coef": (_I, [_I, _I, _P, _LL, _P, _I, _D, _D, _I, _P, _P, _P]),
    "mgb_spd2_profile": (_I, [_I, _P, _P, _LL, _P, _P, _P, _I, _P, _P, _I, _P, _P]),
    "mgb_spd2_profile_bwd_coef": (_I, [_I, _P, _P, _LL, _P, _P, _I, _P, _P, _I, _P, _P, _P]),
    "mgb_radial_table_gram": (_I, [_I, _I, _P, _LL, _LL, _P, _LL, _LL, _I, _I, _I, _P, _P, _LL, _P]),
    "mgb_posterior_pack_bytes": (_LL, [_LL]),
    "mgb_posterior_pack": (_I, [_P, _LL, _LL, _P, _P, _P]),
    "mgb_posterior_retile": (_I, [_P, _LL, _LL, _LL, _P, _P]),
    "mgb_posterior_reduce_workspace_bytes": (_LL, [_LL, _LL]),
    "mgb_posterior_tiled_bytes": (_LL, [_LL, _LL]),
    "mgb_series_gram_tiled": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P]),
    "mgb_posterior_reduce": (_I, [_P, _LL, _LL, _P, _P, _D, _P, _P, _P, _LL, _P]),
    "mgb_posterior_workspace_bytes": (_LL, [_LL, _LL]),
    "mgb_series_posterior": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _D, _D, _P, _P, _P, _LL, _LL, _P]),
    "mgb_series_acquisition_ei": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _P, _LL, _LL, _P, _P, _LL, _P, _LL, _P, _D, _D, _D, _I,
                                       _P, _P, _P, _P, _P, _P]),
    "mgb_radial_acquisition_ei": (_I, [_I, _I, _P, _LL, _LL, _P, _LL, _P, _LL, _LL, _P, _P, _P, _D, _P, _LL, _P, _LL, _P,
                                       _D, _D, _D, _I, _P, _P, _P, _P]),
    "mgb_spd2_acquisition_ei": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P, _D, _P, _LL, _P, _LL, _P, _D, _D, _D, _I,
                                     _P, _P, _P]),
    "mgb_radial_acquisition_ei_step": (_I, [_I, _I, _P, _LL, _LL, _P, _LL, _P, _LL, _LL, _P, _P, _I, _D, _D, _I, _D, _P, _LL,
                                            _P, _LL, _P, _D, _D, _D, _I, _P, _LL, _P, _P, _P, _P]),
    "mgb_spd2_acquisition_ei_step": (_I, [_I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P, _I, _P, _P, _I, _D, _P, _LL, _P, _LL,
                                          _P, _D, _D, _D, _I, _P, _LL, _P, _P, _P]),
    "mgb_int8_available": (_I, []),
    "mgb_posterior_int8_pack_bytes": (_LL, [_LL]),
    "mgb_posterior_int8_pack": (_I, [_P, _LL, _LL, _P, _P]),
    "mgb_posterior_int8_workspace_bytes": (_LL, [_LL, _LL]),
    "mgb_posterior_int8_reduce": (_I, [_P, _LL, _LL, _LL, _P, _P, _D, _D, _P, _P, _P, _LL, _P]),
    "mgb_posterior_use_int8": (_I, [_P, _P, _P]),
    "mgb_posterior_is_int8": (_I, [_P]),
    "mgb_posterior_i8f_pack_bytes": (_LL, [_LL]),
    "mgb_posterior_i8f_pack": (_I, [_P, _LL, _LL, _P, _P]),
    "mgb_posterior_i8f_workspace_bytes": (_LL, [_LL, _LL]),
    "mgb_posterior_i8f_reduce": (_I, [_P, _LL, _LL, _LL, _P, _P, _D, _D, _P, _P, _P, _LL, _P]),
    "mgb_series_posterior_i8f_workspace_bytes": (_LL, [_LL, _LL]),
    "mgb_series_posterior_i8f": (_I, [_P, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P, _D, _D, _P, _P, _P, _LL, _P]),
    "mgb_posterior_i8f_planes_bytes": (_LL, [_LL, _LL]),
    "mgb_posterior_i8f_reduce_debug": (_I, [_P, _LL, _LL, _LL, _P, _P, _D, _D, _P, _P, _P, _LL, _P, _P]),
    "mgb_series_gram_i8_supported": (_I, [_DESC]),
    "mgb_series_gram_i8_workspace_bytes": (_LL, [_LL]),
    "mgb_series_gram_fwd_i8": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _LL, _P, _LL, _P]),
    "mgb_series_gram_fwd_i8_debug": (_I, [_DESC, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _LL, _P, _LL, _P, _P]),
    "mgb_expected_improvement": (_I, [_P, _P, _LL, _D, _I, _P, _P]),
    "mgb_pair_scalars": (_I, [_I, _I, _P, _LL, _LL, _P, _LL, _LL, _P, _P, _P]),
    "mgb_spectral_coef": (_I, [_I, _I, _I, _P, _P, _P, _P, _D, _P, _P, _P, _P, _P]),
    "mgb_series_mll_max_n": (_LL, [_I, _I]),
    "mgb_series_mll_single_cta_max_n": (_LL, [_I, _I]),
    "mgb_dense_max_n": (_LL, []),
    "mgb_dense_fit_workspace_bytes": (_LL, [_LL]),
    "mgb_dense_mll_workspace_bytes": (_LL, [_LL]),
    "mgb_dense_fit_large": (_I, [_P, _LL, _LL, _P, _P, _P, _P, _LL, _P, _P, _P, _P, _LL, _P]),
    "mgb_dense_mll_large": (_I, [_P, _LL, _LL, _P, _P, _P, _P, _LL, _P, _P, _LL, _P]),
    "mgb_series_mll_large": (_I, [_DESC, _P, _LL, _LL, _P, _P, _P, _P, _P, _P, _LL, _P]),
    "mgb_series_fit": (_I, [_DESC, _P, _LL, _LL, _P, _P, _P, _P, _P, _LL, _P, _P, _P, _P]),
    "mgb_series_mll": (_I, [_DESC, _P, _LL, _LL, _P, _P, _P, _P, _P, _P]),
    "mgb_dense_mll": (_I, [_P, _LL, _LL, _P, _P, _P, _P, _LL, _P, _P]),
    "mgb_dense_fit": (_I, [_P, _LL, _LL, _P, _P, _P, _P, _LL, _P, _P, _P, _P]),
    "mgb_circle_coef": (_I, [_I, _I, _P, _P, _I, _P, _P, _P, _P, _P]),
    "mgb_matern_t_grid": (_I, [_P, _D, _D, _I, _P, _P, _P]),
    "mgb_matern_t_weights": (_I, [_P, _P, _P, _I, _D, _P, _P, _P, _P, _P]),
    "mgb_series_gram_host": (_I, [_DESC, _P, _LL, _P, _LL, _P, _P, _I]),
    "mgb_series_posterior_host": (_I, [_DESC, _P, _LL, _P, _LL, _P, _P, _P, _D, _D, _P, _P, _LL, _I]),
    "mgb_posterior_create": (_I, [_DESC, _P, _LL, _P, _P, _P, _D, _D, _LL, _I, C.POINTER(_P)]),
    "mgb_posterior_create_from_device": (_I, [_DESC, _P, _LL, _LL, _P, _P, _D, _D, _LL, _I, C.POINTER(_P)]),
    "mgb_posterior_chunk": (_LL, [_P]),
    "mgb_posterior_eval_host": (_I, [_P, _P, _LL, _P, _P]),
    "mgb_posterior_eval_device": (_I, [_P, _P, _LL, _LL, _P, _P, _P]),
    "mgb_posterior_destroy": (_I, [_P]),
    "mgb_fp64_peak": (_I, [_I, _I, C.POINTER(_D), C.POINTER(_D)]),
}

_lib = None


class MgbError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the .so is absent) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()       # no-op when the content digest of sources + flags matches the stamp
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.mgb_version() != MGB_VERSION:
        raise MgbError("libmgb.so reports version %d, the Python binding expects %d (stale build?)" % (lib.mgb_version(), MGB_VERSION))
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != MGB_OK:
        msg = load().mgb_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError("%s: %s" % (what, msg))
        raise MgbError("%s failed (code %d): %s" % (what, rc, msg))


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def require_cuda_f64(*tensors):
    """The product path runs on the GPU only: anything else is an error, never a fallback."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise MgbError("materngabo_b200 kernels need CUDA tensors (no CPU fallback); got device %s" % t.device)
        if t.dtype != torch.float64:
            raise TypeError("materngabo_b200 kernels compute in float64; got %s" % t.dtype)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise MgbError("tensors on different devices: %s vs %s" % (dev, t.device))
    return dev
