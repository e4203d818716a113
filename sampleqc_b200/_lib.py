"""Loader of the product library libsampleqc_cuda.so (built in-tree by ``__graft_entry__.build()`` /
``python -m sampleqc_b200.build``).  There is no CPU fallback: if the library is missing, or no CUDA
device is usable, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SQC_LIB") or os.path.join(_HERE, "libsampleqc_cuda.so")          # SQC_LIB: a development variant of the same library
_LIB = None

# every symbol include/sampleqc_cuda.h declares
EXPORTS = (
    "sqc_fit_robust", "sqc_fit_mle", "sqc_outlier_maha", "sqc_init_clusts",
    "sqc_session_create", "sqc_session_run", "sqc_session_fetch", "sqc_session_stats", "sqc_session_profile",
    "sqc_session_destroy", "sqc_xchg_export", "sqc_xchg_release",
    "sqc_trim", "sqc_abi_version", "sqc_device_count", "sqc_status_name", "sqc_device_qchisq", "sqc_device_rkeys", "sqc_device_rkeys_share", "sqc_device_expneg", "sqc_mt_jump_poly", "sqc_fp64_peak", "sqc_plan_shards",
    "sqc_resident_create", "sqc_resident_init_clusts", "sqc_resident_subsample", "sqc_resident_fit", "sqc_resident_outliers", "sqc_resident_destroy", "sqc_mmd_fn", "sqc_mmd_pairs",
)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m sampleqc_b200.build` "
            "(sampleqc_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    fa, fo = C.POINTER(A.SqcFitArgs), C.POINTER(A.SqcFitOut)
    for name in ("sqc_fit_robust", "sqc_fit_mle"):
        f = getattr(L, name)
        f.argtypes = [fa, fo, C.c_char_p, C.c_size_t]
        f.restype = C.c_int
    L.sqc_outlier_maha.argtypes = [C.POINTER(A.SqcMahaArgs), A.c_double_p, A.c_uint8_p, A.c_double_p, C.c_char_p, C.c_size_t]
    L.sqc_outlier_maha.restype = C.c_int
    L.sqc_init_clusts.argtypes = [C.POINTER(A.SqcInitArgs), A.c_double_p, A.c_double_p, A.c_int32_p, C.c_char_p, C.c_size_t]
    L.sqc_init_clusts.restype = C.c_int
    L.sqc_session_create.argtypes = [fa, C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.sqc_session_create.restype = C.c_int
    L.sqc_session_run.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.sqc_session_run.restype = C.c_int
    L.sqc_session_fetch.argtypes = [C.c_void_p, fo, C.c_char_p, C.c_size_t]
    L.sqc_session_fetch.restype = C.c_int
    L.sqc_session_stats.argtypes = [C.c_void_p, C.POINTER(A.SqcRunStats)]
    L.sqc_session_stats.restype = C.c_int
    L.sqc_session_profile.argtypes = [C.c_void_p, C.c_int]
    L.sqc_session_profile.restype = C.c_int
    L.sqc_session_destroy.argtypes = [C.c_void_p]
    L.sqc_session_destroy.restype = None
    L.sqc_xchg_export.argtypes = [C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
    L.sqc_xchg_export.restype = C.c_int
    L.sqc_xchg_release.argtypes = [C.c_int]
    L.sqc_xchg_release.restype = C.c_int
    if hasattr(L, "sqc_trim"):
        L.sqc_trim.argtypes = [C.c_int]
        L.sqc_trim.restype = C.c_int
    L.sqc_abi_version.restype = C.c_int
    L.sqc_device_count.restype = C.c_int
    L.sqc_status_name.argtypes = [C.c_int]
    L.sqc_status_name.restype = C.c_char_p
    L.sqc_device_qchisq.argtypes = [A.c_double_p, A.c_double_p, A.c_double_p, C.c_int32, C.c_int, C.c_char_p, C.c_size_t]
    L.sqc_device_qchisq.restype = C.c_int
    if hasattr(L, "sqc_device_expneg"):
        L.sqc_device_expneg.argtypes = [A.c_double_p, A.c_double_p, C.c_int32, C.c_int, C.c_char_p, C.c_size_t]
        L.sqc_device_expneg.restype = C.c_int
    L.sqc_device_rkeys.argtypes = [C.c_uint32, C.c_int64, A.c_int32_p, C.c_int64, C.c_int, C.c_char_p, C.c_size_t]
    L.sqc_device_rkeys.restype = C.c_int
    L.sqc_device_rkeys_share.argtypes = [C.c_uint32, C.c_int64, C.c_int, C.c_int, C.c_int, A.c_int32_p, C.POINTER(C.c_int64), A.c_int32_p, C.c_int, C.c_char_p, C.c_size_t]
    L.sqc_device_rkeys_share.restype = C.c_int
    L.sqc_fp64_peak.argtypes = [C.c_int, A.c_double_p, C.c_char_p, C.c_size_t]
    L.sqc_fp64_peak.restype = C.c_int
    L.sqc_resident_create.argtypes = [A.c_double_p, A.c_int32_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.sqc_resident_create.restype = C.c_int
    L.sqc_resident_init_clusts.argtypes = [C.c_void_p, A.c_double_p, A.c_double_p, A.c_double_p, C.c_int32, C.c_int32, A.c_double_p, A.c_double_p, A.c_int32_p, C.c_char_p, C.c_size_t]
    L.sqc_resident_init_clusts.restype = C.c_int
    L.sqc_resident_subsample.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64, A.c_double_p, C.c_char_p, C.c_size_t]
    L.sqc_resident_subsample.restype = C.c_int
    L.sqc_resident_fit.argtypes = [C.c_void_p, fa, C.c_int, C.c_int32, fo, C.c_char_p, C.c_size_t]
    L.sqc_resident_fit.restype = C.c_int
    L.sqc_resident_outliers.argtypes = [C.c_void_p, C.c_double, A.c_double_p, A.c_uint8_p, A.c_double_p, C.c_char_p, C.c_size_t]
    L.sqc_resident_outliers.restype = C.c_int
    L.sqc_resident_destroy.argtypes = [C.c_void_p]
    L.sqc_resident_destroy.restype = None
    L.sqc_mmd_fn.argtypes = [A.c_double_p, C.c_int64, A.c_double_p, C.c_int64, C.c_int32, C.c_double, A.c_double_p, C.c_int32, C.c_char_p, C.c_size_t]
    L.sqc_mmd_fn.restype = C.c_int
    L.sqc_mmd_pairs.argtypes = [C.POINTER(A.SqcMmdArgs), A.c_double_p, A.c_double_p, C.c_char_p, C.c_size_t]
    L.sqc_mmd_pairs.restype = C.c_int
    L.sqc_plan_shards.argtypes = [A.c_int32_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int64), A.c_int32_p]
    L.sqc_plan_shards.restype = C.c_int
    L.sqc_mt_jump_poly.argtypes = [C.c_uint64, C.c_void_p]
    L.sqc_mt_jump_poly.restype = C.c_int
    if L.sqc_abi_version() != A.SQC_ABI_VERSION:
        raise RuntimeError("libsampleqc_cuda.so ABI version does not match sampleqc_b200._abi")
    _LIB = L
    return L


def check(status: int, err) -> None:
    if status != 0:
        raise A.SampleQCError(status, err.value.decode(errors="replace"))
