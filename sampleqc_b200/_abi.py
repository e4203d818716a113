"""ctypes description of include/sampleqc_cuda.h (the C ABI) plus array marshalling helpers.

This module only *describes* the ABI; it loads nothing.  The product loader is
``sampleqc_b200._lib`` (libsampleqc_cuda.so); the test-only oracle loader lives in ``oracle/``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

SQC_ABI_VERSION = 2
SQC_IPC_HANDLE_BYTES = 64
SQC_N_KERNEL_FAMILIES = 8

SQC_OK = 0
SQC_ERR_BAD_ARG = 1
SQC_ERR_EMPTY_CLUSTER = 2
SQC_ERR_SUBSET_SIZE = 3
SQC_ERR_NOT_SPD = 4
SQC_ERR_SINGULAR = 5
SQC_ERR_NAN = 6
SQC_ERR_SELECT = 7
SQC_ERR_CUDA = -1
SQC_ERR_COMM = -2
SQC_ERR_NOMEM = -3

SQC_RNG_R_COMPAT = 0
SQC_RNG_COUNTER = 1

SQC_VARIANT_ROBUST = 0
SQC_VARIANT_MLE = 1

KERNEL_FAMILIES = ("estep", "partition", "mcd", "update", "rng", "init", "mle_pass", "final")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class SqcFitArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("D", C.c_int32), ("J", C.c_int32), ("K", C.c_int32),
        ("N", C.c_int64),
        ("x", c_double_p),
        ("groups", c_int32_p),
        ("init_z", c_int32_p),
        ("init_gamma", c_double_p),
        ("em_iters", C.c_int32),
        ("mcd_iters", C.c_int32),
        ("mcd_alpha", C.c_double),
        ("seed", C.c_uint32),
        ("track", C.c_int32),
        ("rng_mode", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32),
        ("cell_offset", C.c_int64),
        ("N_total", C.c_int64),
        ("peer_handles", C.c_void_p),
        ("n_gpus", C.c_int32),
    ]


class SqcFitOut(C.Structure):
    _fields_ = [
        ("mu_0", c_double_p), ("alpha_j", c_double_p), ("beta_k", c_double_p), ("sigma_k", c_double_p),
        ("z", c_int32_p), ("p_jk", c_double_p), ("like_1", c_double_p), ("like_2", c_double_p),
        ("gamma_i", c_double_p), ("p_k", c_double_p),
        ("track_alpha_j", c_double_p), ("track_beta_k", c_double_p),
        ("track_sigma_k", c_double_p), ("track_p_jk", c_double_p),
        ("n_iter", C.c_int32), ("mcd_iters_hit", C.c_int32), ("n_zero_like", C.c_int32),
        ("rng_draws", C.c_int64),
        ("rng_state", C.c_int32 * 625),
    ]


class SqcMahaArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("D", C.c_int32), ("J", C.c_int32), ("K", C.c_int32),
        ("N", C.c_int64),
        ("x", c_double_p), ("groups", c_int32_p),
        ("mu_0", c_double_p), ("alpha_j", c_double_p), ("beta_k", c_double_p), ("sigma_k", c_double_p),
        ("alpha", C.c_double),
        ("device", C.c_int32),
    ]


class SqcInitArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("D", C.c_int32), ("J", C.c_int32), ("K", C.c_int32),
        ("N", C.c_int64),
        ("x", c_double_p), ("groups", c_int32_p),
        ("pro", c_double_p), ("mean", c_double_p), ("sigma", c_double_p),
        ("device", C.c_int32),
    ]


class SqcMmdArgs(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("D", C.c_int32), ("n_samples", C.c_int32),
        ("mats", c_double_p), ("mat_rows", C.POINTER(C.c_int64)),
        ("n_pairs", C.c_int64),
        ("pair_i", c_int32_p), ("pair_j", c_int32_p),
        ("n_times", C.c_int32), ("subsample", C.c_int32),
        ("idx_i", c_int32_p), ("idx_j", c_int32_p),
        ("sigma", C.c_double),
        ("device", C.c_int32),
    ]


class SqcRunStats(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int32), ("n_gpu_launches", C.c_int32),
        ("loop_ms", C.c_double), ("init_ms", C.c_double),
        ("family_ms", C.c_double * SQC_N_KERNEL_FAMILIES),
        ("family_launches", C.c_int64 * SQC_N_KERNEL_FAMILIES),
        ("family_bytes", C.c_double * SQC_N_KERNEL_FAMILIES),
        ("mcd_passes", C.c_int64), ("mcd_csteps", C.c_int64),
    ]


class SampleQCError(RuntimeError):
    """Raised where the reference would raise an R error (END_RCPP, src/RcppExports.cpp:17,31)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[sqc status {status}] {message}")
        self.status = status
        self.message = message


def f64_colmajor(a, shape=None) -> np.ndarray:
    arr = np.asarray(a, dtype=np.float64)
    if shape is not None:
        arr = arr.reshape(shape, order="F") if arr.ndim == 1 else arr
    return np.asfortranarray(arr)


def as_i32(a) -> np.ndarray:
    """arma::uvec conversion of the export (src/RcppExports.cpp:21-22,40-41): values are truncated."""
    arr = np.asarray(a)
    if arr.dtype != np.int32:
        arr = arr.astype(np.int64).astype(np.int32) if arr.dtype.kind == "f" else arr.astype(np.int32)
    return np.ascontiguousarray(arr)


def ptr_d(a: np.ndarray | None):
    return a.ctypes.data_as(c_double_p) if a is not None else c_double_p()


def ptr_i(a: np.ndarray | None):
    return a.ctypes.data_as(c_int32_p) if a is not None else c_int32_p()


class FitBuffers:
    """Owns the caller-allocated host outputs of one fit and converts them to the Rcpp List shape."""

    def __init__(self, D, J, K, N, em_iters, variant, track=False, alloc=None):
        """alloc(shape, dtype=..., order=...) may hand out pinned host memory (bench.py's e2e leg)"""
        self.D, self.J, self.K, self.N, self.em_iters = D, J, K, N, max(int(em_iters), 0)
        self.variant, self.track = variant, bool(track)
        z = alloc or np.zeros
        self.mu_0 = z(D)
        self.alpha_j = z((J, D), order="F")
        self.beta_k = z((K, D), order="F")
        self.sigma_k = z((D, D, K), order="F")
        self.z = z(N, dtype=np.int32)
        self.p_jk = z((J, K), order="F")
        self.like_1 = z(self.em_iters)
        self.like_2 = z(self.em_iters)
        self.gamma_i = z((N, K), order="F") if variant == SQC_VARIANT_MLE else None
        self.p_k = z(K) if variant == SQC_VARIANT_MLE else None
        if self.track:
            I = self.em_iters
            self.track_alpha_j = z((J, D, I), order="F")
            self.track_beta_k = z((K, D, I), order="F")
            self.track_sigma_k = z((D * D, K, I), order="F")
            self.track_p_jk = z((J, K, I), order="F")
        else:
            self.track_alpha_j = self.track_beta_k = self.track_sigma_k = self.track_p_jk = None
        self.out = SqcFitOut()
        o = self.out
        o.mu_0, o.alpha_j, o.beta_k, o.sigma_k = ptr_d(self.mu_0), ptr_d(self.alpha_j), ptr_d(self.beta_k), ptr_d(self.sigma_k)
        o.z, o.p_jk, o.like_1, o.like_2 = ptr_i(self.z), ptr_d(self.p_jk), ptr_d(self.like_1), ptr_d(self.like_2)
        o.gamma_i, o.p_k = ptr_d(self.gamma_i), ptr_d(self.p_k)
        o.track_alpha_j, o.track_beta_k = ptr_d(self.track_alpha_j), ptr_d(self.track_beta_k)
        o.track_sigma_k, o.track_p_jk = ptr_d(self.track_sigma_k), ptr_d(self.track_p_jk)

    def as_list(self) -> dict:
        """Same names / shapes as the Rcpp::List of robust_cpp.cpp:562-575 / mle_cpp.cpp:341-356
        (arma column vectors wrap as N x 1 matrices, uvec wraps as double)."""
        res = {
            "D": self.D, "J": self.J, "K": self.K, "N": self.N,
            "mu_0": self.mu_0.reshape(-1, 1),
            "alpha_j": self.alpha_j, "beta_k": self.beta_k, "sigma_k": self.sigma_k,
        }
        if self.variant == SQC_VARIANT_MLE:
            res["gamma_i"] = self.gamma_i
        # 1-based labels as an N x 1 column; kept int32 (Rcpp wraps the arma::uvec as double: as.integer() in R is free,
        # a float64 copy of 10^7 labels here costs more than the D2H transfer)
        res["z"] = self.z.reshape(-1, 1)
        if self.variant == SQC_VARIANT_MLE:
            res["p_k"] = self.p_k.reshape(-1, 1)
        res["p_jk"] = self.p_jk
        res["like_1"] = self.like_1.reshape(-1, 1)
        res["like_2"] = self.like_2.reshape(-1, 1)
        if self.track:
            res["track_alpha_j"] = self.track_alpha_j
            res["track_beta_k"] = self.track_beta_k
            res["track_sigma_k"] = self.track_sigma_k
            res["track_p_jk"] = self.track_p_jk
        # extras (tolerated by every reader of the list)
        res["n_iter"] = int(self.out.n_iter)
        res["rng_draws"] = int(self.out.rng_draws)
        res["mcd_iters_hit"] = int(self.out.mcd_iters_hit)
        res["n_zero_like"] = int(self.out.n_zero_like)
        res["rng_state"] = np.array(self.out.rng_state[:], dtype=np.int32)
        return res


def make_fit_args(x, groups, D, J, K, N, em_iters, *, init_z=None, init_gamma=None, mcd_alpha=0.5,
                  mcd_iters=50, seed=0, track=False, rng_mode=SQC_RNG_R_COMPAT, device=-1,
                  rank=0, world=1, cell_offset=0, N_total=None, peer_handles=None, n_gpus=1):
    """Build a sqc_fit_args plus the list of numpy arrays that must stay alive during the call."""
    keep = []
    xa = f64_colmajor(x)
    if xa.ndim != 2:
        raise ValueError("x must be an N x D matrix")
    ga = as_i32(groups)
    keep += [xa, ga]
    a = SqcFitArgs()
    a.abi_version = SQC_ABI_VERSION
    a.D, a.J, a.K, a.N = int(D), int(J), int(K), int(N)
    a.x, a.groups = ptr_d(xa), ptr_i(ga)
    if init_z is not None:
        za = as_i32(np.asarray(init_z).reshape(-1))
        keep.append(za)
        a.init_z = ptr_i(za)
    if init_gamma is not None:
        gm = f64_colmajor(init_gamma)
        keep.append(gm)
        a.init_gamma = ptr_d(gm)
    a.em_iters, a.mcd_iters, a.mcd_alpha = int(em_iters), int(mcd_iters), float(mcd_alpha)
    a.seed = int(seed) & 0xFFFFFFFF
    a.track, a.rng_mode, a.device = int(bool(track)), int(rng_mode), int(device)
    a.rank, a.world, a.cell_offset = int(rank), int(world), int(cell_offset)
    a.N_total = int(N if N_total is None else N_total)
    a.n_gpus = int(n_gpus)
    if peer_handles is not None:
        ph = np.ascontiguousarray(np.frombuffer(bytes(peer_handles), dtype=np.uint8))
        keep.append(ph)
        a.peer_handles = ph.ctypes.data_as(C.c_void_p)
    return a, keep
