"""ctypes front-end of the CPU oracle (oracle/libsqc_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (sampleqc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from sampleqc_b200 import _abi as A  # noqa: E402  (ABI description only)

_LIB = None
_REF = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsqc_oracle.so")
    src = os.path.join(_HERE, "sqc_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libsqc_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        for name in ("sqc_oracle_fit_robust", "sqc_oracle_fit_mle"):
            f = getattr(L, name)
            f.argtypes = [C.POINTER(A.SqcFitArgs), C.POINTER(A.SqcFitOut), C.c_char_p, C.c_size_t]
            f.restype = C.c_int
        L.sqc_oracle_outlier_maha.argtypes = [C.POINTER(A.SqcMahaArgs), A.c_double_p, A.c_uint8_p, A.c_double_p, C.c_char_p, C.c_size_t]
        L.sqc_oracle_outlier_maha.restype = C.c_int
        L.sqc_oracle_runif.argtypes = [C.c_uint32, A.c_double_p, C.c_int64]
        L.sqc_oracle_rkeys.argtypes = [C.c_uint32, C.c_int64, A.c_int32_p, C.c_int64]
        L.sqc_oracle_counter_key.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.sqc_oracle_counter_key.restype = C.c_int32
        L.sqc_oracle_qchisq.argtypes = [C.c_double, C.c_double]
        L.sqc_oracle_qchisq.restype = C.c_double
        L.sqc_oracle_set_cost_faithful.argtypes = [C.c_int]
        L.sqc_oracle_mmd_fn.argtypes = [A.c_double_p, C.c_int64, A.c_double_p, C.c_int64, C.c_int32, C.c_double, A.c_double_p]
        L.sqc_oracle_mmd_fn.restype = C.c_int
        L.sqc_oracle_fast_mcd.argtypes = [A.c_double_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_uint32,
                                          A.c_double_p, A.c_double_p, A.c_int32_p, A.c_double_p, C.c_char_p, C.c_size_t]
        L.sqc_oracle_fast_mcd.restype = C.c_int
        _LIB = L
    return _LIB


def _run(fn, args, bufs):
    err = C.create_string_buffer(512)
    st = fn(C.byref(args), C.byref(bufs.out), err, 512)
    if st != 0:
        raise A.SampleQCError(st, err.value.decode(errors="replace"))
    return bufs.as_list()


def fit_sampleqc_robust_cpp(x, init_z, groups, D, J, K, N, em_iters, mcd_alpha, mcd_iters, seed, track=False,
                            rng_mode=A.SQC_RNG_R_COMPAT, fn=None):
    args, keep = A.make_fit_args(x, groups, D, J, K, N, em_iters, init_z=init_z, mcd_alpha=mcd_alpha,
                                 mcd_iters=mcd_iters, seed=seed, track=track, rng_mode=rng_mode)
    bufs = A.FitBuffers(D, J, K, N, em_iters, A.SQC_VARIANT_ROBUST, track)
    return _run(fn or lib().sqc_oracle_fit_robust, args, bufs)


def fit_sampleqc_mle_cpp(x, init_gamma_i, groups, D, J, K, N, n_iter, seed, fn=None):
    args, keep = A.make_fit_args(x, groups, D, J, K, N, n_iter, init_gamma=init_gamma_i, seed=seed)
    bufs = A.FitBuffers(D, J, K, N, n_iter, A.SQC_VARIANT_MLE)
    return _run(fn or lib().sqc_oracle_fit_mle, args, bufs)


def outlier_maha(x, groups, mu_0, alpha_j, beta_k, sigma_k, alpha, fn=None):
    xa = A.f64_colmajor(x)
    N, D = xa.shape
    aj, bk, sk = A.f64_colmajor(alpha_j), A.f64_colmajor(beta_k), A.f64_colmajor(sigma_k)
    J, K = aj.shape[0], bk.shape[0]
    ga = A.as_i32(groups)
    m0 = np.ascontiguousarray(np.asarray(mu_0, dtype=np.float64).reshape(-1))
    a = A.SqcMahaArgs()
    a.abi_version, a.D, a.J, a.K, a.N = A.SQC_ABI_VERSION, D, J, K, N
    a.x, a.groups, a.mu_0, a.alpha_j, a.beta_k, a.sigma_k = A.ptr_d(xa), A.ptr_i(ga), A.ptr_d(m0), A.ptr_d(aj), A.ptr_d(bk), A.ptr_d(sk)
    a.alpha, a.device = float(alpha), -1
    maha = np.zeros((N, K), order="F")
    out = np.zeros(N, dtype=np.uint8)
    cut = C.c_double(0)
    err = C.create_string_buffer(512)
    st = (fn or lib().sqc_oracle_outlier_maha)(C.byref(a), A.ptr_d(maha), out.ctypes.data_as(A.c_uint8_p), C.byref(cut), err, 512)
    if st != 0:
        raise A.SampleQCError(st, err.value.decode(errors="replace"))
    return maha, out.astype(bool), cut.value


def mmd_fn(X, Y, sigma) -> float:
    """reference src/mmd_fn.cpp:20-91"""
    xa, ya = A.f64_colmajor(X), A.f64_colmajor(Y)
    if xa.shape[1] != ya.shape[1]:
        raise ValueError("X and Y must have same number of columns")       # :31-33
    out = C.c_double(0)
    if lib().sqc_oracle_mmd_fn(A.ptr_d(xa), xa.shape[0], A.ptr_d(ya), ya.shape[0], xa.shape[1], float(sigma), C.byref(out)) != 0:
        raise A.SampleQCError(1, "sigma must be > 0")
    return out.value


def runif(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n)
    lib().sqc_oracle_runif(seed, A.ptr_d(out), n)
    return out


def rkeys(seed: int, skip: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.int32)
    lib().sqc_oracle_rkeys(seed, skip, A.ptr_i(out), n)
    return out


def counter_key(seed: int, call: int, cell: int) -> int:
    return int(lib().sqc_oracle_counter_key(seed, call, cell))


def qchisq(p: float, df: float) -> float:
    return float(lib().sqc_oracle_qchisq(p, df))


def set_cost_faithful(on: bool) -> None:
    lib().sqc_oracle_set_cost_faithful(int(on))


def fast_mcd(xk, h, mcd_iters, seed):
    xk = np.ascontiguousarray(np.asarray(xk, dtype=np.float64))
    n, D = xk.shape
    loc = np.zeros(D)
    scale = np.zeros((D, D), order="F")
    iters = C.c_int32(0)
    q = C.c_double(0)
    err = C.create_string_buffer(512)
    st = lib().sqc_oracle_fast_mcd(A.ptr_d(xk), n, D, h, mcd_iters, seed, A.ptr_d(loc), A.ptr_d(scale),
                                   C.byref(iters), C.byref(q), err, 512)
    if st != 0:
        raise A.SampleQCError(st, err.value.decode(errors="replace"))
    return loc, scale, iters.value, q.value


# ---- .init_clusts pieces (reference R/fit_sampleqc.R:277-331), plain numpy --------------------------------------
def sample_medians(x, groups, J):
    """grp_medians (:286-289): apply(x[groups == j, ], 2, median) for every sample; J x D"""
    x = np.asarray(x, dtype=np.float64)
    g = np.asarray(groups)
    med = np.full((J, x.shape[1]), np.nan)
    for j in range(J):
        rows = x[g == j]
        if rows.shape[0]:
            med[j] = np.median(rows, axis=0)
    return med


def gmm_classify(x, groups, medians, pro, mean, sigma):
    """posterior (mclust z) and its first maximum (classification - 1) of the median-centred cells (:290, :303-316);
    mean K x D, sigma D x D x K"""
    x = np.asarray(x, dtype=np.float64)
    xc = x - np.asarray(medians)[np.asarray(groups)]
    K, D = len(pro), x.shape[1]
    mean = np.asarray(mean, dtype=np.float64).reshape(K, D)
    sigma = np.asarray(sigma, dtype=np.float64).reshape(D, D, K, order="F")
    ll = np.empty((x.shape[0], K))
    for k in range(K):
        L = np.linalg.cholesky(sigma[:, :, k])
        w = np.linalg.solve(L, (xc - mean[k]).T)
        ll[:, k] = np.log(pro[k]) - np.log(np.diag(L)).sum() - 0.5 * D * np.log(2.0 * np.pi) - 0.5 * (w * w).sum(axis=0)
    m = ll.max(axis=1, keepdims=True)
    e = np.exp(ll - m)
    return e / e.sum(axis=1, keepdims=True), np.argmax(ll, axis=1).astype(np.int32)
