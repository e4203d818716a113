"""Host-side mirror of the reference's Rcpp exports for the EM path, on top of the C ABI.

The names, argument order and return lists are those of the reference's R closures
(reference R/RcppExports.R:7-16 -> src/RcppExports.cpp:16-54):

    fit_sampleqc_robust_cpp(x, init_z, groups, D, J, K, N, em_iters, mcd_alpha, mcd_iters, seed, track=FALSE)
    fit_sampleqc_mle_cpp(x, init_gamma_i, groups, D, J, K, N, n_iter, seed)

plus the post-fit arithmetic of .calc_outliers_dt (reference R/fit_sampleqc.R:419-453) and the
re-centring .fit_one_sampleqc applies to the raw outputs (R/fit_sampleqc.R:242-246).  Errors that the
reference raises as R errors (END_RCPP) surface as SampleQCError.  Everything runs in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A
from ._abi import SampleQCError, SQC_RNG_COUNTER, SQC_RNG_R_COMPAT, SQC_VARIANT_MLE, SQC_VARIANT_ROBUST  # noqa: F401
from ._lib import check, lib


def fit_sampleqc_robust_cpp(x, init_z, groups, D, J, K, N, em_iters, mcd_alpha, mcd_iters, seed, track=False,
                            rng_mode=SQC_RNG_R_COMPAT, device=-1, alloc=None, out=None, n_gpus=1):
    """reference src/fit_sampleQC_robust_cpp.cpp:451-577; returns the Rcpp list as a dict (:562-575).
    `out`: a FitBuffers of matching shape to be reused across calls (e.g. pinned host memory).
    `n_gpus` > 1 (or -1 = all visible): the library shards this one call over the GPUs of the process."""
    args, keep = A.make_fit_args(x, groups, D, J, K, N, em_iters, init_z=init_z, mcd_alpha=mcd_alpha,
                                 mcd_iters=mcd_iters, seed=seed, track=track, rng_mode=rng_mode, device=device, n_gpus=n_gpus)
    bufs = out if out is not None else A.FitBuffers(D, J, K, N, em_iters, SQC_VARIANT_ROBUST, track, alloc=alloc)
    err = C.create_string_buffer(512)
    check(lib().sqc_fit_robust(C.byref(args), C.byref(bufs.out), err, 512), err)
    return bufs.as_list()


def fit_sampleqc_mle_cpp(x, init_gamma_i, groups, D, J, K, N, n_iter, seed, device=-1, alloc=None, n_gpus=1):
    """reference src/fit_sampleQC_mle_cpp.cpp:279-357; returns the Rcpp list as a dict (:341-356)."""
    args, keep = A.make_fit_args(x, groups, D, J, K, N, n_iter, init_gamma=init_gamma_i, seed=seed, device=device, n_gpus=n_gpus)
    bufs = A.FitBuffers(D, J, K, N, n_iter, SQC_VARIANT_MLE, alloc=alloc)
    err = C.create_string_buffer(512)
    check(lib().sqc_fit_mle(C.byref(args), C.byref(bufs.out), err, 512), err)
    return bufs.as_list()


def recentre(fit: dict) -> dict:
    """zero-mean re-centring of mu_0 / alpha_j / beta_k, reference R/fit_sampleqc.R:242-246."""
    out = dict(fit)
    a_means = fit["alpha_j"].mean(axis=0)
    b_means = fit["beta_k"].mean(axis=0)
    out["mu_0"] = fit["mu_0"].reshape(-1) + a_means + b_means
    out["alpha_j"] = fit["alpha_j"] - a_means[None, :]
    out["beta_k"] = fit["beta_k"] - b_means[None, :]
    return out


def calc_outliers(x, groups, mu_0, alpha_j, beta_k, sigma_k, alpha=0.01, device=-1):
    """.calc_outliers_dt (reference R/fit_sampleqc.R:419-453): (maha N x K, outlier N bool, cut)."""
    xa = A.f64_colmajor(x)
    N, D = xa.shape
    aj, bk, sk = A.f64_colmajor(alpha_j), A.f64_colmajor(beta_k), A.f64_colmajor(sigma_k)
    J, K = aj.shape[0], bk.shape[0]
    ga = A.as_i32(groups)
    m0 = np.ascontiguousarray(np.asarray(mu_0, dtype=np.float64).reshape(-1))
    a = A.SqcMahaArgs()
    a.abi_version, a.D, a.J, a.K, a.N = A.SQC_ABI_VERSION, D, J, K, N
    a.x, a.groups, a.mu_0, a.alpha_j, a.beta_k, a.sigma_k = A.ptr_d(xa), A.ptr_i(ga), A.ptr_d(m0), A.ptr_d(aj), A.ptr_d(bk), A.ptr_d(sk)
    a.alpha, a.device = float(alpha), int(device)
    maha = np.zeros((N, K), order="F")
    out = np.zeros(N, dtype=np.uint8)
    cut = C.c_double(0)
    err = C.create_string_buffer(512)
    check(lib().sqc_outlier_maha(C.byref(a), A.ptr_d(maha), out.ctypes.data_as(A.c_uint8_p), C.byref(cut), err, 512), err)
    return maha, out.astype(bool), cut.value


def sample_medians(x, groups, J, device=-1):
    """grp_medians of .init_clusts (reference R/fit_sampleqc.R:286-289): J x D medians, computed on the device."""
    xa = A.f64_colmajor(x)
    N, D = xa.shape
    ga = A.as_i32(groups)
    a = A.SqcInitArgs()
    a.abi_version, a.D, a.J, a.K, a.N = A.SQC_ABI_VERSION, D, J, 0, N
    a.x, a.groups, a.device = A.ptr_d(xa), A.ptr_i(ga), int(device)
    med = np.zeros((J, D), order="F")
    err = C.create_string_buffer(512)
    check(lib().sqc_init_clusts(C.byref(a), A.ptr_d(med), None, None, err, 512), err)
    return med


def gmm_classify(x, groups, J, pro, mean, sigma, want_gamma=True, device=-1):
    """medians + classification of all cells under a mixture fitted on the median-centred subsample
    (summaryMclustBIC(...)$z / $classification, reference R/fit_sampleqc.R:303-316).
    mean: K x D, sigma: D x D x K.  Returns (medians J x D, gamma N x K or None, z 0-based int32)."""
    xa = A.f64_colmajor(x)
    N, D = xa.shape
    ga = A.as_i32(groups)
    pr = np.ascontiguousarray(np.asarray(pro, dtype=np.float64).reshape(-1))
    K = pr.shape[0]
    mu, sg = A.f64_colmajor(np.asarray(mean, dtype=np.float64).reshape(K, D)), A.f64_colmajor(sigma)
    a = A.SqcInitArgs()
    a.abi_version, a.D, a.J, a.K, a.N = A.SQC_ABI_VERSION, D, J, K, N
    a.x, a.groups, a.pro, a.mean, a.sigma, a.device = A.ptr_d(xa), A.ptr_i(ga), A.ptr_d(pr), A.ptr_d(mu), A.ptr_d(sg), int(device)
    med = np.zeros((J, D), order="F")
    gamma = np.zeros((N, K), order="F") if want_gamma else None
    z = np.zeros(N, dtype=np.int32)
    err = C.create_string_buffer(512)
    check(lib().sqc_init_clusts(C.byref(a), A.ptr_d(med), A.ptr_d(gamma) if want_gamma else None, A.ptr_i(z), err, 512), err)
    return med, gamma, z


def init_clusts(x, groups, J, K, N, fit_subsample, seed=0, n_sub=10_000, n_tries=5, device=-1):
    """Mirror of .init_clusts (reference R/fit_sampleqc.R:277-331) with the O(N) work on the device.
    `fit_subsample(x_sample_centred, K, attempt) -> (pro, mean K x D, sigma D x D x K)` is the caller's mixture fit on the
    subsample (mclust "EVE" in the R package, :297-300).  Returns dict(init_gamma N x K, init_z 0-based)."""
    if K == 1:                                                    # :279-283
        return {"init_z": np.zeros(N, dtype=np.int32), "init_gamma": np.zeros((N, 1), order="F")}
    xa = A.f64_colmajor(x)
    ga = A.as_i32(groups)
    rng = np.random.default_rng(seed)
    med = sample_medians(xa, ga, J, device=device)                # :286-289
    for attempt in range(n_tries):                                # :296-309
        idx = rng.choice(N, size=min(N, n_sub), replace=False)
        pro, mean, sigma = fit_subsample(xa[idx] - med[ga[idx]], K, attempt)
        _, gamma, z = gmm_classify(xa, ga, J, pro, mean, sigma, device=device)
        if np.unique(z).size == K:
            return {"init_z": z, "init_gamma": gamma}
    fake = np.exp(rng.standard_normal((N, K)))                    # random fallback, :318-326
    gamma = np.asfortranarray(fake / fake.sum(axis=1, keepdims=True))
    return {"init_z": np.argmax(gamma, axis=1).astype(np.int32), "init_gamma": gamma}


class Session:
    """A resident fit: the cell matrix stays in HBM so the EM loop can be re-run and timed without the
    host<->device copies (sqc_session_* in include/sampleqc_cuda.h)."""

    def __init__(self, x, groups, D, J, K, N, em_iters, *, variant=SQC_VARIANT_ROBUST, init_z=None, init_gamma=None,
                 mcd_alpha=0.5, mcd_iters=50, seed=0, track=False, rng_mode=SQC_RNG_R_COMPAT, device=-1,
                 rank=0, world=1, cell_offset=0, N_total=None, peer_handles=None, raw_pointers=None):
        self.D, self.J, self.K, self.N, self.em_iters = D, J, K, N, em_iters
        self.variant, self.track = variant, track
        args, self._keep = A.make_fit_args(x, groups, D, J, K, N, em_iters, init_z=init_z, init_gamma=init_gamma,
                                           mcd_alpha=mcd_alpha, mcd_iters=mcd_iters, seed=seed, track=track,
                                           rng_mode=rng_mode, device=device, rank=rank, world=world,
                                           cell_offset=cell_offset, N_total=N_total, peer_handles=peer_handles)
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        check(lib().sqc_session_create(C.byref(args), int(variant), C.byref(self._h), err, 512), err)

    def profile(self, on=True):
        lib().sqc_session_profile(self._h, int(on))

    def run(self):
        err = C.create_string_buffer(512)
        check(lib().sqc_session_run(self._h, err, 512), err)

    def fetch(self) -> dict:
        bufs = A.FitBuffers(self.D, self.J, self.K, self.N, self.em_iters, self.variant, self.track)
        err = C.create_string_buffer(512)
        check(lib().sqc_session_fetch(self._h, C.byref(bufs.out), err, 512), err)
        return bufs.as_list()

    def stats(self) -> dict:
        st = A.SqcRunStats()
        lib().sqc_session_stats(self._h, C.byref(st))
        fam = A.KERNEL_FAMILIES
        return {
            "n_iter": st.n_iter, "n_gpu_launches": st.n_gpu_launches, "loop_ms": st.loop_ms, "init_ms": st.init_ms,
            "family_ms": {fam[i]: st.family_ms[i] for i in range(len(fam))},
            "family_launches": {fam[i]: st.family_launches[i] for i in range(len(fam))},
            "family_bytes": {fam[i]: st.family_bytes[i] for i in range(len(fam))},
            "mcd_passes": st.mcd_passes, "mcd_csteps": st.mcd_csteps,
        }

    def close(self):
        if self._h:
            lib().sqc_session_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_qchisq(p, df, device=-1):
    p = np.ascontiguousarray(np.asarray(p, dtype=np.float64).reshape(-1))
    df = np.ascontiguousarray(np.broadcast_to(np.asarray(df, dtype=np.float64), p.shape).copy())
    out = np.zeros_like(p)
    err = C.create_string_buffer(512)
    check(lib().sqc_device_qchisq(A.ptr_d(p), A.ptr_d(df), A.ptr_d(out), p.size, device, err, 512), err)
    return out


def device_rkeys(seed, skip, n, device=-1):
    out = np.zeros(n, dtype=np.int32)
    err = C.create_string_buffer(512)
    check(lib().sqc_device_rkeys(seed, skip, A.ptr_i(out), n, device, err, 512), err)
    return out


def device_rkeys_share(seed, n_total, world, rank, calls, device=-1):
    """rank `rank`'s share of the key stream of `calls` consecutive calls, as a multi-GPU fit generates it
    (returns (calls x Q int32 array, Q, R's state after the last call))"""
    per = -(-n_total // world)
    q_guess = -(-per // 624) * 624
    out = np.zeros((calls, q_guess), dtype=np.int32)
    state = np.zeros(625, dtype=np.int32)
    q = C.c_int64(0)
    err = C.create_string_buffer(512)
    check(lib().sqc_device_rkeys_share(seed, n_total, world, rank, calls, A.ptr_i(out), C.byref(q), A.ptr_i(state), device, err, 512), err)
    assert q.value == q_guess
    return out, int(q.value), state


def device_expneg(x, device=-1):
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
    out = np.zeros_like(x)
    err = C.create_string_buffer(512)
    check(lib().sqc_device_expneg(A.ptr_d(x), A.ptr_d(out), x.size, device, err, 512), err)
    return out


def fp64_peak(device=-1) -> float:
    """measured fp64 FMA instructions per second of the device (x 2 = FLOP/s)"""
    v = C.c_double(0)
    err = C.create_string_buffer(512)
    check(lib().sqc_fp64_peak(int(device), C.byref(v), err, 512), err)
    return v.value


class Resident:
    """The resident pipeline of .fit_one_sampleqc (reference R/fit_sampleqc.R:195-262): .init_clusts -> fit -> re-centring
    -> .calc_outliers_dt on ONE upload of the cells (sqc_resident_* in include/sampleqc_cuda.h)."""

    def __init__(self, x, groups, J, device=-1):
        self._x = A.f64_colmajor(x)
        self.N, self.D = self._x.shape
        self.J = int(J)
        self._g = A.as_i32(groups)
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_create(A.ptr_d(self._x), A.ptr_i(self._g), self.D, self.J, self.N, int(device), C.byref(self._h), err, 512), err)

    def medians(self):
        med = np.zeros((self.J, self.D), order="F")
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_init_clusts(self._h, None, None, None, 0, 0, A.ptr_d(med), None, None, err, 512), err)
        return med

    def subsample(self, idx):
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.int64).reshape(-1))
        out = np.zeros((idx.size, self.D), order="F")
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_subsample(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size, A.ptr_d(out), err, 512), err)
        return out

    def classify(self, pro, mean, sigma, keep_gamma=False, want_gamma=False, want_z=False):
        """classification of all cells under the caller's mixture; the labels stay resident for fit_robust()"""
        pr = np.ascontiguousarray(np.asarray(pro, dtype=np.float64).reshape(-1))
        K = pr.shape[0]
        mu, sg = A.f64_colmajor(np.asarray(mean, dtype=np.float64).reshape(K, self.D)), A.f64_colmajor(sigma)
        gamma = np.zeros((self.N, K), order="F") if want_gamma else None
        z = np.zeros(self.N, dtype=np.int32) if want_z else None
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_init_clusts(self._h, A.ptr_d(pr), A.ptr_d(mu), A.ptr_d(sg), K, int(bool(keep_gamma or want_gamma)), None,
                                             A.ptr_d(gamma) if want_gamma else None, A.ptr_i(z) if want_z else None, err, 512), err)
        return gamma, z

    def fit_robust(self, K, em_iters, mcd_alpha, mcd_iters, seed, track=False, rng_mode=SQC_RNG_R_COMPAT, init_z=None, recentre=True, out=None):
        """fit_sampleqc_robust_cpp on the resident cells (init_z None: the resident classification); parameters re-centred"""
        dummy_x = np.zeros((1, self.D), order="F")
        args, keep = A.make_fit_args(dummy_x, np.zeros(1, dtype=np.int32), self.D, self.J, K, self.N, em_iters, init_z=init_z, mcd_alpha=mcd_alpha,
                                     mcd_iters=mcd_iters, seed=seed, track=track, rng_mode=rng_mode)
        bufs = out if out is not None else A.FitBuffers(self.D, self.J, K, self.N, em_iters, SQC_VARIANT_ROBUST, track)
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_fit(self._h, C.byref(args), SQC_VARIANT_ROBUST, int(bool(recentre)), C.byref(bufs.out), err, 512), err)
        return bufs.as_list()

    def fit_mle(self, K, n_iter, seed, init_gamma=None, recentre=True):
        dummy_x = np.zeros((1, self.D), order="F")
        args, keep = A.make_fit_args(dummy_x, np.zeros(1, dtype=np.int32), self.D, self.J, K, self.N, n_iter, init_gamma=init_gamma, seed=seed)
        bufs = A.FitBuffers(self.D, self.J, K, self.N, n_iter, SQC_VARIANT_MLE)
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_fit(self._h, C.byref(args), SQC_VARIANT_MLE, int(bool(recentre)), C.byref(bufs.out), err, 512), err)
        return bufs.as_list()

    def outliers(self, alpha=0.01, want_maha=True, out_maha=None, out_flags=None):
        maha = out_maha
        if want_maha and maha is None:
            raise ValueError("pass out_maha (N x K, Fortran order) or want_maha=False")
        flags = out_flags if out_flags is not None else np.zeros(self.N, dtype=np.uint8)
        cut = C.c_double(0)
        err = C.create_string_buffer(512)
        check(lib().sqc_resident_outliers(self._h, float(alpha), A.ptr_d(maha) if want_maha else None, flags.ctypes.data_as(A.c_uint8_p), C.byref(cut), err, 512), err)
        return maha, flags.astype(bool) if out_flags is None else flags, cut.value

    def close(self):
        if self._h:
            lib().sqc_resident_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mmd_fn(X, Y, sigma, device=-1) -> float:
    """mmd_fn(X, Y, sigma), reference src/mmd_fn.cpp:20-91 (the Rcpp export of R/RcppExports.R)"""
    xa, ya = A.f64_colmajor(X), A.f64_colmajor(Y)
    if xa.shape[1] != ya.shape[1]:
        raise ValueError("X and Y must have same number of columns")       # :31-33
    out = C.c_double(0)
    err = C.create_string_buffer(512)
    check(lib().sqc_mmd_fn(A.ptr_d(xa), xa.shape[0], A.ptr_d(ya), ya.shape[0], xa.shape[1], float(sigma), C.byref(out), int(device), err, 512), err)
    return out.value


def mmd_subsample_draws(rows, pi, pj, n_times, subsample, seed=22, draw=None):
    """the row picks .dist_fn's sample() would draw (reference R/calc_pairwise_mmds.R:271-277) for every (pair, repetition,
    side): int32 arrays P x n_times x subsample.  `draw(n, k)`: the caller's sampler, called pair by pair, repetition by
    repetition, side i then side j — the order of the R loop; default: numpy, vectorised per sample."""
    P = len(pi)
    idx_i = np.zeros((P, n_times, subsample), dtype=np.int32); idx_j = np.zeros_like(idx_i)
    if draw is not None:
        for p in range(P):
            for t in range(n_times):
                if rows[pi[p]] > subsample:
                    idx_i[p, t] = draw(int(rows[pi[p]]), subsample)
                if rows[pj[p]] > subsample:
                    idx_j[p, t] = draw(int(rows[pj[p]]), subsample)
        return idx_i, idx_j
    rng = np.random.default_rng(seed)
    for side, out in ((pi, idx_i), (pj, idx_j)):
        for s in np.unique(side):
            n = int(rows[s])
            if n <= subsample:
                continue
            where = np.nonzero(side == s)[0]
            for c0 in range(0, where.size, 64):                   # chunks bound the scratch matrix of random keys
                w = where[c0:c0 + 64]
                keys = rng.random((w.size * n_times, n))
                out[w] = np.argpartition(keys, subsample - 1, axis=1)[:, :subsample].reshape(w.size, n_times, subsample)
    return idx_i, idx_j


def calc_mmd_mat(mat_list, n_times, subsample, sigma, draw=None, seed=22, device=-1, return_all=False, timing=None):
    """.calc_mmd_mat (reference R/calc_pairwise_mmds.R:198-247): the symmetric matrix of mean |MMD| over n_times subsample
    pairs for every pair of samples, all pairs in ONE device call (`timing`: a dict that receives the seconds of that call)."""
    import time
    S = len(mat_list)
    mats = [A.f64_colmajor(m) for m in mat_list]
    D = mats[0].shape[1]
    rows = np.array([m.shape[0] for m in mats], dtype=np.int64)
    flat = np.concatenate([m.ravel(order="F") for m in mats])
    pi, pj = np.array([(i, j) for j in range(S) for i in range(j)], dtype=np.int32).reshape(-1, 2).T.copy()      # sample_i < sample_j, :205-206
    P = pi.size
    idx_i, idx_j = mmd_subsample_draws(rows, pi, pj, n_times, subsample, seed=seed, draw=draw)
    a = A.SqcMmdArgs()
    a.abi_version, a.D, a.n_samples = A.SQC_ABI_VERSION, D, S
    a.mats, a.mat_rows = A.ptr_d(flat), rows.ctypes.data_as(C.POINTER(C.c_int64))
    a.n_pairs, a.pair_i, a.pair_j = P, A.ptr_i(pi), A.ptr_i(pj)
    a.n_times, a.subsample, a.idx_i, a.idx_j = int(n_times), int(subsample), A.ptr_i(idx_i), A.ptr_i(idx_j)
    a.sigma, a.device = float(sigma), int(device)
    mean = np.zeros(P); allv = np.zeros((P, n_times))
    err = C.create_string_buffer(512)
    t0 = time.perf_counter()
    check(lib().sqc_mmd_pairs(C.byref(a), A.ptr_d(mean), A.ptr_d(allv), err, 512), err)
    if timing is not None:
        timing["device_call_s"] = time.perf_counter() - t0
    mat = np.zeros((S, S))
    mat[pi, pj] = mean; mat[pj, pi] = mean                        # the reverse versions, :236-239; diagonal 0 (fill = 0, :244)
    return (mat, allv, (pi, pj, idx_i, idx_j)) if return_all else mat
