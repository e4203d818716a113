"""Second, INDEPENDENT restatement of the reference's EM path in plain numpy.  TEST INFRASTRUCTURE ONLY.

Written from the reference sources (cited per function), NOT from oracle/sqc_oracle.cpp: different language, different
linear algebra (LAPACK through numpy instead of hand-written Gauss-Jordan / Cholesky), different Mersenne-Twister
implementation (numpy's MT19937 bit generator, initialised with R's set.seed scrambling), a real sort instead of
nth_element, vectorised pairwise sums instead of serial loops.  tests/test_np_restatement.py requires the C++ oracle
and this module to agree to 1e-9 on every output (same iteration count, same labels): a transcription slip in either
one shows as a disagreement.  The reference itself cannot be built here (R / Rcpp / Armadillo / RcppDist absent), so
this cross-check is what narrows "parity unpinned" (DESIGN.md section 2).

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy.stats import chi2

RAND_MAX = 2147483647.0


class RUnif:
    """R's default RNG after set.seed(seed), reference src/helpers.cpp:10-14:
    initial scrambling of R's RNG.c, then MT19937; unif_rand = word * 2.3283064365386963e-10, kept inside (0, 1)."""

    def __init__(self, seed: int):
        s = np.uint32(seed & 0xFFFFFFFF)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            i_seed = np.empty(625, dtype=np.uint32)
            for j in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                i_seed[j] = s
        self.bg = np.random.MT19937()
        st = self.bg.state
        st["state"]["key"] = i_seed[1:].copy()
        st["state"]["pos"] = 624                     # dummy[0] = 624: regenerate on the first draw
        self.bg.state = st
        self.draws = 0

    def unif(self, n: int) -> np.ndarray:
        w = self.bg.random_raw(n).astype(np.float64)
        u = w * 2.3283064365386963e-10
        lo = 0.5 * 2.328306437080797e-10             # fixup(): 1/(2^32 - 1) / 2
        u = np.where(u <= 0.0, lo, u)
        u = np.where(1.0 - u <= 0.0, 1.0 - lo, u)
        self.draws += n
        return u

    def randperm_keys(self, n: int) -> np.ndarray:
        """arma::randperm under RcppArmadillo: one (int) Rf_runif(0, RAND_MAX) per element"""
        return np.floor(0.0 + (RAND_MAX - 0.0) * self.unif(n)).astype(np.int64)


class CounterKeys:
    """the repo's counter-based alternative (rng_mode = 1, include/sampleqc_cuda.h): key = top 31 bits of
    splitmix64(splitmix64(seed << 32) ^ draw * golden), draw = running index in the same stream R's RNG would serve"""

    def __init__(self, seed: int, n_total: int):
        self.mix = self._mix(np.uint64((seed & 0xFFFFFFFF) << 32))
        self.draws = 0

    @staticmethod
    def _mix(z):
        with np.errstate(over="ignore"):
            z = z + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def randperm_keys(self, n: int) -> np.ndarray:
        draw = np.arange(self.draws, self.draws + n, dtype=np.uint64)
        self.draws += n
        with np.errstate(over="ignore"):
            v = self._mix(self.mix ^ (draw * np.uint64(0x9E3779B97F4A7C15)))
        return (v >> np.uint64(33)).astype(np.int64)


def calc_maha_dists(x, loc, cov):
    """robust_cpp.cpp:11-24: solve(trimatl(chol(cov).t()), (x - loc).t()), squared column norms"""
    L = np.linalg.cholesky(cov)                       # lower factor, cov = L L'
    from scipy.linalg import solve_triangular
    y = solve_triangular(L, (x - loc[None, :]).T, lower=True)
    return (y * y).sum(axis=0)


def fast_mcd(x, h, mcd_iters, keys, trace=None):
    """robust_cpp.cpp:55-108"""
    n, D = x.shape
    if h > n or h < 1:
        raise ValueError("randperm(): h out of range")
    h_idx = np.argsort(keys, kind="stable")[:h]       # :72, ties towards the lower index (repo definition)

    def loc_scale(idx):
        sub = x[idx]
        loc = (sub / h).sum(axis=0)                    # :25-35 (division inside the sum)
        c = sub - loc[None, :]
        return loc, (c[:, :, None] * c[:, None, :] / h).sum(axis=0)   # :36-48, biased 1/h

    loc, scale = loc_scale(h_idx)
    det_new = np.linalg.det(scale)
    det_old = 2.0 * det_new                             # :80
    iters = 0
    while (det_old - det_new > 1e-10) and (det_new > 0) and (iters < mcd_iters):   # :84
        det_old = det_new
        d = calc_maha_dists(x, loc, scale)
        if np.isnan(d).any():
            raise FloatingPointError("sort_index(): detected NaN")
        order = np.argsort(d, kind="stable")
        h_idx = order[:h]
        if trace is not None:
            trace.append(float(d[order[h - 1]]))
        loc, scale = loc_scale(h_idx)
        det_new = np.linalg.det(scale)
        iters += 1
    d = calc_maha_dists(x, loc, scale)                  # :98-103
    q_dists = np.sort(d, kind="stable")[h - 1]
    if trace is not None:
        trace.append(float(q_dists))
    q_chisq = chi2.ppf(h / n, D)
    return loc, scale * q_dists / q_chisq, iters


def dmvnorm(x, mu, S, log):
    """RcppDist::dmvnorm for all rows of x at once; mu is per row (N x D)"""
    D = x.shape[1]
    Si = np.linalg.inv(S)
    v = x - mu
    q = np.einsum("ia,ab,ib->i", v, Si, v)
    ld = -0.5 * D * np.log(2.0 * np.pi) - 0.5 * np.log(np.linalg.det(S)) - 0.5 * q
    return ld if log else np.exp(ld)


def update_p_jk(z, groups, J, K):
    """:322-341"""
    cnt = np.ones((J, K))
    np.add.at(cnt, (groups, z), 1.0)
    return cnt / (K + np.bincount(groups, minlength=J))[:, None]


def update_alpha_j(x, groups, beta_k, scale_k, z, J, K):
    """:116-157"""
    D = x.shape[1]
    inv = np.stack([np.linalg.inv(scale_k[:, :, k]) for k in range(K)])
    numer = np.zeros((J, D)); denom = np.zeros((J, D, D))
    contrib = np.einsum("iab,ib->ia", inv[z], x - beta_k[z])
    np.add.at(numer, groups, contrib)
    np.add.at(denom, groups, inv[z])
    return np.stack([np.linalg.inv(denom[j]) @ numer[j] for j in range(J)])


def update_beta_scale_k(x, groups, alpha_j, z, K, mcd_iters, mcd_alpha, rng, traces=None):
    """:289-318 with restrict_to_x_k :265-283"""
    D = x.shape[1]
    beta = np.zeros((K, D)); scale = np.zeros((D, D, K)); hit = 0
    for k in range(K):
        sel = np.nonzero(z == k)[0]
        if sel.size == 0:
            raise IndexError("x_k.rows(0, -1): empty cluster")
        x_k = x[sel] - alpha_j[groups[sel]]
        h_k = int((sel.size + D + 1) * mcd_alpha)       # :306, truncation
        tr = [] if traces is not None else None
        loc, sc, iters = fast_mcd(x_k, h_k, mcd_iters, rng.randperm_keys(sel.size), tr)
        if traces is not None:
            traces.append((k, sel.size, tr))
        hit += iters == mcd_iters
        beta[k] = loc; scale[:, :, k] = sc
    return beta, scale, hit


def calc_log_likelihood(x, groups, alpha_j, beta_k, sigma_k, p, K):
    """:400-432 (p is J x K) / mle_cpp.cpp:191-222 (p is K): plain densities, no log-sum-exp"""
    like = np.zeros(x.shape[0])
    for k in range(K):
        pk = p[groups, k] if p.ndim == 2 else p[k]
        like += pk * dmvnorm(x, alpha_j[groups] + beta_k[k][None, :], sigma_k[:, :, k], False)
    with np.errstate(divide="ignore"):
        return np.log(like).sum()


def calc_expected_z(x, groups, alpha_j, beta_k, scale_k, p_jk, K):
    """:344-396: first strictly greatest log-density; all -inf -> 0"""
    N = x.shape[0]
    with np.errstate(divide="ignore"):
        ll = np.stack([np.log(p_jk[groups, k]) + dmvnorm(x, alpha_j[groups] + beta_k[k][None, :], scale_k[:, :, k], True) for k in range(K)], axis=1)
    z = np.zeros(N, dtype=np.int64); best = np.full(N, -np.inf)
    for k in range(K):
        better = ll[:, k] > best
        z[better] = k; best[better] = ll[better, k]
    return z


def fit_sampleqc_robust(x, init_z, groups, D, J, K, N, em_iters, mcd_alpha, mcd_iters, seed, rng_mode=0, traces=None):
    """robust_cpp.cpp:451-577; returns the list's numeric members (z 1-based)"""
    if mcd_alpha < 0 or mcd_alpha >= 1:
        raise ValueError("'mcd_alpha' must be between 0 and 1")
    x = np.asarray(x, dtype=np.float64); groups = np.asarray(groups, dtype=np.int64)
    rng = RUnif(seed) if rng_mode == 0 else CounterKeys(seed, N)
    mu_0 = x.mean(axis=0)                                # :478
    old_z = np.asarray(init_z, dtype=np.int64).copy()
    x = x - mu_0[None, :]                                # centre_x, helpers.cpp:105-115
    p_jk = update_p_jk(old_z, groups, J, K)
    alpha_j = np.stack([x[groups == j].sum(axis=0) / max((groups == j).sum(), 1) for j in range(J)])   # helpers.cpp:117-140
    beta_k, scale_k, hit = update_beta_scale_k(x, groups, alpha_j, old_z, K, mcd_iters, mcd_alpha, rng, traces)
    like_1 = np.zeros(max(em_iters, 0)); like_2 = np.zeros(max(em_iters, 0))
    z_delta, it = 1, 0
    new_z = np.zeros(N, dtype=np.int64)
    while z_delta > 0 and it < em_iters:                 # :490
        alpha_j = update_alpha_j(x, groups, beta_k, scale_k, old_z, J, K)
        like_1[it] = calc_log_likelihood(x, groups, alpha_j, beta_k, scale_k, p_jk, K)
        beta_k, scale_k, h2 = update_beta_scale_k(x, groups, alpha_j, old_z, K, mcd_iters, mcd_alpha, rng, traces)
        hit += h2
        p_jk = update_p_jk(old_z, groups, J, K)          # with old_z, :504
        like_2[it] = calc_log_likelihood(x, groups, alpha_j, beta_k, scale_k, p_jk, K)
        new_z = calc_expected_z(x, groups, alpha_j, beta_k, scale_k, p_jk, K)
        z_delta = int((new_z != old_z).sum())
        old_z = new_z
        it += 1
    order = np.argsort(beta_k[:, 0], kind="stable")      # :527-531
    inv = np.empty(K, dtype=np.int64); inv[order] = np.arange(K)
    return {"mu_0": mu_0, "alpha_j": alpha_j, "beta_k": beta_k[order], "sigma_k": scale_k[:, :, order], "z": inv[new_z] + 1,
            "p_jk": p_jk[:, order], "like_1": like_1, "like_2": like_2, "n_iter": it, "mcd_iters_hit": int(hit)}


def fit_sampleqc_mle(x, init_gamma, groups, D, J, K, N, n_iter):
    """mle_cpp.cpp:279-357"""
    x = np.asarray(x, dtype=np.float64); groups = np.asarray(groups, dtype=np.int64)
    g0 = np.asarray(init_gamma, dtype=np.float64)
    mu_0 = x.mean(axis=0)
    x = x - mu_0[None, :]
    alpha_j = np.stack([x[groups == j].sum(axis=0) / max((groups == j).sum(), 1) for j in range(J)])

    def max_beta(alpha_j, g):                             # :65-91
        r = x - alpha_j[groups]
        return (g.T @ r) / g.sum(axis=0)[:, None]

    def max_sigma(alpha_j, beta_k, g):                    # :92-129
        S = np.zeros((D, D, K))
        for k in range(K):
            t = x - alpha_j[groups] - beta_k[k][None, :]
            S[:, :, k] = np.einsum("i,ia,ib->ab", g[:, k], t, t) / g[:, k].sum()
        return S

    def max_alpha(beta_k, sigma_k, g):                    # :26-64
        inv = np.stack([np.linalg.inv(sigma_k[:, :, k]) for k in range(K)])
        numer = np.zeros((J, D)); denom = np.zeros((J, D, D))
        for k in range(K):
            np.add.at(numer, groups, g[:, k][:, None] * ((x - beta_k[k][None, :]) @ inv[k].T))
            np.add.at(denom, groups, g[:, k][:, None, None] * inv[k][None, :, :])
        return np.stack([np.linalg.inv(denom[j]) @ numer[j] for j in range(J)])

    def e_gamma(alpha_j, beta_k, sigma_k, p_k):           # :151-187
        g = np.stack([p_k[k] * dmvnorm(x, alpha_j[groups] + beta_k[k][None, :], sigma_k[:, :, k], False) for k in range(K)], axis=1)
        with np.errstate(invalid="ignore", divide="ignore"):
            return g / g.sum(axis=1, keepdims=True)

    beta_k = max_beta(alpha_j, g0)
    sigma_k = max_sigma(alpha_j, beta_k, g0)
    p_k = g0.sum(axis=0) / g0.sum()                       # :130-147
    like_1 = np.zeros(n_iter); like_2 = np.zeros(n_iter)
    gamma = g0
    for i in range(n_iter):
        gamma = e_gamma(alpha_j, beta_k, sigma_k, p_k)
        alpha_j = max_alpha(beta_k, sigma_k, gamma)
        like_1[i] = calc_log_likelihood(x, groups, alpha_j, beta_k, sigma_k, p_k, K)
        gamma = e_gamma(alpha_j, beta_k, sigma_k, p_k)
        beta_k = max_beta(alpha_j, gamma)
        sigma_k = max_sigma(alpha_j, beta_k, gamma)
        p_k = gamma.sum(axis=0) / gamma.sum()
        like_2[i] = calc_log_likelihood(x, groups, alpha_j, beta_k, sigma_k, p_k, K)
    p_jk = np.zeros((J, K)); np.add.at(p_jk, groups, gamma)   # :248-273
    p_jk = p_jk / p_jk.sum(axis=1, keepdims=True)
    order = np.argsort(beta_k[:, 0], kind="stable")
    # extract_z :224-246: z(i) = idx(k_max), k_max the first strictly greatest gamma
    kmax = np.zeros(N, dtype=np.int64); best = np.full(N, -1.0)
    for k in range(K):
        better = gamma[:, k] > best
        kmax[better] = k; best[better] = gamma[better, k]
    return {"mu_0": mu_0, "alpha_j": alpha_j, "beta_k": beta_k[order], "sigma_k": sigma_k[:, :, order], "gamma_i": gamma,
            "z": order[kmax] + 1, "p_k": p_k[order], "p_jk": p_jk[:, order], "like_1": like_1, "like_2": like_2}
