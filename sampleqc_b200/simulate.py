"""Synthetic QC-metric generator and cluster initialisation for tests and benchmarks.

R is absent from this image, so the benchmark inputs restate the *distributional recipe* of the
reference simulator (reference R/simulate_qcs.R) in numpy — distribution-equivalent, not
stream-equivalent:

  sample sizes   n_j ~ multinomial(N, exp(N(0,1)/2))                         simulate_qcs.R:543-546
  mu_0           (log10 3000, log10 1500, logit .03) + N(0, S(.3,.4,1.5; .98,-.7,-.8))   :191-223
  alpha_j        N(0, S(.25,.25,.6; .9,-.4,-.6)), column-centred              :586-601
  beta_k         N(0, S(.4,.3,1.5; .95,-.6,-.6)), centred, sorted by column 1 :247-270
  Sigma_k        Wishart(df=10, V/10), V = S(sqrt(2.27e-2,1.66e-2,2.93e-1); .99,-.6,-.7)  :299-319
  p_j.           Dirichlet(5,...,5)                                           :672-681
  z_i            Categorical(p_j.)                                            :696-707
  x_i            N(mu_0 + alpha_j + beta_z, Sigma_z)                          :794-797
  outliers       p_out,j ~ Beta(theta p0, theta (1-p0)), p0 = logistic(logit .08 + .3 e),
                 theta = exp(4 + .5 e), p_loss,j = logistic(.5 e + logit p_loss0);
                 outlier cells lose reads binomially and are re-logged        :399-450, :721-746, :823-852

For D = 6 (the CITE-seq style config, which simulate_qcs cannot produce: D is hard-coded to 3 there)
metrics 4-6 are a second, independent copy of the 3-metric recipe.

``init_clusts`` mirrors .init_clusts (reference R/fit_sampleqc.R:277-331) with scikit-learn's
GaussianMixture standing in for mclust "EVE" (the initialisation is an INPUT of the hot path).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def _cov(sd, c12, c13, c23):
    corr = np.array([[1.0, c12, c13], [c12, 1.0, c23], [c13, c23, 1.0]])
    sd = np.asarray(sd, dtype=np.float64)
    return (sd[:, None] * corr) * sd[None, :]


def _logistic(v):
    return 1.0 / (1.0 + np.exp(-v))


def _logit(p):
    return np.log(p / (1.0 - p))


@dataclass
class SimData:
    x: np.ndarray          # N x D, column-major (Fortran order), un-centred
    groups: np.ndarray     # N int32, 0-based sample index, non-decreasing
    z_true: np.ndarray     # N int32
    outlier_true: np.ndarray  # N bool
    mu_0: np.ndarray
    alpha_j: np.ndarray
    beta_k: np.ndarray
    sigma_k: np.ndarray    # K x D x D
    p_jk: np.ndarray
    n_j: np.ndarray

    @property
    def N(self): return self.x.shape[0]

    @property
    def D(self): return self.x.shape[1]

    @property
    def J(self): return int(self.n_j.shape[0])


def _block3(rng, N_j, J, K, groups, z, outliers, p_loss):
    """one 3-metric block of the recipe for all cells"""
    N = groups.shape[0]
    mu_0 = np.array([np.log10(3000.0), np.log10(1500.0), _logit(0.03)]) + \
        rng.multivariate_normal(np.zeros(3), _cov([0.3, 0.4, 1.5], 0.98, -0.7, -0.8))
    alpha = rng.multivariate_normal(np.zeros(3), _cov([0.25, 0.25, 0.6], 0.9, -0.4, -0.6), size=J)
    alpha -= alpha.mean(axis=0, keepdims=True)
    beta = rng.multivariate_normal(np.zeros(3), _cov([0.4, 0.3, 1.5], 0.95, -0.6, -0.6), size=K)
    beta -= beta.mean(axis=0, keepdims=True)
    beta = beta[np.argsort(beta[:, 0], kind="stable")]
    V = _cov(np.sqrt([2.27e-2, 1.66e-2, 2.93e-1]), 0.99, -0.6, -0.7)
    Lv = np.linalg.cholesky(V / 10.0)
    sigma = np.empty((K, 3, 3))
    for k in range(K):                      # Wishart(df=10, V/10) = sum of 10 outer products
        g = Lv @ rng.standard_normal((3, 10))
        sigma[k] = g @ g.T
    x = np.empty((N, 3), order="F")
    base = mu_0[None, :] + alpha[groups]
    for k in range(K):
        idx = np.nonzero(z == k)[0]
        Lk = np.linalg.cholesky(sigma[k])
        x[idx] = base[idx] + beta[k][None, :] + rng.standard_normal((idx.shape[0], 3)) @ Lk.T
    # outliers: loss of non-mito reads and features, re-logged (.sim_outliers :823-852)
    oi = np.nonzero(outliers)[0]
    if oi.size:
        xt = x[oi]
        total_old = np.round(10.0 ** xt[:, 0])
        feats_old = np.round(10.0 ** xt[:, 1])
        mt_prop = _logistic(xt[:, 2])
        non_mt_old = np.round(total_old * (1.0 - mt_prop))
        mt_old = total_old - non_mt_old
        keep = 1.0 - p_loss[groups[oi]]
        non_mt_new = rng.binomial(np.maximum(non_mt_old, 0).astype(np.int64), keep) + 1
        feats_new = rng.binomial(np.maximum(feats_old, 0).astype(np.int64), keep) + 1
        total_new = non_mt_new + mt_old + 1
        x[oi, 0] = np.log10(total_new)
        x[oi, 1] = np.log10(feats_new)
        x[oi, 2] = _logit((mt_old + 1) / total_new)
    return x, mu_0, alpha, beta, sigma


def simulate_qcs(N: int, J: int, K: int, D: int = 3, seed: int = 0, p_out_00: float = 0.08,
                 p_loss_0: float = 0.5, dirichlet: float = 5.0) -> SimData:
    """One sample group of simulate_qcs-style data: N cells over J samples, K true components."""
    if D not in (3, 6):
        raise ValueError("the simulate_qcs recipe is defined for D = 3 (and the D = 6 two-block extension)")
    rng = np.random.default_rng(seed)
    w = np.exp(rng.standard_normal(J) / 2.0)
    n_j = rng.multinomial(N, w / w.sum())
    # a sample with < 2 cells is useless for any fit; move cells from the biggest samples
    while n_j.min() < 2:
        n_j[np.argmin(n_j)] += 2
        n_j[np.argmax(n_j)] -= 2
    groups = np.repeat(np.arange(J, dtype=np.int32), n_j)
    p_jk = rng.dirichlet(np.full(K, dirichlet), size=J)
    z = np.empty(N, dtype=np.int32)
    start = 0
    for j in range(J):
        z[start:start + n_j[j]] = rng.choice(K, size=n_j[j], p=p_jk[j])
        start += n_j[j]
    p0 = _logistic(_logit(p_out_00) + 0.3 * rng.standard_normal())
    theta = np.exp(4.0 + 0.5 * rng.standard_normal())
    p_out = rng.beta(theta * p0, theta * (1.0 - p0), size=J)
    p_loss0 = _logistic(_logit(p_loss_0) + 0.5 * rng.standard_normal())
    p_loss = _logistic(0.5 * rng.standard_normal(J) + _logit(p_loss0))
    outliers = rng.random(N) < p_out[groups]
    blocks = [_block3(rng, n_j, J, K, groups, z, outliers, p_loss) for _ in range(D // 3)]
    x = np.asfortranarray(np.concatenate([b[0] for b in blocks], axis=1))
    mu_0 = np.concatenate([b[1] for b in blocks])
    alpha = np.concatenate([b[2] for b in blocks], axis=1)
    beta = np.concatenate([b[3] for b in blocks], axis=1)
    sigma = np.zeros((K, D, D))
    for bi, b in enumerate(blocks):
        sigma[:, 3 * bi:3 * bi + 3, 3 * bi:3 * bi + 3] = b[4]
    return SimData(x=x, groups=groups, z_true=z, outlier_true=outliers, mu_0=mu_0, alpha_j=alpha,
                   beta_k=beta, sigma_k=sigma, p_jk=p_jk, n_j=n_j)


def sample_medians(x: np.ndarray, groups: np.ndarray, J: int) -> np.ndarray:
    """grp_medians of .init_clusts (R/fit_sampleqc.R:286-289); groups must be non-decreasing"""
    bounds = np.searchsorted(groups, np.arange(J + 1))
    med = np.zeros((J, x.shape[1]))
    for j in range(J):
        if bounds[j + 1] > bounds[j]:
            med[j] = np.median(x[bounds[j]:bounds[j + 1]], axis=0)
    return med


def init_clusts(x: np.ndarray, groups: np.ndarray, J: int, K: int, seed: int = 0, n_sub: int = 10_000,
                n_tries: int = 5, chunk: int = 2_000_000):
    """Mirror of .init_clusts (R/fit_sampleqc.R:277-331): returns (init_z 0-based int32, init_gamma N x K)."""
    N = x.shape[0]
    if K == 1:
        return np.zeros(N, dtype=np.int32), np.zeros((N, 1), order="F")
    from sklearn.mixture import GaussianMixture

    rng = np.random.default_rng(seed)
    xc = x - sample_medians(x, groups, J)[groups]
    init_z = init_gamma = None
    for t in range(n_tries):
        idx = rng.choice(N, size=min(N, n_sub), replace=False)
        gm = GaussianMixture(K, covariance_type="full", random_state=seed + t, n_init=1).fit(xc[idx])
        gamma = np.empty((N, K), order="F")
        for s in range(0, N, chunk):
            gamma[s:s + chunk] = gm.predict_proba(xc[s:s + chunk])
        zz = np.argmax(gamma, axis=1).astype(np.int32)
        if np.unique(zz).size == K:
            init_z, init_gamma = zz, gamma
            break
    if init_z is None:  # random fallback, R/fit_sampleqc.R:318-326
        fake = np.exp(rng.standard_normal((N, K)))
        init_gamma = np.asfortranarray(fake / fake.sum(axis=1, keepdims=True))
        init_z = np.argmax(init_gamma, axis=1).astype(np.int32)
    return init_z, init_gamma


# BASELINE.json configs (C1..C5)
CONFIGS = {
    "C1": dict(N=100_000, J=50, K_true=4, K=2, D=3, variant="robust"),
    "C2": dict(N=1_000_000, J=100, K_true=4, K=4, D=3, variant="robust"),
    "C3": dict(N=10_000_000, J=1000, K_true=4, K=4, D=3, variant="robust"),
    "C4": dict(N=5_000_000, J=500, K_true=6, K=6, D=6, variant="robust"),
    "C5": dict(N=50_000_000, J=5000, K_true=5, K=5, D=3, variant="robust"),
}


def make_config(name: str, scale: float = 1.0, seed: int | None = None):
    """Generate the inputs of a BASELINE config (optionally scaled down by `scale` in N and J).
    Returns (SimData, init_z, init_gamma, fit-params dict)."""
    cfg = CONFIGS[name]
    cid = int(name[1:])
    N = max(int(cfg["N"] * scale), 200)
    J = max(int(round(cfg["J"] * scale)), 2)
    sim = simulate_qcs(N, J, cfg["K_true"], cfg["D"], seed=(1000 + cid) if seed is None else seed)
    init_z, init_gamma = init_clusts(sim.x, sim.groups, J, cfg["K"], seed=cid)
    params = dict(D=cfg["D"], J=J, K=cfg["K"], N=N, em_iters=50, mcd_alpha=0.5, mcd_iters=50, seed=0)
    return sim, init_z, init_gamma, params
