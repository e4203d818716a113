/*
 * sqc_oracle.cpp — CPU ORACLE for the SampleQC EM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A dependency-free C++17 restatement, loop for loop, of the reference's arithmetic:
 *     src/fit_sampleQC_robust_cpp.cpp   (robust hard-assignment EM with FastMCD)
 *     src/fit_sampleQC_mle_cpp.cpp      (soft EM)
 *     src/helpers.cpp                   (seed, centring, initial alpha_j, reorder helpers)
 *     R/fit_sampleqc.R:419-453          (.calc_outliers_dt)
 * Every function cites the reference lines it follows.  Nothing under sampleqc_b200/ (the product)
 * may include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests — the reference holds no golden
 * vectors for this path and cannot be built here (R, Rcpp, Armadillo, RcppDist
 * absent).  What pins this oracle instead:
 *   (1) R's RNG known answers (set.seed(s); runif(3)) — tests/test_oracle_rng.py;
 *   (2) qchisq against scipy.stats.chi2.ppf;
 *   (3) oracle/_ref: the reference's OWN source files compiled against a minimal stand-in for the
 *       absent Armadillo/Rcpp/RcppDist headers (oracle/refshim/), run on the same inputs — this
 *       pins the control flow and every line of reference code, while the third-party numerics
 *       (randperm key recipe, dmvnorm, sort tie order) remain restated from their published
 *       definitions.
 *
 * Third-party arithmetic restated here (none is vendored under /root/reference, no versions pinned
 * in DESCRIPTION:20-58):
 *   R set.seed / unif_rand (Mersenne-Twister, R's initial scrambling)      — Appendix A1
 *   arma::randperm under RcppArmadillo: key_i = (int) Rf_runif(0, RAND_MAX) — Appendix A2
 *   RcppDist::dmvnorm: det(S), S.i() per call; P*exp(-q/2) or logP - q/2    — Appendix A4
 *   arma::inv / det / chol / solve(trimatl) / sort_index / mean             — textbook fp64
 *   R::qchisq                                                              — inverse reg. gamma
 */
#include "../include/sampleqc_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

constexpr int MAXD = 8;

struct OracleError : std::runtime_error {
  int code;
  OracleError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

static int g_cost_faithful = 0;  // 1: recompute det/inverse inside every density call (reference cost)

// ------------------------------------------------------------------------------------------
// R's Mersenne-Twister as reached through set_seed_cpp (helpers.cpp:10-14)  — Appendix A1
// ------------------------------------------------------------------------------------------
struct RMersenne {
  uint32_t mt[624];
  int mti;
  void set_seed(uint32_t seed) {          // R: RNG.c Randomize()/RNG_Init()/FixupSeeds()
    uint32_t s = seed;
    for (int j = 0; j < 50; ++j) s = 69069u * s + 1u;
    uint32_t i_seed[625];
    for (int j = 0; j < 625; ++j) { s = 69069u * s + 1u; i_seed[j] = s; }
    for (int j = 0; j < 624; ++j) mt[j] = i_seed[j + 1];
    mti = 624;                             // dummy[0] = 624: regenerate on first draw
  }
  uint32_t next_u32() {
    if (mti >= 624) {
      const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
      int kk;
      for (kk = 0; kk < 624 - 397; ++kk) {
        uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
      }
      for (; kk < 623; ++kk) {
        uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
      }
      uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
      mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  double unif_rand() {                     // MT_genrand() then fixup()
    double u = (double)next_u32() * 2.3283064365386963e-10;
    const double i2_32m1 = 2.328306437080797e-10;
    if (u <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - u) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return u;
  }
  int randi_key() {                        // arma_rng_alt::randi_val(): (int) Rf_runif(0, RAND_MAX)
    const double rand_max = 2147483647.0;
    double u = unif_rand();                // Rf_runif loops while u<=0 || u>=1: never after fixup
    return (int)(0.0 + (rand_max - 0.0) * u);
  }
};

// counter-based keys (SQC_RNG_COUNTER): identical recipe in the CUDA library
inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// key of the draw with 0-based index `draw` in the fit's key stream (same position as R's stream would
// have: call * N + cluster offset + rank inside the cluster)
inline int counter_key(uint32_t seed, uint32_t call, uint64_t draw) {
  uint64_t m = splitmix64(((uint64_t)seed << 32) ^ (uint64_t)call);
  uint64_t v = splitmix64(m ^ (draw * 0x9E3779B97F4A7C15ull));
  return (int)(v >> 33);
}

// ------------------------------------------------------------------------------------------
// qchisq(p, df) = 2 * qgamma(p, df/2): inverse regularised lower incomplete gamma
// (replaces R::qchisq, robust_cpp.cpp:102 and stats::qchisq, R/fit_sampleqc.R:429)
// ------------------------------------------------------------------------------------------
void reg_gamma_pq(double a, double x, double* P, double* Q) {   // P(a,x), Q(a,x): series / continued fraction
  if (x <= 0) { *P = 0.0; *Q = 1.0; return; }
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 2000; ++n) {
      ap += 1.0; del *= x / ap; sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    *P = sum * std::exp(-x + a * std::log(x) - lg); *Q = 1.0 - *P; return;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 2000; ++i) {
    double an = -i * (i - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    double del = d * c; h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  *Q = std::exp(-x + a * std::log(x) - lg) * h; *P = 1.0 - *Q;
}
double qchisq_oracle(double p, double df) {
  if (!(p >= 0.0) || !(p <= 1.0) || !(df > 0)) return std::numeric_limits<double>::quiet_NaN();
  if (p == 0.0) return 0.0;
  if (p == 1.0) return std::numeric_limits<double>::infinity();
  const double a = 0.5 * df;
  const bool upper = p > 0.5;              // solve on the smaller tail (1 - p is exact for p >= 0.5)
  const double target = upper ? 1.0 - p : p;
  auto resid = [&](double y) { double P, Q; reg_gamma_pq(a, y, &P, &Q); return upper ? (target - Q) : (P - target); };
  // bracket then bisect-safeguarded Newton on y = x/2 (resid is increasing in y)
  double lo = 0.0, hi = std::max(1.0, a);
  while (resid(hi) < 0) { lo = hi; hi *= 2.0; if (hi > 1e300) break; }
  double y = 0.5 * (lo + hi);
  const double lg = std::lgamma(a);
  for (int it = 0; it < 200; ++it) {
    double f = resid(y);
    if (f > 0) hi = y; else lo = y;
    double pdf = std::exp(-y + (a - 1.0) * std::log(y) - lg);
    double yn = (pdf > 0) ? y - f / pdf : 0.5 * (lo + hi);
    if (!(yn > lo && yn < hi)) yn = 0.5 * (lo + hi);
    if (std::fabs(yn - y) <= 4e-16 * std::fabs(yn)) { y = yn; break; }
    y = yn;
  }
  return 2.0 * y;
}

// ------------------------------------------------------------------------------------------
// tiny dense linear algebra (stand-ins for arma::det / inv / chol / solve on D x D)
// matrices are column-major D x D: m[r + D*c]
// ------------------------------------------------------------------------------------------
struct Mat { int n; double a[MAXD * MAXD]; double& operator()(int r, int c) { return a[r + n * c]; }
             double operator()(int r, int c) const { return a[r + n * c]; } };

double mat_det(const Mat& m) {             // arma::det: closed form for tiny, LU otherwise
  const int n = m.n;
  if (n == 1) return m(0, 0);
  if (n == 2) return m(0, 0) * m(1, 1) - m(0, 1) * m(1, 0);
  if (n == 3)
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1))
         - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0))
         + m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  Mat lu = m; double det = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c; double best = std::fabs(lu(c, c));
    for (int r = c + 1; r < n; ++r) if (std::fabs(lu(r, c)) > best) { best = std::fabs(lu(r, c)); piv = r; }
    if (best == 0.0) return 0.0;
    if (piv != c) { for (int k = 0; k < n; ++k) std::swap(lu(c, k), lu(piv, k)); det = -det; }
    det *= lu(c, c);
    for (int r = c + 1; r < n; ++r) {
      double f = lu(r, c) / lu(c, c);
      for (int k = c + 1; k < n; ++k) lu(r, k) -= f * lu(c, k);
    }
  }
  return det;
}
Mat mat_inv(const Mat& m) {                // arma::inv (general): Gauss-Jordan, partial pivoting
  const int n = m.n;
  Mat a = m, inv; inv.n = n;
  for (int i = 0; i < n * n; ++i) inv.a[i] = 0.0;
  for (int i = 0; i < n; ++i) inv(i, i) = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c; double best = std::fabs(a(c, c));
    for (int r = c + 1; r < n; ++r) if (std::fabs(a(r, c)) > best) { best = std::fabs(a(r, c)); piv = r; }
    if (!(best > 0.0) || !std::isfinite(best)) throw OracleError(SQC_ERR_SINGULAR, "inv(): matrix is singular");
    if (piv != c) for (int k = 0; k < n; ++k) { std::swap(a(c, k), a(piv, k)); std::swap(inv(c, k), inv(piv, k)); }
    double d = a(c, c);
    for (int k = 0; k < n; ++k) { a(c, k) /= d; inv(c, k) /= d; }
    for (int r = 0; r < n; ++r) if (r != c) {
      double f = a(r, c);
      if (f != 0.0) for (int k = 0; k < n; ++k) { a(r, k) -= f * a(c, k); inv(r, k) -= f * inv(c, k); }
    }
  }
  return inv;
}
Mat mat_chol_lower(const Mat& m) {         // L with L L' = m   (arma::chol(m).t())
  const int n = m.n; Mat L; L.n = n;
  for (int i = 0; i < n * n; ++i) L.a[i] = 0.0;
  for (int c = 0; c < n; ++c) {
    double s = m(c, c);
    for (int k = 0; k < c; ++k) s -= L(c, k) * L(c, k);
    if (!(s > 0.0) || !std::isfinite(s)) throw OracleError(SQC_ERR_NOT_SPD, "chol(): decomposition failed");
    double d = std::sqrt(s); L(c, c) = d;
    for (int r = c + 1; r < n; ++r) {
      double t = m(r, c);
      for (int k = 0; k < c; ++k) t -= L(r, k) * L(c, k);
      L(r, c) = t / d;
    }
  }
  return L;
}

// RcppDist::dmvnorm for one row (Appendix A4).  det/inverse are hoisted by the callers unless the
// cost-faithful mode asks for the reference's per-call recomputation (identical values either way).
struct Dens { Mat S_inv; double det_S; };
Dens dens_prepare(const Mat& S) { Dens d; d.det_S = mat_det(S); d.S_inv = mat_inv(S); return d; }
inline double quad_form(const double* xc, const Mat& A) {   // X * S_inv * X.t()
  const int n = A.n; double q = 0.0;
  double tmp[MAXD];
  for (int c = 0; c < n; ++c) { double s = 0.0; for (int r = 0; r < n; ++r) s += xc[r] * A(r, c); tmp[c] = s; }
  for (int c = 0; c < n; ++c) q += tmp[c] * xc[c];
  return q;
}
inline double dmvnorm1(const double* x, const double* mean, const Mat& S, const Dens& pre, bool logd) {
  const int m = S.n;
  const Dens* d = &pre; Dens local;
  if (g_cost_faithful) { local = dens_prepare(S); d = &local; }
  double xc[MAXD];
  for (int i = 0; i < m; ++i) xc[i] = x[i] - mean[i];
  const double q = quad_form(xc, d->S_inv);
  const double LN_2PI = 1.837877066409345483560659472811;
  if (logd) {
    double P = -1.0 * (m / 2.0) * LN_2PI - 0.5 * std::log(d->det_S);
    return P - 0.5 * q;
  }
  double P = 1.0 / std::sqrt(std::pow(6.283185307179586476925286766559, (double)m) * d->det_S);
  return P * std::exp(-0.5 * q);
}

// ------------------------------------------------------------------------------------------
// helpers.cpp
// ------------------------------------------------------------------------------------------
struct Problem {
  int D, J, K; int64_t N;
  std::vector<double> x;        // centred, column-major N x D
  std::vector<int32_t> groups;
  double X(int64_t i, int d) const { return x[i + N * d]; }
};

// arma::mean(x, 0): helpers / robust_cpp.cpp:478.  arma accumulates two interleaved partial sums.
void col_means(const double* x, int64_t N, int D, double* mu) {
  for (int d = 0; d < D; ++d) {
    const double* c = x + N * d; double a1 = 0.0, a2 = 0.0; int64_t i;
    for (i = 0; i + 1 < N; i += 2) { a1 += c[i]; a2 += c[i + 1]; }
    if (i < N) a1 += c[i];
    mu[d] = (a1 + a2) / (double)N;
  }
}
// centre_x, helpers.cpp:105-115
void centre_x(const double* x, const double* mu, int64_t N, int D, std::vector<double>& out) {
  out.resize((size_t)N * D);
  for (int d = 0; d < D; ++d) { double m = mu[d]; for (int64_t i = 0; i < N; ++i) out[i + N * d] = x[i + N * d] - m; }
}
// init_alpha_j, helpers.cpp:117-140   (alpha is J x D column-major)
void init_alpha_j(const Problem& P, std::vector<double>& alpha) {
  alpha.assign((size_t)P.J * P.D, 0.0);
  std::vector<uint64_t> n_js(P.J, 0);
  for (int64_t i = 0; i < P.N; ++i) {
    int j = P.groups[i]; n_js[j] += 1;
    for (int d = 0; d < P.D; ++d) alpha[j + P.J * d] += P.X(i, d);
  }
  for (int j = 0; j < P.J; ++j) for (int d = 0; d < P.D; ++d) alpha[j + P.J * d] /= (double)n_js[j];
}

// ------------------------------------------------------------------------------------------
// fit_sampleQC_robust_cpp.cpp statics
// ------------------------------------------------------------------------------------------
// calc_maha_dists, :11-24 — L = chol(cov).t(); solve L Y = (x - loc)'; d = colsum(Y^2)
void calc_maha_dists(const std::vector<double>& xk, int64_t n, int D, const double* loc, const Mat& cov,
                     std::vector<double>& d) {
  Mat L = mat_chol_lower(cov);
  d.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    double y[MAXD]; double s = 0.0;
    for (int r = 0; r < D; ++r) {
      double t = xk[i * D + r] - loc[r];
      for (int c = 0; c < r; ++c) t -= L(r, c) * y[c];
      y[r] = t / L(r, r);
      s += y[r] * y[r];
    }
    d[i] = s;
  }
}
// calc_loc, :25-35 (division inside the sum)
void calc_loc(const std::vector<double>& xk, const std::vector<int64_t>& h_idx, int64_t h, int D, double* loc) {
  for (int d = 0; d < D; ++d) loc[d] = 0.0;
  for (int64_t i = 0; i < h; ++i) { int64_t ix = h_idx[i]; for (int d = 0; d < D; ++d) loc[d] += xk[ix * D + d] / (double)h; }
}
// calc_scale, :36-48 (two-pass, biased 1/h)
void calc_scale(const std::vector<double>& xk, const double* loc, const std::vector<int64_t>& h_idx, int64_t h, int D, Mat& scale) {
  scale.n = D; for (int i = 0; i < D * D; ++i) scale.a[i] = 0.0;
  for (int64_t i = 0; i < h; ++i) {
    int64_t ix = h_idx[i];
    for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b)
      scale(a, b) += (xk[ix * D + a] - loc[a]) * (xk[ix * D + b] - loc[b]) / (double)h;
  }
}
// sort_dists, :49-52 — arma::sort_index ascending, first h.  NaN -> sort_index throws.
// Tie order of the unstable std::sort is unspecified; this repo defines lower index first.
void sort_dists(const std::vector<double>& d, int64_t h, std::vector<int64_t>& h_idx, std::vector<int64_t>& scratch) {
  const int64_t n = (int64_t)d.size();
  for (int64_t i = 0; i < n; ++i) if (std::isnan(d[i])) throw OracleError(SQC_ERR_NAN, "sort_index(): detected NaN");
  scratch.resize(n);
  for (int64_t i = 0; i < n; ++i) scratch[i] = i;
  std::sort(scratch.begin(), scratch.end(), [&](int64_t a, int64_t b) { return d[a] < d[b] || (d[a] == d[b] && a < b); });
  if (h < 1 || h > n) throw OracleError(SQC_ERR_SUBSET_SIZE, "span out of range in sort_dists");
  h_idx.assign(scratch.begin(), scratch.begin() + h);
}

struct KeySource {               // the stream behind arma::randperm (Appendix A2/A3)
  int mode; RMersenne mt; uint32_t seed; uint32_t call; int64_t draws;
  int key(int64_t /*original_cell*/) {
    const int64_t draw = draws++;
    if (mode == SQC_RNG_COUNTER) return counter_key(seed, 0u, (uint64_t)draw);
    return mt.randi_key();
  }
};

// fast_mcd, :55-108.  xk: n x D row-major residuals of one cluster (original cell order),
// cell_ids: original cell index of every row (only used by the counter-based key mode).
struct McdTrace { int iters; double q_dists; };
void fast_mcd(const std::vector<double>& xk, const std::vector<int64_t>& cell_ids, int D, int64_t N, int64_t h,
              int mcd_iters, KeySource& ks, double* loc, Mat& scale, McdTrace* trace, int* hit_max) {
  if (h > N) throw OracleError(SQC_ERR_SUBSET_SIZE, "randperm(): 'M' must be less than or equal to 'N'");
  if (h < 1) throw OracleError(SQC_ERR_SUBSET_SIZE, "fast_mcd: empty h-subset");
  const double eps = 1e-10;
  // h_idx = arma::randperm(N, h), :72 — one key per row, h smallest keys (ties: lower index)
  std::vector<std::pair<int, int64_t>> packets(N);
  for (int64_t i = 0; i < N; ++i) { packets[i].first = ks.key(cell_ids[i]); packets[i].second = i; }
  std::vector<int64_t> h_idx(h), scratch;
  if (h < N) std::partial_sort(packets.begin(), packets.begin() + h, packets.end());
  else       std::sort(packets.begin(), packets.end());
  for (int64_t i = 0; i < h; ++i) h_idx[i] = packets[i].second;
  packets.clear(); packets.shrink_to_fit();

  double loc_tmp[MAXD]; Mat scale_tmp;
  calc_loc(xk, h_idx, h, D, loc_tmp);                       // :75
  calc_scale(xk, loc_tmp, h_idx, h, D, scale_tmp);          // :76
  double det_new = mat_det(scale_tmp);                      // :77
  double det_old = 2 * det_new;                             // :80
  std::vector<double> maha;
  int iters = 0;
  while ((det_old - det_new > eps) & (det_new > 0) & (iters < mcd_iters)) {   // :84
    det_old = det_new;
    calc_maha_dists(xk, N, D, loc_tmp, scale_tmp, maha);    // :87
    sort_dists(maha, h, h_idx, scratch);                    // :88
    calc_loc(xk, h_idx, h, D, loc_tmp);                     // :89
    calc_scale(xk, loc_tmp, h_idx, h, D, scale_tmp);        // :90
    det_new = mat_det(scale_tmp);                           // :91
    iters++;
  }
  if (iters == mcd_iters && hit_max) (*hit_max)++;          // :94-95 (Rprintf)
  calc_maha_dists(xk, N, D, loc_tmp, scale_tmp, maha);      // :98
  sort_dists(maha, h, h_idx, scratch);                      // :99
  double q_dists = maha[h_idx[h - 1]];                      // :100
  double p = (double)h / N;                                 // :101
  double q_chisq = qchisq_oracle(p, D);                     // :102
  for (int i = 0; i < D * D; ++i) scale_tmp.a[i] = scale_tmp.a[i] * q_dists / q_chisq;   // :103
  for (int d = 0; d < D; ++d) loc[d] = loc_tmp[d];
  scale = scale_tmp;
  if (trace) { trace->iters = iters; trace->q_dists = q_dists; }
}

struct Params {
  std::vector<double> alpha;          // J x D col-major
  std::vector<double> beta;           // K x D col-major
  std::vector<Mat>    sigma;          // K
  std::vector<double> p_jk;           // J x K col-major
};

// update_alpha_j, :116-157
void update_alpha_j(const Problem& P, const std::vector<uint32_t>& z, Params& pr) {
  const int D = P.D, J = P.J, K = P.K;
  std::vector<Mat> inv_sigmas(K);
  for (int k = 0; k < K; ++k) inv_sigmas[k] = mat_inv(pr.sigma[k]);              // :126-128
  std::vector<double> numer((size_t)D * J, 0.0);
  std::vector<Mat> denom(J);
  for (int j = 0; j < J; ++j) { denom[j].n = D; for (int i = 0; i < D * D; ++i) denom[j].a[i] = 0.0; }
  for (int64_t i = 0; i < P.N; ++i) {                                            // :131-148
    int j = P.groups[i]; int k = (int)z[i];
    double v[MAXD];
    for (int d = 0; d < D; ++d) v[d] = P.X(i, d) - pr.beta[k + K * d];
    const Mat& A = inv_sigmas[k];
    for (int r = 0; r < D; ++r) { double s = 0.0; for (int c = 0; c < D; ++c) s += A(r, c) * v[c]; numer[r + D * j] += s; }
    for (int e = 0; e < D * D; ++e) denom[j].a[e] += A.a[e];
  }
  for (int j = 0; j < J; ++j) {                                                  // :151-154
    Mat di = mat_inv(denom[j]);
    for (int r = 0; r < D; ++r) { double s = 0.0; for (int c = 0; c < D; ++c) s += di(r, c) * numer[c + D * j]; pr.alpha[j + J * r] = s; }
  }
}

// restrict_to_x_k :265-283 + update_beta_scale_k :289-318
void update_beta_scale_k(const Problem& P, const std::vector<uint32_t>& z, int mcd_iters, double mcd_alpha,
                         KeySource& ks, Params& pr, int* hit_max) {
  const int D = P.D, K = P.K;
  std::vector<double> xk; std::vector<int64_t> ids;
  for (int k = 0; k < K; ++k) {
    xk.clear(); ids.clear();
    for (int64_t i = 0; i < P.N; ++i) if (z[i] == (unsigned)k) {               // :272-278
      for (int d = 0; d < D; ++d) xk.push_back(P.X(i, d) - pr.alpha[P.groups[i] + P.J * d]);
      ids.push_back(i);
    }
    int64_t n_k = (int64_t)ids.size();
    if (n_k == 0) throw OracleError(SQC_ERR_EMPTY_CLUSTER, "empty cluster: x_k.rows(0, -1) out of bounds");   // :307
    int h_k = (int)((n_k + D + 1) * mcd_alpha);                                 // :306 (truncation)
    double loc[MAXD]; Mat sc;
    fast_mcd(xk, ids, D, n_k, h_k, mcd_iters, ks, loc, sc, nullptr, hit_max);   // :307
    for (int d = 0; d < D; ++d) pr.beta[k + K * d] = loc[d];                    // :310
    pr.sigma[k] = sc;                                                           // :311
  }
  ks.call++;
}

// update_p_jk, :322-341
void update_p_jk(const Problem& P, const std::vector<uint32_t>& z, std::vector<double>& p_jk) {
  const int J = P.J, K = P.K;
  p_jk.assign((size_t)J * K, 1.0);
  std::vector<uint64_t> n_js(J, (uint64_t)K);
  for (int64_t i = 0; i < P.N; ++i) { int j = P.groups[i]; n_js[j] += 1; p_jk[j + J * z[i]] += 1; }
  for (int j = 0; j < J; ++j) for (int k = 0; k < K; ++k) p_jk[j + J * k] = p_jk[j + J * k] / (double)n_js[j];
}

// calc_expected_z, :344-396
void calc_expected_z(const Problem& P, const Params& pr, std::vector<uint32_t>& z, int* n_zeros_out) {
  const int D = P.D, J = P.J, K = P.K;
  std::vector<Dens> pre(K);
  for (int k = 0; k < K; ++k) pre[k] = dens_prepare(pr.sigma[k]);
  z.assign(P.N, 0u);
  int n_zeros = 0;
  double xi[MAXD], mu[MAXD], ll[64];
  for (int64_t i = 0; i < P.N; ++i) {
    for (int d = 0; d < D; ++d) xi[d] = P.X(i, d);
    int j = P.groups[i];
    for (int k = 0; k < K; ++k) {                                               // :360-365
      for (int d = 0; d < D; ++d) mu[d] = pr.alpha[j + J * d] + pr.beta[k + K * d];
      ll[k] = std::log(pr.p_jk[j + J * k]) + dmvnorm1(xi, mu, pr.sigma[k], pre[k], true);
    }
    double like_max = -std::numeric_limits<double>::infinity();                 // :371
    for (int k = 0; k < K; ++k) {
      if (ll[k] > like_max) { z[i] = (uint32_t)k; like_max = ll[k]; }           // :375-378
      if (like_max == -std::numeric_limits<double>::infinity()) n_zeros++;      // :380-382
    }
  }
  if (n_zeros_out) *n_zeros_out += n_zeros;
}

// calc_log_likelihood, :400-432
double calc_log_likelihood(const Problem& P, const Params& pr) {
  const int D = P.D, J = P.J, K = P.K;
  std::vector<Dens> pre(K);
  for (int k = 0; k < K; ++k) pre[k] = dens_prepare(pr.sigma[k]);
  double loglike = 0.0, xi[MAXD], mu[MAXD];
  for (int64_t i = 0; i < P.N; ++i) {
    for (int d = 0; d < D; ++d) xi[d] = P.X(i, d);
    double like_i = 0.0; int j = P.groups[i];
    for (int k = 0; k < K; ++k) {
      for (int d = 0; d < D; ++d) mu[d] = pr.alpha[j + J * d] + pr.beta[k + K * d];
      like_i += pr.p_jk[j + J * k] * dmvnorm1(xi, mu, pr.sigma[k], pre[k], false);
    }
    loglike += std::log(like_i);
  }
  return loglike;
}

// sort_index(beta_k.col(0)) :527 (ties: lower index first)
std::vector<int> order_by_first_col(const std::vector<double>& beta, int K) {
  std::vector<int> idx(K);
  for (int k = 0; k < K; ++k) idx[k] = k;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return beta[a] < beta[b]; });
  return idx;
}

void check_common(const sqc_fit_args* a) {
  if (!a || !a->x || !a->groups) throw OracleError(SQC_ERR_BAD_ARG, "null input");
  if (a->D < 1 || a->D > MAXD || a->K < 1 || a->K > 64 || a->J < 1 || a->N < 1) throw OracleError(SQC_ERR_BAD_ARG, "bad dimensions");
  for (int64_t i = 0; i < a->N; ++i) if (a->groups[i] < 0 || a->groups[i] >= a->J) throw OracleError(SQC_ERR_BAD_ARG, "groups out of range");
}

// fit_sampleqc_robust_cpp, :451-577
void fit_robust(const sqc_fit_args* a, sqc_fit_out* o) {
  check_common(a);
  if (!a->init_z) throw OracleError(SQC_ERR_BAD_ARG, "init_z missing");
  const int D = a->D, J = a->J, K = a->K; const int64_t N = a->N; const int em_iters = a->em_iters;
  if ((a->mcd_alpha < 0) | (a->mcd_alpha >= 1)) throw OracleError(SQC_ERR_BAD_ARG, "'mcd_alpha' must be between 0 and 1");   // :471
  KeySource ks; ks.mode = a->rng_mode; ks.seed = a->seed; ks.call = 0; ks.draws = 0;
  ks.mt.set_seed(a->seed);                                                       // :475
  Problem P; P.D = D; P.J = J; P.K = K; P.N = N;
  P.groups.assign(a->groups, a->groups + N);
  std::vector<double> mu_0(D);
  col_means(a->x, N, D, mu_0.data());                                            // :478
  std::vector<uint32_t> old_z(N), new_z(N, 0u);
  for (int64_t i = 0; i < N; ++i) { if (a->init_z[i] < 0 || a->init_z[i] >= K) throw OracleError(SQC_ERR_BAD_ARG, "init_z out of range"); old_z[i] = (uint32_t)a->init_z[i]; }
  centre_x(a->x, mu_0.data(), N, D, P.x);                                        // :480
  Params pr; pr.beta.assign((size_t)K * D, 0.0); pr.sigma.resize(K);
  for (int k = 0; k < K; ++k) { pr.sigma[k].n = D; for (int e = 0; e < D * D; ++e) pr.sigma[k].a[e] = 0.0; }
  update_p_jk(P, old_z, pr.p_jk);                                                // :481
  init_alpha_j(P, pr.alpha);                                                     // :482
  int hit_max = 0, n_zeros = 0;
  update_beta_scale_k(P, old_z, a->mcd_iters, a->mcd_alpha, ks, pr, &hit_max);   // :483
  std::vector<double> like_1(std::max(em_iters, 0), 0.0), like_2(std::max(em_iters, 0), 0.0);
  int64_t z_delta = 1; int iter = 0;
  while ((z_delta > 0) & (iter < em_iters)) {                                    // :490
    update_alpha_j(P, old_z, pr);                                                // :496
    like_1[iter] = calc_log_likelihood(P, pr);                                   // :499
    update_beta_scale_k(P, old_z, a->mcd_iters, a->mcd_alpha, ks, pr, &hit_max); // :502
    update_p_jk(P, old_z, pr.p_jk);                                              // :504
    like_2[iter] = calc_log_likelihood(P, pr);                                   // :506
    calc_expected_z(P, pr, new_z, &n_zeros);                                     // :509
    if (a->track) {                                                              // :511-516
      if (o->track_alpha_j) std::memcpy(o->track_alpha_j + (size_t)iter * J * D, pr.alpha.data(), sizeof(double) * J * D);
      if (o->track_beta_k)  std::memcpy(o->track_beta_k + (size_t)iter * K * D, pr.beta.data(), sizeof(double) * K * D);
      if (o->track_sigma_k) for (int k = 0; k < K; ++k) std::memcpy(o->track_sigma_k + (size_t)iter * D * D * K + (size_t)k * D * D, pr.sigma[k].a, sizeof(double) * D * D);
      if (o->track_p_jk)    std::memcpy(o->track_p_jk + (size_t)iter * J * K, pr.p_jk.data(), sizeof(double) * J * K);
    }
    z_delta = 0;
    for (int64_t i = 0; i < N; ++i) z_delta += (new_z[i] != old_z[i]);           // :519
    old_z = new_z;                                                               // :520
    iter++;
  }
  if (em_iters <= 0) new_z.assign(N, 0u);   // loop never ran: new_z stays zero-filled (:458)
  // :527-531 ascending relabel
  std::vector<int> k_order = order_by_first_col(pr.beta, K);
  if (o->mu_0) std::memcpy(o->mu_0, mu_0.data(), sizeof(double) * D);
  if (o->alpha_j) std::memcpy(o->alpha_j, pr.alpha.data(), sizeof(double) * J * D);
  if (o->beta_k) for (int k = 0; k < K; ++k) for (int d = 0; d < D; ++d) o->beta_k[k + K * d] = pr.beta[k_order[k] + K * d];
  if (o->sigma_k) for (int k = 0; k < K; ++k) std::memcpy(o->sigma_k + (size_t)k * D * D, pr.sigma[k_order[k]].a, sizeof(double) * D * D);
  if (o->p_jk) for (int k = 0; k < K; ++k) for (int j = 0; j < J; ++j) o->p_jk[j + J * k] = pr.p_jk[j + J * k_order[k]];
  if (o->z) {                                                                    // reorder_z :434-445, then + 1
    std::vector<int> inv(K);
    for (int k = 0; k < K; ++k) inv[k_order[k]] = k;
    for (int64_t i = 0; i < N; ++i) o->z[i] = inv[new_z[i]] + 1;
  }
  if (o->like_1) for (int i = 0; i < em_iters; ++i) o->like_1[i] = like_1[i];
  if (o->like_2) for (int i = 0; i < em_iters; ++i) o->like_2[i] = like_2[i];
  if (a->track) {                                                                // :533-537 reorders
    std::vector<double> tmp;
    for (int it = 0; it < em_iters; ++it) {
      if (o->track_beta_k) { double* b = o->track_beta_k + (size_t)it * K * D; tmp.assign(b, b + K * D);
        for (int k = 0; k < K; ++k) for (int d = 0; d < D; ++d) b[k + K * d] = tmp[k_order[k] + K * d]; }
      if (o->track_sigma_k) { double* s = o->track_sigma_k + (size_t)it * D * D * K; tmp.assign(s, s + D * D * K);
        for (int k = 0; k < K; ++k) std::memcpy(s + (size_t)k * D * D, tmp.data() + (size_t)k_order[k] * D * D, sizeof(double) * D * D); }
      if (o->track_p_jk) { double* p = o->track_p_jk + (size_t)it * J * K; tmp.assign(p, p + J * K);
        for (int k = 0; k < K; ++k) for (int j = 0; j < J; ++j) p[j + J * k] = tmp[j + J * k_order[k]]; }
    }
  }
  o->n_iter = iter; o->mcd_iters_hit = hit_max; o->n_zero_like = n_zeros; o->rng_draws = ks.draws;
  o->rng_state[0] = ks.mt.mti;
  for (int i = 0; i < 624; ++i) o->rng_state[i + 1] = (int32_t)ks.mt.mt[i];
}

// ------------------------------------------------------------------------------------------
// fit_sampleQC_mle_cpp.cpp
// ------------------------------------------------------------------------------------------
// calc_expected_gamma_i, :151-187 (global p_k, plain densities, direct normalisation)
void mle_expected_gamma(const Problem& P, const Params& pr, const std::vector<double>& p_k, std::vector<double>& gamma) {
  const int D = P.D, J = P.J, K = P.K; const int64_t N = P.N;
  std::vector<Dens> pre(K);
  for (int k = 0; k < K; ++k) pre[k] = dens_prepare(pr.sigma[k]);
  gamma.resize((size_t)N * K);
  double xi[MAXD], mu[MAXD];
  for (int64_t i = 0; i < N; ++i) {
    for (int d = 0; d < D; ++d) xi[d] = P.X(i, d);
    double gsum = 0.0; int j = P.groups[i];
    for (int k = 0; k < K; ++k) {
      for (int d = 0; d < D; ++d) mu[d] = pr.alpha[j + J * d] + pr.beta[k + K * d];
      double like = p_k[k] * dmvnorm1(xi, mu, pr.sigma[k], pre[k], false);
      gamma[i + N * k] = like; gsum += like;
    }
    for (int k = 0; k < K; ++k) gamma[i + N * k] = gamma[i + N * k] / gsum;
  }
}
// max_alpha_j_given_others, :26-64
void mle_max_alpha(const Problem& P, const std::vector<double>& gamma, Params& pr) {
  const int D = P.D, J = P.J, K = P.K; const int64_t N = P.N;
  std::vector<Mat> inv_sigmas(K);
  for (int k = 0; k < K; ++k) inv_sigmas[k] = mat_inv(pr.sigma[k]);
  std::vector<double> numer((size_t)D * J, 0.0);
  std::vector<Mat> denom(J);
  for (int j = 0; j < J; ++j) { denom[j].n = D; for (int e = 0; e < D * D; ++e) denom[j].a[e] = 0.0; }
  for (int64_t i = 0; i < N; ++i) {
    int j = P.groups[i];
    for (int k = 0; k < K; ++k) {
      double g = gamma[i + N * k]; const Mat& A = inv_sigmas[k]; double v[MAXD];
      for (int d = 0; d < D; ++d) v[d] = P.X(i, d) - pr.beta[k + K * d];
      // gamma * inv_sigma * v: arma evaluates (gamma * inv_sigma) * v as a scaled mat-vec
      for (int r = 0; r < D; ++r) { double s = 0.0; for (int c = 0; c < D; ++c) s += A(r, c) * v[c]; numer[r + D * j] += g * s; }
      for (int e = 0; e < D * D; ++e) denom[j].a[e] += g * A.a[e];
    }
  }
  for (int j = 0; j < J; ++j) {
    Mat di = mat_inv(denom[j]);
    for (int r = 0; r < D; ++r) { double s = 0.0; for (int c = 0; c < D; ++c) s += di(r, c) * numer[c + D * j]; pr.alpha[j + J * r] = s; }
  }
}
// max_beta_k_given_others, :65-91
void mle_max_beta(const Problem& P, const std::vector<double>& gamma, Params& pr) {
  const int D = P.D, J = P.J, K = P.K; const int64_t N = P.N;
  std::vector<double> numer((size_t)K * D, 0.0), denom(K, 0.0);
  for (int64_t i = 0; i < N; ++i) for (int k = 0; k < K; ++k) {
    for (int d = 0; d < D; ++d) numer[k + K * d] += gamma[i + N * k] * (P.X(i, d) - pr.alpha[P.groups[i] + J * d]);
    denom[k] += gamma[i + N * k];
  }
  for (int k = 0; k < K; ++k) for (int d = 0; d < D; ++d) pr.beta[k + K * d] = numer[k + K * d] / denom[k];
}
// max_sigma_k_given_others, :92-129
void mle_max_sigma(const Problem& P, const std::vector<double>& gamma, Params& pr) {
  const int D = P.D, J = P.J, K = P.K; const int64_t N = P.N;
  std::vector<Mat> numer(K); std::vector<double> denom(K, 0.0);
  for (int k = 0; k < K; ++k) { numer[k].n = D; for (int e = 0; e < D * D; ++e) numer[k].a[e] = 0.0; }
  double tmp[MAXD];
  for (int64_t i = 0; i < N; ++i) for (int k = 0; k < K; ++k) {
    double g = gamma[i + N * k];
    for (int d = 0; d < D; ++d) tmp[d] = P.X(i, d) - pr.alpha[P.groups[i] + J * d] - pr.beta[k + K * d];
    for (int d1 = 0; d1 < D; ++d1) for (int d2 = 0; d2 < D; ++d2) numer[k](d1, d2) += g * tmp[d1] * tmp[d2];
    denom[k] += g;
  }
  for (int k = 0; k < K; ++k) { pr.sigma[k].n = D; for (int e = 0; e < D * D; ++e) pr.sigma[k].a[e] = numer[k].a[e] / denom[k]; }
}
// max_p_k_given_others, :130-147
void mle_max_p(const std::vector<double>& gamma, int64_t N, int K, std::vector<double>& p_k) {
  p_k.assign(K, 0.0); double tot = 0.0;
  for (int64_t i = 0; i < N; ++i) for (int k = 0; k < K; ++k) { p_k[k] += gamma[i + N * k]; tot += gamma[i + N * k]; }
  for (int k = 0; k < K; ++k) p_k[k] = p_k[k] / tot;
}
// calc_log_likelihood (mle overload), :191-222
double mle_log_likelihood(const Problem& P, const Params& pr, const std::vector<double>& p_k) {
  const int D = P.D, J = P.J, K = P.K;
  std::vector<Dens> pre(K);
  for (int k = 0; k < K; ++k) pre[k] = dens_prepare(pr.sigma[k]);
  double loglike = 0.0, xi[MAXD], mu[MAXD];
  for (int64_t i = 0; i < P.N; ++i) {
    for (int d = 0; d < D; ++d) xi[d] = P.X(i, d);
    double like_i = 0.0; int j = P.groups[i];
    for (int k = 0; k < K; ++k) {
      for (int d = 0; d < D; ++d) mu[d] = pr.alpha[j + J * d] + pr.beta[k + K * d];
      like_i += p_k[k] * dmvnorm1(xi, mu, pr.sigma[k], pre[k], false);
    }
    loglike += std::log(like_i);
  }
  return loglike;
}

// fit_sampleqc_mle_cpp, :279-357
void fit_mle(const sqc_fit_args* a, sqc_fit_out* o) {
  check_common(a);
  if (!a->init_gamma) throw OracleError(SQC_ERR_BAD_ARG, "init_gamma missing");
  const int D = a->D, J = a->J, K = a->K; const int64_t N = a->N; const int n_iter = a->em_iters;
  Problem P; P.D = D; P.J = J; P.K = K; P.N = N;
  P.groups.assign(a->groups, a->groups + N);
  std::vector<double> mu_0(D);
  col_means(a->x, N, D, mu_0.data());                                            // :297
  centre_x(a->x, mu_0.data(), N, D, P.x);                                        // :298
  std::vector<double> init_gamma(a->init_gamma, a->init_gamma + (size_t)N * K), gamma((size_t)N * K, 0.0), p_k;
  Params pr; pr.beta.assign((size_t)K * D, 0.0); pr.sigma.resize(K);
  init_alpha_j(P, pr.alpha);                                                     // :300
  mle_max_beta(P, init_gamma, pr);                                               // :301
  mle_max_sigma(P, init_gamma, pr);                                              // :302
  mle_max_p(init_gamma, N, K, p_k);                                              // :303 (:299 init_p_k is overwritten)
  std::vector<double> like_1(std::max(n_iter, 0), 0.0), like_2(std::max(n_iter, 0), 0.0);
  for (int i = 0; i < n_iter; ++i) {                                             // :307
    mle_expected_gamma(P, pr, p_k, gamma);                                       // :313
    mle_max_alpha(P, gamma, pr);                                                 // :316
    like_1[i] = mle_log_likelihood(P, pr, p_k);                                  // :317
    mle_expected_gamma(P, pr, p_k, gamma);                                       // :320
    mle_max_beta(P, gamma, pr);                                                  // :323
    mle_max_sigma(P, gamma, pr);                                                 // :324
    mle_max_p(gamma, N, K, p_k);                                                 // :325
    like_2[i] = mle_log_likelihood(P, pr, p_k);                                  // :326
  }
  // extract_p_jk, :248-273
  std::vector<double> p_jk((size_t)J * K, 0.0), p_sums(J, 0.0);
  for (int64_t i = 0; i < N; ++i) { int j = P.groups[i]; for (int k = 0; k < K; ++k) { p_jk[j + J * k] += gamma[i + N * k]; p_sums[j] += gamma[i + N * k]; } }
  for (int j = 0; j < J; ++j) for (int k = 0; k < K; ++k) p_jk[j + J * k] = p_jk[j + J * k] / p_sums[j];
  std::vector<int> k_order = order_by_first_col(pr.beta, K);                     // :334
  if (o->mu_0) std::memcpy(o->mu_0, mu_0.data(), sizeof(double) * D);
  if (o->alpha_j) std::memcpy(o->alpha_j, pr.alpha.data(), sizeof(double) * J * D);
  if (o->beta_k) for (int k = 0; k < K; ++k) for (int d = 0; d < D; ++d) o->beta_k[k + K * d] = pr.beta[k_order[k] + K * d];
  if (o->sigma_k) for (int k = 0; k < K; ++k) std::memcpy(o->sigma_k + (size_t)k * D * D, pr.sigma[k_order[k]].a, sizeof(double) * D * D);
  if (o->z) {                                                                    // extract_z, :224-246: z = idx(k_max) + 1
    for (int64_t i = 0; i < N; ++i) {
      double g_max = -1; int k_max = -1;
      for (int k = 0; k < K; ++k) if (gamma[i + N * k] > g_max) { g_max = gamma[i + N * k]; k_max = k; }
      if (k_max < 0) throw OracleError(SQC_ERR_NAN, "extract_z: index out of bounds (all gamma NaN)");
      o->z[i] = k_order[k_max] + 1;
    }
  }
  if (o->gamma_i) std::memcpy(o->gamma_i, gamma.data(), sizeof(double) * (size_t)N * K);   // :350 (not reordered)
  if (o->p_k) for (int k = 0; k < K; ++k) o->p_k[k] = p_k[k_order[k]];           // :338
  if (o->p_jk) for (int k = 0; k < K; ++k) for (int j = 0; j < J; ++j) o->p_jk[j + J * k] = p_jk[j + J * k_order[k]];   // :339
  if (o->like_1) for (int i = 0; i < n_iter; ++i) o->like_1[i] = like_1[i];
  if (o->like_2) for (int i = 0; i < n_iter; ++i) o->like_2[i] = like_2[i];
  o->n_iter = std::max(n_iter, 0); o->mcd_iters_hit = 0; o->n_zero_like = 0; o->rng_draws = 0;
}

// .calc_outliers_dt, R/fit_sampleqc.R:419-453 ; stats::mahalanobis uses solve(cov)
void outlier_maha(const sqc_maha_args* a, double* maha, uint8_t* outlier, double* maha_cut) {
  const int D = a->D, J = a->J, K = a->K; const int64_t N = a->N;
  if (D < 1 || D > MAXD || K < 1 || K > 64) throw OracleError(SQC_ERR_BAD_ARG, "bad dimensions");
  const double cut = qchisq_oracle(1.0 - a->alpha, (double)D);                   // :429
  if (maha_cut) *maha_cut = cut;
  std::vector<Mat> inv(K);
  for (int k = 0; k < K; ++k) { Mat S; S.n = D; std::memcpy(S.a, a->sigma_k + (size_t)k * D * D, sizeof(double) * D * D); inv[k] = mat_inv(S); }
  for (int64_t i = 0; i < N; ++i) {
    int j = a->groups[i]; double xc[MAXD], v[MAXD];
    for (int d = 0; d < D; ++d) xc[d] = (a->x[i + N * d] - a->mu_0[d]) - a->alpha_j[j + J * d];   // :432
    double mn = std::numeric_limits<double>::infinity();
    for (int k = 0; k < K; ++k) {
      for (int d = 0; d < D; ++d) v[d] = xc[d] - a->beta_k[k + K * d];
      double q = quad_form(v, inv[k]);                                           // rowSums(x %*% cov * x)
      if (maha) maha[i + N * k] = q;
      if (q < mn) mn = q;
    }
    if (outlier) outlier[i] = (mn > cut) ? 1 : 0;                                // :449
  }
}

template <class F> int guarded(F&& f, char* err, size_t errlen) {
  try { f(); if (err && errlen) err[0] = 0; return SQC_OK; }
  catch (const OracleError& e) { if (err && errlen) std::snprintf(err, errlen, "%s", e.what()); return e.code; }
  catch (const std::exception& e) { if (err && errlen) std::snprintf(err, errlen, "%s", e.what()); return SQC_ERR_BAD_ARG; }
}

}  // namespace

extern "C" {

int sqc_oracle_fit_robust(const sqc_fit_args* a, sqc_fit_out* o, char* err, size_t errlen) { return guarded([&] { fit_robust(a, o); }, err, errlen); }
int sqc_oracle_fit_mle(const sqc_fit_args* a, sqc_fit_out* o, char* err, size_t errlen) { return guarded([&] { fit_mle(a, o); }, err, errlen); }
int sqc_oracle_outlier_maha(const sqc_maha_args* a, double* maha, uint8_t* outlier, double* cut, char* err, size_t errlen) {
  return guarded([&] { outlier_maha(a, maha, outlier, cut); }, err, errlen);
}
void sqc_oracle_set_cost_faithful(int on) { g_cost_faithful = on; }
// set.seed(seed); runif(n)
void sqc_oracle_runif(uint32_t seed, double* out, int64_t n) { RMersenne m; m.set_seed(seed); for (int64_t i = 0; i < n; ++i) out[i] = m.unif_rand(); }
// the randperm key stream: skip draws, then n keys
void sqc_oracle_rkeys(uint32_t seed, int64_t skip, int32_t* out, int64_t n) {
  RMersenne m; m.set_seed(seed);
  for (int64_t i = 0; i < skip; ++i) m.next_u32();
  for (int64_t i = 0; i < n; ++i) out[i] = m.randi_key();
}
int32_t sqc_oracle_counter_key(uint32_t seed, uint32_t call, uint64_t cell) { return counter_key(seed, call, cell); }
double sqc_oracle_qchisq(double p, double df) { return qchisq_oracle(p, df); }
// one fast_mcd on an n x D row-major residual matrix (tests): returns status
int sqc_oracle_fast_mcd(const double* xk, int64_t n, int32_t D, int64_t h, int32_t mcd_iters, uint32_t seed,
                        double* loc, double* scale, int32_t* iters, double* q_dists, char* err, size_t errlen) {
  return guarded([&] {
    KeySource ks; ks.mode = SQC_RNG_R_COMPAT; ks.seed = seed; ks.call = 0; ks.draws = 0; ks.mt.set_seed(seed);
    std::vector<double> v(xk, xk + (size_t)n * D); std::vector<int64_t> ids(n);
    for (int64_t i = 0; i < n; ++i) ids[i] = i;
    Mat sc; McdTrace tr; int hit = 0;
    fast_mcd(v, ids, D, n, h, mcd_iters, ks, loc, sc, &tr, &hit);
    std::memcpy(scale, sc.a, sizeof(double) * D * D);
    if (iters) *iters = tr.iters;
    if (q_dists) *q_dists = tr.q_dists;
  }, err, errlen);
}

}  // extern "C"
