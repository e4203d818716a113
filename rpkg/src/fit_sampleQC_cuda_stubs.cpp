// fit_sampleQC_cuda_stubs.cpp — the Rcpp host side of the B200 path.
//
// Drop-in replacements for the BODIES of the two exports of the reference package
//     fit_sampleqc_robust_cpp   (reference src/fit_sampleQC_robust_cpp.cpp:451-577)
//     fit_sampleqc_mle_cpp      (reference src/fit_sampleQC_mle_cpp.cpp:279-357)
// The generated glue (reference src/RcppExports.cpp:16-54, registration table :229-252) and the R closures
// (reference R/RcppExports.R:7-16) stay byte-for-byte as they are: same C++ signatures, same argument
// conversion, same Rcpp::List names / shapes, errors raised through Rcpp::stop (-> END_RCPP -> R error).
// All arithmetic happens in libsampleqc_cuda.so behind the C ABI of include/sampleqc_cuda.h; there is no
// CPU fallback.  This file replaces src/fit_sampleQC_robust_cpp.cpp and src/fit_sampleQC_mle_cpp.cpp in the
// package; src/helpers.cpp, src/mmd_fn.cpp, src/subsample_mmd_fn.cpp are untouched.
//
// SOURCE ONLY in this repository: R, Rcpp and RcppArmadillo are absent from the build image, so the same
// marshalling is exercised through the ctypes mirror sampleqc_b200/api.py in tests/.
//
// [[Rcpp::depends(RcppArmadillo)]]
#include <RcppArmadillo.h>
#include <vector>

#include "sampleqc_cuda.h"

using namespace Rcpp;

namespace {

// R's RNG state after the reference would have consumed (1 + n_iter) * N unif_rand draws, so that code running
// after the fit sees the stream where the reference leaves it (the reference calls set.seed(seed) through
// Rcpp::Function, helpers.cpp:10-14, and Rcpp::RNGScope brackets the export, RcppExports.cpp:38).
// The export runs inside that RNGScope: its destructor calls PutRNGstate(), which OVERWRITES .Random.seed with R's
// internal C state.  Writing the variable alone would therefore be lost; after the assignment GetRNGstate() is
// called so that the internal state is re-read from .Random.seed — PutRNGstate() at scope exit then writes back
// exactly what was set here.
void restore_r_rng(unsigned int seed, const sqc_fit_out& out) {
  Environment base = Environment::namespace_env("base");
  Function set_seed = base["set.seed"];
  set_seed(seed);                                   // selects Mersenne-Twister exactly as the reference does
  Environment global = Environment::global_env();
  IntegerVector rs = clone(as<IntegerVector>(global[".Random.seed"]));   // [0] = RNG kind code, [1] = mti, [2..625] = mt
  for (int i = 0; i < 625; ++i) rs[1 + i] = out.rng_state[i];
  global[".Random.seed"] = rs;
  GetRNGstate();                                    // internal state <- .Random.seed (R_ext/Random.h)
}

// options(SampleQC.n_gpus = n): 1 (default) = one GPU, n > 1 = that many, -1 = all visible (SQC_ALL_GPUS).  The export
// signatures stay those of the reference; fit_sampleqc() needs no change.
int n_gpus_option() {
  Function get_option("getOption");
  return as<int>(get_option("SampleQC.n_gpus", 1));
}

void check_status(int status, const char* err) {
  if (status != SQC_OK) Rcpp::stop("%s [%s]", err, sqc_status_name(status));
}

}  // namespace

// [[Rcpp::export]]
List fit_sampleqc_robust_cpp(arma::mat x, arma::uvec init_z, arma::uvec groups, int D, int J, int K, int N,
                             int em_iters, double mcd_alpha, int mcd_iters, unsigned int seed, bool track = false) {
  // arma::uvec (64-bit unsigned) -> int32, as the library expects
  std::vector<int32_t> g(N), z0(N);
  for (int i = 0; i < N; ++i) { g[i] = (int32_t)groups(i); z0[i] = (int32_t)init_z(i); }
  const int em = em_iters > 0 ? em_iters : 0;

  sqc_fit_args a = {};
  a.abi_version = SQC_ABI_VERSION;
  a.D = D; a.J = J; a.K = K; a.N = N;
  a.x = x.memptr();                                 // column-major N x D, un-centred: handed through as is
  a.groups = g.data(); a.init_z = z0.data();
  a.em_iters = em_iters; a.mcd_iters = mcd_iters; a.mcd_alpha = mcd_alpha;
  a.seed = seed; a.track = track ? 1 : 0;
  a.rng_mode = SQC_RNG_R_COMPAT;                    // arma::randperm keys from R's stream (robust_cpp.cpp:72)
  a.device = -1; a.world = 1; a.N_total = N;
  a.n_gpus = n_gpus_option();                       // options(SampleQC.n_gpus = n): one call sharded over n GPUs by the library

  arma::vec mu_0(D), like_1(em, arma::fill::zeros), like_2(em, arma::fill::zeros);
  arma::mat alpha_j(J, D), beta_k(K, D), p_jk(J, K);
  arma::cube sigma_k(D, D, K);
  std::vector<int32_t> z(N);
  arma::cube track_alpha_j, track_beta_k, track_sigma_k, track_p_jk;
  sqc_fit_out o = {};
  o.mu_0 = mu_0.memptr(); o.alpha_j = alpha_j.memptr(); o.beta_k = beta_k.memptr(); o.sigma_k = sigma_k.memptr();
  o.z = z.data(); o.p_jk = p_jk.memptr(); o.like_1 = like_1.memptr(); o.like_2 = like_2.memptr();
  if (track) {
    track_alpha_j.zeros(J, D, em); track_beta_k.zeros(K, D, em); track_sigma_k.zeros(D * D, K, em); track_p_jk.zeros(J, K, em);
    o.track_alpha_j = track_alpha_j.memptr(); o.track_beta_k = track_beta_k.memptr();
    o.track_sigma_k = track_sigma_k.memptr(); o.track_p_jk = track_p_jk.memptr();
  }
  char err[512];
  check_status(sqc_fit_robust(&a, &o, err, sizeof err), err);
  restore_r_rng(seed, o);
  Rprintf("\ntook %d iterations\n", o.n_iter);      // robust_cpp.cpp:524

  arma::uvec z_out(N);                              // the list's z: 1-based labels, wrapped as double like arma::uvec
  for (int i = 0; i < N; ++i) z_out(i) = (arma::uword)z[i];
  List res = List::create(                          // names and order of robust_cpp.cpp:562-575
      Named("D") = D, Named("J") = J, Named("K") = K, Named("N") = N,
      Named("mu_0") = mu_0, Named("alpha_j") = alpha_j, Named("beta_k") = beta_k, Named("sigma_k") = sigma_k,
      Named("z") = z_out, Named("p_jk") = p_jk, Named("like_1") = like_1, Named("like_2") = like_2);
  if (track) {                                      // :542-559
    res["track_alpha_j"] = track_alpha_j; res["track_beta_k"] = track_beta_k;
    res["track_sigma_k"] = track_sigma_k; res["track_p_jk"] = track_p_jk;
  }
  return res;
}

// [[Rcpp::export]]
List fit_sampleqc_mle_cpp(arma::mat x, arma::mat init_gamma_i, arma::uvec groups, int D, int J, int K, int N,
                          int n_iter, unsigned int seed) {
  std::vector<int32_t> g(N);
  for (int i = 0; i < N; ++i) g[i] = (int32_t)groups(i);
  const int it = n_iter > 0 ? n_iter : 0;

  sqc_fit_args a = {};
  a.abi_version = SQC_ABI_VERSION;
  a.D = D; a.J = J; a.K = K; a.N = N;
  a.x = x.memptr(); a.groups = g.data(); a.init_gamma = init_gamma_i.memptr();
  a.em_iters = n_iter; a.seed = seed; a.device = -1; a.world = 1; a.N_total = N; a.n_gpus = n_gpus_option();

  arma::vec mu_0(D), p_k(K), like_1(it, arma::fill::zeros), like_2(it, arma::fill::zeros);
  arma::mat alpha_j(J, D), beta_k(K, D), p_jk(J, K), gamma_i(N, K);
  arma::cube sigma_k(D, D, K);
  std::vector<int32_t> z(N);
  sqc_fit_out o = {};
  o.mu_0 = mu_0.memptr(); o.alpha_j = alpha_j.memptr(); o.beta_k = beta_k.memptr(); o.sigma_k = sigma_k.memptr();
  o.z = z.data(); o.p_jk = p_jk.memptr(); o.like_1 = like_1.memptr(); o.like_2 = like_2.memptr();
  o.gamma_i = gamma_i.memptr(); o.p_k = p_k.memptr();
  char err[512];
  check_status(sqc_fit_mle(&a, &o, err, sizeof err), err);
  // the reference calls set_seed_cpp(seed) (mle_cpp.cpp:294) and draws nothing afterwards
  Environment base = Environment::namespace_env("base");
  Function set_seed = base["set.seed"];
  set_seed(seed);

  arma::uvec z_out(N);
  for (int i = 0; i < N; ++i) z_out(i) = (arma::uword)z[i];
  return List::create(                              // names and order of mle_cpp.cpp:341-356
      Named("D") = D, Named("J") = J, Named("K") = K, Named("N") = N,
      Named("mu_0") = mu_0, Named("alpha_j") = alpha_j, Named("beta_k") = beta_k, Named("sigma_k") = sigma_k,
      Named("gamma_i") = gamma_i, Named("z") = z_out, Named("p_k") = p_k, Named("p_jk") = p_jk,
      Named("like_1") = like_1, Named("like_2") = like_2);
}

// Optional accelerator for .calc_outliers_dt (reference R/fit_sampleqc.R:419-453): returns the N x K
// squared Mahalanobis distances and the outlier flags.  Not part of the reference's export table — the R
// implementation stays valid; see INTEGRATION.md for the one-line switch.
// [[Rcpp::export]]
List calc_outliers_cuda(arma::mat x, arma::uvec groups, arma::vec mu_0, arma::mat alpha_j, arma::mat beta_k,
                        arma::cube sigma_k, double alpha) {
  const int N = x.n_rows, D = x.n_cols, J = alpha_j.n_rows, K = beta_k.n_rows;
  std::vector<int32_t> g(N);
  for (int i = 0; i < N; ++i) g[i] = (int32_t)groups(i) - 1;     // .calc_outliers_dt receives the 1-based groups
  sqc_maha_args a = {};
  a.abi_version = SQC_ABI_VERSION; a.D = D; a.J = J; a.K = K; a.N = N;
  a.x = x.memptr(); a.groups = g.data(); a.mu_0 = mu_0.memptr(); a.alpha_j = alpha_j.memptr();
  a.beta_k = beta_k.memptr(); a.sigma_k = sigma_k.memptr(); a.alpha = alpha; a.device = -1;
  arma::mat maha(N, K);
  std::vector<uint8_t> out(N);
  double cut = 0.0;
  char err[512];
  check_status(sqc_outlier_maha(&a, maha.memptr(), out.data(), &cut, err, sizeof err), err);
  LogicalVector outlier(N);
  for (int i = 0; i < N; ++i) outlier[i] = out[i] != 0;
  return List::create(Named("maha") = maha, Named("outlier") = outlier, Named("maha_cut") = cut);
}

// Optional accelerator for the O(N) parts of .init_clusts (reference R/fit_sampleqc.R:277-331): grp_medians (:286-289)
// and — given the mixture mclust fitted on the 1e4-cell subsample — the posterior and classification of all N cells
// (what summaryMclustBIC(...)$z / $classification deliver, :303-316).  pro: K, mean: K x D, sigma: D x D x K as
// mclust's parameters$pro / t(parameters$mean) / parameters$variance$sigma.  Pass pro of length 0 for the medians only.
// [[Rcpp::export]]
List init_clusts_cuda(arma::mat x, arma::uvec groups, int J, arma::vec pro, arma::mat mean, arma::cube sigma) {
  const int N = x.n_rows, D = x.n_cols, K = pro.n_elem;
  std::vector<int32_t> g(N);
  for (int i = 0; i < N; ++i) g[i] = (int32_t)groups(i) - 1;     // .init_clusts receives the 1-based groups
  sqc_init_args a = {};
  a.abi_version = SQC_ABI_VERSION; a.D = D; a.J = J; a.K = K; a.N = N;
  a.x = x.memptr(); a.groups = g.data(); a.device = -1;
  arma::mat medians(J, D), gamma;
  std::vector<int32_t> z;
  if (K > 0) { a.pro = pro.memptr(); a.mean = mean.memptr(); a.sigma = sigma.memptr(); gamma.set_size(N, K); z.resize(N); }
  char err[512];
  check_status(sqc_init_clusts(&a, medians.memptr(), K > 0 ? gamma.memptr() : nullptr, K > 0 ? z.data() : nullptr, err, sizeof err), err);
  if (K == 0) return List::create(Named("grp_medians") = medians);
  IntegerVector init_z(N);
  for (int i = 0; i < N; ++i) init_z[i] = z[i];                   // 0-based, as init_z = classification - 1 (:314)
  return List::create(Named("grp_medians") = medians, Named("init_gamma") = gamma, Named("init_z") = init_z);
}

// mmd_fn (reference src/mmd_fn.cpp:20-91): same signature, same errors; the three kernel sums run on the device.
// Replaces src/mmd_fn.cpp in the package (optional: the reference file stays valid).
// [[Rcpp::export]]
double mmd_fn(arma::mat X, arma::mat Y, double sigma) {
  if (sigma <= 0) throw std::invalid_argument("sigma must be > 0");                            // :28-30
  if (X.n_cols != Y.n_cols) throw std::invalid_argument("X and Y must have same number of columns");   // :31-33
  double mmd = 0.0;
  char err[512];
  check_status(sqc_mmd_fn(X.memptr(), X.n_rows, Y.memptr(), Y.n_rows, (int)X.n_cols, sigma, &mmd, -1, err, sizeof err), err);
  return mmd;
}

// .calc_mmd_mat's pair loop (reference R/calc_pairwise_mmds.R:216-233) in ONE device call: mat_list as in R (one
// matrix per sample), pairs (i < j, 1-based as ij_grid holds them), and the row picks .dist_fn's sample() draws
// (:271-277) pre-drawn by the R caller in the loop's order — idx_i / idx_j: n_pairs * n_times * subsample, 1-based,
// ignored for samples with <= subsample rows.  Returns mean(abs(mmd)) per pair (:231, :279).
// [[Rcpp::export]]
NumericVector calc_mmd_pairs_cuda(List mat_list, IntegerVector pair_i, IntegerVector pair_j, int n_times, int subsample,
                                  IntegerVector idx_i, IntegerVector idx_j, double sigma) {
  const int S = mat_list.size(); const R_xlen_t P = pair_i.size();
  std::vector<int64_t> rows(S); std::vector<double> flat; int D = 0;
  for (int s = 0; s < S; ++s) { NumericMatrix m = mat_list[s]; rows[s] = m.nrow(); D = m.ncol(); flat.insert(flat.end(), m.begin(), m.end()); }
  std::vector<int32_t> pi(P), pj(P), ii(idx_i.size()), ij(idx_j.size());
  for (R_xlen_t p = 0; p < P; ++p) { pi[p] = pair_i[p] - 1; pj[p] = pair_j[p] - 1; }
  for (R_xlen_t e = 0; e < idx_i.size(); ++e) ii[e] = idx_i[e] - 1;
  for (R_xlen_t e = 0; e < idx_j.size(); ++e) ij[e] = idx_j[e] - 1;
  sqc_mmd_args a = {};
  a.abi_version = SQC_ABI_VERSION; a.D = D; a.n_samples = S; a.mats = flat.data(); a.mat_rows = rows.data();
  a.n_pairs = P; a.pair_i = pi.data(); a.pair_j = pj.data(); a.n_times = n_times; a.subsample = subsample;
  a.idx_i = ii.empty() ? nullptr : ii.data(); a.idx_j = ij.empty() ? nullptr : ij.data(); a.sigma = sigma; a.device = -1;
  NumericVector mean(P);
  char err[512];
  check_status(sqc_mmd_pairs(&a, mean.begin(), nullptr, err, sizeof err), err);
  return mean;
}
