context("CUDA drop-in for fit_sampleqc_robust_cpp / fit_sampleqc_mle_cpp")
# For the maintainer, run with the CUDA build of the package (rpkg/src/fit_sampleQC_cuda_stubs.cpp in place of
# src/fit_sampleQC_robust_cpp.cpp + src/fit_sampleQC_mle_cpp.cpp, INTEGRATION.md).  NOT run in this repository's CI: the
# image has no R; the same comparisons run there through the C ABI against the CPU restatement (tests/test_gpu_*.py).
#   SAMPLEQC_REF_RDS=reference_fit.rds  -> compare with the reference build's outputs (rpkg/tools/make_reference_fit.R)

ref_f = Sys.getenv("SAMPLEQC_REF_RDS", "")

small_case = function() {
  set.seed(99)
  sims    = simulate_qcs(n_cells = 5e3)
  qc_dt   = make_qc_dt(sims$qc_out, sample_var = 'sample_id', qc_names = sims$expt_params$qc_names)
  x       = as.matrix(qc_dt[, sims$expt_params$qc_names, with = FALSE])
  groups  = as.integer(factor(qc_dt$sample_id))
  D = ncol(x); J = max(groups); K = 2L; N = nrow(x)
  list(x = x, groups = groups, D = D, J = J, K = K, N = N, init = SampleQC:::.init_clusts(x, groups, J, K, N))
}

test_that("the robust export leaves R's RNG where the reference does", {
  # the reference draws one unif_rand per cell per update_beta_scale_k call (arma::randperm, robust_cpp.cpp:72) after
  # set_seed_cpp(seed) (:475): (1 + n_iter) * N draws.  The stub restores that state (restore_r_rng + GetRNGstate).
  cs    = small_case(); seed = 11L
  fit   = with(cs, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 20L, 0.5, 50L, seed))
  u_fit = runif(1)
  n_iter = sum(as.vector(fit$like_2) != 0)                               # the list holds no iteration count; like_* are zero-padded (:562-575)
  set.seed(seed); invisible(runif((1 + n_iter) * cs$N)); u_ref = runif(1)
  expect_identical(u_fit, u_ref)
})

test_that("seeds are replicable, bit for bit", {                        # reference: test-03_fit_sampleqc.R:81-88
  cs    = small_case()
  f1    = with(cs, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 20L, 0.5, 50L, 5L))
  f2    = with(cs, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 20L, 0.5, 50L, 5L))
  expect_identical(f1, f2)
})

test_that("the reference's failure classes come back as R errors", {
  cs    = small_case()
  expect_error(with(cs, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 20L, 1.5, 50L, 1L)),
    "mcd_alpha")                                                          # robust_cpp.cpp:471
  z_bad = cs$init$init_z; z_bad[] = 0L                                   # component 1 empty: x_k.rows(0, -1), :307
  expect_error(with(cs, SampleQC:::fit_sampleqc_robust_cpp(x, z_bad, groups - 1L, D, J, K, N, 20L, 0.5, 50L, 1L)))
})

test_that("outputs equal the reference build's", {
  skip_if(ref_f == "", "SAMPLEQC_REF_RDS not set (make it with rpkg/tools/make_reference_fit.R under the reference build)")
  r     = readRDS(ref_f)
  set.seed(r$seed)
  fit   = with(r, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 50L, 0.5, 50L, seed))
  u     = runif(1)
  expect_identical(names(fit), names(r$robust))                         # List::create names, robust_cpp.cpp:562-575
  expect_identical(sum(as.vector(fit$like_2) != 0), sum(as.vector(r$robust$like_2) != 0))   # same number of EM iterations
  expect_identical(fit$z, r$robust$z)                                   # labels: bit-identical
  for (nn in c("mu_0", "alpha_j", "beta_k", "sigma_k", "p_jk", "like_1", "like_2"))
    expect_equal(fit[[nn]], r$robust[[nn]], tolerance = 1e-8, info = nn)
  expect_identical(u, r$u_after_robust)
  set.seed(r$seed)
  fit_m = with(r, SampleQC:::fit_sampleqc_mle_cpp(x, init$init_gamma, groups - 1L, D, J, K, N, 20L, seed))
  expect_identical(names(fit_m), names(r$mle))                          # mle_cpp.cpp:341-356
  for (nn in c("mu_0", "alpha_j", "beta_k", "sigma_k", "p_k", "gamma_i", "like_1", "like_2"))
    expect_equal(fit_m[[nn]], r$mle[[nn]], tolerance = 1e-8, info = nn)
  expect_identical(fit_m$z, r$mle$z)
})

test_that("fit_sampleqc() is unchanged above the exports, on one GPU and on all of them", {
  skip_if(ref_f == "", "SAMPLEQC_REF_RDS not set")
  # R/fit_sampleqc.R:229-232 calls the export once per sample group: options(SampleQC.n_gpus = n) shards that one call
  r     = readRDS(ref_f)
  fit_1 = with(r, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 50L, 0.5, 50L, seed))
  old   = options(SampleQC.n_gpus = -1L); on.exit(options(old))
  fit_n = with(r, SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 50L, 0.5, 50L, seed))
  expect_identical(fit_n$z, fit_1$z)
  expect_equal(fit_n$beta_k, fit_1$beta_k, tolerance = 1e-10)
})
