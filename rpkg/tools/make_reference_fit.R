# Run ONCE with the unmodified reference build of SampleQC installed: writes the inputs and the reference's outputs of
# the two exports the CUDA stubs replace (src/fit_sampleQC_robust_cpp.cpp:451-577, src/fit_sampleQC_mle_cpp.cpp:279-357)
# to an RDS that tests/testthat/test-06_cuda_dropin.R (run with the CUDA build) compares against.
#   Rscript rpkg/tools/make_reference_fit.R reference_fit.rds
# (Not run in this repository's CI: the image has no R.  Written against the reference's internal API at the cited lines.)
suppressPackageStartupMessages(library(SampleQC))
out_f   = commandArgs(trailingOnly = TRUE)[1]
set.seed(20240101)
sims    = simulate_qcs(n_cells = 2e4)                                  # R/simulate_qcs.R:47
qc_dt   = make_qc_dt(sims$qc_out, sample_var = 'sample_id', qc_names = sims$expt_params$qc_names)
x       = as.matrix(qc_dt[, sims$expt_params$qc_names, with = FALSE])
groups  = as.integer(factor(qc_dt$sample_id))                          # R/fit_sampleqc.R:205-216
D = ncol(x); J = max(groups); K = 2L; N = nrow(x)
init    = SampleQC:::.init_clusts(x, groups, J, K, N)                  # R/fit_sampleqc.R:277-331
seed    = 4321L
set.seed(seed)                                                         # the export re-seeds itself (:475); this pins u_after
fit_r   = SampleQC:::fit_sampleqc_robust_cpp(x, init$init_z, groups - 1L, D, J, K, N, 50L, 0.5, 50L, seed)
u_r     = runif(1)                                                     # R's stream after the fit: (1 + n_iter) * N draws on
set.seed(seed)
fit_m   = SampleQC:::fit_sampleqc_mle_cpp(x, init$init_gamma, groups - 1L, D, J, K, N, 20L, seed)
u_m     = runif(1)
saveRDS(list(x = x, groups = groups, D = D, J = J, K = K, N = N, init = init, seed = seed,
  robust = fit_r, u_after_robust = u_r, mle = fit_m, u_after_mle = u_m), out_f)
