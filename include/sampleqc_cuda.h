/*
 * sampleqc_cuda.h — C ABI of the B200-native SampleQC EM library (libsampleqc_cuda.so).
 *
 * This is the drop-in boundary for ONE path of wmacnair/SampleQC: the multi-sample
 * Gaussian-mixture EM behind the Rcpp exports
 *
 *   fit_sampleqc_robust_cpp(x, init_z, groups, D, J, K, N, em_iters, mcd_alpha,
 *                           mcd_iters, seed, track)     reference src/fit_sampleQC_robust_cpp.cpp:451
 *   fit_sampleqc_mle_cpp(x, init_gamma_i, groups, D, J, K, N, n_iter, seed)
 *                                                       reference src/fit_sampleQC_mle_cpp.cpp:279
 *
 * and the post-fit outlier arithmetic of .calc_outliers_dt (reference R/fit_sampleqc.R:419-453).
 * The Rcpp bodies of those two exports become marshalling stubs that call the entry points
 * below (see INTEGRATION.md and rpkg/src/); the generated src/RcppExports.cpp:16-54 and its
 * registration table :229-252 stay untouched.
 *
 * Conventions
 *   - plain C, no CUDA / torch / Rcpp types in any signature;
 *   - every matrix is COLUMN-MAJOR exactly as R / Armadillo hold it, so the Rcpp stubs can hand
 *     REAL() pointers straight through:  x is N x D,  alpha_j J x D,  beta_k K x D,
 *     sigma_k D x D x K,  p_jk J x K,  gamma_i N x K,  track_* as the reference cubes;
 *   - the caller owns every host pointer (inputs are read-only, outputs caller-allocated);
 *     the library owns all device memory and frees it before the call returns (or at
 *     sqc_session_destroy);
 *   - return value: 0 = ok; >0 = a failure class the reference would also raise as an R error
 *     (bad mcd_alpha, empty cluster, non-SPD scatter, NaN ...); <0 = CUDA / communicator
 *     failure.  A NUL-terminated message is written to err (if err != NULL);
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point returns
 *     SQC_ERR_CUDA.
 */
#ifndef SAMPLEQC_CUDA_H
#define SAMPLEQC_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQC_ABI_VERSION 2

#if defined(__GNUC__)
#define SQC_API __attribute__((visibility("default")))
#else
#define SQC_API
#endif

/* ---- status codes --------------------------------------------------------------------- */
enum {
  SQC_OK                 = 0,
  /* reference-equivalent failure classes (the reference throws / Armadillo throws) */
  SQC_ERR_BAD_ARG        = 1,  /* e.g. mcd_alpha outside [0,1): robust_cpp.cpp:471-472        */
  SQC_ERR_EMPTY_CLUSTER  = 2,  /* n_k == 0: x_k.rows(0,-1), robust_cpp.cpp:307                */
  SQC_ERR_SUBSET_SIZE    = 3,  /* h_k < 1 or h_k > n_k: randperm / span failure, :72, :51     */
  SQC_ERR_NOT_SPD        = 4,  /* chol() of a non-SPD scatter, robust_cpp.cpp:19              */
  SQC_ERR_SINGULAR       = 5,  /* inv() of a singular matrix, :127, :153 / mle_cpp.cpp:39,60  */
  SQC_ERR_NAN            = 6,  /* NaN distance reaching sort_index, :50                       */
  SQC_ERR_SELECT         = 7,  /* exact selection could not be resolved (more exactly-tied
                                  distances than the candidate store holds)                   */
  /* infrastructure failures */
  SQC_ERR_CUDA           = -1,
  SQC_ERR_COMM           = -2,
  SQC_ERR_NOMEM          = -3
};

/* rng_mode: how the random start of every FastMCD (arma::randperm, robust_cpp.cpp:72) is keyed */
enum {
  SQC_RNG_R_COMPAT = 0, /* R's Mersenne-Twister stream after set.seed(seed) (helpers.cpp:10-14),
                           one unif_rand per cell per update_beta_scale_k call, consumed in
                           cluster-then-cell order; key = (int) Rf_runif(0, RAND_MAX)          */
  SQC_RNG_COUNTER  = 1  /* counter-based keys hashed from (seed, position in that same stream);
                           random access, not stream-compatible with R, identical in oracle
                           and CUDA                                                            */
};

/* ---- fit arguments -------------------------------------------------------------------- */
typedef struct sqc_fit_args {
  uint32_t abi_version;     /* SQC_ABI_VERSION                                               */
  int32_t  D, J, K;         /* metrics, samples, mixture components                          */
  int64_t  N;               /* cells                                                         */
  const double*  x;         /* N x D column-major, un-centred (R matrix)                     */
  const int32_t* groups;    /* N, 0-based sample index of every cell                         */
  const int32_t* init_z;    /* robust: N, 0-based initial labels (R sends doubles; the Rcpp
                               stub converts, as arma::uvec conversion does today)           */
  const double*  init_gamma;/* mle: N x K column-major initial responsibilities              */
  int32_t  em_iters;        /* robust: max EM iterations; mle: n_iter (fixed count)          */
  int32_t  mcd_iters;       /* robust: max C-steps per FastMCD                               */
  double   mcd_alpha;       /* robust: subset fraction, h_k = int((n_k + D + 1) * mcd_alpha) */
  uint32_t seed;            /* as received by the export          */
  int32_t  track;           /* robust: fill track_* outputs                                  */
  int32_t  rng_mode;        /* SQC_RNG_*                                                     */
  int32_t  device;          /* CUDA device ordinal; -1 = current device                      */
  /* multi-GPU (cells sharded by sample, one process per GPU).  world <= 1 means single GPU. */
  int32_t  rank, world;
  int64_t  cell_offset;     /* index of this shard's first cell in the full data set         */
  int64_t  N_total;         /* cells over all shards (== N when world <= 1)                  */
  const void* peer_handles; /* world x SQC_IPC_HANDLE_BYTES: sqc_xchg_export() of every rank */
  /* multi-GPU from ONE call (what the Rcpp stubs use: fit_sampleqc() makes a single call per sample group,
   * R/fit_sampleqc.R:229-238).  Only read when world <= 1.  0 or 1: one GPU (`device`); n > 1: the library shards the
   * cells by sample over GPUs 0 .. n-1 of this process (one host thread per GPU, peer access instead of IPC handles);
   * SQC_ALL_GPUS: every visible GPU (at most 8).  Fewer GPUs are used when there are fewer samples than GPUs; a data
   * set whose samples are interleaved cell by cell is fitted on one GPU.  Results are those of the one-GPU fit
   * (sharded fit == oracle, tests/test_gpu_multi.py).  Calls with n_gpus > 1 from different host threads take
   * turns (a GPU has one exchange window per process); single-GPU calls run concurrently.                    */
  int32_t  n_gpus;
} sqc_fit_args;
#define SQC_ALL_GPUS (-1)

/* ---- fit outputs (all caller-allocated; NULL pointers are skipped) --------------------- */
typedef struct sqc_fit_out {
  double*  mu_0;            /* D                                                             */
  double*  alpha_j;         /* J x D                                                         */
  double*  beta_k;          /* K x D, rows ordered by ascending beta_k[,0] (:527-531)        */
  double*  sigma_k;         /* D x D x K                                                     */
  int32_t* z;               /* N, 1-based labels after the final relabel (the list's z)      */
  double*  p_jk;            /* J x K                                                         */
  double*  like_1;          /* em_iters, zero padded                                         */
  double*  like_2;          /* em_iters, zero padded                                         */
  double*  gamma_i;         /* mle: N x K (columns NOT reordered, as in mle_cpp.cpp:350)     */
  double*  p_k;             /* mle: K                                                        */
  double*  track_alpha_j;   /* robust+track: J x D x em_iters                                */
  double*  track_beta_k;    /* robust+track: K x D x em_iters                                */
  double*  track_sigma_k;   /* robust+track: D*D x K x em_iters                              */
  double*  track_p_jk;      /* robust+track: J x K x em_iters                                */
  int32_t  n_iter;          /* EM iterations executed ("took %d iterations", :524)           */
  int32_t  mcd_iters_hit;   /* times fast_mcd reached mcd_iters (:94-95)                     */
  int32_t  n_zero_like;     /* the reference's n_zeros counter (:380-382, :392-393)          */
  int64_t  rng_draws;       /* unif_rand draws consumed = (1 + n_iter) * N_total (robust)    */
  int32_t  rng_state[625];  /* R's .Random.seed payload (mti, mt[624]) after those draws,
                               R_COMPAT mode only; lets the Rcpp stub leave R's RNG where the
                               reference would have left it                                  */
} sqc_fit_out;

/* One-shot fits with HOST buffers: upload, run the whole EM on the device, download.
 * These are what the Rcpp stubs call.  replaces robust_cpp.cpp:451-577 / mle_cpp.cpp:279-357 */
SQC_API int sqc_fit_robust(const sqc_fit_args* args, sqc_fit_out* out, char* err, size_t errlen);
SQC_API int sqc_fit_mle   (const sqc_fit_args* args, sqc_fit_out* out, char* err, size_t errlen);

/* ---- post-fit outlier arithmetic (.calc_outliers_dt, R/fit_sampleqc.R:419-453) ---------- */
typedef struct sqc_maha_args {
  uint32_t abi_version;
  int32_t  D, J, K;
  int64_t  N;
  const double*  x;         /* N x D column-major, un-centred                                */
  const int32_t* groups;    /* N, 0-based                                                    */
  const double*  mu_0;      /* D       (already re-centred as in R/fit_sampleqc.R:242-246)   */
  const double*  alpha_j;   /* J x D                                                         */
  const double*  beta_k;    /* K x D                                                         */
  const double*  sigma_k;   /* D x D x K                                                     */
  double   alpha;           /* outlier level; cut = qchisq(1 - alpha, D)                     */
  int32_t  device;
} sqc_maha_args;

/* maha: N x K column-major (may be NULL), outlier: N bytes 0/1, maha_cut: the chi-square cut */
SQC_API int sqc_outlier_maha(const sqc_maha_args* args, double* maha, uint8_t* outlier,
                     double* maha_cut, char* err, size_t errlen);

/* ---- the step in front of the fit: .init_clusts (R/fit_sampleqc.R:277-331) --------------- */
/* grp_medians (:286-289) of every sample and metric, and - when a mixture is given - the posterior and arg max of ALL
 * N median-centred cells under it (what summaryMclustBIC(...)$z / $classification deliver, :303-316).  The fit of that
 * mixture on the 1e4-cell subsample (mclust) stays with the caller.                                                  */
typedef struct sqc_init_args {
  uint32_t abi_version;
  int32_t  D, J, K;          /* K is ignored when pro == NULL                                 */
  int64_t  N;
  const double*  x;          /* N x D column-major, un-centred                                */
  const int32_t* groups;     /* N, 0-based, any order                                         */
  const double*  pro;        /* K mixing proportions, or NULL: medians only                   */
  const double*  mean;       /* K x D column-major                                            */
  const double*  sigma;      /* D x D x K                                                     */
  int32_t  device;
} sqc_init_args;

/* medians: J x D column-major (NaN for an empty sample); gamma: N x K column-major or NULL; z: N labels, 0-based
 * (the first maximum of a row of gamma), or NULL                                                                    */
SQC_API int sqc_init_clusts(const sqc_init_args* args, double* medians, double* gamma, int32_t* z,
                    char* err, size_t errlen);

/* ---- the resident pipeline: .init_clusts -> fit -> re-centring -> outliers on ONE upload of the cells ------------- */
/* .fit_one_sampleqc (R/fit_sampleqc.R:195-262) runs .init_clusts (:222), fit_sampleqc_*_cpp (:229-238), the re-centring
 * (:242-246) and .calc_outliers_dt (:258) on the same x / groups.  A sqc_resident keeps both in HBM: x crosses PCIe once,
 * the labels of the classification go into the fit without leaving the device, the fitted parameters come back already
 * re-centred and stay with the handle for the outlier pass.  One GPU; the cells of every sample must be contiguous.    */
typedef struct sqc_resident sqc_resident;
SQC_API int  sqc_resident_create(const double* x /* N x D column-major, un-centred */, const int32_t* groups /* N, 0-based */,
                                 int32_t D, int32_t J, int64_t N, int32_t device, sqc_resident** out, char* err, size_t errlen);
/* grp_medians (:286-289) -> medians (J x D, may be NULL); with pro != NULL also the classification of all cells under the
 * mixture fitted on the subsample (:303-316).  gamma (N x K) / z (N, 0-based) may be NULL: labels - and with keep_gamma
 * the responsibilities - stay on the device as the initial values of the next sqc_resident_fit.                       */
SQC_API int  sqc_resident_init_clusts(sqc_resident* r, const double* pro, const double* mean /* K x D */, const double* sigma /* D x D x K */,
                                      int32_t K, int32_t keep_gamma, double* medians, double* gamma, int32_t* z, char* err, size_t errlen);
/* rows idx[0 .. n) of x minus their samples' medians, n x D column-major: the subsample the caller's mixture is fitted on (:291-300) */
SQC_API int  sqc_resident_subsample(sqc_resident* r, const int64_t* idx, int64_t n, double* out, char* err, size_t errlen);
/* the fit on the resident cells.  args->x / groups / D / J / N are ignored; init_z / init_gamma NULL = the resident
 * classification.  recentre != 0: mu_0, alpha_j, beta_k come back re-centred as R/fit_sampleqc.R:242-246 does.       */
SQC_API int  sqc_resident_fit(sqc_resident* r, const sqc_fit_args* args, int variant, int32_t recentre, sqc_fit_out* out, char* err, size_t errlen);
/* .calc_outliers_dt (:419-453) with the (re-centred) parameters of the last sqc_resident_fit; maha (N x K) may be NULL */
SQC_API int  sqc_resident_outliers(sqc_resident* r, double alpha, double* maha, uint8_t* outlier, double* maha_cut, char* err, size_t errlen);
SQC_API void sqc_resident_destroy(sqc_resident* r);

/* ---- sample-sample maximum mean discrepancy --------------------------------------------- */
/* mmd_fn(X, Y, sigma), reference src/mmd_fn.cpp:20-91: X n_x x D, Y n_y x D column-major; *mmd = the signed value.
 * sigma <= 0 -> SQC_ERR_BAD_ARG ("sigma must be > 0", :28-30).                                                        */
SQC_API int sqc_mmd_fn(const double* X, int64_t n_x, const double* Y, int64_t n_y, int32_t D, double sigma, double* mmd,
                       int32_t device, char* err, size_t errlen);
/* .calc_mmd_mat (reference R/calc_pairwise_mmds.R:198-247) in one call: every (i, j) pair x n_times repetitions of
 * abs(mmd_fn(subsample of sample i, subsample of sample j)) (.dist_fn, :265-280) and their mean per pair.  The row picks
 * are the caller's (R's sample() draws them, so R's RNG stream stays the caller's): for pair p, repetition t, the picks of
 * side i are idx_i[(p * n_times + t) * subsample ..], 0-based; a sample with <= subsample rows is used whole and its
 * picks are ignored (:271, :276).                                                                                    */
typedef struct sqc_mmd_args {
  uint32_t abi_version;
  int32_t  D, n_samples;
  const double*  mats;       /* the samples' matrices one after the other, each mat_rows[s] x D column-major           */
  const int64_t* mat_rows;   /* n_samples                                                                              */
  int64_t  n_pairs;
  const int32_t* pair_i;     /* n_pairs, 0-based sample indices                                                        */
  const int32_t* pair_j;
  int32_t  n_times, subsample;
  const int32_t* idx_i;      /* n_pairs x n_times x subsample (may be NULL when no sample exceeds subsample rows)      */
  const int32_t* idx_j;
  double   sigma;
  int32_t  device;
} sqc_mmd_args;
SQC_API int sqc_mmd_pairs(const sqc_mmd_args* args, double* mmd_mean /* n_pairs */, double* mmd_all /* n_pairs x n_times or NULL */,
                          char* err, size_t errlen);

/* ---- resident sessions (benchmarks, multi-GPU drivers) --------------------------------- */
/* A session keeps the cell matrix resident in HBM so that the EM loop can be timed without the
 * host<->device copies, and exposes per-kernel device timings for the roofline report.       */
typedef struct sqc_session sqc_session;

enum { SQC_VARIANT_ROBUST = 0, SQC_VARIANT_MLE = 1 };

/* per-kernel-family accounting of the last sqc_session_run (device times from CUDA events)   */
#define SQC_N_KERNEL_FAMILIES 8
enum {
  SQC_KF_ESTEP     = 0,  /* E-step: densities, argmax / gamma, likelihood, (j,k) statistics   */
  SQC_KF_PARTITION = 1,  /* residual partition by cluster + like_1                            */
  SQC_KF_MCD       = 2,  /* M-step: FastMCD selection + masked moments (persistent kernel)    */
  SQC_KF_UPDATE    = 3,  /* small parameter-update kernels                                    */
  SQC_KF_RNG       = 4,  /* Mersenne-Twister key stream                                       */
  SQC_KF_INIT      = 5,  /* column means, centring, initial statistics                        */
  SQC_KF_MLE_PASS  = 6,  /* MLE fused E/M passes                                              */
  SQC_KF_FINAL     = 7   /* relabel / gather                                                  */
};

typedef struct sqc_run_stats {
  int32_t n_iter;                              /* EM iterations executed                      */
  int32_t n_gpu_launches;                      /* kernels launched by the run                 */
  double  loop_ms;                             /* device time of the EM loop (events)         */
  double  init_ms;                             /* device time of centring / initial M-step    */
  double  family_ms[SQC_N_KERNEL_FAMILIES];    /* summed device time per family               */
  int64_t family_launches[SQC_N_KERNEL_FAMILIES];
  double  family_bytes[SQC_N_KERNEL_FAMILIES]; /* algorithmic bytes moved (DESIGN.md model)   */
  int64_t mcd_passes;                          /* streaming passes executed inside MCD        */
  int64_t mcd_csteps;                          /* lock-step C-steps executed                  */
} sqc_run_stats;

SQC_API int  sqc_session_create (const sqc_fit_args* args, int variant, sqc_session** out,
                         char* err, size_t errlen);      /* allocates + uploads (H2D)          */
SQC_API int  sqc_session_run    (sqc_session* s, char* err, size_t errlen); /* re-runnable: resets state,
                                                             runs init + EM loop on device     */
SQC_API int  sqc_session_fetch  (sqc_session* s, sqc_fit_out* out, char* err, size_t errlen); /* D2H   */
SQC_API int  sqc_session_stats  (const sqc_session* s, sqc_run_stats* stats);
SQC_API int  sqc_session_profile(sqc_session* s, int on);         /* per-launch CUDA-event timing on/off */
SQC_API void sqc_session_destroy(sqc_session* s);

/* ---- multi-GPU exchange bootstrap ------------------------------------------------------ */
/* Each rank allocates its exchange window and exports an IPC handle; the host side gathers all
 * handles (torch.distributed / MPI / files - plumbing) and passes them in sqc_fit_args.        */
#define SQC_IPC_HANDLE_BYTES 64
SQC_API int sqc_xchg_export(int device, void* handle_out /* SQC_IPC_HANDLE_BYTES */, char* err, size_t errlen);
SQC_API int sqc_xchg_release(int device);

/* ---- misc ------------------------------------------------------------------------------ */
/* Finished fits keep their device memory in a per-device cache (allocation costs tens of ms); this
 * releases it.  device < 0: all devices.                                                       */
SQC_API int         sqc_trim(int device);
SQC_API int         sqc_abi_version(void);
SQC_API int         sqc_device_count(void);          /* <0 on CUDA failure                            */
SQC_API const char* sqc_status_name(int status);
/* device-side special functions exposed for tests: qchisq(p, df) evaluated by the same device
 * routine the MCD consistency factor uses (replaces R::qchisq, robust_cpp.cpp:102)            */
SQC_API int sqc_device_qchisq(const double* p, const double* df, double* out, int32_t n, int device,
                      char* err, size_t errlen);
/* the kernels' own exp(x), x <= 0 (density exponents), exposed for accuracy tests               */
SQC_API int sqc_device_expneg(const double* x, double* out, int32_t n, int device, char* err, size_t errlen);
/* the R-compatible Mersenne-Twister key stream generated on the device (tests)               */
SQC_API int sqc_device_rkeys(uint32_t seed, int64_t skip, int32_t* keys_out, int64_t n, int device,
                     char* err, size_t errlen);
/* the same stream as ONE RANK of a multi-GPU fit produces it (tests; runs on one GPU): rank `rank` of `world` generates
 * draws [rank Q, (rank + 1) Q) of each of `calls` consecutive calls of n_total draws (Q = *q_out, a multiple of 624);
 * keys_out holds calls x Q ints, state_out (625 words, may be NULL) R's state after the last call.                     */
SQC_API int sqc_device_rkeys_share(uint32_t seed, int64_t n_total, int world, int rank, int calls, int32_t* keys_out,
                           int64_t* q_out, int32_t* state_out, int device, char* err, size_t errlen);

/* host-only (no GPU needed): the sample-aligned shards a fit with n_gpus > 1 is cut into.  shard_cells: n_gpus + 1 cell
 * boundaries; sample_rank (J, may be NULL): shard that owns every sample code, -1 for a sample without cells.  Returns
 * the number of shards (<= n_gpus; fewer when there are fewer samples), 0 when cells of a sample are interleaved (the
 * fit then runs on one GPU), < 0 on bad arguments                                                                    */
SQC_API int sqc_plan_shards(const int32_t* groups, int64_t N, int32_t J, int32_t n_gpus, int64_t* shard_cells, int32_t* sample_rank);

/* measured fp64 FMA issue rate of the device in FMA instructions per second (x 2 = FLOP/s): the co-limit of the E-step /
 * partition passes next to HBM bandwidth, reported by bench.py beside the HBM roofline          */
SQC_API int sqc_fp64_peak(int device, double* dfma_per_s, char* err, size_t errlen);

/* host-only (no GPU needed): x^e mod phi_MT19937 over GF(2), the jump polynomial that moves the Mersenne-Twister state
 * e words ahead (624 words, bit i of word w = coefficient of x^(32 w + i)); SQC_ERR_COMM if the self-check against the
 * offline-verified table fails                                                                                       */
SQC_API int sqc_mt_jump_poly(uint64_t e, uint32_t* out624);

#ifdef __cplusplus
}
#endif
#endif /* SAMPLEQC_CUDA_H */
