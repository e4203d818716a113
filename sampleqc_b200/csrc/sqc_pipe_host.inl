// sqc_pipe_host.inl — the resident pipeline (included by sqc_lib.cu inside extern "C").
//
// .fit_one_sampleqc (reference R/fit_sampleqc.R:195-262) runs  .init_clusts (:222)  ->  fit_sampleqc_*_cpp (:229-238)  ->
// re-centring (:242-246)  ->  .calc_outliers_dt (:258)  on the same x / groups.  With three independent entry points the
// same N x D matrix crosses PCIe three times.  A sqc_resident keeps x and groups in HBM: the cells are uploaded ONCE,
// every step reads them there, the labels of the classification never leave the device on their way into the fit,
// the fitted parameters come back already re-centred and stay available for the outlier pass.
struct sqc_resident {
  int device = 0, D = 0, J = 0; long long N = 0;
  double* x = nullptr; int* groups = nullptr;                 // device: raw x (N x D column-major), sample index per cell
  double* med = nullptr; bool have_med = false;               // device: J x D medians
  int* z = nullptr; double* gamma = nullptr; int K_cls = 0, K_gam = 0, K_gcls = 0;   // device: classification of all cells under the caller's mixture (K_gcls: K the resident gamma belongs to, 0 = none)
  std::vector<int32_t> h_groups; std::vector<long long> bounds;   // bounds[j + 1] = cells of sample j
  std::vector<int> run_order;                                 // sample codes in order of appearance
  bool runs_ok = true;                                        // every sample is one contiguous run of cells
  // parameters of the last fit, re-centred (host): what .calc_outliers_dt receives
  int K_fit = 0; bool recentred = false; std::vector<double> mu0, alpha, beta, sigma;
  cudaStream_t st = nullptr;
  std::vector<Chunk> chunks;                                  // from the per-device arena the fits use (cudaMalloc of 100s of MB costs tens of ms)
  void* take(size_t bytes) { chunks.push_back(g_chunks.take(device, std::max<size_t>(bytes, 256))); return chunks.back().base; }
  ~sqc_resident() {
    if (device >= 0) cudaSetDevice(device);
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto& c : chunks) g_chunks.give(device, c);
  }
};

int sqc_resident_create(const double* x, const int32_t* groups, int32_t D, int32_t J, int64_t N, int32_t device, sqc_resident** out, char* err, size_t errlen) {
  if (out) *out = nullptr;
  return guarded([&] {
    if (!x || !groups || !out) fail(SQC_ERR_BAD_ARG, "null input");
    if (D < 1 || D > SQC_MAXD || J < 1 || N < 1 || N > 2000000000ll) fail(SQC_ERR_BAD_ARG, "bad dimensions");
    std::unique_ptr<sqc_resident> R(new sqc_resident());
    select_device(device, &R->device);
    R->D = D; R->J = J; R->N = N;
    CK(cudaStreamCreateWithFlags(&R->st, cudaStreamNonBlocking));
    R->x = (double*)R->take((size_t)N * D * sizeof(double)); R->groups = (int*)R->take((size_t)N * sizeof(int)); R->med = (double*)R->take((size_t)J * D * sizeof(double));
    // the one upload of the cells (asynchronous when the caller's memory is pinned); the host scans the groups meanwhile
    CK(cudaMemcpyAsync(R->x, x, (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice, R->st));
    CK(cudaMemcpyAsync(R->groups, groups, (size_t)N * sizeof(int), cudaMemcpyHostToDevice, R->st));
    // host copy + scan of the groups on a few threads (it hides behind the upload): range check, sample sizes, and
    // "every sample is one contiguous run" (chunk-local run lists, merged in order)
    R->h_groups.resize((size_t)N);
    R->bounds.assign((size_t)J + 1, 0);
    {
      const int nth = (N >= (1 << 20)) ? 8 : 1;
      std::vector<std::vector<long long>> cnt(nth, std::vector<long long>(J, 0));
      std::vector<std::vector<int>> runs(nth);
      std::vector<long long> bad(nth, -1); std::vector<char> ovf(nth, 0);
      auto scan = [&](int t) {
        const long long i0 = N * t / nth, i1 = N * (t + 1) / nth;
        memcpy(R->h_groups.data() + i0, groups + i0, (size_t)(i1 - i0) * sizeof(int32_t));
        int prev = -1;
        for (long long i = i0; i < i1; ++i) {
          const int g = groups[i];
          if ((unsigned)g >= (unsigned)J) { bad[t] = i; return; }
          cnt[t][g]++;
          if (g != prev) { if ((long long)runs[t].size() <= J) runs[t].push_back(g); else ovf[t] = 1; prev = g; }
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nth; ++t) th.emplace_back(scan, t);
      scan(0);
      for (auto& x : th) x.join();
      std::vector<char> seen(J, 0);
      int prev = -1;
      for (int t = 0; t < nth; ++t) {
        if (bad[t] >= 0) { cudaStreamSynchronize(R->st); fail(SQC_ERR_BAD_ARG, "groups out of range at cell %lld", bad[t]); }
        if (ovf[t]) R->runs_ok = false;
        for (int j = 0; j < J; ++j) R->bounds[(size_t)j + 1] += cnt[t][j];
        for (int g : runs[t]) { if (g == prev) continue; if (seen[g]) R->runs_ok = false; seen[g] = 1; prev = g; R->run_order.push_back(g); }
      }
    }
    if (!R->runs_ok) { cudaStreamSynchronize(R->st); fail(SQC_ERR_BAD_ARG, "the resident pipeline needs the cells of every sample to be contiguous (as fit_sampleqc() builds x, R/fit_sampleqc.R:205-216)"); }
    CK(cudaStreamSynchronize(R->st));
    *out = R.release();
  }, err, errlen);
}
void sqc_resident_destroy(sqc_resident* r) { delete r; }

namespace {
// start of every sample's run of cells (samples appear in any code order, each in one run)
void resident_run_bounds(const sqc_resident* R, std::vector<long long>& b0, std::vector<long long>& b1) {
  b0.assign(R->J, 0); b1.assign(R->J, 0);
  long long pos = 0;
  for (int g : R->run_order) { b0[g] = pos; pos += R->bounds[(size_t)g + 1]; b1[g] = pos; }
}
void resident_medians(sqc_resident* R) {
  if (R->have_med) return;
  // k_sample_median takes segment bounds b[j] .. b[j + 1]: give it every sample's own run through a 2 J table
  std::vector<long long> b0, b1; resident_run_bounds(R, b0, b1);
  std::vector<long long> tab(2 * (size_t)R->J);
  for (int j = 0; j < R->J; ++j) { tab[2 * (size_t)j] = b0[j]; tab[2 * (size_t)j + 1] = b1[j]; }
  long long* db = nullptr; int* dflag = nullptr;
  auto cleanup = [&] { cudaFree(db); cudaFree(dflag); };
  try {
    CK(cudaMalloc((void**)&db, tab.size() * sizeof(long long))); CK(cudaMalloc((void**)&dflag, sizeof(int)));
    CK(cudaMemsetAsync(dflag, 0, sizeof(int), R->st));
    CK(cudaMemcpyAsync(db, tab.data(), tab.size() * sizeof(long long), cudaMemcpyHostToDevice, R->st));
    k_sample_median_runs<<<R->J * R->D, MED_THREADS, 0, R->st>>>(R->x, db, R->N, R->J, R->med, dflag);
    CK(cudaGetLastError());
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, R->st));
    CK(cudaStreamSynchronize(R->st));
    if (flag) fail(SQC_ERR_NAN, "NaN in x");
  } catch (...) { cleanup(); throw; }
  cleanup();
  R->have_med = true;
}
}  // namespace

// grp_medians (:286-289) and, with a mixture, the classification of all cells (:303-316).  gamma / z may be NULL: the
// labels (and responsibilities) stay on the device as the initial values of the next sqc_resident_fit.
int sqc_resident_init_clusts(sqc_resident* R, const double* pro, const double* mean, const double* sigma, int32_t K, int32_t keep_gamma,
                             double* medians, double* gamma, int32_t* z, char* err, size_t errlen) {
  return guarded([&] {
    if (!R) fail(SQC_ERR_BAD_ARG, "null handle");
    CK(cudaSetDevice(R->device));
    resident_medians(R);
    const int D = R->D, J = R->J; const long long N = R->N;
    if (medians) CK(cudaMemcpyAsync(medians, R->med, (size_t)J * D * sizeof(double), cudaMemcpyDeviceToHost, R->st));
    if (pro) {
      if (K < 1 || K > SQC_MAXK || !mean || !sigma) fail(SQC_ERR_BAD_ARG, "bad mixture");
      const int P = D + D * D + 1;
      std::vector<double> par((size_t)K * P, 0.0);
      for (int k = 0; k < K; ++k) {
        double L[SQC_MAXD * SQC_MAXD], Li[SQC_MAXD * SQC_MAXD];
        if (!small_chol_inv(sigma + (size_t)k * D * D, D, L, Li)) fail(SQC_ERR_NOT_SPD, "mixture covariance %d is not positive definite", k);
        double c = log(pro[k]) - 0.5 * D * 1.8378770664093454835606594728112;
        for (int d = 0; d < D; ++d) { par[(size_t)k * P + d] = mean[k + (size_t)K * d]; c -= log(L[d + D * d]); }
        for (int e = 0; e < D * D; ++e) par[(size_t)k * P + D + e] = Li[e];
        par[(size_t)k * P + D + D * D] = c;
      }
      if (!R->z) R->z = (int*)R->take((size_t)N * sizeof(int));
      const bool want_gamma = keep_gamma || gamma;
      if (want_gamma && (R->K_gam < K || !R->gamma)) { R->gamma = (double*)R->take((size_t)N * K * sizeof(double)); R->K_gam = K; }
      double* dpar = nullptr;
      try {
        CK(cudaMalloc((void**)&dpar, par.size() * sizeof(double)));
        CK(cudaMemcpyAsync(dpar, par.data(), par.size() * sizeof(double), cudaMemcpyHostToDevice, R->st));
        const int grid = (int)std::min<long long>((N + 255) / 256, 148 * 16);
        const size_t sm = (size_t)K * P * sizeof(double);
#define SQC_CASE_CLASSIFY_R(DD) case DD: k_gmm_classify<DD><<<grid, 256, sm, R->st>>>(R->x, R->groups, N, J, K, R->med, dpar, want_gamma ? R->gamma : nullptr, R->z); break;
        switch (D) { SQC_FOR_D(SQC_CASE_CLASSIFY_R) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", D); }
        CK(cudaGetLastError());
        if (gamma) CK(cudaMemcpyAsync(gamma, R->gamma, (size_t)N * K * sizeof(double), cudaMemcpyDeviceToHost, R->st));
        if (z) CK(cudaMemcpyAsync(z, R->z, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, R->st));
        CK(cudaStreamSynchronize(R->st));
      } catch (...) { cudaFree(dpar); throw; }
      cudaFree(dpar);
      R->K_cls = K; R->K_gcls = want_gamma ? K : 0;
    }
    CK(cudaStreamSynchronize(R->st));
  }, err, errlen);
}

// rows idx[0 .. n) of x minus their samples' medians (n x D column-major): the subsample the caller fits its mixture on (:291-300)
int sqc_resident_subsample(sqc_resident* R, const int64_t* idx, int64_t n, double* out, char* err, size_t errlen) {
  return guarded([&] {
    if (!R || !idx || !out || n < 1) fail(SQC_ERR_BAD_ARG, "bad arguments");
    CK(cudaSetDevice(R->device));
    resident_medians(R);
    for (int64_t i = 0; i < n; ++i) if (idx[i] < 0 || idx[i] >= R->N) fail(SQC_ERR_BAD_ARG, "subsample index out of range");
    long long* didx = nullptr; double* dout = nullptr;
    try {
      CK(cudaMalloc((void**)&didx, (size_t)n * sizeof(long long))); CK(cudaMalloc((void**)&dout, (size_t)n * R->D * sizeof(double)));
      CK(cudaMemcpyAsync(didx, idx, (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, R->st));
      k_gather_centred<<<(int)((n + 255) / 256), 256, 0, R->st>>>(R->x, R->groups, R->med, didx, n, R->N, R->J, R->D, dout);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(out, dout, (size_t)n * R->D * sizeof(double), cudaMemcpyDeviceToHost, R->st));
      CK(cudaStreamSynchronize(R->st));
    } catch (...) { cudaFree(didx); cudaFree(dout); throw; }
    cudaFree(didx); cudaFree(dout);
  }, err, errlen);
}

// the fit on the resident cells.  args: x and groups are ignored; init_z / init_gamma NULL = the resident classification.
// recentre != 0: mu_0, alpha_j, beta_k come back re-centred as .fit_one_sampleqc does (R/fit_sampleqc.R:242-246).
int sqc_resident_fit(sqc_resident* R, const sqc_fit_args* args, int variant, int32_t recentre, sqc_fit_out* o, char* err, size_t errlen) {
  sqc_session* S = nullptr;
  int st = guarded([&] {
    if (!R || !args || !o) fail(SQC_ERR_BAD_ARG, "null argument");
    if (variant != SQC_VARIANT_ROBUST && variant != SQC_VARIANT_MLE) fail(SQC_ERR_BAD_ARG, "unknown variant");
    if (args->world > 1) fail(SQC_ERR_BAD_ARG, "the resident pipeline runs on one GPU");
    if (!o->mu_0 || !o->alpha_j || !o->beta_k || !o->sigma_k) fail(SQC_ERR_BAD_ARG, "mu_0, alpha_j, beta_k and sigma_k outputs are needed");
    CK(cudaSetDevice(R->device));
    sqc_fit_args a = *args;
    a.D = R->D; a.J = R->J; a.N = R->N; a.x = R->x; a.groups = R->h_groups.data(); a.device = R->device; a.world = 1; a.N_total = R->N; a.n_gpus = 1;
    if (variant == SQC_VARIANT_ROBUST && !a.init_z) {
      if (!R->z || R->K_cls != a.K) fail(SQC_ERR_BAD_ARG, "no resident classification with K = %d: call sqc_resident_init_clusts first or pass init_z", a.K);
      a.init_z = R->z;                                            // device pointer: copied device to device
    }
    if (variant == SQC_VARIANT_MLE && !a.init_gamma) {
      if (!R->gamma || R->K_gcls != a.K) fail(SQC_ERR_BAD_ARG, "no resident responsibilities with K = %d: call sqc_resident_init_clusts with keep_gamma", a.K);
      a.init_gamma = R->gamma;
    }
    S = session_create(&a, variant);                            // x: device-to-device, the session centres its own copy
    session_run(S);
    session_fetch(S, o);
    const int D = R->D, J = R->J, K = a.K;
    if (recentre) {                                               // R/fit_sampleqc.R:242-246
      for (int d = 0; d < D; ++d) {
        double am = 0.0, bm = 0.0;
        for (int j = 0; j < J; ++j) am += o->alpha_j[j + (size_t)J * d];
        for (int k = 0; k < K; ++k) bm += o->beta_k[k + (size_t)K * d];
        am /= J; bm /= K;
        o->mu_0[d] = o->mu_0[d] + am + bm;
        for (int j = 0; j < J; ++j) o->alpha_j[j + (size_t)J * d] -= am;
        for (int k = 0; k < K; ++k) o->beta_k[k + (size_t)K * d] -= bm;
      }
    }
    R->K_fit = K; R->recentred = recentre != 0;
    R->mu0.assign(o->mu_0, o->mu_0 + D); R->alpha.assign(o->alpha_j, o->alpha_j + (size_t)J * D);
    R->beta.assign(o->beta_k, o->beta_k + (size_t)K * D); R->sigma.assign(o->sigma_k, o->sigma_k + (size_t)K * D * D);
  }, err, errlen);
  delete S;
  return st;
}

// .calc_outliers_dt (R/fit_sampleqc.R:419-453) with the parameters of the last sqc_resident_fit (they must have been
// re-centred: the function subtracts mu_0 and alpha_j from the raw x).  maha may be NULL.
int sqc_resident_outliers(sqc_resident* R, double alpha, double* maha, uint8_t* outlier, double* maha_cut, char* err, size_t errlen) {
  return guarded([&] {
    if (!R || R->K_fit < 1) fail(SQC_ERR_BAD_ARG, "no fit on this handle yet");
    if (!R->recentred) fail(SQC_ERR_BAD_ARG, "the last fit on this handle was not re-centred (recentre = 0): its parameters do not apply to the raw x");
    CK(cudaSetDevice(R->device));
    const int D = R->D, J = R->J, K = R->K_fit; const long long N = R->N;
    std::vector<double> ainv((size_t)K * D * D);
    for (int k = 0; k < K; ++k) if (!small_inv(R->sigma.data() + (size_t)k * D * D, D, ainv.data() + (size_t)k * D * D)) fail(SQC_ERR_SINGULAR, "solve(): system is exactly singular");
    const size_t npar = (size_t)D + (size_t)J * D + (size_t)K * D + (size_t)K * D * D + 4;
    std::vector<double> hp(npar);
    double* q = hp.data();
    memcpy(q, R->mu0.data(), D * sizeof(double)); q += D;
    memcpy(q, R->alpha.data(), (size_t)J * D * sizeof(double)); q += (size_t)J * D;
    memcpy(q, R->beta.data(), (size_t)K * D * sizeof(double)); q += (size_t)K * D;
    memcpy(q, ainv.data(), (size_t)K * D * D * sizeof(double)); q += (size_t)K * D * D;
    q[0] = 1.0 - alpha; q[1] = (double)D; q[2] = 0.0;             // qchisq(1 - alpha, df = D), :429
    double *dpar = nullptr, *dm = nullptr; uint8_t* dout = nullptr;
    auto cleanup = [&] { cudaFree(dpar); cudaFree(dm); cudaFree(dout); };
    try {
      CK(cudaMalloc((void**)&dpar, npar * sizeof(double)));
      if (maha) CK(cudaMalloc((void**)&dm, (size_t)N * K * sizeof(double)));
      if (outlier) CK(cudaMalloc((void**)&dout, (size_t)N));
      CK(cudaMemcpyAsync(dpar, hp.data(), npar * sizeof(double), cudaMemcpyHostToDevice, R->st));
      double* pm = dpar; double* pa = pm + D; double* pb = pa + (size_t)J * D; double* pi = pb + (size_t)K * D; double* pq = pi + (size_t)K * D * D;
      k_qchisq<<<1, 32, 0, R->st>>>(pq, pq + 1, pq + 2, 1);
      CK(cudaGetLastError());
      double cut = 0.0;
      CK(cudaMemcpyAsync(&cut, pq + 2, sizeof(double), cudaMemcpyDeviceToHost, R->st));
      CK(cudaStreamSynchronize(R->st));
      if (maha_cut) *maha_cut = cut;
      int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, R->device);
      const int grid = (int)std::min<long long>((N + 255) / 256, (long long)n_sm * 16);
      const size_t sm = (size_t)K * (D + D * D) * sizeof(double);
#define SQC_CASE_MAHA_R(DD) case DD: k_outlier_maha<DD><<<grid, 256, sm, R->st>>>(R->x, R->groups, N, J, K, pm, pa, pb, pi, cut, dm, dout); break;
      switch (D) { SQC_FOR_D(SQC_CASE_MAHA_R) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", D); }
      CK(cudaGetLastError());
      if (maha) CK(cudaMemcpyAsync(maha, dm, (size_t)N * K * sizeof(double), cudaMemcpyDeviceToHost, R->st));
      if (outlier) CK(cudaMemcpyAsync(outlier, dout, (size_t)N, cudaMemcpyDeviceToHost, R->st));
      CK(cudaStreamSynchronize(R->st));
    } catch (...) { cleanup(); throw; }
    cleanup();
  }, err, errlen);
}
