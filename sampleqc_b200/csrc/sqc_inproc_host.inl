// sqc_inproc_host.inl — sqc_fit_args.n_gpus: ONE call, several GPUs (included by sqc_lib.cu inside its anonymous namespace).
//
// fit_sampleqc() calls fit_sampleqc_*_cpp once per sample group from a single R thread (reference R/fit_sampleqc.R:229-238),
// so the multi-GPU fit has to be reachable from that one call: the library shards the cells by sample (contiguous runs
// of samples with balanced cell counts), starts one host thread per GPU, and every thread builds
// and runs the session of its shard — the same sessions a one-process-per-GPU launch (torchrun) creates, except that
// the peers' exchange windows and key buffers are ordinary device pointers (cudaDeviceEnablePeerAccess; no IPC
// handles, nothing to bootstrap out of band).  Per-sample outputs (alpha_j, p_jk) and per-cell outputs (z, gamma_i) are
// written / scattered into the caller's buffers; the global ones are identical on every rank and taken from rank 0.
struct InprocShard {
  long long i0 = 0, i1 = 0;                  // cells
  std::vector<int> codes;                    // caller's sample code of local sample r
  std::vector<int32_t> groups;               // local sample index of every cell (ascending)
};

// how many GPUs the call may use: n_gpus <= 1 -> one (args.device); n_gpus > 1 -> that many; SQC_ALL_GPUS -> all visible
int resolve_n_gpus(const sqc_fit_args* a) {
  if (a->world > 1) return 1;                // the caller runs one process per GPU already
  int want = a->n_gpus;
  if (want == 0 || want == 1) return 1;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return 1;      // the single-GPU path reports the missing device
  n = std::min(n, XCHG_MAX_WORLD);
  return want < 0 ? n : std::min(want, n);
}

// contiguous runs of samples -> shards with balanced cell counts; false: cells of a sample are interleaved
bool make_shards(const sqc_fit_args* a, int n, std::vector<InprocShard>& out) {
  const long long N = a->N; const int J = a->J;
  struct Run { int code; long long i0, i1; };
  std::vector<Run> runs;
  std::vector<char> seen(J, 0);
  for (long long i = 0; i < N;) {
    const int g = a->groups[i];
    if ((unsigned)g >= (unsigned)J) fail(SQC_ERR_BAD_ARG, "groups out of range at cell %lld", i);
    if (seen[g]) return false;
    seen[g] = 1;
    long long e = i + 1;
    while (e < N && a->groups[e] == g) ++e;
    runs.push_back({g, i, e});
    i = e;
  }
  n = std::max(1, std::min<int>(n, (int)runs.size()));
  out.assign(n, InprocShard());
  size_t r = 0;
  for (int q = 0; q < n; ++q) {
    InprocShard& s = out[q];
    const size_t r_begin = r;
    const long long target = N * (q + 1) / n;
    s.i0 = runs[r].i0;
    // at least one run per shard; leave one for every shard behind this one; a run goes to the shard its middle lies in
    do { s.i1 = runs[r].i1; ++r; }
    while (r < runs.size() && (runs.size() - r) > (size_t)(n - 1 - q) && (q == n - 1 || (runs[r].i0 + runs[r].i1) / 2 <= target));
    s.groups.resize((size_t)(s.i1 - s.i0));
    long long pos = 0;
    for (size_t rr = r_begin; rr < r; ++rr) {
      const int loc = (int)(rr - r_begin);
      s.codes.push_back(runs[rr].code);
      for (long long i = runs[rr].i0; i < runs[rr].i1; ++i) s.groups[(size_t)pos++] = loc;
    }
    if (q == n - 1) for (int j = 0; j < J; ++j) if (!seen[j]) s.codes.push_back(j);      // samples without cells: rows of NaN, as in the reference
  }
  return true;
}

int fit_inproc(const sqc_fit_args* a, sqc_fit_out* o, int variant, int n_req, char* err, size_t errlen) {
  std::vector<InprocShard> sh;
  std::string first_err; int first_code = SQC_OK;
  int st0 = guarded([&] {
    check_args(a, variant);
    if (!o) fail(SQC_ERR_BAD_ARG, "null output");
    if (!make_shards(a, n_req, sh)) sh.clear();
  }, err, errlen);
  if (st0 != SQC_OK) return st0;
  if (sh.size() < 2) return -1000;                                   // one shard / interleaved samples: the caller takes the single-GPU path
  const int W = (int)sh.size(), D = a->D, K = a->K, J = a->J, em = std::max(a->em_iters, 0); const long long N = a->N;
  static std::mutex inproc_mu;                                       // the exchange window of a GPU is one per process:
  std::lock_guard<std::mutex> one_fit(inproc_mu);                    // multi-GPU fits of different host threads take turns
  InprocGroup grp; grp.world = W;
  for (int q = 0; q < W; ++q) grp.device[q] = q;
  struct RankOut { std::vector<double> alpha, p_jk, gamma, tr_alpha, tr_p; sqc_fit_out o{}; int code = SQC_OK; std::string msg; };
  std::vector<RankOut> ro(W);
  std::vector<std::thread> th;
  th.reserve(W);
  for (int q = 0; q < W; ++q) {
    try { th.emplace_back([&, q] {
      RankOut& R = ro[q]; const InprocShard& s = sh[q];
      const int Jl = (int)s.codes.size(); const long long Nl = s.i1 - s.i0;
      sqc_fit_args b = *a;
      b.J = Jl; b.N = Nl; b.x = a->x + s.i0; b.groups = s.groups.data();
      b.init_z = a->init_z ? a->init_z + s.i0 : nullptr; b.init_gamma = a->init_gamma ? a->init_gamma + s.i0 : nullptr;
      b.device = grp.device[q]; b.rank = q; b.world = W; b.cell_offset = s.i0; b.N_total = N; b.peer_handles = nullptr; b.n_gpus = 1;
      sqc_fit_out& lo = R.o;
      R.alpha.resize((size_t)Jl * D); R.p_jk.resize((size_t)Jl * K);
      lo.alpha_j = R.alpha.data(); lo.p_jk = R.p_jk.data();
      lo.z = o->z ? o->z + s.i0 : nullptr;
      if (variant == SQC_VARIANT_MLE && o->gamma_i) { R.gamma.resize((size_t)Nl * K); lo.gamma_i = R.gamma.data(); }
      if (a->track && variant == SQC_VARIANT_ROBUST) {
        if (o->track_alpha_j) { R.tr_alpha.resize((size_t)Jl * D * em); lo.track_alpha_j = R.tr_alpha.data(); }
        if (o->track_p_jk) { R.tr_p.resize((size_t)Jl * K * em); lo.track_p_jk = R.tr_p.data(); }
      }
      if (q == 0) {                                                 // global outputs: identical on every rank
        lo.mu_0 = o->mu_0; lo.beta_k = o->beta_k; lo.sigma_k = o->sigma_k; lo.like_1 = o->like_1; lo.like_2 = o->like_2; lo.p_k = o->p_k;
        lo.track_beta_k = o->track_beta_k; lo.track_sigma_k = o->track_sigma_k;
      }
      sqc_session* S = nullptr;
      char e2[512] = {0};
      R.code = guarded([&] {
        try { S = session_create(&b, variant, &grp, N); }
        catch (...) { grp.abort(); throw; }
        if (!grp.sync(true)) fail(SQC_ERR_COMM, "another GPU of the fit failed while the sessions were built");
        session_run(S);
        session_fetch(S, &lo);
      }, e2, sizeof e2);
      if (R.code != SQC_OK) { grp.abort(); R.msg = e2; }
      delete S;
    }); }
    catch (const std::exception& e) {                               // no thread for this rank: the others give up at their next sync
      grp.abort(); ro[q].code = SQC_ERR_CUDA; ro[q].msg = std::string("could not start a host thread: ") + e.what();
      break;
    }
  }
  for (auto& t : th) t.join();
  for (int q = 0; q < W; ++q) if (ro[q].code != SQC_OK && (first_code == SQC_OK || first_code == SQC_ERR_COMM)) { first_code = ro[q].code; first_err = ro[q].msg; }
  if (first_code != SQC_OK) { if (err && errlen) snprintf(err, errlen, "%s", first_err.c_str()); return first_code; }
  // scatter the per-sample rows back to the caller's sample codes, the per-cell columns to the caller's rows
  for (int q = 0; q < W; ++q) {
    const InprocShard& s = sh[q]; const RankOut& R = ro[q];
    const int Jl = (int)s.codes.size(); const long long Nl = s.i1 - s.i0;
    for (int r = 0; r < Jl; ++r) {
      const int j = s.codes[r];
      if (o->alpha_j) for (int d = 0; d < D; ++d) o->alpha_j[j + (size_t)J * d] = R.alpha[r + (size_t)Jl * d];
      if (o->p_jk) for (int k = 0; k < K; ++k) o->p_jk[j + (size_t)J * k] = R.p_jk[r + (size_t)Jl * k];
      for (int it = 0; it < em; ++it) {
        if (!R.tr_alpha.empty()) for (int d = 0; d < D; ++d) o->track_alpha_j[(size_t)it * J * D + j + (size_t)J * d] = R.tr_alpha[(size_t)it * Jl * D + r + (size_t)Jl * d];
        if (!R.tr_p.empty()) for (int k = 0; k < K; ++k) o->track_p_jk[(size_t)it * J * K + j + (size_t)J * k] = R.tr_p[(size_t)it * Jl * K + r + (size_t)Jl * k];
      }
    }
    if (!R.gamma.empty()) for (int k = 0; k < K; ++k) memcpy(o->gamma_i + (size_t)k * N + s.i0, R.gamma.data() + (size_t)k * Nl, (size_t)Nl * sizeof(double));
  }
  o->n_iter = ro[0].o.n_iter; o->mcd_iters_hit = ro[0].o.mcd_iters_hit; o->rng_draws = ro[0].o.rng_draws;
  memcpy(o->rng_state, ro[0].o.rng_state, sizeof o->rng_state);
  int nz = 0; for (int q = 0; q < W; ++q) nz += ro[q].o.n_zero_like;
  o->n_zero_like = nz;
  if (err && errlen) err[0] = 0;
  return SQC_OK;
}
