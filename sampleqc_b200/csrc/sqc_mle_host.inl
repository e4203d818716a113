// host driver of the MLE variant (reference src/fit_sampleQC_mle_cpp.cpp:279-357); included by sqc_lib.cu
namespace {

__global__ void k_mle_finish(Dev s, MleBufs m, int n_iter) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { s.ctl->iter = n_iter; s.ctl->status = m.iter[1]; }
}
__global__ void k_mle_pack_pk(Dev s, MleBufs m, OutPtrs o, double* out_pk) {
  if ((int)threadIdx.x < s.K) out_pk[threadIdx.x] = m.p_k[o.korder[threadIdx.x]];     // reorder_vector, :338
}

// multi-GPU: the K x NMOM moments summed over the GPUs (rank order: bit-identical everywhere), then the M-step update, in
// one launch.  The log-likelihoods are outputs only: every rank keeps its partial sums and the two vectors are summed
// over the GPUs once, after the last iteration — one exchange per iteration instead of three.
template <int D>
__global__ void __launch_bounds__(256) k_mle_xchg_update(Dev s, MleBufs m, XchgDev x) {
  xchg_sum_f64(x, XCHG_CH_HOST + 2, m.mom, s.K * SQC_NMOM(D));
  mle_update_body<D>(s, m);
}

enum { MLE_S_REDUCE_UPDATE = 0, MLE_S_SAMPLE_ALPHA = 1, MLE_S_SAMPLE = 2 };
// the small kernels between the passes, D compile-time (their solves live in registers)
template <int D> void mle_launch_small(sqc_session* S, int what, double* like, int slot) {
  Dev& s = S->dev; MleBufs& m = S->ml;
  const int J = S->J, K = S->K; const bool multi = S->world > 1;
  switch (what) {
    case MLE_S_REDUCE_UPDATE: {
      // moments + log-likelihood reduced over the tiles, then the M-step update: on one GPU inside the reduction's
      // last block; on several after the all-reduce of the moments, in that kernel
      { Launch L(S, SQC_KF_UPDATE, nullptr, "reduce_update");
        if (multi) k_mle_reduce<D, 0><<<K * SQC_NMOM(D) + 1, 256, 0, S->st>>>(s, m, like, slot);
        else k_mle_reduce<D, 1><<<K * SQC_NMOM(D) + 1, 256, 0, S->st>>>(s, m, like, slot);
        L.done(); }
      if (multi) { Launch L(S, SQC_KF_UPDATE, nullptr, "xchg_update"); k_mle_xchg_update<D><<<1, 256, 0, S->st>>>(s, m, S->xc.dev); L.done(); }
      break; }
    case MLE_S_SAMPLE_ALPHA: {
      Launch L(S, SQC_KF_UPDATE, nullptr, "sample_alpha"); k_mle_sample<D, 1><<<(J * 32 + 255) / 256 + 1, 256, 0, S->st>>>(s, m, like, slot); L.done();
      break; }
    default: {
      Launch L(S, SQC_KF_UPDATE); k_mle_sample<D, 0><<<(J * 32 + 255) / 256, 256, 0, S->st>>>(s, m, nullptr, -1); L.done();
      break; }
  }
}
#define SQC_CASE_MLE_SMALL(DD) case DD: mle_launch_small<DD>(S, what, like, slot); break;
void do_mle_small(sqc_session* S, int what, double* like, int slot) {
  switch (S->D) { SQC_FOR_D(SQC_CASE_MLE_SMALL) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); }
}

void mle_alloc(sqc_session& S, const sqc_fit_args* a, MleBufs& m) {
  const int D = S.D, J = S.J, K = S.K; const long long N = S.N; const int T = S.dev.T;
  double* g0 = S.alloc<double>((size_t)N * K, false);
  if (S.sorted) CK(cudaMemcpy2DAsync(g0, (size_t)N * sizeof(double), a->init_gamma, (size_t)S.ld * sizeof(double), (size_t)N * sizeof(double), K, cudaMemcpyDefault, S.st));
  else {
    std::vector<double> col(N);
    for (int k = 0; k < K; ++k) {
      for (long long i = 0; i < N; ++i) col[i] = a->init_gamma[(size_t)k * S.ld + S.perm[i]];
      CK(cudaMemcpyAsync(g0 + (size_t)k * N, col.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, S.st));
      CK(cudaStreamSynchronize(S.st));
    }
  }
  m.gamma0 = g0;
  m.gamma = S.alloc<double>((size_t)N * K);
  m.t_mom = S.alloc<double>((size_t)T * K * SQC_NMOM(D));
  m.t_ll = S.alloc<double>(T);
  m.mom = S.alloc<double>((size_t)K * SQC_NMOM(D));
  m.g_jk = S.alloc<double>((size_t)J * K); m.x_jk = S.alloc<double>((size_t)J * K * D);
  m.p_k = S.alloc<double>(K + SQC_MAXK); m.coef = S.alloc<double>(K);
  m.iter = S.alloc<int>(4);
}

void mle_run(sqc_session* S) {
  Dev& s = S->dev; MleBufs& m = S->ml;
  const int J = S->J, K = S->K, D = S->D, n_iter = std::max(S->em_iters, 0);
  CK(cudaEventRecord(S->ev_run0, S->st));
  { Launch L(S, SQC_KF_UPDATE); k_ctl_reset<<<1, 32, 0, S->st>>>(s); L.done(); }
  CK(cudaMemcpyAsync(s.alpha, S->alpha0, (size_t)J * D * sizeof(double), cudaMemcpyDeviceToDevice, S->st));   // init_alpha_j, :300
  CK(cudaMemsetAsync(s.beta, 0, (size_t)K * D * sizeof(double), S->st));
  CK(cudaMemsetAsync(s.sigma, 0, (size_t)K * D * D * sizeof(double), S->st));
  CK(cudaMemsetAsync(s.like1, 0, (size_t)(n_iter + 1) * sizeof(double), S->st));
  CK(cudaMemsetAsync(s.like2, 0, (size_t)(n_iter + 1) * sizeof(double), S->st));
  CK(cudaMemsetAsync(m.iter, 0, 4 * sizeof(int), S->st));
  CK(cudaMemsetAsync(s.z, 0, (size_t)S->N, S->st));
  CK(cudaMemsetAsync(m.g_jk, 0, (size_t)J * K * sizeof(double), S->st));
  if (n_iter == 0) CK(cudaMemsetAsync(m.gamma, 0, (size_t)S->N * K * sizeof(double), S->st));            // gamma_i stays zeros, :291
  const bool multi = S->world > 1;
  auto reduce_update = [&](double* like, int slot) { do_mle_small(S, MLE_S_REDUCE_UPDATE, like, slot); };
  // beta_k, Sigma_k, p_k from init_gamma_i (:301-303)
  m.last_store = 0;
  do_mle_pass(S, MLE_INIT);
  reduce_update(s.like1, -1);
  CK(cudaEventRecord(S->ev_init, S->st));
  for (int i = 0; i < n_iter; ++i) {                                                                    // :307
    do_mle_pass(S, MLE_A);                                                                              // :313
    // :316, and like_2 of iteration i-1 (:326) summed by the launch's extra block
    do_mle_small(S, MLE_S_SAMPLE_ALPHA, s.like2, i - 1);
    m.last_store = (i == n_iter - 1) ? 1 : 0;
    do_mle_pass(S, MLE_B);                                                                              // :317, :320
    reduce_update(s.like1, i);                                                                          // :323-325
  }
  if (n_iter > 0) {
    do_mle_small(S, MLE_S_SAMPLE, nullptr, -1);                                                         // extract_p_jk, :331
    do_mle_pass(S, MLE_LIKE);
    { Launch L(S, SQC_KF_UPDATE); k_mle_reduce_ll<<<1, 1024, 0, S->st>>>(s, m, s.like2, n_iter - 1); L.done(); }
  }
  if (multi)                                                          // like_1 / like_2: partial sums until here
    for (int i0 = 0; i0 < n_iter; i0 += 4096) {
      const int n = std::min(4096, n_iter - i0);
      { Launch L(S, SQC_KF_UPDATE); k_xchg_sum_f64<<<1, 256, 0, S->st>>>(S->xc.dev, XCHG_CH_HOST + 3, s.like1 + i0, n, nullptr); L.done(); }
      { Launch L(S, SQC_KF_UPDATE); k_xchg_sum_f64<<<1, 256, 0, S->st>>>(S->xc.dev, XCHG_CH_HOST + 3, s.like2 + i0, n, nullptr); L.done(); }
    }
  { Launch L(S, SQC_KF_UPDATE); k_mle_pjk<<<(J + 127) / 128, 128, 0, S->st>>>(s, m); L.done(); }
  { Launch L(S, SQC_KF_UPDATE); k_mle_finish<<<1, 32, 0, S->st>>>(s, m, n_iter); L.done(); }
  CK(cudaEventRecord(S->ev_loop, S->st));
}

void mle_fetch(sqc_session* S, sqc_fit_out* o) {
  Dev& s = S->dev; MleBufs& m = S->ml;
  const int D = S->D, J = S->J, K = S->K, n_iter = std::max(S->em_iters, 0); const long long N = S->N;
  OutPtrs op = S->optr;
  if (!o->z) op.z = nullptr;
  k_order_components<<<1, 32, 0, S->st>>>(s, op);
  // extract_z writes idx(k_max) (:241), not the inverse permutation
  k_pack_outputs<<<std::max(1, std::min<int>((int)((N + 255) / 256), S->n_sm * 8)), 256, 0, S->st>>>(s, op, 2);
  k_mle_pack_pk<<<1, 32, 0, S->st>>>(s, m, op, m.p_k + K);
  CK(cudaGetLastError());
  std::vector<double> h(S->o_doubles);
  CK(cudaMemcpyAsync(h.data(), S->o_buf, S->o_doubles * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  if (o->p_k) CK(cudaMemcpyAsync(o->p_k, m.p_k + K, (size_t)K * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  if (o->like_1) CK(cudaMemcpyAsync(o->like_1, s.like1, (size_t)n_iter * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  if (o->like_2) CK(cudaMemcpyAsync(o->like_2, s.like2, (size_t)n_iter * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  if (S->sorted) {
    if (o->z) CK(cudaMemcpyAsync(o->z, S->o_z, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, S->st));
    if (o->gamma_i) CK(cudaMemcpyAsync(o->gamma_i, m.gamma, (size_t)N * K * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
  } else {
    std::vector<int> zz(N); std::vector<double> col(N);
    CK(cudaMemcpyAsync(zz.data(), S->o_z, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    if (o->z) for (long long i = 0; i < N; ++i) o->z[S->perm[i]] = zz[i];
    if (o->gamma_i) for (int k = 0; k < K; ++k) {
      CK(cudaMemcpy(col.data(), m.gamma + (size_t)k * N, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost));
      for (long long i = 0; i < N; ++i) o->gamma_i[(size_t)k * N + S->perm[i]] = col[i];
    }
  }
  const double* p = h.data();
  if (o->mu_0) memcpy(o->mu_0, p, sizeof(double) * D); p += D;
  auto rows_out = [&](double* dst, const double* src, int ncol) {    // internal sample order -> the caller's codes (session_create)
    if (S->smap.empty()) { memcpy(dst, src, sizeof(double) * J * ncol); return; }
    for (int c = 0; c < ncol; ++c) for (int r = 0; r < J; ++r) dst[S->smap[r] + (size_t)J * c] = src[r + (size_t)J * c];
  };
  if (o->alpha_j) rows_out(o->alpha_j, p, D); p += (size_t)J * D;
  if (o->beta_k) memcpy(o->beta_k, p, sizeof(double) * K * D); p += (size_t)K * D;
  if (o->sigma_k) memcpy(o->sigma_k, p, sizeof(double) * K * D * D); p += (size_t)K * D * D;
  if (o->p_jk) rows_out(o->p_jk, p, K);
  memset(o->rng_state, 0, sizeof o->rng_state);
  o->n_iter = n_iter; o->mcd_iters_hit = 0; o->n_zero_like = 0; o->rng_draws = 0;
}

}  // namespace
