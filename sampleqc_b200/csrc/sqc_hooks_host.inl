// sqc_hooks_host.inl — test hooks of the C ABI (included by sqc_lib.cu inside extern "C"): the device-side special
// functions and key-stream generators exposed to the parity tests.  Not used by the fits.
int sqc_device_qchisq(const double* p, const double* df, double* out, int32_t n, int device, char* err, size_t errlen) {
  return guarded([&] {
    if (!p || !df || !out || n < 0) fail(SQC_ERR_BAD_ARG, "bad arguments");
    int d; select_device(device, &d);
    double* buf = nullptr; CK(cudaMalloc((void**)&buf, (size_t)3 * std::max(n, 1) * sizeof(double)));
    cudaMemcpy(buf, p, n * sizeof(double), cudaMemcpyHostToDevice); cudaMemcpy(buf + n, df, n * sizeof(double), cudaMemcpyHostToDevice);
    k_qchisq<<<(n + 127) / 128, 128>>>(buf, buf + n, buf + 2 * (size_t)n, n);
    cudaError_t e = cudaMemcpy(out, buf + 2 * (size_t)n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(buf); CK(e);
  }, err, errlen);
}
int sqc_device_expneg(const double* x, double* out, int32_t n, int device, char* err, size_t errlen) {
  return guarded([&] {
    if (!x || !out || n < 0) fail(SQC_ERR_BAD_ARG, "bad arguments");
    int d; select_device(device, &d);
    double* buf = nullptr; CK(cudaMalloc((void**)&buf, (size_t)2 * std::max(n, 1) * sizeof(double)));
    cudaMemcpy(buf, x, n * sizeof(double), cudaMemcpyHostToDevice);
    k_expneg<<<(n + 127) / 128, 128>>>(buf, buf + n, n);
    cudaError_t e = cudaMemcpy(out, buf + n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(buf); CK(e);
  }, err, errlen);
}
int sqc_device_rkeys(uint32_t seed, int64_t skip, int32_t* keys_out, int64_t n, int device, char* err, size_t errlen) {
  return guarded([&] {
    if (!keys_out || n < 0 || skip < 0) fail(SQC_ERR_BAD_ARG, "bad arguments");
    int d; select_device(device, &d);
    uint32_t *st = nullptr, *st2 = nullptr, *st3 = nullptr, *st4 = nullptr, *seq = nullptr, *table = nullptr; int* out = nullptr;
    auto cleanup = [&] { cudaFree(st); cudaFree(st2); cudaFree(st3); cudaFree(st4); cudaFree(seq); cudaFree(table); cudaFree(out); };
    try {
      CK(cudaMalloc((void**)&st, 640 * sizeof(uint32_t))); CK(cudaMalloc((void**)&st2, 640 * sizeof(uint32_t)));
      CK(cudaMalloc((void**)&st3, 640 * sizeof(uint32_t))); CK(cudaMalloc((void**)&st4, 640 * sizeof(uint32_t)));
      CK(cudaMalloc((void**)&seq, (MT_SEQ_WORDS + (size_t)(SQC_MT_JUMP_ROWS + 2) * MT_N) * sizeof(uint32_t))); CK(cudaMalloc((void**)&table, sizeof sqc_mt_jump_table));
      CK(cudaMalloc((void**)&out, (size_t)std::max<int64_t>(std::max(n, skip), 1) * sizeof(int)));
      CK(cudaMemcpy(table, sqc_mt_jump_table, sizeof sqc_mt_jump_table, cudaMemcpyHostToDevice));
      CK(cudaFuncSetAttribute(k_mt_jump, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
      k_mt_seed<<<1, 32>>>(st, seed);
      // skipped draws are generated and discarded as a separate call, so the state hand-over is exercised
      // target_blocks = 1000: the shortest blocks, so that small requests still exercise the jump table
      if (skip > 0) { mt_generate(0, st, st2, st3, st4, seq, table, out, skip, nullptr, nullptr, 1000); std::swap(st, st2); }
      mt_generate(0, st, st2, st3, st4, seq, table, out, n, nullptr, nullptr, n > 2000000 ? 20 : 1000);
      CK(cudaGetLastError());
      CK(cudaMemcpy(keys_out, out, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    } catch (...) { cleanup(); throw; }
    cleanup();
  }, err, errlen);
}

// test hook for the distributed key stream (what one rank of `world` does over `calls` consecutive calls of n_total draws):
// keys_out = calls x Q ints (rank `rank`'s share of every call, Q = *q_out draws per rank), state_out = R's state after the
// last call (identical on every rank).  Runs on one GPU: the pieces (host-derived jump polynomials, k_mt_jump_rows,
// k_mt_fix, share generation) are the ones sqc_fit_robust uses with world > 1.
int sqc_device_rkeys_share(uint32_t seed, int64_t n_total, int world, int rank, int calls, int32_t* keys_out, int64_t* q_out, int32_t* state_out,
                           int device, char* err, size_t errlen) {
  return guarded([&] {
    if (!keys_out || !q_out || n_total < 624LL * 8 * world || world < 1 || world > XCHG_MAX_WORLD || rank < 0 || rank >= world || calls < 1) fail(SQC_ERR_BAD_ARG, "bad arguments");
    int d; select_device(device, &d);
    if (!mtpoly::self_check()) fail(SQC_ERR_COMM, "jump polynomial derivation failed its self-check");
    const long long per = (n_total + world - 1) / world, Q = ((per + MT_N - 1) / MT_N) * MT_N;
    const long long mine = std::max<long long>(0, std::min<long long>(Q, n_total - (long long)rank * Q));
    *q_out = Q;
    uint32_t *st = nullptr, *seq = nullptr, *seq2 = nullptr, *table = nullptr, *rows = nullptr, *jwin = nullptr; int* out = nullptr;
    auto cleanup = [&] { cudaFree(st); cudaFree(seq); cudaFree(seq2); cudaFree(table); cudaFree(rows); cudaFree(jwin); cudaFree(out); };
    try {
      CK(cudaMalloc((void**)&st, 640 * 8 * sizeof(uint32_t)));
      CK(cudaMalloc((void**)&seq, (MT_SEQ_WORDS + (size_t)(SQC_MT_JUMP_ROWS + 2) * MT_N) * sizeof(uint32_t))); CK(cudaMalloc((void**)&seq2, (MT_SEQ_WORDS + 64) * sizeof(uint32_t)));
      CK(cudaMalloc((void**)&table, sizeof sqc_mt_jump_table)); CK(cudaMalloc((void**)&rows, 2 * MT_N * sizeof(uint32_t))); CK(cudaMalloc((void**)&jwin, 2 * MT_N * sizeof(uint32_t)));
      CK(cudaMalloc((void**)&out, (size_t)std::max<long long>(mine, 1) * sizeof(int)));
      CK(cudaMemcpy(table, sqc_mt_jump_table, sizeof sqc_mt_jump_table, cudaMemcpyHostToDevice));
      std::vector<uint32_t> r0(MT_N), r1(MT_N);
      mtpoly::xpow((unsigned long long)(n_total / MT_N) * MT_N - 1ull, r0.data());
      CK(cudaMemcpy(rows, r0.data(), MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice));
      const bool has_rank = rank > 0 && mine > 0;
      if (has_rank) { mtpoly::xpow((unsigned long long)rank * (unsigned long long)Q - 1ull, r1.data()); CK(cudaMemcpy(rows + MT_N, r1.data(), MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice)); }
      CK(cudaFuncSetAttribute(k_mt_jump, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
      CK(cudaFuncSetAttribute(k_mt_jump_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
      k_mt_seed<<<1, 32>>>(st + 640 * 2, seed);
      for (int c = 0; c < calls; ++c) {
        const uint32_t* in = st + 640 * ((c + 2) % 3);
        k_mt_seq<<<1, 256>>>(in, seq2, nullptr);
        k_mt_jump_rows<<<has_rank ? 2 : 1, MT_JUMP_THREADS, MT_SMEM>>>(seq2, rows, jwin, nullptr);
        k_mt_fix<<<1, 256>>>(in, jwin, st + 640 * (c % 3), st + 640 * 5, n_total, has_rank ? 1 : 0, nullptr);
        if (mine > 0) {
          mt_generate(0, st + 640 * 5, st + 640 * 7, st + 640 * 3, st + 640 * 4, seq, table, out, mine, nullptr, nullptr, mine > 2000000 ? 20 : 1000);
          CK(cudaMemcpy(keys_out + (size_t)c * Q, out, (size_t)mine * sizeof(int), cudaMemcpyDeviceToHost));
        }
      }
      CK(cudaGetLastError());
      if (state_out) CK(cudaMemcpy(state_out, st + 640 * ((calls - 1) % 3), 625 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    } catch (...) { cleanup(); throw; }
    cleanup();
  }, err, errlen);
}

// host-only test hook: the jump polynomial x^e mod phi_MT19937 the distributed key stream derives at session creation
// (624 words, bit i of word w = coefficient of x^(32 w + i)); needs no GPU
int sqc_mt_jump_poly(uint64_t e, uint32_t* out624) {
  if (!out624) return SQC_ERR_BAD_ARG;
  mtpoly::xpow((unsigned long long)e, out624);
  return mtpoly::self_check() ? SQC_OK : SQC_ERR_COMM;
}

// measured fp64 FMA throughput of the device (the co-limit of the E-step / partition passes):
// every thread runs 8 independent DFMA chains; returns FMA instructions per second over the whole GPU (x 2 = FLOP/s)
int sqc_fp64_peak(int device, double* dfma_per_s, char* err, size_t errlen) {
  return guarded([&] {
    if (!dfma_per_s) fail(SQC_ERR_BAD_ARG, "null output");
    int d; select_device(device, &d);
    int n_sm = 0; CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, d));
    double* out = nullptr; CK(cudaMalloc((void**)&out, sizeof(double)));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int iters = 8192, grid = n_sm * 2, threads = 1024;
    k_fp64_peak<<<grid, threads>>>(out, 64, 1.0000001);            // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(a)); k_fp64_peak<<<grid, threads>>>(out, iters, 1.0000001); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms = 0; CK(cudaEventElapsedTime(&ms, a, b)); best = std::min(best, ms);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    *dfma_per_s = (double)grid * threads * 8.0 * iters / (best * 1e-3);
  }, err, errlen);
}

// host-only test hook: the shards sqc_fit_robust / sqc_fit_mle cut for n_gpus GPUs (sqc_inproc_host.inl).  shard_cells
// receives n_gpus + 1 cell boundaries, sample_rank[j] the shard that owns sample code j (-1: a sample without cells goes
// to the last shard).  Returns the number of shards (<= n_gpus), 0 when the samples are interleaved (one GPU is used),
// < 0 on bad arguments.  Needs no GPU.
int sqc_plan_shards(const int32_t* groups, int64_t N, int32_t J, int32_t n_gpus, int64_t* shard_cells, int32_t* sample_rank) {
  if (!groups || N < 1 || J < 1 || n_gpus < 1 || n_gpus > XCHG_MAX_WORLD || !shard_cells) return SQC_ERR_BAD_ARG;
  sqc_fit_args a = {}; a.N = N; a.J = J; a.groups = groups;
  std::vector<InprocShard> sh;
  try { if (!make_shards(&a, n_gpus, sh)) return 0; } catch (const SqcFail& e) { return e.code > 0 ? -e.code : e.code; }
  if (sample_rank) for (int j = 0; j < J; ++j) sample_rank[j] = -1;
  for (size_t q = 0; q < sh.size(); ++q) {
    shard_cells[q] = sh[q].i0; shard_cells[q + 1] = sh[q].i1;
    if (sample_rank) for (int c : sh[q].codes) if (sample_rank[c] < 0 && sh[q].i1 > sh[q].i0) {
      // codes of empty samples are appended to the last shard: keep them at -1
      bool has = false;
      for (long long i = sh[q].i0; i < sh[q].i1 && !has; ++i) has = groups[i] == c;
      if (has) sample_rank[c] = (int)q;
    }
  }
  return (int)sh.size();
}
