// sqc_keys_host.inl — host side of the R-compatible key stream (included by sqc_lib.cu inside its anonymous namespace):
// generation waves (mt_generate), the per-call schedule on the side streams (enqueue_* / use_keys / release_keys) and the
// multi-GPU distribution of the stream in per-rank shares (key_stream_setup).  Kernels: sqc_mt.cuh.
// ---- distributed key stream (multi-GPU, SQC_RNG_R_COMPAT) --------------------------------------------------------
// R's stream is one sequence, so a call's N_total keys cost every rank O(N_total) when each generates them all.
// Instead rank r generates draws [r Q, (r + 1) Q) of every call into a buffer the peers map (cudaIpc), and the M-step
// kernel reads a cell's key from its owner over NVLink (4 B per cell in the one or two key rounds of a call).  The
// buffers' IPC handles travel through the exchange window itself, so the C ABI is unchanged.
struct KeyBufLocal { int* p = nullptr; size_t ints = 0; std::vector<int*> retired; };
KeyBufLocal& key_buf_local(int device) { static KeyBufLocal b[64]; return b[device & 63]; }

void key_stream_setup(sqc_session* S) {
  const int W = S->world, r = S->rank;
  { static std::mutex mu; std::lock_guard<std::mutex> lk(mu); if (!mtpoly::self_check()) fail(SQC_ERR_COMM, "distributed key stream: jump polynomial derivation failed its self-check"); }
  const long long per = (S->N_total + W - 1) / W;
  S->key_q = ((per + MT_N - 1) / MT_N) * MT_N;
  S->key_mine = std::max<long long>(0, std::min<long long>(S->key_q, S->N_total - (long long)r * S->key_q));
  // own buffer (two calls in flight), kept per device for the life of the process: its IPC handle stays valid
  KeyBufLocal& kb = key_buf_local(S->device);
  if (kb.ints < 2 * (size_t)S->key_q) {
    // a buffer that has been exported stays allocated: peers keep their mapping of it for the life of the process
    // (IpcCache), and freeing an exported allocation under a live mapping is undefined.  Retired buffers go with
    // sqc_xchg_release.
    if (kb.p) { kb.retired.push_back(kb.p); kb.p = nullptr; kb.ints = 0; }
    CK(cudaMalloc((void**)&kb.p, 2 * (size_t)S->key_q * sizeof(int)));
    kb.ints = 2 * (size_t)S->key_q;
  }
  S->keybuf = kb.p;
  if (S->grp) {
    // in-process ranks: every key buffer is an ordinary device pointer once all of them exist
    CK(cudaDeviceSynchronize());
    if (!S->grp->sync(true)) fail(SQC_ERR_COMM, "another GPU of the fit failed while the key buffers were set up");
    for (int q = 0; q < W; ++q) S->key_peer[q] = key_buf_local(S->grp->device[q]).p;
  } else {
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, kb.p) != cudaSuccess) fail(SQC_ERR_COMM, "cudaIpcGetMemHandle (key buffer) failed");
    // all-gather of the 64-byte handles through the window
    unsigned long long* d_in = S->alloc<unsigned long long>(8); unsigned long long* d_out = S->alloc<unsigned long long>(8 * XCHG_MAX_WORLD);
    CK(cudaMemcpyAsync(d_in, &h, 64, cudaMemcpyHostToDevice, S->st));
    k_xchg_gather_u64<<<1, 256, 0, S->st>>>(S->xc.dev, XCHG_CH_HOST + 0, d_in, d_out, 8, nullptr);
    std::vector<unsigned char> hs((size_t)64 * W);
    CK(cudaMemcpyAsync(hs.data(), d_out, hs.size(), cudaMemcpyDeviceToHost, S->st));
    int st_dev = 0;
    CK(cudaMemcpyAsync(&st_dev, &S->dev.ctl->status, sizeof(int), cudaMemcpyDeviceToHost, S->st));
    CK(cudaStreamSynchronize(S->st));
    if (st_dev == SQC_D_TIMEOUT) fail(SQC_ERR_COMM, "%s", device_status_message(st_dev));      // a peer never sent its handle
    for (int q = 0; q < W; ++q) {
      if (q == r) { S->key_peer[q] = kb.p; continue; }
      try { S->key_peer[q] = (const int*)ipc_cache().open(S->device, hs.data() + (size_t)64 * q); }       // mapped once per process and peer buffer
      catch (const std::exception& e) { fail(SQC_ERR_COMM, "key buffer of rank %d: %s", q, e.what()); }
    }
  }
  // jump polynomials: row 0 = call to call (floor(N_total / 624) chunks), row 1 = start of the call to this rank's share
  static std::map<unsigned long long, std::vector<uint32_t>> rows;      // derived once per distance
  static std::mutex rows_mu;                                            // (the ranks of an in-process fit are threads)
  auto row = [&](unsigned long long words) -> const std::vector<uint32_t>& {
    std::lock_guard<std::mutex> lk(rows_mu);
    auto it = rows.find(words);
    if (it == rows.end()) { std::vector<uint32_t> v(MT_N); mtpoly::xpow(words - 1ull, v.data()); it = rows.emplace(words, std::move(v)).first; }
    return it->second;
  };
  S->mt_rows = S->alloc<uint32_t>(2 * MT_N); S->mt_jwin = S->alloc<uint32_t>(2 * MT_N); S->mt_seq2 = S->alloc<uint32_t>(MT_SEQ_WORDS + 64);
  if (getenv("SQC_RNG_INLINE") && atoi(getenv("SQC_RNG_INLINE"))) S->st_rng2 = S->st;
  else CK(cudaStreamCreateWithFlags(&S->st_rng2, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&S->ev_chain[i], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&S->ev_post_mcd, cudaEventDisableTiming));
  CK(cudaMemcpyAsync(S->mt_rows, row((unsigned long long)(S->N_total / MT_N) * MT_N).data(), MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice, S->st));
  if (r > 0 && S->key_mine > 0)
    CK(cudaMemcpyAsync(S->mt_rows + MT_N, row((unsigned long long)r * (unsigned long long)S->key_q).data(), MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice, S->st));
  CK(cudaStreamSynchronize(S->st));                                      // the rows live in a static map, but keep the copies simple
  CK(cudaFuncSetAttribute(k_mt_jump_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
}

// n keys of R's stream: reads the state st_in (625 words), leaves the state after the n draws in st_out.
// Waves of at most SQC_MT_JUMP_ROWS + 1 blocks; each wave = one sequential 33-chunk kernel + one
// block-parallel kernel; tmp[2] hold the states between waves.
void mt_generate(cudaStream_t stream, const uint32_t* st_in, uint32_t* st_out, uint32_t* tmp0, uint32_t* tmp1, uint32_t* seq,
                 const uint32_t* table, int* out, long long n, const Control* ctl, sqc_session* S, int target_blocks = 49) {
  long long done = 0; int w = 0;
  const uint32_t* cur = st_in;
  if (n <= 0) { CK(cudaMemcpyAsync(st_out, st_in, 625 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream)); return; }
  // block length = mult * 110 chunks; a wave serves up to 148 / mult + 1 blocks.  Longer blocks need fewer GF(2)
  // correlations (the expensive part, which competes with the E-step for the SMs) at the price of a longer
  // sequential tail; measured best at C3: ~49 blocks of 330 chunks (mult = 3).  Never more than 3: the generation
  // of one call has to fit beside one EM iteration.  target_blocks >= 146 asks for the shortest blocks.
  int mult = 1;
  if (target_blocks < SQC_MT_JUMP_ROWS - 2) {
    mult = (int)std::max<long long>(1, (n / MT_BLOCK_WORDS + target_blocks / 2) / target_blocks);
    if (S) mult = std::min(mult, 3);
    mult = std::min(mult, SQC_MT_JUMP_ROWS / 2);
  }
  const long long blk_words = MT_BLOCK_WORDS * mult;
  const long long wave_words = (long long)(SQC_MT_JUMP_ROWS / mult + 1) * blk_words;
  while (done < n) {
    const long long m = std::min(n - done, wave_words - MT_N);     // a wave serves this many draws whatever mti is
    const bool last = done + m >= n;
    uint32_t* nxt = last ? st_out : ((w & 1) ? tmp1 : tmp0);
    const int nb = (int)((MT_N + m + blk_words - 1) / blk_words);  // upper bound on the blocks touched (mti <= 624)
    if (S) { Launch L(S, SQC_KF_RNG, stream, "seq"); k_mt_seq<<<1, 256, 0, stream>>>(cur, seq, ctl); L.done(); }
    else k_mt_seq<<<1, 256, 0, stream>>>(cur, seq, ctl);
    uint32_t* windows = seq + MT_SEQ_WORDS;                        // (SQC_MT_JUMP_ROWS + 1) x 624 behind the sequence
    if (nb > 1) {
      if (S) { Launch L(S, SQC_KF_RNG, stream, "jump"); k_mt_jump<<<nb - 1, MT_JUMP_THREADS, MT_SMEM, stream>>>(cur, seq, table, windows, m, mult, ctl); L.done(); }
      else k_mt_jump<<<nb - 1, MT_JUMP_THREADS, MT_SMEM, stream>>>(cur, seq, table, windows, m, mult, ctl);
    }
    if (S) { Launch L(S, SQC_KF_RNG, stream, "gen"); k_mt_gen<<<nb, MT_GEN_THREADS, 0, stream>>>(cur, nxt, seq, windows, out + done, m, mult, ctl); L.done(); }
    else k_mt_gen<<<nb, MT_GEN_THREADS, 0, stream>>>(cur, nxt, seq, windows, out + done, m, mult, ctl);
    cur = nxt; done += m; ++w;
  }
}

// Keys of update_beta_scale_k call `call` (every call consumes exactly N_total draws).
// They do not depend on the data, so they are produced on a side stream, up to two calls ahead of the EM
// loop; state_after[c] lives in slot c % 3 so that the state after the last EXECUTED call survives.
void enqueue_keys(sqc_session* S, int call, cudaEvent_t after) {
  if (S->rng_mode != SQC_RNG_R_COMPAT) return;
  const int kb = call & 1;
  // the buffer of call was last read by the M-step of call - 2 (key_free), and the generation is meant to run beside
  // the M-step kernel of call - 1, which leaves room for it (`after` is recorded right before that launch)
  CK(cudaStreamWaitEvent(S->st_rng, S->ev_key_free[kb], 0));
  if (after) CK(cudaStreamWaitEvent(S->st_rng, after, 0));
  uint32_t* base = S->mt_state;
  if (S->key_q) {
    // distributed stream: only this rank's share, from the state enqueue_chain left in slot 5 + (call & 1)
    CK(cudaStreamWaitEvent(S->st_rng, S->ev_chain[kb], 0));
    if (S->key_mine > 0)
      mt_generate(S->st_rng, base + 640 * (5 + kb), base + 640 * 7, base + 640 * 3, base + 640 * 4, S->mt_seq, S->mt_table,
                  S->keybuf + (size_t)kb * S->key_q, S->key_mine, S->dev.ctl, S, S->mt_target_blocks);
  } else
  mt_generate(S->st_rng, base + 640 * ((call + 2) % 3), base + 640 * (call % 3), base + 640 * 3, base + 640 * 4, S->mt_seq, S->mt_table,
              S->keybuf + (size_t)kb * S->N_total, S->N_total, S->dev.ctl, S, S->mt_target_blocks);
  CK(cudaEventRecord(S->ev_key_ready[kb], S->st_rng));
}
// distributed stream: the state in front of `call` (slot (call + 2) % 3) -> the state after it (slot call % 3) and the
// state in front of this rank's share (slot 5 + (call & 1)), by jump polynomials.  Data-independent and tiny (one or
// two CTAs, but ~0.2 ms of latency), so it runs on its own stream one call ahead of the generation.
void enqueue_chain(sqc_session* S, int call, cudaEvent_t after) {
  uint32_t* base = S->mt_state;
  if (after) CK(cudaStreamWaitEvent(S->st_rng2, after, 0));
  const uint32_t* in = base + 640 * ((call + 2) % 3);
  const bool has_rank = S->rank > 0 && S->key_mine > 0;
  { Launch L(S, SQC_KF_RNG, S->st_rng2, "chain_seq"); k_mt_seq<<<1, 256, 0, S->st_rng2>>>(in, S->mt_seq2, S->dev.ctl); L.done(); }
  { Launch L(S, SQC_KF_RNG, S->st_rng2, "chain_jump"); k_mt_jump_rows<<<has_rank ? 2 : 1, MT_JUMP_THREADS, MT_SMEM, S->st_rng2>>>(S->mt_seq2, S->mt_rows, S->mt_jwin, S->dev.ctl); L.done(); }
  { Launch L(S, SQC_KF_RNG, S->st_rng2, "chain_fix"); k_mt_fix<<<1, 256, 0, S->st_rng2>>>(in, S->mt_jwin, base + 640 * (call % 3), base + 640 * (5 + (call & 1)), S->N_total, has_rank ? 1 : 0, S->dev.ctl); L.done(); }
  CK(cudaEventRecord(S->ev_chain[call & 1], S->st_rng2));
}
void enqueue_next_keys(sqc_session* S, int call) {
  if (S->rng_mode != SQC_RNG_R_COMPAT || call + 1 > S->em_iters) return;
  CK(cudaEventRecord(S->ev_pre_mcd, S->st));
  enqueue_keys(S, call + 1, S->ev_pre_mcd);
}
void use_keys(sqc_session* S, int call) {
  if (S->rng_mode == SQC_RNG_R_COMPAT) {
    CK(cudaStreamWaitEvent(S->st, S->ev_key_ready[call & 1], 0));
    if (S->key_q) {
      // every rank's share of this call must be complete before any rank's M-step kernel reads it.  For calls 0 and >= 2 the
      // per-iteration exchange in front of the call was enqueued behind the wait for the local share (run_robust) and is the
      // fence; call 1 has no such exchange in front of it (it follows the initial M-step directly): one tiny exchange of its
      // own.  (The peers' reads end before this kernel's last exchange of the call, so the buffer may be refilled once the
      // kernel is done - the same event that frees it locally.)
      if (call == 1) { Launch L(S, SQC_KF_UPDATE, nullptr, "key_fence"); k_xchg_sum_f64<<<1, 32, 0, S->st>>>(S->xc.dev, XCHG_CH_HOST + 0, S->xg_vec + 8, 1, &S->dev.ctl->cont); L.done(); }
      S->mb.keys = S->keybuf + (size_t)(call & 1) * S->key_q;
      for (int q = 0; q < S->world; ++q) S->mb.key_peer[q] = S->key_peer[q] + (size_t)(call & 1) * S->key_q;
      S->mb.key_q = S->key_q; S->mb.key_qinv = 1.0 / (double)S->key_q;
    } else
    S->mb.keys = S->keybuf + (size_t)(call & 1) * S->N_total;
    // the keys of the NEXT call are generated while this call's M-step kernel runs (released only now: kernels that
    // slip onto the SMs during a wait above would hold up the cooperative launch that follows)
    if (!S->rng_after_mcd) enqueue_next_keys(S, call);
    // ... and so is the state chain of call + 2 (two CTAs on the SMs k_mcd leaves free).  Behind the M-step kernel it ran
    // beside the next partition pass, and that pass then never finished before the chain's 0.2 ms jump kernel did
    // (profiles/r2u_timeline_w2_rank0.csv: partition 87 us against 53-70 us without a chain beside it, at any shard size).
    if (S->key_q && call + 2 <= S->em_iters) {
      CK(cudaEventRecord(S->ev_post_mcd, S->st));
      enqueue_chain(S, call + 2, S->ev_post_mcd);             // its share-state slot was last read by the keys of `call`
    }
  } else S->mb.keys = nullptr;
  S->mb.call_mix = counter_call_mix(S->seed, 0u);
  S->mb.stream_base = (long long)call * S->N_total;
  S->mb.call = call;
}
void after_mcd_keys(sqc_session* S, int call) {
  if (S->rng_mode != SQC_RNG_R_COMPAT) return;
  if (S->rng_after_mcd) enqueue_next_keys(S, call);
}
void release_keys(sqc_session* S, int call) {
  if (S->rng_mode != SQC_RNG_R_COMPAT) return;
  CK(cudaEventRecord(S->ev_key_free[call & 1], S->st));
}

