// sqc_stream.cuh — the two per-cell passes of an EM iteration as PERSISTENT kernels fed by the TMA engine.
//
//   k_estep_stream      like_2 (:506) + calc_expected_z (:344-396, :509) + z_delta (:519) + the change of the
//                       per-(sample, component) statistics update_alpha_j / update_p_jk need (reference
//                       src/fit_sampleQC_robust_cpp.cpp:116-157, :322-341)
//   k_partition_stream  like_1 (:499) + restrict_to_x_k for all k at once (:265-283)
//
// Round 1 ran one CTA per 1024-cell tile: ~10 000 short-lived CTAs whose prologue (tile metadata -> parameters ->
// cells, three dependent L2 / HBM round trips) and epilogue (three block barriers) were not overlapped with
// anything; ncu showed 15 % barrier stalls, 13 % long-scoreboard stalls and 14 % of the samples on kernel-parameter
// loads (profiles/r1k_ncu_summary.md).  Here a CTA owns a contiguous range of tiles and keeps two tiles in
// flight: while it computes tile t out of shared memory, the D column segments of tile t + 1 arrive by bulk
// asynchronous copies (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier).  Address arithmetic, the
// metadata chain and the HBM latency leave the inner loop; one block barrier per tile remains.
//
// Statistics.  update_alpha_j / update_p_jk need n_jk and s_jk = sum of x over the cells of (sample j, component k).
// Round 1 re-accumulated them for every cell in every iteration, masked K times.  The E-step now only applies the
// CHANGE: a cell that moves from component a to b subtracts its x from (tile, a) and adds it to (tile, b).  Warps
// without a changed label skip the whole block (the common case after the first iterations); changed cells are
// applied in (cell slot, lane) order per warp and the warps' contributions in warp order, so the result is
// bit-reproducible.  The initial statistics come from k_estep<D, 1> (one full masked pass per fit).
#pragma once
#include "sqc_kernels.cuh"

namespace sqc {

#ifndef SQC_EST2_MINB
#define SQC_EST2_MINB 3
#endif
#ifndef SQC_PART2_MINB
#define SQC_PART2_MINB 3
#endif

template <int D> struct StreamCfg {
  static constexpr int CPT = TileCfg<D>::CPT, TILE = TileCfg<D>::TILE;
  static constexpr int COLS = TILE + 2;                      // one cell of slack on either side: 16-byte aligned bulk copies
  static constexpr int STAGE = D * COLS;                     // doubles per stage
};

// value idx of the per-(sample, component) density constants (layout of ParK<D>)
template <int D> __device__ __forceinline__ double param_value(const Dev& s, int j, int idx) {
  constexpr int SZ = ParK<D>::SZ;
  const int k = idx / SZ, e = idx % SZ;
  if (e < D) return s.alpha[j * D + e] + s.beta[k * D + e];
  if (e < D + SQC_TRI(D)) {
    int t = e - D, r = 0;
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    const int c = t - r * (r + 1) / 2;
    return s.linv[k * D * D + r + D * c];
  }
  if (e == D + SQC_TRI(D)) return s.logp_jk[j * s.K + k] + s.logP[k];          // log(p_jk) + log-normaliser
  if (e == D + SQC_TRI(D) + 1) return s.p_jk[j * s.K + k] * s.Pk[k];            // p_jk * normaliser
  return 0.0;
}

// thread 0: arm the stage's barrier and issue the D column copies of tile t
template <int D> __device__ __forceinline__ void issue_tile(const Dev& s, int t, double* stage, uint64_t* bar) {
  constexpr int COLS = StreamCfg<D>::COLS;
  const long long start = s.tile_start[t]; const int len = s.tile_len[t];
  long long a0[D]; unsigned nb[D]; unsigned total = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const long long e = (long long)d * s.N + start;
    a0[d] = e & ~1ll;
    nb[d] = (unsigned)((((e - a0[d]) + len + 1) & ~1ll) * 8);
    total += nb[d];
  }
  mbar_expect_tx(bar, total);
#pragma unroll
  for (int d = 0; d < D; ++d) bulk_g2s(stage + d * COLS, s.x + a0[d], nb[d], bar);
}

// ---------------------------------------------------------------------------------------------
// E-step
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(EST_THREADS, (D <= 4) ? SQC_EST2_MINB : 2) k_estep_stream(Dev s) {
  constexpr int CPT = StreamCfg<D>::CPT, COLS = StreamCfg<D>::COLS, STAGE = StreamCfg<D>::STAGE, SZ = ParK<D>::SZ, NW = EST_THREADS / 32;
  extern __shared__ __align__(16) unsigned char sm_raw[];
  double* stg = (double*)sm_raw;                               // 2 stages of D column segments
  double* par = stg + 2 * STAGE;                               // K * SZ density constants of the current sample
  double* dw = par + s.K * SZ;                                 // 2 (tile parity) x NW x K x (D + 1) statistic deltas
  __shared__ uint64_t bar[2];
  __shared__ double sh_w[NW];
  __shared__ int sh_i[NW * 2];
  __shared__ int sh_dirty[2];
  if (!s.ctl->cont) return;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, G = gridDim.x, b = blockIdx.x, K = s.K;
  const int t0 = (int)((long long)s.T * b / G), t1 = (int)((long long)s.T * (b + 1) / G);
  const int KS = K * SZ, KD = K * (D + 1);
  if (t0 >= t1) { if (tid == 0) s.t_ll2[b] = 0.0; return; }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  for (int i = tid; i < 2 * NW * KD; i += EST_THREADS) dw[i] = 0.0;
  if (tid < 2) sh_dirty[tid] = -1;
  int start = s.tile_start[t0], len = s.tile_len[t0], j = s.tile_sample[t0];
  for (int idx = tid; idx < KS; idx += EST_THREADS) par[idx] = param_value<D>(s, j, idx);
  __syncthreads();
  if (tid == 0) { issue_tile<D>(s, t0, stg, &bar[0]); if (t0 + 1 < t1) issue_tile<D>(s, t0 + 1, stg + STAGE, &bar[1]); }
  double ll_acc = 0.0; int delta_acc = 0, nzero_acc = 0;
  unsigned phase = 0u;
  for (int t = t0; t < t1; ++t) {
    const int st = (t - t0) & 1;
    int nstart = 0, nlen = 0, nj = j;                          // metadata of the next tile: in flight during this tile's arithmetic
    if (t + 1 < t1) { nstart = s.tile_start[t + 1]; nlen = s.tile_len[t + 1]; nj = s.tile_sample[t + 1]; }
    int zo[CPT]; bool valid[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int i = c * EST_THREADS + tid;
      valid[c] = i < len;
      zo[c] = (int)ld_stream_u8(s.z + (long long)start + (valid[c] ? i : 0));
    }
    mbar_wait(&bar[st], (phase >> st) & 1u); phase ^= (1u << st);
    const double* sg = stg + st * STAGE;
    double xv[CPT][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int off = (int)(((long long)d * s.N + start) & 1ll);
#pragma unroll
      for (int c = 0; c < CPT; ++c) xv[c][d] = sg[d * COLS + off + c * EST_THREADS + tid];
    }
    double best[CPT], like[CPT]; int zn[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) { best[c] = -CUDART_INF; like[c] = 0.0; zn[c] = 0; }
    int nzero = 0;
    for (int k = 0; k < K; ++k) {
      const double* pk = par + k * SZ;
      double m[D], lt[SQC_TRI(D)];
#pragma unroll
      for (int d = 0; d < D; ++d) m[d] = pk[d];
#pragma unroll
      for (int e = 0; e < SQC_TRI(D); ++e) lt[e] = pk[D + e];
      const double lc = pk[D + SQC_TRI(D)], pc = pk[D + SQC_TRI(D) + 1];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        double v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = xv[c][d] - m[d];
        const double q = maha_tri<D>(v, lt);
        const double l = fma(-0.5, q, lc);                     // log(p_jk) + dmvnorm(log = TRUE), :364
        if (l > best[c]) { best[c] = l; zn[c] = k; }           // strict >, first max, :375-378
        if (valid[c] && best[c] == -CUDART_INF) ++nzero;       // :380-382
        like[c] = fma(pc, exp_neg(-0.5 * q), like[c]);         // p_jk * dmvnorm(log = FALSE), :424
      }
    }
    ll_acc += sum_log<CPT>(like, valid);                       // :428
    nzero_acc += nzero;
    bool chg[CPT]; bool anyc = false;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      chg[c] = valid[c] && zn[c] != zo[c];
      anyc = anyc || chg[c];
      if (chg[c]) { s.z[(long long)start + c * EST_THREADS + tid] = (uint8_t)zn[c]; ++delta_acc; }
    }
    if (__any_sync(0xffffffffu, anyc)) {
      // changed cells leave (tile, old label) and join (tile, new label): applied in (cell slot, lane) order
      double* mydw = dw + (size_t)(st * NW + w) * KD;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        unsigned mk = __ballot_sync(0xffffffffu, chg[c]);
        while (mk) {
          const int l = __ffs(mk) - 1; mk &= mk - 1;
          if (lane == l) {
            double* a = mydw + zo[c] * (D + 1); double* bb = mydw + zn[c] * (D + 1);
            a[0] -= 1.0; bb[0] += 1.0;
#pragma unroll
            for (int d = 0; d < D; ++d) { a[1 + d] -= xv[c][d]; bb[1 + d] += xv[c][d]; }
          }
          __syncwarp();
        }
      }
      if (lane == 0) sh_dirty[st] = t;
    }
    __syncthreads();                                           // the stage is free; the warps' deltas are visible
    if (tid == 0 && t + 2 < t1) issue_tile<D>(s, t + 2, stg + st * STAGE, &bar[st]);
    if (sh_dirty[st] == t) {
      for (int idx = tid; idx < KD; idx += EST_THREADS) {
        double acc = 0.0;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) { double* p = dw + (size_t)(st * NW + ww) * KD + idx; acc += *p; *p = 0.0; }
        const int k = idx / (D + 1), e = idx % (D + 1);
        if (e == 0) s.t_nk[t * K + k] += (int)acc;
        else s.t_sk[((long long)t * K + k) * D + (e - 1)] += acc;
      }
    }
    if (nj != j) {                                             // (uniform) next tile belongs to another sample: new constants
      for (int idx = tid; idx < KS; idx += EST_THREADS) par[idx] = param_value<D>(s, nj, idx);
      __syncthreads();
    }
    start = nstart; len = nlen; j = nj;
  }
  // per-CTA sums, fixed order
  ll_acc = warp_sum(ll_acc); delta_acc = warp_sum_i(delta_acc); nzero_acc = warp_sum_i(nzero_acc);
  if (lane == 0) { sh_w[w] = ll_acc; sh_i[w] = delta_acc; sh_i[NW + w] = nzero_acc; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0; int dl = 0, nz = 0;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) { a += sh_w[ww]; dl += sh_i[ww]; nz += sh_i[NW + ww]; }
    s.t_ll2[b] = a;
    if (dl) atomicAdd(&s.ctl->z_delta, dl);
    if (nz) atomicAdd(&s.ctl->n_zero, nz);
  }
}

// ---------------------------------------------------------------------------------------------
// like_1 + stable partition of the residuals by label
// ---------------------------------------------------------------------------------------------
template <int D, int LIKE>
__global__ void __launch_bounds__(EST_THREADS, (D <= 4) ? SQC_PART2_MINB : 2) k_partition_stream(Dev s) {
  constexpr int CPT = StreamCfg<D>::CPT, COLS = StreamCfg<D>::COLS, STAGE = StreamCfg<D>::STAGE, SZ = ParK<D>::SZ, NW = EST_THREADS / 32;
  static_assert(CPT * NW <= 32, "one lane per (cell slot, warp) count in the prefix");
  extern __shared__ __align__(16) unsigned char sm_raw[];
  double* stg = (double*)sm_raw;
  double* par = stg + 2 * STAGE;                               // K * SZ
  double* sh_aj = par + s.K * SZ;                              // D: alpha_j of the current sample
  long long* sh_base = (long long*)(sh_aj + SQC_MAXD);         // K: first position of this tile's cells inside the cluster segments
  int* sh_cnt = (int*)(sh_base + SQC_MAXK);                    // CPT * NW * K counts per (cell slot, warp, component)
  __shared__ uint64_t bar[2];
  __shared__ double sh_w[NW];
  if (!s.ctl->cont) return;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, G = gridDim.x, b = blockIdx.x, K = s.K;
  if (b == 0 && tid < 2 * SQC_MAXK + 1) s.mcd_sync[tid] = 0u;      // epochs, arrivals and status word of the k_mcd launch that follows
  const int t0 = (int)((long long)s.T * b / G), t1 = (int)((long long)s.T * (b + 1) / G);
  const int KS = K * SZ;
  if (t0 >= t1) { if (LIKE && tid == 0) s.t_ll1[b] = 0.0; return; }
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  int start = s.tile_start[t0], len = s.tile_len[t0], j = s.tile_sample[t0];
  if (LIKE) for (int idx = tid; idx < KS; idx += EST_THREADS) par[idx] = param_value<D>(s, j, idx);
  if (tid < D) sh_aj[tid] = s.alpha[j * D + tid];
  __syncthreads();
  if (tid == 0) { issue_tile<D>(s, t0, stg, &bar[0]); if (t0 + 1 < t1) issue_tile<D>(s, t0 + 1, stg + STAGE, &bar[1]); }
  double ll_acc = 0.0;
  unsigned phase = 0u;
  for (int t = t0; t < t1; ++t) {
    const int st = (t - t0) & 1;
    int nstart = 0, nlen = 0, nj = j;
    if (t + 1 < t1) { nstart = s.tile_start[t + 1]; nlen = s.tile_len[t + 1]; nj = s.tile_sample[t + 1]; }
    // base of this tile inside every cluster segment: cells of the earlier samples (s_base) + of the earlier tiles of this sample (t_rel)
    if (tid < K) sh_base[tid] = s.clu_off[tid] + (long long)s.s_base[j * K + tid] + (long long)s.t_rel[t * K + tid];
    int zo[CPT]; bool valid[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int i = c * EST_THREADS + tid;
      valid[c] = i < len;
      zo[c] = valid[c] ? (int)ld_stream_u8(s.z + (long long)start + i) : -1;
    }
    // rank of every cell among the cells of its component inside (cell slot, warp); counts per (slot, warp, component)
    int myrank[CPT];
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const unsigned bal = __ballot_sync(0xffffffffu, zo[c] == k);
        if (zo[c] == k) myrank[c] = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) sh_cnt[(c * NW + w) * K + k] = __popc(bal);
      }
    }
    mbar_wait(&bar[st], (phase >> st) & 1u); phase ^= (1u << st);
    const double* sg = stg + st * STAGE;
    double xv[CPT][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const int off = (int)(((long long)d * s.N + start) & 1ll);
#pragma unroll
      for (int c = 0; c < CPT; ++c) xv[c][d] = sg[d * COLS + off + c * EST_THREADS + tid];
    }
    __syncthreads();                                           // counts, bases visible
    // exclusive prefix over the 32 (slot, warp) entries in tile order, per component: every warp scans for itself
    // (lane e holds entry e), a cell then fetches the prefix of its own entry for its own component
    int pre[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) pre[c] = 0;
    for (int k = 0; k < K; ++k) {
      const int v = (lane < CPT * NW) ? sh_cnt[lane * K + k] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      const int exc = inc - v;
#pragma unroll
      for (int c = 0; c < CPT; ++c) { const int got = __shfl_sync(0xffffffffu, exc, c * NW + w); if (zo[c] == k) pre[c] = got; }
    }
    double aj[D];
#pragma unroll
    for (int d = 0; d < D; ++d) aj[d] = sh_aj[d];
#pragma unroll
    for (int c = 0; c < CPT; ++c) if (valid[c]) {
      const long long pos = sh_base[zo[c]] + pre[c] + myrank[c];
#pragma unroll
      for (int d = 0; d < D; ++d) s.rs[(long long)d * s.rs_stride + pos] = xv[c][d] - aj[d];   // :274
    }
    if (LIKE) {
      double like[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) like[c] = 0.0;
      for (int k = 0; k < K; ++k) {
        const double* pk = par + k * SZ;
        double m[D], lt[SQC_TRI(D)];
#pragma unroll
        for (int d = 0; d < D; ++d) m[d] = pk[d];
#pragma unroll
        for (int e = 0; e < SQC_TRI(D); ++e) lt[e] = pk[D + e];
        const double pc = pk[D + SQC_TRI(D) + 1];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          double v[D];
#pragma unroll
          for (int d = 0; d < D; ++d) v[d] = xv[c][d] - m[d];
          like[c] = fma(pc, exp_neg(-0.5 * maha_tri<D>(v, lt)), like[c]);
        }
      }
      ll_acc += sum_log<CPT>(like, valid);
    }
    __syncthreads();                                           // stage, counts, bases, constants free
    if (tid == 0 && t + 2 < t1) issue_tile<D>(s, t + 2, stg + st * STAGE, &bar[st]);
    if (nj != j) {
      if (LIKE) for (int idx = tid; idx < KS; idx += EST_THREADS) par[idx] = param_value<D>(s, nj, idx);
      if (tid < D) sh_aj[tid] = s.alpha[nj * D + tid];
    }
    start = nstart; len = nlen; j = nj;
  }
  if (LIKE) {
    ll_acc = warp_sum(ll_acc);
    if (lane == 0) sh_w[w] = ll_acc;
    __syncthreads();
    if (tid == 0) { double a = 0.0; for (int ww = 0; ww < NW; ++ww) a += sh_w[ww]; s.t_ll1[b] = a; }
  }
}

template <int D> inline size_t estep_stream_smem(int K) {
  return (size_t)(2 * StreamCfg<D>::STAGE + K * ParK<D>::SZ + 2 * (EST_THREADS / 32) * K * (D + 1)) * sizeof(double);
}
template <int D> inline size_t partition_stream_smem(int K) {
  return (size_t)(2 * StreamCfg<D>::STAGE + K * ParK<D>::SZ + SQC_MAXD) * sizeof(double) + (size_t)SQC_MAXK * sizeof(long long) +
         (size_t)StreamCfg<D>::CPT * (EST_THREADS / 32) * K * sizeof(int);
}

}  // namespace sqc
