// sqc_mcd.cuh — the M-step of the robust fit as ONE persistent cooperative kernel.
//
// Replaces update_beta_scale_k (reference src/fit_sampleQC_robust_cpp.cpp:289-318) and fast_mcd
// (:55-108) for all K clusters in lock step.  Input: the cluster-sorted residual columns `rs`
// (r = x - alpha_j, stable partition by label = restrict_to_x_k :265-283 for every k at once).
//
// The reference's arma::sort_index over n_k distances (:50) is replaced by an EXACT order-statistic
// selection in two streaming rounds:
//   HIST     every cell's key (random start key, or squared Mahalanobis distance as its IEEE bit
//            pattern — monotone for d >= 0) is binned inside a window [wlo, whi) of the sortable bit
//            space; cells under the window are only counted.
//   COLLECT  cells strictly under the bin that holds the h-th smallest key are accumulated into
//            shifted moments (count, sum, upper-triangular scatter); the cells of that one bin are
//            appended to a small candidate list, which the cluster's leader CTA sorts by (key, position)
//            to take exactly the remaining members — ties resolved towards the lower position, the
//            repo's definition of the reference's unspecified tie order.
// A cluster walks through  KEY_HIST -> KEY_COLLECT -> [DIST_HIST -> DIST_COLLECT]* -> final
// DIST_HIST -> DIST_COLLECT(q only) -> DONE; all clusters share every streaming round.
// Between rounds: grid.sync, the K leader CTAs do the tiny per-cluster post-processing
// (histogram scan / candidate sort / moments -> loc, scale, det, Cholesky, loop test :84), grid.sync.
#pragma once
#include "sqc_kernels.cuh"

namespace sqc {

constexpr int MCD_THREADS = 512;                     // one fat CTA per SM: one histogram per SM, 128 registers per thread
constexpr int MCD_CPT = 8;                           // cells per thread per chunk (four 128-bit loads per column)
constexpr int MCD_CHUNK = MCD_THREADS * MCD_CPT;
constexpr int MCD_NB = 4096;                         // histogram bins per window
constexpr int MCD_CAND_CAP = 8192;                   // candidates per cluster (bitonic-sorted in smem)
constexpr size_t MCD_SMEM = (size_t)MCD_CAND_CAP * 12 + 1024;   // sort keys (u64) + positions (u32)

enum { PH_KEY_HIST = 0, PH_KEY_COLLECT = 1, PH_DIST_HIST = 2, PH_DIST_COLLECT = 3, PH_DONE = 4 };

struct McdClu {
  long long n, h, off, prefix;         // cells, subset size, segment offset in rs, key-stream offset
  long long below;                     // cells strictly under the candidate bin (set by the HIST post)
  unsigned long long wlo, whi;         // histogram window (sortable bits)
  unsigned long long blo, bhi;         // candidate bin (sortable bits)
  unsigned long long t_prev;           // last resolved threshold (distance bits), 0 = none
  int shift, phase, is_final, iters, n_hist_rounds, pad;
  double det_old, det_new, q_chisq;
  double cshift[SQC_MAXD];             // moment shift
  double loc[SQC_MAXD];
  double scale[SQC_MAXD * SQC_MAXD];
  double linv[SQC_TRI(SQC_MAXD)];      // packed lower-triangular inverse Cholesky factor (row-major)
};

struct McdWork {
  McdClu clu[SQC_MAXK];
  unsigned long long cnt_below[SQC_MAXK];   // HIST: cells under the window
  unsigned long long cnt_nan[SQC_MAXK];
  unsigned int cand_cnt[SQC_MAXK];
};

struct GridBar { unsigned int count; unsigned int gen; };

struct McdBufs {
  McdWork* work;
  unsigned int* hist;                  // K x MCD_NB
  double* part;                        // grid x K x NMOM(D) per-CTA moment partials
  unsigned long long* cand_bits;       // K x MCD_CAND_CAP
  unsigned int* cand_pos;              // K x MCD_CAND_CAP
  const int* keys;                     // R-compatible key stream of this call (N_total ints), or null
  unsigned long long call_mix;         // counter-mode key mix for this call
  long long stream_base;               // call * N_total: index of this call's first draw
  GridBar* bar;                        // grid barrier words (count must be 0 at launch)
  long long* trace;                    // optional per-round timeline (debug): 8 words per round, word 0 of the buffer = rounds
};
// Sense-reversing grid barrier on two words of global memory (the launch is cooperative, so all CTAs are
// co-resident).  About half the latency of cooperative_groups' grid.sync() for one CTA per SM.
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void grid_barrier(GridBar* bar, unsigned int nblocks, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&bar->count, 1u);
    if (prev == nblocks - 1u) { atomicExch(&bar->count, 0u); __threadfence(); atomicAdd(&bar->gen, 1u); }
    else { while (ld_acquire_u32(&bar->gen) == gen) { } }
    __threadfence();
  }
  ++gen;
  __syncthreads();
}
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ unsigned long long dist_bits(double d) { return (unsigned long long)__double_as_longlong(d); }

// window helpers: choose the smallest shift so that (whi - wlo) >> shift <= MCD_NB
__device__ inline int window_shift(unsigned long long wlo, unsigned long long whi) {
  unsigned long long span = whi - wlo;
  int s = 0;
  while (s < 63 && ((span + ((1ull << s) - 1)) >> s) > (unsigned long long)MCD_NB) ++s;
  return s;
}
__device__ inline void set_dist_window(McdClu& c, double lo, double hi) {
  c.wlo = dist_bits(lo); c.whi = dist_bits(hi);
  c.shift = window_shift(c.wlo, c.whi);
}
__device__ inline void set_full_dist_window(McdClu& c) {       // every finite non-negative double
  c.wlo = 0ull; c.whi = 0x7ff0000000000000ull;
  c.shift = window_shift(c.wlo, c.whi);
}

// in-block bitonic sort of (key, pos) pairs, ascending, lexicographic
__device__ inline void bitonic_sort_pairs(unsigned long long* key, unsigned int* pos, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const unsigned long long ka = key[lo], kb = key[hi];
        const unsigned int pa = pos[lo], pb = pos[hi];
        const bool gt = (ka > kb) || (ka == kb && pa > pb);
        if (gt == asc) { key[lo] = kb; key[hi] = ka; pos[lo] = pb; pos[hi] = pa; }
      }
      __syncthreads();
    }
  }
}

// deterministic block reduction of NM doubles per thread -> out[0..NM) (valid for all threads after return)
template <int NM>
__device__ inline void block_reduce_vec(double (&v)[NM], double* sh /* 32 * NM */, double* out /* NM, shared */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    double a = warp_sum(v[m]);
    if (lane == 0) sh[w * NM + m] = a;
  }
  __syncthreads();
  for (int m = threadIdx.x; m < NM; m += blockDim.x) {
    double a = 0.0;
    for (int ww = 0; ww < nw; ++ww) a += sh[ww * NM + m];
    out[m] = a;
  }
  __syncthreads();
}

template <int D> struct Mom {
  static constexpr int NM = SQC_NMOM(D);
  // accumulate w = r - c into count / sum / upper-triangular scatter
  __device__ static __forceinline__ void add(double (&acc)[NM], const double (&r)[D], const double (&c)[D], bool in) {
    double w[D];
#pragma unroll
    for (int d = 0; d < D; ++d) w[d] = in ? (r[d] - c[d]) : 0.0;
    acc[0] += in ? 1.0 : 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) acc[1 + d] += w[d];
    int e = 1 + D;
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int b = a; b < D; ++b) { acc[e] = fma(w[a], w[b], acc[e]); ++e; }
  }
};

// ---------------------------------------------------------------------------------------------
// leader-side finalisation of a moment set: loc = c + S1/h, scale = S2/h - m m' (biased 1/h, :36-48)
// ---------------------------------------------------------------------------------------------
template <int D>
__device__ inline void moments_to_loc_scale(const double* S, double h, const double* cshift, double* loc, double* scale) {
  double m[D];
  for (int d = 0; d < D; ++d) { m[d] = S[1 + d] / h; loc[d] = cshift[d] + m[d]; }
  int e = 1 + D;
  for (int a = 0; a < D; ++a)
    for (int b = a; b < D; ++b) {
      const double v = S[e++] / h - m[a] * m[b];
      scale[a + D * b] = v; scale[b + D * a] = v;
    }
}
// Cholesky of the scale and its packed inverse (calc_maha_dists :17-21); false -> chol() throws in the reference
template <int D>
__device__ inline bool prepare_dist(McdClu& c) {
  double m[D * D], L[D * D], Li[D * D];
#pragma unroll
  for (int e = 0; e < D * D; ++e) { m[e] = c.scale[e]; L[e] = 0.0; Li[e] = 0.0; }
  bool ok = true;
#pragma unroll
  for (int cc = 0; cc < D; ++cc) {
    double sdiag = m[cc + D * cc];
#pragma unroll
    for (int k = 0; k < cc; ++k) sdiag -= L[cc + D * k] * L[cc + D * k];
    if (!(sdiag > 0.0) || !(sdiag < CUDART_INF)) ok = false;
    const double dd = sqrt(sdiag); L[cc + D * cc] = dd;
#pragma unroll
    for (int r = cc + 1; r < D; ++r) {
      double t = m[r + D * cc];
#pragma unroll
      for (int k = 0; k < cc; ++k) t -= L[r + D * k] * L[cc + D * k];
      L[r + D * cc] = t / dd;
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int cc = 0; cc < D; ++cc) {
    Li[cc + D * cc] = 1.0 / L[cc + D * cc];
#pragma unroll
    for (int r = cc + 1; r < D; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = cc; k < r; ++k) t -= L[r + D * k] * Li[k + D * cc];
      Li[r + D * cc] = t / L[r + D * r];
    }
  }
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int cc = 0; cc <= r; ++cc) c.linv[r * (r + 1) / 2 + cc] = Li[r + D * cc];
#pragma unroll
  for (int d = 0; d < D; ++d) c.cshift[d] = c.loc[d];
  return true;
}

// ---------------------------------------------------------------------------------------------
// the persistent kernel
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(MCD_THREADS, 1) k_mcd(Dev s, McdBufs mb) {
  constexpr int NM = SQC_NMOM(D);
  extern __shared__ __align__(16) unsigned char sh_raw[];
  unsigned long long* sh_key = (unsigned long long*)sh_raw;                  // MCD_CAND_CAP (leader sort)
  unsigned int* sh_pos = (unsigned int*)(sh_raw + (size_t)MCD_CAND_CAP * 8); // MCD_CAND_CAP
  unsigned int* sh_hist = (unsigned int*)sh_raw;                             // MCD_NB (streaming rounds), aliases sh_key
  __shared__ double sh_red[32 * NM];
  __shared__ double sh_out[NM];
  __shared__ long long sh_cstart[SQC_MAXK + 1];
  __shared__ long long sh_scan[33];
  __shared__ int sh_flag[4];
  __shared__ int sh_ph[SQC_MAXK];
  __shared__ long long sh_n[SQC_MAXK];
  __shared__ unsigned long long sh_u64[2];

  if (!s.ctl->cont) return;                    // uniform across the grid: nobody reaches a barrier
  const int tid = threadIdx.x, lane = tid & 31, G = gridDim.x, b = blockIdx.x, K = s.K;
  unsigned int bar_gen = ld_acquire_u32(&mb.bar->gen);
  McdWork* W = mb.work;

  // ---- setup: one leader CTA per cluster ----------------------------------------------------
  if (b < K) {
    const int k = b;
    if (tid == 0) {
      McdClu& c = W->clu[k];
      const long long n = s.clu_n_glob[k];
      c.n = n; c.off = s.clu_off[k]; c.prefix = s.clu_prefix[k];
      c.h = (long long)(int)((double)(unsigned long long)(n + D + 1) * s.mcd_alpha);   // :306 truncation
      c.iters = 0; c.is_final = 0; c.t_prev = 0ull; c.n_hist_rounds = 0; c.below = 0;
      c.det_old = 0.0; c.det_new = 0.0;
      int err = 0;
      if (n <= 0) err = SQC_D_EMPTY_CLUSTER;                       // x_k.rows(0, -1), :307
      else if (c.h > n || c.h < 1) err = SQC_D_SUBSET_SIZE;        // randperm(N, h) with h > N, :72
      if (err) { atomicCAS(&s.ctl->status, 0, err); c.phase = PH_DONE; }
      else {
        c.q_chisq = dev_qchisq((double)c.h / (double)n, D);        // :101-102
        for (int d = 0; d < D; ++d) c.cshift[d] = s.beta[k * D + d];   // previous beta_k (zeros before the first call)
        // key window: the h-th smallest of n uniform keys sits at (h/n) 2^31 +- a few sigma
        const double p = (double)c.h / (double)n, two31 = 2147483648.0;
        const double sig = sqrt(p * (1.0 - p) / (double)n) * two31;
        double lo = p * two31 - 8.0 * sig - 2.0, hi = p * two31 + 8.0 * sig + 2.0;
        if (n <= 65536 || lo < 0.0) lo = 0.0;
        if (n <= 65536 || hi > two31) hi = two31;
        c.wlo = (unsigned long long)lo; c.whi = (unsigned long long)hi;
        c.shift = window_shift(c.wlo, c.whi);
        c.phase = PH_KEY_HIST;
      }
      W->cnt_below[k] = 0ull; W->cnt_nan[k] = 0ull; W->cand_cnt[k] = 0u;
    }
    for (int i = tid; i < MCD_NB; i += MCD_THREADS) mb.hist[k * MCD_NB + i] = 0u;
  }
  __threadfence();
  grid_barrier(mb.bar, (unsigned)G, bar_gen);

  long long rounds = 0;
  for (;;) {
    // ---- active chunk map -------------------------------------------------------------------
    if (tid < K) { sh_ph[tid] = W->clu[tid].phase; sh_n[tid] = W->clu[tid].n; }      // parallel loads, then a tiny serial prefix
    if (tid == 32) sh_flag[3] = s.ctl->status;
    __syncthreads();
    if (tid == 0) {
      long long run = 0, cells = 0, bytes = 0; int any = 0;
      const bool bad = sh_flag[3] != 0;
      for (int k = 0; k < K; ++k) {
        sh_cstart[k] = run;
        const int ph = sh_ph[k];
        if (ph != PH_DONE && !bad) {
          const long long n = sh_n[k];
          run += (n + MCD_CHUNK - 1) / MCD_CHUNK; any = 1; cells += n;
          const bool keyr = (ph == PH_KEY_HIST || ph == PH_KEY_COLLECT);
          bytes += n * ((ph == PH_KEY_HIST ? 0 : 8 * D) + ((keyr && mb.keys) ? 4 : 0));
        }
      }
      sh_cstart[K] = run; sh_flag[0] = any;
      if (b == 0) { s.ctl->mcd_cells += cells; s.ctl->mcd_bytes += bytes; }
    }
    __syncthreads();
    if (!sh_flag[0]) break;
    ++rounds;
    const bool tr = mb.trace && b == 0 && tid == 0 && rounds < 500;
    long long* trow = mb.trace + 8 * rounds;
    if (tr) { trow[0] = gtimer(); int ph = 0; for (int k = 0; k < K; ++k) ph |= (W->clu[k].phase & 7) << (3 * k); trow[4] = ph; trow[5] = sh_cstart[K]; }
    for (int i = tid; i < K * NM; i += MCD_THREADS) mb.part[((long long)b * K) * NM + i] = 0.0;
    const long long total = sh_cstart[K];
    const long long c0 = total * b / G, c1 = total * (b + 1) / G;

    // ---- streaming round ----------------------------------------------------------------------
    long long c = c0;
    while (c < c1) {
      int k = 0;
      while (k + 1 < K && sh_cstart[k + 1] <= c) ++k;
      const long long cend = (sh_cstart[k + 1] < c1) ? sh_cstart[k + 1] : c1;
      const McdClu& cl = W->clu[k];
      const int phase = cl.phase;
      const bool hist_round = (phase == PH_KEY_HIST || phase == PH_DIST_HIST);
      const bool key_round = (phase == PH_KEY_HIST || phase == PH_KEY_COLLECT);
      const long long n = cl.n;
      const unsigned long long lo_b = hist_round ? cl.wlo : cl.blo, hi_b = hist_round ? cl.whi : cl.bhi;
      const int shift = cl.shift;
      double loc[D], cs[D], lt[SQC_TRI(D)];
#pragma unroll
      for (int d = 0; d < D; ++d) { loc[d] = cl.loc[d]; cs[d] = cl.cshift[d]; }
#pragma unroll
      for (int e = 0; e < SQC_TRI(D); ++e) lt[e] = cl.linv[e];
      const bool need_mom = (phase == PH_KEY_COLLECT) || (phase == PH_DIST_COLLECT && !cl.is_final);
      const double* rsb = s.rs + cl.off;
      const long long kbase = cl.prefix;                            // key-stream offset of the cluster's first cell
      if (hist_round) { for (int i = tid; i < MCD_NB; i += MCD_THREADS) sh_hist[i] = 0u; }
      __syncthreads();
      double acc[NM];
#pragma unroll
      for (int m = 0; m < NM; ++m) acc[m] = 0.0;
      unsigned int below = 0, nnan = 0;

      for (long long ch = c; ch < cend; ++ch) {
        const long long e0 = (ch - sh_cstart[k]) * MCD_CHUNK;
        double r[MCD_CPT][D]; unsigned long long bits[MCD_CPT]; bool valid[MCD_CPT];
        long long epos[MCD_CPT];
#pragma unroll
        for (int it = 0; it < MCD_CPT / 2; ++it) {
          const long long e = e0 + (long long)it * (2 * MCD_THREADS) + 2 * tid;
          epos[2 * it] = e; epos[2 * it + 1] = e + 1;
          valid[2 * it] = e < n; valid[2 * it + 1] = (e + 1) < n;
        }
        if (!(key_round && hist_round)) {                           // KEY_HIST needs no residuals
#pragma unroll
          for (int it = 0; it < MCD_CPT / 2; ++it) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              double2 v = make_double2(0.0, 0.0);
              if (valid[2 * it]) v = ld_stream2(rsb + (long long)d * s.rs_stride + epos[2 * it]);
              r[2 * it][d] = v.x; r[2 * it + 1][d] = v.y;
            }
          }
        }
        if (key_round) {
          if (mb.keys) {
#pragma unroll
            for (int it = 0; it < MCD_CPT; ++it) {
              int key = 0;
              if (valid[it]) key = ld_stream_s32(mb.keys + kbase + epos[it]);
              bits[it] = (unsigned long long)(unsigned int)key;
            }
          } else {
#pragma unroll
            for (int it = 0; it < MCD_CPT; ++it)
              bits[it] = (unsigned long long)(unsigned int)counter_key(mb.call_mix, (unsigned long long)(mb.stream_base + kbase + epos[it]));
          }
        } else {
#pragma unroll
          for (int it = 0; it < MCD_CPT; ++it) {
            double v[D];
#pragma unroll
            for (int d = 0; d < D; ++d) v[d] = r[it][d] - loc[d];
            const double dd = maha_tri<D>(v, lt);
            bits[it] = dist_bits(dd);
            if (valid[it] && !(dd == dd)) ++nnan;
          }
        }
        if (hist_round) {
#pragma unroll
          for (int it = 0; it < MCD_CPT; ++it) {
            if (valid[it]) {
              if (bits[it] < lo_b) ++below;
              else if (bits[it] < hi_b) atomicAdd(&sh_hist[(unsigned int)((bits[it] - lo_b) >> shift)], 1u);
            }
          }
        } else {
#pragma unroll
          for (int it = 0; it < MCD_CPT; ++it) {
            const bool in = valid[it] && bits[it] < lo_b;
            const bool cand = valid[it] && bits[it] >= lo_b && bits[it] < hi_b;
            if (need_mom) Mom<D>::add(acc, r[it], cs, in);
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (m) {
              unsigned base = 0;
              if (lane == (__ffs(m) - 1)) base = atomicAdd(&W->cand_cnt[k], (unsigned)__popc(m));
              base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
              if (cand) {
                const unsigned slot = base + __popc(m & ((1u << lane) - 1u));
                if (slot < (unsigned)MCD_CAND_CAP) {
                  mb.cand_bits[(long long)k * MCD_CAND_CAP + slot] = bits[it];
                  mb.cand_pos[(long long)k * MCD_CAND_CAP + slot] = (unsigned int)epos[it];
                }
              }
            }
          }
        }
      }
      // ---- flush this CTA's share of cluster k ------------------------------------------------
      if (hist_round) {
        __syncthreads();
        for (int i = tid; i < MCD_NB; i += MCD_THREADS) { const unsigned v = sh_hist[i]; if (v) atomicAdd(&mb.hist[k * MCD_NB + i], v); }
        unsigned long long bl = (unsigned long long)warp_sum_i((int)below);
        if (lane == 0 && bl) atomicAdd(&W->cnt_below[k], bl);
        __syncthreads();
      } else if (need_mom) {
        block_reduce_vec<NM>(acc, sh_red, sh_out);
        if (tid < NM) mb.part[((long long)b * K + k) * NM + tid] = sh_out[tid];
        __syncthreads();
      }
      const int nn = warp_sum_i((int)nnan);
      if (lane == 0 && nn) atomicAdd(&W->cnt_nan[k], (unsigned long long)nn);
      c = cend;
    }
    if (tr) trow[1] = gtimer();
    __threadfence();
    grid_barrier(mb.bar, (unsigned)G, bar_gen);
    if (tr) trow[2] = gtimer();

    // ---- leaders: per-cluster post-processing -------------------------------------------------
    if (tid == 0) sh_flag[2] = (b < K && W->clu[b].phase != PH_DONE && s.ctl->status == 0) ? 1 : 0;
    __syncthreads();
    if (sh_flag[2]) {
      const int k = b;
      McdClu& cl = W->clu[k];
      const int phase = cl.phase;
      if (W->cnt_nan[k]) {                                          // sort_index(): detected NaN, :50
        if (tid == 0) { atomicCAS(&s.ctl->status, 0, SQC_D_NAN); cl.phase = PH_DONE; }
      } else if (phase == PH_KEY_HIST || phase == PH_DIST_HIST) {
        // locate the bin that holds the h-th smallest key
        const long long below = (long long)W->cnt_below[k];
        constexpr int PER = MCD_NB / MCD_THREADS;
        unsigned int hv[PER]; long long mine = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { hv[i] = mb.hist[k * MCD_NB + tid * PER + i]; mine += hv[i]; }
        long long total_in;
        long long run = block_excl_scan<MCD_THREADS>(mine, sh_scan, &total_in);
        const long long want = cl.h - below;                        // 1-based rank inside the window
        if (tid == 0) { sh_flag[1] = -1; }
        __syncthreads();
        if (want >= 1 && want <= total_in) {
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            if (run < want && want <= run + (long long)hv[i]) { sh_flag[1] = tid * PER + i; sh_u64[0] = (unsigned long long)run; sh_u64[1] = hv[i]; }
            run += hv[i];
          }
        }
        __syncthreads();
        if (tid == 0) {
          cl.n_hist_rounds++;
          const int bin = sh_flag[1];
          if (bin < 0) {
            // the order statistic lies outside the window: widen to the whole side it is on
            if (phase == PH_KEY_HIST) {
              if (want < 1) { cl.whi = cl.wlo; cl.wlo = 0ull; } else { cl.wlo = cl.whi; cl.whi = 1ull << 31; }
            } else {
              if (want < 1) { cl.whi = cl.wlo; cl.wlo = 0ull; } else { cl.wlo = cl.whi; cl.whi = 0x7ff0000000000001ull; }
            }
            cl.shift = window_shift(cl.wlo, cl.whi);
            if (cl.whi <= cl.wlo || cl.n_hist_rounds > 40) { atomicCAS(&s.ctl->status, 0, SQC_D_SELECT); cl.phase = PH_DONE; }
          } else {
            const unsigned long long blo = cl.wlo + ((unsigned long long)bin << cl.shift);
            unsigned long long bhi = blo + (1ull << cl.shift);
            if (bhi > cl.whi) bhi = cl.whi;
            if (sh_u64[1] > (unsigned long long)MCD_CAND_CAP) {
              if (cl.shift == 0 || cl.n_hist_rounds > 40) { atomicCAS(&s.ctl->status, 0, SQC_D_SELECT); cl.phase = PH_DONE; }
              else { cl.wlo = blo; cl.whi = bhi; cl.shift = window_shift(blo, bhi); }       // refine inside the bin
            } else {
              cl.blo = blo; cl.bhi = bhi; cl.below = below + (long long)sh_u64[0];
              cl.phase = (phase == PH_KEY_HIST) ? PH_KEY_COLLECT : PH_DIST_COLLECT;
            }
          }
          W->cnt_below[k] = 0ull; W->cand_cnt[k] = 0u;
        }
        for (int i = tid; i < MCD_NB; i += MCD_THREADS) mb.hist[k * MCD_NB + i] = 0u;
      } else {
        // COLLECT post: exact members of the candidate bin, then the moments
        const int ncand = (int)min(W->cand_cnt[k], (unsigned)MCD_CAND_CAP);
        int P = 32; while (P < ncand) P <<= 1;
        for (int i = tid; i < P; i += MCD_THREADS) {
          if (i < ncand) { sh_key[i] = mb.cand_bits[(long long)k * MCD_CAND_CAP + i]; sh_pos[i] = mb.cand_pos[(long long)k * MCD_CAND_CAP + i]; }
          else { sh_key[i] = ~0ull; sh_pos[i] = ~0u; }
        }
        __syncthreads();
        if (tr) trow[6] = gtimer();
        bitonic_sort_pairs(sh_key, sh_pos, P);
        if (tr) { trow[7] = gtimer(); trow[5] |= ((long long)ncand << 32); }
        const long long take = cl.h - cl.below;                     // candidates that belong to the h-subset
        const bool ok = (take >= 1 && take <= ncand && W->cand_cnt[k] <= (unsigned)MCD_CAND_CAP);
        const unsigned long long thr = ok ? sh_key[take - 1] : 0ull;
        const bool need_mom = (phase == PH_KEY_COLLECT) || !cl.is_final;
        double tot[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) tot[m] = 0.0;
        if (ok && need_mom) {
          double cs[D];
#pragma unroll
          for (int d = 0; d < D; ++d) cs[d] = cl.cshift[d];
          const double* rsb = s.rs + cl.off;
          for (int i = tid; i < (int)take; i += MCD_THREADS) {
            double r[D];
#pragma unroll
            for (int d = 0; d < D; ++d) r[d] = rsb[(long long)d * s.rs_stride + sh_pos[i]];
            Mom<D>::add(tot, r, cs, true);
          }
          // per-CTA partials, fixed order: lane-strided over CTAs, then the warp tree
          const int w = tid >> 5;
#pragma unroll
          for (int m = 0; m < NM; ++m) {
            if ((m & 31) == w) {                                    // warp (m mod 32) owns moment m
              double a = 0.0;
              for (int bb = lane; bb < G; bb += 32) a += mb.part[((long long)bb * K + k) * NM + m];
              a = warp_sum(a);
              if (lane == 0) tot[m] += a;
            }
          }
        }
        __syncthreads();                                            // sh_key / sh_pos no longer needed
        block_reduce_vec<NM>(tot, sh_red, sh_out);
        if (tid == 0) {
          W->cand_cnt[k] = 0u;
          if (!ok) { atomicCAS(&s.ctl->status, 0, SQC_D_SELECT); cl.phase = PH_DONE; }
          else {
            const double h = (double)cl.h;
            bool go_final = false, done = false;
            if (phase == PH_KEY_COLLECT) {
              // random start: loc/scale of the subset (:75-77), det_old = 2 det_new (:80)
              moments_to_loc_scale<D>(sh_out, h, cl.cshift, cl.loc, cl.scale);
              cl.det_new = small_det(cl.scale, D);
              cl.det_old = 2.0 * cl.det_new;
              cl.iters = 0;
            } else if (!cl.is_final) {
              // C-step (:85-92)
              moments_to_loc_scale<D>(sh_out, h, cl.cshift, cl.loc, cl.scale);
              cl.det_old = cl.det_new;
              cl.det_new = small_det(cl.scale, D);
              cl.iters++;
              cl.t_prev = thr;
              atomicAdd((unsigned long long*)&s.ctl->mcd_csteps, 1ull);
            } else {
              // consistency rescale (:98-103): q = h-th smallest distance under the final estimates
              const double q = __longlong_as_double((long long)thr);
              const double f = q / cl.q_chisq;
              for (int e = 0; e < D * D; ++e) { cl.scale[e] = cl.scale[e] * f; s.sigma[k * D * D + e] = cl.scale[e]; }
              for (int d = 0; d < D; ++d) s.beta[k * D + d] = cl.loc[d];
              done = true;
            }
            if (!done) {
              if (sh_out[0] != h) { atomicCAS(&s.ctl->status, 0, SQC_D_SELECT); done = true; }
              const bool cont = ((cl.det_old - cl.det_new) > 1e-10) & (cl.det_new > 0) & (cl.iters < s.mcd_iters);   // :84
              go_final = !cont;
              if (go_final && cl.iters == s.mcd_iters) atomicAdd(&s.ctl->mcd_hit, 1);   // :94-95
            }
            if (done) cl.phase = PH_DONE;
            else if (!prepare_dist<D>(cl)) { atomicCAS(&s.ctl->status, 0, SQC_D_NOT_SPD); cl.phase = PH_DONE; }   // chol(), :19
            else {
              cl.is_final = go_final ? 1 : 0;
              cl.phase = PH_DIST_HIST; cl.n_hist_rounds = 0;
              if (cl.t_prev) {
                // the trimmed scatter of the first C-step shrinks the scale (distances grow by up to
                // 1 / c(p)); later C-steps move the threshold by little.  Wide, cheap window either way.
                const double t = __longlong_as_double((long long)cl.t_prev);
                set_dist_window(cl, t * 0.25, t * 8.0);
              } else {
                set_dist_window(cl, cl.q_chisq * 0.25, cl.q_chisq * 4.0);
              }
            }
          }
        }
      }
    }
    if (tr) { trow[3] = gtimer(); mb.trace[0] = rounds; }
    __threadfence();
    grid_barrier(mb.bar, (unsigned)G, bar_gen);
  }
  if (b == 0 && tid == 0) { s.ctl->mcd_passes += rounds; s.ctl->mcd_call += 1; }
}

}  // namespace sqc
