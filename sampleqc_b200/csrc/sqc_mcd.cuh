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
//            appended, without atomics, to per-warp candidate regions (deterministic order).  The
//            cluster's leader CTA bins the candidates once more (1024 fine bins inside the bin), takes
//            every candidate under the fine bin that holds the h-th key, and sorts only the handful
//            inside it by (key, rank, position) — ties resolved towards the lower cell index, the repo's
//            definition of the reference's unspecified tie order.
// Multi-GPU (cells sharded by sample): the leaders of all ranks exchange the histogram, the fine
// histogram and then ONE message with their moments and tie candidates through peer memory
// (sqc_xchg.cuh) and combine them in rank order, so every rank walks through bit-identical cluster
// states; R-compatible keys are read from the rank that generated them (NVLink).
// A cluster walks through  KEY_HIST -> KEY_COLLECT -> [DIST_HIST -> DIST_COLLECT]* -> final
// DIST_HIST -> DIST_COLLECT(q only) -> DONE; a HIST round is skipped ("direct" COLLECT) when the
// threshold can be predicted well enough (random-key start, settled C-steps).
// Dataflow, no grid-wide barriers: CTAs 0 .. K-1 are the cluster LEADERS, the others are WORKERS.  Leader k publishes
// the state of its cluster under an epoch number; every worker streams its static share of the cluster's chunks
// whenever it sees a new epoch, flushes (histogram / moment partials / candidates) and arrives on a counter;
// when all workers have arrived the leader post-processes (histogram scan / candidate ranking / moments -> loc,
// scale, det, Cholesky, loop test :84) and publishes the next epoch.  Clusters advance independently, so one
// cluster's post-processing overlaps the streaming of the others and a slowly converging cluster does not
// hold grid barriers for everybody.
// Two worker classes.  A small cluster that needs many C-steps is the critical chain of a launch (10 rounds against 4-5
// of the big ones at C3), and with one pool its renewed epoch has to wait until every worker has finished the share of
// a big cluster it is in the middle of (non-preemptive, ~17 us of a ~33 us round).  So while big AND small clusters are
// active, a fixed share of the workers (SQC_MCD_SMALL_WORKERS) serves only the small clusters (n_k <= N /
// SQC_MCD_SMALL_DIV) and the others only the big ones (SQC_MCD_JOIN=1: when one class has finished, its workers join the other).  Who
// serves an epoch is decided by the leader and published in the epoch word ({number, done, all}), so the arrival count
// it waits for and the set of per-CTA partials / candidate regions it reads are known exactly; the chunk dealing, hence
// every reduction order, is a function of that set only: results stay bit-reproducible.
#pragma once
#include "sqc_kernels.cuh"
#include "sqc_xchg.cuh"

namespace sqc {

#ifndef SQC_MCD_THREADS
#define SQC_MCD_THREADS 512      // 16 warps x 128 registers (384 threads, leaving room for the key-stream CTAs, measured slower)
#endif
constexpr int MCD_THREADS = SQC_MCD_THREADS;         // one fat CTA per SM: one histogram per SM
constexpr int MCD_WARPS = MCD_THREADS / 32;
#ifndef SQC_MCD_DIRECT
#define SQC_MCD_DIRECT 1          // 0: always take the HIST round (debug / A-B)
#endif
#ifndef SQC_MCD_DIRECT_KEYCAP
#define SQC_MCD_DIRECT_KEYCAP 12000.0   // same for the random-key start round (the leader walks them in ~20 us)
#endif
#ifndef SQC_MCD_DIRECT_CAP
#define SQC_MCD_DIRECT_CAP 12000.0 // expected candidates a direct round may catch
#endif
#ifndef SQC_MCD_NOISE0
#define SQC_MCD_NOISE0 3.0        // first C-step: half-width of the predicted band in units of 1/sqrt(n) (call-to-call std of the threshold ~1.4/sqrt(n))
#endif
#ifndef SQC_MCD_PREDICT
#define SQC_MCD_PREDICT 1         // 0: no band prediction from the previous update_beta_scale_k calls (debug / A-B)
#endif
#ifndef SQC_MCD_KEY_SIGMAS
#define SQC_MCD_KEY_SIGMAS 3.2    // half-width of the direct key band in standard deviations of the h-th key (a miss costs one regular round)
#endif
#ifndef SQC_MCD_SMALL_DIV
#define SQC_MCD_SMALL_DIV 8       // a cluster with n_k <= N / SQC_MCD_SMALL_DIV is "small" (two worker classes, see k_mcd)
#endif
#ifndef SQC_MCD_SMALL_WORKERS
#define SQC_MCD_SMALL_WORKERS 0.22 // share of the worker CTAs reserved for the small clusters while big ones are active (0: one class)
#endif
constexpr int MCD_NB = 4096;                         // histogram bins per window
constexpr int MCD_NF = 1024;                         // fine bins inside the candidate bin (leader post; a 4 KB exchange on multi-GPU)
constexpr int MCD_CAND_CAP = 8192;                   // a bin holding more cells than this is refined by another HIST round
constexpr int MCD_CAPW = 32;                         // candidate slots per (CTA, warp, cluster) region
constexpr int MCD_RPT = 8;                           // candidate regions per leader thread (grids up to 256 CTAs)
constexpr int MCD_TCACHE = 64;                       // tie candidates whose residuals the leader keeps in shared memory
constexpr int MCD_TINY = 512;                        // tie candidates (inside one fine bin) over all ranks
constexpr size_t MCD_SMEM = (size_t)(MCD_NB + 8) * 4 + (size_t)MCD_TINY * 20 + 256;
template <int D> struct McdCfg { static constexpr int CPT = (D <= 4) ? 8 : 4; static constexpr int CHUNK = MCD_THREADS * CPT; };

constexpr int MCD_HMAX = 64;                         // C-steps per launch whose thresholds are remembered; slot MCD_HMAX = the final q-pass

// Thresholds a cluster reached in the previous update_beta_scale_k calls (one per EM iteration), by C-step index.  The
// random start draws a fresh half of the cluster every call, but the half is large: the h-th distance of C-step i moves
// by ~1/sqrt(n) (i = 0) or far less (i > 0) from one call to the next once the labels settle (measured with
// oracle/np_restatement.py traces).  The leader predicts a narrow band around last call's value and goes straight to a
// COLLECT round; a miss falls back to the regular HIST round, so the selection stays exact.  Ring of 3 slots (call % 3).
struct McdPred {
  int call[3], nsteps[3];
  unsigned long long thr[3][MCD_HMAX + 1];             // distance bits; [MCD_HMAX] = the final quantile q (:98-100)
  double dens[3][MCD_HMAX + 1];                        // candidates per unit of relative width around that threshold
};

enum { PH_KEY_HIST = 0, PH_KEY_COLLECT = 1, PH_DIST_HIST = 2, PH_DIST_COLLECT = 3, PH_DONE = 4 };

struct McdClu {
  long long n, h, off, prefix;         // GLOBAL cells and subset size; local segment offset in rs; key-stream offset of the local segment
  long long n_loc;                     // cells of this cluster on this GPU
  long long below;                     // cells strictly under the candidate bin, over all ranks (set by the HIST post)
  unsigned long long wlo, whi;         // histogram window (sortable bits)
  unsigned long long blo, bhi;         // candidate bin (sortable bits)
  unsigned long long t_prev;           // last resolved threshold (distance bits), 0 = none
  int shift, phase, is_final, iters, n_hist_rounds, direct;   // direct: COLLECT round entered without a HIST round (predicted band)
  int key_shift, serve_all;            // shift of the +-8 sigma key window while a direct key round is tried; serve_all: every worker serves this epoch (else: the cluster's class only)
  double last_rel, dens;               // relative change of the threshold over the last C-step; candidates per unit of relative width
  double band_c, band_w;               // centre and relative half-width of the predicted band of a direct distance round
  double det_old, det_new, q_chisq;
  double cshift[SQC_MAXD];             // moment shift
  double loc[SQC_MAXD];
  double scale[SQC_MAXD * SQC_MAXD];
  double linv[SQC_TRI(SQC_MAXD)];      // packed lower-triangular inverse Cholesky factor (row-major)
};

struct McdWork {
  McdClu clu[SQC_MAXK];
  unsigned long long cnt_below[SQC_MAXK];   // HIST: local cells under the window
  unsigned long long cnt_nan[SQC_MAXK];
};

struct GridBar { unsigned int count; unsigned int gen; };

struct McdBufs {
  McdWork* work;
  unsigned int* hist;                  // K x MCD_NB
  double* part;                        // grid x K x NMOM(D) per-CTA moment partials
  unsigned long long* cand_rec;        // K x regions x MCD_CAPW records of 2 + D words: key bits, position, D residuals
  unsigned int* cand_cnt;              // K x regions
  unsigned int* fhist;                 // K x (MCD_NF + 4): fine histogram of the candidates, built by the workers
  const int* keys;                     // R-compatible key stream of this call (N_total ints), or null
  const int* key_peer[XCHG_MAX_WORLD]; // distributed key stream (multi-GPU): rank r's share of this call's keys (peer memory)
  long long key_q;                     // draws per rank of the distributed stream (0: `keys` holds the whole call)
  double key_qinv;
  unsigned long long call_mix;         // counter-mode key mix for this call
  long long stream_base;               // call * N_total: index of this call's first draw
  McdPred* pred;                       // K: threshold history of the previous calls (band prediction)
  int call;                            // index of this update_beta_scale_k call (0 = initial labelling)
  int join;                            // worker classes: 1 = the workers of a class whose clusters are all done join the other class (the
                                       // moment they do depends on timing: results reproducible to rounding only); 0 = never (default)
  unsigned int* epoch;                 // K: state version of every cluster (0 at launch; the leader publishes 1, 2, ...)
  unsigned int* arrive;                // K: worker arrivals (0 at launch)
  int* status;                         // in-kernel error word (0 at launch): identical on every rank, derived from exchanged data only
  int trace_leader;                    // which leader writes the trace
  long long* trace;                    // optional per-round timeline (debug): 16 words per round, word 0 of the buffer = rounds
  XchgDev xc;                          // peer windows (world <= 1: unused)
};

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
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

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
// MULTI = false compiles the cross-GPU exchanges out: their register needs otherwise add spills to the streaming loop
template <int D, bool MULTI>
__global__ void __launch_bounds__(MCD_THREADS, 1) k_mcd(Dev s, McdBufs mb) {
  constexpr int NM = SQC_NMOM(D), CPT = McdCfg<D>::CPT, CHUNK = McdCfg<D>::CHUNK;
  extern __shared__ __align__(16) unsigned char sh_raw[];
  unsigned int* sh_hist = (unsigned int*)sh_raw;                                       // MCD_NB + 8: streaming histogram / leader fine histogram + extras
  unsigned long long* sh_tkey = (unsigned long long*)(sh_raw + (size_t)(MCD_NB + 8) * 4);    // MCD_TINY tie keys
  unsigned int* sh_tpos = (unsigned int*)(sh_tkey + MCD_TINY);                         // MCD_TINY: rank << 28 ... see below
  unsigned int* sh_tref = sh_tpos + MCD_TINY;                                          // MCD_TINY: region * CAPW + slot (own candidates)
  unsigned int* sh_trank = sh_tref + MCD_TINY;                                         // MCD_TINY: owning rank
  __shared__ double sh_red[32 * NM];
  __shared__ double sh_out[NM];
  __shared__ double sh_tr[MCD_TCACHE * D];      // residuals of the first tie candidates (leader)
  __shared__ unsigned short sh_tord[MCD_TINY];  // tie members in (key, rank, position) order
  __shared__ long long sh_scan[33];
  __shared__ int sh_flag[6];
  __shared__ unsigned long long sh_u64[4];
  __shared__ unsigned long long sh_pT1[MCD_HMAX + 1], sh_pT2[MCD_HMAX + 1];   // leader: thresholds of the previous two calls by C-step index (0 = none)
  __shared__ double sh_pD1[MCD_HMAX + 1];
  __shared__ McdClu sh_clu;                    // worker: snapshot of the state of the cluster being streamed
  __shared__ int sh_ep[SQC_MAXK], sh_eloc[SQC_MAXK], sh_done[SQC_MAXK], sh_off[SQC_MAXK], sh_offc[SQC_MAXK], sh_small[SQC_MAXK];

  if (!s.ctl->cont) return;                    // uniform across the grid (and across ranks): nobody reaches a barrier
  const long long t_entry = mb.trace ? gtimer() : 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, G = gridDim.x, b = blockIdx.x, K = s.K;
  const int world = (MULTI && mb.xc.world > 1) ? mb.xc.world : 1, myrank = world > 1 ? mb.xc.rank : 0;
  // the candidate caps bound the per-rank work of a leader post and the load of the per-warp candidate regions: with the
  // cells spread over `world` ranks the global count may be that much larger (an unbalanced cluster that overflows a
  // region takes the regular fallback round)
  const double cap_scale = (double)world;
  const int regions = G * MCD_WARPS, region = b * MCD_WARPS + wid;
  const int NWK = G - K;                       // worker CTAs; CTAs 0 .. K-1 are the cluster leaders
  McdWork* W = mb.work;
  // worker classes (identical on every CTA: derived from the global cluster sizes)
  if (tid < K) sh_small[tid] = (s.clu_n_glob[tid] * (long long)SQC_MCD_SMALL_DIV <= s.N_total) ? 1 : 0;
  __syncthreads();
  int n_small = 0;
  for (int k2 = 0; k2 < K; ++k2) n_small += sh_small[k2];
  int NS = (int)((double)NWK * SQC_MCD_SMALL_WORKERS + 0.5);
  if (NS < 4) NS = 4;
  const bool split = n_small > 0 && n_small < K && NS < NWK && SQC_MCD_SMALL_WORKERS > 0.0;   // both classes present
  if (!split) NS = 0;

  // ---- setup: one leader CTA per cluster ----------------------------------------------------
  if (b < K) {
    const int k = b;
    if (tid == 0) {
      McdClu& c = sh_clu;                                        // the leader keeps its cluster's state in shared memory
      const long long n = s.clu_n_glob[k];
      c.n = n; c.n_loc = s.clu_n[k]; c.off = s.clu_off[k]; c.prefix = s.clu_prefix[k];
      c.h = (long long)(int)((double)(unsigned long long)(n + D + 1) * s.mcd_alpha);   // :306 truncation
      c.iters = 0; c.is_final = 0; c.t_prev = 0ull; c.n_hist_rounds = 0; c.below = 0; c.direct = 0; c.last_rel = 1.0; c.dens = 0.0;
      c.det_old = 0.0; c.det_new = 0.0;
      int err = 0;
      if (n <= 0) err = SQC_D_EMPTY_CLUSTER;                       // x_k.rows(0, -1), :307
      else if (c.h > n || c.h < 1) err = SQC_D_SUBSET_SIZE;        // randperm(N, h) with h > N, :72
      if (err) { atomicCAS(&(*mb.status), 0, err); c.phase = PH_DONE; }
      else {
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
        // the key threshold is predictable: go straight to a COLLECT round over the band p 2^31 +- SQC_MCD_KEY_SIGMAS sigma
        // when that band holds few enough keys (expected 2 s sqrt(n p (1 - p))); a miss falls back to KEY_HIST
        const double exp_cand = 2.0 * SQC_MCD_KEY_SIGMAS * sqrt((double)n * p * (1.0 - p));
        const double blo = p * two31 - SQC_MCD_KEY_SIGMAS * sig - 1.0, bhi = p * two31 + SQC_MCD_KEY_SIGMAS * sig + 1.0;
        if (SQC_MCD_DIRECT && n > 65536 && exp_cand <= SQC_MCD_DIRECT_KEYCAP * cap_scale && blo > 0.0 && bhi < two31) {
          c.blo = (unsigned long long)blo; c.bhi = (unsigned long long)bhi;
          int fs = 0;
          while (fs < 62 && ((c.bhi - c.blo) >> fs) >= (unsigned long long)MCD_NF) ++fs;
          c.key_shift = c.shift;                                  // kept for the fallback
          c.shift = fs + 10; c.direct = 1; c.below = 0;
          c.phase = PH_KEY_COLLECT;
        }
      }
      W->cnt_below[k] = 0ull; W->cnt_nan[k] = 0ull;
    }
    for (int i = tid; i < MCD_NB; i += MCD_THREADS) mb.hist[k * MCD_NB + i] = 0u;
    // threshold history of the previous two calls (band prediction); this call writes slot call % 3
    {
      McdPred& P = mb.pred[k];
      const int s1 = (mb.call + 2) % 3, s2 = (mb.call + 1) % 3;
      const bool v1 = SQC_MCD_PREDICT && mb.call >= 1 && P.call[s1] == mb.call - 1, v2 = v1 && mb.call >= 2 && P.call[s2] == mb.call - 2;
      const int n1 = v1 ? min(P.nsteps[s1], MCD_HMAX) : 0, n2 = v2 ? min(P.nsteps[s2], MCD_HMAX) : 0;
      for (int i = tid; i <= MCD_HMAX; i += MCD_THREADS) {
        const bool in1 = v1 && (i < n1 || i == MCD_HMAX), in2 = v2 && (i < n2 || i == MCD_HMAX);
        sh_pT1[i] = in1 ? P.thr[s1][i] : 0ull; sh_pD1[i] = in1 ? P.dens[s1][i] : 0.0;
        sh_pT2[i] = in2 ? P.thr[s2][i] : 0ull;
      }
      __syncthreads();
      if (tid == 0) { P.call[mb.call % 3] = mb.call; P.nsteps[mb.call % 3] = 0; P.thr[mb.call % 3][MCD_HMAX] = 0ull; }
    }
  }

  if (b < K) {
    // =========================== leader of cluster k: waits for the workers, post-processes, publishes ===========================
    const int k = b;
    __syncthreads();
    // publish = copy the state to global memory, fence, release the epoch (the workers snapshot it from there)
    // epoch word = {number, done, all}: bit 0 = every worker serves this epoch (else the cluster's class only), bit 1 = done
    const int my_small = sh_small[k];
    auto publish = [&](int epoch_no) {
      if (tid == 0) {
        int all = 1;
        if (split) {
          // Who serves an epoch fixes the chunk dealing and with it the order of every floating-point sum.  By default the
          // classes stay apart for the whole launch, so that order is a function of the data alone: two runs give the
          // same bits ("seeds are replicable", tests/testthat/test-03_fit_sampleqc.R:81-88).  mb.join: the other class
          // joins once all of ITS clusters are done — WHEN this leader notices is a matter of timing.
          if (!mb.join) all = 0;
          else for (int k2 = 0; k2 < K; ++k2) if (sh_small[k2] != my_small && !(ld_acquire_u32(&mb.epoch[k2]) & 2u)) all = 0;
        }
        sh_clu.serve_all = all;
      }
      __syncthreads();
      const unsigned long long* src = (const unsigned long long*)&sh_clu;
      unsigned long long* dst = (unsigned long long*)&W->clu[k];
      for (int i = tid; i < (int)(sizeof(McdClu) / 8); i += MCD_THREADS) dst[i] = src[i];
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_u32(&mb.epoch[k], ((unsigned)epoch_no << 2) | (sh_clu.phase == PH_DONE ? 2u : 0u) | (sh_clu.serve_all ? 1u : 0u));
    };
    int ep = 1;
    publish(1);
    if (mb.trace && b == mb.trace_leader && tid == 0) { mb.trace[1] = t_entry; mb.trace[2] = gtimer(); }
    if (tid == 0) {
      // qchisq(h / n, D) of the consistency factor (:101-102) is first needed after the start rounds: evaluate it
      // while the workers stream the first one
      McdClu& c = sh_clu;
      if (c.phase != PH_DONE) c.q_chisq = dev_qchisq((double)c.h / (double)c.n, D);
    }
    long long my_epochs = 0, my_cells = 0, my_bytes = 0;
    unsigned int want_arrivals = 0;                               // (thread 0) cumulative
    for (;;) {
      if (tid == 0) sh_flag[2] = (sh_clu.phase != PH_DONE) ? 1 : 0;
      __syncthreads();
      if (!sh_flag[2]) break;
      ++my_epochs;
      const bool tr = mb.trace && b == mb.trace_leader && tid == 0 && my_epochs < 500;
      long long* trow = mb.trace + 16 * my_epochs;
      if (tr) { trow[0] = gtimer(); trow[4] = sh_clu.phase; trow[5] = 0; for (int e = 6; e < 16; ++e) trow[e] = 0; trow[13] = sh_clu.direct; trow[14] = sh_clu.iters; }
      if (tid == 0) {
        const int ph = sh_clu.phase; const long long n = sh_clu.n_loc;
        const bool keyr = (ph == PH_KEY_HIST || ph == PH_KEY_COLLECT);
        my_cells += n; my_bytes += n * ((ph == PH_KEY_HIST ? 0 : 8 * D) + ((keyr && mb.keys) ? 4 : 0));
        // every worker of the serving set arrives once per epoch, whether or not it holds cells of this cluster
        want_arrivals += (unsigned)(sh_clu.serve_all ? NWK : (my_small ? NS : NWK - NS));
        // watchdog: one time budget for leaders, workers and exchanges (a worker or a peer died); the leader then stops
        unsigned int spins = 0; long long t0w = 0;
        sh_flag[3] = 0;
        while (ld_acquire_u32(&mb.arrive[k]) < want_arrivals) {
          if ((++spins & 0x3fffu) == 0u) {
            const long long now = gtimer();
            if (!t0w) t0w = now;
            else if (now - t0w > mb.xc.budget_ns || *mb.status == SQC_D_TIMEOUT) { atomicCAS(&(*mb.status), 0, SQC_D_TIMEOUT); sh_flag[3] = 1; break; }
          }
        }
        __threadfence();
      }
      __syncthreads();
      if (sh_flag[3]) {                                            // timed out: no post-processing on partial data, no further exchange
        if (tid == 0) sh_clu.phase = PH_DONE;
        __syncthreads();
        ++ep; publish(ep);
        break;
      }
      if (tr) { trow[1] = gtimer(); trow[2] = trow[1]; }
      {
      McdClu& cl = sh_clu;
      const int phase = cl.phase;
      const bool key_phase = (phase == PH_KEY_HIST || phase == PH_KEY_COLLECT);
      if (phase == PH_KEY_HIST || phase == PH_DIST_HIST) {
        // ---- HIST post: locate the bin that holds the h-th smallest key (over all ranks) ----
        for (int i = tid; i < MCD_NB; i += MCD_THREADS) { sh_hist[i] = mb.hist[k * MCD_NB + i]; mb.hist[k * MCD_NB + i] = 0u; }
        if (tid == 0) {
          const unsigned long long bl = W->cnt_below[k], nn = W->cnt_nan[k];
          // the count travels as two 16-bit limbs (a rank holds < 2^31 cells): their sums over <= 8 ranks cannot wrap,
          // whatever the total number of cells
          sh_hist[MCD_NB] = (unsigned int)(bl & 0xffffull); sh_hist[MCD_NB + 1] = (unsigned int)(bl >> 16); sh_hist[MCD_NB + 2] = nn ? 1u : 0u;
          W->cnt_below[k] = 0ull;
        }
        __syncthreads();
        if constexpr (MULTI) xchg_sum_u32(mb.xc, k, sh_hist, MCD_NB + 3);
        const long long below = (long long)sh_hist[MCD_NB] + ((long long)sh_hist[MCD_NB + 1] << 16);
        const bool has_nan = sh_hist[MCD_NB + 2] != 0;
        constexpr int PER = (MCD_NB + MCD_THREADS - 1) / MCD_THREADS;
        unsigned int hv[PER]; long long mine = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { hv[i] = (tid * PER + i < MCD_NB) ? sh_hist[tid * PER + i] : 0u; mine += hv[i]; }
        long long total_in;
        long long run = block_excl_scan<MCD_THREADS>(mine, sh_scan, &total_in);
        const long long want = cl.h - below;                        // 1-based rank inside the window
        if (tid == 0) sh_flag[1] = -1;
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
          if (has_nan) { atomicCAS(&(*mb.status), 0, SQC_D_NAN); cl.phase = PH_DONE; }          // sort_index(): detected NaN, :50
          else if (bin < 0) {
            // the order statistic lies outside the window: widen to the whole side it is on
            const unsigned long long top = key_phase ? (1ull << 31) : 0x7ff0000000000001ull;
            if (want < 1) { cl.whi = cl.wlo; cl.wlo = 0ull; } else { cl.wlo = cl.whi; cl.whi = top; }
            cl.shift = window_shift(cl.wlo, cl.whi);
            if (cl.whi <= cl.wlo || cl.n_hist_rounds > 40) { atomicCAS(&(*mb.status), 0, SQC_D_SELECT); cl.phase = PH_DONE; }
          } else {
            const unsigned long long blo = cl.wlo + ((unsigned long long)bin << cl.shift);
            unsigned long long bhi = blo + (1ull << cl.shift);
            if (bhi > cl.whi) bhi = cl.whi;
            if (sh_u64[1] > (unsigned long long)MCD_CAND_CAP * (unsigned long long)world && cl.shift > 0 && cl.n_hist_rounds <= 40) {
              cl.wlo = blo; cl.whi = bhi; cl.shift = window_shift(blo, bhi);                 // refine inside the bin
            } else if (sh_u64[1] > (unsigned long long)MCD_CAND_CAP * (unsigned long long)world) {
              atomicCAS(&(*mb.status), 0, SQC_D_SELECT); cl.phase = PH_DONE;
            } else {
              cl.blo = blo; cl.bhi = bhi; cl.below = below + (long long)sh_u64[0];
              cl.phase = key_phase ? PH_KEY_COLLECT : PH_DIST_COLLECT;
            }
          }
        }
      } else {
        // ---- COLLECT post: exact members of the candidate bin, then the moments ---------------
        const bool direct = cl.direct != 0;                          // no HIST round before: `below` comes from the workers' counts
        const bool need_mom = (phase == PH_KEY_COLLECT) || !cl.is_final;
        const bool have_part = need_mom || direct;
        const unsigned long long blo = cl.blo;
        const int fshift = cl.shift > 10 ? cl.shift - 10 : 0;        // MCD_NF = 1024 fine bins inside the candidate bin
        const unsigned int* ccnt = mb.cand_cnt + (long long)k * regions;
        // CTAs that served this epoch: only their partials / candidate regions are current
        const int cta_lo = K + (cl.serve_all ? 0 : (my_small ? 0 : NS)), cta_hi = K + (cl.serve_all ? NWK : (my_small ? NS : NWK));
        // per-CTA moment partials: issue the loads first, use them at the end (warp (m mod warps) owns moment m)
        double part_pre[(NM + MCD_WARPS - 1) / MCD_WARPS];
#pragma unroll
        for (int q = 0; q < (NM + MCD_WARPS - 1) / MCD_WARPS; ++q) {
          const int m = wid + q * MCD_WARPS;
          double a2 = 0.0;
          if (m < NM && have_part) {
            double pv[8];                                            // all loads in flight together (grids up to 256 CTAs)
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int bb = lane + 32 * i; pv[i] = (bb >= cta_lo && bb < cta_hi) ? mb.part[((long long)bb * K + k) * NM + m] : 0.0; }
#pragma unroll
            for (int i = 0; i < 8; ++i) a2 += pv[i];
          }
          part_pre[q] = a2;
        }
        // candidate counts of this thread's regions: same L2 round trip as the partials and the fine histogram
        unsigned int cn[MCD_RPT];
#pragma unroll
        for (int i = 0; i < MCD_RPT; ++i) { const int rg = tid * MCD_RPT + i; cn[i] = (rg >= cta_lo * MCD_WARPS && rg < cta_hi * MCD_WARPS) ? ccnt[rg] : 0u; }   // thread t: regions t * MCD_RPT ..
        // 1. fine histogram of the candidates: built by the workers with atomics on mb.fhist
        {
          constexpr int FPT = (MCD_NF + 4 + MCD_THREADS - 1) / MCD_THREADS;
          unsigned int fv[FPT];
#pragma unroll
          for (int i = 0; i < FPT; ++i) { const int e = tid + i * MCD_THREADS; fv[i] = (e < MCD_NF + 4) ? mb.fhist[k * (MCD_NF + 4) + e] : 0u; }
#pragma unroll
          for (int i = 0; i < FPT; ++i) { const int e = tid + i * MCD_THREADS; if (e < MCD_NF + 4) { sh_hist[e] = fv[i]; mb.fhist[k * (MCD_NF + 4) + e] = 0u; } }
        }
        if (tid == 0) { sh_flag[4] = 0; sh_flag[5] = 0; }
        if (direct && wid == 0) {                                    // cells strictly under the band: moment 0 of the partials (exact: counts < 2^53)
          const double cb = warp_sum(part_pre[0]);
          const unsigned long long cnt = (unsigned long long)cb;
          if (lane == 0) { sh_hist[MCD_NF + 2] = (unsigned int)(cnt & 0xffffull); sh_hist[MCD_NF + 3] = (unsigned int)(cnt >> 16); }   // two 16-bit limbs, as above
        }
        __syncthreads();
        if (tr) trow[8] = gtimer();
        if constexpr (MULTI) xchg_sum_u32(mb.xc, k, sh_hist, MCD_NF + 4);
        const bool overflow = sh_hist[MCD_NF + 1] != 0;
        constexpr int PER = (MCD_NF + MCD_THREADS - 1) / MCD_THREADS;
        unsigned int hv[PER]; long long mine = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { hv[i] = (tid * PER + i < MCD_NF) ? sh_hist[tid * PER + i] : 0u; mine += hv[i]; }
        long long ncand_all;
        long long run = block_excl_scan<MCD_THREADS>(mine, sh_scan, &ncand_all);
        const long long below_all = direct ? ((long long)sh_hist[MCD_NF + 2] + ((long long)sh_hist[MCD_NF + 3] << 16)) : cl.below;
        const long long take = cl.h - below_all;                     // candidates (over all ranks) that belong to the h-subset
        if (tid == 0) sh_flag[1] = -1;
        __syncthreads();
        if (take >= 1 && take <= ncand_all) {
#pragma unroll
          for (int i = 0; i < PER; ++i) {
            if (run < take && take <= run + (long long)hv[i]) { sh_flag[1] = tid * PER + i; sh_u64[0] = (unsigned long long)run; sh_u64[1] = hv[i]; }
            run += hv[i];
          }
        }
        __syncthreads();
        const int fbin = sh_flag[1];
        const long long fbelow = (long long)sh_u64[0];
        const bool ok = !overflow && fbin >= 0 && sh_u64[1] <= (unsigned long long)MCD_TINY;
        if (tr) { trow[6] = gtimer(); trow[5] |= (ncand_all << 32); }
        // 2. candidates under the fine bin are members; those inside it go to the tie list
        double tot[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) tot[m] = 0.0;
        if (ok) {
          double cs[D];
#pragma unroll
          for (int d = 0; d < D; ++d) cs[d] = cl.cshift[d];
          // The candidates are numbered flat — regions in index order, slots in order — and dealt round-robin over the
          // threads: thread t takes candidates t, t + threads, ...  (a prefix of the region counts in shared memory, a
          // binary search per candidate).  However the candidates spread over the regions, the pass takes
          // ceil(candidates / (threads x FB)) L2 round trips (one for a few hundred candidates); walking each region
          // slot by slot took as many as the fullest region had slots.  Which thread adds which member depends on the
          // regions' contents only: reproducible.
          unsigned int mine_c = 0;
#pragma unroll
          for (int i = 0; i < MCD_RPT; ++i) { cn[i] = min(cn[i], (unsigned)MCD_CAPW); mine_c += cn[i]; }
          long long total_c;
          const long long base_c = block_excl_scan<MCD_THREADS>((long long)mine_c, sh_scan, &total_c);
          unsigned int* sh_pref = sh_hist;                           // the fine histogram has been consumed (fbin, fbelow, overflow are in registers / sh_u64)
          {
            unsigned int run_c = (unsigned int)base_c;
#pragma unroll
            for (int i = 0; i < MCD_RPT; ++i) { sh_pref[tid * MCD_RPT + i] = run_c; run_c += cn[i]; }
          }
          static_assert(MCD_THREADS * MCD_RPT <= MCD_NB, "region prefixes fit the histogram buffer");
          if (tid == 0) sh_pref[MCD_THREADS * MCD_RPT] = (unsigned int)total_c;
          __syncthreads();
          const int ntot = (int)total_c;
          constexpr int FB = (D <= 4) ? 4 : 2;                       // records in flight per thread
          for (int f0 = tid; f0 < ntot; f0 += FB * MCD_THREADS) {
            unsigned long long wv[FB][2 + D]; int rgs[FB], sls[FB]; bool live[FB];
#pragma unroll
            for (int q = 0; q < FB; ++q) {
              const int f = f0 + q * MCD_THREADS;
              live[q] = f < ntot;
              int lo = 0, hi = MCD_THREADS * MCD_RPT;               // largest region with prefix <= f (regions without candidates share their successor's prefix)
              if (live[q]) {
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sh_pref[mid] <= (unsigned int)f) lo = mid; else hi = mid; }
              }
              rgs[q] = lo; sls[q] = live[q] ? f - (int)sh_pref[lo] : 0;
              const unsigned long long* rec = mb.cand_rec + (((long long)k * regions + lo) * MCD_CAPW + sls[q]) * (2 + D);
#pragma unroll
              for (int e = 0; e < 2 + D; ++e) wv[q][e] = live[q] ? __ldcg(rec + e) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < FB; ++q) {
              if (!live[q]) continue;
              const unsigned long long bits = wv[q][0];
              const int fb = (int)((bits - blo) >> fshift);
              if (fb < fbin) {
                if (need_mom) {
                  double r[D];
#pragma unroll
                  for (int d = 0; d < D; ++d) r[d] = __longlong_as_double((long long)wv[q][2 + d]);
                  Mom<D>::add(tot, r, cs, true);
                }
              } else if (fb == fbin) {
                const int t = atomicAdd(&sh_flag[4], 1);
                if (t < MCD_TINY) {
                  sh_tkey[t] = bits; sh_tpos[t] = (unsigned int)wv[q][1]; sh_tref[t] = (unsigned int)(rgs[q] * MCD_CAPW + sls[q]); sh_trank[t] = (unsigned)myrank;
                  if (t < MCD_TCACHE) {
#pragma unroll
                    for (int d = 0; d < D; ++d) sh_tr[t * D + d] = __longlong_as_double((long long)wv[q][2 + d]);
                  }
                }
              }
            }
          }
        }
        __syncthreads();
        if (tr) trow[9] = gtimer();
        int ntie = min(sh_flag[4], MCD_TINY);
        // 3. (multi-GPU) ONE exchange per COLLECT post: every rank pushes its moments without the tie members and its
        //    tie list (key, position, residual) to all ranks; each rank then ranks the union and adds the members in
        //    canonical (key, rank, position) order behind the rank-ordered sum of the moments — identical everywhere.
        //    Message lines: [0] = number of ties, [1 .. NM] = moments, then (2 + D) lines per tie.
        [[maybe_unused]] double msum = 0.0;
        [[maybe_unused]] unsigned long long xseq = 0ull;
        constexpr int TIE0 = 1 + NM, TIEL = 2 + D;
        if constexpr (MULTI) {
          if (ok) {
            if (need_mom) {
#pragma unroll
              for (int q = 0; q < (NM + MCD_WARPS - 1) / MCD_WARPS; ++q) {
                const double a = warp_sum(part_pre[q]);
#pragma unroll
                for (int m = 0; m < NM; ++m) if (m == wid + q * MCD_WARPS && lane == 0) tot[m] += a;
              }
            }
            block_reduce_vec<NM>(tot, sh_red, sh_out);
            xseq = xchg_begin(mb.xc, k);
            if (tid == 0) xchg_put(mb.xc, k, xseq, 0, (unsigned)ntie, 0u);
            if (tid < NM) xchg_put_f64(mb.xc, k, xseq, 1 + tid, sh_out[tid]);
            for (int i = tid; i < ntie; i += MCD_THREADS) {
              const long long cb = (long long)k * regions * MCD_CAPW + (long long)sh_tref[i];
              const int l0 = TIE0 + i * TIEL;
              xchg_put(mb.xc, k, xseq, l0, (unsigned)sh_tkey[i], (unsigned)(sh_tkey[i] >> 32));
              xchg_put(mb.xc, k, xseq, l0 + 1, sh_tpos[i], 0u);
#pragma unroll
              for (int d = 0; d < D; ++d)
                xchg_put_f64(mb.xc, k, xseq, l0 + 2 + d, (i < MCD_TCACHE) ? sh_tr[i * D + d] : __longlong_as_double((long long)mb.cand_rec[cb * (2 + D) + 2 + d]));
            }
            ulonglong2 hd[XCHG_MAX_WORLD];
            xchg_wait_all(mb.xc, k, xseq, 0, hd);
            if (tid < NM) {
              ulonglong2 mv[XCHG_MAX_WORLD];
              xchg_wait_all(mb.xc, k, xseq, 1 + tid, mv);
#pragma unroll
              for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < world) msum += line_f64(mv[r]);
            }
#pragma unroll
            for (int r = 0; r < XCHG_MAX_WORLD; ++r) {
              if (r >= world || r == myrank) continue;
              const int cnt = min((int)(unsigned)hd[r].x, MCD_TINY);
              for (int i = tid; i < cnt; i += MCD_THREADS) {
                const ulonglong2 e0 = xchg_wait(mb.xc, k, xseq, r, TIE0 + i * TIEL), e1 = xchg_wait(mb.xc, k, xseq, r, TIE0 + i * TIEL + 1);
                const int t = atomicAdd(&sh_flag[4], 1);
                if (t < MCD_TINY) {
                  sh_tkey[t] = (unsigned long long)(unsigned)e0.x | ((unsigned long long)(unsigned)e0.y << 32);
                  sh_tpos[t] = (unsigned)e1.x; sh_tref[t] = (unsigned)i; sh_trank[t] = (unsigned)r;
                }
              }
            }
            __syncthreads();
            ntie = min(sh_flag[4], MCD_TINY);
          }
        }
        // tiny selection by rank counting: entry i is a member iff fewer than take2 entries precede it in (key, rank, pos) order
        const long long take2 = take - fbelow;
        unsigned long long thr = 0ull;
        [[maybe_unused]] double ttot[MULTI ? NM : 1];
        if (ok) {
          for (int i = tid; i < ntie; i += MCD_THREADS) {
            const unsigned long long ki = sh_tkey[i]; const unsigned ri = sh_trank[i], pi = sh_tpos[i];
            int before = 0;
            for (int j2 = 0; j2 < ntie; ++j2) {
              const unsigned long long kj = sh_tkey[j2]; const unsigned rj = sh_trank[j2], pj = sh_tpos[j2];
              before += (kj < ki) || (kj == ki && (rj < ri || (rj == ri && pj < pi)));
            }
            if (before == (int)take2 - 1) sh_u64[2] = ki;           // the h-th smallest key itself
            if (before < (int)take2) sh_tord[before] = (unsigned short)i;   // members by rank ((key, rank, position) is unique)
          }
          __syncthreads();
          thr = sh_u64[2];
          if (need_mom && tid == 0) {
            // tie members in ascending (key, rank, position) order, so that the summation order is deterministic
            // (single GPU: all entries are this rank's; multi-GPU: the peers' residuals are read from the message)
            double cs[D];
#pragma unroll
            for (int d = 0; d < D; ++d) cs[d] = cl.cshift[d];
            if constexpr (MULTI) {
#pragma unroll
              for (int m = 0; m < NM; ++m) ttot[m] = 0.0;
            }
            const int nmem = (int)min((long long)ntie, take2);
            for (int m2 = 0; m2 < nmem; ++m2) {
              const int best = (int)sh_tord[m2];
              const unsigned ref = sh_tref[best];
              double r[D];
              if (!MULTI || sh_trank[best] == (unsigned)myrank) {
                const long long cb = (long long)k * regions * MCD_CAPW + (long long)ref;
#pragma unroll
                for (int d = 0; d < D; ++d) r[d] = (best < MCD_TCACHE) ? sh_tr[best * D + d] : __longlong_as_double((long long)mb.cand_rec[cb * (2 + D) + 2 + d]);
              } else {
                if constexpr (MULTI) {
#pragma unroll
                  for (int d = 0; d < D; ++d) r[d] = line_f64(xchg_wait(mb.xc, k, xseq, (int)sh_trank[best], TIE0 + (int)ref * TIEL + 2 + d));
                }
              }
              if constexpr (MULTI) Mom<D>::add(ttot, r, cs, true); else Mom<D>::add(tot, r, cs, true);
            }
          }
        }
        if (tr) trow[7] = gtimer();
        if constexpr (MULTI) {
          // moments = rank-ordered sum of the ranks' moments + the tie members (thread 0 is also the reader below)
          __syncthreads();
          if (ok && tid < NM) sh_out[tid] = msum;
          __syncthreads();
          if (ok && need_mom && tid == 0) {
#pragma unroll
            for (int m = 0; m < NM; ++m) sh_out[m] += ttot[m];
          }
          if (tr) trow[10] = gtimer();
        } else {
          // 4. per-CTA partials (loaded at the start), fixed order: lane-strided over CTAs, then the warp tree
          if (ok && need_mom) {
#pragma unroll
            for (int q = 0; q < (NM + MCD_WARPS - 1) / MCD_WARPS; ++q) {
              const double a = warp_sum(part_pre[q]);
#pragma unroll
              for (int m = 0; m < NM; ++m) if (m == wid + q * MCD_WARPS && lane == 0) tot[m] += a;
            }
          }
          __syncthreads();
          if (tr) trow[10] = gtimer();
          block_reduce_vec<NM>(tot, sh_red, sh_out);
        }
        if (tr) trow[11] = gtimer();
        if (tid == 0) {
          if (!ok && direct && key_phase) {
            // the +-4 sigma key band missed (about once in 10^4 cluster starts): regular KEY_HIST round over +-8 sigma
            cl.direct = 0; cl.shift = cl.key_shift;
            cl.phase = PH_KEY_HIST; cl.n_hist_rounds = 0;
          } else if (!ok && direct) {
            // the predicted band missed the h-th distance (or caught too many cells): take the regular HIST round
            cl.direct = 0; cl.last_rel = 1.0;
            cl.phase = PH_DIST_HIST; cl.n_hist_rounds = 0;
            const double t = cl.band_c;
            set_dist_window(cl, t * 0.25, t * 8.0);
          } else if (!ok) {
            // too many candidates in one place (region overflow or a crowded fine bin): refine with another HIST round
            if (fbin >= 0 && cl.shift > 0 && cl.n_hist_rounds <= 40) {
              cl.wlo = cl.blo; cl.whi = cl.bhi; cl.shift = window_shift(cl.wlo, cl.whi);
              cl.phase = key_phase ? PH_KEY_HIST : PH_DIST_HIST;
            } else { atomicCAS(&(*mb.status), 0, SQC_D_SELECT); cl.phase = PH_DONE; }
          } else {
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
              {
                const double tn = __longlong_as_double((long long)thr);
                cl.last_rel = cl.t_prev ? fabs(tn - __longlong_as_double((long long)cl.t_prev)) / tn : 1.0;
                // candidates per unit of relative width of the bin just resolved (bit distance / 2^52 ~ relative distance)
                const double relw = (double)(cl.bhi - cl.blo) * 2.220446049250313e-16;
                cl.dens = relw > 0.0 ? (double)ncand_all / relw : 1e300;
              }
              cl.t_prev = thr;
              if (cl.iters - 1 < MCD_HMAX) { McdPred& P = mb.pred[k]; P.thr[mb.call % 3][cl.iters - 1] = thr; P.dens[mb.call % 3][cl.iters - 1] = cl.dens; }
              if (tr) trow[12] = (long long)thr;
              atomicAdd((unsigned long long*)&s.ctl->mcd_csteps, 1ull);
            } else {
              // consistency rescale (:98-103): q = h-th smallest distance under the final estimates
              const double q = __longlong_as_double((long long)thr);
              {
                McdPred& P = mb.pred[k];
                const double relw = (double)(cl.bhi - cl.blo) * 2.220446049250313e-16;
                P.thr[mb.call % 3][MCD_HMAX] = thr; P.dens[mb.call % 3][MCD_HMAX] = relw > 0.0 ? (double)ncand_all / relw : 1e300;
                P.nsteps[mb.call % 3] = cl.iters;
              }
              if (tr) trow[12] = (long long)thr;
              const double f = q / cl.q_chisq;
              for (int e = 0; e < D * D; ++e) { cl.scale[e] = cl.scale[e] * f; s.sigma[k * D * D + e] = cl.scale[e]; }
              for (int d = 0; d < D; ++d) s.beta[k * D + d] = cl.loc[d];
              done = true;
            }
            if (!done) {
              if (sh_out[0] != h) { atomicCAS(&(*mb.status), 0, SQC_D_SELECT); done = true; }
              const bool cont = ((cl.det_old - cl.det_new) > 1e-10) & (cl.det_new > 0) & (cl.iters < s.mcd_iters);   // :84
              go_final = !cont;
              if (go_final && cl.iters == s.mcd_iters) atomicAdd(&s.ctl->mcd_hit, 1);   // :94-95
            }
            if (done) cl.phase = PH_DONE;
            else if (!prepare_dist<D>(cl)) { atomicCAS(&(*mb.status), 0, SQC_D_NOT_SPD); cl.phase = PH_DONE; }   // chol(), :19
            else {
              cl.is_final = go_final ? 1 : 0;
              cl.phase = PH_DIST_HIST; cl.n_hist_rounds = 0; cl.direct = 0;
              // Skip the HIST round when the threshold can be predicted: a band that should hold the next threshold and
              // few enough cells; the COLLECT round counts the cells under the band itself.  A miss costs one regular
              // round.  Two predictors, the narrower band wins:
              //  (a) inside this call, once the threshold has settled: around the last threshold, +- twice its last move;
              //  (b) from the previous calls (EM iterations): the threshold the same C-step index reached last time,
              //      +- 2.5 x its move between the last two calls + the sampling noise of the random start
              //      (~1/sqrt(n) relative for the first C-step, far less afterwards).
              double wband = 1e300, tband = 0.0;
              if (cl.t_prev && cl.last_rel < 0.02) {
                const double w = 2.0 * cl.last_rel + 5e-4;
                if (cl.dens * 2.0 * w <= SQC_MCD_DIRECT_CAP * cap_scale) { wband = w; tband = __longlong_as_double((long long)cl.t_prev); }
              }
              {
                const int hi = go_final ? MCD_HMAX : (cl.iters < MCD_HMAX ? cl.iters : -1);
                if (hi >= 0 && sh_pT1[hi]) {
                  const double c1 = __longlong_as_double((long long)sh_pT1[hi]);
                  const double noise = ((hi == 0) ? SQC_MCD_NOISE0 : 1.5) / sqrt((double)cl.n);
                  const double w = sh_pT2[hi] ? 2.5 * fabs(c1 - __longlong_as_double((long long)sh_pT2[hi])) / c1 + noise : 6.0 * noise;
                  if (sh_pD1[hi] * 2.0 * w <= SQC_MCD_DIRECT_CAP * cap_scale && w < wband && w < 0.2) { wband = w; tband = c1; }
                }
              }
              //  (c) while the labels still move from call to call the thresholds shift as a whole, but their STEP from one
              //      C-step index to the next repeats: this call's last threshold times the ratio the previous call showed
              //      between the same two indices (only where that ratio is close to 1: the settled tail and the final q-pass).
              if (SQC_MCD_PREDICT && cl.t_prev && cl.iters >= 1) {
                const int hi = go_final ? MCD_HMAX : (cl.iters < MCD_HMAX ? cl.iters : -1), lo = cl.iters - 1;
                if (hi >= 0 && lo < MCD_HMAX && sh_pT1[hi] && sh_pT1[lo]) {
                  const double ratio = __longlong_as_double((long long)sh_pT1[hi]) / __longlong_as_double((long long)sh_pT1[lo]);
                  const double w = 0.5 * fabs(ratio - 1.0) + 3.0 / sqrt((double)cl.n) + 2e-4;
                  const double dens_c = sh_pD1[hi] > 0.0 ? sh_pD1[hi] : cl.dens;
                  if (fabs(ratio - 1.0) < 0.01 && dens_c * 2.0 * w <= SQC_MCD_DIRECT_CAP * cap_scale && w < wband) {
                    wband = w; tband = __longlong_as_double((long long)cl.t_prev) * ratio;
                  }
                }
              }
              if (wband < 1e300 && SQC_MCD_DIRECT) {
                const double t = tband;
                cl.band_c = t; cl.band_w = wband;
                cl.blo = dist_bits(t * (1.0 - wband)); cl.bhi = dist_bits(t * (1.0 + wband));
                int fs = 0;
                while (fs < 62 && ((cl.bhi - cl.blo) >> fs) >= (unsigned long long)MCD_NF) ++fs;
                cl.shift = fs + 10;                                 // the workers bin candidates with shift - 10
                cl.below = 0; cl.direct = 1;
                cl.phase = PH_DIST_COLLECT;
              } else if (cl.t_prev) {
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
      if (tr) { trow[3] = gtimer(); mb.trace[0] = my_epochs; }
      __syncthreads();
      ++ep;
      publish(ep);
    }
    if (tid == 0) {
      atomicAdd((unsigned long long*)&s.ctl->mcd_passes, (unsigned long long)my_epochs);
      atomicAdd((unsigned long long*)&s.ctl->mcd_cells, (unsigned long long)my_cells);
      atomicAdd((unsigned long long*)&s.ctl->mcd_bytes, (unsigned long long)my_bytes);
      if (k == 0) s.ctl->mcd_call += 1;
      if ((*mb.status)) atomicCAS(&s.ctl->status, 0, (*mb.status));
    }
  } else {
    // =========================== worker: streams its share of every cluster whenever that cluster's state is renewed ===========================
    const int w = b - K;
    const int wk_small = (split && w < NS) ? 1 : 0;                // this worker's class
    if (tid < K) { sh_eloc[tid] = 0; sh_done[tid] = 0; }
    if (tid == 0) {                                               // static chunk assignment: round robin over the concatenated chunk lists
      long long run = 0, run_c[2] = {0, 0};
      for (int k = 0; k < K; ++k) {
        const long long nc = (s.clu_n[k] + CHUNK - 1) / CHUNK;
        const int c = sh_small[k], Mc = c ? NS : NWK - NS;
        sh_off[k] = (int)(run % NWK); run += nc;                    // every worker serves
        sh_offc[k] = Mc > 0 ? (int)(run_c[c] % Mc) : 0; run_c[c] += nc;   // the cluster's class only
      }
    }
    __syncthreads();
    unsigned int idle = 0; long long t_idle = 0;
    for (;;) {
      if (tid < K) sh_ep[tid] = (int)ld_acquire_u32(&mb.epoch[tid]);
      __syncthreads();
      // serve ONE renewed cluster per poll, the smallest first: a small cluster that needs many C-steps then cycles at
      // its own (short) period instead of waiting for a sweep over the big ones.  Epochs meant for the other class and
      // finished clusters are only noted.
      if (tid == 0) {
        int sel = -1, nd = 0;
        for (int k = 0; k < K; ++k) {
          if (sh_done[k]) { ++nd; continue; }
          if (sh_ep[k] == sh_eloc[k]) continue;
          if (sh_ep[k] & 2) { sh_done[k] = 1; sh_eloc[k] = sh_ep[k]; ++nd; continue; }
          if (!(sh_ep[k] & 1) && sh_small[k] != wk_small) { sh_eloc[k] = sh_ep[k]; continue; }
          if (sel < 0 || s.clu_n[k] < s.clu_n[sel]) sel = k;
        }
        sh_flag[0] = sel; sh_flag[1] = nd;
      }
      __syncthreads();
      int n_done = sh_flag[1]; bool any_new = false;
      for (int k = sh_flag[0]; k >= 0 && k == sh_flag[0]; k = -1) {
        const int ep = sh_ep[k];
        any_new = true;
        // snapshot of the cluster state (L2 loads: another SM wrote it)
        {
          const unsigned long long* src = (const unsigned long long*)&W->clu[k];
          unsigned long long* dst = (unsigned long long*)&sh_clu;
          for (int i = tid; i < (int)(sizeof(McdClu) / 8); i += MCD_THREADS) dst[i] = __ldcg(src + i);
        }
        __syncthreads();
        // my place in the set of workers that serve this epoch
        const int M = (ep & 1) ? NWK : (wk_small ? NS : NWK - NS), me = (ep & 1) ? w : (wk_small ? w : w - NS);
        const int off_k = (ep & 1) ? sh_off[k] : sh_offc[k];
        const long long nch = (sh_clu.n_loc + CHUNK - 1) / CHUNK;
        const long long c_first = ((long long)me - off_k + M) % M;
      const McdClu& cl = sh_clu;
      const int phase = cl.phase;
      const bool hist_round = (phase == PH_KEY_HIST || phase == PH_DIST_HIST);
      const bool key_round = (phase == PH_KEY_HIST || phase == PH_KEY_COLLECT);
      const long long n = cl.n_loc;
      const unsigned long long lo_b = hist_round ? cl.wlo : cl.blo, hi_b = hist_round ? cl.whi : cl.bhi;
      const int shift = cl.shift;
      double loc[D], cs[D], lt[SQC_TRI(D)];
#pragma unroll
      for (int d = 0; d < D; ++d) { loc[d] = cl.loc[d]; cs[d] = cl.cshift[d]; }
#pragma unroll
      for (int e = 0; e < SQC_TRI(D); ++e) lt[e] = cl.linv[e];
      const bool need_mom = (phase == PH_KEY_COLLECT) || (phase == PH_DIST_COLLECT && (!cl.is_final || cl.direct));
      const double* rsb = s.rs + cl.off;
      const long long kbase = cl.prefix;                            // key-stream offset of the local segment's first cell
      if (hist_round) { for (int i = tid; i < MCD_NB; i += MCD_THREADS) sh_hist[i] = 0u; }
      __syncthreads();
      double acc[NM];
#pragma unroll
      for (int m = 0; m < NM; ++m) acc[m] = 0.0;
      unsigned int below = 0, nnan = 0, wcount = 0;
      const long long cbase = ((long long)k * regions + region) * MCD_CAPW;

      for (long long ch = c_first; ch < nch; ch += M) {
        const long long e0 = ch * CHUNK;
        double r[CPT][D]; unsigned long long bits[CPT]; bool valid[CPT];
        long long epos[CPT];
#pragma unroll
        for (int it = 0; it < CPT / 2; ++it) {
          const long long e = e0 + (long long)it * (2 * MCD_THREADS) + 2 * tid;
          epos[2 * it] = e; epos[2 * it + 1] = e + 1;
          valid[2 * it] = e < n; valid[2 * it + 1] = (e + 1) < n;
        }
        if (!(key_round && hist_round)) {                           // KEY_HIST needs no residuals
#pragma unroll
          for (int it = 0; it < CPT / 2; ++it) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              double2 v = make_double2(0.0, 0.0);
              if (valid[2 * it]) v = ld_stream2(rsb + (long long)d * s.rs_stride + epos[2 * it]);
              r[2 * it][d] = v.x; r[2 * it + 1][d] = v.y;
            }
          }
        }
        if (key_round) {
          if (MULTI && mb.keys && mb.key_q) {
            // the call's stream is generated once per box: draw `pos` lives on rank pos / Q (NVLink read)
#pragma unroll
            for (int it = 0; it < CPT; ++it) {
              int key = 0;
              if (valid[it]) {
                const long long pos = kbase + epos[it];
                int own = (int)((double)pos * mb.key_qinv);
                long long off = pos - (long long)own * mb.key_q;
                if (off < 0) { --own; off += mb.key_q; } else if (off >= mb.key_q) { ++own; off -= mb.key_q; }
                key = ld_stream_s32(mb.key_peer[own] + off);
              }
              bits[it] = (unsigned long long)(unsigned int)key;
            }
          } else if (mb.keys) {
#pragma unroll
            for (int it = 0; it < CPT; ++it) {
              int key = 0;
              if (valid[it]) key = ld_stream_s32(mb.keys + kbase + epos[it]);
              bits[it] = (unsigned long long)(unsigned int)key;
            }
          } else {
#pragma unroll
            for (int it = 0; it < CPT; ++it)
              bits[it] = (unsigned long long)(unsigned int)counter_key(mb.call_mix, (unsigned long long)(mb.stream_base + kbase + epos[it]));
          }
        } else {
#pragma unroll
          for (int it = 0; it < CPT; ++it) {
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
          for (int it = 0; it < CPT; ++it) {
            if (valid[it]) {
              if (bits[it] < lo_b) ++below;
              else if (bits[it] < hi_b) atomicAdd(&sh_hist[(unsigned int)((bits[it] - lo_b) >> shift)], 1u);
            }
          }
        } else {
#pragma unroll
          for (int it = 0; it < CPT; ++it) {
            const bool in = valid[it] && bits[it] < lo_b;
            const bool cand = valid[it] && bits[it] >= lo_b && bits[it] < hi_b;
            if (need_mom) Mom<D>::add(acc, r[it], cs, in);
            const unsigned m = __ballot_sync(0xffffffffu, cand);
            if (m) {                                                // per-warp region, warp-deterministic order, no atomics
              const unsigned slot = wcount + __popc(m & ((1u << lane) - 1u));
              if (cand) {
                // fine bin inside the candidate bin (the leader only reads this histogram); region overflow is flagged there too
                const int fsh = shift > 10 ? shift - 10 : 0;
                atomicAdd(&mb.fhist[k * (MCD_NF + 4) + (slot < (unsigned)MCD_CAPW ? (unsigned int)((bits[it] - lo_b) >> fsh) : (unsigned)(MCD_NF + 1))], 1u);
                if (slot < (unsigned)MCD_CAPW) {
                  unsigned long long* rec = mb.cand_rec + (cbase + slot) * (2 + D);
                  rec[0] = bits[it]; rec[1] = (unsigned long long)epos[it];
#pragma unroll
                  for (int d = 0; d < D; ++d) rec[2 + d] = (unsigned long long)__double_as_longlong(r[it][d]);
                }
              }
              wcount += __popc(m);
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
      } else {
        if (lane == 0) mb.cand_cnt[(long long)k * regions + region] = wcount;
        if (need_mom) {
          block_reduce_vec<NM>(acc, sh_red, sh_out);
          if (tid < NM) mb.part[((long long)b * K + k) * NM + tid] = sh_out[tid];
          __syncthreads();
        }
      }
      const int nn = warp_sum_i((int)nnan);
      if (lane == 0 && nn) atomicAdd(&W->cnt_nan[k], (unsigned long long)nn);
        __threadfence();
        __syncthreads();
        if (tid == 0) { atomicAdd(&mb.arrive[k], 1u); sh_eloc[k] = ep; }
        __syncthreads();
      }
      if (n_done == K) break;
      if (any_new) { idle = 0; t_idle = 0; }
      else if ((++idle & 0x3fffu) == 0u) {                          // watchdog, same budget as the leaders': a leader died
        const long long now = gtimer();
        if (!t_idle) t_idle = now;
        else if (now - t_idle > mb.xc.budget_ns + 5000000000ll) { if (tid == 0) atomicCAS(&(*mb.status), 0, SQC_D_TIMEOUT); break; }
      }
    }
  }
}

}  // namespace sqc
