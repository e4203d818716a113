// sqc_kernels.cuh — sm_100a kernels of the robust multi-sample GMM EM (reference
// src/fit_sampleQC_robust_cpp.cpp) and the shared streaming machinery.
//
// Kernel families (SQC_KF_* in include/sampleqc_cuda.h):
//   k_estep       E-step pass  (calc_log_likelihood :400-432 + calc_expected_z :344-396 + the (j,k)
//                 sufficient statistics that update_alpha_j :116-157 / update_p_jk :322-341 need)
//   k_partition   like_1 pass + stable partition of residuals x - alpha_j by label
//                 (restrict_to_x_k :265-283 for all k in one pass)
//   k_mcd         persistent cooperative kernel: the whole update_beta_scale_k :289-318 /
//                 fast_mcd :55-108 for all clusters in lock step — exact order-statistic selection by
//                 histogram + candidate resolution instead of arma::sort_index, masked moments
//   k_* (update)  alpha_j solves, p_jk, statistics reduction, control
//   (sqc_mt.cuh)  R's Mersenne-Twister stream -> randperm keys, block-parallel by GF(2) jump-ahead
#pragma once
#include <cooperative_groups.h>

#include "sqc_common.cuh"

namespace cg = cooperative_groups;

namespace sqc {

#ifndef SQC_EST_MINB
#define SQC_EST_MINB 3      // resident CTAs per SM the E-step kernel is compiled for (register cap 80; measured best of 2/3/4)
#endif
#ifndef SQC_PART_MINB
#define SQC_PART_MINB 4     // same for the partition kernel (register cap 64)
#endif
constexpr int EST_THREADS = 256;     // E-step / partition CTAs: one sample-aligned tile each
// cells per thread of a tile, by dimension (register budget)
#ifndef SQC_EST_CPT
#define SQC_EST_CPT 4
#endif
template <int D> struct TileCfg { static constexpr int CPT = (D <= 4) ? SQC_EST_CPT : 2; static constexpr int TILE = EST_THREADS * CPT; };

struct Control {
  int iter;          // EM iterations completed
  int z_delta;       // labels changed by the last E-step (robust_cpp.cpp:519)
  int cont;          // loop condition (:490)
  int status;        // SQC_D_*
  int n_zero;        // n_zeros counter (:380-382)
  int mcd_hit;       // fast_mcd reached mcd_iters (:94-95)
  int mcd_call;      // update_beta_scale_k calls so far (RNG stream position = call * N_total)
  int pad;
  long long mcd_passes, mcd_csteps;
  long long mcd_cells;   // sum over streaming rounds of the cells of the clusters still active
  long long mcd_bytes;   // bytes those rounds had to read (residual columns + keys), DESIGN.md
};

// everything a kernel needs, passed by value
struct Dev {
  int D, J, K, T;             // T = number of tiles
  long long N;                // local cells
  long long N_total;
  long long cell_offset;
  // data
  double* x; uint8_t* z;
  const int* tile_start; const int* tile_len; const int* tile_sample; const int* sample_tile_begin; const int* sample_n;
  // parameters
  double* mu0; double* alpha; double* beta; double* sigma; double* linv; double* ainv; double* logP; double* Pk;
  double* p_jk; double* logp_jk;
  double* dens_cons; double* dens_tab; int* dens_slow;     // density constants of the passes (k_after_mstep): K x DensK::SZ | J x K x 33 | J
  // tile partials / reduced statistics
  double* t_ll1; double* t_ll2; int* t_nk; double* t_sk; int* s_base; int* t_rel; unsigned int* done_ctr; unsigned int* mcd_sync; double* t_col;
  int* n_jk; double* s_jk;
  long long* clu_n; long long* clu_off;        // local cells per cluster, segment offsets in rs
  long long* clu_n_glob; long long* clu_prefix; // cells per cluster over all shards; key-stream offset of the local segment
  // cluster-sorted residuals
  double* rs; long long rs_stride;
  // outputs / control
  double* like1; double* like2; Control* ctl;
  double* track_alpha; double* track_beta; double* track_sigma; double* track_p;
  // fit parameters
  int em_iters, mcd_iters, rng_mode, track; double mcd_alpha; unsigned int seed;
};

// ---------------------------------------------------------------------------------------------
// block-level helpers
// ---------------------------------------------------------------------------------------------
template <int NT> __device__ __forceinline__ double block_sum(double v, double* sh /* NT/32 */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < NT / 32; ++i) r += sh[i];     // fixed order: deterministic
  return r;                                                                // valid in thread 0
}
// exclusive scan of one long long per thread; returns the exclusive prefix, *total = block sum (all threads)
template <int NT> __device__ __forceinline__ long long block_excl_scan(long long v, long long* sh /* NT/32 + 1 */, long long* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  __syncthreads();
  if (lane == 31) sh[w] = inc;
  __syncthreads();
  if (threadIdx.x == 0) { long long run = 0; for (int i = 0; i < NT / 32; ++i) { long long t = sh[i]; sh[i] = run; run += t; } sh[NT / 32] = run; }
  __syncthreads();
  *total = sh[NT / 32];
  return sh[w] + inc - v;
}

// squared Mahalanobis norm |Linv v|^2 with Linv packed lower-triangular row-major
template <int D> __device__ __forceinline__ double maha_tri(const double (&v)[D], const double* lt) {
  double q = 0.0;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double y = 0.0;
#pragma unroll
    for (int c = 0; c <= r; ++c) y = fma(lt[r * (r + 1) / 2 + c], v[c], y);
    q = fma(y, y, q);
  }
  return q;
}

// ---------------------------------------------------------------------------------------------
// per-cell mixture densities (RcppDist::dmvnorm at robust_cpp.cpp:364, :424), K components at
// ~22 fp64 instructions per (cell, component) at D = 3.
//
//   * alpha_j is subtracted from the cell ONCE (xp = x - alpha_j); per component the whitened residual is the affine
//     form  y = L_k^-1 xp - c_k  with  c_k = L_k^-1 beta_k  (the reference forms mu_ik = alpha_j + beta_k first; the
//     two differ by rounding only), so the per-component constants L_k^-1, c_k do not depend on the sample;
//   * exp(-q/2) by table + short polynomial with the reduction done on q itself:  k = round(-q 16/ln2),
//     r' = q + k ln2/16 (|r'| <= ln2/32),  exp(-q/2) = 2^(k/32) exp(-r'/2);  exp(-r'/2) by its degree-4 Taylor
//     polynomial (relative error <= 1.3e-12, far inside the 1e-8 bar on the likelihoods — the labels do not use it);
//     2^(j/32), j = k mod 32, comes from a shared-memory table that already carries the component's weight
//     p_jk (2 pi)^(-D/2) det^(-1/2), and 2^(k div 32) is added to the exponent field: 8 fp64 instructions per term
//     including the accumulation (the general exp_neg below: 15);
//   * cells whose likelihood is not safely inside the normal range (every exponent below ~ -530: cells tens of
//     standard deviations from every component; NaN / Inf input) are recomputed by dens_slow with exp_neg, which
//     follows the reference's arithmetic down to the subnormals.
// ---------------------------------------------------------------------------------------------
template <int D> struct DensK { static constexpr int SZ = SQC_TRI(D) + D; };         // L^-1 packed lower (row-major), c = L^-1 beta
constexpr int DENS_TAB = 33;                      // per (sample, component): 32 weighted table entries w 2^(i/32) + the log weight
constexpr int DENS_KMIN = -900 * 32;              // fast path: exp(-q/2) >= 2^-900; the weights lie in (1e-20, 1e20), so the exponent arithmetic cannot wrap
constexpr double DENS_SAFE = 1e-230;              // a likelihood under this is recomputed on the slow path (clamped terms are < 1e-249)

// shared-memory footprint of the constants of K components (doubles): per-component constants + weight tables + alpha_j
template <int D> __host__ __device__ constexpr int dens_doubles(int K) { return K * DensK<D>::SZ + K * DENS_TAB + D + 2; }

// Copy the constants of sample j's tile into shared memory (they are derived once per M-step by k_after_mstep: no
// arithmetic and no dependent loads in the prologue of the ~10^4 tile CTAs).  Needs a barrier after.
template <int D> __device__ __forceinline__ void stage_dens(const Dev& s, int j, double* cons, double* tab, double* aj, bool with_like) {
  const int K = s.K;
  if (with_like) {
    for (int idx = threadIdx.x; idx < K * DensK<D>::SZ; idx += blockDim.x) cons[idx] = s.dens_cons[idx];
    const double* tj = s.dens_tab + (size_t)j * K * DENS_TAB;
    for (int idx = threadIdx.x; idx < K * DENS_TAB; idx += blockDim.x) tab[idx] = tj[idx];
  }
  if (threadIdx.x < D) aj[threadIdx.x] = s.alpha[j * D + threadIdx.x];
}
// the derivation (one warp per sample; Pk, logP: the component normalisers in shared memory)
template <int D> __device__ __forceinline__ void derive_dens_sample(const Dev& s, int j, int lane, const double* Pk, const double* logP) {
  const int K = s.K;
  double* tj = s.dens_tab + (size_t)j * K * DENS_TAB;
  bool bad = false;
  for (int k = 0; k < K; ++k) {
    const double w = s.p_jk[j * K + k] * Pk[k];              // p_jk * normaliser (:424)
    bad = bad || !(w > 1e-20 && w < 1e20);
    tj[k * DENS_TAB + lane] = w * c_exp_tab[lane];
    if (lane == 0) tj[k * DENS_TAB + 32] = s.logp_jk[j * K + k] + logP[k];   // log(p_jk) + log-normaliser (:364)
  }
  if (lane == 0) s.dens_slow[j] = bad ? 1 : 0;               // weights outside the range the fast path's exponent arithmetic is safe for
}
template <int D> __device__ __forceinline__ void derive_dens_component(const Dev& s, int k, const double* Li /* D x D column-major */) {
  constexpr int TRI = SQC_TRI(D), SZ = DensK<D>::SZ;
  double* ck = s.dens_cons + k * SZ;
  for (int r = 0; r < D; ++r) {
    double c = 0.0;
    for (int cc = 0; cc <= r; ++cc) { const double l = Li[r + D * cc]; ck[r * (r + 1) / 2 + cc] = l; c = fma(l, s.beta[k * D + cc], c); }
    ck[TRI + r] = c;
  }
}

// the reference's arithmetic for one cell (mean formed first, general exp with gradual underflow): the slow path.
// Takes plain pointers: a reference to the by-value kernel parameter block would force a copy of it into local memory.
template <int D> __device__ __noinline__ double dens_slow(const double* xcell, long long N, const double* aj, const double* beta, const double* linv,
                                                          const double* pj, const double* Pk, int K) {
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = xcell[(long long)d * N];
  double like = 0.0;
  for (int k = 0; k < K; ++k) {
    double v[D], lt[SQC_TRI(D)];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = x[d] - (aj[d] + beta[k * D + d]);
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) lt[r * (r + 1) / 2 + c] = linv[k * D * D + r + D * c];
    const double q = maha_tri<D>(v, lt);
    like += pj[k] * (Pk[k] * exp_neg(-0.5 * q));
  }
  return like;
}

// squared whitened norm of xp under component constants pk = {L^-1 packed, c}
template <int D> __device__ __forceinline__ double maha_affine(const double (&xp)[D], const double* pk) {
  double q = 0.0;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double y = -pk[SQC_TRI(D) + r];
#pragma unroll
    for (int c = 0; c <= r; ++c) y = fma(pk[r * (r + 1) / 2 + c], xp[c], y);
    q = fma(y, y, q);
  }
  return q;
}
// like += w 2^(j/32) (table entry) * 2^n * exp(-r'/2).  The exponent k is clamped to [DENS_KMIN, 0] as an integer: a
// cell that far out (or with q beyond the range the rounding trick is exact for) contributes ~2^-900 w times a bounded
// polynomial — nothing beside a regular term, and under DENS_SAFE when every component is that far away, which sends
// the cell to dens_slow; NaN / Inf distances give a NaN term and take the same route.
__device__ __forceinline__ double dens_term(double q, const double* tabk, double like) {
  const double magic = 6755399441055744.0;
  const double t = fma(q, -23.083120654223414192, magic);             // -16 / ln 2
  const double kf = t - magic;
  const double rp = fma(kf, 4.33216987849965803e-02, q);              // ln2 / 16:  r' = q + k ln2/16
  double p = fma(rp, 1.0 / 384.0, -1.0 / 48.0);
  p = fma(p, rp, 0.125);
  p = fma(p, rp, -0.5);
  p = fma(p, rp, 1.0);
  const int k = min(max(__double2loint(t), DENS_KMIN), 0);
  const double w = tabk[k & 31];
  const double sc = __hiloint2double(__double2hiint(w) + ((k >> 5) << 20), __double2loint(w));
  return fma(sc, p, like);
}

// ---------------------------------------------------------------------------------------------
// E-step pass.  One CTA per sample-aligned tile.
//   like_2 (:506) + new labels (calc_expected_z :344-396, :509) + z_delta (:519) + the CHANGE of the per-(tile,
//   component) statistics n_k, sum_k x that update_alpha_j :116-157 / update_p_jk :322-341 are computed from:
//   a cell that moves from component a to b subtracts its x from (tile, a) and adds it to (tile, b), applied in
//   (cell slot, lane) order per warp and the warps in warp order — bit-reproducible.  Warps without a changed
//   label (the common case after the first iterations) skip that part.  The initial statistics come from
//   k_tile_stats (one full pass per fit).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(EST_THREADS, (D <= 4) ? SQC_EST_MINB : 2) k_estep(Dev s) {
  constexpr int CPT = TileCfg<D>::CPT, SZ = DensK<D>::SZ, NW = EST_THREADS / 32;
  extern __shared__ double sh_dyn[];
  const int K = s.K, KD = K * (D + 1);
  double* cons = sh_dyn;                                     // K * SZ
  double* tab = cons + K * SZ;                               // K * DENS_TAB
  double* aj = tab + K * DENS_TAB;                           // D (+ pad)
  double* dw = aj + D + 2;                                   // NW * K * (D + 1) statistic deltas (zeroed lazily)
  __shared__ double sh_w[NW];
  __shared__ int sh_i[NW * 2];
  __shared__ int sh_flag[2];                                 // [1] some warp changed labels
  if (!s.ctl->cont) return;
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int j = s.tile_sample[t], start = s.tile_start[t], len = s.tile_len[t];
  // the cell loads go out first; the constants are copied while they are in flight
  double xp[CPT][D]; int zo[CPT]; bool valid[CPT];
  {
    const double* xb = s.x + (long long)start + tid;         // one base per thread: the cell slots and columns are constant offsets
    const uint8_t* zb = s.z + (long long)start + tid;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      valid[c] = c * EST_THREADS + tid < len;
#pragma unroll
      for (int d = 0; d < D; ++d) xp[c][d] = valid[c] ? ld_stream(xb + (long long)d * s.N + c * EST_THREADS) : 0.0;
      zo[c] = valid[c] ? (int)ld_stream_u8(zb + c * EST_THREADS) : 0;
    }
  }
  const bool slow_tile = s.dens_slow[j] != 0;
  if (tid < 2) sh_flag[tid] = 0;
  stage_dens<D>(s, j, cons, tab, aj, true);
  __syncthreads();
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double a = aj[d];
#pragma unroll
    for (int c = 0; c < CPT; ++c) xp[c][d] -= a;
  }
  double best[CPT], like[CPT]; int zn[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) { best[c] = -CUDART_INF; like[c] = 0.0; zn[c] = 0; }
  for (int k = 0; k < K; ++k) {
    const double* pk = cons + k * SZ;
    const double* tk = tab + k * DENS_TAB;
    const double lc = tk[32];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const double q = maha_affine<D>(xp[c], pk);
      const double l = fma(-0.5, q, lc);                     // log(p_jk) + dmvnorm(log = TRUE), :364
      if (l > best[c]) { best[c] = l; zn[c] = k; }           // strict >, first max, :375-378
      like[c] = dens_term(q, tk, like[c]);                   // p_jk * dmvnorm(log = FALSE), :424
    }
  }
  int nzero = 0, delta = 0;
  bool anyc = false, suspect = slow_tile;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    suspect = suspect || (valid[c] && !(like[c] >= DENS_SAFE));
    if (valid[c] && !(best[c] > -CUDART_INF)) nzero += K;    // like_max stays -inf through all K tests, :380-382
    const bool chg = valid[c] && zn[c] != zo[c];
    if (chg) { s.z[(long long)start + c * EST_THREADS + tid] = (uint8_t)zn[c]; ++delta; }
    anyc = anyc || chg;
  }
  if (suspect) {                                             // rare: one test per thread on the fast path
#pragma unroll
    for (int c = 0; c < CPT; ++c)
      if (valid[c] && (slow_tile || !(like[c] >= DENS_SAFE)))
        like[c] = dens_slow<D>(s.x + (long long)start + c * EST_THREADS + tid, s.N, s.alpha + j * D, s.beta, s.linv, s.p_jk + j * K, s.Pk, K);
  }
  double ll = sum_log<CPT>(like, valid);                     // :428
  // labels that moved: statistics of the old component lose the cell, those of the new one gain it
  if (__any_sync(0xffffffffu, anyc)) {
    double* mydw = dw + (size_t)w * KD;
    bool chg[CPT]; int nchg = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) { chg[c] = valid[c] && zn[c] != zo[c]; nchg += __popc(__ballot_sync(0xffffffffu, chg[c])); }
    if (nchg <= 6) {
      // a few cells: applied one by one in (cell slot, lane) order
      for (int e = lane; e < KD; e += 32) mydw[e] = 0.0;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        unsigned mk = __ballot_sync(0xffffffffu, chg[c]);
        while (mk) {
          const int l = __ffs(mk) - 1; mk &= mk - 1;
          if (lane == l) {
            double* a = mydw + zo[c] * (D + 1); double* b = mydw + zn[c] * (D + 1);
            a[0] -= 1.0; b[0] += 1.0;
            const long long gi = (long long)start + c * EST_THREADS + tid;
#pragma unroll
            for (int d = 0; d < D; ++d) { const double xv = s.x[(long long)d * s.N + gi]; a[1 + d] -= xv; b[1 + d] += xv; }
          }
          __syncwarp();
        }
      }
    } else {
      // many cells (the first iterations): per component a signed masked sum over the warp, fixed shuffle tree
      double xv[CPT][D];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const long long gi = (long long)start + c * EST_THREADS + tid;
#pragma unroll
        for (int d = 0; d < D; ++d) xv[c][d] = chg[c] ? s.x[(long long)d * s.N + gi] : 0.0;
      }
      constexpr int NP = (D + 1 <= 2) ? 2 : (D + 1 <= 4) ? 4 : (D + 1 <= 8) ? 8 : 16;
      for (int k = 0; k < K; ++k) {
        double sv[NP];
#pragma unroll
        for (int e = 0; e < NP; ++e) sv[e] = 0.0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const double sg = chg[c] ? ((zn[c] == k ? 1.0 : 0.0) - (zo[c] == k ? 1.0 : 0.0)) : 0.0;
          sv[0] += sg;
#pragma unroll
          for (int d = 0; d < D; ++d) sv[1 + d] = fma(sg, xv[c][d], sv[1 + d]);
        }
        warp_sum_vec<NP>(sv, lane);
        constexpr int STR = 32 / NP;
        if ((lane & (STR - 1)) == 0 && lane / STR < D + 1) mydw[k * (D + 1) + lane / STR] = sv[0];
      }
    }
    if (lane == 0) { sh_flag[1] = 1; sh_i[NW + w] = 1; }
  } else if (lane == 0) sh_i[NW + w] = 0;
  ll = warp_sum(ll); delta = warp_sum_i(delta); nzero = warp_sum_i(nzero);
  if (lane == 0) { sh_w[w] = ll; sh_i[w] = delta; }
  if (nzero && lane == 0) atomicAdd(&s.ctl->n_zero, nzero);
  __syncthreads();
  if (sh_flag[1]) {
    for (int idx = tid; idx < KD; idx += EST_THREADS) {
      double acc = 0.0;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) if (sh_i[NW + ww]) acc += dw[(size_t)ww * KD + idx];
      const int k = idx / (D + 1), e = idx % (D + 1);
      if (e == 0) s.t_nk[t * K + k] += (int)acc;
      else s.t_sk[((long long)t * K + k) * D + (e - 1)] += acc;
    }
  }
  if (tid == 0) {
    double a = 0.0; int dl = 0;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) { a += sh_w[ww]; dl += sh_i[ww]; }
    s.t_ll2[t] = a;
    if (dl) atomicAdd(&s.ctl->z_delta, dl);
  }
}

// statistics of the CURRENT labels, one full pass (initial labelling): per tile and component the count and the sum
// of x, reduced over the tile in fixed order
template <int D>
__global__ void __launch_bounds__(EST_THREADS) k_tile_stats(Dev s) {
  constexpr int CPT = TileCfg<D>::CPT, NW = EST_THREADS / 32;
  extern __shared__ double sh_dyn[];
  double* sh_red = sh_dyn;                                   // NW * K * (D + 1)
  if (!s.ctl->cont) return;
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int start = s.tile_start[t], len = s.tile_len[t];
  double xv[CPT][D]; int zn[CPT]; bool valid[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int i = c * EST_THREADS + tid;
    valid[c] = i < len;
    const long long gi = (long long)start + (valid[c] ? i : 0);
#pragma unroll
    for (int d = 0; d < D; ++d) xv[c][d] = ld_stream(s.x + (long long)d * s.N + gi);
    zn[c] = (int)ld_stream_u8(s.z + gi);
  }
  constexpr int NP = (D + 1 <= 2) ? 2 : (D + 1 <= 4) ? 4 : (D + 1 <= 8) ? 8 : 16;
  for (int k = 0; k < s.K; ++k) {
    double sv[NP];
#pragma unroll
    for (int e = 0; e < NP; ++e) sv[e] = 0.0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const bool in = valid[c] && zn[c] == k;
      sv[0] += in ? 1.0 : 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) sv[1 + d] += in ? xv[c][d] : 0.0;
    }
    warp_sum_vec<NP>(sv, lane);
    constexpr int STR = 32 / NP;
    if ((lane & (STR - 1)) == 0 && lane / STR < D + 1) sh_red[(w * s.K + k) * (D + 1) + lane / STR] = sv[0];
  }
  __syncthreads();
  for (int idx = tid; idx < s.K * (D + 1); idx += EST_THREADS) {
    const int k = idx / (D + 1), e = idx % (D + 1);
    double acc = 0.0;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) acc += sh_red[(ww * s.K + k) * (D + 1) + e];
    if (e == 0) s.t_nk[t * s.K + k] = (int)acc;
    else s.t_sk[((long long)t * s.K + k) * D + (e - 1)] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// like_1 + stable partition by label.  One CTA per tile.
//   r_i = x_i - alpha_{j}  written to rs[d][clu_off[z] + (cells of z in earlier samples / tiles) + rank in tile]
//   (restrict_to_x_k :265-283 keeps original cell order inside each cluster; so does this).
//   LIKE = 1: also the log-likelihood with the new alpha_j and the previous beta/Sigma/p_jk (:499).
// ---------------------------------------------------------------------------------------------
template <int D, int LIKE>
__global__ void __launch_bounds__(EST_THREADS, (D <= 4) ? SQC_PART_MINB : 3) k_partition(Dev s) {   // D > 4: 3 CTAs / 80 registers, measured on C4
  constexpr int CPT = TileCfg<D>::CPT, SZ = DensK<D>::SZ, NW = EST_THREADS / 32;
  extern __shared__ double sh_dyn[];
  const int K = s.K;
  double* cons = sh_dyn;                                     // K * SZ
  double* tab = cons + K * SZ;                               // K * DENS_TAB
  double* aj = tab + K * DENS_TAB;                           // D (+ pad)
  int* sh_cnt = (int*)(aj + D + 2);                          // CPT * NW * K exclusive prefixes
  __shared__ double sh_w[NW];
  if (!s.ctl->cont) return;
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (t == 0 && tid < 2 * SQC_MAXK + 1) s.mcd_sync[tid] = 0u;      // epochs, arrivals and status word of the k_mcd launch that follows
  const int j = s.tile_sample[t], start = s.tile_start[t], len = s.tile_len[t];
  double xp[CPT][D]; int zo[CPT]; bool valid[CPT];
  {
    const double* xb = s.x + (long long)start + tid;
    const uint8_t* zb = s.z + (long long)start + tid;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      valid[c] = c * EST_THREADS + tid < len;
#pragma unroll
      for (int d = 0; d < D; ++d) xp[c][d] = valid[c] ? ld_stream(xb + (long long)d * s.N + c * EST_THREADS) : 0.0;
      zo[c] = valid[c] ? (int)ld_stream_u8(zb + c * EST_THREADS) : -1;
    }
  }
  stage_dens<D>(s, j, cons, tab, aj, LIKE != 0);
  // per (iteration c, warp, component) counts -> exclusive prefix in tile order (c, warp, lane)
  int myrank[CPT];
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const unsigned b = __ballot_sync(0xffffffffu, zo[c] == k);
      if (zo[c] == k) myrank[c] = __popc(b & ((1u << lane) - 1u));
      if (lane == 0) sh_cnt[(c * NW + w) * K + k] = __popc(b);
    }
  }
  __syncthreads();
  if (tid < K) {                                             // tiny serial scan per component
    // base of this tile inside the cluster segment: cells of the earlier samples (s_base) + of the earlier tiles of this sample (t_rel)
    int run = s.s_base[j * K + tid] + s.t_rel[t * K + tid];
    for (int e = 0; e < CPT * NW; ++e) { int v = sh_cnt[e * K + tid]; sh_cnt[e * K + tid] = run; run += v; }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double a = aj[d];
#pragma unroll
    for (int c = 0; c < CPT; ++c) xp[c][d] -= a;             // :274
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < CPT; ++c) if (valid[c]) {
    const long long pos = s.clu_off[zo[c]] + sh_cnt[(c * NW + w) * K + zo[c]] + myrank[c];
#pragma unroll
    for (int d = 0; d < D; ++d) s.rs[(long long)d * s.rs_stride + pos] = xp[c][d];
  }
  if (LIKE) {
    double like[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) like[c] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double* pk = cons + k * SZ;
      const double* tk = tab + k * DENS_TAB;
#pragma unroll
      for (int c = 0; c < CPT; ++c) like[c] = dens_term(maha_affine<D>(xp[c], pk), tk, like[c]);
    }
    const bool slow_tile = s.dens_slow[j] != 0;
    bool suspect = slow_tile;
#pragma unroll
    for (int c = 0; c < CPT; ++c) suspect = suspect || (valid[c] && !(like[c] >= DENS_SAFE));
    if (suspect) {
#pragma unroll
      for (int c = 0; c < CPT; ++c)
        if (valid[c] && (slow_tile || !(like[c] >= DENS_SAFE)))
          like[c] = dens_slow<D>(s.x + (long long)start + c * EST_THREADS + tid, s.N, s.alpha + j * D, s.beta, s.linv, s.p_jk + j * K, s.Pk, K);
    }
    double ll = sum_log<CPT>(like, valid);
    ll = warp_sum(ll);
    if (lane == 0) sh_w[w] = ll;
    __syncthreads();
    if (tid == 0) { double a = 0.0; for (int ww = 0; ww < NW; ++ww) a += sh_w[ww]; s.t_ll1[t] = a; }
  }
}

// ---------------------------------------------------------------------------------------------
// initialisation passes: column sums (arma::mean :478), centring (centre_x helpers.cpp:105-115)
// fused with the per-sample sums of the centred data (init_alpha_j helpers.cpp:117-140)
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(EST_THREADS) k_colsum(Dev s) {
  constexpr int CPT = TileCfg<D>::CPT, NW = EST_THREADS / 32;
  __shared__ double sh[NW * D];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int start = s.tile_start[t], len = s.tile_len[t];
  double acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.0;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int i = c * EST_THREADS + tid;
    if (i < len) {
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] += ld_stream(s.x + (long long)d * s.N + start + i);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) { acc[d] = warp_sum(acc[d]); if (lane == 0) sh[w * D + d] = acc[d]; }
  __syncthreads();
  if (tid < D) { double a = 0.0; for (int ww = 0; ww < NW; ++ww) a += sh[ww * D + tid]; s.t_col[(long long)t * D + tid] = a; }
}
// mu0[d] = sum over tiles (fixed order) / N_total ; single block
__global__ void __launch_bounds__(1024) k_colmean(Dev s) {
  __shared__ double sh[32];
  for (int d = 0; d < s.D; ++d) {
    double a = 0.0;
    for (int t = threadIdx.x; t < s.T; t += 1024) a += s.t_col[(long long)t * s.D + d];
    double r = block_sum<1024>(a, sh);
    if (threadIdx.x == 0) s.mu0[d] = r;       // sum; divided after the (optional) cross-GPU reduction
    __syncthreads();
  }
}
__global__ void k_colmean_finish(Dev s) { if (threadIdx.x < s.D) s.mu0[threadIdx.x] /= (double)s.N_total; }

template <int D>
__global__ void __launch_bounds__(EST_THREADS) k_centre(Dev s) {
  constexpr int CPT = TileCfg<D>::CPT, NW = EST_THREADS / 32;
  __shared__ double sh[NW * D];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int start = s.tile_start[t], len = s.tile_len[t];
  double acc[D], mu[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { acc[d] = 0.0; mu[d] = s.mu0[d]; }
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int i = c * EST_THREADS + tid;
    if (i < len) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        double* p = s.x + (long long)d * s.N + start + i;
        const double v = *p - mu[d];
        *p = v; acc[d] += v;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) { acc[d] = warp_sum(acc[d]); if (lane == 0) sh[w * D + d] = acc[d]; }
  __syncthreads();
  if (tid < D) { double a = 0.0; for (int ww = 0; ww < NW; ++ww) a += sh[ww * D + tid]; s.t_col[(long long)t * D + tid] = a; }
}
// alpha_j = per-sample mean of the centred data; one warp per sample
__global__ void k_init_alpha(Dev s) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= s.J) return;
  if (lane < s.D) {
    double a = 0.0;
    for (int t = s.sample_tile_begin[j]; t < s.sample_tile_begin[j + 1]; ++t) a += s.t_col[(long long)t * s.D + lane];
    s.alpha[j * s.D + lane] = a / (double)s.sample_n[j];
  }
}

// ---------------------------------------------------------------------------------------------
// statistics reduction after an E-step / stats pass:
//   every block: one warp per sample sums its tiles' partials -> n_jk, s_jk
//   the block that finishes last: per-component exclusive scan of n_jk over the samples (-> s_base: the
//   position of a sample's first cell of component k inside the cluster-sorted segment), cluster sizes and
//   segment offsets.  k_partition adds the counts of the earlier tiles of its own sample.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void em_control_block(Dev& s, bool with_like = true, bool with_track_alpha = true);
template <int D> __device__ __forceinline__ void alpha_from_stats(Dev& s, int j);
// exclusive block scans of up to four long long per thread at once (one per component); *total = block sums
template <int NT> __device__ __forceinline__ void block_excl_scan4(long long (&v)[4], long long (*sh)[NT / 32 + 1], long long (&total)[4]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long inc[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    inc[q] = v[q];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, inc[q], o); if (lane >= o) inc[q] += t; }
  }
  __syncthreads();
  if (lane == 31) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sh[q][w] = inc[q];
  }
  __syncthreads();
  if (threadIdx.x < 4) { long long run = 0; for (int i = 0; i < NT / 32; ++i) { long long t = sh[threadIdx.x][i]; sh[threadIdx.x][i] = run; run += t; } sh[threadIdx.x][NT / 32] = run; }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) { total[q] = sh[q][NT / 32]; v[q] = sh[q][w] + inc[q] - v[q]; }
}
// flags: bit 0 (single GPU) = the end-of-iteration bookkeeping rides along (:511-521; one launch less): one extra block
//        sums the tile log-likelihoods beside the sample blocks, the block that finishes last closes the iteration;
//        bit 1 (with bit 0) = alpha_j of the NEXT iteration (update_alpha_j, :496) rides along too: the warp that has just
//        reduced sample j's statistics solves for its alpha_j when the loop goes on — the loop condition of :490 is
//        known at entry (z_delta of the E-step before, the iteration count, the status).
//        (D <= POST_ALPHA_MAXD: beyond it the D x D solve does not fit the 64 registers of a 1024-thread block).
constexpr int POST_ALPHA_MAXD = 4;
template <int D>
__global__ void __launch_bounds__(1024) k_post_stats(Dev s, int n_sample_blocks, int flags) {
  __shared__ long long sh_scan[4][33];
  __shared__ double sh_like[32];
  __shared__ int sh_last;
  if (!s.ctl->cont) return;
  const int tid = threadIdx.x, lane = tid & 31, K = s.K;
  const bool fuse_ctl = flags & 1, fuse_alpha = (D <= POST_ALPHA_MAXD) && (flags & 3) == 3;
  const int it = s.ctl->iter;
  const bool goes_on = (s.ctl->z_delta > 0) && (it + 1 < s.em_iters) && (s.ctl->status == 0);       // what em_control_block will find
  const int n_blocks = n_sample_blocks + (fuse_ctl ? 1 : 0);
  if ((int)blockIdx.x == n_sample_blocks) {                    // the likelihood block (same sums, same order as em_control_block)
    double a1 = 0.0, a2 = 0.0;
    for (int t = tid; t < s.T; t += 1024) { a1 += s.t_ll1[t]; a2 += s.t_ll2[t]; }
    const double r1 = block_sum<1024>(a1, sh_like);
    const double r2 = block_sum<1024>(a2, sh_like);
    if (tid == 0) { s.like1[it] = r1; s.like2[it] = r2; }
  } else {
    const int j = blockIdx.x * 32 + (tid >> 5);
    if (j < s.J) {
      const int t0 = s.sample_tile_begin[j], t1 = s.sample_tile_begin[j + 1];
      const int ne = K * (D + 1);
      // eight tiles' partials in flight at a time (the t_rel stores would otherwise order the t_nk loads one behind the
      // other); the sums run in tile order as before
      for (int e = lane; e < ne; e += 32) {
        const int k = e / (D + 1), c = e % (D + 1);
        if (c == 0) {
          int a = 0;
          for (int tb = t0; tb < t1; tb += 8) {
            int v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (tb + u < t1) ? __ldcg(s.t_nk + (tb + u) * K + k) : 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) if (tb + u < t1) { s.t_rel[(tb + u) * K + k] = a; a += v[u]; }
          }
          s.n_jk[j * K + k] = a;
        } else {
          double a = 0.0;
          for (int tb = t0; tb < t1; tb += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (tb + u < t1) ? __ldcg(s.t_sk + ((long long)(tb + u) * K + k) * D + (c - 1)) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) if (tb + u < t1) a += v[u];
          }
          s.s_jk[((long long)j * K + k) * D + (c - 1)] = a;
        }
      }
      if constexpr (D <= POST_ALPHA_MAXD) if (fuse_alpha) {
        if (s.track && it < s.em_iters && lane < D) s.track_alpha[(long long)it * s.J * D + j + s.J * lane] = s.alpha[j * D + lane];
        __syncwarp();                                          // the statistics above were written by this warp's lanes
        if (goes_on && lane == 0) alpha_from_stats<D>(s, j);
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) sh_last = (atomicAdd(s.done_ctr, 1u) == (unsigned)n_blocks - 1u) ? 1 : 0;
  __syncthreads();
  if (!sh_last) return;
  __threadfence();
  // exclusive scan over the samples for four components at a time (thread t owns a contiguous run of samples)
  const int per = (s.J + 1023) / 1024;
  const int j0 = min(tid * per, s.J), j1 = min(j0 + per, s.J);
  long long off_run = 0, key_run = 0;
  for (int k0 = 0; k0 < K; k0 += 4) {
    long long run[4], total[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      run[q] = 0;
      if (k0 + q < K) for (int j = j0; j < j1; ++j) run[q] += __ldcg(s.n_jk + j * K + k0 + q);
    }
    block_excl_scan4<1024>(run, sh_scan, total);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (k0 + q < K) {
        long long r = run[q];
        for (int j = j0; j < j1; ++j) { s.s_base[j * K + k0 + q] = (int)r; r += __ldcg(s.n_jk + j * K + k0 + q); }
        // single GPU: the key stream of one update_beta_scale_k call runs over the clusters in order (SURVEY Appendix
        // A3); the multi-GPU exchange overwrites clu_n_glob / clu_prefix with the global values
        if (tid == 0) { s.clu_n[k0 + q] = total[q]; s.clu_off[k0 + q] = off_run; s.clu_n_glob[k0 + q] = total[q]; s.clu_prefix[k0 + q] = key_run; }
        off_run += (total[q] + 31) & ~31ll;        // 256-byte aligned cluster segments
        key_run += total[q];
      }
    }
    __syncthreads();
  }
  if (tid == 0) *s.done_ctr = 0u;
  // every other block of this launch has passed its loop-flag test (and read iter / z_delta / status) by now
  if (fuse_ctl) { __syncthreads(); em_control_block(s, false, !fuse_alpha); }
}

// alpha_j from the (j,k) statistics: update_alpha_j :116-157 with the per-cell sums factored out
//   denom_j = sum_k n_jk A_k,  numer_j = sum_k A_k (s_jk - n_jk beta_k),  alpha_j = inv(denom_j) numer_j
// one thread per sample, D compile-time so that the small matrices live in registers
template <int D>
__device__ __forceinline__ bool inv_small(const double (&m)[D * D], double (&inv)[D * D]) {   // Gauss-Jordan, partial pivoting (arma::inv)
  double a[D * D];
#pragma unroll
  for (int i = 0; i < D * D; ++i) { a[i] = m[i]; inv[i] = 0.0; }
#pragma unroll
  for (int i = 0; i < D; ++i) inv[i + D * i] = 1.0;
  bool ok = true;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    int piv = c; double best = fabs(a[c + D * c]);
#pragma unroll
    for (int r = c + 1; r < D; ++r) { const double v = fabs(a[r + D * c]); if (v > best) { best = v; piv = r; } }
    if (!(best > 0.0) || !(best < CUDART_INF)) ok = false;
#pragma unroll
    for (int r = c + 1; r < D; ++r) {
      if (piv == r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double t = a[c + D * k]; a[c + D * k] = a[r + D * k]; a[r + D * k] = t;
          t = inv[c + D * k]; inv[c + D * k] = inv[r + D * k]; inv[r + D * k] = t;
        }
      }
    }
    const double d = a[c + D * c];
#pragma unroll
    for (int k = 0; k < D; ++k) { a[c + D * k] /= d; inv[c + D * k] /= d; }
#pragma unroll
    for (int r = 0; r < D; ++r) {
      if (r != c) {
        const double f = a[r + D * c];
#pragma unroll
        for (int k = 0; k < D; ++k) { a[r + D * k] -= f * a[c + D * k]; inv[r + D * k] -= f * inv[c + D * k]; }
      }
    }
  }
  return ok;
}
template <int D>
__device__ __forceinline__ void alpha_from_stats(Dev& s, int j) {
  const int K = s.K;
  double denom[D * D], numer[D], inv[D * D];
#pragma unroll
  for (int e = 0; e < D * D; ++e) denom[e] = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) numer[d] = 0.0;
  for (int k = 0; k < K; ++k) {
    const double n = (double)s.n_jk[j * K + k];
    const double* A = s.ainv + k * D * D;
    double v[D];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = s.s_jk[((long long)j * K + k) * D + d] - n * s.beta[k * D + d];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double a = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) a += A[r + D * c] * v[c];
      numer[r] += a;
    }
#pragma unroll
    for (int e = 0; e < D * D; ++e) denom[e] += n * A[e];
  }
  if (!inv_small<D>(denom, inv)) { atomicCAS(&s.ctl->status, 0, SQC_D_SINGULAR); return; }   // arma::inv throws, :153
#pragma unroll
  for (int r = 0; r < D; ++r) {
    double a = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) a += inv[r + D * c] * numer[c];
    s.alpha[j * D + r] = a;
  }
}
template <int D>
__global__ void k_update_alpha(Dev s) {
  if (!s.ctl->cont) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s.J) return;
  alpha_from_stats<D>(s, j);
}
// p_jk = (1 + n_jk) / (K + n_j), update_p_jk :322-341
__global__ void k_update_p(Dev s) {
  if (!s.ctl->cont) return;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s.J) return;
  long long nj = s.K;
  for (int k = 0; k < s.K; ++k) nj += s.n_jk[j * s.K + k];
  for (int k = 0; k < s.K; ++k) {
    const double p = (1.0 + (double)s.n_jk[j * s.K + k]) / (double)nj;
    s.p_jk[j * s.K + k] = p; s.logp_jk[j * s.K + k] = log(p);
  }
}

// end-of-iteration bookkeeping (:499,:506,:511-521): deterministic sums of the tile log-likelihoods,
// loop condition, optional tracking.  single block.
__device__ __forceinline__ void em_control_block(Dev& s, bool with_like, bool with_track_alpha) {       // one block of 1024 threads
  __shared__ double sh[32];
  Control* c = s.ctl;
  const int it = c->iter;
  double r1 = 0.0, r2 = 0.0;
  if (with_like) {                                 // (otherwise like1[it] / like2[it] were written by the likelihood block of k_post_stats)
    double a1 = 0.0, a2 = 0.0;
    for (int t = threadIdx.x; t < s.T; t += 1024) { a1 += s.t_ll1[t]; a2 += s.t_ll2[t]; }
    r1 = block_sum<1024>(a1, sh);
    r2 = block_sum<1024>(a2, sh);
  }
  if (s.track && it < s.em_iters) {
    if (with_track_alpha) for (int e = threadIdx.x; e < s.J * s.D; e += 1024) { int j = e / s.D, d = e % s.D; s.track_alpha[(long long)it * s.J * s.D + j + s.J * d] = s.alpha[e]; }
    for (int e = threadIdx.x; e < s.K * s.D; e += 1024) { int k = e / s.D, d = e % s.D; s.track_beta[(long long)it * s.K * s.D + k + s.K * d] = s.beta[e]; }
    for (int e = threadIdx.x; e < s.K * s.D * s.D; e += 1024) s.track_sigma[(long long)it * s.K * s.D * s.D + e] = s.sigma[e];
    for (int e = threadIdx.x; e < s.J * s.K; e += 1024) { int j = e / s.K, k = e % s.K; s.track_p[(long long)it * s.J * s.K + j + s.J * k] = s.p_jk[e]; }
  }
  if (threadIdx.x == 0) {
    // multi-GPU: like sums and z_delta are made global by the host-side exchange before this point
    if (with_like) { s.like1[it] = r1; s.like2[it] = r2; }
    c->iter = it + 1;
    __threadfence();
    c->cont = (c->z_delta > 0) && (it + 1 < s.em_iters) && (c->status == 0);    // :490
    c->z_delta = 0;
  }
}
__global__ void __launch_bounds__(1024) k_em_control(Dev s) {
  if (!s.ctl->cont) return;
  em_control_block(s);
}

// ---------------------------------------------------------------------------------------------
// R's Mersenne-Twister on one CTA: 227-wide wavefronts of the recurrence,
// ping-pong state in shared memory.  st = 624 state words + position.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b) {
  uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// R set.seed(seed) state initialisation (Appendix A1)
__global__ void k_mt_seed(uint32_t* st, uint32_t seed) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t v = seed;
    for (int j = 0; j < 50; ++j) v = 69069u * v + 1u;
    v = 69069u * v + 1u;                        // i_seed[0] (overwritten by mti = 624)
    for (int j = 0; j < 624; ++j) { v = 69069u * v + 1u; st[j] = v; }
    st[624] = 624u;
  }
}

}  // namespace sqc
