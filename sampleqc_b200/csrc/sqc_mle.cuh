// sqc_mle.cuh — the soft EM of fit_sampleqc_mle_cpp (reference src/fit_sampleQC_mle_cpp.cpp:279-357)
// as two fused streaming passes per iteration; the N x K responsibilities are never materialised except
// for the matrix the export returns (gamma_i, :350).
//
//   pass A (MODE_A)   gamma = E(alpha, beta, Sigma, p_k)  calc_expected_gamma_i :151-187  (first call, :313)
//                     -> per-(sample, component) sums  G_jk = sum gamma, X_jk = sum gamma x   (max_alpha_j_given_others :26-64)
//                     -> the log-likelihood of the SAME parameters = like_2 of the previous iteration (:326)
//   pass B (MODE_B)   with the new alpha_j: like_1 (:317), gamma again (:320), and per-component shifted
//                     moments of r = x - alpha_j  (max_beta :65-91, max_sigma :92-129 centred on the NEW beta,
//                     max_p :130-147).  On the last iteration it also stores gamma, its row argmax and its
//                     per-sample sums (extract_z :224-246, extract_p_jk :248-273).
//   MODE_INIT         the same moments from the caller's init_gamma_i (:301-303)
//   MODE_LIKE         log-likelihood only (like_2 of the last iteration)
// Densities are plain (no log-sum-exp), responsibilities are normalised directly (0/0 -> NaN), p_k is
// global — all as in the reference.
#pragma once
#include "sqc_kernels.cuh"

namespace sqc {

enum { MLE_INIT = 0, MLE_A = 1, MLE_B = 2, MLE_LIKE = 3, MLE_B_LAST = 4 };   // MLE_B_LAST: pass B of the last iteration (stores gamma, argmax)

struct MleBufs {
  const double* gamma0 = nullptr;   // N x K column-major, the caller's init_gamma_i
  double* gamma = nullptr;          // N x K column-major, returned gamma_i
  double* t_mom = nullptr;          // K x NMOM(D) x T per-tile partials, tile-minor (pass A uses the first 1 + D of each component)
  double* t_ll = nullptr;           // T per-tile log-likelihood partials
  double* mom = nullptr;            // K x NMOM(D) reduced moments
  double* g_jk = nullptr;           // J x K: sum of gamma per (sample, component)
  double* x_jk = nullptr;           // J x K x D
  double* p_k = nullptr;            // K
  double* coef = nullptr;           // K: p_k * normaliser
  int* iter = nullptr;              // device iteration counter / flags: [0] = iteration, [1] = status, [2] = weights outside the fast densities' range
  int last_store = 0;               // pass B: store gamma / argmax (host sets it for the last iteration: MLE_B_LAST is launched)
};

// warp totals of NM per-lane values, stored to dst[0 .. NM): recursive halving in groups of <= 16 values
template <int NM, int G0 = 0>
__device__ __forceinline__ void warp_sum_store(const double (&acc)[NM], int lane, double* dst) {
  if constexpr (G0 < NM) {
    constexpr int n = (NM - G0 < 16) ? NM - G0 : 16;
    constexpr int NP = n <= 1 ? 1 : n <= 2 ? 2 : n <= 4 ? 4 : n <= 8 ? 8 : 16;
    double v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) v[i] = (i < n) ? acc[G0 + (i < n ? i : 0)] : 0.0;
    warp_sum_vec<NP>(v, lane);
    constexpr int STR = 32 / NP;
    if ((lane & (STR - 1)) == 0 && lane / STR < n) dst[G0 + lane / STR] = v[0];
    warp_sum_store<NM, G0 + 16>(acc, lane, dst);
  }
}

// the reference's arithmetic for one cell (mean formed first, general exp with gradual underflow): the slow path of the
// densities, for cells whose likelihood is not safely inside the normal range and for weights outside the range the fast
// path's exponent arithmetic is safe for.  Plain pointers (see dens_slow).  dcell != NULL: the K densities go to
// dcell[k * dstride].
template <int D> __device__ __noinline__ double mle_dens_slow(const double* xcell, long long N, const double* aj, const double* beta, const double* linv,
                                                              const double* coef, int K, double* dcell, int dstride) {
  double x[D];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = xcell[(long long)d * N];
  double like = 0.0;
  for (int k = 0; k < K; ++k) {
    double v[D], lt[SQC_TRI(D)];
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = x[d] - (aj[d] + beta[k * D + d]);              // mu_ik = alpha_j + beta_k, :168-170
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) lt[r * (r + 1) / 2 + c] = linv[k * D * D + r + D * c];
    const double dens = coef[k] * exp_neg(-0.5 * maha_tri<D>(v, lt));                 // p_k * dmvnorm(x, mu, Sigma), :171
    like += dens;
    if (dcell) dcell[k * dstride] = dens;
  }
  return like;
}

#ifndef SQC_MLE_MINB
#define SQC_MLE_MINB 3
#endif
// One CTA per sample-aligned tile.  The densities use the E-step's arithmetic (sqc_kernels.cuh): affine whitening
// y = L_k^-1 (x - alpha_j) - c_k, exp(-q/2) by table + degree-4 polynomial with the weight p_k (2 pi)^(-D/2) det^(-1/2)
// folded into the table (constants derived once per M-step by mle_update_body), the reference's own arithmetic for the
// cells the fast path is not safe for.  gamma = density / likelihood is one reciprocal per cell and a product per
// component.  Per-component moment sums: registers -> recursive-halving warp sums -> one fixed-order sum over the
// warps per tile (bit-reproducible), tile partials stored tile-minor so that the reductions read them contiguously.
template <int D, int MODE>
__global__ void __launch_bounds__(EST_THREADS, SQC_MLE_MINB) k_mle_pass(Dev s, MleBufs m) {
  constexpr int CPT = TileCfg<D>::CPT, TILE = TileCfg<D>::TILE, SZ = DensK<D>::SZ, NW = EST_THREADS / 32;
  constexpr int NM = (MODE == MLE_A) ? (1 + D) : SQC_NMOM(D), NMF = SQC_NMOM(D);
  extern __shared__ double sh_dyn[];
  const int K = s.K;
  double* cons = sh_dyn;                                     // K * SZ
  double* tab = cons + K * SZ;                               // K * 32
  double* sh_red = tab + K * 32;                             // NW * K * NM
  double* sh_d = sh_red + NW * K * NMF;                      // K * TILE densities (MLE_A / MLE_B)
  __shared__ double sh_w[NW];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int j = s.tile_sample[t], start = s.tile_start[t], len = s.tile_len[t];
  double xv[CPT][D]; bool valid[CPT];
  {
    const double* xb = s.x + (long long)start + tid;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      valid[c] = c * EST_THREADS + tid < len;
#pragma unroll
      for (int d = 0; d < D; ++d) xv[c][d] = valid[c] ? ld_stream(xb + (long long)d * s.N + c * EST_THREADS) : 0.0;
    }
  }
  bool slow_all = false;
  if (MODE != MLE_INIT) {
    for (int idx = tid; idx < K * SZ; idx += EST_THREADS) cons[idx] = s.dens_cons[idx];
    for (int idx = tid; idx < K * 32; idx += EST_THREADS) tab[idx] = s.dens_tab[(idx >> 5) * DENS_TAB + (idx & 31)];
    slow_all = m.iter[2] != 0;
  }
  double aj[D];
#pragma unroll
  for (int d = 0; d < D; ++d) aj[d] = s.alpha[j * D + d];
  __syncthreads();
  // x - alpha_j: what the densities and the moments of pass B / INIT are computed from (pass A sums the raw x)
  double xp[CPT][D];
#pragma unroll
  for (int c = 0; c < CPT; ++c)
#pragma unroll
    for (int d = 0; d < D; ++d) xp[c][d] = xv[c][d] - aj[d];
  double rl[CPT];                                            // 1 / likelihood
#pragma unroll
  for (int c = 0; c < CPT; ++c) rl[c] = 1.0;
  if (MODE != MLE_INIT) {
    double like[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) like[c] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double* pk = cons + k * SZ;
      const double* tk = tab + k * 32;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const double dens = dens_term(maha_affine<D>(xp[c], pk), tk, 0.0);           // p_k * dmvnorm(x, mu, Sigma), :171
        like[c] += dens;
        if (MODE != MLE_LIKE) sh_d[k * TILE + c * EST_THREADS + tid] = dens;
      }
    }
    bool suspect = slow_all; int slowmask = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) suspect = suspect || (valid[c] && !(like[c] >= DENS_SAFE));
    if (suspect) {                                           // rare: one test per thread on the fast path
#pragma unroll
      for (int c = 0; c < CPT; ++c)
        if (valid[c] && (slow_all || !(like[c] >= DENS_SAFE))) {
          like[c] = mle_dens_slow<D>(s.x + (long long)start + c * EST_THREADS + tid, s.N, s.alpha + j * D, s.beta, s.linv, m.coef, K,
                                     (MODE != MLE_LIKE) ? sh_d + c * EST_THREADS + tid : nullptr, TILE);
          slowmask |= 1 << c;
        }
    }
    // log-likelihood of the parameters used (calc_log_likelihood :191-222)
    double ll = sum_log<CPT>(like, valid);
    ll = warp_sum(ll);
    if (lane == 0) sh_w[w] = ll;
    if constexpr (MODE != MLE_LIKE) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) rl[c] = 1.0 / like[c];
      if (slowmask) {                                        // those cells divide as the reference does (:181-183): 0 / 0 = NaN, subnormal likelihoods
#pragma unroll
        for (int c = 0; c < CPT; ++c)
          if (slowmask & (1 << c)) {
            for (int k = 0; k < K; ++k) sh_d[k * TILE + c * EST_THREADS + tid] = sh_d[k * TILE + c * EST_THREADS + tid] / like[c];
            rl[c] = 1.0;
          }
      }
    }
  }
  double gbest[MODE == MLE_B_LAST ? CPT : 1]; int kbest[MODE == MLE_B_LAST ? CPT : 1];
  if constexpr (MODE != MLE_LIKE) {
    // per-component weighted sums over the tile
    if constexpr (MODE == MLE_B_LAST) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) { gbest[c] = -1.0; kbest[c] = -1; }
    }
    for (int k = 0; k < K; ++k) {
      double acc[NM];
#pragma unroll
      for (int e = 0; e < NM; ++e) acc[e] = 0.0;
      double cs[D];
#pragma unroll
      for (int d = 0; d < D; ++d) cs[d] = (MODE == MLE_A) ? 0.0 : s.beta[k * D + d];   // moment shift: current beta_k
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const long long gi = (long long)start + c * EST_THREADS + tid;
        double g;
        if (MODE == MLE_INIT) g = valid[c] ? ld_stream(m.gamma0 + (long long)k * s.N + gi) : 0.0;
        else g = sh_d[k * TILE + c * EST_THREADS + tid] * rl[c];                       // :181-183
        if constexpr (MODE == MLE_B_LAST) {
          if (valid[c]) {
            m.gamma[(long long)k * s.N + gi] = g;
            if (g > gbest[c]) { gbest[c] = g; kbest[c] = k; }                          // extract_z :234-238
          }
        }
        if (!valid[c]) g = 0.0;
        double wv[D];
#pragma unroll
        for (int d = 0; d < D; ++d) wv[d] = (MODE == MLE_A) ? xv[c][d] : xp[c][d] - cs[d];
        acc[0] += g;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[1 + d] = fma(g, wv[d], acc[1 + d]);
        if (MODE != MLE_A) {
          int e = 1 + D;
#pragma unroll
          for (int a = 0; a < D; ++a) {
            const double ga = g * wv[a];
#pragma unroll
            for (int b = a; b < D; ++b) { acc[e] = fma(ga, wv[b], acc[e]); ++e; }
          }
        }
      }
      warp_sum_store<NM>(acc, lane, sh_red + (w * K + k) * NM);
    }
  }
  __syncthreads();
  if constexpr (MODE != MLE_LIKE) {
    for (int idx = tid; idx < K * NM; idx += EST_THREADS) {
      const int k = idx / NM, e = idx % NM;
      double a = 0.0;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) a += sh_red[(ww * K + k) * NM + e];
      m.t_mom[((size_t)k * NMF + e) * s.T + t] = a;
    }
  }
  if (MODE != MLE_INIT && tid == 0) { double a = 0.0; for (int ww = 0; ww < NW; ++ww) a += sh_w[ww]; m.t_ll[t] = a; }
  if constexpr (MODE == MLE_B_LAST) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) if (valid[c]) {
      if (kbest[c] < 0) { atomicCAS(&m.iter[1], 0, SQC_D_NAN); kbest[c] = 0; }  // idx(-1): index out of bounds
      s.z[(long long)start + c * EST_THREADS + tid] = (uint8_t)kbest[c];
    }
  }
}

// lower Cholesky factor and its inverse (small_chol_inv, sqc_common.cuh) with every index static: registers
template <int D>
__device__ __forceinline__ bool chol_inv_static(const double (&m)[D * D], double (&L)[D * D], double (&Li)[D * D]) {
#pragma unroll
  for (int i = 0; i < D * D; ++i) { L[i] = 0.0; Li[i] = 0.0; }
  bool ok = true;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double sdiag = m[c + D * c];
#pragma unroll
    for (int k = 0; k < c; ++k) sdiag -= L[c + D * k] * L[c + D * k];
    ok = ok && (sdiag > 0.0) && (sdiag < CUDART_INF);
    const double d = sqrt(sdiag); L[c + D * c] = d;
#pragma unroll
    for (int r = c + 1; r < D; ++r) {
      double t = m[r + D * c];
#pragma unroll
      for (int k = 0; k < c; ++k) t -= L[r + D * k] * L[c + D * k];
      L[r + D * c] = t / d;
    }
  }
#pragma unroll
  for (int c = 0; c < D; ++c) {                              // forward substitution for the inverse, column by column
    Li[c + D * c] = 1.0 / L[c + D * c];
#pragma unroll
    for (int r = c + 1; r < D; ++r) {
      double t = 0.0;
#pragma unroll
      for (int k = c; k < r; ++k) t -= L[r + D * k] * Li[k + D * c];
      Li[r + D * c] = t / L[r + D * r];
    }
  }
  return ok;
}

// per-sample reduction of the pass-A partials and the alpha_j solve (max_alpha_j_given_others :26-64):
//   denom_j = sum_k G_jk inv(Sigma_k), numer_j = sum_k inv(Sigma_k) (X_jk - G_jk beta_k)
// one warp per sample; WITH_ALPHA = 0 only reduces G_jk (extract_p_jk).  D is a template parameter so that the solve of
// lane 0 runs in registers (with a run-time D it was 3000 instructions on local memory, most of the kernel).
// like != NULL: one more block (the last) sums the tiles' log-likelihood partials into like[slot] (slot < 0: discard)
template <int D, int WITH_ALPHA>
__global__ void __launch_bounds__(256) k_mle_sample(Dev s, MleBufs m, double* like, int slot) {
  if (like && blockIdx.x == gridDim.x - 1) {
    __shared__ double sh[8];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int t = threadIdx.x;
    for (; t + 768 < s.T; t += 1024) { a0 += m.t_ll[t]; a1 += m.t_ll[t + 256]; a2 += m.t_ll[t + 512]; a3 += m.t_ll[t + 768]; }
    for (; t < s.T; t += 256) a0 += m.t_ll[t];
    const double r = block_sum<256>((a0 + a1) + (a2 + a3), sh);
    if (threadIdx.x == 0 && slot >= 0) like[slot] = r;
    return;
  }
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= s.J) return;
  constexpr int NMF = SQC_NMOM(D);
  const int K = s.K;
  const int t0 = s.sample_tile_begin[j], t1 = s.sample_tile_begin[j + 1];
  for (int e = lane; e < K * (1 + D); e += 32) {
    const int k = e / (1 + D), c = e % (1 + D);
    double a = 0.0;
    const double* src = m.t_mom + ((size_t)k * NMF + c) * s.T;
    for (int t = t0; t < t1; ++t) a += src[t];
    if (c == 0) m.g_jk[j * K + k] = a; else m.x_jk[((long long)j * K + k) * D + (c - 1)] = a;
  }
  if (!WITH_ALPHA) return;
  __syncwarp();
  if (lane == 0) {
    double denom[D * D], numer[D], inv[D * D];
#pragma unroll
    for (int e = 0; e < D * D; ++e) denom[e] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) numer[d] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double g = m.g_jk[j * K + k];
      const double* A = s.ainv + k * D * D;
      double v[D];
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = m.x_jk[((long long)j * K + k) * D + d] - g * s.beta[k * D + d];
#pragma unroll
      for (int r = 0; r < D; ++r) {
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) a += A[r + D * c] * v[c];
        numer[r] += a;
      }
#pragma unroll
      for (int e = 0; e < D * D; ++e) denom[e] += g * A[e];
    }
    if (!inv_small<D>(denom, inv)) { atomicCAS(&m.iter[1], 0, SQC_D_SINGULAR); return; }   // inv(), :60
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double a = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) a += inv[r + D * c] * numer[c];
      s.alpha[j * D + r] = a;
    }
  }
}

template <int D> __device__ __forceinline__ void mle_update_body(const Dev& s, const MleBufs& m);
// fixed-order reduction over tiles: block b < K * NMOM sums moment b = k * NMOM + e (contiguous over the tiles); block
// K * NMOM sums the log-likelihood partials into like[slot] (slot < 0: discard).
// UPDATE = 1 (single GPU): the block that finishes last goes on to the M-step update — one launch less (several GPUs:
// k_mle_xchg_update)
template <int D, int UPDATE>
__global__ void __launch_bounds__(256) k_mle_reduce(Dev s, MleBufs m, double* like, int slot) {   // 256 threads: the update's matrices want registers
  __shared__ double sh[8];
  __shared__ int sh_last;
  const int b = blockIdx.x, nm = s.K * SQC_NMOM(D);
  const double* src = (b == nm) ? m.t_ll : m.t_mom + (size_t)b * s.T;
  double acc[8];                                             // independent loads in flight: the kernel is nothing but their latency
#pragma unroll
  for (int u = 0; u < 8; ++u) acc[u] = 0.0;
  int t = threadIdx.x;
  for (; t + 7 * 256 < s.T; t += 8 * 256) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] += src[t + u * 256];
  }
  for (; t < s.T; t += 256) acc[0] += src[t];
  const double r = block_sum<256>(((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7])), sh);
  if (threadIdx.x == 0) {
    if (b == nm) { if (slot >= 0) like[slot] = r; }
    else m.mom[b] = r;
  }
  if (!UPDATE) return;
  if (threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(&m.iter[3], 1);
    sh_last = (prev == (int)gridDim.x - 1);
    if (sh_last) { m.iter[3] = 0; __threadfence(); }
  }
  __syncthreads();
  if (sh_last) mle_update_body<D>(s, m);
}

__global__ void __launch_bounds__(1024) k_mle_reduce_ll(Dev s, MleBufs m, double* like, int slot) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int t = threadIdx.x; t < s.T; t += 1024) a += m.t_ll[t];
  const double r = block_sum<1024>(a, sh);
  if (threadIdx.x == 0 && slot >= 0) like[slot] = r;
}

// beta_k, Sigma_k, p_k from the reduced moments (:65-147), then the derived constants of the passes.  Every thread of ONE
// block calls it; thread k < K works on component k with the small matrices in registers, then the whole block fills
// the K x 32 weighted exponential tables.
template <int D>
__device__ __forceinline__ void mle_update_body(const Dev& s, const MleBufs& m) {
  constexpr int NMF = SQC_NMOM(D), TRI = SQC_TRI(D);
  const int K = s.K;
  __shared__ double sh_tot;
  __shared__ double sh_coef[SQC_MAXK];
  __shared__ int sh_bad;
  if (threadIdx.x == 0) { double a = 0.0; for (int k = 0; k < K; ++k) a += m.mom[k * NMF]; sh_tot = a; sh_bad = 0; }
  __syncthreads();
  const int k = threadIdx.x;
  if (k < K) {
    const double* S = m.mom + k * NMF;
    double mo[NMF];
#pragma unroll
    for (int e = 0; e < NMF; ++e) mo[e] = S[e];
    const double g = mo[0];
    double mm[D], Sg[D * D];
#pragma unroll
    for (int d = 0; d < D; ++d) mm[d] = mo[1 + d] / g;
    {
      int e = 1 + D;
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int b = a; b < D; ++b) { const double v = mo[e] / g - mm[a] * mm[b]; Sg[a + D * b] = v; Sg[b + D * a] = v; ++e; }
    }
    double bk[D];
#pragma unroll
    for (int d = 0; d < D; ++d) { bk[d] = s.beta[k * D + d] + mm[d]; s.beta[k * D + d] = bk[d]; }
#pragma unroll
    for (int e = 0; e < D * D; ++e) s.sigma[k * D * D + e] = Sg[e];
    const double pk = g / sh_tot;
    m.p_k[k] = pk;
    double L[D * D], Li[D * D], Ai[D * D];
    bool good = true;
    if (!inv_small<D>(Sg, Ai)) { atomicCAS(&m.iter[1], 0, SQC_D_SINGULAR); good = false; }
    else if (!chol_inv_static<D>(Sg, L, Li)) { atomicCAS(&m.iter[1], 0, SQC_D_NOT_SPD); good = false; }
    double coef = 0.0;
    if (good) {
      const double det = small_det(Sg, D);
#pragma unroll
      for (int e = 0; e < D * D; ++e) { s.linv[k * D * D + e] = Li[e]; s.ainv[k * D * D + e] = Ai[e]; }
      const double nrm = 1.0 / sqrt(pow(6.283185307179586476925286766559, (double)D) * det);
      s.Pk[k] = nrm;
      coef = pk * nrm;
      m.coef[k] = coef;
      // constants of the passes' fast densities (sqc_kernels.cuh): L^-1 packed, c = L^-1 beta
      double* ck = s.dens_cons + k * (TRI + D);
#pragma unroll
      for (int r = 0; r < D; ++r) {
        double c = 0.0;
#pragma unroll
        for (int cc = 0; cc <= r; ++cc) { const double l = Li[r + D * cc]; ck[r * (r + 1) / 2 + cc] = l; c = fma(l, bk[cc], c); }
        ck[TRI + r] = c;
      }
      if (!(coef > 1e-20 && coef < 1e20)) sh_bad = 1;       // outside the range the fast path's exponent arithmetic is safe for
    }
    sh_coef[k] = coef;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < K * 32; idx += blockDim.x) s.dens_tab[(idx >> 5) * DENS_TAB + (idx & 31)] = sh_coef[idx >> 5] * c_exp_tab[idx & 31];   // the weight folded into the exp table
  if (threadIdx.x == 0) m.iter[2] = sh_bad;
}

// p_jk = G_jk / sum_k G_jk (extract_p_jk :248-273), stored where the robust path keeps p_jk
__global__ void k_mle_pjk(Dev s, MleBufs m) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= s.J) return;
  double tot = 0.0;
  for (int k = 0; k < s.K; ++k) tot += m.g_jk[j * s.K + k];
  for (int k = 0; k < s.K; ++k) s.p_jk[j * s.K + k] = m.g_jk[j * s.K + k] / tot;
}

template <int D> size_t mle_smem(int K, int mode) {
  size_t n = (size_t)K * DensK<D>::SZ + (size_t)K * 32 + (size_t)(EST_THREADS / 32) * K * SQC_NMOM(D);
  if (mode == MLE_A || mode == MLE_B || mode == MLE_B_LAST) n += (size_t)K * TileCfg<D>::TILE;
  return n * sizeof(double);
}
template <int D> void mle_launch_pass(const Dev& s, const MleBufs& m, int mode, cudaStream_t st) {
  const size_t sm = mle_smem<D>(s.K, mode);
  switch (mode) {
    case MLE_INIT: k_mle_pass<D, MLE_INIT><<<s.T, EST_THREADS, sm, st>>>(s, m); break;
    case MLE_A:
      cudaFuncSetAttribute(k_mle_pass<D, MLE_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      k_mle_pass<D, MLE_A><<<s.T, EST_THREADS, sm, st>>>(s, m); break;
    case MLE_B:
      if (m.last_store) {
        cudaFuncSetAttribute(k_mle_pass<D, MLE_B_LAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        k_mle_pass<D, MLE_B_LAST><<<s.T, EST_THREADS, sm, st>>>(s, m);
      } else {
        cudaFuncSetAttribute(k_mle_pass<D, MLE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        k_mle_pass<D, MLE_B><<<s.T, EST_THREADS, sm, st>>>(s, m);
      }
      break;
    default: k_mle_pass<D, MLE_LIKE><<<s.T, EST_THREADS, sm, st>>>(s, m); break;
  }
}

}  // namespace sqc
