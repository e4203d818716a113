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

enum { MLE_INIT = 0, MLE_A = 1, MLE_B = 2, MLE_LIKE = 3 };

struct MleBufs {
  const double* gamma0 = nullptr;   // N x K column-major, the caller's init_gamma_i
  double* gamma = nullptr;          // N x K column-major, returned gamma_i
  double* t_mom = nullptr;          // T x K x NMOM(D) per-tile partials (pass A uses the first 1 + D of each)
  double* t_ll = nullptr;           // T per-tile log-likelihood partials
  double* mom = nullptr;            // K x NMOM(D) reduced moments
  double* g_jk = nullptr;           // J x K: sum of gamma per (sample, component)
  double* x_jk = nullptr;           // J x K x D
  double* p_k = nullptr;            // K
  double* coef = nullptr;           // K: p_k * normaliser
  int* iter = nullptr;              // device iteration counter / flags: [0] = iteration, [1] = status
  int last_store = 0;               // pass B: store gamma / argmax (host sets it for the last iteration)
};

template <int D> struct MlePar { static constexpr int SZ = D + SQC_TRI(D) + 1; };

#ifndef SQC_MLE_MINB
#define SQC_MLE_MINB 3
#endif
template <int D, int MODE>
__global__ void __launch_bounds__(EST_THREADS, SQC_MLE_MINB) k_mle_pass(Dev s, MleBufs m) {
  constexpr int CPT = TileCfg<D>::CPT, TILE = TileCfg<D>::TILE, SZ = MlePar<D>::SZ, NW = EST_THREADS / 32;
  constexpr int NM = (MODE == MLE_A) ? (1 + D) : SQC_NMOM(D);
  extern __shared__ double sh_dyn[];
  double* sh_par = sh_dyn;                                   // K * SZ
  double* sh_red = sh_par + s.K * SZ;                        // NW * NM
  double* sh_d = sh_red + NW * SQC_NMOM(D);                  // K * TILE densities (not for MLE_LIKE / MLE_INIT)
  __shared__ double sh_w[NW];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, K = s.K;
  const int j = s.tile_sample[t], start = s.tile_start[t], len = s.tile_len[t];
  if (MODE != MLE_INIT) {
    for (int idx = tid; idx < K * SZ; idx += EST_THREADS) {
      const int k = idx / SZ, e = idx % SZ;
      double v;
      if (e < D) v = s.alpha[j * D + e] + s.beta[k * D + e];          // mu_ik = alpha_j + beta_k, :168-170
      else if (e < D + SQC_TRI(D)) {
        int tt = e - D, r = 0;
        while ((r + 1) * (r + 2) / 2 <= tt) ++r;
        v = s.linv[k * D * D + r + D * (tt - r * (r + 1) / 2)];
      } else v = m.coef[k];
      sh_par[idx] = v;
    }
  }
  double xv[CPT][D]; bool valid[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int i = c * EST_THREADS + tid;
    valid[c] = i < len;
    const long long gi = (long long)start + (valid[c] ? i : 0);
#pragma unroll
    for (int d = 0; d < D; ++d) xv[c][d] = ld_stream(s.x + (long long)d * s.N + gi);
  }
  __syncthreads();
  double like[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) like[c] = 0.0;
  if (MODE != MLE_INIT) {
    for (int k = 0; k < K; ++k) {
      const double* pk = sh_par + k * SZ;
      double mu[D], lt[SQC_TRI(D)];
#pragma unroll
      for (int d = 0; d < D; ++d) mu[d] = pk[d];
#pragma unroll
      for (int e = 0; e < SQC_TRI(D); ++e) lt[e] = pk[D + e];
      const double coef = pk[D + SQC_TRI(D)];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        double v[D];
#pragma unroll
        for (int d = 0; d < D; ++d) v[d] = xv[c][d] - mu[d];
        const double dens = coef * exp_neg(-0.5 * maha_tri<D>(v, lt));   // p_k * dmvnorm(x, mu, Sigma), :171
        like[c] += dens;
        if (MODE != MLE_LIKE) sh_d[k * TILE + c * EST_THREADS + tid] = dens;
      }
    }
    // log-likelihood of the parameters used (calc_log_likelihood :191-222)
    double ll = sum_log<CPT>(like, valid);
    ll = warp_sum(ll);
    if (lane == 0) sh_w[w] = ll;
    __syncthreads();
    if (tid == 0) { double a = 0.0; for (int ww = 0; ww < NW; ++ww) a += sh_w[ww]; m.t_ll[t] = a; }
    if (MODE == MLE_LIKE) return;
  }
  // per-component weighted sums over the tile
  double aj[D];
#pragma unroll
  for (int d = 0; d < D; ++d) aj[d] = (MODE == MLE_A) ? 0.0 : s.alpha[j * D + d];
  double gbest[CPT]; int kbest[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) { gbest[c] = -1.0; kbest[c] = -1; }
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
      else g = sh_d[k * TILE + c * EST_THREADS + tid] / like[c];               // :181-183
      if (MODE == MLE_B && m.last_store && valid[c]) {
        m.gamma[(long long)k * s.N + gi] = g;
        if (g > gbest[c]) { gbest[c] = g; kbest[c] = k; }                      // extract_z :234-238
      }
      if (!valid[c]) g = 0.0;
      double wv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) wv[d] = (xv[c][d] - aj[d]) - cs[d];
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
#pragma unroll
    for (int e = 0; e < NM; ++e) { const double a = warp_sum(acc[e]); if (lane == 0) sh_red[w * NM + e] = a; }
    __syncthreads();
    if (tid < NM) {
      double a = 0.0;
#pragma unroll
      for (int ww = 0; ww < NW; ++ww) a += sh_red[ww * NM + tid];
      m.t_mom[((long long)t * K + k) * SQC_NMOM(D) + tid] = a;
    }
    __syncthreads();
  }
  if (MODE == MLE_B && m.last_store) {
#pragma unroll
    for (int c = 0; c < CPT; ++c) if (valid[c]) {
      if (kbest[c] < 0) { atomicCAS(&m.iter[1], 0, SQC_D_NAN); kbest[c] = 0; }  // idx(-1): index out of bounds
      s.z[(long long)start + c * EST_THREADS + tid] = (uint8_t)kbest[c];
    }
  }
}

// per-sample reduction of the pass-A partials and the alpha_j solve (max_alpha_j_given_others :26-64):
//   denom_j = sum_k G_jk inv(Sigma_k), numer_j = sum_k inv(Sigma_k) (X_jk - G_jk beta_k)
// one warp per sample; WITH_ALPHA = 0 only reduces G_jk (extract_p_jk)
template <int WITH_ALPHA>
__global__ void k_mle_sample(Dev s, MleBufs m) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= s.J) return;
  const int D = s.D, K = s.K, NMF = SQC_NMOM(D);
  const int t0 = s.sample_tile_begin[j], t1 = s.sample_tile_begin[j + 1];
  for (int e = lane; e < K * (1 + D); e += 32) {
    const int k = e / (1 + D), c = e % (1 + D);
    double a = 0.0;
    for (int t = t0; t < t1; ++t) a += m.t_mom[((long long)t * K + k) * NMF + c];
    if (c == 0) m.g_jk[j * K + k] = a; else m.x_jk[((long long)j * K + k) * D + (c - 1)] = a;
  }
  if (!WITH_ALPHA) return;
  __syncwarp();
  if (lane == 0) {
    double denom[SQC_MAXD * SQC_MAXD], numer[SQC_MAXD], inv[SQC_MAXD * SQC_MAXD];
    for (int e = 0; e < D * D; ++e) denom[e] = 0.0;
    for (int d = 0; d < D; ++d) numer[d] = 0.0;
    for (int k = 0; k < K; ++k) {
      const double g = m.g_jk[j * K + k];
      const double* A = s.ainv + k * D * D;
      double v[SQC_MAXD];
      for (int d = 0; d < D; ++d) v[d] = m.x_jk[((long long)j * K + k) * D + d] - g * s.beta[k * D + d];
      for (int r = 0; r < D; ++r) { double a = 0.0; for (int c = 0; c < D; ++c) a += A[r + D * c] * v[c]; numer[r] += a; }
      for (int e = 0; e < D * D; ++e) denom[e] += g * A[e];
    }
    if (!small_inv(denom, D, inv)) { atomicCAS(&m.iter[1], 0, SQC_D_SINGULAR); return; }   // inv(), :60
    for (int r = 0; r < D; ++r) { double a = 0.0; for (int c = 0; c < D; ++c) a += inv[r + D * c] * numer[c]; s.alpha[j * D + r] = a; }
  }
}

// fixed-order reduction over tiles: block k sums the NMOM moments of component k; block K sums the
// log-likelihood partials into like[slot] (slot < 0: discard)
__global__ void __launch_bounds__(1024) k_mle_reduce(Dev s, MleBufs m, double* like, int slot) {
  __shared__ double sh[32];
  const int K = s.K, NMF = SQC_NMOM(s.D);
  if ((int)blockIdx.x == K) {
    double a = 0.0;
    for (int t = threadIdx.x; t < s.T; t += 1024) a += m.t_ll[t];
    const double r = block_sum<1024>(a, sh);
    if (threadIdx.x == 0 && slot >= 0) like[slot] = r;
    return;
  }
  const int k = blockIdx.x;
  for (int e = 0; e < NMF; ++e) {
    double a = 0.0;
    for (int t = threadIdx.x; t < s.T; t += 1024) a += m.t_mom[((long long)t * K + k) * NMF + e];
    const double r = block_sum<1024>(a, sh);
    if (threadIdx.x == 0) m.mom[k * NMF + e] = r;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_mle_reduce_ll(Dev s, MleBufs m, double* like, int slot) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int t = threadIdx.x; t < s.T; t += 1024) a += m.t_ll[t];
  const double r = block_sum<1024>(a, sh);
  if (threadIdx.x == 0 && slot >= 0) like[slot] = r;
}

// beta_k, Sigma_k, p_k from the reduced moments (:65-147), then the derived density constants
__global__ void k_mle_update(Dev s, MleBufs m, int derive_only) {
  const int D = s.D, K = s.K, NMF = SQC_NMOM(D);
  __shared__ double tot;
  if (threadIdx.x == 0) { double a = 0.0; for (int k = 0; k < K; ++k) a += m.mom[k * NMF]; tot = a; }
  __syncthreads();
  const int k = threadIdx.x;
  if (k >= K) return;
  if (!derive_only) {
    const double* S = m.mom + k * NMF;
    const double g = S[0];
    double mm[SQC_MAXD];
    for (int d = 0; d < D; ++d) mm[d] = S[1 + d] / g;
    int e = 1 + D;
    for (int a = 0; a < D; ++a)
      for (int b = a; b < D; ++b) {
        const double v = S[e++] / g - mm[a] * mm[b];
        s.sigma[k * D * D + a + D * b] = v; s.sigma[k * D * D + b + D * a] = v;
      }
    for (int d = 0; d < D; ++d) s.beta[k * D + d] = s.beta[k * D + d] + mm[d];
    m.p_k[k] = g / tot;
  }
  const double* Sg = s.sigma + k * D * D;
  double L[SQC_MAXD * SQC_MAXD], Li[SQC_MAXD * SQC_MAXD], Ai[SQC_MAXD * SQC_MAXD];
  if (!small_inv(Sg, D, Ai)) { atomicCAS(&m.iter[1], 0, SQC_D_SINGULAR); return; }
  if (!small_chol_inv(Sg, D, L, Li)) { atomicCAS(&m.iter[1], 0, SQC_D_NOT_SPD); return; }
  const double det = small_det(Sg, D);
  for (int e2 = 0; e2 < D * D; ++e2) { s.linv[k * D * D + e2] = Li[e2]; s.ainv[k * D * D + e2] = Ai[e2]; }
  s.Pk[k] = 1.0 / sqrt(pow(6.283185307179586476925286766559, (double)D) * det);
  m.coef[k] = m.p_k[k] * s.Pk[k];
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
  size_t n = (size_t)K * MlePar<D>::SZ + (size_t)(EST_THREADS / 32) * SQC_NMOM(D);
  if (mode == MLE_A || mode == MLE_B) n += (size_t)K * TileCfg<D>::TILE;
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
      cudaFuncSetAttribute(k_mle_pass<D, MLE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      k_mle_pass<D, MLE_B><<<s.T, EST_THREADS, sm, st>>>(s, m); break;
    default: k_mle_pass<D, MLE_LIKE><<<s.T, EST_THREADS, sm, st>>>(s, m); break;
  }
}

}  // namespace sqc
