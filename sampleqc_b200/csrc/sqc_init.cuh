// sqc_init.cuh — the device side of .init_clusts (reference R/fit_sampleqc.R:277-331), the step in front of the fit:
//   * grp_medians (:286-289): the median of every QC metric inside every sample  ->  k_sample_median
//   * the classification of ALL N median-centred cells under the K-component Gaussian mixture that was fitted on a
//     1e4-cell subsample (:303-316: summaryMclustBIC(...)$z and $classification)       ->  k_gmm_classify
// The subsample fit itself (mclust "EVE", 1e4 x D) stays on the host side of the boundary: it is O(1e4), and its
// optimiser is third-party code; what is O(N) — medians and the classify-all pass — runs here.
#pragma once
#include "sqc_common.cuh"

namespace sqc {

// order-preserving map double -> uint64 (negative values: all bits flipped, others: sign bit set)
__device__ __forceinline__ unsigned long long f64_order_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_order_key(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

// One CTA per (sample j, metric d): exact median of the sample's cells (stats::median: the middle order statistic, or
// the mean of the two middle ones).  Radix select over the 64-bit order keys, 11 bits per pass (6 passes, the segment
// stays in L2), then one pass for the upper middle element when n is even.  Empty samples give NaN (median(numeric(0)) = NA).
constexpr int MED_THREADS = 256, MED_BITS = 11, MED_BINS = 1 << MED_BITS;
__global__ void __launch_bounds__(MED_THREADS) k_sample_median(const double* x /* N x D */, const long long* bounds /* J + 1 */, long long N, int J,
                                                             double* med /* J x D col-major */, int* nan_flag) {
  __shared__ unsigned int hist[MED_BINS];
  __shared__ unsigned int sh_scan[MED_THREADS];
  __shared__ unsigned long long sh_prefix;
  __shared__ long long sh_rank;
  __shared__ unsigned long long sh_min[MED_THREADS / 32];
  __shared__ unsigned int sh_cnt[MED_THREADS / 32];
  const int j = blockIdx.x % J, d = blockIdx.x / J, tid = threadIdx.x;
  const long long b0 = bounds[j], n = bounds[j + 1] - b0;
  if (n <= 0) { if (tid == 0) med[j + (long long)J * d] = __longlong_as_double(0x7ff8000000000000ll); return; }
  const double* seg = x + (long long)d * N + b0;
  const long long k1 = (n - 1) / 2, k2 = n / 2;                  // 0-based ranks of the middle element(s)
  if (tid == 0) { sh_prefix = 0ull; sh_rank = k1; }
  __syncthreads();
  // ---- radix select of rank k1 ----
  for (int shift = 64 - MED_BITS; ; shift -= MED_BITS) {
    const int sh = shift < 0 ? 0 : shift;                       // last pass: the remaining low bits (64 = 5 * 11 + 9)
    const int nb = shift < 0 ? (1 << (MED_BITS + shift)) : MED_BINS;
    const unsigned long long himask = (sh + (shift < 0 ? MED_BITS + shift : MED_BITS)) >= 64 ? 0ull : ~((1ull << (sh + (shift < 0 ? MED_BITS + shift : MED_BITS))) - 1ull);
    for (int i = tid; i < MED_BINS; i += MED_THREADS) hist[i] = 0u;
    __syncthreads();
    const unsigned long long prefix = sh_prefix;
    for (long long i = tid; i < n; i += MED_THREADS) {
      const double v = seg[i];
      if (v != v) { *nan_flag = 1; continue; }
      const unsigned long long key = f64_order_key(v);
      if ((key & himask) == prefix) atomicAdd(&hist[(unsigned int)((key >> sh) & (unsigned long long)(nb - 1))], 1u);
    }
    __syncthreads();
    // locate the bin that holds the wanted rank: per-thread partial sums of MED_BINS / MED_THREADS bins, scan, refine
    constexpr int PER = MED_BINS / MED_THREADS;
    unsigned int mine = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) mine += hist[tid * PER + i];
    sh_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
      long long want = sh_rank, run = 0;
      int t = 0;
      while (t < MED_THREADS - 1 && run + (long long)sh_scan[t] <= want) { run += sh_scan[t]; ++t; }
      int bin = t * PER;
      while (bin < t * PER + PER - 1 && run + (long long)hist[bin] <= want) { run += hist[bin]; ++bin; }
      sh_rank = want - run;
      sh_prefix = prefix | ((unsigned long long)bin << sh);
    }
    __syncthreads();
    if (shift <= 0) break;
  }
  const unsigned long long key1 = sh_prefix;
  double v1 = f64_from_order_key(key1), v2 = v1;
  if (k2 != k1) {
    // rank k2 = k1 + 1: the same value if enough cells tie with it, else the smallest value above it
    unsigned int cnt_le = 0; unsigned long long mn = ~0ull;
    for (long long i = tid; i < n; i += MED_THREADS) {
      const double v = seg[i];
      if (v != v) continue;
      const unsigned long long key = f64_order_key(v);
      if (key <= key1) ++cnt_le; else if (key < mn) mn = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, o);
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, mn, o);
      mn = other < mn ? other : mn;
    }
    if ((tid & 31) == 0) { sh_cnt[tid >> 5] = cnt_le; sh_min[tid >> 5] = mn; }
    __syncthreads();
    if (tid == 0) {
      long long c = 0; unsigned long long m = ~0ull;
      for (int w = 0; w < MED_THREADS / 32; ++w) { c += sh_cnt[w]; m = sh_min[w] < m ? sh_min[w] : m; }
      if (c <= k2) v2 = f64_from_order_key(m);
    }
  }
  if (tid == 0) med[j + (long long)J * d] = (k2 != k1) ? (v1 + v2) * 0.5 : v1;     // mean of the two middle values (exact halving)
}
// the same over explicit runs (samples in any order, each one contiguous run of cells): the resident pipeline
__global__ void __launch_bounds__(MED_THREADS) k_sample_median_runs(const double* x /* N x D */, const long long* bounds /* 2 J: start and end of every sample's run */, long long N, int J,
                                                             double* med /* J x D col-major */, int* nan_flag) {
  __shared__ unsigned int hist[MED_BINS];
  __shared__ unsigned int sh_scan[MED_THREADS];
  __shared__ unsigned long long sh_prefix;
  __shared__ long long sh_rank;
  __shared__ unsigned long long sh_min[MED_THREADS / 32];
  __shared__ unsigned int sh_cnt[MED_THREADS / 32];
  const int j = blockIdx.x % J, d = blockIdx.x / J, tid = threadIdx.x;
  const long long b0 = bounds[2 * j], n = bounds[2 * j + 1] - b0;
  if (n <= 0) { if (tid == 0) med[j + (long long)J * d] = __longlong_as_double(0x7ff8000000000000ll); return; }
  const double* seg = x + (long long)d * N + b0;
  const long long k1 = (n - 1) / 2, k2 = n / 2;                  // 0-based ranks of the middle element(s)
  if (tid == 0) { sh_prefix = 0ull; sh_rank = k1; }
  __syncthreads();
  // ---- radix select of rank k1 ----
  for (int shift = 64 - MED_BITS; ; shift -= MED_BITS) {
    const int sh = shift < 0 ? 0 : shift;                       // last pass: the remaining low bits (64 = 5 * 11 + 9)
    const int nb = shift < 0 ? (1 << (MED_BITS + shift)) : MED_BINS;
    const unsigned long long himask = (sh + (shift < 0 ? MED_BITS + shift : MED_BITS)) >= 64 ? 0ull : ~((1ull << (sh + (shift < 0 ? MED_BITS + shift : MED_BITS))) - 1ull);
    for (int i = tid; i < MED_BINS; i += MED_THREADS) hist[i] = 0u;
    __syncthreads();
    const unsigned long long prefix = sh_prefix;
    for (long long i = tid; i < n; i += MED_THREADS) {
      const double v = seg[i];
      if (v != v) { *nan_flag = 1; continue; }
      const unsigned long long key = f64_order_key(v);
      if ((key & himask) == prefix) atomicAdd(&hist[(unsigned int)((key >> sh) & (unsigned long long)(nb - 1))], 1u);
    }
    __syncthreads();
    // locate the bin that holds the wanted rank: per-thread partial sums of MED_BINS / MED_THREADS bins, scan, refine
    constexpr int PER = MED_BINS / MED_THREADS;
    unsigned int mine = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) mine += hist[tid * PER + i];
    sh_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
      long long want = sh_rank, run = 0;
      int t = 0;
      while (t < MED_THREADS - 1 && run + (long long)sh_scan[t] <= want) { run += sh_scan[t]; ++t; }
      int bin = t * PER;
      while (bin < t * PER + PER - 1 && run + (long long)hist[bin] <= want) { run += hist[bin]; ++bin; }
      sh_rank = want - run;
      sh_prefix = prefix | ((unsigned long long)bin << sh);
    }
    __syncthreads();
    if (shift <= 0) break;
  }
  const unsigned long long key1 = sh_prefix;
  double v1 = f64_from_order_key(key1), v2 = v1;
  if (k2 != k1) {
    // rank k2 = k1 + 1: the same value if enough cells tie with it, else the smallest value above it
    unsigned int cnt_le = 0; unsigned long long mn = ~0ull;
    for (long long i = tid; i < n; i += MED_THREADS) {
      const double v = seg[i];
      if (v != v) continue;
      const unsigned long long key = f64_order_key(v);
      if (key <= key1) ++cnt_le; else if (key < mn) mn = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt_le += __shfl_xor_sync(0xffffffffu, cnt_le, o);
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, mn, o);
      mn = other < mn ? other : mn;
    }
    if ((tid & 31) == 0) { sh_cnt[tid >> 5] = cnt_le; sh_min[tid >> 5] = mn; }
    __syncthreads();
    if (tid == 0) {
      long long c = 0; unsigned long long m = ~0ull;
      for (int w = 0; w < MED_THREADS / 32; ++w) { c += sh_cnt[w]; m = sh_min[w] < m ? sh_min[w] : m; }
      if (c <= k2) v2 = f64_from_order_key(m);
    }
  }
  if (tid == 0) med[j + (long long)J * d] = (k2 != k1) ? (v1 + v2) * 0.5 : v1;     // mean of the two middle values (exact halving)
}
// rows idx[i] of x minus the medians of their samples: the subsample of .init_clusts (:291-296), n x D column-major
__global__ void k_gather_centred(const double* x, const int* groups, const double* med, const long long* idx, long long n, long long N, int J, int D, double* out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long c = idx[i]; const int j = groups[c];
  for (int d = 0; d < D; ++d) out[i + n * d] = x[(long long)d * N + c] - med[j + (long long)J * d];
}


// Posterior of every median-centred cell under the mixture sum_k pro_k N(mean_k, Sigma_k) and its arg max
// (mclust: z = estep(...)$z, classification = map(z): the first maximum).  Parameters per cluster in shared memory:
// mean (D), inverse Cholesky factor (lower, D x D), c_k = log pro_k - sum log diag L_k - D/2 log 2 pi.
template <int D>
__global__ void __launch_bounds__(256) k_gmm_classify(const double* x, const int* groups, long long N, int J, int K, const double* med /* J x D */,
                                                      const double* par /* K x (D + D*D + 1) */, double* gamma /* N x K or null */, int* z /* N or null */) {
  extern __shared__ double sh[];
  constexpr int P = D + D * D + 1;
  for (int e = threadIdx.x; e < K * P; e += blockDim.x) sh[e] = par[e];
  __syncthreads();
  const long long gsz = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gsz) {
    const int j = ld_stream_s32(groups + i);
    double xc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xc[d] = ld_stream(x + (long long)d * N + i) - med[j + (long long)J * d];   // :290
    double lmax = -CUDART_INF; int best = 0;
    double ld[SQC_MAXK];
    for (int k = 0; k < K; ++k) {
      const double* pk = sh + k * P;
      double v[D];
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = xc[d] - pk[d];
      double q = 0.0;
#pragma unroll
      for (int r = 0; r < D; ++r) {                 // w = Linv v (lower triangular), q = |w|^2
        double t = 0.0;
#pragma unroll
        for (int c = 0; c <= r; ++c) t += pk[D + r + D * c] * v[c];
        q += t * t;
      }
      const double l = pk[D + D * D] - 0.5 * q;
      ld[k] = l;
      if (l > lmax) { lmax = l; best = k; }         // strict >: the first maximum, like map() / which.max
    }
    if (gamma) {
      double den = 0.0;
      for (int k = 0; k < K; ++k) { ld[k] = exp(ld[k] - lmax); den += ld[k]; }
      for (int k = 0; k < K; ++k) gamma[i + N * k] = ld[k] / den;
    }
    if (z) z[i] = best;
  }
}

}  // namespace sqc
