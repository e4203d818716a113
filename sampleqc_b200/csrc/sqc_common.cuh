// sqc_common.cuh — device-side building blocks shared by all kernels of libsampleqc_cuda.so.
//
// Data layout in HBM:
//   x      D columns of N fp64, structure-of-arrays (R's column-major matrix as is), centred in place
//   z      N uint8 labels
//   tiles  sample-aligned runs of <= TILE cells; a tile never crosses a sample boundary, so the sample
//          index (and with it alpha_j, p_j.) is uniform per tile and `groups` is never streamed
//   rs     D columns of cluster-sorted residuals r = x - alpha_j (stable partition by label)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#define SQC_MAXD 8
#define SQC_MAXK 16
#define SQC_TRI(D) ((D) * ((D) + 1) / 2)
#define SQC_NMOM(D) (1 + (D) + SQC_TRI(D))          // count, sum, upper-triangular scatter

// device-side status (mirrors include/sampleqc_cuda.h)
#define SQC_D_OK 0
#define SQC_D_BAD_ARG 1
#define SQC_D_EMPTY_CLUSTER 2
#define SQC_D_SUBSET_SIZE 3
#define SQC_D_NOT_SPD 4
#define SQC_D_SINGULAR 5
#define SQC_D_NAN 6
#define SQC_D_SELECT 7
#define SQC_D_PEER 8          // multi-GPU: another rank failed
#define SQC_D_TIMEOUT 9       // a wait on another CTA / another rank ran out of its time budget
#ifndef SQC_WAIT_BUDGET_NS
#define SQC_WAIT_BUDGET_NS 60000000000ll   // one budget for every in-kernel wait (leader, workers, exchanges), %globaltimer based
#endif

namespace sqc {

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// ---------------------------------------------------------------------------------------------
// streaming loads: the cell matrix is read once per pass, keep it out of L1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream2(const double* p) {          // 128-bit, 16-byte aligned
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned ld_stream_u8(const uint8_t* p) {
  unsigned v;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ---------------------------------------------------------------------------------------------
// bulk asynchronous copies global -> shared (TMA engine, cp.async.bulk; SASS UBLKCP) completed on an mbarrier:
// the tile movers of the E-step / partition passes.  One thread arms the barrier with the byte count and issues
// the copies; every thread of the CTA waits on the barrier's phase parity before it reads the stage.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {   // 16-byte aligned, bytes % 16 == 0
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// exp(x) for x <= 0 (the density exponents -q/2):  x = (32 n + j) ln2/32 + r,  exp(x) = 2^n * 2^(j/32) * exp(r).
// k = 32 n + j by the 1.5 * 2^52 rounding trick, two-step Cody-Waite reduction (|r| <= ln2/64), degree-6 Taylor
// polynomial, one table look-up (32 doubles, L1-resident), scaling by exponent arithmetic.  <= 2 ulp from the
// correctly rounded value, including the gradual underflow below -708 (two-step scaling); 11 fp64 instructions.
__device__ const double c_exp_tab[32] = {
  1, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
  1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
  1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
  1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
  1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
  1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.6457554781539649,
  1.681792830507429, 1.7186192981224779, 1.7562521603732995, 1.7947090750031072,
  1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};
__device__ __forceinline__ double exp_neg(double x) {
  const double magic = 6755399441055744.0;
  const double t = fma(x, 46.166241308446828384, magic);              // 32 / ln 2
  const int k = __double2loint(t);
  const double kf = t - magic;
  double r = fma(kf, -2.16608493865351192653e-02, x);                 // ln2/32, high part (32 significant bits)
  r = fma(kf, -5.96317165397058656257e-12, r);                        // ln2/32, low part
  double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  p *= __ldg(&c_exp_tab[k & 31]);
  const int n = k >> 5;
  if (!(x >= -746.0)) return (x != x) ? x : 0.0;                      // underflow to zero (also -inf); NaN stays NaN
  if (n >= -1020) return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  const double s1 = __hiloint2double((n + 1000 + 1023) << 20, 0);     // 2^(n + 1000), normal
  return (p * s1) * 9.33263618503218878990e-302;                      // * 2^-1000: the hardware rounds the subnormal
}
// sum of log(v[c]) over the valid cells of a thread: one log of the product when that cannot over- / underflow
template <int CPT>
__device__ __forceinline__ double sum_log(const double (&v)[CPT], const bool (&valid)[CPT]) {
  double prod = 1.0; bool safe = true;
#pragma unroll
  for (int c = 0; c < CPT; ++c) { const double u = valid[c] ? v[c] : 1.0; prod *= u; safe = safe && (u > 1e-70) && (u < 1e70); }
  if (safe) return log(prod);
  double a = 0.0;
#pragma unroll
  for (int c = 0; c < CPT; ++c) if (valid[c]) a += log(v[c]);
  return a;
}

// sum over the warp of NV doubles per lane with recursive halving: every step halves the number of values a lane
// carries, so the shuffle count is about 2 NV instead of 5 NV.  NP = NV padded to a power of two (<= 16).
// On return, lane l with (l & (32 / NP - 1)) == 0 holds the total of value index  l / (32 / NP)  in v[0].
template <int NP>
__device__ __forceinline__ void warp_sum_vec(double (&v)[NP], int lane) {
  static_assert(NP == 1 || NP == 2 || NP == 4 || NP == 8 || NP == 16, "NP must be a power of two <= 16");
  int mask = 16;
#pragma unroll
  for (int half = NP / 2; half >= 1; half >>= 1) {
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = up ? v[i] : v[i + half];
      const double keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
    mask >>= 1;
  }
#pragma unroll
  for (; mask >= 1; mask >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], mask);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(0xffffffffu, v); }    // one REDUX instead of five shuffle + add pairs
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// tiny dense linear algebra on D x D column-major matrices, run by ONE thread
// (stand-ins for arma::det / inv / chol used at robust_cpp.cpp:19,77,91,127,153)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline double small_det(const double* m, int n) {
  if (n == 1) return m[0];
  if (n == 2) return m[0] * m[3] - m[2] * m[1];
  if (n == 3)
    return m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
  double lu[SQC_MAXD * SQC_MAXD];
  for (int i = 0; i < n * n; ++i) lu[i] = m[i];
  double det = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c; double best = fabs(lu[c + n * c]);
    for (int r = c + 1; r < n; ++r) if (fabs(lu[r + n * c]) > best) { best = fabs(lu[r + n * c]); piv = r; }
    if (best == 0.0) return 0.0;
    if (piv != c) { for (int k = 0; k < n; ++k) { double t = lu[c + n * k]; lu[c + n * k] = lu[piv + n * k]; lu[piv + n * k] = t; } det = -det; }
    det *= lu[c + n * c];
    for (int r = c + 1; r < n; ++r) {
      double f = lu[r + n * c] / lu[c + n * c];
      for (int k = c + 1; k < n; ++k) lu[r + n * k] -= f * lu[c + n * k];
    }
  }
  return det;
}
// general inverse by Gauss-Jordan with partial pivoting; false if singular
__host__ __device__ inline bool small_inv(const double* m, int n, double* inv) {
  double a[SQC_MAXD * SQC_MAXD];
  for (int i = 0; i < n * n; ++i) { a[i] = m[i]; inv[i] = 0.0; }
  for (int i = 0; i < n; ++i) inv[i + n * i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c; double best = fabs(a[c + n * c]);
    for (int r = c + 1; r < n; ++r) if (fabs(a[r + n * c]) > best) { best = fabs(a[r + n * c]); piv = r; }
    if (!(best > 0.0) || !(best < __builtin_huge_val())) return false;
    if (piv != c) for (int k = 0; k < n; ++k) {
      double t = a[c + n * k]; a[c + n * k] = a[piv + n * k]; a[piv + n * k] = t;
      t = inv[c + n * k]; inv[c + n * k] = inv[piv + n * k]; inv[piv + n * k] = t;
    }
    double d = a[c + n * c];
    for (int k = 0; k < n; ++k) { a[c + n * k] /= d; inv[c + n * k] /= d; }
    for (int r = 0; r < n; ++r) if (r != c) {
      double f = a[r + n * c];
      if (f != 0.0) for (int k = 0; k < n; ++k) { a[r + n * k] -= f * a[c + n * k]; inv[r + n * k] -= f * inv[c + n * k]; }
    }
  }
  return true;
}
// lower Cholesky factor L (L L' = m) and its inverse Linv (lower); false if not SPD
__host__ __device__ inline bool small_chol_inv(const double* m, int n, double* L, double* Linv) {
  for (int i = 0; i < n * n; ++i) { L[i] = 0.0; Linv[i] = 0.0; }
  for (int c = 0; c < n; ++c) {
    double s = m[c + n * c];
    for (int k = 0; k < c; ++k) s -= L[c + n * k] * L[c + n * k];
    if (!(s > 0.0) || !(s < __builtin_huge_val())) return false;
    double d = sqrt(s); L[c + n * c] = d;
    for (int r = c + 1; r < n; ++r) {
      double t = m[r + n * c];
      for (int k = 0; k < c; ++k) t -= L[r + n * k] * L[c + n * k];
      L[r + n * c] = t / d;
    }
  }
  // forward substitution for the inverse, column by column
  for (int c = 0; c < n; ++c) {
    Linv[c + n * c] = 1.0 / L[c + n * c];
    for (int r = c + 1; r < n; ++r) {
      double t = 0.0;
      for (int k = c; k < r; ++k) t -= L[r + n * k] * Linv[k + n * c];
      Linv[r + n * c] = t / L[r + n * r];
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------------
// qchisq(p, df), integer df: replaces R::qchisq (robust_cpp.cpp:102) / stats::qchisq
// (R/fit_sampleqc.R:429).  Half-integer-shape regularised incomplete gamma in closed form
// (lower tail by series, upper tail by the finite erfc / Poisson sums), safeguarded Newton.
// ---------------------------------------------------------------------------------------------
__device__ inline double gamma_p_series(double a, double y) {         // P(a, y), positive-term series
  if (y <= 0.0) return 0.0;
  double term = 1.0 / a, sum = term, ap = a;
  for (int n = 0; n < 4000; ++n) { ap += 1.0; term *= y / ap; sum += term; if (term < sum * 1e-17) break; }
  return sum * exp(-y + a * log(y) - lgamma(a));
}
__device__ inline double gamma_q_halfint(int df, double y) {          // Q(df/2, y), finite positive sum
  if (y <= 0.0) return 1.0;
  const double ey = exp(-y);
  if ((df & 1) == 0) {
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < df / 2; ++i) { term *= y / i; sum += term; }
    return ey * sum;
  }
  const double sy = sqrt(y);
  double q = erfc(sy);
  double term = ey * sy * 1.1283791670955125738961589031215;         // y^(1/2) e^-y / Gamma(3/2)
  for (int i = 0; i < df / 2; ++i) { q += term; term *= y / (1.5 + i); }
  return q;
}
__device__ inline double dev_qchisq(double p, int df) {
  if (!(p >= 0.0) || !(p <= 1.0) || df < 1) return CUDART_NAN;
  if (p == 0.0) return 0.0;
  if (p == 1.0) return CUDART_INF;
  const double a = 0.5 * df;
  const bool upper = p > 0.5;                      // solve on the smaller tail (1 - p is exact for p >= 0.5)
  const double target = upper ? 1.0 - p : p;
  const double lg = lgamma(a);
  // residual, increasing in y = x / 2
  auto resid = [&](double y) { return upper ? (target - gamma_q_halfint(df, y)) : (gamma_p_series(a, y) - target); };
  // bracket, then bisection-safeguarded Newton started from Wilson-Hilferty when that lies inside
  double lo = 0.0, hi = fmax(1.0, a);
  while (resid(hi) < 0.0) { lo = hi; hi *= 2.0; if (hi > 1e300) break; }
  const double zq = normcdfinv(p), c = 2.0 / (9.0 * df), w = 1.0 - c + zq * sqrt(c);
  double y = 0.5 * df * w * w * w;
  if (!(y > lo && y < hi)) y = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    const double f = resid(y);
    if (f > 0.0) hi = y; else lo = y;
    const double pdf = exp(-y + (a - 1.0) * log(y) - lg);
    double yn = (pdf > 0.0) ? y - f / pdf : 0.5 * (lo + hi);
    if (!(yn > lo && yn < hi)) yn = 0.5 * (lo + hi);
    if (fabs(yn - y) <= 4e-16 * fabs(yn)) { y = yn; break; }
    y = yn;
  }
  return 2.0 * y;
}

// ---------------------------------------------------------------------------------------------
// random keys for the FastMCD start (arma::randperm, robust_cpp.cpp:72)
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__host__ __device__ inline int counter_key(uint64_t call_mix, uint64_t cell) {
  uint64_t v = splitmix64(call_mix ^ (cell * 0x9E3779B97F4A7C15ull));
  return (int)(v >> 33);
}
__host__ __device__ inline uint64_t counter_call_mix(uint32_t seed, uint32_t call) {
  return splitmix64(((uint64_t)seed << 32) ^ (uint64_t)call);
}
// Mersenne-Twister output word -> randperm key:  (int) Rf_runif(0, RAND_MAX) after R's fixup()
__device__ __forceinline__ int mt_word_to_key(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  double u = (double)y * 2.3283064365386963e-10;
  const double i2_32m1 = 2.328306437080797e-10;
  if (u <= 0.0) u = 0.5 * i2_32m1;
  if ((1.0 - u) <= 0.0) u = 1.0 - 0.5 * i2_32m1;
  return (int)(2147483647.0 * u);
}

}  // namespace sqc
