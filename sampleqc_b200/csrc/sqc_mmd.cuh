// sqc_mmd.cuh — maximum mean discrepancy between samples: mmd_fn (reference src/mmd_fn.cpp:20-91)
// and the pair loop around it, .calc_mmd_mat / .dist_fn (reference R/calc_pairwise_mmds.R:198-280).
//
//   mmd(X, Y) = 2 / (n_x (n_x - 1)) sum_{i<j} k(x_i, x_j) + 2 / (n_y (n_y - 1)) sum_{i<j} k(y_i, y_j) - 2 / (n_x n_y) sum_{i,j} k(x_i, y_j),
//   k(a, b) = exp(-|a - b|^2 / sigma)          (the first two norms are 1 when a matrix has one row, :57-61, :73-77)
//
// A job = one (X, Y) pair of row sets: matrices are column-major with a leading dimension, rows optionally picked through an
// index list (the subsample .dist_fn draws with sample(), :265-277 — drawn by the caller, so R's RNG stream is the caller's).
// One thread owns one row i: the x-x terms j > i and all x-y terms of row i of X, the y-y terms j > i of row i of Y; the
// row-j loads are uniform over the CTA (broadcast, L1/L2 resident).  A job is cut into slices of MMD_THREADS rows; per-slice
// sums are reduced in fixed order, then over the slices in order: bit-reproducible.
#pragma once
#include "sqc_common.cuh"

namespace sqc {

constexpr int MMD_THREADS = 128;

struct MmdJob {
  const double* x; const double* y;      // column-major matrices
  const int* xi; const int* yi;          // row picks (or null: rows 0 .. n - 1)
  long long ldx, ldy;                    // leading dimensions
  int nx, ny;
  int slice0, nslices;                   // this job's partial sums live at part[3 * (slice0 + s)]
};

template <int D>
__global__ void __launch_bounds__(MMD_THREADS) k_mmd_slices(const MmdJob* jobs, const int* slice_job, double sigma, double* part) {
  __shared__ double sh[3 * (MMD_THREADS / 32)];
  const int sl = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const MmdJob jb = jobs[slice_job[sl]];
  const int s = sl - jb.slice0;
  const int i = s * MMD_THREADS + tid;
  const double inv = -1.0 / sigma;
  double xx = 0.0, yy = 0.0, xy = 0.0;
  auto row = [](const double* m, const int* pick, long long ld, int r, double (&v)[D]) {
    const long long rr = pick ? (long long)pick[r] : (long long)r;
#pragma unroll
    for (int d = 0; d < D; ++d) v[d] = __ldg(m + rr + ld * d);
  };
  auto kern = [&](const double (&a)[D], const double (&b)[D]) {
    double e = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) { const double t = a[d] - b[d]; e = fma(t, t, e); }
    return exp(e * inv);                                  // std::exp(-1 * euc / sigma), :52
  };
  if (i < jb.nx) {
    double a[D], b[D];
    row(jb.x, jb.xi, jb.ldx, i, a);
    for (int j = i + 1; j < jb.nx; ++j) { row(jb.x, jb.xi, jb.ldx, j, b); xx += kern(a, b); }     // :46-54
    for (int j = 0; j < jb.ny; ++j) { row(jb.y, jb.yi, jb.ldy, j, b); xy += kern(a, b); }         // :79-86
  }
  if (i < jb.ny) {
    double a[D], b[D];
    row(jb.y, jb.yi, jb.ldy, i, a);
    for (int j = i + 1; j < jb.ny; ++j) { row(jb.y, jb.yi, jb.ldy, j, b); yy += kern(a, b); }     // :62-70
  }
  xx = warp_sum(xx); yy = warp_sum(yy); xy = warp_sum(xy);
  if (lane == 0) { sh[3 * w] = xx; sh[3 * w + 1] = yy; sh[3 * w + 2] = xy; }
  __syncthreads();
  if (tid < 3) {
    double a = 0.0;
#pragma unroll
    for (int ww = 0; ww < MMD_THREADS / 32; ++ww) a += sh[3 * ww + tid];
    part[3 * (size_t)sl + tid] = a;
  }
}

// per job: the slices in order, the norms (:56-61, :72-77, :87), xx + yy - 2 xy (:90); absolute value as .dist_fn takes
// it (R/calc_pairwise_mmds.R:279) when `absolute`; then per group of n_times consecutive jobs their mean (:226-231)
__global__ void k_mmd_finish(const MmdJob* jobs, int n_jobs, const double* part, int absolute, int n_times, double* per_job, double* mean_out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_groups = n_times > 0 ? n_jobs / n_times : 0;
  if (n_times > 0 ? g >= n_groups : g >= n_jobs) return;
  const int j0 = n_times > 0 ? g * n_times : g, j1 = n_times > 0 ? j0 + n_times : g + 1;
  double acc = 0.0;
  for (int j = j0; j < j1; ++j) {
    const MmdJob jb = jobs[j];
    double xx = 0.0, yy = 0.0, xy = 0.0;
    for (int s = 0; s < jb.nslices; ++s) { xx += part[3 * (size_t)(jb.slice0 + s)]; yy += part[3 * (size_t)(jb.slice0 + s) + 1]; xy += part[3 * (size_t)(jb.slice0 + s) + 2]; }
    const double nxx = jb.nx == 1 ? 1.0 : 2.0 / jb.nx / (jb.nx - 1), nyy = jb.ny == 1 ? 1.0 : 2.0 / jb.ny / (jb.ny - 1);
    xx *= nxx; yy *= nyy; xy = xy / jb.ny / jb.nx;
    double m = xx + yy - 2 * xy;
    if (absolute) m = fabs(m);
    if (per_job) per_job[j] = m;
    acc += m;
  }
  if (mean_out && n_times > 0) mean_out[g] = acc / n_times;
}

}  // namespace sqc
