// sqc_lib.cu — host side of libsampleqc_cuda.so: the C ABI of include/sampleqc_cuda.h.
//
// Replaces the bodies of fit_sampleqc_robust_cpp (reference src/fit_sampleQC_robust_cpp.cpp:451-577)
// and fit_sampleqc_mle_cpp (src/fit_sampleQC_mle_cpp.cpp:279-357).  The host only validates, lays the
// data out in HBM, enqueues the whole EM (every kernel is predicated on the device-side loop flag, so
// nothing returns to the host until convergence) and copies the results back.  There is no CPU
// fallback: without a CUDA device every compute entry point returns SQC_ERR_CUDA.
#include <algorithm>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/sampleqc_cuda.h"
#include "sqc_mcd.cuh"
#include "sqc_mt.cuh"
#include "sqc_mle.cuh"
#include "sqc_xchg.cuh"
#include "sqc_init.cuh"
#include "sqc_mmd.cuh"

#ifndef SQC_MCD_RESERVE_SMS
#define SQC_MCD_RESERVE_SMS 24   // SMs k_mcd leaves to the key-stream kernels (R-compatible keys only), see session_create
#endif
#ifndef SQC_FOR_D
#define SQC_FOR_D(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#endif

using namespace sqc;

namespace {

struct SqcFail {
  int code; std::string msg;
};
[[noreturn]] void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  throw SqcFail{code, buf};
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      fail(e_ == cudaErrorMemoryAllocation ? SQC_ERR_NOMEM : SQC_ERR_CUDA, "CUDA: %s (%s) at %s:%d", \
           cudaGetErrorString(e_), #call, __FILE__, __LINE__);                                     \
  } while (0)

const char* device_status_message(int st) {
  switch (st) {
    case SQC_ERR_BAD_ARG: return "'mcd_alpha' must be between 0 and 1";
    case SQC_ERR_EMPTY_CLUSTER: return "empty cluster: x_k.rows(0, -1) out of bounds (robust_cpp.cpp:307)";
    case SQC_ERR_SUBSET_SIZE: return "randperm(): 'M' must be less than or equal to 'N' (h-subset size out of range, robust_cpp.cpp:72)";
    case SQC_ERR_NOT_SPD: return "chol(): decomposition failed (robust_cpp.cpp:19)";
    case SQC_ERR_SINGULAR: return "inv(): matrix is singular";
    case SQC_ERR_NAN: return "sort_index(): detected NaN (robust_cpp.cpp:50)";
    case SQC_ERR_SELECT: return "exact selection could not be resolved (too many exactly tied distances)";
    case SQC_D_PEER: return "another rank of the multi-GPU fit reported an error";
    case SQC_D_TIMEOUT: return "timed out waiting for another GPU of the fit (or for a worker CTA): a peer failed or never started";
    default: return "device error";
  }
}

// ---------------------------------------------------------------------------------------------
// small kernels that only the host driver needs
// ---------------------------------------------------------------------------------------------
__global__ void k_ctl_reset(Dev s) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    Control* c = s.ctl;
    c->iter = 0; c->z_delta = 0; c->cont = 1; c->status = 0; c->n_zero = 0; c->mcd_hit = 0; c->mcd_call = 0; c->pad = 0;
    c->mcd_passes = 0; c->mcd_csteps = 0; c->mcd_cells = 0; c->mcd_bytes = 0;
  }
}
__global__ void k_ctl_start_loop(Dev s) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { Control* c = s.ctl; c->cont = (s.em_iters > 0) && (c->status == 0); c->z_delta = 0; }
}
// after an M-step: block 0 derives the per-component constants (inverse Cholesky factor, inverse, normalisers —
// what RcppDist::dmvnorm recomputes on every call, SURVEY Appendix A4, hoisted); the other blocks update
// p_jk = (1 + n_jk) / (K + n_j) (update_p_jk :322-341) when with_p is set
template <int D>
__device__ void derive_component(Dev& s, int k, bool store, double* shPk, double* shlogP) {
  double S[D * D], Ai[D * D];
#pragma unroll
  for (int e = 0; e < D * D; ++e) S[e] = s.sigma[k * D * D + e];
  const bool ok1 = inv_small<D>(S, Ai);
  double L[SQC_MAXD * SQC_MAXD], Li[SQC_MAXD * SQC_MAXD];
  const bool ok2 = small_chol_inv(S, D, L, Li);
  const double det = small_det(S, D);
  shlogP[k] = -1.0 * (D / 2.0) * 1.837877066409345483560659472811 - 0.5 * log(det);
  shPk[k] = 1.0 / sqrt(pow(6.283185307179586476925286766559, (double)D) * det);
  if (!store) return;
  if (!ok1) { atomicCAS(&s.ctl->status, 0, SQC_D_SINGULAR); return; }
  if (!ok2) { atomicCAS(&s.ctl->status, 0, SQC_D_NOT_SPD); return; }
#pragma unroll
  for (int e = 0; e < D * D; ++e) { s.linv[k * D * D + e] = Li[e]; s.ainv[k * D * D + e] = Ai[e]; }
  s.logP[k] = shlogP[k]; s.Pk[k] = shPk[k];
  derive_dens_component<D>(s, k, Li);
}
// after an M-step.  Every block derives the component normalisers for itself (K tiny factorisations; block 0 also
// stores the per-component constants: inverse Cholesky factor, inverse, normalisers — what RcppDist::dmvnorm
// recomputes on every call, SURVEY Appendix A4, hoisted); then one warp per sample: p_jk = (1 + n_jk) / (K + n_j)
// (update_p_jk :322-341) when with_p is set, and the sample's weighted exp tables / log weights for the two passes.
template <int D>
__global__ void __launch_bounds__(256) k_after_mstep(Dev s, int with_p) {
  __shared__ double shPk[SQC_MAXK], shlogP[SQC_MAXK];
  if (!s.ctl->cont) return;
  const int K = s.K, lane = threadIdx.x & 31;
  if ((int)threadIdx.x < K) derive_component<D>(s, threadIdx.x, blockIdx.x == 0, shPk, shlogP);
  __syncthreads();
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= s.J) return;
  if (with_p) {
    long long nj = K;
    for (int k = 0; k < K; ++k) nj += s.n_jk[j * K + k];
    if (lane < K) {
      const double p = (1.0 + (double)s.n_jk[j * K + lane]) / (double)nj;
      s.p_jk[j * K + lane] = p; s.logp_jk[j * K + lane] = log(p);
    }
    __syncwarp();
  }
  derive_dens_sample<D>(s, j, lane, shPk, shlogP);
}
// final relabel by ascending beta_k[,0] (robust_cpp.cpp:527-531, reorder_z :434-445) and packing of
// every output into the column-major shapes of the Rcpp list
struct OutPtrs { double* mu0; double* alpha; double* beta; double* sigma; double* p_jk; int* z; int* korder; };
__global__ void k_order_components(Dev s, OutPtrs o) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int idx[SQC_MAXK];
    for (int k = 0; k < s.K; ++k) idx[k] = k;
    for (int a = 1; a < s.K; ++a) {             // stable insertion sort (ties: lower index first)
      int v = idx[a], b = a - 1;
      while (b >= 0 && s.beta[idx[b] * s.D] > s.beta[v * s.D]) { idx[b + 1] = idx[b]; --b; }
      idx[b + 1] = v;
    }
    for (int k = 0; k < s.K; ++k) o.korder[k] = idx[k];
    for (int k = 0; k < s.K; ++k) o.korder[SQC_MAXK + idx[k]] = k;     // inverse map
  }
}
// zmode: 0 = inverse relabel (reorder_z), 1 = loop never ran (labels all zero), 2 = idx(k_max) (mle extract_z :241)
__global__ void k_pack_outputs(Dev s, OutPtrs o, int zmode) {
  const int D = s.D, J = s.J, K = s.K;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  for (long long e = gid; e < (long long)J * D; e += gsz) { int j = (int)(e / D), d = (int)(e % D); o.alpha[j + (long long)J * d] = s.alpha[e]; }
  for (long long e = gid; e < (long long)K * D; e += gsz) { int k = (int)(e / D), d = (int)(e % D); o.beta[k + K * d] = s.beta[o.korder[k] * D + d]; }
  for (long long e = gid; e < (long long)K * D * D; e += gsz) { int k = (int)(e / (D * D)), r = (int)(e % (D * D)); o.sigma[e] = s.sigma[o.korder[k] * D * D + r]; }
  for (long long e = gid; e < (long long)J * K; e += gsz) { int j = (int)(e / K), k = (int)(e % K); o.p_jk[j + (long long)J * k] = s.p_jk[j * K + o.korder[k]]; }
  for (long long e = gid; e < D; e += gsz) o.mu0[e] = s.mu0[e];
  if (o.z) for (long long i = gid; i < s.N; i += gsz)
    o.z[i] = (zmode == 1) ? (o.korder[SQC_MAXK + 0] + 1) : (zmode == 2) ? (o.korder[s.z[i]] + 1) : (o.korder[SQC_MAXK + s.z[i]] + 1);
}
__global__ void k_widen_labels(const int* in, uint8_t* out, long long n, int K, int* bad) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  for (long long i = gid; i < n; i += gsz) { int v = in[i]; if (v < 0 || v >= K) { *bad = 1; v = 0; } out[i] = (uint8_t)v; }
}
__global__ void k_expneg(const double* x, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = exp_neg(x[i]);
}
__global__ void __launch_bounds__(1024) k_fp64_peak(double* out, int iters, double m) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, 1e-9); a1 = fma(a1, m, 1e-9); a2 = fma(a2, m, 1e-9); a3 = fma(a3, m, 1e-9);
    a4 = fma(a4, m, 1e-9); a5 = fma(a5, m, 1e-9); a6 = fma(a6, m, 1e-9); a7 = fma(a7, m, 1e-9);
  }
  const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (r == 12345.678) *out = r;
}
__global__ void k_qchisq(const double* p, const double* df, double* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dev_qchisq(p[i], (int)df[i]);
}
// .calc_outliers_dt (reference R/fit_sampleqc.R:419-453): N x K squared Mahalanobis distances with
// stats::mahalanobis' solve(cov) form, row minimum against qchisq(1 - alpha, D)
template <int D>
__global__ void __launch_bounds__(256) k_outlier_maha(const double* x, const int* groups, long long N, int J, int K,
                                                      const double* mu0, const double* alpha /* J x D col-major */,
                                                      const double* beta /* K x D col-major */, const double* ainv /* K x D x D */,
                                                      double cut, double* maha, uint8_t* outlier) {
  extern __shared__ double sh[];                 // K * (D + D*D)
  for (int e = threadIdx.x; e < K * D; e += blockDim.x) { int k = e / D, d = e % D; sh[k * (D + D * D) + d] = beta[k + K * d]; }
  for (int e = threadIdx.x; e < K * D * D; e += blockDim.x) { int k = e / (D * D), r = e % (D * D); sh[k * (D + D * D) + D + r] = ainv[e]; }
  __syncthreads();
  const long long gsz = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gsz) {
    const int j = ld_stream_s32(groups + i);
    double xc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) xc[d] = (ld_stream(x + (long long)d * N + i) - mu0[d]) - alpha[j + (long long)J * d];   // :432
    double mn = CUDART_INF;
    for (int k = 0; k < K; ++k) {
      const double* pk = sh + k * (D + D * D);
      double v[D];
#pragma unroll
      for (int d = 0; d < D; ++d) v[d] = xc[d] - pk[d];
      double q = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {                // rowSums((x %*% cov) * x), stats::mahalanobis
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < D; ++r) t += v[r] * pk[D + r + D * c];
        q += t * v[c];
      }
      if (maha) maha[i + N * k] = q;
      if (q < mn) mn = q;
    }
    if (outlier) outlier[i] = (mn > cut) ? 1 : 0;   // :449
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// device memory: sessions carve their buffers out of a few large chunks; finished sessions hand the chunks to a
// per-device cache instead of cudaFree (cudaMalloc / cudaFree of GBs cost tens of ms and synchronise the device),
// so repeated fits — one per sample group in fit_sampleqc(K_list = ...) — pay for the allocation once.
// sqc_trim() releases the cache.
// ---------------------------------------------------------------------------------------------
namespace {
struct Chunk { char* base; size_t bytes; };
struct ChunkCache {
  std::mutex mu; std::vector<Chunk> free_chunks[64];
  Chunk take(int device, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu);
    auto& v = free_chunks[device & 63];
    int best = -1;
    for (int i = 0; i < (int)v.size(); ++i) if (v[i].bytes >= bytes && (best < 0 || v[i].bytes < v[best].bytes)) best = i;
    if (best >= 0) { Chunk c = v[best]; v.erase(v.begin() + best); return c; }
    // nothing fits: drop the cached chunks (they would only fragment the device) and allocate
    for (auto& c : v) cudaFree(c.base);
    v.clear();
    Chunk c{nullptr, bytes};
    cudaError_t e = cudaMalloc((void**)&c.base, bytes);
    if (e != cudaSuccess) fail(e == cudaErrorMemoryAllocation ? SQC_ERR_NOMEM : SQC_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return c;
  }
  void give(int device, Chunk c) { std::lock_guard<std::mutex> lk(mu); free_chunks[device & 63].push_back(c); }
  void trim(int device) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& c : free_chunks[device & 63]) cudaFree(c.base);
    free_chunks[device & 63].clear();
  }
};
ChunkCache g_chunks;
}  // namespace

// ---------------------------------------------------------------------------------------------
// one sharded fit inside ONE process (sqc_fit_args.n_gpus > 1): a host thread and a device per rank.  The ranks meet at
// a few host-side sync points while their sessions are built (windows allocated and cleared, key buffers allocated);
// a rank that fails says so there, and every rank gives up before an exchange kernel could wait for it.
// ---------------------------------------------------------------------------------------------
struct InprocGroup {
  int world = 1; int device[XCHG_MAX_WORLD] = {};
  std::mutex mu; std::condition_variable cv; int arrived = 0; unsigned long long gen = 0; bool failed = false;
  bool sync(bool ok) {                                  // true iff no rank has failed so far
    std::unique_lock<std::mutex> lk(mu);
    if (!ok) failed = true;
    if (failed) { cv.notify_all(); return false; }
    const unsigned long long g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g || failed; });
    return !failed;
  }
  void abort() { std::lock_guard<std::mutex> lk(mu); failed = true; cv.notify_all(); }
};

// ---------------------------------------------------------------------------------------------
// session
// ---------------------------------------------------------------------------------------------
struct sqc_session {
  int D = 0, J = 0, K = 0, variant = 0, device = 0, tile = 0;
  long long N = 0, N_total = 0;
  int em_iters = 0, mcd_iters = 0, rng_mode = 0, track = 0, rank = 0, world = 1;
  double mcd_alpha = 0.5; unsigned seed = 0;
  cudaStream_t st = nullptr, st_rng = nullptr, st_rng2 = nullptr;
  cudaEvent_t ev_key_ready[2] = {nullptr, nullptr}, ev_key_free[2] = {nullptr, nullptr}, ev_pre_mcd = nullptr, ev_post_mcd = nullptr, ev_chain[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  Dev dev{};
  McdBufs mb{};
  MleBufs ml{};
  Xchg xc{};
  InprocGroup* grp = nullptr;            // non-null: the ranks are threads of this process
  uint8_t* z0 = nullptr; double* alpha0 = nullptr; double* gamma0 = nullptr;
  int mt_target_blocks = 49;
  bool rng_after_mcd = false;          // generate the next call's keys behind (not beside) the M-step kernel
  unsigned long long* xg_gather = nullptr; double* xg_vec = nullptr;
  uint32_t* mt_state = nullptr; uint32_t* mt_seq = nullptr; uint32_t* mt_table = nullptr; int* keybuf = nullptr;
  // distributed key stream (multi-GPU, R-compatible keys): this rank generates draws [rank * key_q, (rank + 1) * key_q) of every call
  long long key_q = 0, key_mine = 0; uint32_t* mt_rows = nullptr; uint32_t* mt_jwin = nullptr; uint32_t* mt_seq2 = nullptr; const int* key_peer[XCHG_MAX_WORLD] = {}; void* key_mapped[XCHG_MAX_WORLD] = {};
  double* o_buf = nullptr; int* o_z = nullptr; int* o_korder = nullptr; OutPtrs optr{};
  size_t o_doubles = 0;
  int mcd_grid = 0, n_sm = 0;
  long long ld = 0;                      // leading dimension of the caller's x / init_gamma
  bool sorted = true; std::vector<long long> perm;      // perm[i] = original index of the cell stored at i (cells of a sample were interleaved)
  std::vector<int> smap;                                // non-empty: internal sample r is the caller's sample smap[r] (contiguous runs in non-ascending code order)
  bool profile = false;
  struct Ev { cudaEvent_t a, b; int fam; const char* tag = nullptr; };
  std::vector<Ev> evs; size_t ev_used = 0;
  cudaEvent_t ev_run0 = nullptr, ev_init = nullptr, ev_loop = nullptr;
  sqc_run_stats stats{};
  bool ran = false;
  Control h_ctl{};

  std::vector<Chunk> chunks; size_t chunk_used = 0, first_chunk_bytes = 64ull << 20;
  template <class T> T* alloc(size_t n, bool zero = true) {
    const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255;
    static const bool no_arena = getenv("SQC_NO_ARENA") != nullptr;     // debug: one allocation per buffer
    if (no_arena || chunks.empty() || chunk_used + bytes > chunks.back().bytes) {
      chunks.push_back(g_chunks.take(device, no_arena ? bytes : std::max(bytes, chunks.empty() ? first_chunk_bytes : (size_t)(64ull << 20))));
      chunk_used = 0;
    }
    void* p = chunks.back().base + chunk_used;
    chunk_used += bytes;
    if (zero) CK(cudaMemsetAsync(p, 0, bytes, st));
    return (T*)p;
  }
  ~sqc_session() {
    if (device >= 0) cudaSetDevice(device);
    if (st) cudaStreamSynchronize(st);
    if (st_rng) cudaStreamSynchronize(st_rng);
    if (st_rng2) cudaStreamSynchronize(st_rng2);
    xchg_close(xc);
    for (auto& c : chunks) g_chunks.give(device, c);
    for (auto& e : evs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (ev_run0) cudaEventDestroy(ev_run0);
    if (ev_init) cudaEventDestroy(ev_init);
    if (ev_loop) cudaEventDestroy(ev_loop);
    for (int i = 0; i < 2; ++i) { if (ev_key_ready[i]) cudaEventDestroy(ev_key_ready[i]); if (ev_key_free[i]) cudaEventDestroy(ev_key_free[i]); }
    if (ev_pre_mcd) cudaEventDestroy(ev_pre_mcd);
    for (int i = 0; i < 2; ++i) if (ev_chain[i]) cudaEventDestroy(ev_chain[i]);
    if (ev_post_mcd) cudaEventDestroy(ev_post_mcd);
    if (st_rng2 && st_rng2 != st) cudaStreamDestroy(st_rng2);
    if (st_rng && st_rng != st) cudaStreamDestroy(st_rng);
    if (st) cudaStreamDestroy(st);
  }
};

namespace {
void mle_alloc(sqc_session&, const sqc_fit_args*, sqc::MleBufs&);
void mle_run(sqc_session*);
void mle_fetch(sqc_session*, sqc_fit_out*);
void xchg_allreduce_host(sqc_session* S, double* buf, int n);
void key_stream_setup(sqc_session* S);

// per-launch accounting: count every launch, time it with events when profiling is on
struct Launch {
  sqc_session* S; int fam; sqc_session::Ev* ev = nullptr; cudaStream_t on;
  Launch(sqc_session* s, int f, cudaStream_t stream_ = nullptr, const char* tag = nullptr) : S(s), fam(f), on(stream_) {
    S->stats.n_gpu_launches++; S->stats.family_launches[fam]++;
    if (S->profile) {
      if (S->ev_used == S->evs.size()) { sqc_session::Ev e; CK(cudaEventCreate(&e.a)); CK(cudaEventCreate(&e.b)); e.fam = f; S->evs.push_back(e); }
      ev = &S->evs[S->ev_used++]; ev->fam = f; ev->tag = tag;
      CK(cudaEventRecord(ev->a, stream()));
    }
  }
  cudaStream_t stream() const { return on ? on : (fam == SQC_KF_RNG ? S->st_rng : S->st); }
  void done() { if (ev) CK(cudaEventRecord(ev->b, stream())); CK(cudaGetLastError()); }
};

template <int D> void launch_estep(sqc_session* S, int mode) {
  const Dev& s = S->dev;
  if (mode == 0) {
    const size_t sm = (size_t)(dens_doubles<D>(s.K) + (EST_THREADS / 32) * s.K * (D + 1)) * sizeof(double);
    Launch L(S, SQC_KF_ESTEP);
    k_estep<D><<<s.T, EST_THREADS, sm, S->st>>>(s);
    L.done();
  } else {                                                   // statistics of the initial labelling (one full pass per fit)
    const size_t sm = (size_t)((EST_THREADS / 32) * s.K * (D + 1)) * sizeof(double);
    Launch L(S, SQC_KF_INIT);
    k_tile_stats<D><<<s.T, EST_THREADS, sm, S->st>>>(s);
    L.done();
  }
}
template <int D> void launch_partition(sqc_session* S, int like) {
  const Dev& s = S->dev;
  const size_t sm = (size_t)dens_doubles<D>(s.K) * sizeof(double) + (size_t)TileCfg<D>::CPT * (EST_THREADS / 32) * s.K * sizeof(int);
  Launch L(S, SQC_KF_PARTITION);
  if (like) k_partition<D, 1><<<s.T, EST_THREADS, sm, S->st>>>(s);
  else k_partition<D, 0><<<s.T, EST_THREADS, sm, S->st>>>(s);
  L.done();
}
template <int D> void launch_mcd(sqc_session* S) {
  Launch L(S, SQC_KF_MCD);
  void* args[] = {(void*)&S->dev, (void*)&S->mb};
  void* fn = S->world > 1 ? (void*)k_mcd<D, true> : (void*)k_mcd<D, false>;
  // cooperative launch = guaranteed co-residency of the leader / worker CTAs that wait on each other's flags.  (A plain
  // launch was tried so that the key-stream kernels could run beside it; they still did not — the two kernels want
  // different shared-memory carve-outs — so the guarantee is kept.)
  CK(cudaLaunchCooperativeKernel(fn, dim3(S->mcd_grid), dim3(MCD_THREADS), args, MCD_SMEM, S->st));
  L.done();
}
template <int D> int mcd_max_grid(int n_sm) {
  CK(cudaFuncSetAttribute(k_mcd<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCD_SMEM));
  CK(cudaFuncSetAttribute(k_mcd<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MCD_SMEM));
  int per = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_mcd<D, true>, MCD_THREADS, MCD_SMEM));
  if (per < 1) fail(SQC_ERR_CUDA, "k_mcd<%d> does not fit on an SM", D);
  return std::min(n_sm, (MCD_NB - MCD_NF - 16) / MCD_WARPS);   // one CTA per SM (the prefix of grid x warps candidate regions sits behind the fine histogram)
}
template <int D> void launch_init_passes(sqc_session* S) {
  const Dev& s = S->dev;
  { Launch L(S, SQC_KF_INIT); k_colsum<D><<<s.T, EST_THREADS, 0, S->st>>>(s); L.done(); }
  { Launch L(S, SQC_KF_INIT); k_colmean<<<1, 1024, 0, S->st>>>(s); L.done(); }
  if (S->world > 1) xchg_allreduce_host(S, s.mu0, s.D);
  { Launch L(S, SQC_KF_INIT); k_colmean_finish<<<1, 32, 0, S->st>>>(s); L.done(); }
  { Launch L(S, SQC_KF_INIT); k_centre<D><<<s.T, EST_THREADS, 0, S->st>>>(s); L.done(); }
  { Launch L(S, SQC_KF_INIT); k_init_alpha<<<(s.J * 32 + 255) / 256, 256, 0, S->st>>>(s); L.done(); }
}

// ---- host-stream level exchanges (multi-GPU) -----------------------------------------------------
// after k_post_stats: global cluster sizes and the key-stream offset of the local segment of every cluster
// (the stream of one call runs over clusters in order, inside a cluster in cell order,
// and lower ranks hold the lower cell indices); the tile partial sums of the log-likelihoods become global sums
// (slot 0 of t_ll1 / t_ll2), z_delta and the status are agreed on.
// The per-iteration host-stream exchanges in ONE kernel (one CTA): cluster sizes of every rank -> global sizes and
// key-stream offsets (k_layout_finish), and with_like: (like_1, like_2, z_delta, n_zero, status) summed in rank order,
// then the end-of-iteration control (k_like_global + k_em_control).
__global__ void __launch_bounds__(1024) k_iter_exchange(Dev s, XchgDev x, int with_like) {
  __shared__ double sh[32];
  __shared__ unsigned long long sh_n[XCHG_MAX_WORLD * SQC_MAXK];
  __shared__ double sh_d[XCHG_MAX_WORLD * 5];
  __shared__ double sh_g[5];
  if (!s.ctl->cont) return;
  const int tid = threadIdx.x, K = s.K, W = x.world, ch = XCHG_CH_HOST + 1;
  double r1 = 0.0, r2 = 0.0;
  if (with_like) {
    double a1 = 0.0, a2 = 0.0;
    for (int t = tid; t < s.T; t += 1024) { a1 += s.t_ll1[t]; a2 += s.t_ll2[t]; }
    r1 = block_sum<1024>(a1, sh);
    r2 = block_sum<1024>(a2, sh);
  }
  const unsigned long long seq = xchg_begin(x, ch);
  if (tid < K) { const unsigned long long v = (unsigned long long)s.clu_n[tid]; xchg_put(x, ch, seq, tid, (unsigned int)v, (unsigned int)(v >> 32)); }
  if (tid == 0 && with_like) {                     // (block_sum leaves the total in thread 0)
    xchg_put_f64(x, ch, seq, K + 0, r1); xchg_put_f64(x, ch, seq, K + 1, r2); xchg_put_f64(x, ch, seq, K + 2, (double)s.ctl->z_delta);
    xchg_put_f64(x, ch, seq, K + 3, (double)s.ctl->n_zero); xchg_put_f64(x, ch, seq, K + 4, (double)s.ctl->status);
  }
  if (tid < K * W) {
    const int r = tid / K, k = tid % K;
    const ulonglong2 v = xchg_wait(x, ch, seq, r, k);
    sh_n[r * K + k] = (unsigned long long)(unsigned int)v.x | ((unsigned long long)(unsigned int)v.y << 32);
  } else if (with_like && tid >= 512 && tid < 512 + 5 * W) {
    const int r = (tid - 512) / 5, e = (tid - 512) % 5;
    sh_d[r * 5 + e] = line_f64(xchg_wait(x, ch, seq, r, K + e));
  }
  __syncthreads();
  if (tid == 0) {
    long long run = 0;
    for (int k = 0; k < K; ++k) {
      long long tot = 0, before = 0;
      for (int r = 0; r < W; ++r) { const long long v = (long long)sh_n[r * K + k]; tot += v; if (r < x.rank) before += v; }
      s.clu_n_glob[k] = tot; s.clu_prefix[k] = run + before; run += tot;
    }
  }
  if (!with_like) return;
  if (tid < 5) { double g = 0.0; for (int r = 0; r < W; ++r) g += sh_d[r * 5 + tid]; sh_g[tid] = g; }
  __syncthreads();
  if (tid == 0) {
    s.ctl->z_delta = (int)sh_g[2];
    if (sh_g[4] != 0.0 && s.ctl->status == 0) s.ctl->status = SQC_D_PEER;            // a peer failed (its own message says why)
  }
  for (int t = tid; t < s.T; t += 1024) { s.t_ll1[t] = (t == 0) ? sh_g[0] : 0.0; s.t_ll2[t] = (t == 0) ? sh_g[1] : 0.0; }
  __syncthreads();
  em_control_block(s);
}
void xchg_iteration(sqc_session* S, Dev& s, int with_like) {
  Launch L(S, SQC_KF_UPDATE, nullptr, "iter_exchange"); k_iter_exchange<<<1, 1024, 0, S->st>>>(s, S->xc.dev, with_like); L.done();
}

void xchg_allreduce_host(sqc_session* S, double* buf, int n) {
  Launch L(S, SQC_KF_UPDATE); k_xchg_sum_f64<<<1, 256, 0, S->st>>>(S->xc.dev, XCHG_CH_HOST + 0, buf, n, nullptr); L.done();
}

#define SQC_CASE_ESTEP(DD) case DD: launch_estep<DD>(S, mode); break;
void do_estep(sqc_session* S, int mode) { switch (S->D) { SQC_FOR_D(SQC_CASE_ESTEP) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); } }
#define SQC_CASE_PART(DD) case DD: launch_partition<DD>(S, like); break;
void do_partition(sqc_session* S, int like) { switch (S->D) { SQC_FOR_D(SQC_CASE_PART) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); } }
#define SQC_CASE_MCD(DD) case DD: launch_mcd<DD>(S); break;
void do_mcd(sqc_session* S) { switch (S->D) { SQC_FOR_D(SQC_CASE_MCD) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); } }
#define SQC_CASE_GRID(DD) case DD: return mcd_max_grid<DD>(n_sm);
int do_mcd_grid(int D, int n_sm) { switch (D) { SQC_FOR_D(SQC_CASE_GRID) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", D); } }
#define SQC_CASE_INIT(DD) case DD: launch_init_passes<DD>(S); break;
void do_init_passes(sqc_session* S) { switch (S->D) { SQC_FOR_D(SQC_CASE_INIT) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); } }
#define SQC_CASE_ALPHA(DD) case DD: k_update_alpha<DD><<<(S->J + 127) / 128, 128, 0, S->st>>>(S->dev); break;
void do_update_alpha(sqc_session* S) {
  Launch L(S, SQC_KF_UPDATE);
  switch (S->D) { SQC_FOR_D(SQC_CASE_ALPHA) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); }
  L.done();
}
#define SQC_CASE_POST(DD) case DD: k_post_stats<DD><<<nsb + (flags & 1), 1024, 0, S->st>>>(S->dev, nsb, flags); break;
void do_post_stats(sqc_session* S, int flags) {
  const int nsb = (S->J + 31) / 32;
  Launch L(S, SQC_KF_UPDATE, nullptr, "post_stats");
  switch (S->D) { SQC_FOR_D(SQC_CASE_POST) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); }
  L.done();
}
#define SQC_CASE_AFTER(DD) case DD: k_after_mstep<DD><<<(S->J + 7) / 8, 256, 0, S->st>>>(S->dev, with_p); break;
void do_after_mstep(sqc_session* S, int with_p) {
  Launch L(S, SQC_KF_UPDATE);
  switch (S->D) { SQC_FOR_D(SQC_CASE_AFTER) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); }
  L.done();
}
#define SQC_CASE_MLE(DD) case DD: mle_launch_pass<DD>(S->dev, S->ml, which, S->st); break;
void do_mle_pass(sqc_session* S, int which) {
  static const char* tags[] = {"pass_init", "pass_a", "pass_b", "pass_like"};
  Launch L(S, SQC_KF_MLE_PASS, nullptr, tags[which & 3]);
  switch (S->D) { SQC_FOR_D(SQC_CASE_MLE) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", S->D); }
  L.done();
}
int tile_cells(int D) { return D <= 4 ? TileCfg<3>::TILE : TileCfg<6>::TILE; }

void select_device(int device, int* chosen) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n < 1) fail(SQC_ERR_CUDA, "no usable CUDA device (%s): this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  int d = device;
  if (d < 0) CK(cudaGetDevice(&d));
  if (d >= n) fail(SQC_ERR_CUDA, "device %d out of range (%d visible)", d, n);
  CK(cudaSetDevice(d));
  *chosen = d;
}

void check_args(const sqc_fit_args* a, int variant) {
  if (!a) fail(SQC_ERR_BAD_ARG, "null arguments");
  if (a->abi_version != SQC_ABI_VERSION) fail(SQC_ERR_BAD_ARG, "ABI version mismatch: caller %u, library %d", a->abi_version, SQC_ABI_VERSION);
  if (!a->x || !a->groups) fail(SQC_ERR_BAD_ARG, "null input");
  if (a->D < 1 || a->D > SQC_MAXD) fail(SQC_ERR_BAD_ARG, "D must be in 1..%d", SQC_MAXD);
  if (a->K < 1 || a->K > SQC_MAXK) fail(SQC_ERR_BAD_ARG, "K must be in 1..%d", SQC_MAXK);
  if (a->J < 1 || a->N < 1) fail(SQC_ERR_BAD_ARG, "bad dimensions");
  if (a->N > 2000000000ll) fail(SQC_ERR_BAD_ARG, "N too large for one shard");
  if (variant == SQC_VARIANT_ROBUST) {
    if (!a->init_z) fail(SQC_ERR_BAD_ARG, "init_z missing");
    if ((a->mcd_alpha < 0) | (a->mcd_alpha >= 1) | !(a->mcd_alpha == a->mcd_alpha))
      fail(SQC_ERR_BAD_ARG, "'mcd_alpha' must be between 0 and 1");                    // robust_cpp.cpp:471-472
  } else if (!a->init_gamma) fail(SQC_ERR_BAD_ARG, "init_gamma missing");
  if (a->rng_mode != SQC_RNG_R_COMPAT && a->rng_mode != SQC_RNG_COUNTER) fail(SQC_ERR_BAD_ARG, "unknown rng_mode");
}

double now_ms();
struct PhaseTimer {
  bool on; double t0; const char* what[16]; double dt[16]; int n = 0;
  PhaseTimer() : on(getenv("SQC_TIMING") != nullptr), t0(0) { if (on) t0 = now_ms(); }
  void mark(const char* w) { if (!on || n >= 16) return; const double t = now_ms(); what[n] = w; dt[n++] = t - t0; t0 = t; }
  void dump(const char* head) { if (!on) return; fprintf(stderr, "[sqc timing] %s:", head); for (int i = 0; i < n; ++i) fprintf(stderr, " %s %.2f", what[i], dt[i]); fprintf(stderr, "\n"); }
};
sqc_session* session_create(const sqc_fit_args* a, int variant, InprocGroup* grp = nullptr, long long ld = 0) {
  PhaseTimer pt;
  check_args(a, variant);
  if (ld <= 0) ld = a->N;                               // leading dimension of the host matrices x / init_gamma (in-process shards are row ranges)
  std::unique_ptr<sqc_session> up(new sqc_session());
  sqc_session* S = up.get();
  select_device(a->device, &S->device);
  CK(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
  CK(cudaDeviceGetAttribute(&S->n_sm, cudaDevAttrMultiProcessorCount, S->device));
  S->D = a->D; S->J = a->J; S->K = a->K; S->N = a->N; S->variant = variant; S->grp = grp; S->ld = ld;
  S->world = a->world > 1 ? a->world : 1; S->rank = S->world > 1 ? a->rank : 0;
  S->N_total = S->world > 1 ? a->N_total : a->N;
  S->em_iters = a->em_iters; S->mcd_iters = a->mcd_iters; S->mcd_alpha = a->mcd_alpha; S->seed = a->seed;
  S->rng_mode = a->rng_mode; S->track = a->track ? 1 : 0;
  S->profile = getenv("SQC_PROFILE") && atoi(getenv("SQC_PROFILE")) != 0;
  const int D = S->D, J = S->J, K = S->K; const long long N = S->N;
  const int em = std::max(S->em_iters, 0);

  // one chunk for (almost) everything: generous estimate of what the session will carve out
  {
    const double n = (double)N, nt = (double)S->N_total;
    double est = n * (8.0 * D * 2 + 16) + nt * 8.0 + 64.0 * (K + 2) * 8 * D;
    est += (n / tile_cells(D) + J + 16) * (K * (16.0 + 8 * D) + 64 + 8.0 * D + (variant == SQC_VARIANT_MLE ? 8.0 * K * SQC_NMOM(D) : 0.0));
    est += (double)K * 256 * MCD_WARPS * MCD_CAPW * (12.0 + 8 * D) + (double)J * (K * (24.0 + 8 * D) + 16.0 * D) + (32 << 20);
    if (variant == SQC_VARIANT_MLE) est += n * 16.0 * K;
    if (S->track) est += (double)em * (J * (D + K) + K * D * (D + 1)) * 8.0;
    S->first_chunk_bytes = (size_t)est;
  }
  pt.mark("setup");
  // ---- upload x first (asynchronous from pinned memory), scan the sample layout on the host meanwhile ----
  Dev& s = S->dev;
  s.x = S->alloc<double>((size_t)N * D, false);
  CK(cudaMemcpy2DAsync(s.x, (size_t)N * sizeof(double), a->x, (size_t)ld * sizeof(double), (size_t)N * sizeof(double), D, cudaMemcpyDefault, S->st));   // (a device pointer in the resident pipeline)
  // Cells of a sample must be contiguous in HBM (tiles never cross a sample boundary).  Three cases:
  //   groups non-decreasing                     -> as is;
  //   every sample one contiguous run, any order -> as is, samples relabelled by order of appearance (rows of alpha_j /
  //      p_jk are put back in code order on fetch).  This is what fit_sampleqc() sends: x is an rbind over samples, but
  //      groups = as.integer(factor(sample_ids)) numbers them alphabetically (R/fit_sampleqc.R:205-216).  Cell order is
  //      untouched, so the randperm keys pair with the same cells as in the reference (restrict_to_x_k keeps cell order);
  //   cells of a sample interleaved             -> stable counting sort by sample, once.  The fit is valid, but the
  //      R-compatible key stream is then consumed in the sorted cell order, not the caller's.
  std::vector<long long> n_j(J, 0);
  bool sorted = true, contiguous = true;
  std::vector<int> run_code;                          // sample codes in order of appearance (while contiguous)
  {
    const int nth = (N >= (1 << 20)) ? 4 : 1;           // a few host threads: the scan hides behind the upload of x
    std::vector<std::vector<long long>> cnt(nth, std::vector<long long>(J, 0));
    std::vector<std::vector<int>> runs(nth);
    std::vector<long long> bad(nth, -1); std::vector<char> srt(nth, 1), ovf(nth, 0);
    auto scan = [&](int t) {
      const long long i0 = N * t / nth, i1 = N * (t + 1) / nth;
      long long* c = cnt[t].data();
      int prev = i0 ? a->groups[i0 - 1] : -1;
      for (long long i = i0; i < i1; ++i) {
        const int g = a->groups[i];
        if ((unsigned)g >= (unsigned)J) { bad[t] = i; return; }
        c[g]++;
        if (g != prev) {
          if (g < prev) srt[t] = 0;
          if ((long long)runs[t].size() <= J) runs[t].push_back(g); else ovf[t] = 1;
        }
        prev = g;
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nth; ++t) th.emplace_back(scan, t);
    scan(0);
    for (auto& x : th) x.join();
    for (int t = 0; t < nth; ++t) {
      if (bad[t] >= 0) { cudaStreamSynchronize(S->st); fail(SQC_ERR_BAD_ARG, "groups out of range at cell %lld", bad[t]); }
      if (!srt[t]) sorted = false;
      if (ovf[t]) contiguous = false;
      for (int j = 0; j < J; ++j) n_j[j] += cnt[t][j];
      for (int g : runs[t]) if ((long long)run_code.size() <= J) run_code.push_back(g); else contiguous = false;
    }
    if (!sorted && contiguous) {                        // one run per sample?
      std::vector<char> seen(J, 0);
      for (int g : run_code) { if (seen[g]) { contiguous = false; break; } seen[g] = 1; }
    }
  }
  pt.mark("x_enqueue+scan");
  S->sorted = sorted || contiguous;                     // cells stay where they are
  S->smap.clear();
  if (!sorted && contiguous) {                          // relabel the samples by appearance; empty samples go last
    std::vector<char> seen(J, 0);
    for (int g : run_code) { S->smap.push_back(g); seen[g] = 1; }
    for (int j = 0; j < J; ++j) if (!seen[j]) S->smap.push_back(j);
    std::vector<long long> n_int(J);
    for (int r = 0; r < J; ++r) n_int[r] = n_j[S->smap[r]];
    n_j.swap(n_int);
  } else if (!sorted) {                                 // stable counting sort by sample
    if (S->rng_mode == SQC_RNG_R_COMPAT && variant == SQC_VARIANT_ROBUST) {
      static bool warned = false;
      if (!warned) { warned = true; fprintf(stderr, "sampleqc_cuda: cells of a sample are not contiguous; they are sorted by sample once, and the R-compatible random-start keys follow that sorted cell order\n"); }
    }
    std::vector<long long> start(J + 1, 0);
    for (int j = 0; j < J; ++j) start[j + 1] = start[j] + n_j[j];
    S->perm.resize(N);
    for (long long i = 0; i < N; ++i) S->perm[start[a->groups[i]]++] = i;
  }
  sorted = S->sorted;
  const int TILE = tile_cells(D);
  S->tile = TILE;
  std::vector<int> t_start, t_len, t_sample, s_tb(J + 1, 0), s_n(J);
  {
    long long pos = 0;
    for (int j = 0; j < J; ++j) {
      s_tb[j] = (int)t_start.size(); s_n[j] = (int)n_j[j];
      for (long long o = 0; o < n_j[j]; o += TILE) { t_start.push_back((int)(pos + o)); t_len.push_back((int)std::min<long long>(TILE, n_j[j] - o)); t_sample.push_back(j); }
      pos += n_j[j];
    }
    s_tb[J] = (int)t_start.size();
  }
  const int T = (int)t_start.size();

  s.D = D; s.J = J; s.K = K; s.T = T; s.N = N; s.N_total = S->N_total; s.cell_offset = S->world > 1 ? a->cell_offset : 0;
  s.em_iters = S->em_iters; s.mcd_iters = S->mcd_iters; s.rng_mode = S->rng_mode; s.track = S->track; s.mcd_alpha = S->mcd_alpha; s.seed = S->seed;
  s.z = S->alloc<uint8_t>(N);
  auto up_i = [&](const std::vector<int>& v) { int* p = S->alloc<int>(v.size(), false); CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, S->st)); return p; };
  s.tile_start = up_i(t_start); s.tile_len = up_i(t_len); s.tile_sample = up_i(t_sample); s.sample_tile_begin = up_i(s_tb); s.sample_n = up_i(s_n);
  pt.mark("tiles");
  CK(cudaStreamSynchronize(S->st));                 // the host vectors above go out of scope
  pt.mark("sync_x_upload");
  s.mu0 = S->alloc<double>(D); s.alpha = S->alloc<double>((size_t)J * D); s.beta = S->alloc<double>((size_t)K * D);
  s.sigma = S->alloc<double>((size_t)K * D * D); s.linv = S->alloc<double>((size_t)K * D * D); s.ainv = S->alloc<double>((size_t)K * D * D);
  s.logP = S->alloc<double>(K); s.Pk = S->alloc<double>(K);
  s.p_jk = S->alloc<double>((size_t)J * K); s.logp_jk = S->alloc<double>((size_t)J * K);
  s.dens_cons = S->alloc<double>((size_t)K * (SQC_TRI(SQC_MAXD) + SQC_MAXD)); s.dens_tab = S->alloc<double>((size_t)J * K * DENS_TAB); s.dens_slow = S->alloc<int>(J);
  s.t_ll1 = S->alloc<double>(T); s.t_ll2 = S->alloc<double>(T);
  s.t_nk = S->alloc<int>((size_t)T * K); s.t_sk = S->alloc<double>((size_t)T * K * D); s.s_base = S->alloc<int>((size_t)J * K); s.t_rel = S->alloc<int>((size_t)T * K); s.done_ctr = S->alloc<unsigned int>(4);
  s.t_col = S->alloc<double>((size_t)T * D);
  s.n_jk = S->alloc<int>((size_t)J * K); s.s_jk = S->alloc<double>((size_t)J * K * D);
  s.clu_n = S->alloc<long long>(K); s.clu_off = S->alloc<long long>(K); s.clu_n_glob = S->alloc<long long>(K); s.clu_prefix = S->alloc<long long>(K);
  s.like1 = S->alloc<double>(em + 1); s.like2 = S->alloc<double>(em + 1);
  s.ctl = S->alloc<Control>(1);
  S->xc.dev.status = &s.ctl->status;                  // (multi-GPU) exchange waits report a timeout here
  S->xc.dev.budget_ns = getenv("SQC_WAIT_BUDGET_MS") ? std::max(1LL, atoll(getenv("SQC_WAIT_BUDGET_MS"))) * 1000000LL : SQC_WAIT_BUDGET_NS;
  S->alpha0 = S->alloc<double>((size_t)J * D);
  if (S->track && variant == SQC_VARIANT_ROBUST) {
    s.track_alpha = S->alloc<double>((size_t)J * D * em); s.track_beta = S->alloc<double>((size_t)K * D * em);
    s.track_sigma = S->alloc<double>((size_t)K * D * D * em); s.track_p = S->alloc<double>((size_t)J * K * em);
  }
  // outputs, packed: mu0 | alpha | beta | sigma | p_jk
  S->o_doubles = (size_t)D + (size_t)J * D + (size_t)K * D + (size_t)K * D * D + (size_t)J * K;
  S->o_buf = S->alloc<double>(S->o_doubles); S->o_z = S->alloc<int>(N, false); S->o_korder = S->alloc<int>(2 * SQC_MAXK);
  S->optr.mu0 = S->o_buf; S->optr.alpha = S->optr.mu0 + D; S->optr.beta = S->optr.alpha + (size_t)J * D;
  S->optr.sigma = S->optr.beta + (size_t)K * D; S->optr.p_jk = S->optr.sigma + (size_t)K * D * D; S->optr.z = S->o_z; S->optr.korder = S->o_korder;

  // ---- upload ----------------------------------------------------------------------------------
  if (!sorted) {                                    // the optimistic upload above was in the wrong order
    CK(cudaStreamSynchronize(S->st));
    std::vector<double> col(N);
    for (int d = 0; d < D; ++d) {
      for (long long i = 0; i < N; ++i) col[i] = a->x[(size_t)d * ld + S->perm[i]];
      CK(cudaMemcpyAsync(s.x + (size_t)d * N, col.data(), (size_t)N * sizeof(double), cudaMemcpyHostToDevice, S->st));
      CK(cudaStreamSynchronize(S->st));
    }
  }
  if (S->world > 1) {
    try {
      if (grp) {
        // in-process ranks: clear the own window, enable peer access, meet, then take the peers' windows as plain pointers
        xchg_ensure(S->device, true);
        for (int r = 0; r < S->world; ++r) if (grp->device[r] != S->device) {
          cudaError_t e = cudaDeviceEnablePeerAccess(grp->device[r], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
          else if (e != cudaSuccess) throw std::runtime_error(std::string("cudaDeviceEnablePeerAccess failed: ") + cudaGetErrorString(e));
        }
        CK(cudaStreamSynchronize(S->st));               // x is on the device: a failure so far is local
        if (!grp->sync(true)) fail(SQC_ERR_COMM, "another GPU of the fit failed while the sessions were built");
        xchg_open_inproc(S->xc, S->device, S->rank, S->world, grp->device);
      } else xchg_open(S->xc, S->device, S->rank, S->world, a->peer_handles);
    }
    catch (const SqcFail&) { throw; }
    catch (const std::exception& e) { fail(SQC_ERR_COMM, "%s", e.what()); }
    S->xg_gather = S->alloc<unsigned long long>((size_t)XCHG_MAX_WORLD * SQC_MAXK); S->xg_vec = S->alloc<double>(16);
  }

  if (variant == SQC_VARIANT_ROBUST) {
    S->z0 = S->alloc<uint8_t>(N, false);
    int* bad = S->alloc<int>(1);
    {                                               // labels travel as int32 (the export's arma::uvec); narrow on the device
      int* tmp = S->o_z;                               // the output label buffer doubles as the int32 staging area
      if (sorted) CK(cudaMemcpyAsync(tmp, a->init_z, (size_t)N * sizeof(int), cudaMemcpyDefault, S->st));
      else { std::vector<int> zz(N); for (long long i = 0; i < N; ++i) zz[i] = a->init_z[S->perm[i]]; CK(cudaMemcpy(tmp, zz.data(), (size_t)N * sizeof(int), cudaMemcpyHostToDevice)); }
      k_widen_labels<<<std::min<long long>((N + 255) / 256, 4096), 256, 0, S->st>>>(tmp, S->z0, N, K, bad);
      int hb = 0; CK(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, S->st));
      CK(cudaStreamSynchronize(S->st));
      if (hb) fail(SQC_ERR_BAD_ARG, "init_z out of range");
    }
    s.rs_stride = ((N + 32ll * (K + 1)) + 31) & ~31ll;
    s.rs = S->alloc<double>((size_t)s.rs_stride * D);
    S->mcd_grid = do_mcd_grid(D, S->n_sm);
    if (S->rng_mode == SQC_RNG_R_COMPAT) {
      // R-compatible keys: the M-step kernel leaves some SMs to the key-stream kernels of the NEXT call, which run beside
      // it on the side stream (a CTA of k_mcd owns a whole SM, so they cannot share one; without free SMs the generation
      // falls onto the E-step / partition passes and onto the critical path: +50 % per iteration at C3).  k_mcd is
      // HBM / latency bound, not SM bound.
      // No more than the generation has blocks (one CTA each): a rank of a multi-GPU fit generates only its share of a call.
      const bool shares = S->world > 1 && S->N_total >= 624LL * 8 * S->world && !getenv("SQC_KEYS_REPLICATED");
      const long long share = shares ? (S->N_total + S->world - 1) / S->world : S->N_total;
      // a rank's share is cut into about as many blocks as SMs are reserved, so that the jump kernel (one CTA per block,
      // 157 us each) runs in ONE wave: with 73 blocks of a 5 M-key share on 24 SMs it took three (0.42 ms) and the chain
      // fell behind k_mcd onto the partition pass and the key fence (profiles/r2s_timeline_w2_rank0.csv)
      if (shares && !getenv("SQC_MT_BLOCKS")) S->mt_target_blocks = SQC_MCD_RESERVE_SMS;
      const long long blk = MT_BLOCK_WORDS * std::max<long long>(1, std::min<long long>(3, (share / MT_BLOCK_WORDS + S->mt_target_blocks / 2) / S->mt_target_blocks));
      int reserve = getenv("SQC_MCD_RESERVE") ? atoi(getenv("SQC_MCD_RESERVE")) : (int)std::min<long long>(SQC_MCD_RESERVE_SMS, (share + blk - 1) / blk + 1) + (shares ? 2 : 0);   // + the two CTAs of the state chain
      reserve = std::max(0, std::min(reserve, S->mcd_grid - 2 * K - 8));
      S->mcd_grid -= reserve;
    }
    McdBufs& mb = S->mb;
    mb.work = S->alloc<McdWork>(1);
    mb.hist = S->alloc<unsigned int>((size_t)K * MCD_NB);
    mb.part = S->alloc<double>((size_t)S->mcd_grid * K * SQC_NMOM(D));
    const size_t nreg = (size_t)K * S->mcd_grid * MCD_WARPS;                 // candidate regions: one per (cluster, CTA, warp)
    mb.cand_rec = S->alloc<unsigned long long>(nreg * MCD_CAPW * (2 + D), false);
    mb.cand_cnt = S->alloc<unsigned int>(nreg);
    mb.fhist = S->alloc<unsigned int>((size_t)K * (MCD_NF + 4));
    mb.pred = S->alloc<McdPred>(K);
    mb.xc = S->xc.dev;
    mb.epoch = S->alloc<unsigned int>(2 * SQC_MAXK + 4); mb.arrive = mb.epoch + SQC_MAXK; mb.status = (int*)(mb.epoch + 2 * SQC_MAXK);
    s.mcd_sync = mb.epoch;
    mb.join = getenv("SQC_MCD_JOIN") ? atoi(getenv("SQC_MCD_JOIN")) : 0;     // 1: worker classes join (faster tail, sums reproducible to rounding only)
    mb.trace_leader = getenv("SQC_MCD_TRACE") ? atoi(getenv("SQC_MCD_TRACE")) : 0;
    mb.trace = getenv("SQC_MCD_TRACE") ? S->alloc<long long>(16 * 512) : nullptr;
    if (S->rng_mode == SQC_RNG_R_COMPAT) {
      S->mt_state = S->alloc<uint32_t>(640 * 8);
      if (S->world > 1 && S->N_total >= 624LL * 8 * S->world && !getenv("SQC_KEYS_REPLICATED")) key_stream_setup(S);
      else S->keybuf = S->alloc<int>(2 * (size_t)S->N_total, false);
      if (getenv("SQC_RNG_INLINE") && atoi(getenv("SQC_RNG_INLINE"))) S->st_rng = S->st;      // debug: no side stream
      else CK(cudaStreamCreateWithFlags(&S->st_rng, cudaStreamNonBlocking));
      if (getenv("SQC_RNG_AFTER_MCD")) S->rng_after_mcd = atoi(getenv("SQC_RNG_AFTER_MCD")) != 0;
      if (getenv("SQC_MT_BLOCKS")) S->mt_target_blocks = std::max(1, atoi(getenv("SQC_MT_BLOCKS")));
      for (int i = 0; i < 2; ++i) { CK(cudaEventCreateWithFlags(&S->ev_key_ready[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&S->ev_key_free[i], cudaEventDisableTiming)); }
      CK(cudaEventCreateWithFlags(&S->ev_pre_mcd, cudaEventDisableTiming));
      S->mt_seq = S->alloc<uint32_t>(MT_SEQ_WORDS + (size_t)(SQC_MT_JUMP_ROWS + 2) * MT_N);
      S->mt_table = S->alloc<uint32_t>((size_t)SQC_MT_JUMP_ROWS * MT_N, false);
      CK(cudaMemcpyAsync(S->mt_table, sqc_mt_jump_table, sizeof sqc_mt_jump_table, cudaMemcpyHostToDevice, S->st));
      CK(cudaFuncSetAttribute(k_mt_jump, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MT_SMEM));
    }
  } else {
    mle_alloc(*S, a, S->ml);
  }
  pt.mark("allocs+labels");
  // ---- centring: mu_0 = colMeans(x) (:478), x <- x - mu_0 (centre_x), alpha_j = per-sample means ----
  S->stats = sqc_run_stats{};
  do_init_passes(S);
  CK(cudaMemcpyAsync(S->alpha0, s.alpha, (size_t)J * D * sizeof(double), cudaMemcpyDeviceToDevice, S->st));
  CK(cudaStreamSynchronize(S->st));
  if (S->world > 1) {                                 // the column means were exchanged: a peer that never showed up is known by now
    int st_dev = 0;
    CK(cudaMemcpy(&st_dev, &s.ctl->status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st_dev == SQC_D_TIMEOUT) fail(SQC_ERR_COMM, "%s", device_status_message(st_dev));
  }
  CK(cudaEventCreate(&S->ev_run0)); CK(cudaEventCreate(&S->ev_init)); CK(cudaEventCreate(&S->ev_loop));
  pt.mark("init_passes");
  pt.dump("create");
  return up.release();
}

#include "sqc_keys_host.inl"

void run_robust(sqc_session* S) {
  Dev& s = S->dev;
  const int J = S->J;
  CK(cudaEventRecord(S->ev_run0, S->st));
  { Launch L(S, SQC_KF_UPDATE); k_ctl_reset<<<1, 32, 0, S->st>>>(s); L.done(); }
  CK(cudaMemcpyAsync(s.z, S->z0, (size_t)S->N, cudaMemcpyDeviceToDevice, S->st));
  CK(cudaMemcpyAsync(s.alpha, S->alpha0, (size_t)J * S->D * sizeof(double), cudaMemcpyDeviceToDevice, S->st));   // init_alpha_j, :482
  CK(cudaMemsetAsync(s.beta, 0, (size_t)S->K * S->D * sizeof(double), S->st));
  CK(cudaMemsetAsync(s.sigma, 0, (size_t)S->K * S->D * S->D * sizeof(double), S->st));
  CK(cudaMemsetAsync(s.like1, 0, (size_t)(std::max(S->em_iters, 0) + 1) * sizeof(double), S->st));
  CK(cudaMemsetAsync(s.like2, 0, (size_t)(std::max(S->em_iters, 0) + 1) * sizeof(double), S->st));
  CK(cudaMemsetAsync(S->mb.pred, 0xFF, (size_t)S->K * sizeof(McdPred), S->st));          // no threshold history: call[] = -1
  if (S->mt_state) {                                                                    // set_seed_cpp, :475
    CK(cudaEventRecord(S->ev_key_free[0], S->st)); CK(cudaEventRecord(S->ev_key_free[1], S->st));   // after k_ctl_reset
    if (S->key_q) {
      CK(cudaStreamWaitEvent(S->st_rng2, S->ev_key_free[0], 0));
      { Launch L(S, SQC_KF_RNG, S->st_rng2, "seed"); k_mt_seed<<<1, 32, 0, S->st_rng2>>>(S->mt_state + 640 * 2, S->seed); L.done(); }
      enqueue_chain(S, 0, nullptr);
      if (1 <= S->em_iters) enqueue_chain(S, 1, nullptr);
    } else {
      CK(cudaStreamWaitEvent(S->st_rng, S->ev_key_free[0], 0));
      { Launch L(S, SQC_KF_RNG); k_mt_seed<<<1, 32, 0, S->st_rng>>>(S->mt_state + 640 * 2, S->seed); L.done(); }
    }
    enqueue_keys(S, 0, nullptr);
  }
  // single GPU: the end-of-iteration bookkeeping and the next iteration's alpha_j ride in k_post_stats (SQC_FUSE_ALPHA=0: own launch)
  static const bool fuse_alpha_env = !(getenv("SQC_FUSE_ALPHA") && atoi(getenv("SQC_FUSE_ALPHA")) == 0);
  const int loop_flags = S->world > 1 ? 0 : ((fuse_alpha_env && S->D <= POST_ALPHA_MAXD) ? 3 : 1);
  auto post_stats = [&](int flags) { do_post_stats(S, flags); };
  auto m_step = [&](int call, int like) {
    do_partition(S, like);
    use_keys(S, call);
    do_mcd(S);
    after_mcd_keys(S, call);
    release_keys(S, call);
    do_after_mstep(S, like);                     // derived constants; inside the loop also p_jk (:504)
  };
  // initial labelling: p_jk (:481), alpha_j (:482), beta_k / Sigma_k (:483)
  do_estep(S, 1);
  post_stats(0);
  // distributed key stream: every rank's share of a call has to be complete before any rank's M-step kernel reads it.  The
  // per-iteration exchange that precedes the call doubles as that fence: the stream waits for the local share first.
  if (S->key_q) CK(cudaStreamWaitEvent(S->st, S->ev_key_ready[0], 0));
  if (S->world > 1) xchg_iteration(S, s, 0);
  { Launch L(S, SQC_KF_UPDATE); k_update_p<<<(J + 127) / 128, 128, 0, S->st>>>(s); L.done(); }
  m_step(0, 0);
  { Launch L(S, SQC_KF_UPDATE); k_ctl_start_loop<<<1, 32, 0, S->st>>>(s); L.done(); }
  CK(cudaEventRecord(S->ev_init, S->st));
  for (int it = 0; it < S->em_iters; ++it) {                                            // :490
    if (it == 0 || !(loop_flags & 2)) do_update_alpha(S);                               // :496 (later iterations: inside the k_post_stats before)
    m_step(it + 1, 1);                                                                  // like_1 :499, M-step :502, p_jk :504
    do_estep(S, 0);                                                                     // :506, :509, :519
    post_stats(loop_flags);                                                             // single GPU: :511-521 and the next :496 fused in
    if (S->key_q && it + 2 <= S->em_iters) CK(cudaStreamWaitEvent(S->st, S->ev_key_ready[(it + 2) & 1], 0));   // keys of the next call (fence, see above)
    if (S->world > 1) xchg_iteration(S, s, 1);     // cluster layout of the next M-step + like / z_delta / status + loop control
  }
  CK(cudaEventRecord(S->ev_loop, S->st));
}

void session_run(sqc_session* S) {
  CK(cudaSetDevice(S->device));
  S->stats = sqc_run_stats{};
  S->ev_used = 0;
  if (S->variant == SQC_VARIANT_ROBUST) run_robust(S);
  else mle_run(S);
  CK(cudaMemcpyAsync(&S->h_ctl, S->dev.ctl, sizeof(Control), cudaMemcpyDeviceToHost, S->st));
  CK(cudaStreamSynchronize(S->st));
  if (S->st_rng) CK(cudaStreamSynchronize(S->st_rng));
  if (S->st_rng2) CK(cudaStreamSynchronize(S->st_rng2));
  S->ran = true;
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, S->ev_run0, S->ev_init)); S->stats.init_ms = ms;
  CK(cudaEventElapsedTime(&ms, S->ev_init, S->ev_loop)); S->stats.loop_ms = ms;
  S->stats.n_iter = S->h_ctl.iter;
  S->stats.mcd_passes = S->h_ctl.mcd_passes; S->stats.mcd_csteps = S->h_ctl.mcd_csteps;
  for (size_t i = 0; i < S->ev_used; ++i) {
    CK(cudaEventElapsedTime(&ms, S->evs[i].a, S->evs[i].b));
    S->stats.family_ms[S->evs[i].fam] += ms;
  }
  if (const char* tl = S->ev_used ? getenv("SQC_TIMELINE") : nullptr) {                          // debug: one CSV row per launch (ms since the run began)
    static const char* fam_names[] = {"estep", "partition", "mcd", "update", "rng", "init", "mle_pass", "final"};
    char path[512]; snprintf(path, sizeof path, "%s.rank%d.csv", tl, S->rank);
    if (FILE* f = fopen(path, "w")) {
      fprintf(f, "launch,family,tag,t0_ms,t1_ms\n");
      for (size_t i = 0; i < S->ev_used; ++i) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, S->ev_run0, S->evs[i].a); cudaEventElapsedTime(&t1, S->ev_run0, S->evs[i].b);
        fprintf(f, "%zu,%s,%s,%.4f,%.4f\n", i, fam_names[S->evs[i].fam], S->evs[i].tag ? S->evs[i].tag : "", t0, t1);
      }
      fclose(f);
    }
  }
  // algorithmic bytes: one pass = N (8 D + 5) bytes; E-step writes N more
  const double pass = (double)S->N * (8.0 * S->D + 5.0);
  const int it = S->h_ctl.iter;
  if (S->variant == SQC_VARIANT_ROBUST) {
    S->stats.family_bytes[SQC_KF_ESTEP] = (double)it * (pass + (double)S->N);
    S->stats.family_bytes[SQC_KF_PARTITION] = (double)(it + 1) * (pass + 8.0 * S->D * (double)S->N);
    S->stats.family_bytes[SQC_KF_MCD] = (double)S->h_ctl.mcd_bytes;
  } else {
    S->stats.family_bytes[SQC_KF_MLE_PASS] = (double)(2 * it + 1) * pass;
  }
  if (S->h_ctl.status != 0) fail((S->h_ctl.status == SQC_D_PEER || S->h_ctl.status == SQC_D_TIMEOUT) ? SQC_ERR_COMM : S->h_ctl.status, "%s", device_status_message(S->h_ctl.status));
}

void reorder_track(const sqc_session* S, sqc_fit_out* o, const int* korder) {
  const int D = S->D, J = S->J, K = S->K, em = std::max(S->em_iters, 0);
  std::vector<double> tmp;
  for (int it = 0; it < em; ++it) {                 // robust_cpp.cpp:533-537
    if (o->track_beta_k) { double* b = o->track_beta_k + (size_t)it * K * D; tmp.assign(b, b + K * D);
      for (int k = 0; k < K; ++k) for (int d = 0; d < D; ++d) b[k + K * d] = tmp[korder[k] + K * d]; }
    if (o->track_sigma_k) { double* g = o->track_sigma_k + (size_t)it * D * D * K; tmp.assign(g, g + D * D * K);
      for (int k = 0; k < K; ++k) memcpy(g + (size_t)k * D * D, tmp.data() + (size_t)korder[k] * D * D, sizeof(double) * D * D); }
    if (o->track_p_jk) { double* p = o->track_p_jk + (size_t)it * J * K; tmp.assign(p, p + J * K);
      for (int k = 0; k < K; ++k) for (int j = 0; j < J; ++j) p[j + J * k] = tmp[j + J * korder[k]]; }
  }
}

void session_fetch(sqc_session* S, sqc_fit_out* o) {
  if (!o) fail(SQC_ERR_BAD_ARG, "null output");
  if (!S->ran) fail(SQC_ERR_BAD_ARG, "sqc_session_fetch before sqc_session_run");
  CK(cudaSetDevice(S->device));
  Dev& s = S->dev;
  const int D = S->D, J = S->J, K = S->K, em = std::max(S->em_iters, 0); const long long N = S->N;
  if (S->variant == SQC_VARIANT_MLE) { mle_fetch(S, o); return; }
  OutPtrs op = S->optr;
  if (!o->z) op.z = nullptr;
  k_order_components<<<1, 32, 0, S->st>>>(s, op);
  k_pack_outputs<<<std::max(1, std::min<int>((int)((N + 255) / 256), S->n_sm * 8)), 256, 0, S->st>>>(s, op, S->em_iters <= 0 ? 1 : 0);
  CK(cudaGetLastError());
  std::vector<double> h(S->o_doubles);
  int korder[2 * SQC_MAXK];
  CK(cudaMemcpyAsync(h.data(), S->o_buf, S->o_doubles * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  CK(cudaMemcpyAsync(korder, S->o_korder, sizeof korder, cudaMemcpyDeviceToHost, S->st));
  if (o->z) {
    if (S->sorted) CK(cudaMemcpyAsync(o->z, S->o_z, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, S->st));
    else { std::vector<int> zz(N); CK(cudaMemcpyAsync(zz.data(), S->o_z, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, S->st)); CK(cudaStreamSynchronize(S->st));
           for (long long i = 0; i < N; ++i) o->z[S->perm[i]] = zz[i]; }
  }
  if (o->like_1) CK(cudaMemcpyAsync(o->like_1, s.like1, (size_t)em * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  if (o->like_2) CK(cudaMemcpyAsync(o->like_2, s.like2, (size_t)em * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  if (S->track) {
    if (o->track_alpha_j) CK(cudaMemcpyAsync(o->track_alpha_j, s.track_alpha, (size_t)J * D * em * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    if (o->track_beta_k) CK(cudaMemcpyAsync(o->track_beta_k, s.track_beta, (size_t)K * D * em * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    if (o->track_sigma_k) CK(cudaMemcpyAsync(o->track_sigma_k, s.track_sigma, (size_t)K * D * D * em * sizeof(double), cudaMemcpyDeviceToHost, S->st));
    if (o->track_p_jk) CK(cudaMemcpyAsync(o->track_p_jk, s.track_p, (size_t)J * K * em * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  }
  if (S->mt_state) {
    uint32_t st[625];
    CK(cudaMemcpyAsync(st, S->mt_state + 640 * (S->h_ctl.iter % 3), sizeof st, cudaMemcpyDeviceToHost, S->st));   // state after call n_iter
    CK(cudaStreamSynchronize(S->st));
    o->rng_state[0] = (int32_t)st[624];
    for (int i = 0; i < 624; ++i) o->rng_state[i + 1] = (int32_t)st[i];
  } else memset(o->rng_state, 0, sizeof o->rng_state);
  CK(cudaStreamSynchronize(S->st));
  const double* p = h.data();
  if (o->mu_0) memcpy(o->mu_0, p, sizeof(double) * D); p += D;
  auto rows_out = [&](double* dst, const double* src, int ncol) {    // J x ncol column-major; internal sample order -> the caller's codes
    if (S->smap.empty()) { memcpy(dst, src, sizeof(double) * J * ncol); return; }
    for (int c = 0; c < ncol; ++c) for (int r = 0; r < J; ++r) dst[S->smap[r] + (size_t)J * c] = src[r + (size_t)J * c];
  };
  if (o->alpha_j) rows_out(o->alpha_j, p, D); p += (size_t)J * D;
  if (o->beta_k) memcpy(o->beta_k, p, sizeof(double) * K * D); p += (size_t)K * D;
  if (o->sigma_k) memcpy(o->sigma_k, p, sizeof(double) * K * D * D); p += (size_t)K * D * D;
  if (o->p_jk) rows_out(o->p_jk, p, K);
  if (S->track) reorder_track(S, o, korder);
  if (S->track && !S->smap.empty()) {
    std::vector<double> tmp;
    for (int it = 0; it < em; ++it) {
      if (o->track_alpha_j) { double* b = o->track_alpha_j + (size_t)it * J * D; tmp.assign(b, b + (size_t)J * D); rows_out(b, tmp.data(), D); }
      if (o->track_p_jk) { double* b = o->track_p_jk + (size_t)it * J * K; tmp.assign(b, b + (size_t)J * K); rows_out(b, tmp.data(), K); }
    }
  }
  o->n_iter = S->h_ctl.iter; o->mcd_iters_hit = S->h_ctl.mcd_hit; o->n_zero_like = S->h_ctl.n_zero;
  o->rng_draws = (long long)(1 + S->h_ctl.iter) * S->N_total;
}

template <class F> int guarded(F&& f, char* err, size_t errlen) {
  try { f(); if (err && errlen) err[0] = 0; return SQC_OK; }
  catch (const SqcFail& e) { if (err && errlen) snprintf(err, errlen, "%s", e.msg.c_str()); return e.code; }
  catch (const std::bad_alloc&) { if (err && errlen) snprintf(err, errlen, "host allocation failed"); return SQC_ERR_NOMEM; }
  catch (const std::exception& e) { if (err && errlen) snprintf(err, errlen, "%s", e.what()); return SQC_ERR_CUDA; }
}

double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

int one_shot(const sqc_fit_args* a, sqc_fit_out* o, int variant, char* err, size_t errlen) {
  sqc_session* S = nullptr;
  const bool timing = getenv("SQC_TIMING") != nullptr;
  double t0 = now_ms(), t1 = t0, t2 = t0, t3 = t0;
  int st = guarded([&] {
    if (!o) fail(SQC_ERR_BAD_ARG, "null output");
    S = session_create(a, variant); t1 = now_ms();
    session_run(S); t2 = now_ms();
    session_fetch(S, o); t3 = now_ms();
  }, err, errlen);
  delete S;
  if (timing) fprintf(stderr, "[sqc timing] create %.2f ms  run %.2f ms  fetch %.2f ms  destroy %.2f ms\n", t1 - t0, t2 - t1, t3 - t2, now_ms() - t3);
  return st;
}

#include "sqc_inproc_host.inl"

}  // namespace

#include "sqc_mle_host.inl"

extern "C" {

// n_gpus > 1 (or SQC_ALL_GPUS): the library shards the call over the GPUs of this process (sqc_inproc_host.inl)
static int fit_entry(const sqc_fit_args* a, sqc_fit_out* o, int variant, char* err, size_t errlen) {
  if (a && a->abi_version == SQC_ABI_VERSION) {
    const int n = resolve_n_gpus(a);
    if (n > 1) { const int st = fit_inproc(a, o, variant, n, err, errlen); if (st != -1000) return st; }
  }
  return one_shot(a, o, variant, err, errlen);
}
int sqc_fit_robust(const sqc_fit_args* a, sqc_fit_out* o, char* err, size_t errlen) { return fit_entry(a, o, SQC_VARIANT_ROBUST, err, errlen); }
int sqc_fit_mle(const sqc_fit_args* a, sqc_fit_out* o, char* err, size_t errlen) { return fit_entry(a, o, SQC_VARIANT_MLE, err, errlen); }

int sqc_session_create(const sqc_fit_args* a, int variant, sqc_session** out, char* err, size_t errlen) {
  if (out) *out = nullptr;
  return guarded([&] {
    if (!out) fail(SQC_ERR_BAD_ARG, "null session pointer");
    if (variant != SQC_VARIANT_ROBUST && variant != SQC_VARIANT_MLE) fail(SQC_ERR_BAD_ARG, "unknown variant");
    *out = session_create(a, variant);
  }, err, errlen);
}
int sqc_session_run(sqc_session* s, char* err, size_t errlen) { return guarded([&] { if (!s) fail(SQC_ERR_BAD_ARG, "null session"); session_run(s); }, err, errlen); }
int sqc_session_fetch(sqc_session* s, sqc_fit_out* o, char* err, size_t errlen) { return guarded([&] { if (!s) fail(SQC_ERR_BAD_ARG, "null session"); session_fetch(s, o); }, err, errlen); }
int sqc_session_stats(const sqc_session* s, sqc_run_stats* st) { if (!s || !st) return SQC_ERR_BAD_ARG; *st = s->stats; return SQC_OK; }
// debug: per-round timeline of the LAST k_mcd launch (only when SQC_MCD_TRACE is set at session creation)
SQC_API int sqc_session_mcd_trace(sqc_session* s, long long* out, int n_words) {
  if (!s || !out || !s->mb.trace) return SQC_ERR_BAD_ARG;
  cudaSetDevice(s->device);
  return cudaMemcpy(out, s->mb.trace, sizeof(long long) * std::min(n_words, 16 * 512), cudaMemcpyDeviceToHost) == cudaSuccess ? SQC_OK : SQC_ERR_CUDA;
}
int sqc_session_profile(sqc_session* s, int on) { if (!s) return SQC_ERR_BAD_ARG; s->profile = on != 0; return SQC_OK; }
void sqc_session_destroy(sqc_session* s) { delete s; }

int sqc_outlier_maha(const sqc_maha_args* a, double* maha, uint8_t* outlier, double* maha_cut, char* err, size_t errlen) {
  return guarded([&] {
    if (!a || !a->x || !a->groups || !a->mu_0 || !a->alpha_j || !a->beta_k || !a->sigma_k) fail(SQC_ERR_BAD_ARG, "null input");
    if (a->abi_version != SQC_ABI_VERSION) fail(SQC_ERR_BAD_ARG, "ABI version mismatch");
    const int D = a->D, J = a->J, K = a->K; const long long N = a->N;
    if (D < 1 || D > SQC_MAXD || K < 1 || K > 64 || J < 1 || N < 1) fail(SQC_ERR_BAD_ARG, "bad dimensions");
    int dev; select_device(a->device, &dev);
    // solve(cov) on the host: K tiny inverses (stats::mahalanobis)
    std::vector<double> ainv((size_t)K * D * D);
    for (int k = 0; k < K; ++k) if (!small_inv(a->sigma_k + (size_t)k * D * D, D, ainv.data() + (size_t)k * D * D)) fail(SQC_ERR_SINGULAR, "solve(): system is exactly singular");
    // device buffers from the per-device arena the fits use; copies on a stream (asynchronous from pinned memory), the
    // host checks the groups while x is on its way
    cudaStream_t st = nullptr;
    std::vector<Chunk> chunks;
    auto take = [&](size_t bytes) { chunks.push_back(g_chunks.take(dev, std::max<size_t>(bytes, 256))); return (void*)chunks.back().base; };
    auto cleanup = [&] { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } for (auto& c : chunks) g_chunks.give(dev, c); };
    try {
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      const size_t npar = (size_t)D + (size_t)J * D + (size_t)K * D + (size_t)K * D * D + 4;
      double* dx = (double*)take((size_t)N * D * sizeof(double)); int* dg = (int*)take((size_t)N * sizeof(int));
      double* dpar = (double*)take(npar * sizeof(double));
      double* dm = maha ? (double*)take((size_t)N * K * sizeof(double)) : nullptr;
      uint8_t* dout = outlier ? (uint8_t*)take((size_t)N) : nullptr;
      CK(cudaMemcpyAsync(dx, a->x, (size_t)N * D * sizeof(double), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(dg, a->groups, (size_t)N * sizeof(int), cudaMemcpyHostToDevice, st));
      std::vector<double> hp(npar);
      double* q = hp.data();
      memcpy(q, a->mu_0, D * sizeof(double)); q += D;
      memcpy(q, a->alpha_j, (size_t)J * D * sizeof(double)); q += (size_t)J * D;
      memcpy(q, a->beta_k, (size_t)K * D * sizeof(double)); q += (size_t)K * D;
      memcpy(q, ainv.data(), (size_t)K * D * D * sizeof(double)); q += (size_t)K * D * D;
      q[0] = 1.0 - a->alpha; q[1] = (double)D; q[2] = 0.0;               // qchisq(1 - alpha, df = D), :429
      CK(cudaMemcpyAsync(dpar, hp.data(), npar * sizeof(double), cudaMemcpyHostToDevice, st));
      for (long long i = 0; i < N; ++i) if ((unsigned)a->groups[i] >= (unsigned)J) fail(SQC_ERR_BAD_ARG, "groups out of range at cell %lld", i);
      double* pm = dpar; double* pa = pm + D; double* pb = pa + (size_t)J * D; double* pi = pb + (size_t)K * D; double* pq = pi + (size_t)K * D * D;
      k_qchisq<<<1, 32, 0, st>>>(pq, pq + 1, pq + 2, 1);
      double cut = 0.0;
      CK(cudaMemcpyAsync(&cut, pq + 2, sizeof(double), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (maha_cut) *maha_cut = cut;
      int n_sm = 148; CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
      const int grid = (int)std::min<long long>((N + 255) / 256, (long long)n_sm * 16);   // a few waves of 256-thread CTAs, grid-stride
      const size_t sm = (size_t)K * (D + D * D) * sizeof(double);
#define SQC_CASE_MAHA(DD) case DD: k_outlier_maha<DD><<<grid, 256, sm, st>>>(dx, dg, N, J, K, pm, pa, pb, pi, cut, dm, dout); break;
      switch (D) { SQC_FOR_D(SQC_CASE_MAHA) default: fail(SQC_ERR_BAD_ARG, "D = %d not compiled in", D); }
      CK(cudaGetLastError());
      if (maha) CK(cudaMemcpyAsync(maha, dm, (size_t)N * K * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (outlier) CK(cudaMemcpyAsync(outlier, dout, (size_t)N, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    } catch (...) { cleanup(); throw; }
    cleanup();
  }, err, errlen);
}

#include "sqc_init_host.inl"
#include "sqc_pipe_host.inl"
#include "sqc_mmd_host.inl"

int sqc_xchg_export(int device, void* handle_out, char* err, size_t errlen) {
  return guarded([&] {
    if (!handle_out) fail(SQC_ERR_BAD_ARG, "null handle buffer");
    int d; select_device(device, &d);
    try { xchg_export(d, handle_out); } catch (const std::exception& e) { fail(SQC_ERR_COMM, "%s", e.what()); }
  }, err, errlen);
}
int sqc_xchg_release(int device) {
  // call it on every rank once no rank will run another multi-GPU fit: it closes this process' mappings of the peers'
  // windows / key buffers and frees its own
  const int st = xchg_release(device);
  KeyBufLocal& kb = key_buf_local(device);
  for (int* p : kb.retired) cudaFree(p);
  kb.retired.clear();
  if (kb.p) { cudaFree(kb.p); kb.p = nullptr; kb.ints = 0; }
  return st;
}

int sqc_trim(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return SQC_ERR_CUDA;
  for (int d = 0; d < n; ++d) if (device < 0 || device == d) { cudaSetDevice(d); g_chunks.trim(d); }
  return SQC_OK;
}
int sqc_abi_version(void) { return SQC_ABI_VERSION; }
int sqc_device_count(void) { int n = 0; cudaError_t e = cudaGetDeviceCount(&n); return e == cudaSuccess ? n : -1; }
const char* sqc_status_name(int st) {
  switch (st) {
    case SQC_OK: return "SQC_OK"; case SQC_ERR_BAD_ARG: return "SQC_ERR_BAD_ARG"; case SQC_ERR_EMPTY_CLUSTER: return "SQC_ERR_EMPTY_CLUSTER";
    case SQC_ERR_SUBSET_SIZE: return "SQC_ERR_SUBSET_SIZE"; case SQC_ERR_NOT_SPD: return "SQC_ERR_NOT_SPD"; case SQC_ERR_SINGULAR: return "SQC_ERR_SINGULAR";
    case SQC_ERR_NAN: return "SQC_ERR_NAN"; case SQC_ERR_SELECT: return "SQC_ERR_SELECT"; case SQC_ERR_CUDA: return "SQC_ERR_CUDA";
    case SQC_ERR_COMM: return "SQC_ERR_COMM"; case SQC_ERR_NOMEM: return "SQC_ERR_NOMEM"; default: return "SQC_ERR_UNKNOWN";
  }
}
#include "sqc_hooks_host.inl"

}  // extern "C"
