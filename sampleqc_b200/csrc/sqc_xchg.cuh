// sqc_xchg.cuh — cross-GPU exchange through peer memory (NVLink 5 / NVSwitch), one process per GPU.
//
// Cells are sharded by sample: alpha_j, p_jk and the labels never leave their GPU; what
// the ranks must agree on are a few small vectors per step (histograms of the exact selection, candidate
// lists, K x (1 + D + D^2) moments, log-likelihood sums).  Those are exchanged INSIDE the kernels that
// produce them: every rank owns a window in its HBM (cudaMalloc, exported with cudaIpcGetMemHandle and mapped
// by all peers), publishes its vector there and a monotonically increasing sequence number behind it
// (st.release.sys), then reads every peer's copy over NVLink (ld.acquire.sys on the flag, ld.relaxed.sys on
// the data) and combines the copies in RANK ORDER — so all ranks hold bit-identical results and take
// identical control-flow decisions without any host round trip or NCCL launch.  Messages are <= 32 KB:
// what matters is latency, not the 900 GB/s links.  Slots are double-buffered by sequence parity: a rank can
// publish exchange q + 1 only after it has read everybody's q, so a slot is never overwritten while a peer
// may still read it.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "sqc_common.cuh"

namespace sqc {

constexpr int XCHG_MAX_WORLD = 8;
constexpr int XCHG_CHANNELS = SQC_MAXK + 4;          // one per cluster leader + host-stream level channels
constexpr int XCHG_CH_HOST = SQC_MAXK;               // first host-stream channel
constexpr int XCHG_SLOT_BYTES = 32 * 1024;
constexpr size_t XCHG_WINDOW_BYTES = 4096 + (size_t)XCHG_CHANNELS * 2 * XCHG_SLOT_BYTES;

struct XchgDev {
  int world, rank;
  unsigned char* peer[XCHG_MAX_WORLD];               // window base of every rank (own window: local pointer)
  unsigned long long* next_seq;                      // XCHG_CHANNELS counters in LOCAL memory (same value on all ranks)
};

__device__ __forceinline__ unsigned long long* xw_flag(unsigned char* base, int ch) { return (unsigned long long*)base + ch; }
__device__ __forceinline__ unsigned char* xw_slot(unsigned char* base, int ch, unsigned long long seq) {
  return base + 4096 + ((size_t)ch * 2 + (size_t)(seq & 1ull)) * XCHG_SLOT_BYTES;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_sys_u32(const unsigned int* p) {
  unsigned int v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}

// Block-wide: publish n_words 32-bit words from `src` (shared or global) on channel ch, wait for every rank.
// Returns the sequence number; afterwards xchg_peer_slot(x, ch, seq, r) is readable for every rank r.
__device__ inline unsigned long long xchg_publish(const XchgDev& x, int ch, const unsigned int* src, int n_words) {
  __shared__ unsigned long long sh_seq;
  if (threadIdx.x == 0) sh_seq = ++x.next_seq[ch];
  __syncthreads();
  const unsigned long long seq = sh_seq;
  unsigned int* dst = (unsigned int*)xw_slot(x.peer[x.rank], ch, seq);
  for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = src[i];
  __syncthreads();                                   // the block's stores happen-before thread 0's release (cumulativity)
  if (threadIdx.x == 0) { __threadfence_system(); st_release_sys_u64(xw_flag(x.peer[x.rank], ch), seq); }
  if ((int)threadIdx.x < x.world) {
    const unsigned long long* f = xw_flag(x.peer[threadIdx.x], ch);
    while (ld_acquire_sys_u64(f) < seq) { }
  }
  __syncthreads();
  return seq;
}
__device__ __forceinline__ const unsigned int* xchg_peer_slot(const XchgDev& x, int ch, unsigned long long seq, int r) {
  return (const unsigned int*)xw_slot(x.peer[r], ch, seq);
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const unsigned int* p) {      // 16-byte aligned; not volatile: loads may overlap
  uint4 v; asm("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// sum of 32-bit counters over ranks, in place (block-wide).  n is rounded up to a multiple of 4 words (the slots
// are 16-byte aligned and larger than any message); every thread issues all its remote 128-bit loads before it
// uses any of them, so one exchange costs about one NVLink round trip.
__device__ inline void xchg_sum_u32(const XchgDev& x, int ch, unsigned int* buf, int n) {
  const unsigned long long seq = xchg_publish(x, ch, buf, n);
  const int n4 = (n + 3) >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    uint4 v[XCHG_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) v[r] = ld_relaxed_sys_v4(xchg_peer_slot(x, ch, seq, r) + 4 * i);
    uint4 a = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
    if (4 * i + 0 < n) buf[4 * i + 0] = a.x;
    if (4 * i + 1 < n) buf[4 * i + 1] = a.y;
    if (4 * i + 2 < n) buf[4 * i + 2] = a.z;
    if (4 * i + 3 < n) buf[4 * i + 3] = a.w;
  }
  __syncthreads();
}
// sum of doubles over ranks in rank order (bit-identical on every rank), in place (block-wide)
__device__ inline void xchg_sum_f64(const XchgDev& x, int ch, double* buf, int n) {
  const unsigned long long seq = xchg_publish(x, ch, (const unsigned int*)buf, 2 * n);
  const int n2 = (n + 1) >> 1;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    uint4 v[XCHG_MAX_WORLD];
#pragma unroll
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) v[r] = ld_relaxed_sys_v4(xchg_peer_slot(x, ch, seq, r) + 4 * i);
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) { a0 += __hiloint2double((int)v[r].y, (int)v[r].x); a1 += __hiloint2double((int)v[r].w, (int)v[r].z); }
    buf[2 * i] = a0;
    if (2 * i + 1 < n) buf[2 * i + 1] = a1;
  }
  __syncthreads();
}

// host-stream level reductions (one CTA); used between the streaming kernels of the EM loop
__global__ void k_xchg_sum_f64(XchgDev x, int ch, double* buf, int n, const int* cont) {
  if (cont && !*cont) return;
  xchg_sum_f64(x, ch, buf, n);
}
// all-gather of n 64-bit words per rank: out[r * n + i]
__global__ void k_xchg_gather_u64(XchgDev x, int ch, const unsigned long long* in, unsigned long long* out, int n, const int* cont) {
  if (cont && !*cont) return;
  const unsigned long long seq = xchg_publish(x, ch, (const unsigned int*)in, 2 * n);
  for (int i = threadIdx.x; i < n * x.world; i += blockDim.x) {
    const int r = i / n, e = i % n;
    out[i] = ld_relaxed_sys_u64((const unsigned long long*)xchg_peer_slot(x, ch, seq, r) + e);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct Xchg {
  XchgDev dev{};                 // passed by value to kernels
  int world = 1, rank = 0;
  void* mapped[XCHG_MAX_WORLD] = {};
  bool open = false;
};

struct XchgLocal { unsigned char* window = nullptr; unsigned long long* next_seq = nullptr; };
inline XchgLocal& xchg_local(int device) { static XchgLocal w[64]; return w[device & 63]; }

// allocate (once per device and process) this rank's window and export its IPC handle
inline void xchg_export(int device, void* handle_out) {
  XchgLocal& L = xchg_local(device);
  if (!L.window) {
    if (cudaMalloc((void**)&L.window, XCHG_WINDOW_BYTES) != cudaSuccess) throw std::runtime_error("exchange window allocation failed");
    cudaMemset(L.window, 0, XCHG_WINDOW_BYTES);
    cudaMalloc((void**)&L.next_seq, XCHG_CHANNELS * sizeof(unsigned long long));
    cudaMemset(L.next_seq, 0, XCHG_CHANNELS * sizeof(unsigned long long));
    cudaDeviceSynchronize();
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, L.window) != cudaSuccess) throw std::runtime_error("cudaIpcGetMemHandle failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof h);
}
inline int xchg_release(int device) {
  XchgLocal& L = xchg_local(device);
  if (L.window) { cudaSetDevice(device); cudaFree(L.window); cudaFree(L.next_seq); L.window = nullptr; L.next_seq = nullptr; }
  return 0;
}
inline void xchg_open(Xchg& X, int device, int rank, int world, const void* handles) {
  if (world > XCHG_MAX_WORLD) throw std::runtime_error("world size above XCHG_MAX_WORLD");
  if (!handles) throw std::runtime_error("peer_handles missing for a multi-GPU fit");
  XchgLocal& L = xchg_local(device);
  if (!L.window) throw std::runtime_error("sqc_xchg_export must be called before a multi-GPU session is created");
  X.world = world; X.rank = rank;
  X.dev.world = world; X.dev.rank = rank; X.dev.next_seq = L.next_seq;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { X.dev.peer[r] = L.window; continue; }
    cudaIpcMemHandle_t h; memcpy(&h, (const char*)handles + (size_t)r * 64, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e));
    X.mapped[r] = p; X.dev.peer[r] = (unsigned char*)p;
  }
  X.open = true;
}
inline void xchg_close(Xchg& X) {
  if (!X.open) return;
  for (int r = 0; r < X.world; ++r) if (X.mapped[r]) { cudaIpcCloseMemHandle(X.mapped[r]); X.mapped[r] = nullptr; }
  X.open = false;
}

}  // namespace sqc
