// sqc_xchg.cuh — cross-GPU exchange through peer memory (NVLink 5 / NVSwitch), one process per GPU.
//
// Cells are sharded by sample: alpha_j, p_jk and the labels never leave their GPU; what
// the ranks must agree on are a few small vectors per step (histograms of the exact selection, candidate
// lists, K x (1 + D + D^2) moments, log-likelihood sums).  Those are exchanged INSIDE the kernels that
// produce them.  Every rank owns a window in its HBM (cudaMalloc, exported with cudaIpcGetMemHandle and mapped
// by all peers).  Messages are <= 80 KB, so what matters is latency, not the 900 GB/s links; the protocol is
// therefore a PUSH of self-validating lines:
//   * a line is 16 bytes = two 64-bit words, each {32 data bits, 32-bit sequence number of the exchange};
//   * the sender stores its lines straight into every peer's window (st.relaxed.sys.v2.u64, fire and forget,
//     one NVLink crossing, no fence and no separate flag);
//   * the receiver polls its OWN memory until both halves of a line carry the expected sequence number, so a
//     line is complete the moment it is seen, whatever the order the stores arrive in.
// All ranks combine the copies in RANK ORDER, so they hold bit-identical results and take identical
// control-flow decisions without any host round trip or NCCL launch.  Lines live in (channel, sequence parity,
// source rank) slots: a rank sends exchange q + 2 only after it has received everybody's q + 1, which they sent
// after consuming q — so a slot is never overwritten while its owner may still read it.
#pragma once
#include <cstdint>
#include <cstring>
#include <array>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "sqc_common.cuh"

namespace sqc {

constexpr int XCHG_MAX_WORLD = 8;
constexpr int XCHG_CHANNELS = SQC_MAXK + 4;          // one per cluster leader + host-stream level channels
constexpr int XCHG_CH_HOST = SQC_MAXK;               // first host-stream channel
constexpr int XCHG_LINE_CAP = 5632;                  // lines per slot: >= 1 + 45 + 512 * (2 + 8) (tie message of k_mcd at D = 8)
constexpr size_t XCHG_WINDOW_BYTES = (size_t)XCHG_CHANNELS * 2 * XCHG_MAX_WORLD * XCHG_LINE_CAP * 16;

struct XchgDev {
  int world, rank;
  unsigned char* peer[XCHG_MAX_WORLD];               // window base of every rank (own window: local pointer)
  unsigned long long* next_seq;                      // XCHG_CHANNELS counters in LOCAL memory (same value on all ranks)
  int* status;                                       // the session's device status word: a wait that runs out of its budget says so here
  long long budget_ns;                               // time budget of every in-kernel wait (SQC_WAIT_BUDGET_NS; SQC_WAIT_BUDGET_MS in the environment overrides)
};

// slot of `src`'s message of exchange `seq` on channel `ch`, inside the window at `base`
__device__ __forceinline__ ulonglong2* xw_line(unsigned char* base, int ch, unsigned long long seq, int src) {
  return (ulonglong2*)base + (((size_t)ch * 2 + (size_t)(seq & 1ull)) * XCHG_MAX_WORLD + (size_t)src) * XCHG_LINE_CAP;
}
__device__ __forceinline__ void st_line(ulonglong2* p, unsigned int w0, unsigned int w1, unsigned int f) {
  const unsigned long long a = (unsigned long long)w0 | ((unsigned long long)f << 32), b = (unsigned long long)w1 | ((unsigned long long)f << 32);
  asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_line(const ulonglong2* p) {
  ulonglong2 v; asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ bool line_ready(const ulonglong2& v, unsigned int f) { return (unsigned int)(v.x >> 32) == f && (unsigned int)(v.y >> 32) == f; }
__device__ __forceinline__ double line_f64(const ulonglong2& v) { return __hiloint2double((int)(unsigned int)v.y, (int)(unsigned int)v.x); }

// Block-wide: number of the next exchange on channel ch (identical on every rank: all ranks run the same exchanges)
__device__ inline unsigned long long xchg_begin(const XchgDev& x, int ch) {
  __shared__ unsigned long long sh_seq;
  if (threadIdx.x == 0) sh_seq = ++x.next_seq[ch];
  __syncthreads();
  return sh_seq;
}
// one line to every rank (own window included: the receive path is the same for all sources)
__device__ __forceinline__ void xchg_put(const XchgDev& x, int ch, unsigned long long seq, int line, unsigned int w0, unsigned int w1) {
#pragma unroll
  for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) st_line(xw_line(x.peer[r], ch, seq, x.rank) + line, w0, w1, (unsigned int)seq);
}
__device__ __forceinline__ void xchg_put_f64(const XchgDev& x, int ch, unsigned long long seq, int line, double v) {
  xchg_put(x, ch, seq, line, (unsigned int)__double2loint(v), (unsigned int)__double2hiint(v));
}
// wait for line `line` of rank r's message: local polling, tight at first, then with a 1 us back-off.  A peer that never
// sends: after SQC_WAIT_BUDGET_NS the wait gives up, writes SQC_D_TIMEOUT into the session's status word (every later
// wait then returns at once) and the fit ends with SQC_ERR_COMM — no trap, the CUDA context survives.
__device__ __forceinline__ ulonglong2 xchg_wait(const XchgDev& x, int ch, unsigned long long seq, int r, int line) {
  const ulonglong2* p = xw_line(x.peer[x.rank], ch, seq, r) + line;
  ulonglong2 v = ld_line(p);
  unsigned int spins = 0; long long t0 = 0;
  while (!line_ready(v, (unsigned int)seq)) {
    if ((++spins & 0xffu) == 0u) {
      if (*(volatile int*)x.status == SQC_D_TIMEOUT) break;      // somebody has given up already: every wait returns at once
      if (spins > (1u << 16)) {
        __nanosleep(1000);
        const long long now = gtimer();
        if (!t0) t0 = now;
        else if (now - t0 > x.budget_ns) { atomicCAS(x.status, 0, SQC_D_TIMEOUT); break; }
      }
    }
    v = ld_line(p);
  }
  return v;
}
// the same line of every rank: all loads are issued before the first is examined
__device__ __forceinline__ void xchg_wait_all(const XchgDev& x, int ch, unsigned long long seq, int line, ulonglong2 (&v)[XCHG_MAX_WORLD]) {
#pragma unroll
  for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) v[r] = ld_line(xw_line(x.peer[x.rank], ch, seq, r) + line);
#pragma unroll
  for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world && !line_ready(v[r], (unsigned int)seq)) v[r] = xchg_wait(x, ch, seq, r, line);
}

// sum of 32-bit counters over ranks, in place (block-wide).  A thread sends and later overwrites the same words.
__device__ __noinline__ void xchg_sum_u32(const XchgDev& x, int ch, unsigned int* buf, int n) {
  const unsigned long long seq = xchg_begin(x, ch);
  const int nl = (n + 1) >> 1;
  for (int i = threadIdx.x; i < nl; i += blockDim.x) xchg_put(x, ch, seq, i, buf[2 * i], (2 * i + 1 < n) ? buf[2 * i + 1] : 0u);
  for (int i = threadIdx.x; i < nl; i += blockDim.x) {
    ulonglong2 v[XCHG_MAX_WORLD];
    xchg_wait_all(x, ch, seq, i, v);
    unsigned int a0 = 0u, a1 = 0u;
#pragma unroll
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) { a0 += (unsigned int)v[r].x; a1 += (unsigned int)v[r].y; }
    buf[2 * i] = a0;
    if (2 * i + 1 < n) buf[2 * i + 1] = a1;
  }
  __syncthreads();
}
// sum of doubles over ranks in rank order (bit-identical on every rank), in place (block-wide); one double per line
__device__ __noinline__ void xchg_sum_f64(const XchgDev& x, int ch, double* buf, int n) {
  const unsigned long long seq = xchg_begin(x, ch);
  for (int i = threadIdx.x; i < n; i += blockDim.x) xchg_put_f64(x, ch, seq, i, buf[i]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    ulonglong2 v[XCHG_MAX_WORLD];
    xchg_wait_all(x, ch, seq, i, v);
    double a = 0.0;
#pragma unroll
    for (int r = 0; r < XCHG_MAX_WORLD; ++r) if (r < x.world) a += line_f64(v[r]);
    buf[i] = a;
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
  const unsigned long long seq = xchg_begin(x, ch);
  for (int i = threadIdx.x; i < n; i += blockDim.x) xchg_put(x, ch, seq, i, (unsigned int)in[i], (unsigned int)(in[i] >> 32));
  for (int i = threadIdx.x; i < n * x.world; i += blockDim.x) {
    const int r = i / n, e = i % n;
    const ulonglong2 v = xchg_wait(x, ch, seq, r, e);
    out[i] = (unsigned long long)(unsigned int)v.x | ((unsigned long long)(unsigned int)v.y << 32);
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

// allocate (once per device and process) this rank's window; reset = clear the window and the sequence counters.
// Every multi-GPU session starts from a cleared window and sequence number 0 on every rank (a rank that failed or was
// restarted therefore cannot leave the counters of the ranks out of step, and stale lines of earlier fits cannot
// match): each rank clears its own window before the ranks meet (in-process: a host barrier; one process per GPU:
// the all-gather of the handles that follows sqc_xchg_export).
inline void xchg_ensure(int device, bool reset) {
  XchgLocal& L = xchg_local(device);
  if (!L.window) {
    if (cudaMalloc((void**)&L.window, XCHG_WINDOW_BYTES) != cudaSuccess) throw std::runtime_error("exchange window allocation failed");
    if (cudaMalloc((void**)&L.next_seq, XCHG_CHANNELS * sizeof(unsigned long long)) != cudaSuccess) throw std::runtime_error("exchange counter allocation failed");
    reset = true;
  }
  if (reset) {
    cudaMemset(L.window, 0, XCHG_WINDOW_BYTES);
    cudaMemset(L.next_seq, 0, XCHG_CHANNELS * sizeof(unsigned long long));
    if (cudaDeviceSynchronize() != cudaSuccess) throw std::runtime_error("exchange window reset failed");
  }
}
inline void xchg_export(int device, void* handle_out) {
  xchg_ensure(device, true);
  XchgLocal& L = xchg_local(device);
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, L.window) != cudaSuccess) throw std::runtime_error("cudaIpcGetMemHandle failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, sizeof h);
}
// Peer mappings are kept for the life of the process, keyed by the 64 bytes of the IPC handle (sqc_comm semantics):
// opening and closing 7 windows and 7 key buffers per fit cost more than the fit itself at 8 GPUs.  A handle names one
// allocation of one peer process: a peer that restarts or re-allocates presents a new handle and gets a new mapping; the
// stale one is never used again and is closed by sqc_xchg_release(device).
struct IpcCache {
  std::mutex mu; std::map<std::array<unsigned char, 64>, void*> map[64];
  void* open(int device, const void* handle64) {
    std::array<unsigned char, 64> key; memcpy(key.data(), handle64, 64);
    std::lock_guard<std::mutex> lk(mu);
    auto& m = map[device & 63];
    auto it = m.find(key);
    if (it != m.end()) return it->second;
    cudaIpcMemHandle_t h; memcpy(&h, handle64, sizeof h);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); throw std::runtime_error(std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e)); }
    m.emplace(key, p);
    return p;
  }
  void close_all(int device) {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : map[device & 63]) cudaIpcCloseMemHandle(kv.second);
    map[device & 63].clear();
  }
};
inline IpcCache& ipc_cache() { static IpcCache c; return c; }

inline int xchg_release(int device) {
  cudaSetDevice(device);
  ipc_cache().close_all(device);
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
  X.open = true;                                       // from here on xchg_close releases what has been mapped (also after a failure half way)
  for (int r = 0; r < world; ++r) {
    if (r == rank) { X.dev.peer[r] = L.window; continue; }
    X.dev.peer[r] = (unsigned char*)ipc_cache().open(device, (const char*)handles + (size_t)r * 64);   // mapped once per process and peer
  }
}
// one process, one host thread per GPU: the peers' windows are ordinary device pointers (peer access enabled by the caller)
inline void xchg_open_inproc(Xchg& X, int device, int rank, int world, const int* devices) {
  if (world > XCHG_MAX_WORLD) throw std::runtime_error("world size above XCHG_MAX_WORLD");
  X.world = world; X.rank = rank;
  X.dev.world = world; X.dev.rank = rank; X.dev.next_seq = xchg_local(device).next_seq;
  for (int r = 0; r < world; ++r) {
    XchgLocal& L = xchg_local(devices[r]);
    if (!L.window) throw std::runtime_error("exchange window of a peer is missing");
    X.dev.peer[r] = L.window;
  }
  X.open = true;                                       // nothing mapped: xchg_close has nothing to release
}
inline void xchg_close(Xchg& X) {
  if (!X.open) return;
  for (int r = 0; r < X.world; ++r) X.mapped[r] = nullptr;       // (the mappings belong to the process-wide cache)
  X.open = false;
}

}  // namespace sqc
