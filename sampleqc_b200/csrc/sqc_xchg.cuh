// sqc_xchg.cuh — cross-GPU exchange over peer memory (NVLink / NVSwitch).  STUB: single GPU only for now.
#pragma once
#include <cuda_runtime.h>
struct sqc_session;
namespace sqc {
struct XchgDev { int world, rank; };
struct Xchg { XchgDev* dev = nullptr; int world = 1, rank = 0; };
inline void xchg_open(Xchg&, int, int, int, const void*) { throw std::runtime_error("multi-GPU exchange not built"); }
inline void xchg_close(Xchg&) {}
inline void xchg_allreduce_host(Xchg&, cudaStream_t, double*, int) {}
inline void xchg_cluster_layout(sqc_session*, Dev&) {}
inline void xchg_allreduce_like(sqc_session*, Dev&) {}
struct MleBufs;
inline void xchg_allreduce_mle(sqc_session*, Dev&, MleBufs&, int, double*, int) {}
inline void xchg_export(int, void*) { throw std::runtime_error("multi-GPU exchange not built"); }
inline int xchg_release(int) { return 0; }
}  // namespace sqc
