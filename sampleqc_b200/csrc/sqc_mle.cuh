// sqc_mle.cuh — soft-EM (fit_sampleqc_mle_cpp) passes.  STUB.
#pragma once
#include "sqc_kernels.cuh"
namespace sqc {
struct MleBufs { double* gamma = nullptr; };
template <int D> void mle_launch_pass(const Dev&, const MleBufs&, int, cudaStream_t) {}
}  // namespace sqc
