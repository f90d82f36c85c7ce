#pragma once
#include "common.cuh"
namespace psn {
// RAII: records a CUDA event pair around the kernels launched in its scope when profiling is enabled.
struct ProfScope {
  ProfScope(int tag, long long rows, cudaStream_t st);
  ~ProfScope();
  int idx_;
  cudaStream_t st_;
};
}  // namespace psn
