// common.cuh — error reporting and launch accounting shared by the translation units of libgo2b200.so
#pragma once
#include <cuda_runtime.h>

namespace go2 {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t err, const char* file, int line);
void count_launch();
}  // namespace go2

#define GO2_CUDA_OK(expr)                                                      \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return go2::set_cuda_error(_e, __FILE__, __LINE__); \
  } while (0)
