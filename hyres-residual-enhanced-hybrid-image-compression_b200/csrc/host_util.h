// Host-side helpers shared by the C-ABI translation units: error recording and
// CUDA status checks. No exception leaves the library; failures become negative
// HYRES_ERR_* codes plus a message retrievable through hyres_last_error().
#pragma once
#include <cuda_runtime.h>

int hy_fail(int code, const char* msg);
// Every kernel launch of the library is counted (hyres_launch_count): bench.py reports it.
void hy_count_launch();

#define HY_CUDA(expr)                                                   \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) return hy_fail(-2, cudaGetErrorString(_e));  \
  } while (0)
