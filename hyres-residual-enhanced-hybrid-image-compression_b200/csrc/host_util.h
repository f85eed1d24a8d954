// Host-side helpers shared by the C-ABI translation units: error recording and
// CUDA status checks. No exception leaves the library; failures become negative
// HYRES_ERR_* codes plus a message retrievable through hyres_last_error().
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>

int hy_fail(int code, const char* msg);
// Every kernel launch of the library is counted (hyres_launch_count): bench.py reports it.
void hy_count_launch();

// nsplit codes of the split-precision layers (hyres_b200.h): 1..3 bf16 parts, or two half parts (2 | 16)
inline bool hy_split_code_ok(int code) { return (code >= 1 && code <= 3) || code == (2 | 16); }

#define HY_CUDA(expr)                                                   \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) return hy_fail(-2, cudaGetErrorString(_e));  \
  } while (0)

// One-time initialisation per DEVICE (function attributes such as the dynamic shared-memory opt-in are per device,
// and one process may drive several GPUs): `done` says whether the current device has been initialised, `mark`
// records it after the initialisation succeeded (two threads racing both initialise; that is harmless).
struct HyPerDevice {
  std::atomic<uint64_t> mask{0};
  static uint64_t bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
  }
  bool done() const { return (mask.load(std::memory_order_acquire) & bit()) != 0; }
  void mark() { mask.fetch_or(bit(), std::memory_order_release); }
};

// Launch with the programmatic-stream-serialization attribute (see pdl_wait in common.cuh); HYRES_NO_PDL=1
// falls back to plain stream order.
template <typename P>
inline cudaError_t hy_launch_pdl(void (*kernel)(P), int grid, int block, int smem, cudaStream_t stream, const P& params) {
  static const bool no_pdl = getenv("HYRES_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, params);
}
