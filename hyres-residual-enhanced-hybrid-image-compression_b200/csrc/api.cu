// Library-level entry points of the C-ABI: version, device probe, last error.
#include <cstdio>
#include <cstring>

#include "host_util.h"
#include "hyres_b200.h"

#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void hy_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int hy_fail(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
  return code;
}

extern "C" {

int hyres_version(void) { return 100; }

long long hyres_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* hyres_last_error(void) { return g_err; }

// The product path has no CPU fallback: anything but an sm_100 device is an error.
int hyres_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) return hy_fail(HYRES_ERR_CUDA, "no CUDA device visible");
  if (device < 0 || device >= count) return hy_fail(HYRES_ERR_ARG, "device index out of range");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) {
    char msg[96];
    snprintf(msg, sizeof msg, "device is sm_%d%d; this library is sm_100a only", major, minor);
    return hy_fail(HYRES_ERR_UNSUPPORTED, msg);
  }
  return HYRES_OK;
}

}  // extern "C"
