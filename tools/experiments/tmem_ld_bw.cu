// Microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM with 4 / 8 / 16 warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../hyres-residual-enhanced-hybrid-image-compression_b200/csrc/common.cuh"

template <int X>
__global__ void k(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { hy::tmem_alloc(hy::smem_u32(&slot), 512); hy::tmem_relinquish(); }
  hy::tc_fence_before(); __syncthreads(); hy::tc_fence_after();
  const uint32_t tb = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (X == 32) {
      uint32_t r[32];
      hy::tmem_ld32(tb + ((i * 32) & 255), r);
      hy::tmem_ld_fence32(r);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    } else {
      uint32_t r[16];
      hy::tmem_ld16(tb + ((i * 16) & 255), r);
      hy::tmem_ld_fence(r);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  hy::tc_fence_before(); __syncthreads();
  if (warp == 0) { hy::tc_fence_after(); hy::tmem_dealloc(slot, 512); }
}

int main() {
  long long* d; uint32_t* s;
  cudaMalloc(&d, 148 * 8); cudaMalloc(&s, 148 * 1024 * 4);
  const int iters = 2000;
  for (int warps : {1, 4, 8, 16}) {
    for (int x : {16, 32}) {
      if (x == 32) k<32><<<148, warps * 32>>>(iters, d, s); else k<16><<<148, warps * 32>>>(iters, d, s);
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      double clk = (double)h[0];
      double bytes = (double)iters * warps * 32 * x * 4;
      printf("warps %2d x%d: %.0f clk, %.1f B/clk/SM, %.1f clk per ld per warp  (%s)\n", warps, x, clk, bytes / clk, clk / iters,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
