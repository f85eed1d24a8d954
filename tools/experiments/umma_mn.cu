// Experiment: MN-major B operand for tcgen05.mma (kind::f16, SWIZZLE_128B).
//   D[m][n] = sum_k U[m][k] * T[k][n],  A = U [128][64] K-major, B = T stored as [k rows][64 n] (128 B rows, 16 B
//   chunks XOR-swizzled by row & 7 -- exactly what a TMA box (64 channels, pixels...) writes), i.e. N is the
//   contiguous dimension of B.  Instruction descriptor bit 16 = "B is MN-major"; per MMA (K = 16) the B start
//   address advances by 16 rows = 2048 B.  Which of LBO / SBO carries the 8-row-group stride (1024 B)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I <csrc> -o umma_mn umma_mn.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* a_lin, const __nv_bfloat16* b_lin, float* out,
                                        int lbo, int sbo, int nk) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - hy::smem_u32(raw));
  const uint32_t a_s = base, b_s = base + 128 * 128, bar = b_s + 64 * 128, slot = bar + 8;
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(a_lin)[r * 8 + c];
  }
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + 128 * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(b_lin)[r * 8 + c];
  }
  hy::fence_async_smem();
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (threadIdx.x < 32) { hy::tmem_alloc(slot, 64); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc = hy::umma_idesc_bf16(128, 64) | (1u << 16);  // B MN-major
    for (int kk = 0; kk < nk; ++kk) {
      uint64_t bd = hy::umma_desc_sw128(b_s + kk * 2048, sbo);
      bd = (bd & ~(static_cast<uint64_t>(0x3fff) << 16)) | (static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16);
      hy::umma_bf16(tmem, hy::umma_desc_sw128(a_s + kk * 32), bd, idesc, kk ? 1u : 0u);
    }
    hy::umma_commit(bar);
  }
  hy::mbar_wait(bar, 0);
  hy::tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    hy::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    hy::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[threadIdx.x * 64 + c0 + i] = __uint_as_float(r[i]);
  }
  hy::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 64); }
}

int main() {
  std::vector<__nv_bfloat16> a(128 * 64), b(64 * 64);
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) a[r * 64 + c] = __float2bfloat16(static_cast<float>((r * 7 + c * 3) % 13) - 6.f);
  for (int kx = 0; kx < 64; ++kx)
    for (int n = 0; n < 64; ++n) b[kx * 64 + n] = __float2bfloat16(static_cast<float>((kx * 5 + n * 11) % 17) - 8.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 128 * 128 + 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * 64);
  const int cand[][2] = {{16, 1024}, {1024, 1024}, {1024, 16}, {2048, 1024}, {1024, 2048}, {128, 1024}, {1024, 128}, {8192, 1024}};
  for (int nk = 2; nk <= 4; nk += 2)
  for (auto& c : cand) {
    k<<<1, 128, smem>>>(da, db, dout, c[0], c[1], nk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("lbo=%d sbo=%d CUDA error %s\n", c[0], c[1], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        float ref = 0;
        for (int kx = 0; kx < nk * 16; ++kx) ref += __bfloat162float(a[m * 64 + kx]) * __bfloat162float(b[kx * 64 + n]);
        if (ref != out[m * 64 + n]) ++bad;
      }
    printf("K=%d lbo=%5d sbo=%5d mismatches=%d\n", nk * 16, c[0], c[1], bad);
  }
  return 0;
}
