// Experiment: can a SWIZZLE_128B K-major UMMA A-operand descriptor start at an arbitrary
// 128-byte row (not 1024 B aligned) of a buffer written with absolute-address swizzle?
// Tries base_offset = 0 and base_offset = (addr >> 7) & 7 for row shifts 0..17.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I <csrc> -o umma_shift umma_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

constexpr int kRows = 320;  // rows in the smem A buffer (64 bf16 = 128 B each)

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* a_lin, const __nv_bfloat16* b_lin, float* out,
                                        int shift, int use_base_off, int sbo_rows) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - hy::smem_u32(raw));
  const uint32_t a_s = base, b_s = base + kRows * 128, bar = b_s + 64 * 128, slot = bar + 8;
  // write A [kRows][64] and B [64][64] with absolute-address 128B swizzle (16B chunk ^= row&7)
  for (int i = threadIdx.x; i < kRows * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(a_lin)[r * 8 + c];
    *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    const uint4 v = reinterpret_cast<const uint4*>(b_lin)[r * 8 + c];
    *reinterpret_cast<uint4*>(gen + kRows * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
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
    const uint32_t idesc = hy::umma_idesc_bf16(128, 64);
    const uint32_t a0 = a_s + shift * 128;
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t ad = hy::umma_desc_sw128(a0 + kk * 32);
      // 8-row groups sbo_rows*128 B apart (8 = dense; 10 = a patch that is 10 positions wide)
      ad = (ad & ~(static_cast<uint64_t>(0x3fff) << 32)) | (static_cast<uint64_t>((sbo_rows * 128) >> 4) << 32);
      if (use_base_off) ad |= static_cast<uint64_t>((a0 >> 7) & 7) << 49;
      hy::umma_bf16(tmem, ad, hy::umma_desc_sw128(b_s + kk * 32), idesc, kk ? 1u : 0u);
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
  std::vector<__nv_bfloat16> a(kRows * 64), b(64 * 64);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < 64; ++c) a[r * 64 + c] = __float2bfloat16(static_cast<float>((r * 7 + c * 3) % 61) - 30.f);
  for (int n = 0; n < 64; ++n)
    for (int c = 0; c < 64; ++c) b[n * 64 + c] = __float2bfloat16(static_cast<float>((n * 5 + c * 11) % 17) - 8.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  const int smem = kRows * 128 + 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * 64);
  for (int sbo = 8; sbo <= 10; sbo += 2)
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int shift = 0; shift < 23; ++shift) {
      if (sbo == 10 && ubo == 1) continue;
      k<<<1, 128, smem>>>(da, db, dout, shift, ubo, sbo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("base_off=%d shift=%d CUDA error %s\n", ubo, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          const int row = shift + (m >> 3) * sbo + (m & 7);
          for (int c = 0; c < 64; ++c) ref += __bfloat162float(a[row * 64 + c]) * __bfloat162float(b[n * 64 + c]);
          if (ref != out[m * 64 + n]) ++bad;
        }
      printf("sbo_rows=%d base_off=%d shift=%2d mismatches=%d\n", sbo, ubo, shift, bad);
    }
  return 0;
}
