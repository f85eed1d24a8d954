// Experiment: tcgen05.mma with the A operand in tensor memory (written by the epilogue warps with tcgen05.st).
//   D[m][n] = sum_k A[m][k] * B[n][k],  A [128][64] bf16 packed two K elements per 32-bit TMEM column (lane = row m),
//   B [N][64] K-major SWIZZLE_128B in shared memory.  Per K = 16 MMA the A address advances by 8 columns.
//   order = 0: element 2j in the low half of column j; order = 1: in the high half.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I <csrc> -o umma_ts umma_ts.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(128) k(const __nv_bfloat16* a_lin, const __nv_bfloat16* b_lin, float* out, int N, int order) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - hy::smem_u32(raw));
  const uint32_t b_s = base, bar = b_s + 128 * 128, slot = bar + 8;
  for (int i = threadIdx.x; i < N * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(b_lin)[r * 8 + c];
  }
  hy::fence_async_smem();
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (threadIdx.x < 32) { hy::tmem_alloc(slot, 256); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const int warp = threadIdx.x >> 5;
  const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t colA = 192;
  {
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a_lin + threadIdx.x * 64);
    uint32_t v[16];
    for (int h = 0; h < 2; ++h) {
      for (int j = 0; j < 16; ++j) {
        uint32_t u = arow[h * 16 + j];
        if (order) u = (u >> 16) | (u << 16);
        v[j] = u;
      }
      tmem_st16(t_lane + colA + h * 16, v);
    }
    tmem_st_wait();
  }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = hy::umma_idesc_bf16(128, N);
    for (int kk = 0; kk < 4; ++kk) umma_ts_bf16(tmem, tmem + colA + kk * 8, hy::umma_desc_sw128(b_s + kk * 32), idesc, kk ? 1u : 0u);
    hy::umma_commit(bar);
  }
  hy::mbar_wait(bar, 0);
  hy::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    hy::tmem_ld16(t_lane + c0, r);
    hy::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[threadIdx.x * N + c0 + i] = __uint_as_float(r[i]);
  }
  hy::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 256); }
}

int main() {
  const int N = 128;
  std::vector<__nv_bfloat16> a(128 * 64), b(N * 64);
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) a[r * 64 + c] = __float2bfloat16(static_cast<float>((r * 7 + c * 3) % 13) - 6.f);
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < 64; ++c) b[n * 64 + c] = __float2bfloat16(static_cast<float>((n * 5 + c * 11) % 17) - 8.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 128 * 128 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(128 * N);
  for (int order = 0; order < 2; ++order) {
    k<<<1, 128, smem>>>(da, db, dout, N, order);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("order=%d CUDA error %s\n", order, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int c = 0; c < 64; ++c) ref += __bfloat162float(a[m * 64 + c]) * __bfloat162float(b[n * 64 + c]);
        if (ref != out[m * N + n]) ++bad;
      }
    printf("order=%d mismatches=%d of %d\n", order, bad, 128 * N);
  }
  return 0;
}
