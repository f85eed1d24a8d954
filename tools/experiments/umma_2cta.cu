// Experiment: tcgen05.mma.cta_group::2 (M = 256 over an SM pair) -- operand split, TMEM placement, multicast
// commit, and the sustained rate against cta_group::1.
//   * each CTA of the pair holds its own 128 x 64 A tile and N/2 rows of the B tile (same shared-memory offsets);
//   * the leader CTA (cluster rank 0) issues; D rows 0..127 land in CTA 0's TMEM, rows 128..255 in CTA 1's;
//   * tcgen05.commit ... multicast::cluster signals the barrier at the same offset in both CTAs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I <csrc> -o umma_2cta umma_2cta.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"

namespace cg = cooperative_groups;

__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// A: [2 CTAs][128][64] bf16 row-major, B: [N][64] (rows 0..N/2-1 go to CTA 0, the rest to CTA 1), out: [256][N]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k2(const __nv_bfloat16* a_lin, const __nv_bfloat16* b_lin, float* out, int N, int iters, long long* clk) {
  extern __shared__ uint8_t raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - hy::smem_u32(raw));
  const uint32_t a_s = base, b_s = base + 16384, bar = b_s + 16384, slot = bar + 8;
  const int half = N / 2;
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(a_lin)[(rank * 128 + r) * 8 + c];
  }
  for (int i = threadIdx.x; i < half * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(gen + 16384 + r * 128 + ((c ^ (r & 7)) << 4)) = reinterpret_cast<const uint4*>(b_lin)[(rank * half + r) * 8 + c];
  }
  hy::fence_async_smem();
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(slot, 256); tmem_relinquish2(); }
  hy::tc_fence_before();
  __syncthreads();
  cluster.sync();  // both CTAs' operands and barriers are ready
  hy::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = hy::umma_idesc_bf16(256, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int kk = 0; kk < 4; ++kk)
        umma2_bf16(tmem, hy::umma_desc_sw128(a_s + kk * 32), hy::umma_desc_sw128(b_s + kk * 32), idesc, (it | kk) ? 1u : 0u);
    umma2_commit_mc(bar, 0x3);
    hy::mbar_wait(bar, 0);
    clk[0] = clock64() - t0;
  }
  hy::mbar_wait(bar, 0);
  hy::tc_fence_after();
  const int warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    hy::tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    hy::tmem_ld_fence(r);
    for (int i = 0; i < 16; ++i) out[(rank * 128 + threadIdx.x) * N + c0 + i] = __uint_as_float(r[i]);
  }
  hy::tc_fence_before();
  __syncthreads();
  cluster.sync();
  if (threadIdx.x < 32) { hy::tc_fence_after(); tmem_dealloc2(tmem, 256); }
}

int main() {
  std::vector<__nv_bfloat16> a(256 * 64), b(256 * 64);
  for (int r = 0; r < 256; ++r)
    for (int c = 0; c < 64; ++c) a[r * 64 + c] = __float2bfloat16(static_cast<float>((r * 7 + c * 3) % 61) - 30.f);
  for (int n = 0; n < 256; ++n)
    for (int c = 0; c < 64; ++c) b[n * 64 + c] = __float2bfloat16(static_cast<float>((n * 5 + c * 11) % 17) - 8.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  long long* dclk;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dout, 256 * 256 * 4); cudaMalloc(&dclk, 8);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 16384 + 16384 + 64 + 1024;
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> out(256 * 256);
  for (int N : {64, 128, 256}) {
    k2<<<2, 128, smem>>>(da, db, dout, N, 1, dclk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dout, 256 * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 256; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int c = 0; c < 64; ++c) ref += __bfloat162float(a[m * 64 + c]) * __bfloat162float(b[n * 64 + c]);
        if (ref != out[m * N + n]) ++bad;
      }
    printf("cta_group::2 M=256 N=%3d K=64: mismatches=%d of %d\n", N, bad, 256 * N);
    // rate: 74 clusters, 256 x 4 MMAs each
    k2<<<148, 128, smem>>>(da, db, dout, N, 256, dclk);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d rate run: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    long long c;
    cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
    const double per = static_cast<double>(c) / 1024.0;
    printf("  rate: %.1f clk per 256 x %d x 16 MMA -> %.0f MAC/clk/SM (%.2f of 4096)\n", per, N, 256.0 * N * 16 / per / 2,
           256.0 * N * 16 / per / 2 / 4096);
  }
  return 0;
}
