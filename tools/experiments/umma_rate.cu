// Experiment: sustained tcgen05.mma rate (cta_group::1, kind::f16, SS operands, SWIZZLE_128B K-major) as a
// function of N, with A tiles walked the way the convolution kernels walk them (a new 16 KB tile every 4 MMAs, or
// one shifted patch).  One CTA per SM on every SM, `iters` groups of 4 MMAs (K = 64) issued back to back by one
// thread, one commit at the end; reports clocks per 128 x N x 16 MMA and the implied fraction of 4096 MAC/clk/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I <csrc> -o umma_rate umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include "common.cuh"

__global__ void __launch_bounds__(128) k(long long* out, int N, int iters, int a_tiles, int a_step_bytes, int sbo,
                                        int ncommit) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 8 * 16384, bar = b_s + 2 * 32768, slot = bar + 8;
  for (int i = threadIdx.x; i < (8 * 16384 + 2 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - hy::smem_u32(raw)))[i] = 0x3c003c00u;
  hy::fence_async_smem();
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (threadIdx.x < 32) { hy::tmem_alloc(slot, 512); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc = hy::umma_idesc_bf16(128, N);
    const uint32_t hi_a = hy::desc_hi_sw128(sbo), hi_b = hy::desc_hi_sw128();
    const long long t0 = clock64();
    uint32_t par = 0;
    for (int i = 0; i < iters; ++i) {
      const uint32_t a_lo = hy::desc_lo(a_s + (i % a_tiles) * a_step_bytes);
      const uint32_t b_lo = hy::desc_lo(b_s + (ncommit == -3 ? 0 : (i & 1) * 32768));
      const uint32_t d = ncommit == -2 ? tmem + (i & 3) * 128 : tmem + ((i & 1) ? 256 : 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        hy::umma_bf16(d, hy::desc_pack(a_lo + 2 * kk, hi_a), hy::desc_pack(b_lo + 2 * kk, hi_b), idesc,
                      ncommit == -1 ? (kk ? 1u : 0u) : 1u);
      if (ncommit > 0 && (i % ncommit) == ncommit - 1) {
        hy::umma_commit(bar);
        hy::mbar_wait(bar, par);
        par ^= 1u;
      }
    }
    const long long t1 = clock64();
    hy::umma_commit(bar);
    hy::mbar_wait(bar, par);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  hy::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 512); }
}

int main() {
  if (getenv("SKIP_MAIN")) return 0;
  long long* d;
  cudaMalloc(&d, 148 * 2 * 8);
  const int smem = 8 * 16384 + 2 * 32768 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<long long> h(148 * 2);
  const int iters = 2000;
  struct Cfg { const char* name; int a_tiles, a_step, sbo, ncommit; };
  const Cfg cfgs[] = {
      {"same A tile", 1, 0, 1024, 0},
      {"8 A tiles round robin", 8, 16384, 1024, 0},
      {"3x3 patch taps (PW=10, shifted rows)", 9, 128, 1280, 0},
      {"8 A tiles, commit+wait every 9 groups", 8, 16384, 1024, 9},
      {"same A tile, first-of-group overwrites", 1, 0, 1024, -1},
      {"same A tile, D cycles over 4 ranges", 1, 0, 1024, -2},
      {"same A tile, B tile fixed", 1, 0, 1024, -3},
  };
  for (const Cfg& c : cfgs)
    for (int N : {16, 64, 128, 256}) {
      if (c.ncommit == -2 && N > 128) continue;
      for (int grid : {148}) {
        k<<<grid, 128, smem>>>(d, N, iters, c.a_tiles, c.a_step, c.sbo, c.ncommit);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
        long long worst = 0, issue = 0;
        for (int i = 0; i < grid; ++i) { worst = std::max(worst, h[2 * i + 1]); issue = std::max(issue, h[2 * i]); }
        const double per = static_cast<double>(worst) / (iters * 4);
        printf("%-40s N=%3d grid=%3d  clk/MMA %.1f (issue %.1f)  MAC/clk/SM %.0f  frac %.2f\n", c.name, N, grid, per,
               static_cast<double>(issue) / (iters * 4), 128.0 * N * 16 / per, 128.0 * N * 16 / per / 4096);
      }
    }
  return 0;
}

// ---- second experiment: the fused residual-unit kernel's G2 issue pattern, literally ----
// 36 MMAs (9 taps x 4 k-steps, N = 64) on one shifted patch, compile-time offsets, then commit + wait.
__global__ void __launch_bounds__(128) g2(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 32768, bar = b_s + 9 * 8192, slot = bar + 8;
  for (int i = threadIdx.x; i < (32768 + 9 * 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - hy::smem_u32(raw)))[i] = 0x3c003c00u;
  hy::fence_async_smem();
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (threadIdx.x < 32) { hy::tmem_alloc(slot, 512); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    const uint32_t idesc64 = hy::umma_idesc_bf16(128, 64);
    constexpr uint32_t hi = hy::desc_hi_sw128(), hi_t1 = hy::desc_hi_sw128(10 * 128);
    const uint32_t t_lo = hy::desc_lo(a_s), w2_lo = hy::desc_lo(b_s);
    long long issue = 0, total = 0;
    uint32_t par = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            hy::umma_bf16(tmem + 128, hy::desc_pack(t_lo + (((r * 10 + s) * 128 + kk * 32) >> 4), hi_t1),
                          hy::desc_pack(w2_lo + (((s * 3 + r) * 8192 + kk * 32) >> 4), hi), idesc64, (s | r | kk) ? 1u : 0u);
      hy::umma_commit(bar);
      const long long t1 = clock64();
      hy::mbar_wait(bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      issue += t1 - t0;
      total += t2 - t0;
    }
    out[blockIdx.x * 2] = issue / reps;
    out[blockIdx.x * 2 + 1] = total / reps;
  }
  hy::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 512); }
}

struct G2Runner {
  G2Runner() {
    long long* d;
    cudaMalloc(&d, 148 * 16);
    const int smem = 32768 + 9 * 8192 + 1024 + 64;
    cudaFuncSetAttribute(g2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    g2<<<148, 128, smem>>>(d, 200);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("RU G2 pattern (36 MMAs N=64, commit, wait): issue %lld clk, total %lld clk => %.1f clk/MMA (%s)\n", h[0], h[1],
           h[1] / 36.0, cudaGetErrorString(e));
  }
} g2_runner;

// ---- third experiment: converged warp, leader lane decided once (MODE 2) or elect per MMA (MODE 1) ----
template <int MODE, int N, int NMMA>
__global__ void __launch_bounds__(128) conv_issue(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 8 * 16384, bar = b_s + 2 * 32768, slot = bar + 8;
  for (int i = threadIdx.x; i < (8 * 16384 + 2 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - hy::smem_u32(raw)))[i] = 0x3c003c00u;
  hy::fence_async_smem();
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (threadIdx.x < 32) { hy::tmem_alloc(slot, 512); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x < 32) {
    const uint32_t leader = hy::elect_leader();
    const uint32_t idesc = hy::umma_idesc_bf16(128, N);
    constexpr uint32_t hi = hy::desc_hi_sw128();
    const uint32_t a_lo = hy::desc_lo(a_s), b_lo = hy::desc_lo(b_s);
    long long issue = 0, total = 0;
    uint32_t par = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < NMMA; ++i)
        hy::umma_issue<MODE>(tmem + (i / 16 % 2) * 256, hy::desc_pack(a_lo + ((((i / 4) % 8) * 16384 + (i % 4) * 32) >> 4), hi),
                             hy::desc_pack(b_lo + ((((i / 4) % 2) * 32768 + (i % 4) * 32) >> 4), hi), idesc, i % 16 ? 1u : 0u, leader);
      hy::umma_commit_mode<MODE>(bar, leader);
      const long long t1 = clock64();
      hy::mbar_wait(bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      issue += t1 - t0;
      total += t2 - t0;
    }
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2] = issue / reps;
      out[blockIdx.x * 2 + 1] = total / reps;
    }
  }
  hy::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 512); }
}

template <int MODE, int N, int NMMA>
void run_conv_issue(long long* d) {
  const int smem = 8 * 16384 + 2 * 32768 + 1024 + 64;
  cudaFuncSetAttribute(conv_issue<MODE, N, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  conv_issue<MODE, N, NMMA><<<148, 128, smem>>>(d, 100);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("mode %d N=%3d %3d MMAs + commit + wait: issue %5lld clk, total %5lld clk => %.1f clk/MMA, frac of 4096 MAC/clk %.2f (%s)\n", MODE, N, NMMA,
         h[0], h[1], (double)h[1] / NMMA, 128.0 * N * 16 * NMMA / h[1] / 4096, cudaGetErrorString(e));
}

struct ConvIssueRunner {
  ConvIssueRunner() {
    long long* d;
    cudaMalloc(&d, 148 * 16);
    run_conv_issue<1, 64, 64>(d);
    run_conv_issue<2, 64, 64>(d);
    run_conv_issue<2, 16, 64>(d);
    run_conv_issue<2, 32, 64>(d);
    run_conv_issue<2, 64, 16>(d);
    run_conv_issue<2, 64, 128>(d);
    run_conv_issue<2, 96, 64>(d);
    run_conv_issue<2, 128, 64>(d);
    run_conv_issue<2, 192, 64>(d);
    run_conv_issue<2, 256, 64>(d);
    run_conv_issue<1, 128, 64>(d);
    run_conv_issue<1, 256, 64>(d);
  }
} conv_issue_runner;

// ---- fourth experiment: provably warp-uniform role dispatch (shfl-derived warp index), descriptors in plain
// integer arithmetic so that ptxas can keep them in uniform registers ----
template <int N, int NMMA>
__global__ void __launch_bounds__(128) uni_issue(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 8 * 16384, bar = b_s + 2 * 32768, slot = bar + 8;
  for (int i = threadIdx.x; i < (8 * 16384 + 2 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - hy::smem_u32(raw)))[i] = 0x3c003c00u;
  hy::fence_async_smem();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (warp == 0) { hy::tmem_alloc(slot, 512); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem_v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_v) : "r"(slot));
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_v, 0);
  if (warp == 0) {
    const uint32_t leader = hy::elect_leader();
    const uint32_t idesc = hy::umma_idesc_bf16(128, N);
    constexpr uint64_t hi = static_cast<uint64_t>(hy::desc_hi_sw128()) << 32;
    const uint64_t a_d = hi | hy::desc_lo(a_s), b_d = hi | hy::desc_lo(b_s);
    long long issue = 0, total = 0;
    uint32_t par = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < NMMA; ++i)
        hy::umma_issue<2>(tmem + (i / 16 % 2) * 256, a_d + ((((i / 4) % 8) * 16384 + (i % 4) * 32) >> 4),
                          b_d + ((((i / 4) % 2) * 32768 + (i % 4) * 32) >> 4), idesc, i % 16 ? 1u : 0u, leader);
      hy::umma_commit_mode<2>(bar, leader);
      const long long t1 = clock64();
      hy::mbar_wait(bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      issue += t1 - t0;
      total += t2 - t0;
    }
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2] = issue / reps;
      out[blockIdx.x * 2 + 1] = total / reps;
    }
  }
  hy::tc_fence_before();
  __syncthreads();
  if (warp == 0) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 512); }
}

template <int N, int NMMA>
void run_uni_issue(long long* d) {
  const int smem = 8 * 16384 + 2 * 32768 + 1024 + 64;
  cudaFuncSetAttribute(uni_issue<N, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  uni_issue<N, NMMA><<<148, 128, smem>>>(d, 100);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("uniform N=%3d %3d MMAs + commit + wait: issue %5lld clk, total %5lld clk => %.1f clk/MMA, frac of 4096 MAC/clk %.2f (%s)\n", N, NMMA,
         h[0], h[1], (double)h[1] / NMMA, 128.0 * N * 16 * NMMA / h[1] / 4096, cudaGetErrorString(e));
}

struct UniIssueRunner {
  UniIssueRunner() {
    long long* d;
    cudaMalloc(&d, 148 * 16);
    run_uni_issue<16, 64>(d);
    run_uni_issue<32, 64>(d);
    run_uni_issue<64, 64>(d);
    run_uni_issue<64, 128>(d);
    run_uni_issue<96, 64>(d);
    run_uni_issue<128, 64>(d);
    run_uni_issue<256, 64>(d);
  }
} uni_issue_runner;

// ---- fifth experiment: the resident-weights kernel's dynamic issue loop (steps in the parameter block) ----
struct DynStep { int32_t a_off16, b_off16, d_col, acc; };
struct alignas(64) DynParams { DynStep steps[64]; int32_t nsteps, N, reps, pad; long long* out; };

template <int STYLE>  // 0: lane 0 only (branch on threadIdx), 1: converged warp, shfl-derived warp id, leader predicate
__global__ void __launch_bounds__(128) dyn_issue(const __grid_constant__ DynParams p) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (hy::smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 8 * 16384, bar = b_s + 2 * 32768, slot = bar + 8;
  for (int i = threadIdx.x; i < (8 * 16384 + 2 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(raw + (base - hy::smem_u32(raw)))[i] = 0x3c003c00u;
  hy::fence_async_smem();
  const int warp = STYLE ? __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) : (threadIdx.x >> 5);
  if (threadIdx.x == 0) { hy::mbar_init(bar, 1); hy::mbar_fence_init(); }
  if (warp == 0) { hy::tmem_alloc(slot, 512); hy::tmem_relinquish(); }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem_v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_v) : "r"(slot));
  const uint32_t tmem = STYLE ? __shfl_sync(0xffffffffu, tmem_v, 0) : tmem_v;
  if (warp == 0 && (STYLE || (threadIdx.x & 31) == 0)) {
    const uint32_t leader = STYLE ? hy::elect_leader() : 1u;
    const uint32_t idesc = hy::umma_idesc_bf16(128, p.N);
    constexpr uint64_t hi = static_cast<uint64_t>(hy::desc_hi_sw128()) << 32;
    const uint64_t a_d = hi | hy::desc_lo(a_s), b_d = hi | hy::desc_lo(b_s);
    long long issue = 0, total = 0;
    uint32_t par = 0;
    for (int rep = 0; rep < p.reps; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < p.nsteps; ++i) {
        const DynStep st = p.steps[i];
        const uint64_t a0 = a_d + st.a_off16, b0 = b_d + st.b_off16;
        const uint32_t d = tmem + st.d_col;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (STYLE) hy::umma_issue<2>(d, a0 + 2 * k, b0 + 2 * k, idesc, static_cast<uint32_t>(st.acc | k), leader);
          else hy::umma_bf16(d, a0 + 2 * k, b0 + 2 * k, idesc, static_cast<uint32_t>(st.acc | k));
        }
      }
      if (STYLE) hy::umma_commit_mode<2>(bar, leader); else hy::umma_commit(bar);
      const long long t1 = clock64();
      hy::mbar_wait(bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      issue += t1 - t0;
      total += t2 - t0;
    }
    if ((threadIdx.x & 31) == 0) {
      p.out[blockIdx.x * 2] = issue / p.reps;
      p.out[blockIdx.x * 2 + 1] = total / p.reps;
    }
  }
  hy::tc_fence_before();
  __syncthreads();
  if (warp == 0) { hy::tc_fence_after(); hy::tmem_dealloc(tmem, 512); }
}

struct DynRunner {
  DynRunner() {
    long long* d;
    cudaMalloc(&d, 148 * 16);
    const int smem = 8 * 16384 + 2 * 32768 + 1024 + 64;
    cudaFuncSetAttribute(dyn_issue<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dyn_issue<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {16, 64, 128})
      for (int style = 0; style < 2; ++style) {
        DynParams p;
        memset(&p, 0, sizeof p);
        p.nsteps = 9; p.N = N; p.reps = 100; p.out = d;
        for (int i = 0; i < 9; ++i) {
          p.steps[i].a_off16 = ((i / 3) * 10 + i % 3) * 128 >> 4;
          p.steps[i].b_off16 = ((i % 4) * N * 128) >> 4;
          p.steps[i].d_col = 0;
          p.steps[i].acc = i ? 1 : 0;
        }
        if (style) dyn_issue<1><<<148, 128, smem>>>(p); else dyn_issue<0><<<148, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("dynamic loop style %d N=%3d 36 MMAs + commit + wait: issue %5lld clk, total %5lld clk => %.1f clk/MMA (%s)\n", style, N,
               h[0], h[1], h[1] / 36.0, cudaGetErrorString(e));
      }
  }
} dyn_runner;
