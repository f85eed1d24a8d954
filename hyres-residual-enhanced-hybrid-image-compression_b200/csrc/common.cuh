// Shared device helpers for the HyRES B200 hot path: raw PTX wrappers for
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM) and small math.
// sm_100a only. No CUTLASS dependency: the descriptor encodings are written out
// here (bit layouts follow the PTX ISA "matrix descriptor" / "instruction
// descriptor" tables for tcgen05).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

#define HYRES_OK 0
#define HYRES_ERR_ARG -1
#define HYRES_ERR_CUDA -2
#define HYRES_ERR_DRIVER -3
#define HYRES_ERR_UNSUPPORTED -4
#define HYRES_ERR_STATE -5

namespace hy {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug must fault (trap) rather than hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && (clock64() - t0) > 4000000000LL) {
      printf("hyres: mbarrier wait timed out: block %d of %d (%d threads), thread %d, barrier 0x%x parity %u\n",
             blockIdx.x, gridDim.x, blockDim.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Named barrier over `nthreads` threads of the CTA (id 1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still draining.  Every kernel of the convolution path lets its
// successor in right away (pdl_launch_dependents, first instruction) and runs its own prologue -- barrier
// init, TMEM allocation, the TMA load of its weights: nothing the predecessor can write -- before pdl_wait,
// after which the predecessor has completed and its memory is visible.  All threads wait: even the epilogue
// stores must not land before the predecessor finished reading (the allocator may have recycled its inputs).
// ----------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a tile (no shared memory, no barrier): issued a few tiles ahead so that the real load, which can
// only start once its shared-memory stage is free, finds its data in L2 instead of paying the HBM latency.
__device__ __forceinline__ void tma_prefetch_4d(const void* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* map, uint32_t src, int c0, int c1, int c2,
                                             int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 operands, fp32 accumulate, cta_group::1.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-converged variants: every lane of the issuing warp executes the surrounding (warp-uniform)
// address arithmetic, so the descriptors live in uniform registers, and one elected lane issues.
// (A loop that runs inside `if (lane == 0)` makes the compiler rebuild every descriptor in vector
// registers and move it to the uniform file per instruction: measured ~64 clk per MMA issued, twice the
// tensor time of a 128x64x16 MMA.)  elect.sync picks the same lane every time, so tcgen05.commit issued
// the same way tracks these MMAs.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}"
      ::"r"(bar)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05 ops of this thread finished.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// Wait for outstanding tcgen05.ld; the registers are in/out operands so that the compiler
// cannot schedule their first use above the wait.
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld with the 32 destination registers as in/out operands (see tmem_ld_fence).
__device__ __forceinline__ void tmem_ld_fence32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]),
                 "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]),
                 "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                 "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]),
                 "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B, bf16:
// rows are 128 B apart, 8-row groups SBO=1024 B apart, LBO unused (encoded 1).
//   [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2
// The swizzle is a function of the absolute shared-memory address (measured:
// tools/experiments/umma_shift.cu), so `saddr` may be any 128 B row of a buffer written with
// chunk ^= (addr >> 7) & 7, and `sbo` (byte distance between successive 8-row groups) need
// not be 1024: a patch that is P positions wide uses sbo = P * 128 and any tap of a
// convolution is just another start row.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t sbo = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// The same descriptor split into words: the start-address field is the only part that changes between
// the MMAs of a tile, and since every shared-memory address is < 256 KB, advancing it by `bytes` is a
// plain add of bytes >> 4 to the low word -- one integer add per operand per MMA in the issue loop.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3fffu) | (1u << 16); }
__host__ __device__ constexpr uint32_t desc_hi_sw128(uint32_t sbo = 1024) { return (sbo >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// Descriptor as plain integer arithmetic (no inline asm): in a converged warp whose role was picked through a
// warp-uniform value (warp index broadcast by __shfl_sync), ptxas keeps such descriptors and their per-MMA
// offsets in uniform registers (UIADD3.64), so an MMA costs one UTCHMMA plus one or two uniform adds.
__device__ __forceinline__ uint64_t desc_u64(uint32_t saddr, uint32_t sbo = 1024) {
  return (static_cast<uint64_t>(desc_hi_sw128(sbo)) << 32) | desc_lo(saddr);
}
// Packed fp32 pair add (FADD2): (a, b) += (c, d).
__device__ __forceinline__ void add2(float& a, float& b, float c, float d) {
  asm("{\n\t"
      ".reg .b64 x, y;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %3};\n\t"
      "add.rn.f32x2 x, x, y;\n\t"
      "mov.b64 {%0, %1}, x;\n\t"
      "}"
      : "+f"(a), "+f"(b)
      : "f"(c), "f"(d));
}
// max(x, 0) on a packed bf16 pair (HMNMX2): ReLU commutes with the bf16 rounding, so relu(pack(a, b)) ==
// pack(relu(a), relu(b)).
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t u) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(u), "r"(0u));
  return r;
}
// MMA issue in one of three styles (MODE): 0 = the calling code runs in a single lane; 1 = converged warp,
// elect.sync per MMA; 2 = converged warp, `leader` (1 in exactly one lane) decided once by the caller.
template <int MODE>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate, uint32_t leader) {
  if (MODE == 0) {
    umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  } else if (MODE == 1) {
    umma_bf16_elect(tmem_d, adesc, bdesc, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
        : "memory");
  }
}
template <int MODE>
__device__ __forceinline__ void umma_commit_mode(uint32_t bar, uint32_t leader) {
  if (MODE == 0) {
    umma_commit(bar);
  } else if (MODE == 1) {
    umma_commit_elect(bar);
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}"
        ::"r"(bar), "r"(leader)
        : "memory");
  }
}
// tensor-memory stores (epilogue -> TMEM): 8 columns per call, lane = this thread's lane of its warp's quadrant
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from tensor memory (bf16, two K elements per 32-bit column,
// element 2j in the low half of column j, lane = row; 8 columns per K = 16 instruction) -- measured in
// tools/experiments/umma_ts.cu.  Converged-warp form: `leader` is 1 in exactly one lane.
__device__ __forceinline__ void umma_ts_issue(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_leader() {
  uint32_t is_leader;
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, e;\n\t"
      "}"
      : "=r"(is_leader));
  return is_leader;
}
// Instruction descriptor, kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ inline uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// Same with A = B = IEEE half (format code 0): the half-part split-precision layers.
__host__ __device__ inline uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ----------------------------------------------------------------------------
// split-precision parts of an fp32 value (hyres_b200.h: hyres_conv_create_split)
// ----------------------------------------------------------------------------
constexpr int kSplitF16 = 16;            // format flag in an nsplit code (HYRES_SPLIT_F16)
constexpr float kF16LoScale = 2048.f;    // second half part = half((v - p0) * 2^11)
constexpr float kF16LoInv = 1.f / 2048.f;
// bf16 parts: v = p0 + p1 + p2 (+ O(2^-25 |v|)), each rounded to nearest; the residuals are exact in fp32
__device__ __forceinline__ void split_bf16x3(float v, uint16_t& p0, uint16_t& p1, uint16_t& p2) {
  const __nv_bfloat16 b0 = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(b0);
  const __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(b1);
  p0 = __bfloat16_as_ushort(b0);
  p1 = __bfloat16_as_ushort(b1);
  p2 = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
}
// half parts: v = p0 + p1 * 2^-11 to 2^-23 |v| (or 2^-36 absolute when p1 is subnormal); the residual v - p0 is
// exact in fp32 and stays in the range of p0 after scaling.  |v| >= 65520 overflows to infinity (and the
// products to NaN): loud, not silent.
__device__ __forceinline__ void split_f16x2(float v, uint16_t& p0, uint16_t& p1) {
  const __half h0 = __float2half_rn(v);
  const float r = (v - __half2float(h0)) * kF16LoScale;
  p0 = __half_as_ushort(h0);
  p1 = __half_as_ushort(__float2half_rn(r));
}

// ----------------------------------------------------------------------------
// small math
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// Branch-free approximate transcendentals for epilogues whose result is rounded to bf16 (or compared
// at 2e-4): the IEEE-accurate divide / sqrt expand to long dependent sequences with slow-path calls
// that serialise the 16-wide epilogue chunks (measured: the gate epilogue was 3x over its HBM floor).
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.f + fast_ex2(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_rsqrt(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Order-independent accumulation of block partials into a device double: every partial is first rounded to a
// multiple of 2^-18, so as long as |sum| < 2^34 every intermediate sum is exactly representable in a double and
// the additions commute -- the result does not depend on the order in which the blocks arrive (run-to-run
// deterministic), at a rounding cost of 2^-19 per block (1e-9 relative on the rate / distortion sums).
__device__ __forceinline__ void atomic_add_exact(double* out, double partial) {
  atomicAdd(out, rint(partial * 262144.0) * (1.0 / 262144.0));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace hy
