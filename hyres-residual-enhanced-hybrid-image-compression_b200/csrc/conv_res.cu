// Persistent implicit-GEMM convolution with shared-memory-resident weights.
//
// The generic kernel (conv_tc.cu) gives each CTA one output tile and streams the weights of
// every tap through a ring; for the many small-K layers of the codec (1x1, 3x3 64->64, the
// 3-channel output layers) that makes a tile's life a serial latency chain and re-reads the
// weights from L2 once per tile.  This kernel is the path for every layer whose packed weights
// fit in shared memory next to the activation stages:
//
//   * one CTA per SM, weights loaded once; tiles (16x8 output positions) are walked grid-stride;
//   * A operand: ONE halo patch per 64-channel chunk per tile ((16+(R-1)d) x (8+(S-1)d)
//     positions, TMA, image borders zero-filled by the TMA unit).  Every tap (r,s) is the same
//     patch addressed through a UMMA descriptor whose start row is r*d*PW + s*d and whose 8-row
//     groups are PW*128 B apart -- no im2col and no per-column re-load;
//   * transposed 5x5/stride-2 convs: the four sub-pixel phases share the one input patch and
//     accumulate into four TMEM column ranges;
//   * TMEM accumulators are double buffered; two epilogue warp groups take alternate tiles, so
//     the tensor pipe, the TMA loads and two epilogues are in flight at once;
//   * epilogue operands (skip / gate / GDN inputs) arrive by TMA into the stage, the bf16 result
//     is written over them in place and leaves by TMA store (coalesced, clipped at the edge);
//   * GDN: the kernel squares its own A tile in shared memory (x0_square), so producers need not
//     write x^2 to HBM.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kTH = 16, kTW = 8;
constexpr int kThreads = 384;  // warps 0-3 / 4-7: epilogue groups 0 / 1; warp 8: TMA loads; warp 9: MMA;
                               // warps 10 / 11: TMA stores + stage release of group 0 / 1
constexpr int kThreadsSq = 512;  // + warps 12-15: x0_square transform (GDN layers only)
constexpr int kMaxSteps = 64;
constexpr int kMaxStages = 6;
constexpr int kSmemLimit = 227 * 1024;

// One 64-deep k-step (four MMAs), with everything the issue loop needs precomputed on the host: the
// loop is one 16-byte constant load and two adds per operand (measured: the issue thread, not the
// tensor pipe, bounded the many-tap / narrow-N layers when it rebuilt descriptors per MMA).
struct Step {
  int32_t a_off16;  // (chunk * a_chunk_bytes + tap start row * 128) >> 4, relative to the stage base
  int32_t b_off16;  // (k-slot * BN * 128) >> 4, relative to the weight base
  int32_t d_col;    // phase * BN: accumulator column offset
  int32_t acc;      // 0 for the first step of a phase (overwrite), 1 afterwards
};

#define RES_STAMP(itv, slot) do { if (p.trace && blockIdx.x == 0 && (itv) < 64) p.trace[(itv) * 16 + (slot)] = clock64(); } while (0)

struct alignas(64) ResParams {
  CUtensorMap mapA, mapW, mapOut, mapAux0, mapAux1;
  CUtensorMap mapT2, mapT3, mapU;  // up-add mode: half / quarter resolution addends and the interpolation matrices
  Step steps[kMaxSteps];
  int32_t nsteps, nphase, nchunk_in, nchunk_out;
  int32_t PW, PH, org_h, org_w;
  int32_t a_chunk_bytes, stage_bytes, o_off, x1_off, NA;
  int32_t w_bytes, nslots, BN;
  int32_t stage_tx_bytes;
  int32_t tmem_cols;
  int32_t Hv, Wv, OH, OW, out_mul;
  int32_t tiles_w, tiles_per_img, ntiles;
  int32_t cout, epi, act;
  float slope;
  int32_t a_square, has_aux0, has_aux1, store_bf16;
  int32_t pf_extra;  // L2 prefetch distance beyond the ring, in tiles; < 0: no prefetch
  int32_t up, u_bytes, t2_off, t3_off;  // up-add mode (see conv_res_try_run)
  int32_t decoupled, stg_base_off, stg_per_grp;  // result staging outside the TMA ring (1 or 2 buffers per epilogue group)
  int32_t cta_limit;  // > 0: at most this many CTAs
  long long* trace;  // optional: clock64 stamps of CTA 0 (tools/experiments/res_trace.py), 16 slots per tile
  const float* bias;
  const float* pixscale;
  float* out_f32;
  long long f32_sb, f32_sh, f32_sw, f32_sc;
};

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t sq_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  t = __hmul2(t, t);
  return *reinterpret_cast<uint32_t*>(&t);
}
// (a, b) = (a, b) * s + (c, d) and (a, b) *= s as packed fp32 pairs (FFMA2 / FMUL2)
__device__ __forceinline__ void fma2(float& a, float& b, float s, float c, float d) {
  asm("{\n\t"
      ".reg .b64 x, y, z;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %2};\n\t"
      "mov.b64 z, {%3, %4};\n\t"
      "fma.rn.f32x2 x, x, y, z;\n\t"
      "mov.b64 {%0, %1}, x;\n\t"
      "}"
      : "+f"(a), "+f"(b)
      : "f"(s), "f"(c), "f"(d));
}
__device__ __forceinline__ void mul2(float& a, float& b, float s) {
  asm("{\n\t"
      ".reg .b64 x, y;\n\t"
      "mov.b64 x, {%0, %1};\n\t"
      "mov.b64 y, {%2, %2};\n\t"
      "mul.rn.f32x2 x, x, y;\n\t"
      "mov.b64 {%0, %1}, x;\n\t"
      "}"
      : "+f"(a), "+f"(b)
      : "f"(s));
}
__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = hy::bf16_lo(u[i]);
    f[2 * i + 1] = hy::bf16_hi(u[i]);
  }
}

// EPI / ACT >= 0: the epilogue mode / activation are compile-time constants (straight-line chunk code: measured,
// the generic epilogue spends a third of a chunk's ~500 clk in uniform branches and instruction fetch); -1: taken
// from the parameters at run time.
template <int EPI, int ACT>
__global__ void __launch_bounds__(kThreadsSq, 1) conv_res_kernel(const __grid_constant__ ResParams p) {
  const int epi = EPI >= 0 ? EPI : p.epi;
  const int act = ACT >= 0 ? ACT : p.act;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;
  const uint32_t u_base = base + p.w_bytes;  // interpolation matrices (up-add mode), resident like the weights
  const uint32_t st_base = u_base + p.u_bytes;
  const uint32_t stg_base = st_base + p.stg_base_off;
  const uint32_t bias_base = stg_base + (p.decoupled ? 2 * p.stg_per_grp * p.nchunk_out * 16384 : 0);
  const uint32_t bar_base = bias_base + 1024;
  // barriers: W_FULL | A_FULL[4] | A_EMPTY[4] | A_READY[4] | ACC_FULL[2] | ACC_EMPTY[2] ; tmem slot
  const uint32_t W_FULL = bar_base;
  const uint32_t A_FULL = bar_base + 8;
  const uint32_t A_EMPTY = A_FULL + 8 * kMaxStages;
  const uint32_t A_READY = A_EMPTY + 8 * kMaxStages;
  const uint32_t ACC_FULL = A_READY + 8 * kMaxStages;
  const uint32_t ACC_EMPTY = ACC_FULL + 16;
  // STAGED[2]: an epilogue group's result tile is staged (and its operands consumed); STAGED_ACK[2]: the group's store
  // warp has seen that.  The group waits for the acknowledgement of tile k - 1 before it signals tile k, so the
  // barrier can never complete two phases while the store warp is still busy with an earlier tile (under memory
  // pressure from kernels of other streams its TMA-store read wait can outlast two epilogues; a waiter that misses a
  // phase never wakes up -- seen as a watchdog trap when compress and decompress ran concurrently on several streams).
  const uint32_t STAGED = ACC_EMPTY + 16;
  const uint32_t STAGED_ACK = STAGED + 16;
  const uint32_t STG_FREE = STAGED_ACK + 16;  // [2][2]: decoupled mode: a staging buffer of the group has been read by its store
  const bool store_role = p.store_bf16 || !p.decoupled;  // otherwise the store warps have nothing to send or release
  const uint32_t tmem_slot = STG_FREE + 32;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    hy::mbar_init(W_FULL, 1);
    for (int i = 0; i < p.NA; ++i) {
      hy::mbar_init(A_FULL + 8 * i, 1);
      // coupled: MMA commit + the store warp's release (the stage also holds the epilogue operands / the result);
      // decoupled: the MMA commit alone frees the activation patch
      hy::mbar_init(A_EMPTY + 8 * i, p.decoupled ? 1 : 2);
      hy::mbar_init(A_READY + 8 * i, 128);  // x0_square transform
    }
    for (int i = 0; i < 2; ++i) {
      hy::mbar_init(ACC_FULL + 8 * i, 1);
      hy::mbar_init(ACC_EMPTY + 8 * i, 128);
      hy::mbar_init(STG_FREE + 16 * i, 1);
      hy::mbar_init(STG_FREE + 16 * i + 8, 1);
    }
    for (int i = 0; i < 2; ++i) {
      hy::mbar_init(STAGED + 8 * i, 128);
      hy::mbar_init(STAGED_ACK + 8 * i, 1);
    }
    hy::mbar_fence_init();
    // the weights do not depend on the predecessor kernel: their load starts before the dependency wait
    if (p.nslots) hy::tma_prefetch_desc(&p.mapW);
    hy::mbar_arrive_expect_tx(W_FULL, p.w_bytes + p.u_bytes);
    for (int s = 0; s < p.nslots; ++s) hy::tma_load_2d(w_base + s * p.BN * 128, &p.mapW, W_FULL, s * 64, 0);
    if (p.up) {
      hy::tma_prefetch_desc(&p.mapU);
      hy::tma_load_2d(u_base, &p.mapU, W_FULL, 0, 0);
      hy::tma_load_2d(u_base + 16384, &p.mapU, W_FULL, 0, 128);
    }
  }
  {
    float* sb = reinterpret_cast<float*>(smem_raw + (bias_base - hy::smem_u32(smem_raw)));
    if (p.bias)
      for (int i = threadIdx.x; i < p.BN; i += blockDim.x) sb[i] = __ldg(p.bias + i);  // bias padded to BN
  }
  if (warp == 8 && lane == 0) {
    hy::tma_prefetch_desc(&p.mapA);
    if (p.nslots) hy::tma_prefetch_desc(&p.mapW);
    if (p.has_aux0) hy::tma_prefetch_desc(&p.mapAux0);
    if (p.has_aux1) hy::tma_prefetch_desc(&p.mapAux1);
    if (p.up) {
      hy::tma_prefetch_desc(&p.mapT2);
      hy::tma_prefetch_desc(&p.mapT3);
    }
  }
  if (warp == 9) {
    hy::tmem_alloc(tmem_slot, p.tmem_cols);
    hy::tmem_relinquish();
  }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  // Dependents (the next kernel of this stream) may be scheduled from here on -- only AFTER this CTA owns its
  // tensor memory: a dependent that lands on the same SM allocates TMEM in its prologue and then waits for this
  // grid to finish, so it must never be able to take the columns this CTA still has to allocate.
  hy::pdl_launch_dependents();
  uint32_t tmem_base_v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base_v) : "r"(tmem_slot));
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_v, 0);
  hy::pdl_wait();  // everything above is independent of the predecessor kernel

  auto tile_origin = [&](int t, int& b_img, int& h0, int& w0) {
    b_img = t / p.tiles_per_img;
    const int rem = t - b_img * p.tiles_per_img;
    const int th = rem / p.tiles_w;
    h0 = th * kTH;
    w0 = (rem - th * p.tiles_w) * kTW;
  };
  const int acc_cols = p.nphase * p.BN;

  if (warp == 8) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t par = 0;
      auto prefetch = [&](int t) {
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        for (int c = 0; c < p.nchunk_in; ++c) hy::tma_prefetch_4d(&p.mapA, c * 64, w0 + p.org_w, h0 + p.org_h, b_img);
        if (p.has_aux0)
          for (int c = 0; c < p.nchunk_out; ++c) hy::tma_prefetch_4d(&p.mapAux0, c * 64, w0, h0, b_img);
        if (p.has_aux1)
          for (int c = 0; c < p.nchunk_out; ++c) hy::tma_prefetch_4d(&p.mapAux1, c * 64, w0, h0, b_img);
      };
      // the ring is shallow when a stage also holds the epilogue operands (gate: two stages): tiles beyond the ring
      // are prefetched into L2, so a freed stage refills at L2 latency
      const int ahead = (p.NA + p.pf_extra) * static_cast<int>(gridDim.x);
      for (int k = 0; k < p.pf_extra; ++k) {
        const int t = blockIdx.x + (p.NA + k) * static_cast<int>(gridDim.x);
        if (p.pf_extra >= 0 && t < p.ntiles) prefetch(t);
      }
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        const uint32_t sb = st_base + stage * p.stage_bytes;
        hy::mbar_wait(A_EMPTY + 8 * stage, par ^ 1u);
        RES_STAMP((t - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x), 0);  // load issued
        hy::mbar_arrive_expect_tx(A_FULL + 8 * stage, p.stage_tx_bytes);
        for (int c = 0; c < p.nchunk_in; ++c)
          hy::tma_load_4d(sb + c * p.a_chunk_bytes, &p.mapA, A_FULL + 8 * stage, c * 64, w0 + p.org_w, h0 + p.org_h, b_img);
        if (p.has_aux0)
          for (int c = 0; c < p.nchunk_out; ++c)
            hy::tma_load_4d(sb + p.o_off + c * 16384, &p.mapAux0, A_FULL + 8 * stage, c * 64, w0, h0, b_img);
        if (p.has_aux1)
          for (int c = 0; c < p.nchunk_out; ++c)
            hy::tma_load_4d(sb + p.x1_off + c * 16384, &p.mapAux1, A_FULL + 8 * stage, c * 64, w0, h0, b_img);
        if (p.up) {
          // addend patches of the tile, read from tensors padded by one replicated pixel: tile rows h0 .. h0+15
          // interpolate half-resolution rows h0/2-1 .. h0/2+8 (padded coordinate h0/2), quarter rows h0/4-1 .. h0/4+4
          hy::tma_load_4d(sb + p.t2_off, &p.mapT2, A_FULL + 8 * stage, 0, w0 >> 1, h0 >> 1, b_img);
          hy::tma_load_4d(sb + p.t3_off, &p.mapT3, A_FULL + 8 * stage, 0, w0 >> 2, h0 >> 2, b_img);
        }
        if (p.pf_extra >= 0 && t + ahead < p.ntiles) prefetch(t + ahead);
        if (++stage == p.NA) { stage = 0; par ^= 1u; }
      }
    }
  } else if (warp == 9) {
    // ============================ MMA issuer ============================
    // The whole warp runs the loop converged (descriptor arithmetic in uniform registers); one elected lane issues.
    {
      const uint32_t leader = hy::elect_leader();
      const uint32_t idesc = hy::umma_idesc_bf16(128, p.BN);
      const uint64_t w_d = hy::desc_u64(w_base);
      const uint32_t sbo_a = p.PW * 128;
      hy::mbar_wait(W_FULL, 0);
      int stage = 0, it = 0;
      uint32_t par = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint64_t a_d0 = hy::desc_u64(st_base + stage * p.stage_bytes, sbo_a);
        const uint32_t d0 = tmem_base + buf * acc_cols;
        hy::mbar_wait((p.a_square ? A_READY : A_FULL) + 8 * stage, par);
        if (lane == 0) RES_STAMP(it, 1);  // operands landed
        hy::mbar_wait(ACC_EMPTY + 8 * buf, ((it >> 1) & 1) ^ 1u);
        if (lane == 0) RES_STAMP(it, 2);  // accumulator free
        hy::tc_fence_after();
        for (int i = 0; i < p.nsteps; ++i) {
          const Step st = p.steps[i];
          const uint64_t a_d = a_d0 + static_cast<uint32_t>(st.a_off16), b_d = w_d + static_cast<uint32_t>(st.b_off16);
          const uint32_t d = d0 + st.d_col;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            hy::umma_issue<2>(d, a_d + 2 * k, b_d + 2 * k, idesc, static_cast<uint32_t>(st.acc | k), leader);
        }
        if (p.up) {
          // acc += U2 . T2patch + U4 . T3patch: bilinear x2 / x4 up-sampling as two small GEMMs.  A = the constant
          // interpolation matrix [128 tile positions][64 patch pixels] (K-major); B = the patch as TMA wrote it,
          // [patch pixel][64 channels], i.e. N-contiguous: an MN-major operand (idesc bit 16; 16 pixels = 2048 B
          // per K step; measured in tools/experiments/umma_mn.cu).
          const uint32_t idesc_mn = hy::umma_idesc_bf16(128, 64) | (1u << 16);
          // up == 2 (statistics mode): no weight GEMM; the two up-sampled tensors go to columns [0, 64) and [64, 128)
          const uint32_t d1 = p.up == 2 ? d0 + 64 : d0;
          const uint64_t u2_d = hy::desc_u64(u_base), u4_d = hy::desc_u64(u_base + 16384);
          const uint64_t t2_d = hy::desc_u64(st_base + stage * p.stage_bytes + p.t2_off);
          const uint64_t t3_d = hy::desc_u64(st_base + stage * p.stage_bytes + p.t3_off);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            hy::umma_issue<2>(d0, u2_d + 2 * k, t2_d + 128 * k, idesc_mn, (p.up == 2 && k == 0) ? 0u : 1u, leader);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            hy::umma_issue<2>(d1, u4_d + 2 * k, t3_d + 128 * k, idesc_mn, (p.up == 2 && k == 0) ? 0u : 1u, leader);
        }
        hy::umma_commit_mode<2>(A_EMPTY + 8 * stage, leader);
        hy::umma_commit_mode<2>(ACC_FULL + 8 * buf, leader);
        if (lane == 0) RES_STAMP(it, 3);  // MMAs issued
        if (++stage == p.NA) { stage = 0; par ^= 1u; }
      }
    }
  } else if (warp == 10 || warp == 11) {
    // ============================ store warps ============================
    // One per epilogue group: waits until the group has staged a tile, sends it with a TMA store and releases the
    // stage once the store has read it.  (Measured with the clock64 trace: when a thread of the group did this,
    // its warp -- and through the group barrier the whole group -- lost ~880 clk per tile waiting for the store.)
    if (lane == 0 && store_role) {
      const int grp = warp - 10;
      for (int it = grp;; it += 2) {
        const int t = blockIdx.x + it * static_cast<int>(gridDim.x);
        if (t >= p.ntiles) break;
        const int stage = it % p.NA;
        const uint32_t sb = st_base + stage * p.stage_bytes;
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        const int sk = p.stg_per_grp == 2 ? ((it >> 1) & 1) : 0;  // staging buffer of this tile (decoupled mode)
        const uint32_t src = p.decoupled ? stg_base + (grp * p.stg_per_grp + sk) * p.nchunk_out * 16384 : sb + p.o_off;
        hy::mbar_wait(STAGED + 8 * grp, (it >> 1) & 1);
        hy::mbar_arrive(STAGED_ACK + 8 * grp);
        if (p.store_bf16) {
          for (int c = 0; c < p.nchunk_out; ++c) hy::tma_store_4d(&p.mapOut, src + c * 16384, c * 64, w0, h0, b_img);
          hy::tma_store_commit();
          hy::tma_store_wait_read<0>();
        }
        hy::mbar_arrive(p.decoupled ? STG_FREE + 16 * grp + 8 * sk : A_EMPTY + 8 * stage);
        RES_STAMP(it, 8);  // store read done, stage released
      }
      if (p.store_bf16) hy::tma_store_wait_all<0>();
    }
  } else if (warp >= 12) {
    // ============================ x0_square transform (GDN) ============================
    // GDN operand: square the activation tile in place (bf16 RN of the exact product, the same value a
    // producer-side x*x store would have held), then release it to the MMA warp.  Its own four warps, so the
    // transform of tile t+1 runs under the MMAs of tile t and the epilogues of tiles t-1, t-2.
    if (p.a_square) {
      const int row = threadIdx.x - 384;
      int stage = 0;
      uint32_t par = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const uint32_t sb = st_base + stage * p.stage_bytes;
        hy::mbar_wait(A_FULL + 8 * stage, par);
        for (int c = 0; c < p.nchunk_in; ++c) {
          const uint32_t r0 = sb + c * p.a_chunk_bytes + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // element-wise and in place, so any visiting order works: rotate the 16 B chunk by the row so
            // that the 8 rows of a quarter-warp touch 8 different bank groups (rows are 128 B apart)
            const uint32_t a = r0 + ((j ^ (row & 7)) << 4);
            uint4 v = lds128(a);
            v.x = sq_bf16x2(v.x); v.y = sq_bf16x2(v.y); v.z = sq_bf16x2(v.z); v.w = sq_bf16x2(v.w);
            sts128(a, v);
          }
        }
        hy::fence_async_smem();
        hy::mbar_arrive(A_READY + 8 * stage);
        if (++stage == p.NA) { stage = 0; par ^= 1u; }
      }
    }
  } else {
    // ============================ epilogue groups ============================
    const int grp = warp >> 2;
    const int row = threadIdx.x & 127;  // TMEM lane == tile position
    const int ti = row >> 3, tj = row & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t sw = row & 7;
    const int nchunk16 = p.BN >> 4;
    const bool need0 = p.has_aux0 != 0;
    const bool pixscale1 = epi == HYRES_EPI_PIXSCALE && p.nphase == 1;  // one scale per tile position
    for (int it = grp;; it += 2) {
      const int t = blockIdx.x + it * static_cast<int>(gridDim.x);
      if (t >= p.ntiles) break;
      const int stage = it % p.NA;
      const uint32_t par = (it / p.NA) & 1;
      const uint32_t sb = st_base + stage * p.stage_bytes;
      int b_img, h0, w0;
      tile_origin(t, b_img, h0, w0);
      const int hv = h0 + ti, wv = w0 + tj;
      const bool valid = hv < p.Hv && wv < p.Wv;
      // the per-pixel scale is issued before the waits so that its latency hides under them
      float ps = 0.f;
      if (pixscale1 && valid) ps = __ldg(p.pixscale + (static_cast<long long>(b_img) * p.OH + hv) * p.OW + wv);
      if (need0) hy::mbar_wait(A_FULL + 8 * stage, par);  // acquire the TMA-written epilogue operands
      if (row == 0) RES_STAMP(it, 4);  // epilogue group at the loop top
      hy::mbar_wait(ACC_FULL + 8 * grp, (it >> 1) & 1);
      hy::tc_fence_after();
      if (row == 0) RES_STAMP(it, 5);  // accumulator seen

      const uint32_t acc0 = t_lane + grp * acc_cols;
      const int total = p.nphase * nchunk16;
      const int sk = p.stg_per_grp == 2 ? ((it >> 1) & 1) : 0;
      const uint32_t so_row = (p.decoupled ? stg_base + (grp * p.stg_per_grp + sk) * p.nchunk_out * 16384 : sb + p.o_off) + row * 128;
      const uint32_t sx_row = sb + p.x1_off + row * 128;
      // decoupled: the tile that last used this staging buffer must have left it (first use passes at once); with two
      // buffers per group that store was issued two tiles of the group ago
      if (p.decoupled && p.store_bf16)
        hy::mbar_wait(STG_FREE + 16 * grp + 8 * sk, (((it >> 1) / p.stg_per_grp) & 1) ^ 1u);
      if (epi == HYRES_EPI_STATS) {
        // channel mean / max over [s1 | up2(s2) | up4(s3)] of this thread's position: the 128 up-sampled channels sit
        // in the accumulator (fp32, as the reference interpolates), s1 in the activation tile
        float sum = 0.f, sum1 = 0.f, mx = -3.0e38f;
        uint32_t ra[16], rb[16];
        auto fold = [&](const uint32_t (&r)[16]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a = __uint_as_float(r[2 * i]), b = __uint_as_float(r[2 * i + 1]);
            hy::add2(sum, sum1, a, b);
            mx = fmaxf(mx, fmaxf(a, b));
          }
        };
        hy::tmem_ld16(acc0, ra);
        hy::tmem_ld_fence(ra);
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          hy::tmem_ld16(acc0 + (q + 1) * 16, rb);
          fold(ra);
          hy::tmem_ld_fence(rb);
          if (q + 2 < 8) hy::tmem_ld16(acc0 + (q + 2) * 16, ra);
          fold(rb);
          if (q + 2 < 8) hy::tmem_ld_fence(ra);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 w = lds128(sb + row * 128 + ((j ^ sw) << 4));
          const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float a = hy::bf16_lo(u[i]), b = hy::bf16_hi(u[i]);
            hy::add2(sum, sum1, a, b);
            mx = fmaxf(mx, fmaxf(a, b));
          }
        }
        if (valid)
          reinterpret_cast<float2*>(p.out_f32)[(static_cast<long long>(b_img) * p.OH + hv) * p.OW + wv] =
              make_float2((sum + sum1) * (1.f / 192.f), mx);
      }
      int ph = 0, n = 0;  // phase / first channel of the chunk, advanced incrementally (no division per chunk)
      auto chunk = [&](const uint32_t (&r)[16]) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        const uint32_t ba = bias_base + n * 4;
        if (epi == HYRES_EPI_PIXSCALE) {
          if (!pixscale1) {
            const long long opix = (static_cast<long long>(b_img) * p.OH + hv * p.out_mul + (ph >> 1)) * p.OW + wv * p.out_mul + (ph & 1);
            ps = valid ? __ldg(p.pixscale + opix) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = lds_f4(ba + i * 16);
            fma2(v[4 * i], v[4 * i + 1], ps, b4.x, b4.y);
            fma2(v[4 * i + 2], v[4 * i + 3], ps, b4.z, b4.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = lds_f4(ba + i * 16);
            hy::add2(v[4 * i], v[4 * i + 1], b4.x, b4.y);
            hy::add2(v[4 * i + 2], v[4 * i + 3], b4.z, b4.w);
          }
        }
        // smem address of this thread's 16 channels inside a [128 rows][64 ch] swizzled chunk
        const uint32_t coff = (n >> 6) * 16384;
        const uint32_t j0 = (n & 63) >> 3;
        const uint32_t c0 = coff + ((j0 ^ sw) << 4), c1 = coff + (((j0 + 1) ^ sw) << 4);
        const uint32_t oa0 = so_row + c0, oa1 = so_row + c1;
        if (need0) {
          float x[16];
          unpack16(lds128(oa0), lds128(oa1), x);
          if (epi == HYRES_EPI_ADD) {
#pragma unroll
            for (int i = 0; i < 8; ++i) hy::add2(v[2 * i], v[2 * i + 1], x[2 * i], x[2 * i + 1]);
          } else if (epi == HYRES_EPI_GATE) {
            float a[16];
            unpack16(lds128(sx_row + c0), lds128(sx_row + c1), a);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(a[i], hy::fast_sigmoid(v[i]), x[i]);
          } else if (epi == HYRES_EPI_GDN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = x[i] * hy::fast_rsqrt(v[i]);
          } else if (epi == HYRES_EPI_IGDN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = x[i] * hy::fast_sqrt(v[i]);
          }
        }
        bool relu_packed = false;
        if (act == HYRES_ACT_PRELU) {
          if (p.slope >= 0.f && p.slope <= 1.f) {
            // 0 <= slope <= 1: prelu(v) = max(v, slope * v)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float t0 = v[2 * i], t1 = v[2 * i + 1];
              mul2(t0, t1, p.slope);
              v[2 * i] = fmaxf(v[2 * i], t0);
              v[2 * i + 1] = fmaxf(v[2 * i + 1], t1);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] >= 0.f ? v[i] : v[i] * p.slope;
          }
        } else if (act == HYRES_ACT_RELU) {
          // ReLU commutes with the bf16 rounding: applied on the packed pairs when only bf16 leaves the kernel
          if (p.store_bf16 && !p.out_f32) {
            relu_packed = true;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
        } else if (act == HYRES_ACT_CLAMP01) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fminf(fmaxf(v[i], 0.f), 1.f);
        }
        if (p.store_bf16) {
          uint32_t u[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) u[i] = hy::pack_bf16(v[2 * i], v[2 * i + 1]);
          if (relu_packed) {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = hy::relu_bf16x2(u[i]);
          }
          sts128(oa0, make_uint4(u[0], u[1], u[2], u[3]));
          sts128(oa1, make_uint4(u[4], u[5], u[6], u[7]));
        }
        if (p.out_f32 && valid && n < p.cout) {
          const int oh = hv * p.out_mul + (ph >> 1), ow = wv * p.out_mul + (ph & 1);
          float* o = p.out_f32 + b_img * p.f32_sb + oh * p.f32_sh + ow * p.f32_sw + n * p.f32_sc;
          if (n + 16 <= p.cout && p.f32_sc == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (n + i < p.cout) o[i * p.f32_sc] = v[i];
          }
        }
        n += 16;
        if (n == p.BN) { n = 0; ++ph; }
      };
      // two register buffers: chunk q + 1 is in flight from TMEM while chunk q is processed (columns of successive
      // phases are contiguous)
      uint32_t ra[16], rb[16];
      if (epi != HYRES_EPI_STATS) {
      hy::tmem_ld16(acc0, ra);
      hy::tmem_ld_fence(ra);
      }
      for (int q = 0; q < (epi == HYRES_EPI_STATS ? 0 : total); q += 2) {
        if (q + 1 < total) hy::tmem_ld16(acc0 + (q + 1) * 16, rb);
        chunk(ra);
        if (q + 1 < total) {
          hy::tmem_ld_fence(rb);
          if (q + 2 < total) hy::tmem_ld16(acc0 + (q + 2) * 16, ra);
          chunk(rb);
          if (q + 2 < total) hy::tmem_ld_fence(ra);
        }
      }
      hy::tc_fence_before();
      hy::mbar_arrive(ACC_EMPTY + 8 * grp);
      if (row == 0) RES_STAMP(it, 6);  // chunks done
      if (p.store_bf16) hy::fence_async_smem();
      if (store_role) {  // this thread's part is staged and its operands are consumed
        if (it >= 2) hy::mbar_wait(STAGED_ACK + 8 * grp, ((it >> 1) - 1) & 1);  // the store warp took the previous tile
        hy::mbar_arrive(STAGED + 8 * grp);
      }
    }
  }

  hy::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// pitch_w / pitch_h: pixels per row / rows per image of the allocation when it is larger than the W x H view
// (a view into the interior of a padded tensor); 0 = dense.
int encode_map4(CUtensorMap* m, const void* ptr, int C, int ld, int B, int H, int W, int box_w, int box_h,
                int pitch_w = 0, int pitch_h = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return hy_fail(HYRES_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if (!pitch_w) pitch_w = W;
  if (!pitch_h) pitch_h = H;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)pitch_w * ld * 2, (cuuint64_t)pitch_h * pitch_w * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(act C=%d ld=%d B=%d H=%d W=%d box %dx%d) -> %d", C, ld, B, H, W,
             box_w, box_h, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

template <int EPI, int ACT>
cudaError_t launch_res(int grid, int threads, int smem, cudaStream_t stream, const ResParams& p) {
  static HyPerDevice attr;
  if (!attr.done()) {
    const cudaError_t e = cudaFuncSetAttribute(conv_res_kernel<EPI, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (e != cudaSuccess) return e;
    attr.mark();
  }
  return hy_launch_pdl(conv_res_kernel<EPI, ACT>, grid, threads, smem, stream, p);
}

// Interpolation matrices of the up-add mode, [256][64] bf16 on the device: rows 0..127 = U2, rows 128..255 = U4.
// Row = tile position ti * 8 + tj of a 16 x 8 tile.  U2 column = pixel pr * 6 + pc of the 10 x 6 (+ padding) patch of
// the half-resolution tensor whose origin is one pixel up-left of the tile's first source pixel; U4 column = pixel
// pr * 4 + pc of the 8 x 4 quarter-resolution patch.  Entries are the bilinear weights of
// F.interpolate(scale_factor = 2 / 4, mode = "bilinear", align_corners = False)
// (models/layers/enhancement.py:101-104): source coordinate (o + 0.5) / s - 0.5; the clamp at the image border is
// realised by the replicated one-pixel border of the padded source tensors.  All weights are j / 64: exact in bf16.
const void* up_table() {
  static void* tab[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return nullptr;
  if (tab[dev]) return tab[dev];
  std::vector<__nv_bfloat16> h(256 * 64, __float2bfloat16(0.f));
  for (int ti = 0; ti < kTH; ++ti)
    for (int tj = 0; tj < kTW; ++tj) {
      const int row = ti * kTW + tj;
      for (int s = 0; s < 2; ++s) {
        const float sc = s ? 4.f : 2.f;
        const int pw = s ? 4 : 6;
        const float py = (ti + 0.5f) / sc - 0.5f + 1.f, px = (tj + 0.5f) / sc - 0.5f + 1.f;  // patch origin = source - 1
        const int y0 = static_cast<int>(py), x0 = static_cast<int>(px);
        const float ly = py - y0, lx = px - x0;
        __nv_bfloat16* u = h.data() + (s * 128 + row) * 64;
        u[y0 * pw + x0] = __float2bfloat16((1.f - ly) * (1.f - lx));
        u[y0 * pw + x0 + 1] = __float2bfloat16((1.f - ly) * lx);
        u[(y0 + 1) * pw + x0] = __float2bfloat16(ly * (1.f - lx));
        u[(y0 + 1) * pw + x0 + 1] = __float2bfloat16(ly * lx);
      }
    }
  void* d = nullptr;
  if (cudaMalloc(&d, h.size() * 2) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  tab[dev] = d;
  return d;
}

// Launch with the epilogue variant of p.epi / p.act.
int res_dispatch(const ResParams& p, int smem, cudaStream_t stream) {
  const int grid = std::min(p.ntiles, p.cta_limit > 0 ? std::min(p.cta_limit, num_sms()) : num_sms());
  const int threads = p.a_square ? kThreadsSq : kThreads;
  hy_count_launch();
  // the combinations the codec uses are compiled straight-line; anything else runs the generic instance
#define RES_CASE(E, A)                                                \
  if (p.epi == (E) && p.act == (A)) {                                 \
    HY_CUDA((launch_res<E, A>(grid, threads, smem, stream, p)));      \
    return HYRES_OK;                                                  \
  }
  RES_CASE(HYRES_EPI_LINEAR, HYRES_ACT_NONE)
  RES_CASE(HYRES_EPI_LINEAR, HYRES_ACT_RELU)
  RES_CASE(HYRES_EPI_LINEAR, HYRES_ACT_PRELU)
  RES_CASE(HYRES_EPI_ADD, HYRES_ACT_NONE)
  RES_CASE(HYRES_EPI_ADD, HYRES_ACT_RELU)
  RES_CASE(HYRES_EPI_GATE, HYRES_ACT_NONE)
  RES_CASE(HYRES_EPI_GDN, HYRES_ACT_NONE)
  RES_CASE(HYRES_EPI_IGDN, HYRES_ACT_NONE)
  RES_CASE(HYRES_EPI_PIXSCALE, HYRES_ACT_PRELU)
  RES_CASE(HYRES_EPI_STATS, HYRES_ACT_NONE)
#undef RES_CASE
  HY_CUDA((launch_res<-1, -1>(grid, threads, smem, stream, p)));
  return HYRES_OK;
}

}  // namespace

// Runs the layer on the resident-weights kernel when it qualifies (*handled = 1); otherwise
// leaves *handled = 0 and the caller falls through to the streaming kernel.
int conv_res_try_run(hyres_conv* c, const hyres_conv_io* io, cudaStream_t stream, int* handled) {
  *handled = 0;
  const bool deconv = c->kind == HYRES_DECONV_K5S2;
  if (!deconv && (c->stride != 1 || c->cin1 != 0)) return HYRES_OK;
  if (io->out_sq) return HYRES_OK;
  if (deconv && io->out_bf16) return HYRES_OK;  // interleaved bf16 stores stay on the streaming kernel
  if (io->ld_x0 && io->ld_x0 < c->cin0) return hy_fail(HYRES_ERR_ARG, "conv_run: ld_x0 smaller than the channel count");
  const int BN = c->cout_pad;
  if (BN > 256 || (BN % 16) || 2 * c->nphase * BN > 512) return HYRES_OK;
  const int nslots = c->ktot / 64;
  const long long w_bytes = static_cast<long long>(nslots) * BN * 128;
  if (w_bytes > 150 * 1024) return HYRES_OK;

  ResParams p;
  memset(&p, 0, sizeof p);
  // geometry of the shared patch
  int org_h, org_w, PH, PW;
  if (deconv) {
    org_h = org_w = -1; PH = kTH + 2; PW = kTW + 2;
  } else {
    org_h = org_w = -c->pad; PH = kTH + (c->R - 1) * c->dil; PW = kTW + (c->S - 1) * c->dil;
  }
  if (PW > 256 || PH > 256) return HYRES_OK;
  // tap list from the layer's plan
  struct { int phase, chunk, a_row, slot; } steps_tmp[kMaxSteps];
  int ns = 0;
  for (int ph = 0; ph < c->nphase; ++ph)
    for (int g = c->ph_begin[ph]; g < c->ph_begin[ph] + c->ph_count[ph]; ++g) {
      const TapGroup& tg = c->groups[g];
      for (int t = 0; t < tg.ntaps; ++t) {
        if (ns >= kMaxSteps) return HYRES_OK;
        const int ro = tg.dh + tg.tap_row[t] - org_h, co = tg.dw - org_w;
        if (ro < 0 || co < 0 || ro + kTH > PH || co + kTW > PW) return hy_fail(HYRES_ERR_STATE, "conv_res: tap outside patch");
        steps_tmp[ns].phase = ph;
        steps_tmp[ns].chunk = tg.c_off / 64;
        steps_tmp[ns].a_row = ro * PW + co;
        steps_tmp[ns].slot = tg.kslot0 + t;
        ++ns;
      }
    }
  p.nsteps = ns;
  p.nphase = c->nphase;
  p.nchunk_in = (c->cin0 + 63) / 64;
  p.nchunk_out = (c->cout + 63) / 64;
  p.PW = PW; p.PH = PH; p.org_h = org_h; p.org_w = org_w;
  p.a_chunk_bytes = (PH * PW * 128 + 1023) / 1024 * 1024;
  {
    uint32_t started = 0;
    for (int i = 0; i < ns; ++i) {
      p.steps[i].a_off16 = (steps_tmp[i].chunk * p.a_chunk_bytes + steps_tmp[i].a_row * 128) >> 4;
      p.steps[i].b_off16 = (steps_tmp[i].slot * BN * 128) >> 4;
      p.steps[i].d_col = steps_tmp[i].phase * BN;
      p.steps[i].acc = (started >> steps_tmp[i].phase) & 1u;
      started |= 1u << steps_tmp[i].phase;
    }
  }
  const bool need0 = io->epi == HYRES_EPI_ADD || io->epi == HYRES_EPI_GATE || io->epi == HYRES_EPI_GDN ||
                     io->epi == HYRES_EPI_IGDN;
  const bool need1 = io->epi == HYRES_EPI_GATE;
  p.has_aux0 = need0; p.has_aux1 = need1;
  p.store_bf16 = io->out_bf16 ? 1 : 0;
  p.a_square = io->x0_square ? 1 : 0;
  {
    // HYRES_RES_TRACE=<device pointer, hex>: 64 x 16 clock64 stamps of CTA 0 (tools/experiments/res_trace.py)
    const char* e = getenv("HYRES_RES_TRACE");
    p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
  if (p.a_square && (PH != kTH || PW != kTW)) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: x0_square needs a 1x1 layer");
  const bool up = io->up_t2 != nullptr || io->up_t3 != nullptr;
  if (up) {
    if (!io->up_t2 || !io->up_t3) return hy_fail(HYRES_ERR_ARG, "conv_run: up-add needs both up_t2 and up_t3");
    if (deconv || BN != 64 || c->nphase != 1 || (io->H % 32) || (io->W % 32))
      return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: up-add needs a stride-1 layer with 64 output channels and H, W multiples of 32");
  }
  if (io->out_pad < 0 || io->out_pad > 8 || (io->out_pad && (!io->out_bf16 || deconv)))
    return hy_fail(HYRES_ERR_ARG, "conv_run: out_pad needs a bf16 output of a stride-1 layer");
  if (deconv && (need0 || need1)) return HYRES_OK;
  // Without epilogue operands the result staging lives outside the ring (one buffer per epilogue group): the MMA
  // commit alone frees a patch, so the ring runs several tiles ahead of the epilogues instead of waiting for the
  // TMA store of the tile that last used the stage (measured: the store's read adds ~1000 clk to a stage's life).
  p.decoupled = (!need0 && !need1 && p.store_bf16) ? 1 : 0;
  int stage = p.nchunk_in * p.a_chunk_bytes;
  p.o_off = stage;
  if ((need0 || p.store_bf16) && !p.decoupled) stage += p.nchunk_out * 16384;
  p.x1_off = stage;
  if (need1) stage += p.nchunk_out * 16384;
  constexpr int kT2Rows = 6 * 11, kT3Rows = 4 * 8;  // 10 x 6 live pixels + 4 padding rows of the K = 64 GEMM; 6 x 4 live
  p.up = up;
  p.u_bytes = up ? 32768 : 0;
  p.t2_off = stage;
  if (up) stage += (kT2Rows * 128 + 1023) / 1024 * 1024;
  p.t3_off = stage;
  if (up) stage += kT3Rows * 128;
  p.stage_bytes = stage;
  p.stage_tx_bytes = p.nchunk_in * PH * PW * 128 + (need0 ? p.nchunk_out * 16384 : 0) + (need1 ? p.nchunk_out * 16384 : 0) +
                     (up ? (kT2Rows + kT3Rows) * 128 : 0);
  const int fixed0 = static_cast<int>(w_bytes) + p.u_bytes + 1024 /*bias*/ + 512 /*barriers*/ + 1024 /*alignment*/;
  // two staging buffers per group when the ring keeps >= 3 stages next to them: the group then never waits for the
  // read of its previous store (~1000 clk)
  p.stg_per_grp = 1;
  if (p.decoupled && (kSmemLimit - fixed0 - 4 * p.nchunk_out * 16384) / stage >= 3) p.stg_per_grp = 2;
  const int stg_bytes = p.decoupled ? 2 * p.stg_per_grp * p.nchunk_out * 16384 : 0;
  const int fixed = fixed0 + stg_bytes;
  int NA = (kSmemLimit - fixed) / stage;
  // a single stage (no load / compute overlap) is still accepted for the gate at C = 192: its three operand tiles
  // (144 KB) leave room for one stage only, and the streaming kernel's per-thread global loads of the two epilogue
  // operands are far slower (measured 115 us against ~35 us for 16x64x96x192)
  static const bool allow1 = getenv("HYRES_RES_NO_NA1") == nullptr;
  if (NA < 2 && !(NA == 1 && need1 && allow1)) return HYRES_OK;
  NA = std::min(NA, kMaxStages);
  p.NA = NA;
  {
    // L2 prefetch of the tiles beyond the ring pays only when the ring is shallow (gate: two stages, GDN: three); with a deep
    // ring the prefetch operations queue in front of the loads the MMAs are waiting for (measured: 1x1 64->64 at
    // 16x512x768 0.296 -> 0.259 ms, channel statistics 0.321 -> 0.242 ms without it).  HYRES_RES_PF overrides.
    static const char* e = getenv("HYRES_RES_PF");
    p.pf_extra = e ? atoi(e) : (NA <= 3 ? 0 : -1);
  }
  p.stg_base_off = NA * stage;
  p.w_bytes = static_cast<int>(w_bytes);
  p.nslots = nslots;
  p.BN = BN;
  int cols = 32;
  while (cols < 2 * c->nphase * BN) cols <<= 1;
  p.tmem_cols = cols;

  int OH, OW;
  hyres_conv_out_size(c, io->H, io->W, &OH, &OW);
  p.OH = OH; p.OW = OW;
  p.out_mul = deconv ? 2 : 1;
  p.Hv = OH / p.out_mul; p.Wv = OW / p.out_mul;
  p.tiles_w = (p.Wv + kTW - 1) / kTW;
  p.tiles_per_img = p.tiles_w * ((p.Hv + kTH - 1) / kTH);
  p.ntiles = p.tiles_per_img * io->B;
  p.cout = c->cout;
  p.epi = io->epi; p.act = io->act; p.slope = io->slope;
  p.bias = c->d_bias;
  p.cta_limit = io->cta_limit;
  p.pixscale = io->pixscale;
  p.out_f32 = io->out_f32;
  p.f32_sb = io->f32_sb; p.f32_sh = io->f32_sh; p.f32_sw = io->f32_sw; p.f32_sc = io->f32_sc;

  const int ld_x0 = io->ld_x0 ? io->ld_x0 : c->cin0;
  int rc = encode_map4(&p.mapA, io->x0, c->cin0, ld_x0, io->B, io->H, io->W, PW, PH);
  if (rc != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapW, c->d_w, c->ktot, c->cout_pad, BN)) != HYRES_OK) return rc;
  if (p.store_bf16) {
    // out_pad: the output view is the interior of a tensor padded by out_pad pixels on every side
    const int op = io->out_pad;
    const uint8_t* optr = static_cast<const uint8_t*>(io->out_bf16) + (static_cast<int64_t>(op) * (OW + 2 * op) + op) * io->ld_out * 2;
    if ((rc = encode_map4(&p.mapOut, optr, c->cout, io->ld_out, io->B, OH, OW, kTW, kTH, OW + 2 * op, OH + 2 * op)) != HYRES_OK) return rc;
  }
  if (up) {
    const void* tab = up_table();
    if (!tab) return hy_fail(HYRES_ERR_CUDA, "conv_run: cannot allocate the interpolation table");
    if ((rc = encode_map4(&p.mapT2, io->up_t2, 64, 64, io->B, OH / 2 + 2, OW / 2 + 2, 6, 11)) != HYRES_OK) return rc;
    if ((rc = encode_map4(&p.mapT3, io->up_t3, 64, 64, io->B, OH / 4 + 2, OW / 4 + 2, 4, 8)) != HYRES_OK) return rc;
    if ((rc = encode_w_map(&p.mapU, tab, 64, 256, 128)) != HYRES_OK) return rc;
  }
  if (need0 && (rc = encode_map4(&p.mapAux0, io->aux0, c->cout, io->ld_aux0, io->B, OH, OW, kTW, kTH)) != HYRES_OK) return rc;
  if (need1 && (rc = encode_map4(&p.mapAux1, io->aux1, c->cout, io->ld_aux1, io->B, OH, OW, kTW, kTH)) != HYRES_OK) return rc;

  const int rc2 = res_dispatch(p, fixed + NA * stage, stream);
  if (rc2 != HYRES_OK) return rc2;
  *handled = 1;
  return HYRES_OK;
}

// Channel mean / max over the virtual concat [f1 | up2(s2) | up4(s3)] on the resident-weights kernel in its
// statistics mode: the two bilinear up-samplings are tensor-core GEMMs (interpolation matrix x TMA patch), the
// epilogue folds the 128 accumulator columns and the 64 channels of the f1 tile of its position.
extern "C" int hyres_refine_stats3_tc(const void* f1, const void* s2_padded, const void* s3_padded, float* stats, int B,
                                      int H, int W, void* stream_v) {
  if (!f1 || !s2_padded || !s3_padded || !stats || B <= 0) return hy_fail(HYRES_ERR_ARG, "stats3_tc: bad argument");
  if ((H % 32) || (W % 32)) return hy_fail(HYRES_ERR_UNSUPPORTED, "stats3_tc: H and W must be multiples of 32");
  const void* tab = up_table();
  if (!tab) return hy_fail(HYRES_ERR_CUDA, "stats3_tc: cannot allocate the interpolation table");
  ResParams p;
  memset(&p, 0, sizeof p);
  constexpr int kT2Rows = 6 * 11, kT3Rows = 4 * 8;
  p.nsteps = 0; p.nphase = 1; p.nchunk_in = 1; p.nchunk_out = 0;
  p.PW = kTW; p.PH = kTH; p.org_h = p.org_w = 0;
  p.a_chunk_bytes = kTH * kTW * 128;
  p.up = 2; p.u_bytes = 32768;
  p.o_off = p.x1_off = p.a_chunk_bytes;
  p.t2_off = p.a_chunk_bytes;
  p.t3_off = p.t2_off + (kT2Rows * 128 + 1023) / 1024 * 1024;
  p.stage_bytes = p.t3_off + kT3Rows * 128;
  p.stage_tx_bytes = p.a_chunk_bytes + (kT2Rows + kT3Rows) * 128;
  const int fixed = p.u_bytes + 1024 /*bias*/ + 512 /*barriers*/ + 1024 /*alignment*/;
  p.NA = std::min((kSmemLimit - fixed) / p.stage_bytes, kMaxStages);
  p.stg_base_off = p.NA * p.stage_bytes;
  p.stg_per_grp = 1;
  p.BN = 128; p.tmem_cols = 256;
  p.Hv = p.OH = H; p.Wv = p.OW = W; p.out_mul = 1;
  p.tiles_w = W / kTW; p.tiles_per_img = p.tiles_w * (H / kTH); p.ntiles = p.tiles_per_img * B;
  p.epi = HYRES_EPI_STATS; p.act = HYRES_ACT_NONE;
  p.out_f32 = stats;
  {
    static const char* e = getenv("HYRES_RES_PF");
    p.pf_extra = e ? atoi(e) : -1;  // deep ring: no L2 prefetch (see conv_res_try_run)
    const char* tr = getenv("HYRES_RES_TRACE");
    p.trace = tr ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 16)) : nullptr;
  }
  int rc = encode_map4(&p.mapA, f1, 64, 64, B, H, W, kTW, kTH);
  if (rc != HYRES_OK) return rc;
  if ((rc = encode_map4(&p.mapT2, s2_padded, 64, 64, B, H / 2 + 2, W / 2 + 2, 6, 11)) != HYRES_OK) return rc;
  if ((rc = encode_map4(&p.mapT3, s3_padded, 64, 64, B, H / 4 + 2, W / 4 + 2, 4, 8)) != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapU, tab, 64, 256, 128)) != HYRES_OK) return rc;
  return res_dispatch(p, fixed + p.NA * p.stage_bytes, static_cast<cudaStream_t>(stream_v));
}
