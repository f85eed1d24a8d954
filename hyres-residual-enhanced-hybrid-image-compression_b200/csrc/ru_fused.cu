// Fused bottleneck residual unit on Blackwell tensor cores:
//
//     out = [ReLU]( x + conv1x1_{64->128}( ReLU( conv3x3_{64->64}( ReLU( conv1x1_{128->64}(x) ))))) )
//
// i.e. the reference's ResidualUnit (models/layers/attention.py:16-33, final ReLU) and
// compressai's ResidualBottleneckBlock (models/checkerboard.py:38,42,51,55, no final ReLU)
// at C = 128.  These 16 blocks are 40 % of the MACs of the codec and, launched as three
// convolutions each, are bound by HBM round trips of the two 64-channel intermediates and by
// per-tile latency.  Here one persistent CTA per SM keeps all three weight matrices (104 KB)
// resident in shared memory and walks 16x8 output tiles:
//
//   TMA   x halo patch 18x10 positions x 128 ch (two 64-ch SWIZZLE_128B chunks, image borders
//         zero-filled by the TMA unit), double buffered;
//   G1    tcgen05.mma  [180(->256) x 128] . W1^T -> TMEM (2 x 64 columns)
//   E1    TMEM -> +bias, ReLU, zero outside the image, bf16 -> smem patch t1 [180][64] (swizzled)
//   G2    nine taps, each ONE descriptor on the same t1 patch: start row (r*10+s), 8-row
//         groups 10 rows (1280 B) apart -> TMEM (64 columns).  No im2col, no duplicate of t1.
//   E2    TMEM -> +bias, ReLU, bf16 -> smem t2 [128][64]
//   G3    tcgen05.mma  [128 x 64] . W3^T -> TMEM (128 columns)
//   E3    TMEM -> +bias + skip (centre of the x patch, still in smem) [ReLU] -> bf16 staged
//         over the x buffer -> TMA store (clipped at the image edge by the TMA unit).
//
// HBM traffic is the algorithmic minimum (read x once + halo, write out once); the next
// tile's TMA load and the previous tile's TMA store overlap the three GEMMs.
//
// Eight epilogue warps: warp w reads TMEM lane quadrant (w & 3) and column half (w >> 2), so each
// epilogue stage is half as long as with four warps.  The three GEMMs use disjoint TMEM column
// ranges (G1 [0,128), G2 [128,192), G3 [256,384)), so G1 of the next tile is issued between G2 and
// G3 of the current one and runs while the epilogue warps are busy with E2 / E3.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

#define RU_STAMP(slot) do { if (p.trace && blockIdx.x == 0 && it < 64) p.trace[it * 16 + (slot)] = clock64(); } while (0)

namespace {

constexpr int kTH = 16, kTW = 8;        // output tile
constexpr int kPW = 10;                 // patch width (positions)
constexpr int kNP = 180;                // patch positions (18 x 10)
constexpr uint32_t kColG1 = 0, kColG2 = 128, kColT2 = 192, kColG3 = 256, kTmemCols = 512;

// shared-memory map (bytes from the 1024-aligned base)
constexpr uint32_t kW1 = 0;             // [2 k-chunks][64 n][128 B]
constexpr uint32_t kW2 = 16384;         // [9 taps][64 n][128 B]
constexpr uint32_t kW3 = 90112;         // [128 n][128 B]
constexpr uint32_t kWBytes = 106496;
constexpr uint32_t kXChunk = 23552;     // 180 rows x 128 B, padded to a multiple of 1024
constexpr uint32_t kXBytes = kNP * 128; // bytes one TMA chunk load delivers
constexpr uint32_t kX = kWBytes;                 // x halo patch: 2 chunks (one buffer: it only feeds G1)
constexpr uint32_t kT1 = kX + 2 * kXChunk;       // t1 patch [180][64]
constexpr uint32_t kT2 = kT1 + kXChunk;          // (unused since t2 lives in tensor memory; kept so the layout below is unchanged)
constexpr uint32_t kStg = kT2 + 16384;           // skip tile in, result out: 2 chunks [128][64]
constexpr uint32_t kBars = kStg + 32768;
constexpr uint32_t kBias = kBars + 256;
constexpr uint32_t kSmemUsed = kBias + 1024;

enum { W_FULL = 0, X_FULL, X_FREE, G1_DONE, T1_READY, G2_DONE, T2_READY, G3_DONE, ACC3_FREE, SKIP_FULL, STAGED, STG_FREE,
       NUM_BARS };

struct alignas(64) RuParams {
  CUtensorMap mapX, mapSkip, mapOut, mapW1, mapW2, mapW3;
  // b1 [64] | b2 [64] | b3 [128] as launch parameters: the epilogues read them through the constant cache with
  // warp-uniform indices, not through shared memory (whose 128 B/clk port bounds this kernel: the broadcast bias
  // loads were ~290 wavefronts of a tile's ~4600)
  float bias[256];
  int32_t H, W, tiles_w, tiles_per_img, ntiles, final_relu;
  long long* trace;  // optional: clock64 stamps of CTA 0 (tools/experiments), 16 slots per tile
};

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// 16 accumulator columns -> (+bias) ReLU -> 16 bf16 in q[0..1]; `keep` = false zeroes the row.
__device__ __forceinline__ void bias_relu_pack16(uint32_t (&r)[16], const float (&b)[16], bool keep, uint4 (&q)[2]) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) hy::add2(v[2 * i], v[2 * i + 1], b[g * 8 + 2 * i], b[g * 8 + 2 * i + 1]);
    q[g].x = hy::relu_bf16x2(hy::pack_bf16(v[0], v[1]));
    q[g].y = hy::relu_bf16x2(hy::pack_bf16(v[2], v[3]));
    q[g].z = hy::relu_bf16x2(hy::pack_bf16(v[4], v[5]));
    q[g].w = hy::relu_bf16x2(hy::pack_bf16(v[6], v[7]));
    if (!keep) q[g] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Software pipeline across tiles (i = this CTA's i-th tile):
//   epilogue warps :  E1(i)  E3(i-1)  E2(i)   E1(i+1)  E3(i)  E2(i+1) ...
//   tensor pipe    :         G2(i)  G1(i+1)   G3(i)           G2(i+1) ...
// The 3x3 GEMM of tile i and the first GEMM of tile i+1 run under the (long) output epilogue of tile i-1, and
// G3(i) under E1(i+1); every buffer has one producer and one consumer phase per tile, tracked by an mbarrier.
// CS: column splits = epilogue warps per TMEM lane quadrant (2 -> 8 epilogue warps, 4 -> 16).
template <int CS>
__global__ void __launch_bounds__(128 * CS + 96, 1) ru_fused_kernel(const __grid_constant__ RuParams p) {
  constexpr int kEpiThreads = 128 * CS;
  constexpr int kThreads = kEpiThreads + 96;
  constexpr int kWarpLoad = 4 * CS, kWarpMma = 4 * CS + 1, kWarpStore = 4 * CS + 2;  // one warp each
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  auto bar = [&](int i) { return base + kBars + 8u * i; };
  const uint32_t tmem_slot = base + kBars + 128;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction: uniform role branches
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    hy::mbar_init(bar(W_FULL), 1);
    hy::mbar_init(bar(X_FULL), 1);
    hy::mbar_init(bar(X_FREE), 1);
    hy::mbar_init(bar(G1_DONE), 1);
    hy::mbar_init(bar(G2_DONE), 1);
    hy::mbar_init(bar(G3_DONE), 1);
    hy::mbar_init(bar(SKIP_FULL), 1);
    hy::mbar_init(bar(STG_FREE), 1);
    hy::mbar_init(bar(T1_READY), kEpiThreads);
    hy::mbar_init(bar(T2_READY), kEpiThreads);
    hy::mbar_init(bar(ACC3_FREE), kEpiThreads);
    hy::mbar_init(bar(STAGED), kEpiThreads);
    hy::mbar_fence_init();
    // the weights do not depend on the predecessor kernel: their load starts before the dependency wait
    hy::mbar_arrive_expect_tx(bar(W_FULL), kWBytes);
    hy::tma_load_2d(base + kW1, &p.mapW1, bar(W_FULL), 0, 0);
    hy::tma_load_2d(base + kW1 + 8192, &p.mapW1, bar(W_FULL), 64, 0);
    for (int s = 0; s < 9; ++s) hy::tma_load_2d(base + kW2 + s * 8192, &p.mapW2, bar(W_FULL), s * 64, 0);
    hy::tma_load_2d(base + kW3, &p.mapW3, bar(W_FULL), 0, 0);
  }
  if (warp == kWarpLoad && lane == 0) {
    hy::tma_prefetch_desc(&p.mapX);
    hy::tma_prefetch_desc(&p.mapSkip);
    hy::tma_prefetch_desc(&p.mapW1);
    hy::tma_prefetch_desc(&p.mapW2);
    hy::tma_prefetch_desc(&p.mapW3);
  }
  if (warp == kWarpStore && lane == 0) hy::tma_prefetch_desc(&p.mapOut);
  if (warp == kWarpMma) {
    hy::tmem_alloc(tmem_slot, kTmemCols);
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
  const int stride = static_cast<int>(gridDim.x);
  const int first = static_cast<int>(blockIdx.x);

  if (warp == kWarpLoad) {
    // ============================ TMA load producer ============================
    if (lane == 0) {
      auto load_x = [&](int t) {
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        hy::mbar_arrive_expect_tx(bar(X_FULL), 2 * kXBytes);
        hy::tma_load_4d(base + kX, &p.mapX, bar(X_FULL), 0, w0 - 1, h0 - 1, b_img);
        hy::tma_load_4d(base + kX + kXChunk, &p.mapX, bar(X_FULL), 64, w0 - 1, h0 - 1, b_img);
      };
      if (first < p.ntiles) load_x(first);
      int it = 0;
      for (int t = first; t < p.ntiles; t += stride, ++it) {
        // x(i+1): the patch buffer only feeds G1, so it is free as soon as G1(i) has been executed
        if (t + stride < p.ntiles) {
          hy::mbar_wait(bar(X_FREE), it & 1);
          load_x(t + stride);
        }
        // skip(i): the centre of x(i) again (an L2 hit), into the staging tile the result will leave from
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        if (it > 0) hy::mbar_wait(bar(STG_FREE), (it - 1) & 1);
        hy::mbar_arrive_expect_tx(bar(SKIP_FULL), 32768);
        hy::tma_load_4d(base + kStg, &p.mapSkip, bar(SKIP_FULL), 0, w0, h0, b_img);
        hy::tma_load_4d(base + kStg + 16384, &p.mapSkip, bar(SKIP_FULL), 64, w0, h0, b_img);
      }
    }
  } else if (warp == kWarpMma) {
    // ============================ MMA issuer ============================
    // The whole warp runs the loop converged (descriptors stay in uniform registers); one elected lane issues.
    {
      const uint32_t leader = hy::elect_leader();
      const uint32_t idesc64 = hy::umma_idesc_bf16(128, 64);
      const uint32_t idesc128 = hy::umma_idesc_bf16(128, 128);
      const uint64_t w1_d = hy::desc_u64(base + kW1), w2_d = hy::desc_u64(base + kW2), w3_d = hy::desc_u64(base + kW3);
      const uint64_t x_d = hy::desc_u64(base + kX), t1_d = hy::desc_u64(base + kT1, kPW * 128), t2_d = hy::desc_u64(base + kT2);
      // G1: t1 = x . W1^T over the 180-position patch (2 blocks of 128 rows; rows >= 180 unused)
      auto issue_g1 = [&]() {
        hy::tc_fence_after();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              hy::umma_issue<2>(tmem_base + kColG1 + blk * 64, x_d + ((kc * kXChunk + blk * 16384 + k * 32) >> 4),
                                w1_d + ((kc * 8192 + k * 32) >> 4), idesc64, (kc | k) ? 1u : 0u, leader);
        hy::umma_commit_mode<2>(bar(G1_DONE), leader);
        hy::umma_commit_mode<2>(bar(X_FREE), leader);
      };
      hy::mbar_wait(bar(W_FULL), 0);
      int it = 0;
      if (first < p.ntiles) {
        hy::mbar_wait(bar(X_FULL), 0);
        issue_g1();
      }
      for (int t = first; t < p.ntiles; t += stride, ++it) {
        const uint32_t ph = it & 1;
        const bool has_next = t + stride < p.ntiles;
        // G2: 3x3 over the t1 patch in smem; tap (r,s) = start row r*10+s, row groups 1280 B apart
        hy::mbar_wait(bar(T1_READY), ph);
        hy::tc_fence_after();
        if (lane == 0) RU_STAMP(8);
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              hy::umma_issue<2>(tmem_base + kColG2, t1_d + (((r * kPW + s) * 128 + k * 32) >> 4),
                                w2_d + (((s * 3 + r) * 8192 + k * 32) >> 4), idesc64, (s | r | k) ? 1u : 0u, leader);
        hy::umma_commit_mode<2>(bar(G2_DONE), leader);
        if (lane == 0) RU_STAMP(9);
        // G1 of the next tile rides behind G2 (its columns were drained by E1 of this tile) and runs under E3 / E2
        bool g1_ahead = false;
        if (has_next && __all_sync(0xffffffffu, hy::mbar_try_wait(bar(X_FULL), (it + 1) & 1))) {
          issue_g1();
          g1_ahead = true;
        }
        // G3: 1x1 expand of t2
        hy::mbar_wait(bar(T2_READY), ph);
        if (it > 0) hy::mbar_wait(bar(ACC3_FREE), (it - 1) & 1);  // E3 of the previous tile drained [256,384)
        hy::tc_fence_after();
        if (lane == 0) RU_STAMP(10);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          hy::umma_ts_issue(tmem_base + kColG3, tmem_base + kColT2 + k * 8, w3_d + ((k * 32) >> 4), idesc128, k ? 1u : 0u, leader);
        hy::umma_commit_mode<2>(bar(G3_DONE), leader);
        if (lane == 0) RU_STAMP(11);
        if (has_next && !g1_ahead) {
          hy::mbar_wait(bar(X_FULL), (it + 1) & 1);
          issue_g1();
        }
      }
    }
  } else if (warp == kWarpStore) {
    // ============================ TMA store ============================
    if (lane == 0) {
      int it = 0;
      for (int t = first; t < p.ntiles; t += stride, ++it) {
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        hy::mbar_wait(bar(STAGED), it & 1);
        hy::tma_store_4d(&p.mapOut, base + kStg, 0, w0, h0, b_img);
        hy::tma_store_4d(&p.mapOut, base + kStg + 16384, 64, w0, h0, b_img);
        hy::tma_store_commit();
        hy::tma_store_wait_read<0>();           // the staging tile may be refilled with the next skip
        hy::mbar_arrive(bar(STG_FREE));
      }
      hy::tma_store_wait_all<0>();
    }
  } else {
    // ============================ epilogue warps ============================
    // warp w: TMEM lane quadrant (w & 3), column split csel = w >> 2 of CS
    constexpr int U12 = 64 / CS / 16;   // 16-column units of the 64-wide t1 / t2 rows per warp: 2 (CS=2), 1 (CS=4)
    constexpr int U3 = 128 / CS / 16;   // 16-column units of the 128-wide output row per warp: 4, 2
    const int q = warp & 3;
    const int csel = warp >> 2;
    const int tid = q * 32 + lane; // TMEM lane == GEMM row
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    auto load_bias16 = [&](int off, float (&b)[16]) {
#pragma unroll
      for (int i = 0; i < 16; ++i) b[i] = p.bias[off + i];
    };

    // E3: + bias + skip, in place in the staging tile; this warp: columns [csel*128/CS, +128/CS)
    auto epilogue3 = [&](int it) {
      const uint32_t ph = it & 1;
      uint32_t r[U3][16];
      hy::mbar_wait(bar(G3_DONE), ph);
      hy::tc_fence_after();
      if (threadIdx.x == 0) RU_STAMP(5);
      const int c0 = csel * (128 / CS);  // first output channel of this warp
#pragma unroll
      for (int u = 0; u < U3; ++u) hy::tmem_ld16(t_lane + kColG3 + c0 + u * 16, r[u]);
      hy::mbar_wait(bar(SKIP_FULL), ph);
      const uint32_t row = base + kStg + (c0 >> 6) * 16384 + tid * 128;
      const uint32_t sw = tid & 7;
      const int j0 = (c0 & 63) >> 3;   // first 16-byte chunk of this warp inside the 64-channel row
      uint4 sk[2 * U3];
#pragma unroll
      for (int j = 0; j < 2 * U3; ++j) sk[j] = lds128(row + (((j0 + j) ^ sw) << 4));
#pragma unroll
      for (int u = 0; u < U3; ++u) hy::tmem_ld_fence(r[u]);  // the first waits; all tie the registers to the wait
#pragma unroll
      for (int u = 0; u < U3; ++u) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int j = 2 * u + g;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[u][g * 8 + i]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            hy::add2(v[2 * i], v[2 * i + 1], p.bias[128 + c0 + j * 8 + 2 * i], p.bias[128 + c0 + j * 8 + 2 * i + 1]);
          hy::add2(v[0], v[1], hy::bf16_lo(sk[j].x), hy::bf16_hi(sk[j].x));
          hy::add2(v[2], v[3], hy::bf16_lo(sk[j].y), hy::bf16_hi(sk[j].y));
          hy::add2(v[4], v[5], hy::bf16_lo(sk[j].z), hy::bf16_hi(sk[j].z));
          hy::add2(v[6], v[7], hy::bf16_lo(sk[j].w), hy::bf16_hi(sk[j].w));
          uint4 o;
          o.x = hy::pack_bf16(v[0], v[1]);
          o.y = hy::pack_bf16(v[2], v[3]);
          o.z = hy::pack_bf16(v[4], v[5]);
          o.w = hy::pack_bf16(v[6], v[7]);
          if (p.final_relu) {
            o.x = hy::relu_bf16x2(o.x); o.y = hy::relu_bf16x2(o.y); o.z = hy::relu_bf16x2(o.z); o.w = hy::relu_bf16x2(o.w);
          }
          sts128(row + (((j0 + j) ^ sw) << 4), o);
        }
      }
      hy::tc_fence_before();
      hy::mbar_arrive(bar(ACC3_FREE));  // TMEM columns [256,384) may be overwritten by the next G3
      hy::fence_async_smem();
      hy::mbar_arrive(bar(STAGED));
      if (threadIdx.x == 0) RU_STAMP(6);
    };

    int it = 0;
    for (int t = first; t < p.ntiles; t += stride, ++it) {
      const uint32_t ph = it & 1;
      int b_img, h0, w0;
      tile_origin(t, b_img, h0, w0);
      uint4 qv[2];

      // ---- E1: t1 patch (this warp: 64/CS of the 64 channels); the t1 buffer was released by G2_DONE(i-1) ----
      if (threadIdx.x == 0) RU_STAMP(0);
      hy::mbar_wait(bar(G1_DONE), ph);
      hy::tc_fence_after();
      if (threadIdx.x == 0) RU_STAMP(1);
      {
        uint32_t ra[U12][16], rb[U12][16];
#pragma unroll
        for (int u = 0; u < U12; ++u) hy::tmem_ld16(t_lane + kColG1 + csel * (64 / CS) + u * 16, ra[u]);
        if (q < 2) {  // rows 128.. exist in lanes 0..51 only
#pragma unroll
          for (int u = 0; u < U12; ++u) hy::tmem_ld16(t_lane + kColG1 + 64 + csel * (64 / CS) + u * 16, rb[u]);
        }
#pragma unroll
        for (int u = 0; u < U12; ++u) {
          hy::tmem_ld_fence(ra[u]);
          if (q < 2) hy::tmem_ld_fence(rb[u]);
        }
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          if (blk == 1 && q >= 2) break;  // warp-uniform
          const int pp = blk * 128 + tid;
          const int pr = pp / kPW, pq = pp - pr * kPW;
          const int hh = h0 - 1 + pr, ww = w0 - 1 + pq;
          const bool live = pp < kNP;
          const bool keep = live && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;  // conv2 zero-pads t1, not x
          const uint32_t row = base + kT1 + pp * 128;
          const uint32_t sw = pp & 7;
#pragma unroll
          for (int u = 0; u < U12; ++u) {
            const int c0 = csel * (64 / CS) + u * 16;
            float bb[16];
            load_bias16(c0, bb);
            bias_relu_pack16(blk ? rb[u] : ra[u], bb, keep, qv);
            if (live) {
              sts128(row + ((((c0 >> 3)) ^ sw) << 4), qv[0]);
              sts128(row + ((((c0 >> 3) + 1) ^ sw) << 4), qv[1]);
            }
          }
        }
      }
      hy::fence_async_smem();
      hy::tc_fence_before();
      hy::mbar_arrive(bar(T1_READY));
      if (threadIdx.x == 0) RU_STAMP(2);

      // ---- E3 of the previous tile, under G2(i) / G1(i+1) ----
      if (it > 0) epilogue3(it - 1);

      // ---- E2: t2 (the t2 buffer was released by G3_DONE(i-1), observed in E3 above) ----
      hy::mbar_wait(bar(G2_DONE), ph);
      hy::tc_fence_after();
      if (threadIdx.x == 0) RU_STAMP(3);
      {
        // t2 goes to tensor memory (columns [192, 224): two channels per column) and G3 reads it from there as its A
        // operand: 16 KB of shared-memory writes and 16 KB of MMA operand reads per tile leave the shared-memory port
        uint32_t ra[U12][16];
#pragma unroll
        for (int u = 0; u < U12; ++u) hy::tmem_ld16(t_lane + kColG2 + csel * (64 / CS) + u * 16, ra[u]);
#pragma unroll
        for (int u = 0; u < U12; ++u) hy::tmem_ld_fence(ra[u]);
#pragma unroll
        for (int u = 0; u < U12; ++u) {
          const int c0 = csel * (64 / CS) + u * 16;
          float bb[16];
          load_bias16(64 + c0, bb);
          bias_relu_pack16(ra[u], bb, true, qv);
          const uint32_t w8[8] = {qv[0].x, qv[0].y, qv[0].z, qv[0].w, qv[1].x, qv[1].y, qv[1].z, qv[1].w};
          hy::tmem_st8(t_lane + kColT2 + (c0 >> 1), w8);
        }
      }
      hy::tmem_st_wait();
      hy::tc_fence_before();
      hy::mbar_arrive(bar(T2_READY));
      if (threadIdx.x == 0) RU_STAMP(4);
    }
    if (it > 0) epilogue3(it - 1);
  }

  hy::tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, kTmemCols);
  }
}

int encode_act4d(CUtensorMap* m, const void* ptr, int C, int ld, int B, int H, int W, int box_w, int box_h) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return hy_fail(HYRES_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(ru act C=%d ld=%d B=%d H=%d W=%d) -> %d", C, ld, B, H, W, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

bool is_conv(const hyres_conv* c, int cin, int cout, int k) {
  return c && c->kind == HYRES_CONV && c->cin0 == cin && c->cin1 == 0 && c->cout == cout && c->R == k && c->S == k &&
         c->stride == 1 && c->dil == 1 && c->pad == k / 2 && c->tap_mask.empty() && c->nsplit == 1;
}

}  // namespace

extern "C" {

int hyres_ru_supported(const hyres_conv* c1, const hyres_conv* c2, const hyres_conv* c3) {
  return is_conv(c1, 128, 64, 1) && is_conv(c2, 64, 64, 3) && is_conv(c3, 64, 128, 1) ? 1 : 0;
}

int hyres_ru_run(const hyres_conv* c1, const hyres_conv* c2, const hyres_conv* c3, const hyres_ru_io* io,
                 void* stream_v) {
  if (!io || !io->x || !io->out) return hy_fail(HYRES_ERR_ARG, "ru_run: null argument");
  if (!hyres_ru_supported(c1, c2, c3))
    return hy_fail(HYRES_ERR_UNSUPPORTED, "ru_run: needs 1x1 128->64, 3x3 64->64 (stride 1, pad 1), 1x1 64->128");
  if (io->B <= 0 || io->H <= 0 || io->W <= 0) return hy_fail(HYRES_ERR_ARG, "ru_run: empty input");
  if (io->ld_x < 128 || io->ld_out < 128 || (io->ld_x % 8) || (io->ld_out % 8))
    return hy_fail(HYRES_ERR_ARG, "ru_run: channel strides must be >= 128 and multiples of 8");
  if (io->x == io->out) return hy_fail(HYRES_ERR_ARG, "ru_run: in-place operation is not supported (halo reads)");
  RuParams p;
  memset(&p, 0, sizeof p);
  int rc = encode_act4d(&p.mapX, io->x, 128, io->ld_x, io->B, io->H, io->W, kPW, kTH + 2);
  if (rc != HYRES_OK) return rc;
  rc = encode_act4d(&p.mapSkip, io->x, 128, io->ld_x, io->B, io->H, io->W, kTW, kTH);
  if (rc != HYRES_OK) return rc;
  rc = encode_act4d(&p.mapOut, io->out, 128, io->ld_out, io->B, io->H, io->W, kTW, kTH);
  if (rc != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapW1, c1->d_w, c1->ktot, c1->cout_pad, 64)) != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapW2, c2->d_w, c2->ktot, c2->cout_pad, 64)) != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapW3, c3->d_w, c3->ktot, c3->cout_pad, 128)) != HYRES_OK) return rc;
  if (c1->h_bias.size() < 64 || c2->h_bias.size() < 64 || c3->h_bias.size() < 128)
    return hy_fail(HYRES_ERR_STATE, "ru_run: layer without a host bias copy");
  std::copy(c1->h_bias.begin(), c1->h_bias.begin() + 64, p.bias);
  std::copy(c2->h_bias.begin(), c2->h_bias.begin() + 64, p.bias + 64);
  std::copy(c3->h_bias.begin(), c3->h_bias.begin() + 128, p.bias + 128);
  p.H = io->H; p.W = io->W;
  p.tiles_w = (io->W + kTW - 1) / kTW;
  p.tiles_per_img = p.tiles_w * ((io->H + kTH - 1) / kTH);
  p.ntiles = p.tiles_per_img * io->B;
  p.final_relu = io->final_relu ? 1 : 0;
  {
    // HYRES_RU_TRACE=<device pointer, hex>: 64 x 16 clock64 stamps of CTA 0 (tools/experiments/ru_trace.py)
    static const char* e = getenv("HYRES_RU_TRACE");
    p.trace = e ? reinterpret_cast<long long*>(strtoull(e, nullptr, 16)) : nullptr;
  }
  const int smem = kSmemUsed + 1024;
  static const int cs = [] { const char* e = getenv("HYRES_RU_CS"); const int v = e ? atoi(e) : 0; return v == 4 ? 4 : 2; }();
  static HyPerDevice attr;
  if (!attr.done()) {
    HY_CUDA(cudaFuncSetAttribute(ru_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    HY_CUDA(cudaFuncSetAttribute(ru_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.mark();
  }
  const int grid = std::min(p.ntiles, num_sms());
  hy_count_launch();
  if (cs == 2) HY_CUDA(hy_launch_pdl(ru_fused_kernel<2>, grid, 128 * 2 + 96, smem, static_cast<cudaStream_t>(stream_v), p));
  else HY_CUDA(hy_launch_pdl(ru_fused_kernel<4>, grid, 128 * 4 + 96, smem, static_cast<cudaStream_t>(stream_v), p));
  return HYRES_OK;
}

}  // extern "C"
