// Implicit-GEMM convolution for the HyRES codec on Blackwell tensor cores.
//
// Replaces every nn.Conv2d / nn.ConvTranspose2d of the reference hot path
// (models/checkerboard.py:35-88, models/layers/attention.py:16-47,
// models/layers/enhancement.py:60-85, compressai GDN / ResidualBottleneckBlock).
//
// Design (B200-first, not a cuDNN translation):
//  * activations NHWC bf16, weights packed K-major bf16, fp32 accumulation in TMEM;
//  * one CTA owns a (16*MT) x 8 patch of output positions of one image and BN output
//    channels: MT accumulators of 128 x BN fp32 in tensor memory;
//  * A operand: TMA loads a *halo patch* [(16*MT+halo) rows][8 cols][64 ch] once per
//    (kernel column, 64-channel chunk); every vertical tap of that column is the same
//    patch shifted by whole rows = 1 KB, so the UMMA descriptor just moves its start
//    address (stays 1024 B aligned -> canonical SWIZZLE_128B K-major layout).  Image
//    borders are the TMA out-of-bounds zero fill: no padding pass, no predication;
//  * stride-2 convs read plain NHWC through a 5-D view (2C, W/2, 2, H/2, B): the
//    parity of a tap selects the inner offset / parity coordinate, so they are stride-1
//    patch loads too.  Transposed convs are 4 sub-pixel phases (9/6/6/4 taps) writing
//    interleaved output positions (grid.z = phase);
//  * B operand: [BN][64] weight tiles streamed through their own mbarrier ring;
//  * warp roles: warps 0-3 epilogue (TMEM lane == output position), warp 4 TMA
//    producer, warp 5 TMEM allocator + single-thread tcgen05.mma issuer;
//  * fused epilogues: bias, ReLU/PReLU/clamp, residual add, attention gate,
//    GDN/IGDN normalisation, per-pixel scale; bf16 / fp32 / squared outputs.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kTileW = 8;     // output columns per tile
constexpr int kSubH = 16;     // output rows per 128-row accumulator
// Warp roles.  Plain layers: warps 0-3 / 4-7 = epilogue groups 0 / 1, warp 8 = TMA producer, warp 9 = MMA issuer.
// Split-precision layers (kSplit): their fp32 epilogue is the critical path and latency-bound (ncu: 2.5 warps per
// scheduler active, 0.4 eligible), so each group has EIGHT warps -- warps w and w + 4 of a group read the same TMEM
// lane quarter and take alternate 32-column chunks -- 16 epilogue warps, warp 16 = producer, warp 17 = MMA issuer.
template <bool kSplit> struct Roles {
  static constexpr int kEpiWarps = kSplit ? 16 : 8;
  static constexpr int kWarpProd = kEpiWarps, kWarpMma = kEpiWarps + 1;
  static constexpr int kThreads = (kEpiWarps + 2) * 32;
};
constexpr int kMaxNBlocks = 8;
constexpr int kSmemLimit = 227 * 1024;

struct alignas(64) ConvParams {
  CUtensorMap mapA0;
  CUtensorMap mapA1;
  CUtensorMap mapB;
  const TapGroup* groups;
  int32_t ph_begin[4];
  int32_t ph_count[4];
  int32_t nb_n0[kMaxNBlocks];  // first output channel of each N block
  int32_t nb_bn[kMaxNBlocks];  // its width (multiple of b_box_rows, <= 256)
  int32_t n_blocks, b_box_rows, bn_max;
  int32_t nphase, tiles_w, tiles_h, nitems;
  int32_t OHv, OWv;          // extent of the iterated output grid (per phase for deconv)
  int32_t OH, OW;            // true output extent
  int32_t out_mul;           // 1, or 2 for transposed conv
  int32_t MT, NA, NB;
  int32_t a_stage_bytes, b_stage_bytes, patch_rows;
  int32_t tmem_cols, acc_cols, nbuf, acc_empty_count, fast_epi;
  int32_t nacc, acc_stride;  // split-precision layers: accumulators per tile, TMEM columns between them
  int32_t acc_mask[4];       // per phase: which accumulators hold a sum
  int32_t split_epi, split_mode, out_nsplit, split_square;
  int32_t split_f16, out_f16;  // half-part format of the operands / of the parts written by the epilogue
  const float* aux0_f32;
  const float* aux1_f32;
  __nv_bfloat16* out_split;
  int32_t cout;              // live output channels
  int32_t epi, act;
  float slope;
  const float* bias;
  const __nv_bfloat16* aux0;
  const __nv_bfloat16* aux1;
  const float* pixscale;
  int32_t ld_aux0, ld_aux1;
  __nv_bfloat16* out_bf16;
  __nv_bfloat16* out_sq;
  int32_t ld_out, ld_sq;
  float* out_f32;
  long long f32_sb, f32_sh, f32_sw, f32_sc;
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == HYRES_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HYRES_ACT_PRELU) return v >= 0.f ? v : v * slope;
  if (act == HYRES_ACT_CLAMP01) return fminf(fmaxf(v, 0.f), 1.f);
  return v;
}

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* p, const float (&f)[16]) {
  uint4 a, b;
  a.x = hy::pack_bf16(f[0], f[1]);
  a.y = hy::pack_bf16(f[2], f[3]);
  a.z = hy::pack_bf16(f[4], f[5]);
  a.w = hy::pack_bf16(f[6], f[7]);
  b.x = hy::pack_bf16(f[8], f[9]);
  b.y = hy::pack_bf16(f[10], f[11]);
  b.z = hy::pack_bf16(f[12], f[13]);
  b.w = hy::pack_bf16(f[14], f[15]);
  reinterpret_cast<uint4*>(p)[0] = a;
  reinterpret_cast<uint4*>(p)[1] = b;
}

struct Item {
  int b_img, h0, w0, phase, n0, bn;
};

// Persistent streaming implicit GEMM: one CTA per SM walks work items
// (output tile of MT x 128 positions) x (sub-pixel phase) x (block of <= 256 output channels),
// N-block fastest so that the CTAs running side by side share the same activation patch in L2.
// The A patches and the weight tiles stream through two deep TMA rings that keep filling across
// item boundaries; the accumulators are double buffered in TMEM whenever 2 * MT * BN <= 512 columns,
// and two epilogue warp groups drain them while the next item's MMAs are issued.
template <bool kSplit>
__global__ void __launch_bounds__(Roles<kSplit>::kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kWarpProd = Roles<kSplit>::kWarpProd, kWarpMma = Roles<kSplit>::kWarpMma;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + p.NA * p.a_stage_bytes;
  const uint32_t bar_base = b_base + p.NB * p.b_stage_bytes;
  // barrier slots (8 B each): a_full[NA] a_empty[NA] b_full[NB] b_empty[NB] acc_full[2] acc_empty[2] ; tmem slot
  const uint32_t a_full = bar_base;
  const uint32_t a_empty = a_full + 8 * p.NA;
  const uint32_t b_full = a_empty + 8 * p.NA;
  const uint32_t b_empty = b_full + 8 * p.NB;
  const uint32_t acc_full = b_empty + 8 * p.NB;
  const uint32_t acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.NA; ++i) {
      hy::mbar_init(a_full + 8 * i, 1);
      hy::mbar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < p.NB; ++i) {
      hy::mbar_init(b_full + 8 * i, 1);
      hy::mbar_init(b_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      hy::mbar_init(acc_full + 8 * i, 1);
      hy::mbar_init(acc_empty + 8 * i, p.acc_empty_count);
    }
    hy::mbar_fence_init();
  }
  if (warp == kWarpProd && lane == 0) {
    hy::tma_prefetch_desc(&p.mapA0);
    hy::tma_prefetch_desc(&p.mapA1);
    hy::tma_prefetch_desc(&p.mapB);
  }
  if (warp == kWarpMma) {
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

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  auto decode = [&](int item) {
    Item it;
    const int nb = item % p.n_blocks;
    int rest = item / p.n_blocks;
    it.phase = rest % p.nphase;
    rest /= p.nphase;
    it.b_img = rest / tiles_per_img;
    const int t_in = rest - it.b_img * tiles_per_img;
    const int th = t_in / p.tiles_w;
    it.h0 = th * (kSubH * p.MT);
    it.w0 = (t_in - th * p.tiles_w) * kTileW;
    it.n0 = p.nb_n0[nb];
    it.bn = p.nb_bn[nb];
    return it;
  };

  if (warp == kWarpProd) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const uint32_t a_bytes = p.patch_rows * (kTileW * 128);
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
        const Item it = decode(item);
        const int g_begin = p.ph_begin[it.phase], g_count = p.ph_count[it.phase];
        const uint32_t b_bytes = it.bn * 128;
        for (int g = 0; g < g_count; ++g) {
          const TapGroup tg = p.groups[g_begin + g];
          hy::mbar_wait(a_empty + 8 * sa, pa ^ 1u);
          hy::mbar_arrive_expect_tx(a_full + 8 * sa, a_bytes);
          hy::tma_load_5d(a_base + sa * p.a_stage_bytes, tg.src ? &p.mapA1 : &p.mapA0,
                          a_full + 8 * sa, tg.c_off, it.w0 + tg.dw, tg.hpar, it.h0 + tg.dh, it.b_img);
          for (int t = 0; t < tg.ntaps; ++t) {
            hy::mbar_wait(b_empty + 8 * sb, pb ^ 1u);
            hy::mbar_arrive_expect_tx(b_full + 8 * sb, b_bytes);
            const uint32_t dst = b_base + sb * p.b_stage_bytes;
            for (int r = 0; r < it.bn; r += p.b_box_rows)
              hy::tma_load_2d(dst + r * 128, &p.mapB, b_full + 8 * sb, (tg.kslot0 + t) * 64, it.n0 + r);
            if (++sb == p.NB) { sb = 0; pb ^= 1u; }
          }
          if (++sa == p.NA) { sa = 0; pa ^= 1u; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop converged (descriptor arithmetic in uniform registers); one elected lane issues.
    {
      const uint32_t leader = hy::elect_leader();
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it_n = 0;
      for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it_n) {
        const Item it = decode(item);
        const int g_begin = p.ph_begin[it.phase], g_count = p.ph_count[it.phase];
        const uint32_t idesc = p.split_f16 ? hy::umma_idesc_f16(128, it.bn) : hy::umma_idesc_bf16(128, it.bn);
        const int buf = p.nbuf == 2 ? (it_n & 1) : 0;
        const uint32_t use = p.nbuf == 2 ? (it_n >> 1) : it_n;  // how often this buffer was used before
        hy::mbar_wait(acc_empty + 8 * buf, (use & 1u) ^ 1u);
        hy::tc_fence_after();
        const uint32_t d_item = tmem_base + buf * p.acc_cols;
        uint32_t used = 0;  // bit a: accumulator a already holds a partial sum (accumulate flag)
        for (int g = 0; g < g_count; ++g) {
          const TapGroup tg = p.groups[g_begin + g];
          hy::mbar_wait(a_full + 8 * sa, pa);
          hy::tc_fence_after();
          const uint32_t a_stage = a_base + sa * p.a_stage_bytes;
          for (int t = 0; t < tg.ntaps; ++t) {
            hy::mbar_wait(b_full + 8 * sb, pb);
            hy::tc_fence_after();
            const uint64_t b_d = hy::desc_u64(b_base + sb * p.b_stage_bytes);
            const uint64_t a_d = hy::desc_u64(a_stage + p.groups[g_begin + g].tap_row[t] * (kTileW * 128));
            const int acc = p.nacc > 1 ? p.groups[g_begin + g].tap_acc[t] : 0;
            const uint32_t first = (used >> acc) & 1u;
            for (int sub = 0; sub < p.MT; ++sub) {
              const uint64_t a_sub = a_d + ((sub * (kSubH * kTileW * 128)) >> 4);
              const uint32_t d_tmem = d_item + acc * p.acc_stride + sub * p.bn_max;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                hy::umma_issue<2>(d_tmem, a_sub + 2 * k, b_d + 2 * k, idesc, first | static_cast<uint32_t>(k), leader);
            }
            used |= 1u << acc;
            hy::umma_commit_mode<2>(b_empty + 8 * sb, leader);
            if (++sb == p.NB) { sb = 0; pb ^= 1u; }
          }
          hy::umma_commit_mode<2>(a_empty + 8 * sa, leader);
          if (++sa == p.NA) { sa = 0; pa ^= 1u; }
        }
        hy::umma_commit_mode<2>(acc_full + 8 * buf, leader);
      }
    }
  } else {
    // ===================== epilogue groups (warps 0-3, 4-7) =====================
    // With two accumulator buffers the groups take alternate items; with one (MT * BN > 256) both
    // groups drain the same item, sub-tile `sub` going to group (sub & 1).
    // Software-pipelined over 16-column chunks: the TMEM load and the global loads of the
    // epilogue operands (skip / gate / GDN inputs) of chunk c+1 are in flight while chunk c is
    // finished and stored, and the first chunk's operands are requested before the
    // accumulator is complete, so one global-load latency is exposed per item, not per chunk.
    const int grp = kSplit ? (warp >> 3) : (warp >> 2);
    const int row = (warp & 3) * 32 + lane;  // TMEM lane == tile row
    const int ti = row >> 3;
    const int tj = row & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const bool need0 = p.epi == HYRES_EPI_ADD || p.epi == HYRES_EPI_GATE || p.epi == HYRES_EPI_GDN ||
                       p.epi == HYRES_EPI_IGDN;
    const bool need1 = p.epi == HYRES_EPI_GATE;
    int it_n = 0;
    for (int item = blockIdx.x; item < p.nitems; item += gridDim.x, ++it_n) {
      int buf, sub0, sub_step;
      uint32_t use;
      if (p.nbuf == 2) {
        if ((it_n & 1) != grp) continue;
        buf = it_n & 1; use = it_n >> 1; sub0 = 0; sub_step = 1;
      } else {
        if (grp >= p.MT) continue;
        buf = 0; use = it_n; sub0 = grp; sub_step = 2;
      }
      const Item it = decode(item);
      const int ph_p = it.phase >> 1, ph_q = it.phase & 1;
      const int nchunk = it.bn >> 4;
      const int nsub = (p.MT - sub0 + sub_step - 1) / sub_step;
      const int total = nsub * nchunk;
      const uint32_t t_item = t_lane + buf * p.acc_cols;
      if constexpr (kSplit) {
        // split-precision layer: v = sum of the tile's accumulators (cross products first, fp32 round to nearest)
        // + bias, the fp32 element-wise stage (skip / gate / GDN), ReLU -> fp32 row and / or its 16-bit parts.
        // TMEM hands every thread one position (row) of 32 channels; global memory wants a warp to touch few
        // 128-byte lines per instruction.  Each epilogue warp therefore owns a 4.5 KB shared-memory tile through
        // which operands and results are transposed: global accesses use the "line layout" (instruction j, lane l ->
        // row 4j + l/8, 16-byte piece l%8: four rows of 128 B per instruction instead of 32 scattered pieces).
        // The layer's few MMAs leave this epilogue as the critical path (ncu: issue-bound on 8 warps), so every
        // address below is a per-sub-tile base plus compile-time multiples of two strides -- no per-piece 64-bit
        // index arithmetic.
        const float lo = p.act == HYRES_ACT_RELU ? 0.f : -3.402823466e38f;
        const int nch = it.bn >> 5;
        const int steps = nsub * nch;
        const uint32_t mask = static_cast<uint32_t>(p.acc_mask[it.phase]);
        float* stg = reinterpret_cast<float*>(smem_raw + (bar_base + 512 - hy::smem_u32(smem_raw))) + warp * (32 * 36);
        const int wq = warp & 3;
        const int qrow = lane >> 3, qpc = lane & 7;  // line layout
        const int prow = lane >> 2, ppc = lane & 3;  // parts layout: instruction j -> row 8j + l/4, 16-byte piece l%4
        const bool coalesced = p.f32_sc == 1 && (p.cout & 31) == 0;
        const bool has_aux = p.split_mode != HYRES_SPLIT_COPY;
        const bool aux_staged = has_aux && coalesced;
        const int om = p.out_mul;
        const long long row_px = static_cast<long long>(om) * p.OW;  // output pixels between two tile rows
        // tile row `tr` (0..15 within the sub-tile), tile column `tc` -> output pixel index
        auto pixel = [&](int sub, int tr, int tc) -> long long {
          const int hv = it.h0 + sub * kSubH + tr, wv = it.w0 + tc;
          return (static_cast<long long>(it.b_img) * p.OH + (hv * om + ph_p)) * p.OW + (wv * om + ph_q);
        };
        // line layout of sub-tile `sub`: piece j sits at pixel  pix0 + (j >> 1) * row_px + (j & 1) * 4 * om ;
        // bit j of the mask says whether that pixel exists
        auto line_base = [&](int sub, long long& pix0, uint32_t& ok) {
          const int tr0 = wq * 4;
          pix0 = pixel(sub, tr0, qrow);
          const int rows_ok = min(max(p.OHv - (it.h0 + sub * kSubH + tr0), 0), 4);
          const uint32_t cols = (it.w0 + qrow < p.OWv ? 0x55u : 0u) | (it.w0 + 4 + qrow < p.OWv ? 0xaau : 0u);
          ok = ((1u << (2 * rows_ok)) - 1u) & cols;
        };
        auto chunk_of = [&](int q, int& sub, int& cb) {
          const int si = q / nch;
          sub = sub0 + si * sub_step;
          cb = (q - si * nch) << 5;
        };
        const int aux_row_e = static_cast<int>(row_px) * p.cout, aux_col_e = 4 * om * p.cout;  // element strides
        // operand of chunk q in the line layout; chunk q + 1 is requested before chunk q's accumulators are read
        auto fetch = [&](const float* src, int q, float4 (&dst)[8]) {
          if (q >= steps) return;
          int sub, cb;
          chunk_of(q, sub, cb);
          if (it.n0 + cb >= p.cout) return;
          long long pix0;
          uint32_t ok;
          line_base(sub, pix0, ok);
          const float* s0 = src + pix0 * p.cout + (it.n0 + cb + qpc * 4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((ok >> j) & 1u) dst[j] = __ldg(reinterpret_cast<const float4*>(s0 + (j >> 1) * aux_row_e + (j & 1) * aux_col_e));
          }
        };
        // line layout -> the warp's tile; every thread then reads its own row (8 float4 = 32 channels) from there
        auto stage = [&](const float4 (&src)[8]) {
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(stg + (4 * j + qrow) * 36 + qpc * 4) = src[j];
          __syncwarp();
        };
        float4 an[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) an[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // the two warps of a lane quarter take alternate chunks: this one q = half, half + 2, ...
        const int half = (warp >> 2) & 1;
        if (aux_staged) fetch(p.aux0_f32, half, an);
        hy::mbar_wait(acc_full + 8 * buf, use & 1u);
        hy::tc_fence_after();
        if (half >= steps) {  // a single-chunk item: nothing to read for this warp, release the accumulator at once
          hy::tc_fence_before();
          hy::mbar_arrive(acc_empty + 8 * buf);
        }
        const int f32_row_e = static_cast<int>(row_px * p.f32_sw), f32_col_e = static_cast<int>(4 * om * p.f32_sw);
        const int nparts = p.out_nsplit;
        const int sp_px = nparts * p.cout;  // elements per pixel of the parts tensor
        const int sp_row_e = static_cast<int>(row_px) * sp_px;
        int cur_sub = -1;
        long long pix_line = 0, pix_own = 0, pix_parts = 0;
        uint32_t ok_line = 0, ok_parts = 0;
        bool valid_row = false;
        for (int q = half; q < steps; q += 2) {
          int sub, cb;
          chunk_of(q, sub, cb);
          if (sub != cur_sub) {  // per sub-tile: the three pixel bases and their validity
            cur_sub = sub;
            line_base(sub, pix_line, ok_line);
            const int tr_own = wq * 4 + (lane >> 3), tc_own = lane & 7;
            pix_own = pixel(sub, tr_own, tc_own);
            valid_row = it.h0 + sub * kSubH + tr_own < p.OHv && it.w0 + tc_own < p.OWv;
            pix_parts = pixel(sub, wq * 4, prow);
            const int rows_ok = min(max(p.OHv - (it.h0 + sub * kSubH + wq * 4), 0), 4);
            ok_parts = it.w0 + prow < p.OWv ? ((1u << rows_ok) - 1u) : 0u;
          }
          const int n = it.n0 + cb;
          if (aux_staged) {
            stage(an);
            fetch(p.aux0_f32, q + 2, an);
          }
          const bool row_ok = valid_row && n < p.cout;
          const float4* g0 = reinterpret_cast<const float4*>(p.aux0_f32 + pix_own * p.cout + n);  // used when !aux_staged
          const float4* g1 = reinterpret_cast<const float4*>(p.aux1_f32 + pix_own * p.cout + n);  // gate only
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
          for (int a = p.nacc - 1; a >= 0; --a) {
            if (!((mask >> a) & 1u)) continue;
            uint32_t r[32];
            hy::tmem_ld32(t_item + a * p.acc_stride + sub * p.bn_max + cb, r);
            hy::tmem_ld_fence32(r);
            // half parts: the last accumulator holds the cross products, scaled by 2^11 (exact power of two)
            const float sc = (p.split_f16 && a == p.nacc - 1) ? hy::kF16LoInv : 1.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(r[i]), sc, v[i]);
          }
          if (q + 2 >= steps) {  // this warp's last chunk
            hy::tc_fence_before();
            hy::mbar_arrive(acc_empty + 8 * buf);
          }
          if (n >= p.cout) continue;  // warp-uniform
          const float4* bq = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bq + i);
            float4 t = make_float4(v[4 * i] + b4.x, v[4 * i + 1] + b4.y, v[4 * i + 2] + b4.z, v[4 * i + 3] + b4.w);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (aux_staged) a = *reinterpret_cast<const float4*>(stg + lane * 36 + i * 4);
            else if (has_aux && row_ok) a = __ldg(g0 + i);
            if (p.split_mode == HYRES_SPLIT_ADD) {
              t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
            } else if (p.split_mode == HYRES_SPLIT_GATE) {
              const float4 g = row_ok ? __ldg(g1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
              t.x = g.x * (1.f / (1.f + expf(-t.x))) + a.x;
              t.y = g.y * (1.f / (1.f + expf(-t.y))) + a.y;
              t.z = g.z * (1.f / (1.f + expf(-t.z))) + a.z;
              t.w = g.w * (1.f / (1.f + expf(-t.w))) + a.w;
            } else if (p.split_mode == HYRES_SPLIT_GDN) {
              t.x = a.x * (1.f / sqrtf(t.x)); t.y = a.y * (1.f / sqrtf(t.y));
              t.z = a.z * (1.f / sqrtf(t.z)); t.w = a.w * (1.f / sqrtf(t.w));
            } else if (p.split_mode == HYRES_SPLIT_IGDN) {
              t.x = a.x * sqrtf(t.x); t.y = a.y * sqrtf(t.y); t.z = a.z * sqrtf(t.z); t.w = a.w * sqrtf(t.w);
            }
            v[4 * i] = fmaxf(t.x, lo); v[4 * i + 1] = fmaxf(t.y, lo);
            v[4 * i + 2] = fmaxf(t.z, lo); v[4 * i + 3] = fmaxf(t.w, lo);
          }
          if (p.out_f32) {
            if (coalesced) {
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(stg + lane * 36 + i * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              __syncwarp();
              float* o0 = p.out_f32 + pix_line * p.f32_sw + (n + qpc * 4);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if ((ok_line >> j) & 1u)
                  *reinterpret_cast<float4*>(o0 + (j >> 1) * f32_row_e + (j & 1) * f32_col_e) =
                      *reinterpret_cast<const float4*>(stg + (4 * j + qrow) * 36 + qpc * 4);
            } else if (valid_row) {
              float* o = p.out_f32 + it.b_img * p.f32_sb + (pix_own / p.OW % p.OH) * p.f32_sh + (pix_own % p.OW) * p.f32_sw + n * p.f32_sc;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.cout) o[i * p.f32_sc] = v[i];
            }
          }
          if (p.out_split) {  // cout is a multiple of 32 here (checked by the entry point)
            if (p.split_square) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= v[i];
            }
            uint4* stg4 = reinterpret_cast<uint4*>(stg);
            __nv_bfloat16* sp0 = p.out_split + pix_parts * sp_px + (n + ppc * 8);
            for (int part = 0; part < nparts; ++part) {
              uint32_t w[16];
              if (p.out_f16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
                  const float2 f = __half22float2(h);
                  v[2 * i] = (v[2 * i] - f.x) * hy::kF16LoScale;  // exact residual, back in p0's range
                  v[2 * i + 1] = (v[2 * i + 1] - f.y) * hy::kF16LoScale;
                  w[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                  w[i] = *reinterpret_cast<const uint32_t*>(&h);
                  v[2 * i] -= __uint_as_float(w[i] << 16);          // exact: the residual of a round-to-nearest bf16
                  v[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
                }
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) stg4[lane * 5 + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
              __syncwarp();
              __nv_bfloat16* sp = sp0 + part * p.cout;
#pragma unroll
              for (int j = 0; j < 4; ++j)  // instruction j, lane l -> row 8j + l/4 (tile row wq*4 + j), piece l%4: 8 rows of 64 B
                if ((ok_parts >> j) & 1u)
                  *reinterpret_cast<uint4*>(sp + j * sp_row_e) = stg4[(8 * j + prow) * 5 + ppc];
            }
          }
        }
        continue;
      }
      if constexpr (!kSplit) {
      if (p.fast_epi) {
        // bias (+ReLU) -> bf16 / fp32 rows: 32 accumulator columns per step, TMEM loads double buffered
        const float lo = p.act == HYRES_ACT_RELU ? 0.f : -3.402823466e38f;
        const int nch = it.bn >> 5;
        const int steps = nsub * nch;
        uint32_t r0[32], r1[32];
        hy::mbar_wait(acc_full + 8 * buf, use & 1u);
        hy::tc_fence_after();
        hy::tmem_ld32(t_item + sub0 * p.bn_max, r0);
        auto body = [&](int q, uint32_t (&r)[32]) {
          const int si = q / nch;
          const int sub = sub0 + si * sub_step;
          const int n = it.n0 + ((q - si * nch) << 5);
          const int hv = it.h0 + sub * kSubH + ti, wv = it.w0 + tj;
          if (hv >= p.OHv || wv >= p.OWv || n >= p.cout) return;
          const int oh = hv * p.out_mul + ph_p, ow = wv * p.out_mul + ph_q;
          const long long opix = (static_cast<long long>(it.b_img) * p.OH + oh) * p.OW + ow;
          float v[32];
          const float4* bq = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(bq + i);
            v[4 * i] = fmaxf(__uint_as_float(r[4 * i]) + b4.x, lo);
            v[4 * i + 1] = fmaxf(__uint_as_float(r[4 * i + 1]) + b4.y, lo);
            v[4 * i + 2] = fmaxf(__uint_as_float(r[4 * i + 2]) + b4.z, lo);
            v[4 * i + 3] = fmaxf(__uint_as_float(r[4 * i + 3]) + b4.w, lo);
          }
          const bool full = n + 32 <= p.cout;
          if (p.out_bf16) {
            __nv_bfloat16* o = p.out_bf16 + opix * p.ld_out + n;
            if (full) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<uint4*>(o)[i] = make_uint4(hy::pack_bf16(v[8 * i], v[8 * i + 1]), hy::pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                                            hy::pack_bf16(v[8 * i + 4], v[8 * i + 5]), hy::pack_bf16(v[8 * i + 6], v[8 * i + 7]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.cout) o[i] = __float2bfloat16_rn(v[i]);
            }
          }
          if (p.out_f32) {
            float* o = p.out_f32 + it.b_img * p.f32_sb + oh * p.f32_sh + ow * p.f32_sw + n * p.f32_sc;
            if (full && p.f32_sc == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.cout) o[i * p.f32_sc] = v[i];
            }
          }
        };
        auto next_col = [&](int q) {
          const int si = q / nch;
          return static_cast<uint32_t>((sub0 + si * sub_step) * p.bn_max + ((q - si * nch) << 5));
        };
        for (int q = 0; q < steps; q += 2) {
          hy::tmem_ld_fence32(r0);
          if (q + 1 < steps) {
            hy::tmem_ld32(t_item + next_col(q + 1), r1);
          } else {
            hy::tc_fence_before();
            hy::mbar_arrive(acc_empty + 8 * buf);
          }
          body(q, r0);
          if (q + 1 < steps) {
            hy::tmem_ld_fence32(r1);
            if (q + 2 < steps) {
              hy::tmem_ld32(t_item + next_col(q + 2), r0);
            } else {
              hy::tc_fence_before();
              hy::mbar_arrive(acc_empty + 8 * buf);
            }
            body(q + 1, r1);
          }
        }
        continue;
      }
      auto geom = [&](int q, long long& opix, int& oh, int& ow, bool& valid, int& n, uint32_t& tcol) {
        const int si = q / nchunk;
        const int sub = sub0 + si * sub_step;
        const int ch = q - si * nchunk;
        const int hv = it.h0 + sub * kSubH + ti;
        const int wv = it.w0 + tj;
        n = it.n0 + ch * 16;
        tcol = sub * p.bn_max + ch * 16;
        valid = (hv < p.OHv) && (wv < p.OWv) && (n < p.cout);
        oh = hv * p.out_mul + ph_p;
        ow = wv * p.out_mul + ph_q;
        opix = (static_cast<long long>(it.b_img) * p.OH + oh) * p.OW + ow;
      };
      auto fetch_aux = [&](int q, uint4 (&x0)[2], uint4 (&x1)[2]) {
        long long opix; int oh, ow, n; bool valid; uint32_t tcol;
        geom(q, opix, oh, ow, valid, n, tcol);
        if (!valid) return;
        if (need0) {
          const uint4* g = reinterpret_cast<const uint4*>(p.aux0 + opix * p.ld_aux0 + n);
          x0[0] = __ldg(g); x0[1] = __ldg(g + 1);
        }
        if (need1) {
          const uint4* g = reinterpret_cast<const uint4*>(p.aux1 + opix * p.ld_aux1 + n);
          x1[0] = __ldg(g); x1[1] = __ldg(g + 1);
        }
      };
      auto unpack = [](const uint4 (&x)[2], float (&f)[16]) {
        const uint32_t u[8] = {x[0].x, x[0].y, x[0].z, x[0].w, x[1].x, x[1].y, x[1].z, x[1].w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { f[2 * i] = hy::bf16_lo(u[i]); f[2 * i + 1] = hy::bf16_hi(u[i]); }
      };
      uint4 a0c[2] = {}, a1c[2] = {}, a0n[2] = {}, a1n[2] = {};
      uint32_t rc[16];
      fetch_aux(0, a0c, a1c);
      hy::mbar_wait(acc_full + 8 * buf, use & 1u);
      hy::tc_fence_after();
      {
        long long opix; int oh, ow, n; bool valid; uint32_t tcol;
        geom(0, opix, oh, ow, valid, n, tcol);
        hy::tmem_ld16(t_item + tcol, rc);
        hy::tmem_ld_fence(rc);
      }
      for (int q = 0; q < total; ++q) {
        if (q + 1 < total) fetch_aux(q + 1, a0n, a1n);
        long long opix; int oh, ow, n; bool valid; uint32_t tcol;
        geom(q, opix, oh, ow, valid, n, tcol);
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rc[i]);
        if (q + 1 < total) {
          long long opix1; int oh1, ow1, n1; bool valid1; uint32_t tcol1;
          geom(q + 1, opix1, oh1, ow1, valid1, n1, tcol1);
          hy::tmem_ld16(t_item + tcol1, rc);
        } else {
          // every accumulator column of this thread's share now sits in registers: release the buffer
          hy::tc_fence_before();
          hy::mbar_arrive(acc_empty + 8 * buf);
        }
        if (valid) {
          const bool full16 = (n + 16 <= p.cout);
          if (p.epi == HYRES_EPI_PIXSCALE) {
            const float ps = __ldg(p.pixscale + opix);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= ps;
          }
          {
            const float4* bq = reinterpret_cast<const float4*>(p.bias + n);  // bias padded to cout_pad
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(bq + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
          if (p.epi == HYRES_EPI_ADD) {
            float a[16];
            unpack(a0c, a);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += a[i];
          } else if (p.epi == HYRES_EPI_GATE) {
            float a[16], x[16];
            unpack(a1c, a);
            unpack(a0c, x);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(a[i], hy::fast_sigmoid(v[i]), x[i]);
          } else if (p.epi == HYRES_EPI_GDN) {
            float x[16];
            unpack(a0c, x);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = x[i] * hy::fast_rsqrt(v[i]);
          } else if (p.epi == HYRES_EPI_IGDN) {
            float x[16];
            unpack(a0c, x);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = x[i] * hy::fast_sqrt(v[i]);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = apply_act(v[i], p.act, p.slope);

          if (p.out_bf16) {
            __nv_bfloat16* o = p.out_bf16 + opix * p.ld_out + n;
            if (full16) {
              store16_bf16(o, v);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < p.cout) o[i] = __float2bfloat16_rn(v[i]);
            }
          }
          if (p.out_sq) {
            __nv_bfloat16* o = p.out_sq + opix * p.ld_sq + n;
            float sq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              // square of the *stored* (bf16-rounded) activation: the GDN operand is x^2 of
              // the tensor the next layer sees.
              const float xb = __bfloat162float(__float2bfloat16_rn(v[i]));
              sq[i] = xb * xb;
            }
            if (full16) {
              store16_bf16(o, sq);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < p.cout) o[i] = __float2bfloat16_rn(sq[i]);
            }
          }
          if (p.out_f32) {
            float* o = p.out_f32 + it.b_img * p.f32_sb + oh * p.f32_sh + ow * p.f32_sw + n * p.f32_sc;
            if (full16 && p.f32_sc == 1) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(o)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (n + i < p.cout) o[i * p.f32_sc] = v[i];
            }
          }
        }
        if (q + 1 < total) {
          hy::tmem_ld_fence(rc);
          a0c[0] = a0n[0]; a0c[1] = a0n[1]; a1c[0] = a1n[0]; a1c[1] = a1n[1];
        }
      }
      }  // !kSplit
    }
  }

  hy::tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------

int choose_bn(int cout) {
  if (cout <= 16) return 16;
  int pad = (cout + 15) / 16 * 16;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (pad % bn == 0) return bn;
  return 16;
}

}  // namespace

namespace {

void build_plan(hyres_conv* c) {
  c->groups.clear();
  c->slots.clear();
  auto live = [&](int r, int s) { return c->tap_mask.empty() || c->tap_mask[r * c->S + s] != 0; };
  const int nsrc = c->cin1 > 0 ? 2 : 1;
  const int P = c->nsplit;
  int extra = 0;
  int lead = 0, phase_mask = 0;
  // One group = one activation patch (input part `part`, 64-channel chunk, kernel column) and the B tiles that
  // meet it: every live kernel row, times the weight parts 0 .. P-1-part of a split-precision layer (the small
  // products are issued first).
  auto add_group = [&](int src, int c_off, int dw, int hpar, const std::vector<std::pair<int, int>>& rows,
                       int chunk, int s, int part) {
    // rows: (row offset in view rows, r)
    if (rows.empty()) return;
    int mn = rows[0].first, mx = rows[0].first;
    for (auto& t : rows) { mn = std::min(mn, t.first); mx = std::max(mx, t.first); }
    TapGroup g{};
    g.src = src; g.c_off = c_off; g.dw = dw; g.dh = mn; g.hpar = hpar;
    g.kslot0 = static_cast<int>(c->slots.size());
    int n = 0;
    for (int wpart = P - 1 - part; wpart >= 0; --wpart)
      for (size_t i = 0; i < rows.size(); ++i) {
        g.tap_row[n] = rows[i].first - mn;
        // leading products round robin over the first nacc - 1 accumulators, cross products into the last
        int acc = 0;
        if (c->nacc > 1) acc = (part == 0 && wpart == 0) ? (lead++ % (c->nacc - 1)) : c->nacc - 1;
        g.tap_acc[n++] = acc;
        phase_mask |= 1 << acc;
        c->slots.push_back({src, chunk, rows[i].second, s, wpart});
      }
    g.ntaps = n;
    extra = std::max(extra, mx - mn);
    c->groups.push_back(g);
  };
  if (c->kind == HYRES_CONV && c->stride == 1) {
    c->nphase = 1;
    c->ph_begin[0] = 0;
    for (int part = P - 1; part >= 0; --part)
      for (int src = 0; src < nsrc; ++src) {
        const int cin = src ? c->cin1 : c->cin0;
        for (int ch = 0; ch < (cin + 63) / 64; ++ch)
          for (int s = 0; s < c->S; ++s) {
            std::vector<std::pair<int, int>> rows;
            for (int r = 0; r < c->R; ++r)
              if (live(r, s)) rows.push_back({r * c->dil - c->pad, r});
            add_group(src, part * cin + ch * 64, s * c->dil - c->pad, 0, rows, ch, s, part);
          }
      }
    c->ph_count[0] = static_cast<int>(c->groups.size());
    c->acc_mask[0] = phase_mask;
  } else if (c->kind == HYRES_CONV && c->stride == 2) {
    c->nphase = 1;
    c->ph_begin[0] = 0;
    for (int part = P - 1; part >= 0; --part)
      for (int ch = 0; ch < (c->cin0 + 63) / 64; ++ch)
        for (int s = 0; s < c->S; ++s) {
          const int ds = s - c->pad;
          const int q = ds & 1;
          const int wq = (ds - q) / 2;
          for (int par = 0; par < 2; ++par) {
            std::vector<std::pair<int, int>> rows;
            for (int r = 0; r < c->R; ++r) {
              const int dr = r - c->pad;
              if ((dr & 1) != par || !live(r, s)) continue;
              rows.push_back({(dr - par) / 2, r});
            }
            add_group(0, q * P * c->cin0 + part * c->cin0 + ch * 64, wq, par, rows, ch, s, part);
          }
        }
    c->ph_count[0] = static_cast<int>(c->groups.size());
    c->acc_mask[0] = phase_mask;
  } else {  // HYRES_DECONV_K5S2
    c->nphase = 4;
    for (int ph = 0; ph < 4; ++ph) {
      const int pp = ph >> 1, qq = ph & 1;
      c->ph_begin[ph] = static_cast<int>(c->groups.size());
      lead = 0;
      phase_mask = 0;
      for (int part = P - 1; part >= 0; --part)
        for (int ch = 0; ch < (c->cin0 + 63) / 64; ++ch)
          for (int s = qq; s < 5; s += 2) {
            std::vector<std::pair<int, int>> rows;
            for (int r = pp; r < 5; r += 2) rows.push_back({(pp + 2 - r) / 2, r});
            add_group(0, part * c->cin0 + ch * 64, (qq + 2 - s) / 2, 0, rows, ch, s, part);
          }
      c->ph_count[ph] = static_cast<int>(c->groups.size()) - c->ph_begin[ph];
      c->acc_mask[ph] = phase_mask;
    }
  }
  c->extra_rows = extra;
  c->ktot = static_cast<int>(c->slots.size()) * 64;
  c->macs_per_pos = static_cast<int64_t>(c->ktot) * c->cout_pad / (c->nphase);
}

// part `part` of an fp32 weight.  bf16 parts: v = p0 + p1 + p2 (+ an error below 2^-24 |v|), each rounded to nearest;
// half parts: p0 = half(v), p1 = half((v - p0) * 2^11).  Returned as raw 16-bit patterns.
inline __nv_bfloat16 split_part(float v, int part, bool f16) {
  if (f16) {
    const __half h0 = __float2half_rn(v);
    const __half h = part == 0 ? h0 : __float2half_rn((v - __half2float(h0)) * 2048.f);
    return __ushort_as_bfloat16(__half_as_ushort(h));
  }
  __nv_bfloat16 b = __float2bfloat16(v);
  for (int i = 0; i < part; ++i) {
    v -= __bfloat162float(b);
    b = __float2bfloat16(v);
  }
  return b;
}

void pack_weights(const hyres_conv* c, const float* w, std::vector<__nv_bfloat16>& out) {
  out.assign(static_cast<size_t>(c->cout_pad) * c->ktot, __float2bfloat16(0.f));
  const int RS = c->R * c->S;
  for (size_t ks = 0; ks < c->slots.size(); ++ks) {
    const auto& sl = c->slots[ks];
    const int cin_src = sl.src ? c->cin1 : c->cin0;
    const int cbase = (sl.src ? c->cin0 : 0) + sl.chunk * 64;
    for (int cc = 0; cc < 64; ++cc) {
      if (sl.chunk * 64 + cc >= cin_src) break;
      const int ci = cbase + cc;
      for (int n = 0; n < c->cout; ++n) {
        float v;
        if (c->kind == HYRES_DECONV_K5S2)
          v = w[(static_cast<size_t>(ci) * c->cout + n) * RS + sl.r * c->S + sl.s];
        else
          v = w[(static_cast<size_t>(n) * c->w_cin_total + ci) * RS + sl.r * c->S + sl.s];
        out[static_cast<size_t>(n) * c->ktot + ks * 64 + cc] = split_part(v, sl.wpart, c->split_f16);
      }
    }
  }
}

// Device-side packing (training: the weights change every step and live on the GPU): one thread per packed element.
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                    const hyres_conv::Slot* __restrict__ slots, int cout, int cout_pad, int ktot,
                                    int cin0, int cin1, int w_cin_total, int RS, int S, int deconv, int f16) {
  const long long total = static_cast<long long>(cout_pad) * ktot;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / ktot);
    const int k = static_cast<int>(idx - static_cast<long long>(n) * ktot);
    const hyres_conv::Slot sl = slots[k >> 6];
    const int cc = k & 63;
    const int cin_src = sl.src ? cin1 : cin0;
    float v = 0.f;
    if (n < cout && sl.chunk * 64 + cc < cin_src) {
      const int ci = (sl.src ? cin0 : 0) + sl.chunk * 64 + cc;
      v = deconv ? w[(static_cast<size_t>(ci) * cout + n) * RS + sl.r * S + sl.s]
                 : w[(static_cast<size_t>(n) * w_cin_total + ci) * RS + sl.r * S + sl.s];
    }
    if (f16) {
      uint16_t p0, p1;
      hy::split_f16x2(v, p0, p1);
      out[idx] = __ushort_as_bfloat16(sl.wpart ? p1 : p0);
      continue;
    }
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    for (int i = 0; i < sl.wpart; ++i) {
      v -= __bfloat162float(b);
      b = __float2bfloat16_rn(v);
    }
    out[idx] = b;
  }
}

__global__ void pad_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int cout, int cout_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cout_pad) out[i] = (b && i < cout) ? b[i] : 0.f;
}

int upload_weights(hyres_conv* c, const float* weight, const float* bias) {
  std::vector<__nv_bfloat16> packed;
  pack_weights(c, weight, packed);
  HY_CUDA(cudaMemcpy(c->d_w, packed.data(), packed.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  std::vector<float> b(c->cout_pad, 0.f);
  if (bias) std::copy(bias, bias + c->cout, b.begin());
  HY_CUDA(cudaMemcpy(c->d_bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
  c->h_bias = b;
  if (c->nsplit == 1 && conv_sc_applicable(c)) {
    conv_sc_pack(c, weight, packed);
    if (!c->d_w_tap) {
      c->w_tap_elems = static_cast<int64_t>(packed.size());
      HY_CUDA(cudaMalloc(&c->d_w_tap, packed.size() * sizeof(__nv_bfloat16)));
    }
    HY_CUDA(cudaMemcpy(c->d_w_tap, packed.data(), packed.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  }
  return HYRES_OK;
}

int encode_act_map(CUtensorMap* m, const void* ptr, int C, int B, int H, int W, int stride2, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return hy_fail(HYRES_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[5], strides[4];
  const cuuint64_t es = 2;
  if (!stride2) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = B;
    strides[0] = C * es; strides[1] = (cuuint64_t)W * C * es; strides[2] = (cuuint64_t)W * C * es;
    strides[3] = (cuuint64_t)H * W * C * es;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
    strides[0] = 2 * C * es; strides[1] = (cuuint64_t)W * C * es; strides[2] = 2ull * W * C * es;
    strides[3] = (cuuint64_t)H * W * C * es;
  }
  cuuint32_t box[5] = {64, (cuuint32_t)kTileW, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(act C=%d B=%d H=%d W=%d s2=%d rows=%d) -> %d", C, B, H, W,
             stride2, box_rows, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

}  // namespace

int encode_w_map(CUtensorMap* m, const void* ptr, int ktot, int cout_pad, int bn) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return hy_fail(HYRES_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)cout_pad};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[128];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(weights ktot=%d cout=%d bn=%d) -> %d", ktot, cout_pad, bn, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

// SMs the persistent kernels leave alone (hyres_set_reserved_sms): grids are sized to the rest
static std::atomic<int> g_reserved_sms{0};

extern "C" int hyres_set_reserved_sms(int n) {
  if (n < 0) return hy_fail(HYRES_ERR_ARG, "set_reserved_sms: negative count");
  return g_reserved_sms.exchange(n);
}

int num_sms() {
  static std::atomic<int> cache[64];  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  std::atomic<int>& slot = cache[dev & 63];
  int n = slot.load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    slot.store(n, std::memory_order_relaxed);
  }
  return std::max(n / 2, n - g_reserved_sms.load(std::memory_order_relaxed));
}

extern "C" {

int hyres_conv_create(hyres_conv** out, int kind, int cin0, int cin1, int w_cin_total, int cout, int R, int S,
                      int stride, int pad, int dil, const float* weight, const float* bias,
                      const uint8_t* tap_mask) {
  return hyres_conv_create_split(out, kind, cin0, cin1, w_cin_total, cout, R, S, stride, pad, dil, weight, bias,
                                 tap_mask, 1);
}

int hyres_conv_create_split(hyres_conv** out, int kind, int cin0, int cin1, int w_cin_total, int cout, int R, int S,
                            int stride, int pad, int dil, const float* weight, const float* bias,
                            const uint8_t* tap_mask, int nsplit) {
  if (!out) return hy_fail(HYRES_ERR_ARG, "conv_create: null argument");
  if (!hy_split_code_ok(nsplit)) return hy_fail(HYRES_ERR_ARG, "conv_create: nsplit must be 1, 2, 3 or 2 | HYRES_SPLIT_F16");
  const bool split_f16 = (nsplit & HYRES_SPLIT_F16) != 0;
  nsplit &= 15;
  if (kind != HYRES_CONV && kind != HYRES_DECONV_K5S2) return hy_fail(HYRES_ERR_ARG, "conv_create: bad kind");
  if (cin0 <= 0 || cin1 < 0 || cout <= 0 || (cin0 % 8) || (cin1 % 8))
    return hy_fail(HYRES_ERR_ARG, "conv_create: channel counts must be positive multiples of 8");
  if (kind == HYRES_DECONV_K5S2) {
    if (R != 5 || S != 5 || cin1 != 0) return hy_fail(HYRES_ERR_UNSUPPORTED, "deconv: only k5 s2 p2 op1");
    stride = 2; pad = 2; dil = 1;
  } else {
    if (stride != 1 && stride != 2) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv: stride must be 1 or 2");
    if (stride == 2 && (dil != 1 || cin1 != 0 || pad != R / 2 || R != S))
      return hy_fail(HYRES_ERR_UNSUPPORTED, "conv stride 2: needs dil=1, pad=k/2, one input");
    if (R > kMaxRows || S > 7 || dil < 1) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv: kernel too large");
    if (stride == 1 && 2 * pad != dil * (R - 1))
      return hy_fail(HYRES_ERR_UNSUPPORTED, "conv stride 1: only 'same' padding");
  }
  if (cin1 > 0 && (cin0 % 64)) return hy_fail(HYRES_ERR_UNSUPPORTED, "two-input conv: cin0 must be a multiple of 64");
  hyres_conv* c = new hyres_conv();
  c->kind = kind; c->cin0 = cin0; c->cin1 = cin1;
  c->w_cin_total = w_cin_total > 0 ? w_cin_total : cin0 + cin1;
  c->cout = cout; c->R = R; c->S = S; c->stride = stride; c->pad = pad; c->dil = dil;
  c->nsplit = nsplit;
  c->split_f16 = split_f16;
  if (tap_mask) c->tap_mask.assign(tap_mask, tap_mask + R * S);
  c->BN = choose_bn(cout);
  c->cout_pad = (cout + c->BN - 1) / c->BN * c->BN;
  if (nsplit > 1) {
    // one accumulator for the cross products plus one per ~72 leading MMAs (K / 16) of the longest chain, at most 3
    static const int nacc_env = [] { const char* e = getenv("HYRES_SPLIT_NACC"); return e ? atoi(e) : 0; }();
    int live = 0;
    for (int i = 0; i < R * S; ++i) live += (!tap_mask || tap_mask[i]) ? 1 : 0;
    if (kind == HYRES_DECONV_K5S2) live = 9;  // the largest sub-pixel phase
    const int lead = live * ((cin0 + 63) / 64 + (cin1 + 63) / 64) * 4;
    c->lead_mmas = lead;
    // short chains (K <= 256: at most 96 MMAs with the cross products) stay below one ulp in a single accumulator
    int nacc = (lead <= 16 && !split_f16) ? 1 : 1 + std::min(3, (lead + 71) / 72);  // half parts: never shared
    if (nacc_env >= (split_f16 ? 2 : 1) && nacc_env <= 4) nacc = nacc_env;
    c->nacc = nacc;
    c->bn_cap = nacc <= 2 ? 256 : 128;
    if (c->cout_pad > c->bn_cap && (c->cout_pad % 64)) { c->nacc = 2; c->bn_cap = 256; }  // N blocks are multiples of 64
  }
  build_plan(c);
  if (c->groups.empty()) { delete c; return hy_fail(HYRES_ERR_ARG, "conv_create: no live taps"); }
  cudaError_t e;
  e = cudaMalloc(&c->d_groups, c->groups.size() * sizeof(TapGroup));
  if (e == cudaSuccess) e = cudaMemcpy(c->d_groups, c->groups.data(), c->groups.size() * sizeof(TapGroup), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&c->d_w, static_cast<size_t>(c->cout_pad) * c->ktot * sizeof(__nv_bfloat16));
  if (e == cudaSuccess) e = cudaMalloc(&c->d_bias, c->cout_pad * sizeof(float));
  if (e != cudaSuccess) { hyres_conv_destroy(c); return hy_fail(HYRES_ERR_CUDA, cudaGetErrorString(e)); }
  if (conv_sc_applicable(c) && c->nsplit == 1) {
    c->w_tap_elems = conv_sc_packed_elems(c);
    e = cudaMalloc(&c->d_w_tap, c->w_tap_elems * sizeof(__nv_bfloat16));
    if (e != cudaSuccess) { hyres_conv_destroy(c); return hy_fail(HYRES_ERR_CUDA, cudaGetErrorString(e)); }
  }
  // weight == NULL: an empty layer whose packed operands arrive through hyres_conv_import_packed
  if (weight) {
    int rc = upload_weights(c, weight, bias);
    if (rc != HYRES_OK) { hyres_conv_destroy(c); return rc; }
  } else {
    c->h_bias.assign(c->cout_pad, 0.f);
  }
  *out = c;
  return HYRES_OK;
}

int64_t hyres_conv_packed_elems(const hyres_conv* c, int which) {
  if (!c) return 0;
  if (which == 0) return static_cast<int64_t>(c->cout_pad) * c->ktot;
  if (which == 1) return c->d_w_tap ? static_cast<int64_t>(c->w_tap_elems) : 0;
  return which == 2 ? c->cout_pad : 0;
}

int hyres_conv_export_packed(const hyres_conv* c, void* w, void* w_tap, float* bias) {
  if (!c || !w || !bias || (c->d_w_tap && !w_tap)) return hy_fail(HYRES_ERR_ARG, "conv_export_packed: null argument");
  HY_CUDA(cudaMemcpy(w, c->d_w, static_cast<size_t>(c->cout_pad) * c->ktot * sizeof(__nv_bfloat16), cudaMemcpyDeviceToHost));
  if (c->d_w_tap) HY_CUDA(cudaMemcpy(w_tap, c->d_w_tap, c->w_tap_elems * sizeof(__nv_bfloat16), cudaMemcpyDeviceToHost));
  HY_CUDA(cudaMemcpy(bias, c->d_bias, c->cout_pad * sizeof(float), cudaMemcpyDeviceToHost));
  return HYRES_OK;
}

int hyres_conv_import_packed(hyres_conv* c, const void* w, const void* w_tap, const float* bias) {
  if (!c || !w || !bias || (c->d_w_tap && !w_tap)) return hy_fail(HYRES_ERR_ARG, "conv_import_packed: null argument");
  HY_CUDA(cudaMemcpy(c->d_w, w, static_cast<size_t>(c->cout_pad) * c->ktot * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  if (c->d_w_tap) HY_CUDA(cudaMemcpy(c->d_w_tap, w_tap, c->w_tap_elems * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice));
  HY_CUDA(cudaMemcpy(c->d_bias, bias, c->cout_pad * sizeof(float), cudaMemcpyHostToDevice));
  c->h_bias.assign(bias, bias + c->cout_pad);
  return HYRES_OK;
}

int hyres_conv_update(hyres_conv* c, const float* weight, const float* bias) {
  if (!c || !weight) return hy_fail(HYRES_ERR_ARG, "conv_update: null argument");
  c->w_tap_valid = true;
  return upload_weights(c, weight, bias);
}

int hyres_conv_update_device(hyres_conv* c, const float* weight_dev, const float* bias_dev, void* stream_v) {
  if (!c || !weight_dev) return hy_fail(HYRES_ERR_ARG, "conv_update_device: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  if (!c->d_slots) {
    HY_CUDA(cudaMalloc(&c->d_slots, c->slots.size() * sizeof(hyres_conv::Slot)));
    HY_CUDA(cudaMemcpy(c->d_slots, c->slots.data(), c->slots.size() * sizeof(hyres_conv::Slot), cudaMemcpyHostToDevice));
  }
  const long long total = static_cast<long long>(c->cout_pad) * c->ktot;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8));
  hy_count_launch();
  pack_weights_kernel<<<grid, 256, 0, st>>>(weight_dev, c->d_w, c->d_slots, c->cout, c->cout_pad, c->ktot, c->cin0,
                                           c->cin1, c->w_cin_total, c->R * c->S, c->S,
                                           c->kind == HYRES_DECONV_K5S2 ? 1 : 0, c->split_f16 ? 1 : 0);
  HY_CUDA(cudaGetLastError());
  hy_count_launch();
  pad_bias_kernel<<<(c->cout_pad + 255) / 256, 256, 0, st>>>(bias_dev, c->d_bias, c->cout, c->cout_pad);
  HY_CUDA(cudaGetLastError());
  c->w_tap_valid = false;  // the tap-major copy of a three-output-channel layer is not refreshed: use the generic kernel
  return HYRES_OK;
}

void hyres_conv_destroy(hyres_conv* c) {
  if (!c) return;
  if (c->d_groups) cudaFree(c->d_groups);
  if (c->d_slots) cudaFree(c->d_slots);
  if (c->d_wg_items) cudaFree(c->d_wg_items);
  if (c->d_w) cudaFree(c->d_w);
  if (c->d_w_tap) cudaFree(c->d_w_tap);
  if (c->d_bias) cudaFree(c->d_bias);
  delete c;
}

int64_t hyres_conv_macs_per_pos(const hyres_conv* c) { return c ? c->macs_per_pos : 0; }

int hyres_conv_out_size(const hyres_conv* c, int H, int W, int* OH, int* OW) {
  if (!c || !OH || !OW) return hy_fail(HYRES_ERR_ARG, "conv_out_size: null argument");
  if (c->kind == HYRES_DECONV_K5S2) { *OH = 2 * H; *OW = 2 * W; }
  else if (c->stride == 2) { *OH = H / 2; *OW = W / 2; }
  else { *OH = H; *OW = W; }
  return HYRES_OK;
}

int hyres_conv_run(hyres_conv* c, const hyres_conv_io* io, void* stream_v) {
  if (!c || !io || !io->x0) return hy_fail(HYRES_ERR_ARG, "conv_run: null argument");
  if (c->cin1 > 0 && !io->x1) return hy_fail(HYRES_ERR_ARG, "conv_run: second input missing");
  if (io->B <= 0 || io->H <= 0 || io->W <= 0) return hy_fail(HYRES_ERR_ARG, "conv_run: empty input");
  if (c->kind == HYRES_CONV && c->stride == 2 && ((io->H | io->W) & 1))
    return hy_fail(HYRES_ERR_ARG, "conv_run: stride-2 conv needs even H and W");
  if ((io->epi == HYRES_EPI_ADD || io->epi == HYRES_EPI_GDN || io->epi == HYRES_EPI_IGDN || io->epi == HYRES_EPI_GATE) && !io->aux0)
    return hy_fail(HYRES_ERR_ARG, "conv_run: aux0 missing for epilogue");
  if (io->epi == HYRES_EPI_GATE && !io->aux1) return hy_fail(HYRES_ERR_ARG, "conv_run: aux1 missing for gate");
  if (io->epi == HYRES_EPI_PIXSCALE && !io->pixscale) return hy_fail(HYRES_ERR_ARG, "conv_run: pixscale missing");
  if (io->aux0 && (io->ld_aux0 % 8)) return hy_fail(HYRES_ERR_ARG, "conv_run: ld_aux0 must be a multiple of 8");
  if (io->aux1 && (io->ld_aux1 % 8)) return hy_fail(HYRES_ERR_ARG, "conv_run: ld_aux1 must be a multiple of 8");
  if (io->out_bf16 && (io->ld_out % 8)) return hy_fail(HYRES_ERR_ARG, "conv_run: ld_out must be a multiple of 8");
  if (io->out_sq && (io->ld_sq % 8)) return hy_fail(HYRES_ERR_ARG, "conv_run: ld_sq must be a multiple of 8");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);

  // layers whose weights fit in shared memory run on the persistent resident-weights kernel
  const bool res_only = io->out_pad != 0 || io->up_t2 != nullptr || io->up_t3 != nullptr;
  static const bool no_sc = getenv("HYRES_NO_SC") != nullptr;
  const bool split = c->nsplit > 1;  // split-precision layers run on the streaming kernel only
  if (split && (res_only || io->x0_square || io->epi != HYRES_EPI_LINEAR || io->out_bf16 || io->out_sq))
    return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: split-precision layers take the linear epilogue and write fp32 / parts");
  if (split) {
    if (io->act != HYRES_ACT_NONE && io->act != HYRES_ACT_RELU)
      return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: split-precision layers support ReLU only");
    if (!io->out_f32 && !io->out_split) return hy_fail(HYRES_ERR_ARG, "conv_run: split layer without an output");
    if (io->out_split && (!hy_split_code_ok(io->out_nsplit) || (c->cout % 32)))
      return hy_fail(HYRES_ERR_ARG, "conv_run: out_split needs a valid nsplit code and a multiple of 32 output channels");
    const int m = io->split_mode;
    if (m != HYRES_SPLIT_COPY && m != HYRES_SPLIT_ADD && m != HYRES_SPLIT_GATE && m != HYRES_SPLIT_GDN && m != HYRES_SPLIT_IGDN)
      return hy_fail(HYRES_ERR_ARG, "conv_run: bad split_mode");
    if (m != HYRES_SPLIT_COPY && !io->aux0_f32) return hy_fail(HYRES_ERR_ARG, "conv_run: aux0_f32 missing");
    if (m == HYRES_SPLIT_GATE && !io->aux1_f32) return hy_fail(HYRES_ERR_ARG, "conv_run: aux1_f32 missing");
    if (m != HYRES_SPLIT_COPY && (c->cout % 32)) return hy_fail(HYRES_ERR_ARG, "conv_run: fused split stage needs a multiple of 32 output channels");
  } else if (io->out_split || io->split_mode) {
    return hy_fail(HYRES_ERR_ARG, "conv_run: out_split / split_mode need a split-precision layer");
  }
  if (!no_sc && !res_only && !split) {
    int handled = 0;
    const int rc = conv_sc_try_run(c, io, stream, &handled);
    if (rc != HYRES_OK || handled) return rc;
  }
  static const bool no_res = getenv("HYRES_NO_RES") != nullptr;
  if (!no_res && !split) {
    int handled = 0;
    const int rc = conv_res_try_run(c, io, stream, &handled);
    if (rc != HYRES_OK || handled) return rc;
  }
  if (io->x0_square) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: x0_square is only available on resident-weight 1x1 layers");
  if (res_only) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: out_pad / up-add are only available on resident-weight layers");
  if (io->ld_x0 && io->ld_x0 != c->nsplit * c->cin0) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: strided x0 is only available on resident-weight layers");

  ConvParams p;
  memset(&p, 0, sizeof p);
  int OH, OW;
  hyres_conv_out_size(c, io->H, io->W, &OH, &OW);
  p.OH = OH; p.OW = OW;
  p.out_mul = c->kind == HYRES_DECONV_K5S2 ? 2 : 1;
  p.OHv = OH / p.out_mul; p.OWv = OW / p.out_mul;
  p.nphase = c->nphase;

  // N blocks: <= 256 output channels each, balanced, multiples of the weight TMA box
  static const int bn_cap_env = [] { const char* e = getenv("HYRES_TC_BNMAX"); const int v = e ? atoi(e) : 0; return (v >= 64 && v <= 256 && v % 64 == 0) ? v : 256; }();
  const int bn_cap = std::min(bn_cap_env, c->bn_cap);
  if (c->cout_pad <= bn_cap || (c->cout_pad <= 256 && c->cout_pad % 64)) {
    p.n_blocks = 1; p.nb_n0[0] = 0; p.nb_bn[0] = c->cout_pad;
  } else {
    if (c->cout_pad % 64) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: wide layers need a multiple of 64 output channels");
    const int nblk = (c->cout_pad + bn_cap - 1) / bn_cap;
    const int bn = ((c->cout_pad + nblk - 1) / nblk + 63) / 64 * 64;
    int n0 = 0, k = 0;
    while (n0 < c->cout_pad) {
      if (k >= kMaxNBlocks) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: too many output channels");
      p.nb_n0[k] = n0; p.nb_bn[k] = std::min(bn, c->cout_pad - n0);
      n0 += p.nb_bn[k]; ++k;
    }
    p.n_blocks = k;
  }
  p.bn_max = 0;
  for (int k = 0; k < p.n_blocks; ++k) p.bn_max = std::max(p.bn_max, p.nb_bn[k]);
  p.fast_epi = io->epi == HYRES_EPI_LINEAR && (io->act == HYRES_ACT_NONE || io->act == HYRES_ACT_RELU) && !io->out_sq;
  for (int k = 0; k < p.n_blocks; ++k)
    if (p.nb_bn[k] % 32) p.fast_epi = 0;
  p.b_box_rows = 64;
  for (int k = 0; k < p.n_blocks; ++k)
    while (p.nb_bn[k] % p.b_box_rows) p.b_box_rows >>= 1;   // cout_pad is a multiple of 16

  // Sub-tiles per item.  Taller items halve the weight traffic per MAC (the L2 -> SM path, about
  // 42 B/clk/SM, bounds a 128 x 256 item at under half of the tensor peak); they are used when the
  // layer still yields at least two waves of items.
  const int tiles_w = (p.OWv + kTileW - 1) / kTileW;
  auto items_for = [&](int mt) {
    return static_cast<long long>(io->B) * tiles_w * ((p.OHv + kSubH * mt - 1) / (kSubH * mt)) * c->nphase * p.n_blocks;
  };
  const int nacc = c->nacc;
  if (nacc * p.bn_max > 512) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: split layer too wide for its accumulators");
  p.nacc = nacc;
  int mt = io->mt_hint;
  if (mt != 1 && mt != 2 && mt != 4) mt = (p.bn_max * 2 <= 512 && items_for(2) >= 2LL * num_sms()) ? 2 : 1;
  while (mt > 1 && nacc * mt * p.bn_max > 512) mt >>= 1;
  // few MMAs per tile: the fused fp32 epilogue of a split layer dominates, so keep two accumulator buffers and let
  // the two epilogue groups drain one tile each under the next tile's MMAs
  if (split && c->lead_mmas <= 16)
    while (mt > 1 && 2 * nacc * mt * p.bn_max > 512) mt >>= 1;
  p.MT = mt;
  p.acc_stride = mt * p.bn_max;
  p.acc_cols = nacc * p.acc_stride;
  p.nbuf = 2 * p.acc_cols <= 512 ? 2 : 1;
  p.acc_empty_count = ((p.nbuf == 2 || mt == 1) ? 128 : 256) * (split ? 2 : 1);  // split layers: 8 warps per group
  for (int i = 0; i < 4; ++i) p.acc_mask[i] = c->acc_mask[i];
  int cols = 32;
  while (cols < p.nbuf * p.acc_cols) cols <<= 1;
  p.tmem_cols = cols;
  p.patch_rows = kSubH * mt + c->extra_rows;
  if (p.patch_rows > 256) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: patch too tall");
  p.a_stage_bytes = p.patch_rows * kTileW * 128;
  p.b_stage_bytes = p.bn_max * 128;
  // rings: fill shared memory; at least two patches and four weight tiles in flight
  int total_groups = 0, total_btiles = 0;
  for (int ph = 0; ph < c->nphase; ++ph) {
    total_groups += c->ph_count[ph];
    for (int g = 0; g < c->ph_count[ph]; ++g) total_btiles += c->groups[c->ph_begin[ph] + g].ntaps;
  }
  // split layers: eight 4.5 KB transposition tiles for the epilogue warps behind the barrier block
  const int fixed = 512 /*barriers*/ + 1024 /*alignment*/ + (split ? Roles<true>::kEpiWarps * 32 * 36 * 4 : 0);
  const double taps_per_group = static_cast<double>(total_btiles) / total_groups;
  p.NA = 2; p.NB = 2;
  auto smem_need = [&]() { return p.NA * p.a_stage_bytes + p.NB * p.b_stage_bytes + fixed; };
  if (smem_need() > kSmemLimit) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: tile does not fit shared memory");
  for (;;) {
    // grow the ring that is shallower in units of k-groups covered; fall back to the other one
    const bool prefer_a = p.NA * taps_per_group <= p.NB;
    bool grown = false;
    for (int attempt = 0; attempt < 2 && !grown; ++attempt) {
      const bool grow_a = (attempt == 0) == prefer_a;
      int& n = grow_a ? p.NA : p.NB;
      const int cap = grow_a ? 8 : 12;
      if (n >= cap) continue;
      ++n;
      if (smem_need() > kSmemLimit) --n; else grown = true;
    }
    if (!grown) break;
  }
  p.tiles_w = tiles_w;
  p.tiles_h = (p.OHv + kSubH * mt - 1) / (kSubH * mt);
  const long long nitems = items_for(mt);
  if (nitems > 0x7fffffffLL) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: too many work items");
  p.nitems = static_cast<int>(nitems);
  for (int i = 0; i < 4; ++i) { p.ph_begin[i] = c->ph_begin[i]; p.ph_count[i] = c->ph_count[i]; }
  p.groups = c->d_groups;
  p.cout = c->cout;
  p.epi = io->epi; p.act = io->act; p.slope = io->slope;
  p.bias = c->d_bias;
  p.aux0 = static_cast<const __nv_bfloat16*>(io->aux0); p.ld_aux0 = io->ld_aux0;
  p.aux1 = static_cast<const __nv_bfloat16*>(io->aux1); p.ld_aux1 = io->ld_aux1;
  p.pixscale = io->pixscale;
  p.out_bf16 = static_cast<__nv_bfloat16*>(io->out_bf16); p.ld_out = io->ld_out;
  p.out_sq = static_cast<__nv_bfloat16*>(io->out_sq); p.ld_sq = io->ld_sq;
  p.split_epi = split ? 1 : 0;
  p.split_mode = io->split_mode; p.out_nsplit = io->out_nsplit & 15; p.split_square = io->split_square;
  p.split_f16 = c->split_f16 ? 1 : 0;
  p.out_f16 = (io->out_nsplit & HYRES_SPLIT_F16) ? 1 : 0;
  p.aux0_f32 = io->aux0_f32; p.aux1_f32 = io->aux1_f32;
  p.out_split = static_cast<__nv_bfloat16*>(io->out_split);
  p.out_f32 = io->out_f32;
  p.f32_sb = io->f32_sb; p.f32_sh = io->f32_sh; p.f32_sw = io->f32_sw; p.f32_sc = io->f32_sc;

  const int s2 = (c->kind == HYRES_CONV && c->stride == 2) ? 1 : 0;
  int rc = encode_act_map(&p.mapA0, io->x0, c->nsplit * c->cin0, io->B, io->H, io->W, s2, p.patch_rows);
  if (rc != HYRES_OK) return rc;
  if (c->cin1 > 0) rc = encode_act_map(&p.mapA1, io->x1, c->nsplit * c->cin1, io->B, io->H, io->W, 0, p.patch_rows);
  else p.mapA1 = p.mapA0;
  if (rc != HYRES_OK) return rc;
  rc = encode_w_map(&p.mapB, c->d_w, c->ktot, c->cout_pad, p.b_box_rows);
  if (rc != HYRES_OK) return rc;

  const int smem = smem_need();
  static HyPerDevice attr;
  if (!attr.done()) {
    HY_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    HY_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr.mark();
  }
  const int grid = std::min(p.nitems, io->cta_limit > 0 ? std::min(io->cta_limit, num_sms()) : num_sms());
  hy_count_launch();
  if (split) HY_CUDA(hy_launch_pdl(conv_tc_kernel<true>, grid, Roles<true>::kThreads, smem, stream, p));
  else HY_CUDA(hy_launch_pdl(conv_tc_kernel<false>, grid, Roles<false>::kThreads, smem, stream, p));
  return HYRES_OK;
}

}  // extern "C"
