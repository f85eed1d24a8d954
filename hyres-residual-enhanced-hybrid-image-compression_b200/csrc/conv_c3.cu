// First-layer convolutions of the two 3-channel inputs, fused with the residual arithmetic around them:
//
//   g_a.0            conv 5x5 / stride 2, 3 -> 128   on  residual = x - jpeg      (models/hyres.py:48,96 +
//                                                                                   models/checkerboard.py:36)
//   refine.conv_in   conv 3x3 / stride 1, 3 -> 64 + PReLU  on  x0 = jpeg + r_hat  (models/hyres.py:62,127 +
//                                                                                   models/layers/enhancement.py:60,89)
//
// With three input channels the layers are bound by HBM, not by the tensor pipe: the earlier path wrote an
// im2col tensor (128 / 256 B per position) and read it back through a 1x1 GEMM.  Here the im2col row of every
// output position is built on chip: the 128 threads of a 4 x 32 position tile first fill a small bf16 patch of
// src = a +- b (tile + halo, zero outside the image; coalesced row runs; src is written out once, by the tile
// that owns the pixel), then every thread reads its 27 / 75 taps from that patch at compile-time offsets,
// converts to bf16 and stores its row into the SWIZZLE_128B K-major A tile; tcgen05.mma multiplies by the
// shared-memory-resident weights into TMEM (the bias rides in two spare K columns as a bf16 hi + lo pair
// against constant-one activations); the epilogue (PReLU) stages bf16 rows over the dead A tile and leaves by
// TMA store.
// Two 128-thread groups per CTA alternate tiles so that one group's gather overlaps the other's MMA / epilogue;
// several CTAs share an SM.  HBM traffic is the algorithmic minimum: read the two images once, write src and
// the activation once.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kTileH = 4, kTileW = 32;
constexpr int kThreads = 288;  // warps 0-3 / 4-7: worker groups 0 / 1; warp 8: weights TMA + MMA issue

struct alignas(64) C3Params {
  CUtensorMap mapW, mapOut;
  const float* a;
  const float* b;
  float* sum_out;
  const float* bias;
  float sign, slope;
  int32_t act;
  int32_t B, H, W, OH, OW;
  int32_t tiles_w, tiles_per_img, ntiles;
};

__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float act_fn(float v, int act, float slope) {
  if (act == HYRES_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HYRES_ACT_PRELU) return v >= 0.f ? v : v * slope;
  if (act == HYRES_ACT_CLAMP01) return fminf(fmaxf(v, 0.f), 1.f);
  return v;
}

// KS: kernel size (3: stride 1 pad 1; 5: stride 2 pad 2), NOUT: output channels (multiple of 64).
template <int KS, int STRIDE, int NOUT>
__global__ void __launch_bounds__(kThreads, KS == 3 ? 3 : 2) conv_c3_kernel(const __grid_constant__ C3Params p) {
  constexpr int PAD = KS / 2;
  constexpr int KLIVE = KS * KS * 3;           // 27 / 75
  constexpr int NK16 = (KLIVE + 15) / 16;      // k-steps of 16: 2 / 5
  constexpr int KCH = (NK16 * 16 + 63) / 64;   // 64-wide chunks of the A tile: 1 / 2
  constexpr int NCHUNK16B = NK16 * 2;          // 16-byte chunks of a row the MMAs read
  constexpr int OCH = NOUT / 64;               // 64-channel chunks of the output staging
  constexpr uint32_t kBufBytes = (KCH > OCH ? KCH : OCH) * 16384;  // A tile, later the output staging
  constexpr uint32_t kWBytes = KCH * NOUT * 128;
  constexpr int RH = (kTileH - 1) * STRIDE + KS, RW = (kTileW - 1) * STRIDE + KS;  // src patch: 6 x 34 / 11 x 67
  constexpr uint32_t kPatchBytes = (3 * RH * RW * 2 + 127) / 128 * 128;  // bf16
  static_assert(KLIVE + 2 <= NK16 * 16, "two spare K columns carry the bias");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;
  const uint32_t buf_base = base + kWBytes;                 // [2 groups][kBufBytes]
  const uint32_t patch_base = buf_base + 2 * kBufBytes;     // [2 groups][3][RH][RW] bf16
  const uint32_t bar_base = patch_base + 2 * kPatchBytes;
  const uint32_t W_FULL = bar_base, A_READY = bar_base + 8, ACC_FULL = bar_base + 24, tmem_slot = bar_base + 40;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    hy::mbar_init(W_FULL, 1);
    for (int g = 0; g < 2; ++g) {
      hy::mbar_init(A_READY + 8 * g, 128);
      hy::mbar_init(ACC_FULL + 8 * g, 1);
    }
    hy::mbar_fence_init();
  }
  if (warp == 8) {
    if (lane == 0) {
      hy::tma_prefetch_desc(&p.mapW);
      hy::tma_prefetch_desc(&p.mapOut);
    }
    hy::tmem_alloc(tmem_slot, 2 * NOUT);
    hy::tmem_relinquish();
  }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  // Dependents (the next kernel of this stream) may be scheduled from here on -- only AFTER this CTA owns its
  // tensor memory: a dependent that lands on the same SM allocates TMEM in its prologue and then waits for this
  // grid to finish, so it must never be able to take the columns this CTA still has to allocate.
  hy::pdl_launch_dependents();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  hy::pdl_wait();  // (the weight / bias loads below are cheap; everything dependent comes after this)

  // tile k of this CTA is global tile blockIdx.x + k * gridDim.x; group (k & 1) owns it
  if (warp == 8) {
    // ===================== weights + MMA issue =====================
    if (lane == 0) {
      hy::mbar_arrive_expect_tx(W_FULL, kWBytes);
      for (int c = 0; c < KCH; ++c) hy::tma_load_2d(w_base + c * NOUT * 128, &p.mapW, W_FULL, c * 64, 0);
    }
    hy::mbar_wait(W_FULL, 0);
    // bias[n] -> weight columns KLIVE, KLIVE + 1 of row n as bf16 hi + lo (the A tile holds 1.0 there)
    for (int n = lane; n < NOUT; n += 32) {
      const float bv = __ldg(p.bias + n);
      const __nv_bfloat16 hi_b = __float2bfloat16_rn(bv);
      const __nv_bfloat16 lo_b = __float2bfloat16_rn(bv - __bfloat162float(hi_b));
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        constexpr int kk0 = KLIVE;
        const int kcol = kk0 + e;
        const uint32_t addr = w_base + (kcol >> 6) * NOUT * 128 + n * 128 + ((((kcol & 63) >> 3) ^ (n & 7)) << 4) + (kcol & 7) * 2;
        const unsigned short bits = __bfloat16_as_ushort(e ? lo_b : hi_b);
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(bits) : "memory");
      }
    }
    hy::fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      const uint32_t idesc = hy::umma_idesc_bf16(128, NOUT);
      constexpr uint32_t hi = hy::desc_hi_sw128();
      int k = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++k) {
        const int g = k & 1;
        const uint32_t a_lo = hy::desc_lo(buf_base + g * kBufBytes), w_lo = hy::desc_lo(w_base);
        hy::mbar_wait(A_READY + 8 * g, (k >> 1) & 1);
        hy::tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < NK16; ++kk)
          hy::umma_bf16(tmem_base + g * NOUT, hy::desc_pack(a_lo + (((kk >> 2) * 16384 + (kk & 3) * 32) >> 4), hi),
                        hy::desc_pack(w_lo + (((kk >> 2) * NOUT * 128 + (kk & 3) * 32) >> 4), hi), idesc, kk ? 1u : 0u);
        hy::umma_commit(ACC_FULL + 8 * g);
      }
    }
  } else {
    // ===================== worker groups: gather -> A tile, then epilogue =====================
    const int g = warp >> 2;
    const int row = threadIdx.x & 127;          // tile position == A row == TMEM lane
    const int tr = row >> 5, tc = row & 31;     // (warp & 3, lane)
    const uint32_t buf = buf_base + g * kBufBytes;
    const uint32_t sw = row & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * NOUT;
    const uint32_t patch = patch_base + g * kPatchBytes;
    const long long plane = static_cast<long long>(p.H) * p.W;
    const int epi_mode = p.act == HYRES_ACT_NONE ? 0 : (p.act == HYRES_ACT_PRELU && p.slope >= 0.f && p.slope <= 1.f ? 1 : 2);
    int k = g, it = 0;
    for (int t = blockIdx.x + g * gridDim.x; t < p.ntiles; t += 2 * gridDim.x, k += 2, ++it) {
      const int b_img = t / p.tiles_per_img;
      const int rem = t - b_img * p.tiles_per_img;
      const int th = rem / p.tiles_w;
      const int h0 = th * kTileH, w0 = (rem - th * p.tiles_w) * kTileW;
      // ---- fill the fp32 patch of src = a +- b (tile + halo, zero outside the image) ----
      const float* pa = p.a + static_cast<long long>(b_img) * 3 * plane;
      const bool has_b = p.b != nullptr;
      const float* pb = has_b ? p.b + static_cast<long long>(b_img) * 3 * plane : pa;
      float* so = p.sum_out ? p.sum_out + static_cast<long long>(b_img) * 3 * plane : nullptr;
      const int ihb = h0 * STRIDE - PAD, iwb = w0 * STRIDE - PAD;
      const int plane_i = static_cast<int>(plane);
      // warp w fills patch rows w, w + 4, ... (a row = one (channel, image row) run of RW values): the row
      // terms are warp-uniform, the column terms are computed once per tile.  Every load of the tile is
      // issued before the first use (read-only path, so they may pass the src stores).
      constexpr int CSTEPS = (RW + 31) / 32, QROWS = (3 * RH + 3) / 4;
      int col_off[CSTEPS];       // iw, or -1 outside the image / the patch
#pragma unroll
      for (int j = 0; j < CSTEPS; ++j) {
        const int cc = tc + 32 * j, iw = iwb + cc;
        col_off[j] = (cc < RW && static_cast<unsigned>(iw) < static_cast<unsigned>(p.W)) ? iw : -1;
      }
      float xv[QROWS][CSTEPS], xb[QROWS][CSTEPS];
#pragma unroll
      for (int q = 0; q < QROWS; ++q) {
        const int pr = tr + 4 * q;
        const int c = pr / RH, rr = pr - c * RH;
        const int ih = ihb + rr;
        const bool rok = pr < 3 * RH && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H);
        const int roff = c * plane_i + ih * p.W;
#pragma unroll
        for (int j = 0; j < CSTEPS; ++j) {
          const bool ok = rok && col_off[j] >= 0;
          xv[q][j] = ok ? __ldg(pa + roff + col_off[j]) : 0.f;
          xb[q][j] = (ok && has_b) ? __ldg(pb + roff + col_off[j]) : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < QROWS; ++q)
#pragma unroll
        for (int j = 0; j < CSTEPS; ++j) xv[q][j] = fmaf(p.sign, xb[q][j], xv[q][j]);
#pragma unroll
      for (int q = 0; q < QROWS; ++q) {
        const int pr = tr + 4 * q;
        const int c = pr / RH, rr = pr - c * RH;
        const int ih = ihb + rr;
        const bool rok = pr < 3 * RH && static_cast<unsigned>(ih) < static_cast<unsigned>(p.H);
        const bool rown = so != nullptr && rok && rr >= PAD && rr < PAD + kTileH * STRIDE;
        const int roff = c * plane_i + ih * p.W;
#pragma unroll
        for (int j = 0; j < CSTEPS; ++j) {
          const int cc = tc + 32 * j;
          // src leaves once, from the tile that owns the pixel
          if (rown && col_off[j] >= 0 && cc >= PAD && cc < PAD + kTileW * STRIDE) so[roff + col_off[j]] = xv[q][j];
          if (pr < 3 * RH && cc < RW) {
            const unsigned short hb = __bfloat16_as_ushort(__float2bfloat16_rn(xv[q][j]));
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(patch + (pr * RW + cc) * 2), "h"(hb) : "memory");
          }
        }
      }
      // the group's buffer still feeds the previous tile's TMA store
      if (it > 0 && row == 0) hy::tma_store_wait_read<0>();
      hy::named_bar_sync(1 + g, 128);
      // ---- this position's A row: k = (r*KS+s)*3+c <- patch[c][tr*S + r][tc*S + s]; bias columns = 1 ----
      {
        const uint32_t tbase = patch + ((tr * STRIDE) * RW + tc * STRIDE) * 2;
#pragma unroll
        for (int j = 0; j < NCHUNK16B; ++j) {
          uint32_t w[4];
#pragma unroll
          for (int e2 = 0; e2 < 4; ++e2) {
            uint32_t half[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int kk = 8 * j + 2 * e2 + h;
              if (kk < KLIVE) {
                const int c = kk % 3, rs = kk / 3, r = rs / KS, sx = rs % KS;
                unsigned short u;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(tbase + ((c * RH + r) * RW + sx) * 2));
                half[h] = u;
              } else {
                half[h] = kk < KLIVE + 2 ? 0x3f80u : 0u;  // bf16 1.0 in the two bias columns
              }
            }
            w[e2] = half[0] | (half[1] << 16);
          }
          sts128(buf + (j >> 3) * 16384 + row * 128 + (((j & 7) ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
        }
      }
      hy::fence_async_smem();
      hy::mbar_arrive(A_READY + 8 * g);

      // ---- epilogue: TMEM -> bias, activation -> bf16 rows staged over the A tile -> TMA store ----
      hy::mbar_wait(ACC_FULL + 8 * g, it & 1);
      hy::tc_fence_after();
      auto epilogue = [&](auto mode_tag) {
        constexpr int MODE = decltype(mode_tag)::value;
        uint32_t r0[32], r1[32];
        hy::tmem_ld32(t_lane, r0);
#pragma unroll
        for (int q = 0; q < NOUT / 32; ++q) {
          uint32_t(&cur)[32] = (q & 1) ? r1 : r0;
          uint32_t(&nxt)[32] = (q & 1) ? r0 : r1;
          hy::tmem_ld_fence32(cur);
          if (q + 1 < NOUT / 32) hy::tmem_ld32(t_lane + (q + 1) * 32, nxt);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f[i] = __uint_as_float(cur[j * 8 + i]);
              if (MODE == 1) f[i] = fmaxf(f[i], f[i] * p.slope);  // PReLU with 0 <= slope <= 1
              if (MODE == 2) f[i] = act_fn(f[i], p.act, p.slope);
            }
            uint4 o;
            o.x = hy::pack_bf16(f[0], f[1]); o.y = hy::pack_bf16(f[2], f[3]);
            o.z = hy::pack_bf16(f[4], f[5]); o.w = hy::pack_bf16(f[6], f[7]);
            const int n = q * 32 + j * 8;
            sts128(buf + (n >> 6) * 16384 + row * 128 + ((((n & 63) >> 3) ^ sw) << 4), o);
          }
        }
      };
      if (epi_mode == 0) epilogue(std::integral_constant<int, 0>{});
      else if (epi_mode == 1) epilogue(std::integral_constant<int, 1>{});
      else epilogue(std::integral_constant<int, 2>{});
      hy::tc_fence_before();
      hy::fence_async_smem();
      hy::named_bar_sync(1 + g, 128);
      if (row == 0) {
#pragma unroll
        for (int c = 0; c < OCH; ++c) hy::tma_store_4d(&p.mapOut, buf + c * 16384, c * 64, w0, h0, b_img);
        hy::tma_store_commit();
      }
    }
    if (row == 0) hy::tma_store_wait_all<0>();
  }

  hy::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, 2 * NOUT);
  }
}

int encode_out_map(CUtensorMap* m, const void* ptr, int C, int ld, int B, int H, int W) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return hy_fail(HYRES_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kTileW, (cuuint32_t)kTileH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(conv3ch out C=%d ld=%d B=%d H=%d W=%d) -> %d", C, ld, B, H, W, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

template <int KS, int STRIDE, int NOUT>
int launch(const C3Params& p, cudaStream_t stream) {
  constexpr int NK16 = (KS * KS * 3 + 15) / 16;
  constexpr int KCH = (NK16 * 16 + 63) / 64, OCH = NOUT / 64;
  constexpr int buf = (KCH > OCH ? KCH : OCH) * 16384;
  constexpr int RH = (kTileH - 1) * STRIDE + KS, RW = (kTileW - 1) * STRIDE + KS;
  constexpr int patch = (3 * RH * RW * 2 + 127) / 128 * 128;
  constexpr int smem = KCH * NOUT * 128 + 2 * buf + 2 * patch + 64 + 1024;
  static int per_sm = 0;
  if (!per_sm) {
    auto kern = conv_c3_kernel<KS, STRIDE, NOUT>;
    HY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    HY_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // CTAs per SM: registers, shared memory and the 512 TMEM columns (2 * NOUT each)
    cudaFuncAttributes fa;
    HY_CUDA(cudaFuncGetAttributes(&fa, kern));
    const int by_regs = 65536 / (std::max(fa.numRegs, 1) * kThreads);
    per_sm = std::max(1, std::min({by_regs, (227 * 1024) / (smem + 1024), 512 / (2 * NOUT), 4}));
  }
  const int grid = std::max(1, std::min((p.ntiles + 1) / 2, num_sms() * per_sm));
  hy_count_launch();
  HY_CUDA(hy_launch_pdl(conv_c3_kernel<KS, STRIDE, NOUT>, grid, kThreads, smem, stream, p));
  return HYRES_OK;
}

}  // namespace

extern "C" int hyres_conv3ch_run(hyres_conv* c, int ksize, int stride, const float* a, const float* b, int sign,
                                 float* sum_out, void* out_bf16, int ld_out, int B, int H, int W, int act, float slope,
                                 void* stream_v) {
  if (!c || !a || !out_bf16) return hy_fail(HYRES_ERR_ARG, "conv3ch_run: null argument");
  if (B <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "conv3ch_run: empty input");
  const bool k3 = ksize == 3 && stride == 1, k5 = ksize == 5 && stride == 2;
  if (!k3 && !k5) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv3ch_run: only 3x3 / stride 1 and 5x5 / stride 2");
  // the layer is the 1x1 GEMM over the im2col ordering k = (r*K+s)*3 + c, K padded to 64 / 128
  if (c->kind != HYRES_CONV || c->R != 1 || c->S != 1 || c->cin1 != 0 || c->cin0 != (k3 ? 64 : 128))
    return hy_fail(HYRES_ERR_ARG, "conv3ch_run: layer must be the 1x1 GEMM over the im2col ordering (cin 64 / 128)");
  if (c->cout != (k3 ? 64 : 128) || c->cout_pad != c->cout)
    return hy_fail(HYRES_ERR_UNSUPPORTED, "conv3ch_run: 3x3 needs 64 output channels, 5x5 / stride 2 needs 128");
  if (k5 && ((H | W) & 1)) return hy_fail(HYRES_ERR_ARG, "conv3ch_run: stride-2 conv needs even H and W");
  if (ld_out < c->cout || (ld_out % 8)) return hy_fail(HYRES_ERR_ARG, "conv3ch_run: ld_out must be >= cout and a multiple of 8");
  if (sum_out && !b) return hy_fail(HYRES_ERR_ARG, "conv3ch_run: sum_out without a second operand");
  C3Params p;
  memset(&p, 0, sizeof p);
  p.a = a; p.b = b; p.sum_out = sum_out; p.bias = c->d_bias;
  p.sign = sign < 0 ? -1.f : 1.f; p.slope = slope; p.act = act;
  p.B = B; p.H = H; p.W = W;
  p.OH = k5 ? H / 2 : H; p.OW = k5 ? W / 2 : W;
  p.tiles_w = (p.OW + kTileW - 1) / kTileW;
  p.tiles_per_img = p.tiles_w * ((p.OH + kTileH - 1) / kTileH);
  const long long nt = static_cast<long long>(p.tiles_per_img) * B;
  if (nt > 0x7fffffffLL) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv3ch_run: too many tiles");
  p.ntiles = static_cast<int>(nt);
  int rc = encode_w_map(&p.mapW, c->d_w, c->ktot, c->cout_pad, c->cout);
  if (rc != HYRES_OK) return rc;
  if ((rc = encode_out_map(&p.mapOut, out_bf16, c->cout, ld_out, B, p.OH, p.OW)) != HYRES_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  return k3 ? launch<3, 1, 64>(p, stream) : launch<5, 2, 128>(p, stream);
}
