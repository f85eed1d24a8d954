// Convolutions with three output channels: g_s.8 (transposed 5x5 / stride 2, 128 -> 3,
// models/checkerboard.py:57) and refine.fusion.2 (3x3, 64 -> 3, models/layers/enhancement.py:85).
//
// As tap-by-tap implicit GEMMs these layers waste the tensor pipe: every tap is an MMA with N = 16 (3 live
// columns) whose cost is the shared-memory read of its 128 x 64 A tile, 36 / 100 of them per tile.  Here the
// taps move into the N dimension instead: ONE GEMM per tile multiplies the input patch (tile + halo, 180
// positions) by all taps' weights at once,
//
//     D[q][tap * 3 + c] = sum_k x[q][k] * w[tap][c][k]          N = 27 -> 32  /  75 -> 80,
//
// and the epilogue finishes the convolution as a gather: every output pixel sums the D entries of the
// (patch position, tap) pairs that land on it (9 for the 3x3; 9 / 6 / 6 / 4 for the four sub-pixel phases
// of the transposed conv), read from a small fp32 scratch in shared memory, adds the bias and writes fp32
// (NCHW or any strides) directly.  A tile costs 8 / 16 MMAs instead of 36 / 100.
//
// Persistent, one CTA per SM: TMA patch ring, weights resident, double-buffered TMEM accumulators, two
// epilogue groups on alternate tiles.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kTH = 16, kTW = 8, kPW = 10, kNP = 180;
constexpr int kThreads = 320;  // warps 0-3 / 4-7: epilogue groups; warp 8: TMA; warp 9: MMA
constexpr uint32_t kChunk = 23040;  // 180 rows x 128 B (the swizzle follows absolute address bits: no 1 KB padding needed)

struct alignas(64) ScParams {
  CUtensorMap mapA, mapW;
  const float* bias;
  float* out;
  long long sb, sh, sw, sc;  // output strides in elements
  int32_t H, W;              // input extent
  int32_t tiles_w, tiles_per_img, ntiles;
  int32_t act, vec2;
};

template <int MODE> struct Cfg;
template <> struct Cfg<0> {  // conv 3x3, stride 1, pad 1, Cin 64
  static constexpr int KCH = 1, NT = 9, N = 32, NA = 4;
};
template <> struct Cfg<1> {  // transposed conv 5x5, stride 2, pad 2, output_padding 1, Cin 128
  static constexpr int KCH = 2, NT = 25, N = 80, NA = 2;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_sc_kernel(const __grid_constant__ ScParams p) {
  using C = Cfg<MODE>;
  constexpr int NJ = C::NT * 3;                        // live accumulator columns
  constexpr uint32_t kWBytes = C::KCH * C::N * 128;
  constexpr uint32_t kStage = C::KCH * kChunk;
  constexpr uint32_t kPBytes = (kNP * NJ * 4 + 127) / 128 * 128;
  constexpr uint32_t kSlack = 32768 - kChunk;
  constexpr int ACCW = 2 * C::N;                       // two 128-row blocks cover the 180-position patch
  constexpr int TMEM_COLS = 2 * ACCW <= 128 ? 128 : (2 * ACCW <= 256 ? 256 : 512);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;
  const uint32_t st_base = base + ((kWBytes + 1023u) & ~1023u);
  const uint32_t p_base = st_base + C::NA * kStage + kSlack;  // slack: block 1 of the last chunk reads rows up to 255
  const uint32_t bar_base = p_base + 2 * kPBytes;
  const uint32_t W_FULL = bar_base, A_FULL = bar_base + 8, A_EMPTY = A_FULL + 8 * C::NA, ACC_FULL = A_EMPTY + 8 * C::NA,
                 ACC_EMPTY = ACC_FULL + 16, tmem_slot = ACC_EMPTY + 16;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    hy::mbar_init(W_FULL, 1);
    for (int i = 0; i < C::NA; ++i) {
      hy::mbar_init(A_FULL + 8 * i, 1);
      hy::mbar_init(A_EMPTY + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      hy::mbar_init(ACC_FULL + 8 * i, 1);
      hy::mbar_init(ACC_EMPTY + 8 * i, 128);
    }
    hy::mbar_fence_init();
    // the weights do not depend on the predecessor kernel: their load starts before the dependency wait
    hy::mbar_arrive_expect_tx(W_FULL, kWBytes);
    for (int c = 0; c < C::KCH; ++c) hy::tma_load_2d(w_base + c * C::N * 128, &p.mapW, W_FULL, c * 64, 0);
  }
  if (warp == 8 && lane == 0) hy::tma_prefetch_desc(&p.mapA);
  if (warp == 9) {
    hy::tmem_alloc(tmem_slot, TMEM_COLS);
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
  hy::pdl_wait();  // everything above is independent of the predecessor kernel

  auto tile_origin = [&](int t, int& b_img, int& h0, int& w0) {
    b_img = t / p.tiles_per_img;
    const int rem = t - b_img * p.tiles_per_img;
    const int th = rem / p.tiles_w;
    h0 = th * kTH;
    w0 = (rem - th * p.tiles_w) * kTW;
  };

  if (warp == 8) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
        const int stage = it % C::NA;
        const uint32_t par = (it / C::NA) & 1;
        int b_img, h0, w0;
        tile_origin(t, b_img, h0, w0);
        hy::mbar_wait(A_EMPTY + 8 * stage, par ^ 1u);
        hy::mbar_arrive_expect_tx(A_FULL + 8 * stage, C::KCH * kNP * 128);
        for (int c = 0; c < C::KCH; ++c)
          hy::tma_load_4d(st_base + stage * kStage + c * kChunk, &p.mapA, A_FULL + 8 * stage, c * 64, w0 - 1, h0 - 1, b_img);
      }
    }
  } else if (warp == 9) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = hy::umma_idesc_bf16(128, C::N);
      constexpr uint32_t hi = hy::desc_hi_sw128();
      const uint32_t w_lo = hy::desc_lo(w_base);
      hy::mbar_wait(W_FULL, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
        const int stage = it % C::NA, buf = it & 1;
        const uint32_t a_lo = hy::desc_lo(st_base + stage * kStage);
        hy::mbar_wait(A_FULL + 8 * stage, (it / C::NA) & 1);
        hy::mbar_wait(ACC_EMPTY + 8 * buf, ((it >> 1) & 1) ^ 1u);
        hy::tc_fence_after();
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
          for (int kc = 0; kc < C::KCH; ++kc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              hy::umma_bf16(tmem_base + buf * ACCW + blk * C::N,
                            hy::desc_pack(a_lo + ((kc * kChunk + blk * 16384 + k * 32) >> 4), hi),
                            hy::desc_pack(w_lo + ((kc * C::N * 128 + k * 32) >> 4), hi), idesc, (kc | k) ? 1u : 0u);
        hy::umma_commit(A_EMPTY + 8 * stage);
        hy::umma_commit(ACC_FULL + 8 * buf);
      }
    }
  } else {
    // ============================ epilogue groups ============================
    const int grp = warp >> 2;
    const int tid = threadIdx.x & 127;  // TMEM lane; also the output position (ti, tj) of the gather phase
    const int ti = tid >> 3, tj = tid & 7;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + grp * ACCW;
    const uint32_t P = p_base + grp * kPBytes;
    const float b0 = __ldg(p.bias), b1 = __ldg(p.bias + 1), b2 = __ldg(p.bias + 2);
    for (int it = grp;; it += 2) {
      const int t = blockIdx.x + it * static_cast<int>(gridDim.x);
      if (t >= p.ntiles) break;
      int b_img, h0, w0;
      tile_origin(t, b_img, h0, w0);
      hy::mbar_wait(ACC_FULL + 8 * grp, (it >> 1) & 1);
      hy::tc_fence_after();
      // ---- phase 1: D rows of the 180 patch positions -> fp32 scratch P[q][j] (row stride NJ words, odd) ----
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const int q = blk * 128 + tid;
        if (blk == 1 && (warp & 3) >= 2) break;  // rows 128..179 live in lanes 0..51 (warp-uniform)
#pragma unroll
        for (int c0 = 0; c0 < NJ; c0 += 16) {
          uint32_t r[16];
          hy::tmem_ld16(t_lane + blk * C::N + c0, r);
          hy::tmem_ld_fence(r);
          if (q < kNP) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < NJ) asm volatile("st.shared.b32 [%0], %1;" ::"r"(P + (q * NJ + c0 + i) * 4), "r"(r[i]) : "memory");
          }
        }
      }
      hy::tc_fence_before();
      hy::mbar_arrive(ACC_EMPTY + 8 * grp);
      hy::named_bar_sync(1 + grp, 128);
      // ---- phase 2: gather the (patch position, tap) pairs of every output pixel ----
      auto ldp = [&](int q, int j) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(P + (q * NJ + j) * 4));
        return v;
      };
      if (MODE == 0) {
        const int oh = h0 + ti, ow = w0 + tj;
        float a0 = b0, a1 = b1, a2 = b2;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int q = (ti + r) * kPW + tj + s, j = (r * 3 + s) * 3;
            a0 += ldp(q, j);
            a1 += ldp(q, j + 1);
            a2 += ldp(q, j + 2);
          }
        if (p.act == HYRES_ACT_CLAMP01) {
          a0 = fminf(fmaxf(a0, 0.f), 1.f); a1 = fminf(fmaxf(a1, 0.f), 1.f); a2 = fminf(fmaxf(a2, 0.f), 1.f);
        }
        if (oh < p.H && ow < p.W) {
          float* o = p.out + b_img * p.sb + oh * p.sh + ow * p.sw;
          o[0] = a0; o[p.sc] = a1; o[2 * p.sc] = a2;
        }
      } else {
        // output (2*ih + ph, 2*iw + pw) of input position (ih, iw) = (h0 + ti, w0 + tj):
        //   rows r = ph, ph + 2, (ph + 4) contribute from input row ih + (ph + 2 - r) / 2, i.e. patch row ti + 1 + d, d = 1, 0, -1
        const int ih = h0 + ti, iw = w0 + tj;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          float acc[2][3];
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            acc[pw][0] = b0; acc[pw][1] = b1; acc[pw][2] = b2;
#pragma unroll
            for (int r = ph; r < 5; r += 2)
#pragma unroll
              for (int s = pw; s < 5; s += 2) {
                const int q = (ti + 1 + (ph + 2 - r) / 2) * kPW + tj + 1 + (pw + 2 - s) / 2, j = (r * 5 + s) * 3;
                acc[pw][0] += ldp(q, j);
                acc[pw][1] += ldp(q, j + 1);
                acc[pw][2] += ldp(q, j + 2);
              }
          }
          if (p.act == HYRES_ACT_CLAMP01) {
#pragma unroll
            for (int pw = 0; pw < 2; ++pw)
#pragma unroll
              for (int c = 0; c < 3; ++c) acc[pw][c] = fminf(fmaxf(acc[pw][c], 0.f), 1.f);
          }
          if (ih < p.H && iw < p.W) {
            float* o = p.out + b_img * p.sb + (2 * ih + ph) * p.sh + (2 * iw) * p.sw;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              if (p.vec2) {
                *reinterpret_cast<float2*>(o + c * p.sc) = make_float2(acc[0][c], acc[1][c]);
              } else {
                o[c * p.sc] = acc[0][c];
                o[c * p.sc + p.sw] = acc[1][c];
              }
            }
          }
        }
      }
      hy::named_bar_sync(1 + grp, 128);  // the scratch may be overwritten by this group's next tile
    }
  }

  hy::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int encode_patch_map(CUtensorMap* m, const void* ptr, int C, int ld, int B, int H, int W) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return hy_fail(HYRES_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)kPW, (cuuint32_t)(kTH + 2), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(conv_sc act C=%d ld=%d B=%d H=%d W=%d) -> %d", C, ld, B, H, W, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

template <int MODE>
int launch(const ScParams& p, cudaStream_t stream) {
  using C = Cfg<MODE>;
  constexpr int NJ = C::NT * 3;
  constexpr int smem = ((C::KCH * C::N * 128 + 1023) & ~1023) + C::NA * C::KCH * kChunk + (32768 - kChunk) +
                       2 * ((kNP * NJ * 4 + 127) / 128 * 128) + 256 + 1024;
  static_assert(smem <= 227 * 1024, "conv_sc: shared memory budget");
  static HyPerDevice attr;
  if (!attr.done()) {
    HY_CUDA(cudaFuncSetAttribute(conv_sc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr.mark();
  }
  const int grid = std::min(p.ntiles, num_sms());
  hy_count_launch();
  HY_CUDA(hy_launch_pdl(conv_sc_kernel<MODE>, grid, kThreads, smem, stream, p));
  return HYRES_OK;
}

}  // namespace

// Tap-major packed weights [N][Cin] (row j = tap * 3 + c, K-major bf16) for the layers this kernel serves;
// returns false when the layer is not one of them.
bool conv_sc_applicable(const hyres_conv* c) {
  if (c->cout != 3 || c->cin1 != 0 || !c->tap_mask.empty()) return false;
  if (c->kind == HYRES_DECONV_K5S2) return c->cin0 == 128;
  return c->kind == HYRES_CONV && c->R == 3 && c->S == 3 && c->stride == 1 && c->dil == 1 && c->pad == 1 && c->cin0 == 64;
}

int64_t conv_sc_packed_elems(const hyres_conv* c) {
  return static_cast<int64_t>(c->kind == HYRES_DECONV_K5S2 ? 80 : 32) * c->cin0;
}

void conv_sc_pack(const hyres_conv* c, const float* w, std::vector<__nv_bfloat16>& out) {
  const bool deconv = c->kind == HYRES_DECONV_K5S2;
  const int NT = deconv ? 25 : 9, N = deconv ? 80 : 32, cin = c->cin0;
  out.assign(static_cast<size_t>(N) * cin, __float2bfloat16(0.f));
  for (int tap = 0; tap < NT; ++tap)
    for (int co = 0; co < 3; ++co)
      for (int k = 0; k < cin; ++k) {
        // nn.Conv2d weight [Cout][Cin][R][S]; nn.ConvTranspose2d weight [Cin][Cout][R][S]
        const float v = deconv ? w[(static_cast<size_t>(k) * 3 + co) * NT + tap] : w[(static_cast<size_t>(co) * cin + k) * NT + tap];
        out[static_cast<size_t>(tap * 3 + co) * cin + k] = __float2bfloat16(v);
      }
}

int conv_sc_try_run(hyres_conv* c, const hyres_conv_io* io, cudaStream_t stream, int* handled) {
  *handled = 0;
  if (!c->d_w_tap || !c->w_tap_valid || !conv_sc_applicable(c)) return HYRES_OK;
  if (io->out_bf16 || io->out_sq || !io->out_f32 || io->epi != HYRES_EPI_LINEAR) return HYRES_OK;
  if (io->act != HYRES_ACT_NONE && io->act != HYRES_ACT_CLAMP01) return HYRES_OK;
  if (io->x0_square || (io->ld_x0 && io->ld_x0 != c->cin0)) return HYRES_OK;
  const bool deconv = c->kind == HYRES_DECONV_K5S2;
  ScParams p;
  memset(&p, 0, sizeof p);
  p.bias = c->d_bias;
  p.out = io->out_f32;
  p.sb = io->f32_sb; p.sh = io->f32_sh; p.sw = io->f32_sw; p.sc = io->f32_sc;
  p.H = io->H; p.W = io->W;
  p.tiles_w = (io->W + kTW - 1) / kTW;
  p.tiles_per_img = p.tiles_w * ((io->H + kTH - 1) / kTH);
  const long long nt = static_cast<long long>(p.tiles_per_img) * io->B;
  if (nt > 0x7fffffffLL) return HYRES_OK;
  p.ntiles = static_cast<int>(nt);
  p.act = io->act;
  p.vec2 = (p.sw == 1 && !(reinterpret_cast<uintptr_t>(p.out) & 7) && !(p.sh & 1) && !(p.sb & 1) && !(p.sc & 1)) ? 1 : 0;
  int rc = encode_patch_map(&p.mapA, io->x0, c->cin0, c->cin0, io->B, io->H, io->W);
  if (rc != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapW, c->d_w_tap, c->cin0, deconv ? 80 : 32, deconv ? 80 : 32)) != HYRES_OK) return rc;
  rc = deconv ? launch<1>(p, stream) : launch<0>(p, stream);
  if (rc == HYRES_OK) *handled = 1;
  return rc;
}
