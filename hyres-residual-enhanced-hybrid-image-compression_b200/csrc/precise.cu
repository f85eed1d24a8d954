// Element-wise side of the split-precision (fp32-equivalent) trunk.
//
// The entropy-critical half of the codec -- g_a, h_a, h_s, context_prediction, param_aggregation
// (models/checkerboard.py:35-45,61-88) -- decides integer symbols and CDF indexes, which must equal the fp32
// reference's (north star: bit-exact symbols).  Those layers therefore keep fp32 activations in HBM and run
// on the bf16 tensor cores as split products: every fp32 value is carried as nsplit bf16 parts
// (v = p0 + p1 + p2, 8 mantissa bits each, round-to-nearest residuals) and the convolution kernel multiplies
// part i of the activations with parts 0 .. nsplit-1-i of the weights into one fp32 accumulator
// (conv_tc.cu: hyres_conv_create_split).  The kernels here produce the parts and do the non-linear work
// between two convolutions in plain fp32 with IEEE division / sqrt / expf (no fast-math intrinsics):
//   copy / ReLU, residual add, attention gate a * sigmoid(b) + x (models/layers/attention.py:44-47),
//   GDN x * rsqrt(gamma x^2 + beta) and its inverse, the x^2 operand of GDN, round-about-the-median of the
//   hyper-latent, and the im2col of the 3-channel first layer.
// HBM-bound: float4 in, 8-byte bf16x4 out per part, grid-stride.
#include <cstdint>

#include "common.cuh"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kBlock = 256;

inline int grid_for(int64_t n, int per_block, int cap = 148 * 16) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return static_cast<int>(g);
}

// v -> its parts in the format of an nsplit code: 1..3 bf16 parts, or two half parts (2 | HYRES_SPLIT_F16)
__device__ __forceinline__ void split_code(float v, bool f16, uint16_t& p0, uint16_t& p1, uint16_t& p2) {
  if (f16) {
    hy::split_f16x2(v, p0, p1);
    p2 = 0;
  } else {
    hy::split_bf16x3(v, p0, p1, p2);
  }
}

__device__ __forceinline__ uint32_t pack2(uint16_t a, uint16_t b) {
  return static_cast<uint32_t>(a) | (static_cast<uint32_t>(b) << 16);
}

__device__ __forceinline__ float sigmoid_ieee(float x) { return 1.f / (1.f + expf(-x)); }

template <int MODE>
__device__ __forceinline__ float combine(float in, float a0, float a1, float ch) {
  if (MODE == HYRES_SPLIT_COPY) return in;
  if (MODE == HYRES_SPLIT_ADD) return in + a0;
  if (MODE == HYRES_SPLIT_GATE) return a1 * sigmoid_ieee(in) + a0;
  if (MODE == HYRES_SPLIT_GDN) return a0 * (1.f / sqrtf(in));
  if (MODE == HYRES_SPLIT_IGDN) return a0 * sqrtf(in);
  if (MODE == HYRES_SPLIT_SQUARE) return in * in;
  return rintf(in - ch) + ch;  // HYRES_SPLIT_ROUND_CHAN
}

// rows x C fp32 (channel-contiguous) -> fp32 result and/or its bf16 parts [rows][nsplit*C]
template <int MODE>
__global__ void split_kernel(const float* __restrict__ in, const float* __restrict__ aux0,
                             const float* __restrict__ aux1, const float* __restrict__ chan, int relu,
                             float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_split, int nsplit,
                             int64_t rows, int C) {
  const int C4 = C >> 2;
  const int64_t total = rows * C4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  constexpr bool kAux0 = MODE == HYRES_SPLIT_ADD || MODE == HYRES_SPLIT_GATE || MODE == HYRES_SPLIT_GDN ||
                         MODE == HYRES_SPLIT_IGDN;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += stride) {
    const int64_t row = t / C4;
    const int c = static_cast<int>(t - row * C4) * 4;
    const int64_t e = row * C + c;
    const float4 x = __ldg(reinterpret_cast<const float4*>(in + e));
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, m = a;
    if (kAux0) a = __ldg(reinterpret_cast<const float4*>(aux0 + e));
    if (MODE == HYRES_SPLIT_GATE) b = __ldg(reinterpret_cast<const float4*>(aux1 + e));
    if (MODE == HYRES_SPLIT_ROUND_CHAN) m = __ldg(reinterpret_cast<const float4*>(chan + c));
    float v[4] = {combine<MODE>(x.x, a.x, b.x, m.x), combine<MODE>(x.y, a.y, b.y, m.y),
                  combine<MODE>(x.z, a.z, b.z, m.z), combine<MODE>(x.w, a.w, b.w, m.w)};
    if (relu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = fmaxf(v[k], 0.f);
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + e) = make_float4(v[0], v[1], v[2], v[3]);
    if (out_split) {
      const int P = nsplit & 15;
      const bool f16 = (nsplit & hy::kSplitF16) != 0;
      uint16_t p[3][4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split_code(v[k], f16, p[0][k], p[1][k], p[2][k]);
      __nv_bfloat16* o = out_split + row * (static_cast<int64_t>(P) * C) + c;
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (q < P)
          *reinterpret_cast<uint2*>(o + static_cast<int64_t>(q) * C) =
              make_uint2(pack2(p[q][0], p[q][1]), pack2(p[q][2], p[q][3]));
    }
  }
}

__global__ void sub_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
    o[i] = __ldg(a + i) - __ldg(b + i);
}

// 5x5 / stride 2 / pad 2 im2col of a 3-channel fp32 NCHW image as bf16 parts:
//   A[b, i, j, part*128 + (r*5+s)*3 + c] = part(x[b, c, 2i + r - 2, 2j + s - 2])   (0 outside, 0 for k >= 75)
// One thread writes 8 consecutive k of every part.
__global__ void im2col5s2_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a, int nsplit, int B,
                                       int H, int W) {
  const int OH = H / 2, OW = W / 2;
  constexpr int G = 128 / 8;
  const int64_t total = static_cast<int64_t>(B) * OH * OW * G;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t plane = static_cast<int64_t>(H) * W;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += stride) {
    const int g = static_cast<int>(t % G);
    const int64_t pix = t / G;
    const int j = static_cast<int>(pix % OW);
    const int64_t bi = pix / OW;
    const int i = static_cast<int>(bi % OH);
    const int b = static_cast<int>(bi / OH);
    const int P = nsplit & 15;
    const bool f16 = (nsplit & hy::kSplitF16) != 0;
    uint16_t p[3][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g * 8 + e;
      float v = 0.f;
      if (k < 75) {
        const int c = k % 3, rs = k / 3;
        const int r = rs / 5, s = rs - r * 5;
        const int ih = 2 * i + r - 2, iw = 2 * j + s - 2;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W)
          v = __ldg(x + (static_cast<int64_t>(b) * 3 + c) * plane + static_cast<int64_t>(ih) * W + iw);
      }
      split_code(v, f16, p[0][e], p[1][e], p[2][e]);
    }
    __nv_bfloat16* o = a + pix * (static_cast<int64_t>(P) * 128) + g * 8;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q < P)
        *reinterpret_cast<uint4*>(o + q * 128) = make_uint4(pack2(p[q][0], p[q][1]), pack2(p[q][2], p[q][3]),
                                                            pack2(p[q][4], p[q][5]), pack2(p[q][6], p[q][7]));
  }
}

// out[b, p, c] = float(sym[b, c, p]) + chan[c]: integer symbols in the coder's (B,C,h,w) order -> fp32 NHWC
__global__ void symbols_to_nhwc_kernel(const int32_t* __restrict__ sym, const float* __restrict__ chan,
                                       float* __restrict__ out, int hw, int C) {
  extern __shared__ int32_t itile[];  // [32][C+1]
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * 32;
  const int np = min(32, hw - p0);
  const int ld = C + 1;
  for (int t = threadIdx.x; t < np * C; t += kBlock) {
    const int c = t / np, px = t - c * np;
    itile[px * ld + c] = __ldg(sym + (static_cast<int64_t>(b) * C + c) * hw + p0 + px);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < np * C; t += kBlock) {
    const int px = t / C, c = t - px * C;
    out[(static_cast<int64_t>(b) * hw + p0 + px) * C + c] =
        static_cast<float>(itile[px * ld + c]) + (chan ? __ldg(chan + c) : 0.f);
  }
}

template <int MODE>
void launch_split(const float* in, const float* aux0, const float* aux1, const float* chan, int relu, float* out_f32,
                  void* out_split, int nsplit, int64_t rows, int C, cudaStream_t st) {
  const int64_t total = rows * (C / 4);
  split_kernel<MODE><<<grid_for(total, kBlock), kBlock, 0, st>>>(
      in, aux0, aux1, chan, relu, out_f32, static_cast<__nv_bfloat16*>(out_split), nsplit, rows, C);
}

}  // namespace

extern "C" {

int hyres_split_f32(const float* in, int64_t rows, int C, int mode, const float* aux0, const float* aux1,
                    const float* chan, int relu, float* out_f32, void* out_split, int nsplit, void* stream_v) {
  if (!in || rows <= 0 || C <= 0 || (C & 3))
    return hy_fail(HYRES_ERR_ARG, "split_f32: bad argument (C must be a multiple of 4)");
  if (!out_f32 && !out_split) return hy_fail(HYRES_ERR_ARG, "split_f32: no output");
  if (out_split && !hy_split_code_ok(nsplit))
    return hy_fail(HYRES_ERR_ARG, "split_f32: nsplit must be 1, 2, 3 or 2 | HYRES_SPLIT_F16");
  const bool need0 = mode == HYRES_SPLIT_ADD || mode == HYRES_SPLIT_GATE || mode == HYRES_SPLIT_GDN ||
                     mode == HYRES_SPLIT_IGDN;
  if (need0 && !aux0) return hy_fail(HYRES_ERR_ARG, "split_f32: aux0 missing");
  if (mode == HYRES_SPLIT_GATE && !aux1) return hy_fail(HYRES_ERR_ARG, "split_f32: aux1 missing");
  if (mode == HYRES_SPLIT_ROUND_CHAN && !chan) return hy_fail(HYRES_ERR_ARG, "split_f32: per-channel vector missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
#define HY_SPLIT_CASE(M) \
  case M: launch_split<M>(in, aux0, aux1, chan, relu, out_f32, out_split, nsplit, rows, C, st); break;
  switch (mode) {
    HY_SPLIT_CASE(HYRES_SPLIT_COPY)
    HY_SPLIT_CASE(HYRES_SPLIT_ADD)
    HY_SPLIT_CASE(HYRES_SPLIT_GATE)
    HY_SPLIT_CASE(HYRES_SPLIT_GDN)
    HY_SPLIT_CASE(HYRES_SPLIT_IGDN)
    HY_SPLIT_CASE(HYRES_SPLIT_SQUARE)
    HY_SPLIT_CASE(HYRES_SPLIT_ROUND_CHAN)
    default: return hy_fail(HYRES_ERR_ARG, "split_f32: unknown mode");
  }
#undef HY_SPLIT_CASE
  hy_count_launch();
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_residual_im2col5s2_split(const float* x, const float* jpeg, float* residual, void* a_out, int nsplit, int B,
                                   int H, int W, void* stream_v) {
  if (!x || !a_out || B <= 0 || H <= 0 || W <= 0 || ((H | W) & 1) || !hy_split_code_ok(nsplit))
    return hy_fail(HYRES_ERR_ARG, "residual_im2col5s2_split: bad argument");
  if (jpeg && !residual) return hy_fail(HYRES_ERR_ARG, "residual_im2col5s2_split: residual buffer missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const float* src = x;
  if (jpeg) {
    const int64_t n = static_cast<int64_t>(B) * 3 * H * W;
    hy_count_launch();
    sub_kernel<<<grid_for(n, kBlock), kBlock, 0, st>>>(x, jpeg, residual, n);
    HY_CUDA(cudaGetLastError());
    src = residual;
  }
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * (W / 2) * 16;
  hy_count_launch();
  im2col5s2_split_kernel<<<grid_for(total, kBlock), kBlock, 0, st>>>(src, static_cast<__nv_bfloat16*>(a_out), nsplit,
                                                                    B, H, W);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_symbols_to_nhwc_f32(const int32_t* symbols, const float* chan, float* out, int B, int h, int w, int C,
                              void* stream_v) {
  if (!symbols || !out || B <= 0 || h <= 0 || w <= 0 || C <= 0)
    return hy_fail(HYRES_ERR_ARG, "symbols_to_nhwc_f32: bad argument");
  const int smem = 32 * (C + 1) * 4;
  if (smem > 48 * 1024)
    HY_CUDA(cudaFuncSetAttribute(symbols_to_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((h * w + 31) / 32, B);
  hy_count_launch();
  symbols_to_nhwc_kernel<<<grid, kBlock, smem, static_cast<cudaStream_t>(stream_v)>>>(symbols, chan, out, h * w, C);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"
