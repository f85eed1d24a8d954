// Entropy-model kernels: the checkerboard quantiser, GaussianConditional likelihood /
// symbols / CDF indexes, and the EntropyBottleneck (factorised prior).
//
// Reference behaviour restated (cited in include/hyres_b200.h):
//   * models/checkerboard.py:106-142 -- anchor / non-anchor tensors are zero-filled
//     *full-size* tensors and every position of both passes is quantised and coded;
//   * compressai GaussianConditional: scale lower bound 0.11, likelihood lower bound 1e-9,
//     Phi(t) = 0.5 * erfc(-t / sqrt(2)), indexes = #{table[j] < scale, j < n-1};
//   * compressai EntropyBottleneck: filters (3,3,3,3) cumulative MLP per channel.
//
// Layout: inputs are NHWC fp32 ([B,h,w,C], the conv epilogue's native layout);
// integer / likelihood outputs the API returns in NCHW are transposed through shared
// memory so both the read and the write side are coalesced (32 positions x C channels
// per block).  HBM-bound: no tensor cores.
#include <cstdint>

#include "common.cuh"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kPix = 32;  // positions per block
constexpr int kPixEb = 8; // positions per block of the (small, math-heavy) EntropyBottleneck kernels: 4x the blocks

// counter-based uniform noise in (-0.5, 0.5): splitmix64 of (seed, element index)
__device__ __forceinline__ float uniform_noise(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  // 24 random bits -> (0,1) open interval -> shift
  const float u = (static_cast<float>(static_cast<uint32_t>(z >> 40)) + 0.5f) * (1.0f / 16777216.0f);
  return u - 0.5f;
}

__device__ __forceinline__ float std_cumulative(float t) {
  // compressai: half * erfc(const * inputs), const = -(2 ** -0.5)
  return 0.5f * erfcf(-0.70710678118654752440f * t);
}

__device__ __forceinline__ float gauss_likelihood(float v_abs, float scale, float scale_bound, float lik_bound) {
  const float s = fmaxf(scale, scale_bound);
  const float upper = std_cumulative((0.5f - v_abs) / s);
  const float lower = std_cumulative((-0.5f - v_abs) / s);
  return fmaxf(upper - lower, lik_bound);
}

__device__ __forceinline__ int scale_index(float scale, float scale_bound, const float* __restrict__ table, int n) {
  const float s = fmaxf(scale, scale_bound);
  // bucketize(s, table[:n-1], right=False): number of entries strictly below s
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (table[mid] < s) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ bool is_anchor(int i, int j) { return ((i + j) & 1) == 0; }

// ---------------------------------------------------------------------------
// quantiser pass (forward): yq = round(part - mu) + mu   or   part + noise
// ---------------------------------------------------------------------------
__global__ void gc_quant_pass_kernel(const float* __restrict__ y, const float* __restrict__ params, int pass,
                                     int mode, uint64_t seed, float* __restrict__ yq_f32,
                                     __nv_bfloat16* __restrict__ yq_bf16, int64_t npix, int h, int w, int M) {
  const int M4 = M >> 2;
  const int64_t total = npix * M4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += stride) {
    const int64_t pix = t / M4;
    const int c = static_cast<int>(t - pix * M4) * 4;
    const int j = static_cast<int>(pix % w);
    const int i = static_cast<int>((pix / w) % h);
    const bool own = (pass == 0) == is_anchor(i, j);
    float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (own) yv = __ldg(reinterpret_cast<const float4*>(y + pix * M + c));
    float4 r;
    if (mode == 0) {
      const float4 mu = __ldg(reinterpret_cast<const float4*>(params + pix * 2 * M + M + c));
      r.x = rintf(yv.x - mu.x) + mu.x;
      r.y = rintf(yv.y - mu.y) + mu.y;
      r.z = rintf(yv.z - mu.z) + mu.z;
      r.w = rintf(yv.w - mu.w) + mu.w;
    } else {
      const uint64_t e = static_cast<uint64_t>(pix) * M + c;
      r.x = yv.x + uniform_noise(seed, e);
      r.y = yv.y + uniform_noise(seed, e + 1);
      r.z = yv.z + uniform_noise(seed, e + 2);
      r.w = yv.w + uniform_noise(seed, e + 3);
    }
    if (yq_f32) *reinterpret_cast<float4*>(yq_f32 + pix * M + c) = r;
    if (yq_bf16) {
      uint2 o;
      o.x = hy::pack_bf16(r.x, r.y);
      o.y = hy::pack_bf16(r.z, r.w);
      *reinterpret_cast<uint2*>(yq_bf16 + pix * M + c) = o;
    }
  }
}

// ---------------------------------------------------------------------------
// y_hat = yq_a + yq_na ; likelihood of y under summed params -> NCHW ; sum log2
// ---------------------------------------------------------------------------
__global__ void gc_merge_likelihood_kernel(const float* __restrict__ y, const float* __restrict__ pa,
                                           const float* __restrict__ pna, const float* __restrict__ yq_a,
                                           const float* __restrict__ yq_na, int mode, uint64_t seed,
                                           __nv_bfloat16* __restrict__ y_hat, float* __restrict__ lik_nchw,
                                           double* __restrict__ sum_log2, int hw, int M, float scale_bound,
                                           float lik_bound) {
  extern __shared__ float tile[];  // [kPix][M+1]
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPix;
  const int np = min(kPix, hw - p0);
  const int ld = M + 1;
  float acc = 0.f;
  const int M4 = M >> 2;  // M % 4 == 0 (checked by the entry point): four channels per thread, 16-byte loads
  for (int t = threadIdx.x; t < np * M4; t += kThreads) {
    const int px = t / M4, c = (t - px * M4) * 4;
    const int64_t pix = static_cast<int64_t>(b) * hw + p0 + px;
    const int64_t e = pix * M + c;
    const float4 yv = __ldg(reinterpret_cast<const float4*>(y + e));
    const float4 sa = __ldg(reinterpret_cast<const float4*>(pa + pix * 2 * M + c));
    const float4 sn = __ldg(reinterpret_cast<const float4*>(pna + pix * 2 * M + c));
    const float4 ma = __ldg(reinterpret_cast<const float4*>(pa + pix * 2 * M + M + c));
    const float4 mn = __ldg(reinterpret_cast<const float4*>(pna + pix * 2 * M + M + c));
    if (y_hat) {
      const float4 qa = __ldg(reinterpret_cast<const float4*>(yq_a + e)), qn = __ldg(reinterpret_cast<const float4*>(yq_na + e));
      uint2 o;
      o.x = hy::pack_bf16(qa.x + qn.x, qa.y + qn.y);
      o.y = hy::pack_bf16(qa.z + qn.z, qa.w + qn.w);
      *reinterpret_cast<uint2*>(y_hat + e) = o;
    }
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    const float ss[4] = {sa.x + sn.x, sa.y + sn.y, sa.z + sn.z, sa.w + sn.w};
    const float mm[4] = {ma.x + mn.x, ma.y + mn.y, ma.z + mn.z, ma.w + mn.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float outv;
      if (mode == 0) outv = rintf(yy[k] - mm[k]) + mm[k];
      else outv = yy[k] + uniform_noise(seed ^ 0xA5A5A5A5DEADBEEFull, static_cast<uint64_t>(e + k));
      const float lik = gauss_likelihood(fabsf(outv - mm[k]), ss[k], scale_bound, lik_bound);
      tile[px * ld + c + k] = lik;
      acc += log2f(lik);
    }
  }
  __syncthreads();
  if (lik_nchw) {
    for (int t = threadIdx.x; t < np * M; t += kThreads) {
      const int c = t / np, px = t - c * np;
      lik_nchw[(static_cast<int64_t>(b) * M + c) * hw + p0 + px] = tile[px * ld + c];
    }
  }
  if (sum_log2) {
    __shared__ float part[kThreads / 32];
    acc = hy::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int k = 0; k < kThreads / 32; ++k) s += part[k];
      hy::atomic_add_exact(sum_log2, s);
    }
  }
}

// ---------------------------------------------------------------------------
// symbols / indexes (compress) -- integer outputs in (B,M,h,w) order
// ---------------------------------------------------------------------------
// `rows` (optional): the host coder's packed-table layout [3][n_scales] = first entry, offset, escape bin of every CDF
// row (hyres_rans_table_layout).  With it the kernel also writes, in the indexes' order,
//   encoder (y given):  `slots`  -- base + (symbol - offset) for a value inside the table, -(row + 1) otherwise;
//   decoder (y == NULL): `slots` holds CODES -- at the structurally zero positions of pass `pass` the symbol is
//     round(-mean), known without decoding: bit 30 | packed entry when it is inside the table; the plain row index
//     at every other position.
__global__ void gc_symbols_kernel(const float* __restrict__ y, const float* __restrict__ params, int pass,
                                  const float* __restrict__ table, int n_scales, float scale_bound,
                                  int32_t* __restrict__ symbols, int32_t* __restrict__ indexes,
                                  float* __restrict__ yq_f32, __nv_bfloat16* __restrict__ yq_bf16, int h, int w,
                                  int M, const int32_t* __restrict__ rows, int32_t* __restrict__ slots) {
  extern __shared__ int32_t itile[];  // [2 or 3][kPix][M+1]
  __shared__ float s_table[64];
  __shared__ int32_t s_rows[3 * 64];
  if (rows && threadIdx.x < 3 * 64) {
    const int k = threadIdx.x / 64, r = threadIdx.x - k * 64;
    s_rows[threadIdx.x] = r < n_scales ? rows[k * n_scales + r] : 0;
  }
  const int hw = h * w;
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPix;
  const int np = min(kPix, hw - p0);
  const int ld = M + 1;
  int32_t* t_sym = itile;
  int32_t* t_idx = itile + kPix * ld;
  int32_t* t_slot = itile + 2 * kPix * ld;  // only there (and only touched) when `slots` is given
  if (threadIdx.x < 64) s_table[threadIdx.x] = threadIdx.x < n_scales ? table[threadIdx.x] : 3.4e38f;
  __syncthreads();
  for (int t = threadIdx.x; t < np * M; t += kThreads) {
    const int px = t / M, c = t - px * M;
    const int p = p0 + px;
    const int i = p / w, j = p - i * w;
    const int64_t pix = static_cast<int64_t>(b) * hw + p;
    const int64_t e = pix * M + c;
    const bool own = (y != nullptr) && ((pass == 0) == is_anchor(i, j));
    const float yv = own ? __ldg(y + e) : 0.f;
    const float sc = __ldg(params + pix * 2 * M + c);
    const float mu = __ldg(params + pix * 2 * M + M + c);
    const float q = rintf(yv - mu);
    const int32_t sym = static_cast<int32_t>(q);
    const int32_t ci = scale_index(sc, scale_bound, s_table, n_scales);
    if (slots) {
      const int32_t value = sym - s_rows[64 + ci];
      const bool inside = value >= 0 && value < s_rows[128 + ci];
      int32_t sl;
      if (y != nullptr) sl = inside ? s_rows[ci] + value : -(ci + 1);
      else sl = ((pass == 0) != is_anchor(i, j) && inside) ? ((1 << 30) | (s_rows[ci] + value)) : ci;
      t_slot[px * ld + c] = sl;
    }
    t_sym[px * ld + c] = sym;
    t_idx[px * ld + c] = ci;
    const float deq = q + mu;  // dequantize: float(symbol) + mean
    if (yq_f32) yq_f32[e] = deq;
    if (yq_bf16) yq_bf16[e] = __float2bfloat16_rn(deq);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < np * M; t += kThreads) {
    const int c = t / np, px = t - c * np;
    const int64_t o = (static_cast<int64_t>(b) * M + c) * hw + p0 + px;
    if (symbols) symbols[o] = t_sym[px * ld + c];
    if (indexes) indexes[o] = t_idx[px * ld + c];
    if (slots) slots[o] = t_slot[px * ld + c];
  }
}

// decoder: yq = float(symbol) + mean, symbols in (B,M,h,w) order -> NHWC
// pass >= 0: the symbols at the structurally zero positions of that checkerboard pass are not read (the decoder
// does not produce them, hyres_gc_codes): they are round(-mean) by construction (models/checkerboard.py:106-110, Q1).
__global__ void gc_dequant_kernel(const int32_t* __restrict__ symbols, const float* __restrict__ params,
                                  float* __restrict__ yq_f32, __nv_bfloat16* __restrict__ yq_bf16, int hw, int M,
                                  int pass, int w) {
  extern __shared__ int32_t itile[];  // [kPix][M+1]
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPix;
  const int np = min(kPix, hw - p0);
  const int ld = M + 1;
  for (int t = threadIdx.x; t < np * M; t += kThreads) {
    const int c = t / np, px = t - c * np;
    itile[px * ld + c] = __ldg(symbols + (static_cast<int64_t>(b) * M + c) * hw + p0 + px);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < np * M; t += kThreads) {
    const int px = t / M, c = t - px * M;
    const int64_t pix = static_cast<int64_t>(b) * hw + p0 + px;
    const int64_t e = pix * M + c;
    const float mu = __ldg(params + pix * 2 * M + M + c);
    float q = static_cast<float>(itile[px * ld + c]);
    if (pass >= 0) {
      const int p = p0 + px;
      const int i = p / w, j = p - i * w;
      if ((pass == 0) != is_anchor(i, j)) q = rintf(0.f - mu);
    }
    const float deq = q + mu;
    if (yq_f32) yq_f32[e] = deq;
    if (yq_bf16) yq_bf16[e] = __float2bfloat16_rn(deq);
  }
}

// ---------------------------------------------------------------------------
// EntropyBottleneck
// ---------------------------------------------------------------------------
struct EbChan {
  float m0[3], m1[9], m2[9], m3[9], m4[3];
  float b0[3], b1[3], b2[3], b3[3], b4[1];
  float f0[3], f1[3], f2[3], f3[3];
};
static_assert(sizeof(EbChan) == 58 * 4, "EbChan layout");

__device__ __forceinline__ float eb_logits(const EbChan& p, float x) {
  float h[3], g[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    h[k] = p.m0[k] * x + p.b0[k];
    h[k] += p.f0[k] * tanhf(h[k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g[k] = p.m1[3 * k] * h[0] + p.m1[3 * k + 1] * h[1] + p.m1[3 * k + 2] * h[2] + p.b1[k];
    g[k] += p.f1[k] * tanhf(g[k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    h[k] = p.m2[3 * k] * g[0] + p.m2[3 * k + 1] * g[1] + p.m2[3 * k + 2] * g[2] + p.b2[k];
    h[k] += p.f2[k] * tanhf(h[k]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g[k] = p.m3[3 * k] * h[0] + p.m3[3 * k + 1] * h[1] + p.m3[3 * k + 2] * h[2] + p.b3[k];
    g[k] += p.f3[k] * tanhf(g[k]);
  }
  return p.m4[0] * g[0] + p.m4[1] * g[1] + p.m4[2] * g[2] + p.b4[0];
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// mode bit0: likelihood input uses noise (module.training); bit1: z_hat output is the noisy value
__global__ void eb_forward_kernel(const float* __restrict__ z, const EbChan* __restrict__ prm,
                                  const float* __restrict__ medians, int mode, uint64_t seed, float lik_bound,
                                  __nv_bfloat16* __restrict__ zhat_bf16, float* __restrict__ zhat_nchw,
                                  float* __restrict__ lik_nchw, int32_t* __restrict__ symbols,
                                  double* __restrict__ sum_log2, int hw, int C) {
  extern __shared__ float tile[];  // [3][kPixEb][C+1]: lik, zhat, symbol(bits)
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPixEb;
  const int np = min(kPixEb, hw - p0);
  const int ld = C + 1;
  float* t_lik = tile;
  float* t_zh = tile + kPixEb * ld;
  int32_t* t_sym = reinterpret_cast<int32_t*>(tile + 2 * kPixEb * ld);
  float acc = 0.f;
  for (int t = threadIdx.x; t < np * C; t += kThreads) {
    const int px = t / C, c = t - px * C;
    const int64_t e = (static_cast<int64_t>(b) * hw + p0 + px) * C + c;
    const float zv = __ldg(z + e);
    const float med = __ldg(medians + c);
    const float q = rintf(zv - med);
    const float deq = q + med;
    const float noisy = zv + uniform_noise(seed, static_cast<uint64_t>(e));
    const float v = (mode & 1) ? noisy : deq;
    const float zh = (mode & 2) ? noisy : deq;
    const EbChan& p = prm[c];
    const float lower = eb_logits(p, v - 0.5f);
    const float upper = eb_logits(p, v + 0.5f);
    const float lik = fmaxf(sigmoidf_acc(upper) - sigmoidf_acc(lower), lik_bound);
    acc += log2f(lik);
    t_lik[px * ld + c] = lik;
    t_zh[px * ld + c] = zh;
    t_sym[px * ld + c] = static_cast<int32_t>(q);
    if (zhat_bf16) zhat_bf16[e] = __float2bfloat16_rn(zh);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < np * C; t += kThreads) {
    const int c = t / np, px = t - c * np;
    const int64_t o = (static_cast<int64_t>(b) * C + c) * hw + p0 + px;
    if (lik_nchw) lik_nchw[o] = t_lik[px * ld + c];
    if (zhat_nchw) zhat_nchw[o] = t_zh[px * ld + c];
    if (symbols) symbols[o] = t_sym[px * ld + c];
  }
  if (sum_log2) {
    __shared__ float part[kThreads / 32];
    acc = hy::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int k = 0; k < kThreads / 32; ++k) s += part[k];
      hy::atomic_add_exact(sum_log2, s);
    }
  }
}

__global__ void eb_dequant_kernel(const int32_t* __restrict__ symbols, const float* __restrict__ medians,
                                  __nv_bfloat16* __restrict__ zhat, int hw, int C) {
  extern __shared__ int32_t itile[];
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kPixEb;
  const int np = min(kPixEb, hw - p0);
  const int ld = C + 1;
  for (int t = threadIdx.x; t < np * C; t += kThreads) {
    const int c = t / np, px = t - c * np;
    itile[px * ld + c] = __ldg(symbols + (static_cast<int64_t>(b) * C + c) * hw + p0 + px);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < np * C; t += kThreads) {
    const int px = t / C, c = t - px * C;
    const int64_t e = (static_cast<int64_t>(b) * hw + p0 + px) * C + c;
    zhat[e] = __float2bfloat16_rn(static_cast<float>(itile[px * ld + c]) + __ldg(medians + c));
  }
}

int set_smem(const void* fn, int bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return hy_fail(HYRES_ERR_CUDA, cudaGetErrorString(e));
  }
  return HYRES_OK;
}

}  // namespace

extern "C" {

int hyres_gc_quant_pass(const float* y, const float* params, int pass, int mode, uint64_t seed, float* yq_f32,
                        void* yq_bf16, int B, int h, int w, int M, void* stream_v) {
  if (!y || (!params && mode == 0) || B <= 0 || h <= 0 || w <= 0 || M <= 0 || (M & 3))
    return hy_fail(HYRES_ERR_ARG, "gc_quant_pass: bad argument");
  const int64_t npix = static_cast<int64_t>(B) * h * w;
  const int64_t total = npix * (M / 4);
  int grid = static_cast<int>(std::min<int64_t>((total + kThreads - 1) / kThreads, 148 * 16));
  hy_count_launch();
  gc_quant_pass_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_v)>>>(
      y, params, pass, mode, seed, yq_f32, static_cast<__nv_bfloat16*>(yq_bf16), npix, h, w, M);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_gc_merge_likelihood(const float* y, const float* params_a, const float* params_na, const float* yq_a,
                              const float* yq_na, int mode, uint64_t seed, void* y_hat_bf16, float* lik_nchw,
                              double* sum_log2, int B, int h, int w, int M, void* stream_v) {
  if (!y || !params_a || !params_na || B <= 0 || h <= 0 || w <= 0 || M <= 0 || (M & 3))
    return hy_fail(HYRES_ERR_ARG, "gc_merge_likelihood: bad argument (M must be a positive multiple of 4)");
  if (y_hat_bf16 && (!yq_a || !yq_na)) return hy_fail(HYRES_ERR_ARG, "gc_merge_likelihood: yq_a/yq_na required");
  const int hw = h * w;
  const int smem = kPix * (M + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(gc_merge_likelihood_kernel), smem);
  if (rc) return rc;
  dim3 grid((hw + kPix - 1) / kPix, B);
  hy_count_launch();
  gc_merge_likelihood_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      y, params_a, params_na, yq_a, yq_na, mode, seed, static_cast<__nv_bfloat16*>(y_hat_bf16), lik_nchw, sum_log2,
      hw, M, 0.11f, 1e-9f);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_gc_symbols(const float* y, const float* params, int pass, const float* scale_table, int n_scales,
                     float scale_bound, int32_t* symbols, int32_t* indexes, float* yq_f32, void* yq_bf16, int B,
                     int h, int w, int M, const int32_t* coder_rows, int32_t* slots, void* stream_v) {
  if (!y || !params || !scale_table || n_scales < 2 || n_scales > 64 || B <= 0 || h <= 0 || w <= 0 || M <= 0)
    return hy_fail(HYRES_ERR_ARG, "gc_symbols: bad argument");
  if ((slots != nullptr) != (coder_rows != nullptr)) return hy_fail(HYRES_ERR_ARG, "gc_symbols: slots and coder_rows go together");
  const int smem = (slots ? 3 : 2) * kPix * (M + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(gc_symbols_kernel), smem);
  if (rc) return rc;
  dim3 grid((h * w + kPix - 1) / kPix, B);
  hy_count_launch();
  gc_symbols_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      y, params, pass, scale_table, n_scales, scale_bound, symbols, indexes, yq_f32,
      static_cast<__nv_bfloat16*>(yq_bf16), h, w, M, coder_rows, slots);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_gc_codes(const float* params, int pass, const float* scale_table, int n_scales, float scale_bound,
                   const int32_t* coder_rows, int32_t* indexes, int32_t* codes, int B, int h, int w, int M,
                   void* stream_v) {
  if (!params || !scale_table || !coder_rows || !codes || n_scales < 2 || n_scales > 64 || B <= 0 || h <= 0 || w <= 0 ||
      M <= 0 || (pass != 0 && pass != 1))
    return hy_fail(HYRES_ERR_ARG, "gc_codes: bad argument");
  const int smem = 3 * kPix * (M + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(gc_symbols_kernel), smem);
  if (rc) return rc;
  dim3 grid((h * w + kPix - 1) / kPix, B);
  hy_count_launch();
  gc_symbols_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      nullptr, params, pass, scale_table, n_scales, scale_bound, nullptr, indexes, nullptr, nullptr, h, w, M, coder_rows,
      codes);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_gc_indexes(const float* params, const float* scale_table, int n_scales, float scale_bound,
                     int32_t* indexes, int B, int h, int w, int M, void* stream_v) {
  if (!params || !scale_table || !indexes || n_scales < 2 || n_scales > 64 || B <= 0 || h <= 0 || w <= 0 || M <= 0)
    return hy_fail(HYRES_ERR_ARG, "gc_indexes: bad argument");
  const int smem = 2 * kPix * (M + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(gc_symbols_kernel), smem);
  if (rc) return rc;
  dim3 grid((h * w + kPix - 1) / kPix, B);
  hy_count_launch();
  gc_symbols_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      nullptr, params, 0, scale_table, n_scales, scale_bound, nullptr, indexes, nullptr, nullptr, h, w, M, nullptr,
      nullptr);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_gc_dequant(const int32_t* symbols, const float* params, float* yq_f32, void* yq_bf16, int B, int h,
                     int w, int M, int pass, void* stream_v) {
  if (!symbols || !params || B <= 0 || h <= 0 || w <= 0 || M <= 0)
    return hy_fail(HYRES_ERR_ARG, "gc_dequant: bad argument");
  const int smem = kPix * (M + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(gc_dequant_kernel), smem);
  if (rc) return rc;
  dim3 grid((h * w + kPix - 1) / kPix, B);
  hy_count_launch();
  gc_dequant_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      symbols, params, yq_f32, static_cast<__nv_bfloat16*>(yq_bf16), h * w, M, pass, w);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_eb_forward(const float* z, const float* eb_params, const float* medians, int mode, uint64_t seed,
                     float lik_bound, void* zhat_bf16, float* zhat_nchw, float* lik_nchw, int32_t* symbols,
                     double* sum_log2, int B, int h, int w, int C, void* stream_v) {
  if (!z || !eb_params || !medians || B <= 0 || h <= 0 || w <= 0 || C <= 0)
    return hy_fail(HYRES_ERR_ARG, "eb_forward: bad argument");
  const int smem = 3 * kPixEb * (C + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(eb_forward_kernel), smem);
  if (rc) return rc;
  dim3 grid((h * w + kPixEb - 1) / kPixEb, B);
  hy_count_launch();
  eb_forward_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      z, reinterpret_cast<const EbChan*>(eb_params), medians, mode, seed, lik_bound,
      static_cast<__nv_bfloat16*>(zhat_bf16), zhat_nchw, lik_nchw, symbols, sum_log2, h * w, C);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_eb_dequant(const int32_t* symbols, const float* medians, void* zhat_bf16, int B, int h, int w, int C,
                     void* stream_v) {
  if (!symbols || !medians || !zhat_bf16 || B <= 0 || h <= 0 || w <= 0 || C <= 0)
    return hy_fail(HYRES_ERR_ARG, "eb_dequant: bad argument");
  const int smem = kPixEb * (C + 1) * 4;
  int rc = set_smem(reinterpret_cast<const void*>(eb_dequant_kernel), smem);
  if (rc) return rc;
  dim3 grid((h * w + kPixEb - 1) / kPixEb, B);
  hy_count_launch();
  eb_dequant_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream_v)>>>(
      symbols, medians, static_cast<__nv_bfloat16*>(zhat_bf16), h * w, C);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"
