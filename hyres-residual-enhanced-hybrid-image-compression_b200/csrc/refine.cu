// Memory-bound pieces of MultiScaleRefine (models/layers/enhancement.py:15-40,87-112):
//   * SEBlock: global average pool (deterministic two-stage), 64->4->64 FC, sigmoid, scale;
//     fused with the bilinear 1/2 (= 2x2 mean) and 1/4 (= mean of the centre 2x2 of each
//     4x4 block) down-samplings of the scaled feature map -- one read of feat, three writes;
//   * bilinear x2 / x4 up-sampling of the half / quarter branches written straight into
//     their channel slices of the 192-channel concat buffer, fused with the channel
//     mean / max of SpatialAttention -- the concat is never re-read for the statistics;
//   * the 7x7 (2->1) conv + sigmoid of SpatialAttention.  Its per-pixel scale is applied
//     in the epilogue of fusion.0 (HYRES_EPI_PIXSCALE): W*(multi*att) == att*(W*multi).
// NHWC bf16 activations, 16-byte (8-channel) vectors per thread.
#include <cstdint>
#include <cstring>

#include "common.cuh"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kPoolBlocks = 64;  // partial-sum blocks per image

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = hy::bf16_lo(u.x); f[1] = hy::bf16_hi(u.x);
  f[2] = hy::bf16_lo(u.y); f[3] = hy::bf16_hi(u.y);
  f[4] = hy::bf16_lo(u.z); f[5] = hy::bf16_hi(u.z);
  f[6] = hy::bf16_lo(u.w); f[7] = hy::bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = hy::pack_bf16(f[0], f[1]);
  u.y = hy::pack_bf16(f[2], f[3]);
  u.z = hy::pack_bf16(f[4], f[5]);
  u.w = hy::pack_bf16(f[6], f[7]);
  return u;
}
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// stage 1: per (image, slab) channel sums.  C == 64: 8 threads cover one position.
__global__ void se_pool_partial(const __nv_bfloat16* __restrict__ feat, float* __restrict__ partial, int hw) {
  const int b = blockIdx.y;
  const int slab = blockIdx.x;
  const int g = threadIdx.x & 7;        // channel group (8 channels)
  const int lane_pix = threadIdx.x >> 3;  // 32 positions per sweep
  const int per = (hw + kPoolBlocks - 1) / kPoolBlocks;
  const int p_begin = slab * per;
  const int p_end = min(hw, p_begin + per);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = p_begin + lane_pix; p < p_end; p += kThreads / 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(feat + (static_cast<int64_t>(b) * hw + p) * 64) + g);
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
  __shared__ float red[kThreads / 8][64 + 1];
#pragma unroll
  for (int k = 0; k < 8; ++k) red[lane_pix][g * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
    for (int r = 0; r < kThreads / 8; ++r) s += red[r][threadIdx.x];
    partial[(static_cast<int64_t>(b) * kPoolBlocks + slab) * 64 + threadIdx.x] = s;
  }
}

// stage 2: fixed-order reduction of the partials -> mean
__global__ void se_pool_final(const float* __restrict__ partial, float* __restrict__ pooled, int hw) {
  const int b = blockIdx.x;
  const int c = threadIdx.x;
  float s = 0.f;
  for (int k = 0; k < kPoolBlocks; ++k) s += partial[(static_cast<int64_t>(b) * kPoolBlocks + k) * 64 + c];
  pooled[b * 64 + c] = s / static_cast<float>(hw);
}

// feat_s = feat * sigmoid(fc2 relu(fc1 pooled)); feat_h = avg2x2(feat_s); feat_q = centre2x2 of 4x4
// One thread: a 4x4 block of positions x 8 channels.
__global__ void se_scale_down_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ pooled,
                                     const float* __restrict__ fc1, const float* __restrict__ fc2, int Cr,
                                     __nv_bfloat16* __restrict__ feat_s, __nv_bfloat16* __restrict__ feat_h,
                                     __nv_bfloat16* __restrict__ feat_q, int H, int W) {
  __shared__ float s_scale[64];
  __shared__ float s_hidden[16];
  const int b = blockIdx.y;
  const int bw = W / 4, bh = H / 4;
  const int64_t nblk = static_cast<int64_t>(bw) * bh * 8;
  const int64_t img = static_cast<int64_t>(b) * H * W;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  // the 16 loads of this thread's (first) block are in flight while the block evaluates the SE gate: the two tiny
  // matrix-vector products are ~130 dependent L2 accesses, which otherwise precede every block's first load
  uint4 u[4][4];
  auto load_block = [&](int64_t tt) {
    const int g = static_cast<int>(tt & 7);
    const int64_t blk = tt >> 3;
    const int bj = static_cast<int>(blk % bw), bi = static_cast<int>(blk / bw);
#pragma unroll
    for (int dy = 0; dy < 4; ++dy)
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int64_t pix = img + static_cast<int64_t>(bi * 4 + dy) * W + (bj * 4 + dx);
        u[dy][dx] = __ldg(reinterpret_cast<const uint4*>(feat + pix * 64) + g);
      }
  };
  if (t < nblk) load_block(t);
  if (threadIdx.x < Cr) {
    float a = 0.f;
    for (int c = 0; c < 64; ++c) a += fc1[threadIdx.x * 64 + c] * pooled[b * 64 + c];
    s_hidden[threadIdx.x] = fmaxf(a, 0.f);
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float a = 0.f;
    for (int r = 0; r < Cr; ++r) a += fc2[threadIdx.x * Cr + r] * s_hidden[r];
    s_scale[threadIdx.x] = 1.f / (1.f + expf(-a));
  }
  __syncthreads();
  for (; t < nblk; t += stride) {
    const int g = static_cast<int>(t & 7);
    const int64_t blk = t >> 3;
    const int bj = static_cast<int>(blk % bw), bi = static_cast<int>(blk / bw);
    float sc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sc[k] = s_scale[g * 8 + k];
    float half_acc[2][2][8];
    float q_acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      q_acc[k] = 0.f;
      half_acc[0][0][k] = half_acc[0][1][k] = half_acc[1][0][k] = half_acc[1][1][k] = 0.f;
    }
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int64_t pix = img + static_cast<int64_t>(bi * 4 + dy) * W + (bj * 4 + dx);
        float f[8];
        unpack8(u[dy][dx], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = round_bf16(f[k] * sc[k]);
        reinterpret_cast<uint4*>(feat_s + pix * 64)[g] = pack8(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          half_acc[dy >> 1][dx >> 1][k] += f[k];
          if ((dy == 1 || dy == 2) && (dx == 1 || dx == 2)) q_acc[k] += f[k];
        }
      }
    }
    if (t + stride < nblk) load_block(t + stride);  // next block of this thread, under the stores of this one
    const int Wh = W / 2, Wq = W / 4;
    const int64_t img_h = static_cast<int64_t>(b) * (H / 2) * Wh, img_q = static_cast<int64_t>(b) * (H / 4) * Wq;
#pragma unroll
    for (int hy_ = 0; hy_ < 2; ++hy_)
#pragma unroll
      for (int hx = 0; hx < 2; ++hx) {
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = half_acc[hy_][hx][k] * 0.25f;
        const int64_t pix = img_h + static_cast<int64_t>(bi * 2 + hy_) * Wh + (bj * 2 + hx);
        reinterpret_cast<uint4*>(feat_h + pix * 64)[g] = pack8(o);
      }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = q_acc[k] * 0.25f;
    reinterpret_cast<uint4*>(feat_q + (img_q + static_cast<int64_t>(bi) * Wq + bj) * 64)[g] = pack8(o);
  }
}

__global__ void replicate_border_kernel(uint4* __restrict__ t, int B, int Hp, int Wp, int c8) {
  const int per_img = 2 * Wp + 2 * (Hp - 2);
  const int64_t total = static_cast<int64_t>(B) * per_img * c8;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % c8);
  const int64_t r = i / c8;
  const int k = static_cast<int>(r % per_img);
  const int64_t b = r / per_img;
  int y, x;
  if (k < Wp) { y = 0; x = k; }
  else if (k < 2 * Wp) { y = Hp - 1; x = k - Wp; }
  else { const int q = k - 2 * Wp; y = 1 + (q >> 1); x = (q & 1) ? Wp - 1 : 0; }
  const int sy = min(max(y, 1), Hp - 2), sx = min(max(x, 1), Wp - 2);
  t[((b * Hp + y) * Wp + x) * c8 + c] = t[((b * Hp + sy) * Wp + sx) * c8 + c];
}

// att = sigmoid(conv7x7(stats; zero pad 3, no bias)); stats [B,H,W,2], weights [1][2][7][7]
// One thread: four horizontally adjacent outputs.  A 7x7 window per output read from shared memory costs 98 + 98
// loads (tile + weights) per output and made the kernel shared-memory bound (0.12 ms for 75 MB of traffic); with a
// sliding 10-wide row segment per kernel row (5 LDS.128) and the weights fetched four at a time it is 35 + 25 loads
// per FOUR outputs.  The sum runs in the same (r, s) order as before.
constexpr int kAttW = 128, kAttH = 8;
__global__ void __launch_bounds__(256) spatial_att_kernel(const float* __restrict__ stats, const float* __restrict__ w7,
                                                          float* __restrict__ att, int H, int W) {
  __shared__ __align__(16) float sw[100];
  __shared__ __align__(16) float2 tile[kAttH + 6][kAttW + 8];
  if (threadIdx.x < 98) sw[threadIdx.x] = w7[threadIdx.x];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kAttW, y0 = blockIdx.y * kAttH;
  const int64_t img = static_cast<int64_t>(b) * H * W;
  for (int t = threadIdx.x; t < (kAttH + 6) * (kAttW + 6); t += blockDim.x) {
    const int ty = t / (kAttW + 6), tx = t - ty * (kAttW + 6);
    const int y = y0 + ty - 3, x = x0 + tx - 3;
    float2 v = make_float2(0.f, 0.f);
    if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(reinterpret_cast<const float2*>(stats) + img + static_cast<int64_t>(y) * W + x);
    tile[ty][tx] = v;
  }
  __syncthreads();
  const int lx = (threadIdx.x & 31) * 4, ly = threadIdx.x >> 5;
  const int x = x0 + lx, y = y0 + ly;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    float2 row[10];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float4 q = *reinterpret_cast<const float4*>(&tile[ly + r][lx + 2 * i]);
      row[2 * i] = make_float2(q.x, q.y);
      row[2 * i + 1] = make_float2(q.z, q.w);
    }
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const float wm = sw[r * 7 + s], wx = sw[49 + r * 7 + s];
#pragma unroll
      for (int o = 0; o < 4; ++o) a[o] += wm * row[o + s].x + wx * row[o + s].y;
    }
  }
  if (y < H) {
#pragma unroll
    for (int o = 0; o < 4; ++o)
      if (x + o < W) att[img + static_cast<int64_t>(y) * W + x + o] = 1.f / (1.f + expf(-a[o]));
  }
}

}  // namespace

extern "C" {

int hyres_refine_se_pool(const void* feat, float* scratch, float* pooled, int B, int H, int W, int C, void* stream_v) {
  if (!feat || !scratch || !pooled || B <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "se_pool: bad argument");
  if (C != 64) return hy_fail(HYRES_ERR_UNSUPPORTED, "se_pool: C must be 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  hy_count_launch();
  se_pool_partial<<<dim3(kPoolBlocks, B), kThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(feat), scratch, H * W);
  hy_count_launch();
  se_pool_final<<<B, 64, 0, st>>>(scratch, pooled, H * W);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_refine_se_scale_down(const void* feat, const float* pooled, const float* fc1, const float* fc2, int C,
                               int Cr, void* feat_s, void* feat_h, void* feat_q, int B, int H, int W,
                               void* stream_v) {
  if (!feat || !pooled || !fc1 || !fc2 || !feat_s || !feat_h || !feat_q || B <= 0)
    return hy_fail(HYRES_ERR_ARG, "se_scale_down: bad argument");
  if (C != 64 || Cr < 1 || Cr > 16) return hy_fail(HYRES_ERR_UNSUPPORTED, "se_scale_down: C must be 64, Cr <= 16");
  if ((H & 3) || (W & 3)) return hy_fail(HYRES_ERR_ARG, "se_scale_down: H and W must be multiples of 4");
  const int64_t nblk = static_cast<int64_t>(H / 4) * (W / 4) * 8;
  // one resident wave (148 SMs x 8 blocks of 256 threads) over all images: the SE gate is evaluated once per block
  const int per_img = std::max(1, (148 * 8) / B);
  int gx = static_cast<int>(std::min<int64_t>((nblk + kThreads - 1) / kThreads, per_img));
  hy_count_launch();
  se_scale_down_kernel<<<dim3(gx, B), kThreads, 0, static_cast<cudaStream_t>(stream_v)>>>(
      static_cast<const __nv_bfloat16*>(feat), pooled, fc1, fc2, Cr, static_cast<__nv_bfloat16*>(feat_s),
      static_cast<__nv_bfloat16*>(feat_h), static_cast<__nv_bfloat16*>(feat_q), H, W);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_replicate_border(void* t, int B, int Hp, int Wp, int C, void* stream_v) {
  if (!t || B <= 0 || Hp < 3 || Wp < 3) return hy_fail(HYRES_ERR_ARG, "replicate_border: bad argument");
  if (C % 8) return hy_fail(HYRES_ERR_ARG, "replicate_border: C must be a multiple of 8");
  const int64_t total = static_cast<int64_t>(B) * (2 * Wp + 2 * (Hp - 2)) * (C / 8);
  hy_count_launch();
  replicate_border_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      static_cast<uint4*>(t), B, Hp, Wp, C / 8);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_refine_spatial_att(const float* stats, const float* w7x7, float* att, int B, int H, int W, void* stream_v) {
  if (!stats || !w7x7 || !att || B <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "spatial_att: bad argument");
  dim3 grid((W + kAttW - 1) / kAttW, (H + kAttH - 1) / kAttH, B);
  hy_count_launch();
  spatial_att_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(stats, w7x7, att, H, W);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"
