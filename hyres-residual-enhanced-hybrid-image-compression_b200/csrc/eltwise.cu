// HBM-bound kernels of the HyRES hot path: the final clamp, NHWC<->NCHW layout changes and loss
// reductions (the residual / add-back arithmetic is fused into the first-layer kernels, conv_c3.cu).
// All are coalesced and vectorised; none stages through shared memory unless it
// transposes.  Reference call sites are cited per entry point in include/hyres_b200.h.
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

__global__ void final_clamp_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o,
                                   int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i);
    const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
    float4 r;
    r.x = fminf(fmaxf(x.x + y.x, 0.f), 1.f);
    r.y = fminf(fmaxf(x.y + y.y, 0.f), 1.f);
    r.z = fminf(fmaxf(x.z + y.z, 0.f), 1.f);
    r.w = fminf(fmaxf(x.w + y.w, 0.f), 1.f);
    reinterpret_cast<float4*>(o)[i] = r;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
    o[i] = fminf(fmaxf(a[i] + b[i], 0.f), 1.f);
}

template <typename TIn, typename TOut>
__global__ void transpose_tiles(const TIn* __restrict__ in, TOut* __restrict__ out, int rows, int cols) {
  // in: [batch][rows][cols] -> out: [batch][cols][rows]
  __shared__ float tile[32][33];
  const int64_t boff = static_cast<int64_t>(blockIdx.z) * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int r = r0 + dy, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[dy][threadIdx.x] = static_cast<float>(in[boff + static_cast<int64_t>(r) * cols + c]);
  }
  __syncthreads();
  for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
    const int c = c0 + dy, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[boff + static_cast<int64_t>(c) * rows + r] = static_cast<TOut>(tile[threadIdx.x][dy]);
  }
}

__global__ void add_to_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                   __nv_bfloat16* __restrict__ o, int64_t n) {
  const int64_t n4 = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg(reinterpret_cast<const float4*>(a) + i);
    const float4 y = __ldg(reinterpret_cast<const float4*>(b) + i);
    uint2 r;
    r.x = hy::pack_bf16(x.x + y.x, x.y + y.y);
    r.y = hy::pack_bf16(x.z + y.z, x.w + y.w);
    reinterpret_cast<uint2*>(o)[i] = r;
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride)
    o[i] = __float2bfloat16_rn(a[i] + b[i]);
}

__device__ __forceinline__ void block_accumulate(double v, double* out) {
  __shared__ double part[kBlock / 32];
  v = hy::warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = threadIdx.x < kBlock / 32 ? part[threadIdx.x] : 0.0;
    s = hy::warp_sum_d(s);
    if (threadIdx.x == 0) hy::atomic_add_exact(out, s);
  }
}

// src/losses/rd_loss.py:23-44 from the three device sums (one thread; the float operations in the reference's order):
// out = [y_bpp, z_bpp, residual_bpp, bpp, mse * 255^2, lambda * mse + bpp]
__global__ void rd_loss_finalize_kernel(const double* __restrict__ sy, const double* __restrict__ sz,
                                        const double* __restrict__ se, const float* __restrict__ jpeg_bpp,
                                        double num_pixels, double num_elems, float lmbda, float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  const float y = static_cast<float>(-sy[0] / num_pixels);
  const float z = static_cast<float>(-sz[0] / num_pixels);
  const float res = y + z;
  const float bpp = res + (jpeg_bpp ? jpeg_bpp[0] : 0.f);
  const float mse = static_cast<float>(se[0] / num_elems) * 65025.f;
  out[0] = y; out[1] = z; out[2] = res; out[3] = bpp; out[4] = mse;
  out[5] = lmbda * mse + bpp;
}

__global__ void sqdiff_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, double* out) {
  float acc = 0.f;
  double dacc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int it = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    const float d = __ldg(a + i) - __ldg(b + i);
    acc = fmaf(d, d, acc);
    if ((++it & 63) == 0) { dacc += acc; acc = 0.f; }
  }
  block_accumulate(dacc + acc, out);
}

__global__ void log2_kernel(const float* __restrict__ x, int64_t n, double* out) {
  float acc = 0.f;
  double dacc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int it = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    acc += log2f(__ldg(x + i));
    if ((++it & 63) == 0) { dacc += acc; acc = 0.f; }
  }
  block_accumulate(dacc + acc, out);
}

}  // namespace

extern "C" {

int hyres_final_clamp(const float* x0, const float* refined, float* x_hat, int64_t n, void* stream_v) {
  if (!x0 || !refined || !x_hat || n < 0) return hy_fail(HYRES_ERR_ARG, "final_clamp: bad argument");
  if (n == 0) return HYRES_OK;
  hy_count_launch();
  final_clamp_kernel<<<grid_for(n, kBlock * 4), kBlock, 0, static_cast<cudaStream_t>(stream_v)>>>(x0, refined, x_hat, n);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_add_to_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream_v) {
  if (!a || !b || !out_bf16 || n < 0) return hy_fail(HYRES_ERR_ARG, "add_to_bf16: bad argument");
  if (n == 0) return HYRES_OK;
  hy_count_launch();
  add_to_bf16_kernel<<<grid_for(n, kBlock * 4), kBlock, 0, static_cast<cudaStream_t>(stream_v)>>>(
      a, b, static_cast<__nv_bfloat16*>(out_bf16), n);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_nchw_f32_to_nhwc_bf16(const float* in, void* out, int B, int C, int H, int W, void* stream_v) {
  if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "nchw_to_nhwc: bad argument");
  const int hw = H * W;
  dim3 grid((hw + 31) / 32, (C + 31) / 32, B), block(32, 8);
  hy_count_launch();
  transpose_tiles<float, __nv_bfloat16><<<grid, block, 0, static_cast<cudaStream_t>(stream_v)>>>(
      in, static_cast<__nv_bfloat16*>(out), C, hw);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_nhwc_to_nchw_f32(const float* in, float* out, int B, int C, int H, int W, void* stream_v) {
  if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "nhwc_to_nchw: bad argument");
  const int hw = H * W;
  dim3 grid((C + 31) / 32, (hw + 31) / 32, B), block(32, 8);
  hy_count_launch();
  transpose_tiles<float, float><<<grid, block, 0, static_cast<cudaStream_t>(stream_v)>>>(in, out, hw, C);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_nhwc_bf16_to_nchw_f32(const void* in, float* out, int B, int C, int H, int W, void* stream_v) {
  if (!in || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "nhwc_bf16_to_nchw: bad argument");
  const int hw = H * W;
  dim3 grid((C + 31) / 32, (hw + 31) / 32, B), block(32, 8);
  hy_count_launch();
  transpose_tiles<__nv_bfloat16, float><<<grid, block, 0, static_cast<cudaStream_t>(stream_v)>>>(
      static_cast<const __nv_bfloat16*>(in), out, hw, C);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_reduce_sqdiff(const float* a, const float* b, int64_t n, double* out, void* stream_v) {
  if (!a || !b || !out || n < 0) return hy_fail(HYRES_ERR_ARG, "reduce_sqdiff: bad argument");
  if (n == 0) return HYRES_OK;
  hy_count_launch();
  sqdiff_kernel<<<grid_for(n, kBlock * 8, 148 * 4), kBlock, 0, static_cast<cudaStream_t>(stream_v)>>>(a, b, n, out);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_rd_loss_finalize(const double* sum_log2_y, const double* sum_log2_z, const double* sum_sq_err,
                           const float* jpeg_bpp, double num_pixels, double num_elems, float lmbda, float* out6,
                           void* stream_v) {
  if (!sum_log2_y || !sum_log2_z || !sum_sq_err || !out6 || num_pixels <= 0 || num_elems <= 0)
    return hy_fail(HYRES_ERR_ARG, "rd_loss_finalize: bad argument");
  hy_count_launch();
  rd_loss_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_v)>>>(sum_log2_y, sum_log2_z, sum_sq_err, jpeg_bpp,
                                                                            num_pixels, num_elems, lmbda, out6);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_reduce_log2(const float* x, int64_t n, double* out, void* stream_v) {
  if (!x || !out || n < 0) return hy_fail(HYRES_ERR_ARG, "reduce_log2: bad argument");
  if (n == 0) return HYRES_OK;
  hy_count_launch();
  log2_kernel<<<grid_for(n, kBlock * 8, 148 * 4), kBlock, 0, static_cast<cudaStream_t>(stream_v)>>>(x, n, out);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"
