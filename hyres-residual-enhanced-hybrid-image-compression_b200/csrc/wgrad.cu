// Weight gradient of a convolution on the tcgen05 tensor cores (training step, src/utils/engine.py:50-53:
// loss.backward() reaches every nn.Conv2d / nn.ConvTranspose2d of models/checkerboard.py:35-88 and
// models/layers/enhancement.py:60-85).
//
//   dW[co, ci, r, s] = sum over (b, oh, ow) of  g[b, oh, ow, co] * x[b, oh*stride + r*dil - pad, ow*stride + s*dil - pad, ci]
//
// is a GEMM whose reduction dimension is the output POSITION: M = co, N = ci, K = positions.  Both operands live in
// HBM as NHWC bf16, i.e. with the reduction dimension outermost -- for the tensor core both are "MN-major" operands
// (instruction descriptor bits 15 / 16), and a TMA box (64 channels x 8 columns x rows) lands in shared memory in
// exactly the canonical MN-major SWIZZLE_128B layout: 64-channel atoms of eight 128-byte K rows, atoms along K
// 1024 B apart (SBO), the second 64-channel atom of M one box further (LBO).  No transposition anywhere.
//
//  * one CTA owns a block of 128 output channels, up to eight (64-input-channel, tap) accumulators of 128 x 64 fp32
//    in tensor memory (all 512 columns), and a contiguous range of 16 x 8 position tiles; the accumulators stay in
//    TMEM for the CTA's whole life;
//  * per tile the producer warp loads the gradient tile (two boxes) and, per kernel column, one halo patch of the
//    input; every vertical tap is that patch shifted by whole image rows (1 KB), exactly the forward kernel's trick,
//    and the forward layer's own tap-group table drives it (stride-2 layers read the same 5-D parity view);
//  * image borders, ragged tiles and channel counts below 128 / 64 are TMA out-of-bounds zero fill;
//  * CTAs that share an accumulator but cover different position ranges write partial sums, which a second kernel
//    adds in a fixed order straight into PyTorch's weight layout (deterministic, no atomics).
// A transposed convolution's weight gradient is the same kernel with the roles of input and output gradient
// swapped (it is the weight gradient of the stride-2 convolution that is its data-gradient).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kTileW = 8, kTileH = 16;  // 128 output positions per tile = K of 128 per tile
constexpr int kThreads = 192;           // warps 0-3: epilogue, warp 4: TMA producer, warp 5: MMA issuer
constexpr int kMaxAcc = 8;              // 8 accumulators x 64 columns = the 512 TMEM columns
constexpr int kMaxGroups = 3;           // input patches per tile
constexpr int kABytes = 2 * kTileH * kTileW * 128;  // gradient tile: two 64-channel boxes
constexpr int kSmemLimit = 227 * 1024;

struct WgItem {
  int32_t g_first, g_count, n_acc;
  int32_t acc_first[kMaxGroups];
  int32_t kslot[kMaxAcc];
};

struct alignas(64) WgParams {
  CUtensorMap mapG;
  CUtensorMap mapX;
  const TapGroup* groups;
  const WgItem* items;
  float* partial;
  int32_t n_items, n_co_blocks, n_split, n_kslots;
  int32_t tiles_w, tiles_h, ntiles;
  int32_t b_bytes, stage_bytes, nstage;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
  // MN-major SWIZZLE_128B operand: LBO = byte distance between 64-element atoms along M / N, SBO = 1024 (8 K rows)
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + p.nstage * p.stage_bytes;
  const uint32_t full = bar_base, empty = full + 8 * p.nstage, done = empty + 8 * p.nstage, tmem_slot = done + 8;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int item_i = bid % p.n_items;
  bid /= p.n_items;
  const int cb = bid % p.n_co_blocks;
  const int split = bid / p.n_co_blocks;
  const WgItem it = p.items[item_i];
  const int t0 = static_cast<int>(static_cast<long long>(split) * p.ntiles / p.n_split);
  const int t1 = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / p.n_split);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      hy::mbar_init(full + 8 * i, 1);
      hy::mbar_init(empty + 8 * i, 1);
    }
    hy::mbar_init(done, 1);
    hy::mbar_fence_init();
  }
  if (warp == 4 && lane == 0) {
    hy::tma_prefetch_desc(&p.mapG);
    hy::tma_prefetch_desc(&p.mapX);
  }
  if (warp == 5) {
    hy::tmem_alloc(tmem_slot, 512);
    hy::tmem_relinquish();
  }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
  uint32_t tmem_base_v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base_v) : "r"(tmem_slot));
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_v, 0);

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t bytes = kABytes + it.g_count * p.b_bytes;
      for (int t = t0; t < t1; ++t) {
        const int b_img = t / tiles_per_img;
        const int r = t - b_img * tiles_per_img;
        const int th = r / p.tiles_w;
        const int h0 = th * kTileH, w0 = (r - th * p.tiles_w) * kTileW;
        hy::mbar_wait(empty + 8 * s, ph ^ 1u);
        hy::mbar_arrive_expect_tx(full + 8 * s, bytes);
        const uint32_t st = base + s * p.stage_bytes;
        hy::tma_load_5d(st, &p.mapG, full + 8 * s, cb * 128, w0, 0, h0, b_img);
        hy::tma_load_5d(st + kABytes / 2, &p.mapG, full + 8 * s, cb * 128 + 64, w0, 0, h0, b_img);
        for (int g = 0; g < it.g_count; ++g) {
          const TapGroup tg = p.groups[it.g_first + g];
          hy::tma_load_5d(st + kABytes + g * p.b_bytes, &p.mapX, full + 8 * s, tg.c_off, w0 + tg.dw, tg.hpar,
                          h0 + tg.dh, b_img);
        }
        if (++s == p.nstage) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (converged warp, one elected lane issues) =====================
    const uint32_t leader = hy::elect_leader();
    // M = 128 (both 64-channel atoms of the gradient tile), N = 64, A and B MN-major
    const uint32_t idesc = hy::umma_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
    int s = 0;
    uint32_t ph = 0;
    for (int t = t0; t < t1; ++t) {
      hy::mbar_wait(full + 8 * s, ph);
      hy::tc_fence_after();
      const uint32_t st = base + s * p.stage_bytes;
      const uint32_t acc_flag = t > t0 ? 1u : 0u;
      for (int g = 0; g < it.g_count; ++g) {
        const int gi = it.g_first + g;
        const int ntaps = p.groups[gi].ntaps;
        const uint32_t bst = st + kABytes + g * p.b_bytes;
        for (int tp = 0; tp < ntaps; ++tp) {
          const uint32_t d_tmem = tmem_base + (it.acc_first[g] + tp) * 64;
          const uint32_t brow = bst + p.groups[gi].tap_row[tp] * (kTileW * 128);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            hy::umma_issue<2>(d_tmem, desc_mn(st + k * 2048, kABytes / 2), desc_mn(brow + k * 2048, 1024), idesc,
                              acc_flag | static_cast<uint32_t>(k), leader);
        }
      }
      hy::umma_commit_mode<2>(empty + 8 * s, leader);
      if (++s == p.nstage) { s = 0; ph ^= 1u; }
    }
    hy::umma_commit_mode<2>(done, leader);
  } else {
    // ===================== epilogue: accumulators -> partial sums =====================
    hy::mbar_wait(done, 0);
    hy::tc_fence_after();
    const int row = warp * 32 + lane;  // output channel within the block == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    for (int a = 0; a < it.n_acc; ++a) {
      float* o = p.partial + ((static_cast<size_t>(split) * p.n_kslots + it.kslot[a]) * p.n_co_blocks + cb) * (128 * 64) +
                 row * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        hy::tmem_ld32(t_lane + a * 64 + half * 32, r);
        hy::tmem_ld_fence32(r);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          reinterpret_cast<float4*>(o + half * 32)[i] =
              make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                          __uint_as_float(r[4 * i + 3]));
      }
    }
  }
  hy::tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, 512);
  }
}

// dW[co][ci][r][s] (PyTorch layout of the planned convolution) = sum over position splits, in split order
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const hyres_conv::Slot* __restrict__ slots,
                                    int n_split, int n_kslots, int n_co_blocks, int cout, int cin, int RS, int S,
                                    float* __restrict__ out) {
  const long long per_split = static_cast<long long>(n_kslots) * n_co_blocks * 128 * 64;
  const long long total = per_split;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cil = static_cast<int>(idx & 63);
    const int row = static_cast<int>((idx >> 6) & 127);
    long long rest = idx >> 13;
    const int cb = static_cast<int>(rest % n_co_blocks);
    const int ks = static_cast<int>(rest / n_co_blocks);
    const hyres_conv::Slot sl = slots[ks];
    const int co = cb * 128 + row, ci = sl.chunk * 64 + cil;
    if (co >= cout || ci >= cin) continue;
    float acc = 0.f;
    for (int sp = 0; sp < n_split; ++sp) acc += partial[sp * per_split + idx];
    out[(static_cast<size_t>(co) * cin + ci) * RS + sl.r * S + sl.s] = acc;
  }
}

int encode_map(CUtensorMap* m, const void* ptr, int C, int B, int H, int W, int stride2, int box_rows) {
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
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(wgrad C=%d B=%d H=%d W=%d s2=%d rows=%d) -> %d", C, B, H, W, stride2,
             box_rows, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
  return HYRES_OK;
}

struct Plan {
  std::vector<WgItem> items;
  int n_co_blocks, n_kslots, patch_rows, b_bytes, stage_bytes, nstage;
};

bool supported(const hyres_conv* c) {
  return c && c->kind == HYRES_CONV && c->nsplit == 1 && c->cin1 == 0 && (c->cin0 % 8) == 0 && (c->cout % 8) == 0 &&
         c->nphase == 1;
}

void make_plan(const hyres_conv* c, Plan& pl) {
  pl.items.clear();
  WgItem cur{};
  cur.g_first = 0;
  const int ng = c->ph_count[0];
  for (int g = 0; g < ng; ++g) {
    const TapGroup& tg = c->groups[c->ph_begin[0] + g];
    if (cur.g_count == kMaxGroups || cur.n_acc + tg.ntaps > kMaxAcc) {
      pl.items.push_back(cur);
      cur = WgItem{};
      cur.g_first = g;
    }
    cur.acc_first[cur.g_count] = cur.n_acc;
    for (int t = 0; t < tg.ntaps; ++t) cur.kslot[cur.n_acc++] = tg.kslot0 + t;
    ++cur.g_count;
  }
  if (cur.g_count) pl.items.push_back(cur);
  pl.n_co_blocks = (c->cout + 127) / 128;
  pl.n_kslots = static_cast<int>(c->slots.size());
  pl.patch_rows = kTileH + c->extra_rows;
  pl.b_bytes = pl.patch_rows * kTileW * 128;
  pl.stage_bytes = kABytes + kMaxGroups * pl.b_bytes;
  int maxg = 1;
  for (auto& it : pl.items) maxg = std::max(maxg, it.g_count);
  pl.stage_bytes = kABytes + maxg * pl.b_bytes;
  pl.nstage = std::min(6, (kSmemLimit - 2048) / pl.stage_bytes);
}

int choose_split(const Plan& pl, int ntiles) {
  const int ctas = static_cast<int>(pl.items.size()) * pl.n_co_blocks;
  int split = (2 * num_sms() + ctas - 1) / ctas;
  split = std::max(1, std::min(split, std::min(ntiles, 64)));
  return split;
}

}  // namespace

extern "C" {

int hyres_wgrad_supported(const hyres_conv* c) { return supported(c) ? 1 : 0; }

int64_t hyres_wgrad_workspace_bytes(const hyres_conv* c, int B, int H, int W) {
  if (!supported(c) || B <= 0 || H <= 0 || W <= 0) return 0;
  Plan pl;
  make_plan(c, pl);
  int OH, OW;
  hyres_conv_out_size(c, H, W, &OH, &OW);
  const int ntiles = B * ((OH + kTileH - 1) / kTileH) * ((OW + kTileW - 1) / kTileW);
  const int split = choose_split(pl, ntiles);
  return static_cast<int64_t>(split) * pl.n_kslots * pl.n_co_blocks * 128 * 64 * 4;
}

int hyres_wgrad_run(hyres_conv* c, const void* x, const void* gout, int B, int H, int W, float* dw, void* workspace,
                    void* stream_v) {
  if (!supported(c)) return hy_fail(HYRES_ERR_UNSUPPORTED, "wgrad_run: layer not supported");
  if (!x || !gout || !dw || !workspace || B <= 0 || H <= 0 || W <= 0) return hy_fail(HYRES_ERR_ARG, "wgrad_run: bad argument");
  if (c->stride == 2 && ((H | W) & 1)) return hy_fail(HYRES_ERR_ARG, "wgrad_run: stride-2 layers need even H and W");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  Plan pl;
  make_plan(c, pl);
  if (pl.nstage < 2) return hy_fail(HYRES_ERR_UNSUPPORTED, "wgrad_run: tile does not fit shared memory");
  int OH, OW;
  hyres_conv_out_size(c, H, W, &OH, &OW);
  WgParams p;
  memset(&p, 0, sizeof p);
  p.tiles_w = (OW + kTileW - 1) / kTileW;
  p.tiles_h = (OH + kTileH - 1) / kTileH;
  p.ntiles = B * p.tiles_w * p.tiles_h;
  p.n_items = static_cast<int>(pl.items.size());
  p.n_co_blocks = pl.n_co_blocks;
  p.n_kslots = pl.n_kslots;
  p.n_split = choose_split(pl, p.ntiles);
  p.b_bytes = pl.b_bytes;
  p.stage_bytes = pl.stage_bytes;
  p.nstage = pl.nstage;
  p.groups = c->d_groups + c->ph_begin[0];
  // the plan's item table lives with the layer (uploaded on first use: a step that is being captured into a CUDA
  // graph must not copy from pageable host memory); workspace = partial sums
  if (!c->d_wg_items) {
    HY_CUDA(cudaMalloc(&c->d_wg_items, pl.items.size() * sizeof(WgItem)));
    HY_CUDA(cudaMemcpy(c->d_wg_items, pl.items.data(), pl.items.size() * sizeof(WgItem), cudaMemcpyHostToDevice));
  }
  p.items = static_cast<const WgItem*>(c->d_wg_items);
  p.partial = static_cast<float*>(workspace);
  if (!c->d_slots) {
    HY_CUDA(cudaMalloc(&c->d_slots, c->slots.size() * sizeof(hyres_conv::Slot)));
    HY_CUDA(cudaMemcpy(c->d_slots, c->slots.data(), c->slots.size() * sizeof(hyres_conv::Slot), cudaMemcpyHostToDevice));
  }
  int rc = encode_map(&p.mapG, gout, c->cout, B, OH, OW, 0, kTileH);
  if (rc != HYRES_OK) return rc;
  rc = encode_map(&p.mapX, x, c->cin0, B, H, W, c->stride == 2 ? 1 : 0, pl.patch_rows);
  if (rc != HYRES_OK) return rc;
  const int smem = p.nstage * p.stage_bytes + 2048;
  static HyPerDevice attr;
  if (!attr.done()) {
    HY_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr.mark();
  }
  const int grid = p.n_items * p.n_co_blocks * p.n_split;
  hy_count_launch();
  wgrad_kernel<<<grid, kThreads, smem, st>>>(p);
  HY_CUDA(cudaGetLastError());
  HY_CUDA(cudaMemsetAsync(dw, 0, static_cast<size_t>(c->cout) * c->w_cin_total * c->R * c->S * sizeof(float), st));
  const long long total = static_cast<long long>(p.n_kslots) * p.n_co_blocks * 128 * 64;
  hy_count_launch();
  wgrad_reduce_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 8)), 256, 0, st>>>(
      p.partial, c->d_slots, p.n_split, p.n_kslots, p.n_co_blocks, c->cout, c->w_cin_total, c->R * c->S, c->S, dw);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Bias gradient: db[c] = sum over rows of g[row][c] (g: bf16 [rows][C], C a multiple of 8).  HBM-bound column sums:
// a block owns a slab of rows, a thread 8 channels (16-byte loads) of every (blockDim.y)-th row; slabs are reduced in
// shared memory, slab partials in a second, fixed-order pass (deterministic).
// ---------------------------------------------------------------------------------------------------------------
namespace {

__global__ void colsum_partial_kernel(const __nv_bfloat16* __restrict__ g, long long rows, int C, int rows_per_block,
                                      float* __restrict__ partial) {
  extern __shared__ float red[];  // [blockDim.y][C]
  const int c8 = threadIdx.x;     // group of 8 channels
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c8 * 8 < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(g + r * C) + c8);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += hy::bf16_lo(u[i]);
        acc[2 * i + 1] += hy::bf16_hi(u[i]);
      }
    }
  if (c8 * 8 < C)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.y * C + c8 * 8 + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < C; c += blockDim.x * blockDim.y) {
    float s = 0.f;
    for (int y = 0; y < blockDim.y; ++y) s += red[y * C + c];
    partial[static_cast<long long>(blockIdx.x) * C + c] = s;
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[static_cast<long long>(b) * C + c];
  out[c] = s;
}

}  // namespace

extern "C" {

int64_t hyres_colsum_workspace_bytes(int64_t rows, int C) {
  if (rows <= 0 || C <= 0) return 0;
  const int64_t nblocks = std::min<int64_t>(148 * 4, (rows + 63) / 64);
  return nblocks * C * 4;
}

int hyres_colsum_bf16(const void* g, int64_t rows, int C, float* out, void* workspace, void* stream_v) {
  if (!g || !out || !workspace || rows <= 0 || C <= 0 || (C % 8) || C > 2048)
    return hy_fail(HYRES_ERR_ARG, "colsum_bf16: bad argument (C must be a multiple of 8, at most 2048)");
  cudaStream_t st = static_cast<cudaStream_t>(stream_v);
  const int nblocks = static_cast<int>(std::min<int64_t>(148 * 4, (rows + 63) / 64));
  const int rows_per_block = static_cast<int>((rows + nblocks - 1) / nblocks);
  const int tx = C / 8;                              // threads along channels
  const int ty = std::max(1, std::min(1024 / tx, 48 * 1024 / (C * 4)));  // row lanes (shared memory: ty * C floats)
  hy_count_launch();
  colsum_partial_kernel<<<nblocks, dim3(tx, ty), ty * C * 4, st>>>(static_cast<const __nv_bfloat16*>(g), rows, C,
                                                                  rows_per_block, static_cast<float*>(workspace));
  HY_CUDA(cudaGetLastError());
  hy_count_launch();
  colsum_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(static_cast<const float*>(workspace), nblocks, C, out);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"
