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

#include "common.cuh"
#include "conv_priv.h"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr int kTH = 16, kTW = 8;
constexpr int kThreads = 320;  // warps 0-3 / 4-7: epilogue groups 0 / 1; warp 8: TMA loads; warp 9: MMA
constexpr int kThreadsSq = 448;  // + warps 10-13: x0_square transform (GDN layers only)
constexpr int kMaxSteps = 64;
constexpr int kMaxStages = 4;
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

struct alignas(64) ResParams {
  CUtensorMap mapA, mapW, mapOut, mapAux0, mapAux1;
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
__device__ __forceinline__ float act_fn(float v, int act, float slope) {
  if (act == HYRES_ACT_RELU) return fmaxf(v, 0.f);
  if (act == HYRES_ACT_PRELU) return v >= 0.f ? v : v * slope;
  if (act == HYRES_ACT_CLAMP01) return fminf(fmaxf(v, 0.f), 1.f);
  return v;
}
__device__ __forceinline__ void unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = hy::bf16_lo(u[i]);
    f[2 * i + 1] = hy::bf16_hi(u[i]);
  }
}

__global__ void __launch_bounds__(kThreadsSq, 1) conv_res_kernel(const __grid_constant__ ResParams p) {
  hy::pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (hy::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base;
  const uint32_t st_base = base + p.w_bytes;
  const uint32_t bias_base = st_base + p.NA * p.stage_bytes;
  const uint32_t bar_base = bias_base + 1024;
  // barriers: W_FULL | A_FULL[4] | A_EMPTY[4] | A_READY[4] | ACC_FULL[2] | ACC_EMPTY[2] ; tmem slot
  const uint32_t W_FULL = bar_base;
  const uint32_t A_FULL = bar_base + 8;
  const uint32_t A_EMPTY = A_FULL + 8 * kMaxStages;
  const uint32_t A_READY = A_EMPTY + 8 * kMaxStages;
  const uint32_t ACC_FULL = A_READY + 8 * kMaxStages;
  const uint32_t ACC_EMPTY = ACC_FULL + 16;
  const uint32_t tmem_slot = ACC_EMPTY + 16;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    hy::mbar_init(W_FULL, 1);
    for (int i = 0; i < p.NA; ++i) {
      hy::mbar_init(A_FULL + 8 * i, 1);
      hy::mbar_init(A_EMPTY + 8 * i, 2);    // MMA commit + the epilogue group's release
      hy::mbar_init(A_READY + 8 * i, 128);  // x0_square transform
    }
    for (int i = 0; i < 2; ++i) {
      hy::mbar_init(ACC_FULL + 8 * i, 1);
      hy::mbar_init(ACC_EMPTY + 8 * i, 128);
    }
    hy::mbar_fence_init();
    // the weights do not depend on the predecessor kernel: their load starts before the dependency wait
    hy::tma_prefetch_desc(&p.mapW);
    hy::mbar_arrive_expect_tx(W_FULL, p.w_bytes);
    for (int s = 0; s < p.nslots; ++s) hy::tma_load_2d(w_base + s * p.BN * 128, &p.mapW, W_FULL, s * 64, 0);
  }
  {
    float* sb = reinterpret_cast<float*>(smem_raw + (bias_base - hy::smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < p.BN; i += blockDim.x) sb[i] = __ldg(p.bias + i);  // bias padded to BN
  }
  if (warp == 8 && lane == 0) {
    hy::tma_prefetch_desc(&p.mapA);
    hy::tma_prefetch_desc(&p.mapW);
    if (p.has_aux0) hy::tma_prefetch_desc(&p.mapAux0);
    if (p.has_aux1) hy::tma_prefetch_desc(&p.mapAux1);
  }
  if (warp == 9) {
    hy::tmem_alloc(tmem_slot, p.tmem_cols);
    hy::tmem_relinquish();
  }
  hy::tc_fence_before();
  __syncthreads();
  hy::tc_fence_after();
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
        hy::mbar_arrive_expect_tx(A_FULL + 8 * stage, p.stage_tx_bytes);
        for (int c = 0; c < p.nchunk_in; ++c)
          hy::tma_load_4d(sb + c * p.a_chunk_bytes, &p.mapA, A_FULL + 8 * stage, c * 64, w0 + p.org_w, h0 + p.org_h, b_img);
        if (p.has_aux0)
          for (int c = 0; c < p.nchunk_out; ++c)
            hy::tma_load_4d(sb + p.o_off + c * 16384, &p.mapAux0, A_FULL + 8 * stage, c * 64, w0, h0, b_img);
        if (p.has_aux1)
          for (int c = 0; c < p.nchunk_out; ++c)
            hy::tma_load_4d(sb + p.x1_off + c * 16384, &p.mapAux1, A_FULL + 8 * stage, c * 64, w0, h0, b_img);
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
        hy::mbar_wait(ACC_EMPTY + 8 * buf, ((it >> 1) & 1) ^ 1u);
        hy::tc_fence_after();
        for (int i = 0; i < p.nsteps; ++i) {
          const Step st = p.steps[i];
          const uint64_t a_d = a_d0 + static_cast<uint32_t>(st.a_off16), b_d = w_d + static_cast<uint32_t>(st.b_off16);
          const uint32_t d = d0 + st.d_col;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            hy::umma_issue<2>(d, a_d + 2 * k, b_d + 2 * k, idesc, static_cast<uint32_t>(st.acc | k), leader);
        }
        hy::umma_commit_mode<2>(A_EMPTY + 8 * stage, leader);
        hy::umma_commit_mode<2>(ACC_FULL + 8 * buf, leader);
        if (++stage == p.NA) { stage = 0; par ^= 1u; }
      }
    }
  } else if (warp >= 10) {
    // ============================ x0_square transform (GDN) ============================
    // GDN operand: square the activation tile in place (bf16 RN of the exact product, the same value a
    // producer-side x*x store would have held), then release it to the MMA warp.  Its own four warps, so the
    // transform of tile t+1 runs under the MMAs of tile t and the epilogues of tiles t-1, t-2.
    if (p.a_square) {
      const int row = threadIdx.x - 320;
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
    const bool need0 = p.has_aux0 != 0, need1 = p.has_aux1 != 0;
    for (int it = grp;; it += 2) {
      const int t = blockIdx.x + it * static_cast<int>(gridDim.x);
      if (t >= p.ntiles) break;
      const int stage = it % p.NA;
      const uint32_t par = (it / p.NA) & 1;
      const uint32_t sb = st_base + stage * p.stage_bytes;
      int b_img, h0, w0;
      tile_origin(t, b_img, h0, w0);
      if (need0) hy::mbar_wait(A_FULL + 8 * stage, par);  // acquire the TMA-written epilogue operands
      hy::mbar_wait(ACC_FULL + 8 * grp, (it >> 1) & 1);
      hy::tc_fence_after();

      const int hv = h0 + ti, wv = w0 + tj;
      const bool valid = hv < p.Hv && wv < p.Wv;
      const uint32_t acc0 = t_lane + grp * acc_cols;
      const int total = p.nphase * nchunk16;
      uint32_t rc[16];
      hy::tmem_ld16(acc0, rc);
      hy::tmem_ld_fence(rc);
      for (int q = 0; q < total; ++q) {
        const int ph = q / nchunk16;
        const int n = (q - ph * nchunk16) << 4;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rc[i]);
        if (q + 1 < total) {
          hy::tmem_ld16(acc0 + (q + 1) * 16, rc);  // columns of successive phases are contiguous
        } else {
          // nothing left to read from TMEM once the last chunk sits in registers
        }
        const int oh = hv * p.out_mul + (ph >> 1), ow = wv * p.out_mul + (ph & 1);
        const long long opix = (static_cast<long long>(b_img) * p.OH + oh) * p.OW + ow;
        if (p.epi == HYRES_EPI_PIXSCALE) {
          const float ps = valid ? __ldg(p.pixscale + opix) : 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= ps;
        }
        {
          const uint32_t ba = bias_base + n * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = lds_f4(ba + i * 16);
            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          }
        }
        // smem address of this thread's 16 channels inside a [128 rows][64 ch] swizzled chunk
        const uint32_t coff = (n >> 6) * 16384 + row * 128;
        const uint32_t j0 = (n & 63) >> 3;
        const uint32_t oa0 = sb + p.o_off + coff + ((j0 ^ sw) << 4);
        const uint32_t oa1 = sb + p.o_off + coff + (((j0 + 1) ^ sw) << 4);
        if (need0) {
          float x[16];
          unpack16(lds128(oa0), lds128(oa1), x);
          if (p.epi == HYRES_EPI_ADD) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += x[i];
          } else if (p.epi == HYRES_EPI_GATE) {
            float a[16];
            unpack16(lds128(sb + p.x1_off + coff + ((j0 ^ sw) << 4)), lds128(sb + p.x1_off + coff + (((j0 + 1) ^ sw) << 4)), a);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(a[i], hy::fast_sigmoid(v[i]), x[i]);
          } else if (p.epi == HYRES_EPI_GDN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = x[i] * hy::fast_rsqrt(v[i]);
          } else if (p.epi == HYRES_EPI_IGDN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = x[i] * hy::fast_sqrt(v[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = act_fn(v[i], p.act, p.slope);
        if (p.store_bf16) {
          uint4 a, b;
          a.x = hy::pack_bf16(v[0], v[1]); a.y = hy::pack_bf16(v[2], v[3]);
          a.z = hy::pack_bf16(v[4], v[5]); a.w = hy::pack_bf16(v[6], v[7]);
          b.x = hy::pack_bf16(v[8], v[9]); b.y = hy::pack_bf16(v[10], v[11]);
          b.z = hy::pack_bf16(v[12], v[13]); b.w = hy::pack_bf16(v[14], v[15]);
          sts128(oa0, a);
          sts128(oa1, b);
        }
        if (p.out_f32 && valid && n < p.cout) {
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
        if (q + 1 < total) hy::tmem_ld_fence(rc);
      }
      hy::tc_fence_before();
      hy::mbar_arrive(ACC_EMPTY + 8 * grp);
      if (p.store_bf16) hy::fence_async_smem();
      hy::named_bar_sync(1 + grp, 128);  // staging complete / every thread done with the stage's operands
      if (row == 0) {
        if (p.store_bf16) {
          for (int c = 0; c < p.nchunk_out; ++c)
            hy::tma_store_4d(&p.mapOut, sb + p.o_off + c * 16384, c * 64, w0, h0, b_img);
          hy::tma_store_commit();
          hy::tma_store_wait_read<0>();
        }
        hy::mbar_arrive(A_EMPTY + 8 * stage);
      }
    }
    if (row == 0 && p.store_bf16) hy::tma_store_wait_all<0>();
  }

  hy::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    hy::tc_fence_after();
    hy::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int encode_map4(CUtensorMap* m, const void* ptr, int C, int ld, int B, int H, int W, int box_w, int box_h) {
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
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled(act C=%d ld=%d B=%d H=%d W=%d box %dx%d) -> %d", C, ld, B, H, W,
             box_w, box_h, (int)r);
    return hy_fail(HYRES_ERR_DRIVER, msg);
  }
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
    static const int pf = [] { const char* e = getenv("HYRES_RES_PF"); return e ? atoi(e) : 0; }();
    p.pf_extra = pf;
  }
  if (p.a_square && (PH != kTH || PW != kTW)) return hy_fail(HYRES_ERR_UNSUPPORTED, "conv_run: x0_square needs a 1x1 layer");
  if (deconv && (need0 || need1)) return HYRES_OK;
  int stage = p.nchunk_in * p.a_chunk_bytes;
  p.o_off = stage;
  if (need0 || p.store_bf16) stage += p.nchunk_out * 16384;
  p.x1_off = stage;
  if (need1) stage += p.nchunk_out * 16384;
  p.stage_bytes = stage;
  p.stage_tx_bytes = p.nchunk_in * PH * PW * 128 + (need0 ? p.nchunk_out * 16384 : 0) + (need1 ? p.nchunk_out * 16384 : 0);
  const int fixed = static_cast<int>(w_bytes) + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment*/;
  int NA = (kSmemLimit - fixed) / stage;
  if (NA < 2) return HYRES_OK;
  NA = std::min(NA, kMaxStages);
  p.NA = NA;
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
  p.pixscale = io->pixscale;
  p.out_f32 = io->out_f32;
  p.f32_sb = io->f32_sb; p.f32_sh = io->f32_sh; p.f32_sw = io->f32_sw; p.f32_sc = io->f32_sc;

  const int ld_x0 = io->ld_x0 ? io->ld_x0 : c->cin0;
  int rc = encode_map4(&p.mapA, io->x0, c->cin0, ld_x0, io->B, io->H, io->W, PW, PH);
  if (rc != HYRES_OK) return rc;
  if ((rc = encode_w_map(&p.mapW, c->d_w, c->ktot, c->cout_pad, BN)) != HYRES_OK) return rc;
  if (p.store_bf16 && (rc = encode_map4(&p.mapOut, io->out_bf16, c->cout, io->ld_out, io->B, OH, OW, kTW, kTH)) != HYRES_OK) return rc;
  if (need0 && (rc = encode_map4(&p.mapAux0, io->aux0, c->cout, io->ld_aux0, io->B, OH, OW, kTW, kTH)) != HYRES_OK) return rc;
  if (need1 && (rc = encode_map4(&p.mapAux1, io->aux1, c->cout, io->ld_aux1, io->B, OH, OW, kTW, kTH)) != HYRES_OK) return rc;

  const int smem = fixed + NA * stage;
  static int smem_set = 0;
  if (!smem_set) {
    HY_CUDA(cudaFuncSetAttribute(conv_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    smem_set = 1;
  }
  const int grid = std::min(p.ntiles, num_sms());
  hy_count_launch();
  HY_CUDA(hy_launch_pdl(conv_res_kernel, grid, p.a_square ? kThreadsSq : kThreads, smem, stream, p));
  *handled = 1;
  return HYRES_OK;
}
