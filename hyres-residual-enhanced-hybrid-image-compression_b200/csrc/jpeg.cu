// The JPEG stage of ResidualJPEGCompression on the device (models/utils/turbo_jpeg_compression.py:17-77).
//
// The reference round-trips every image through libjpeg-turbo on the CPU (TurboJPEG.encode, PyTurboJPEG defaults:
// the RGB array is read as BGR, 4:2:2, baseline Huffman with the Annex K tables; TurboJPEG.decode: ISLOW IDCT, fancy
// up-sampling) and uses two things from it: the decoded pixels and the file size.  Both are integer functions of
// the input bytes, so they are reproduced here bit for bit (oracle/jpeg_oracle.py is the restatement, pinned to
// libjpeg-turbo 3.1.2 byte-for-byte):
//
//   jpeg_color_fwd    x fp32 NCHW -> .byte() -> YCbCr (jccolor.c fixed point) -> h2v1 chroma down-sample
//   jpeg_dct          8 threads per 8x8 block: ISLOW FDCT (jfdctint.c), quantise (jcdctmgr.c), store the
//                     coefficients in zig-zag order, de-quantise, ISLOW IDCT (jidctint.c), range limit
//   jpeg_color_inv    fancy h2v1 up-sample (jdsample.c) + YCbCr -> RGB (jdcolor.c) -> u8 / 255 fp32 NCHW
//   jpeg_block_bits   Huffman code length of every block in scan order (jchuff.c encode_one_block)
//   jpeg_scan_offsets exclusive prefix sum per image
//   jpeg_block_write  the same walk, writing the bits at the block's offset (atomicOr: order independent)
//   jpeg_file_size    0xFF bytes of the padded scan (they are stuffed with 0x00) -> file size
//
// All of it is byte / integer work bound by HBM (about 40 B per pixel in total); none of it touches the host.
#include <cstdint>
#include <cstring>
#include <initializer_list>

#include "common.cuh"
#include "host_util.h"
#include "hyres_b200.h"

namespace {

// ---- Annex K tables (jcparam.c / jutils.c), natural order ----
const uint8_t kStdLumQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                              14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                              18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                              49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kStdChrQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                              99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                              99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t kDcLumBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChrBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChrBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

constexpr int kHeaderBytes = 623;  // SOI + APP0 + 2 DQT + SOF0 + 4 DHT + SOS (jcmarker.c), independent of the image
constexpr int kMaxBlockBits = 1728;  // 63 x (16 + 10) + 9 + 11 rounded up to a multiple of 64

// Huffman tables as (code << 8 | length), index = symbol; [0]: luminance, [1]: chrominance
struct HuffTables {
  uint32_t dc[2][12];
  uint32_t ac[2][256];
};
struct QuantTables {
  uint16_t q[2][64];  // natural order
};

__constant__ uint8_t c_zigzag[64];
__device__ HuffTables g_huff;

void derive(const uint8_t* bits, const uint8_t* vals, uint32_t* table) {
  // jchuff.c jpeg_make_c_derived_tbl: canonical codes in order of increasing length
  uint32_t code = 0;
  int k = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int i = 0; i < bits[len - 1]; ++i) table[vals[k++]] = (code++ << 8) | static_cast<uint32_t>(len);
    code <<= 1;
  }
}

void quant_tables_host(int quality, QuantTables* t) {
  // jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE)
  int q = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
  const int scale = q < 50 ? 5000 / q : 200 - 2 * q;
  for (int c = 0; c < 2; ++c)
    for (int i = 0; i < 64; ++i) {
      long v = ((c ? kStdChrQ[i] : kStdLumQ[i]) * static_cast<long>(scale) + 50) / 100;
      t->q[c][i] = static_cast<uint16_t>(v < 1 ? 1 : (v > 255 ? 255 : v));
    }
}

int upload_tables() {
  static bool done[64] = {};
  int dev = 0;
  HY_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && done[dev]) return HYRES_OK;
  HuffTables h;
  memset(&h, 0, sizeof h);
  derive(kDcLumBits, kDcVals, h.dc[0]);
  derive(kDcChrBits, kDcVals, h.dc[1]);
  derive(kAcLumBits, kAcLumVals, h.ac[0]);
  derive(kAcChrBits, kAcChrVals, h.ac[1]);
  HY_CUDA(cudaMemcpyToSymbol(g_huff, &h, sizeof h));
  HY_CUDA(cudaMemcpyToSymbol(c_zigzag, kZigzag, 64));
  if (dev < 64) done[dev] = true;
  return HYRES_OK;
}

// ------------------------------------------------------------------------------------------------
// colour conversion + chroma down-sampling.  One thread = two horizontally adjacent pixels.
// The array's channel 0 is what libjpeg reads as B (PyTurboJPEG's TJPF_BGR default on an RGB array).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int to_byte(float v) {
  v = fminf(fmaxf(v, 0.f), 1.f);
  return static_cast<int>(v * 255.f);  // torch: (clamp(x) * 255).byte() truncates
}

__global__ void jpeg_color_fwd_kernel(const float* __restrict__ x, uint8_t* __restrict__ yp, uint8_t* __restrict__ cbp,
                                      uint8_t* __restrict__ crp, int B, int H, int W) {
  const int W2 = W >> 1;
  const int64_t total = static_cast<int64_t>(B) * H * W2;
  const int64_t plane = static_cast<int64_t>(H) * W;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(t % W2);
    const int64_t bh = t / W2;
    const int h = static_cast<int>(bh % H);
    const int64_t b = bh / H;
    const float* p = x + (b * 3) * plane + static_cast<int64_t>(h) * W + 2 * j;
    const float2 c0 = __ldg(reinterpret_cast<const float2*>(p));
    const float2 c1 = __ldg(reinterpret_cast<const float2*>(p + plane));
    const float2 c2 = __ldg(reinterpret_cast<const float2*>(p + 2 * plane));
    int yv[2], cb[2], cr[2];
    const int bb[2] = {to_byte(c0.x), to_byte(c0.y)}, gg[2] = {to_byte(c1.x), to_byte(c1.y)},
              rr[2] = {to_byte(c2.x), to_byte(c2.y)};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      yv[i] = (19595 * rr[i] + 38470 * gg[i] + 7471 * bb[i] + 32768) >> 16;
      cb[i] = (-11059 * rr[i] - 21709 * gg[i] + 32768 * bb[i] + (128 << 16) + 32767) >> 16;
      cr[i] = (32768 * rr[i] - 27439 * gg[i] - 5329 * bb[i] + (128 << 16) + 32767) >> 16;
    }
    const int bias = j & 1;  // jcsample.c h2v1_downsample: 0, 1, 0, 1, ...
    *reinterpret_cast<uchar2*>(yp + bh * W + 2 * j) = make_uchar2(static_cast<uint8_t>(yv[0]), static_cast<uint8_t>(yv[1]));
    cbp[bh * W2 + j] = static_cast<uint8_t>((cb[0] + cb[1] + bias) >> 1);
    crp[bh * W2 + j] = static_cast<uint8_t>((cr[0] + cr[1] + bias) >> 1);
  }
}

// ------------------------------------------------------------------------------------------------
// 8x8 block transform: FDCT -> quantise -> (store) -> de-quantise -> IDCT.  libjpeg's ISLOW integer transforms.
// ------------------------------------------------------------------------------------------------
constexpr int CB = 13, P1 = 2;
constexpr int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633;
constexpr int F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

template <bool FIRST>
__device__ __forceinline__ void fdct8(const int (&d)[8], int (&o)[8]) {
  const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
  const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  constexpr int sh = FIRST ? CB - P1 : CB + P1;
  if (FIRST) {
    o[0] = (t10 + t11) << P1;
    o[4] = (t10 - t11) << P1;
  } else {
    o[0] = descale(t10 + t11, P1);
    o[4] = descale(t10 - t11, P1);
  }
  int z1 = (t12 + t13) * F_0_541;
  o[2] = descale(z1 + t13 * F_0_765, sh);
  o[6] = descale(z1 - t12 * F_1_847, sh);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * F_1_175;
  const int a4 = t4 * F_0_298, a5 = t5 * F_2_053, a6 = t6 * F_3_072, a7 = t7 * F_1_501;
  z1 = -z1 * F_0_899;
  z2 = -z2 * F_2_562;
  z3 = -z3 * F_1_961 + z5;
  z4 = -z4 * F_0_390 + z5;
  o[7] = descale(a4 + z1 + z3, sh);
  o[5] = descale(a5 + z2 + z4, sh);
  o[3] = descale(a6 + z2 + z3, sh);
  o[1] = descale(a7 + z1 + z4, sh);
}

template <bool FIRST>
__device__ __forceinline__ void idct8(const int (&c)[8], int (&o)[8]) {
  int z2 = c[2], z3 = c[6];
  int z1 = (z2 + z3) * F_0_541;
  const int e2 = z1 - z3 * F_1_847, e3 = z1 + z2 * F_0_765;
  const int e0 = (c[0] + c[4]) << CB, e1 = (c[0] - c[4]) << CB;
  const int t10 = e0 + e3, t13 = e0 - e3, t11 = e1 + e2, t12 = e1 - e2;
  int t0 = c[7], t1 = c[5], t2 = c[3], t3 = c[1];
  z1 = t0 + t3;
  z2 = t1 + t2;
  z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * F_1_175;
  t0 *= F_0_298;
  t1 *= F_2_053;
  t2 *= F_3_072;
  t3 *= F_1_501;
  z1 = -z1 * F_0_899;
  z2 = -z2 * F_2_562;
  z3 = -z3 * F_1_961 + z5;
  z4 = -z4 * F_0_390 + z5;
  t0 += z1 + z3;
  t1 += z2 + z4;
  t2 += z2 + z3;
  t3 += z1 + z4;
  constexpr int sh = FIRST ? CB - P1 : CB + P1 + 3;
  o[0] = descale(t10 + t3, sh);
  o[7] = descale(t10 - t3, sh);
  o[1] = descale(t11 + t2, sh);
  o[6] = descale(t11 - t2, sh);
  o[2] = descale(t12 + t1, sh);
  o[5] = descale(t12 - t1, sh);
  o[3] = descale(t13 + t0, sh);
  o[4] = descale(t13 - t0, sh);
}

// Block numbering of one image: [0, nY) luminance blocks in raster order, then nY/2 Cb, then nY/2 Cr.
// planes / recon: Y [B][H][W] | Cb [B][H][W/2] | Cr [B][H][W/2] (u8); coef: [B][2 nY][64] int16, zig-zag order.
__global__ void __launch_bounds__(256)
jpeg_dct_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ recon, int16_t* __restrict__ coef,
                const QuantTables qt, int B, int H, int W) {
  __shared__ int tile[32][2][72];
  const int lane8 = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;
  const int nY = (H >> 3) * (W >> 3);
  const int64_t per_img = 2 * static_cast<int64_t>(nY);
  const int64_t total = per_img * B;
  const int64_t ysz = static_cast<int64_t>(B) * H * W;
  const int64_t nwarp_blocks = (total + 3) >> 2;  // four blocks per warp
  for (int64_t wb = blockIdx.x * 8 + (threadIdx.x >> 5); wb < nwarp_blocks; wb += static_cast<int64_t>(gridDim.x) * 8) {
    const int64_t blk = wb * 4 + ((threadIdx.x >> 3) & 3);
    const bool active = blk < total;
    int comp = 0, pw = W;
    int64_t base = 0;
    if (active) {
      const int64_t b = blk / per_img;
      int r = static_cast<int>(blk - b * per_img);
      if (r >= nY) {
        comp = 1;
        r -= nY;
        pw = W >> 1;
        base = ysz;
        if (r >= (nY >> 1)) {
          r -= nY >> 1;
          base += ysz >> 1;
        }
      }
      const int bw = pw >> 3;
      const int by = r / bw, bx = r - by * bw;
      base += (b * H + by * 8) * pw + bx * 8;
    }
    int (*A)[9] = reinterpret_cast<int(*)[9]>(tile[slot][0]);
    int (*Q)[9] = reinterpret_cast<int(*)[9]>(tile[slot][1]);
    int d[8], o[8];
    // ---- forward, pass 1: this thread's row ----
    if (active) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(planes + base + static_cast<int64_t>(lane8) * pw));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        d[i] = static_cast<int>((raw.x >> (8 * i)) & 0xFF) - 128;
        d[4 + i] = static_cast<int>((raw.y >> (8 * i)) & 0xFF) - 128;
      }
      fdct8<true>(d, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) A[lane8][i] = o[i];
    }
    __syncwarp();
    // ---- forward, pass 2: this thread's column; quantise ----
    int q[8];
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = A[i][lane8];
      fdct8<false>(d, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int qv = qt.q[comp][k * 8 + lane8];
        const int div = qv << 3;  // the FDCT output carries a factor 8
        const int a = o[k] < 0 ? -o[k] : o[k];
        const int v = (a + (div >> 1)) / div;
        q[k] = o[k] < 0 ? -v : v;
        Q[k][lane8] = q[k];
        d[k] = q[k] * qv;  // de-quantised, for the inverse transform
      }
    }
    __syncwarp();
    if (active) {
      // coefficients out, zig-zag order: this thread writes positions 8*lane8 .. 8*lane8 + 7 (16 bytes)
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n0 = c_zigzag[lane8 * 8 + 2 * i], n1 = c_zigzag[lane8 * 8 + 2 * i + 1];
        const uint32_t lo = static_cast<uint16_t>(static_cast<int16_t>(Q[n0 >> 3][n0 & 7]));
        const uint32_t hi = static_cast<uint16_t>(static_cast<int16_t>(Q[n1 >> 3][n1 & 7]));
        w[i] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4*>(coef + blk * 64 + lane8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
      // ---- inverse, pass 1: this thread's column ----
      idct8<true>(d, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) A[i][lane8] = o[i];
    }
    __syncwarp();
    if (active) {
      // ---- inverse, pass 2: this thread's row; range limit ----
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = A[lane8][i];
      idct8<false>(d, o);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        lo |= static_cast<uint32_t>(min(max(o[i] + 128, 0), 255)) << (8 * i);
        hi |= static_cast<uint32_t>(min(max(o[4 + i] + 128, 0), 255)) << (8 * i);
      }
      *reinterpret_cast<uint2*>(recon + base + static_cast<int64_t>(lane8) * pw) = make_uint2(lo, hi);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// fancy h2v1 up-sampling + YCbCr -> RGB -> fp32 / 255.  One thread = the two pixels of one chroma sample.
// ------------------------------------------------------------------------------------------------
__global__ void jpeg_color_inv_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ cbp,
                                      const uint8_t* __restrict__ crp, float* __restrict__ out, int B, int H, int W) {
  const int W2 = W >> 1;
  const int64_t total = static_cast<int64_t>(B) * H * W2;
  const int64_t plane = static_cast<int64_t>(H) * W;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(t % W2);
    const int64_t bh = t / W2;
    const int h = static_cast<int>(bh % H);
    const int64_t b = bh / H;
    const uint8_t* cbr = cbp + bh * W2;
    const uint8_t* crr = crp + bh * W2;
    const int jl = j > 0 ? j - 1 : 0, jr = j < W2 - 1 ? j + 1 : W2 - 1;
    const int cb1 = cbr[j], cr1 = crr[j];
    int cb[2], cr[2];
    // jdsample.c h2v1_fancy_upsample; the first and last output columns copy their sample
    cb[0] = j > 0 ? (3 * cb1 + cbr[jl] + 1) >> 2 : cb1;
    cr[0] = j > 0 ? (3 * cr1 + crr[jl] + 1) >> 2 : cr1;
    cb[1] = j < W2 - 1 ? (3 * cb1 + cbr[jr] + 2) >> 2 : cb1;
    cr[1] = j < W2 - 1 ? (3 * cr1 + crr[jr] + 2) >> 2 : cr1;
    const uchar2 yy = *reinterpret_cast<const uchar2*>(yp + bh * W + 2 * j);
    const int yv[2] = {yy.x, yy.y};
    float c0[2], c1[2], c2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int u = cb[i] - 128, v = cr[i] - 128;
      const int r = yv[i] + ((91881 * v + 32768) >> 16);
      const int bl = yv[i] + ((116130 * u + 32768) >> 16);
      const int g = yv[i] + ((-22554 * u - 46802 * v + 32768) >> 16);
      c0[i] = static_cast<float>(min(max(bl, 0), 255)) / 255.0f;
      c1[i] = static_cast<float>(min(max(g, 0), 255)) / 255.0f;
      c2[i] = static_cast<float>(min(max(r, 0), 255)) / 255.0f;
    }
    float* p = out + (b * 3) * plane + static_cast<int64_t>(h) * W + 2 * j;
    *reinterpret_cast<float2*>(p) = make_float2(c0[0], c0[1]);
    *reinterpret_cast<float2*>(p + plane) = make_float2(c1[0], c1[1]);
    *reinterpret_cast<float2*>(p + 2 * plane) = make_float2(c2[0], c2[1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Huffman coding of one block (jchuff.c encode_one_block).  Scan order of an image (4:2:2 interleaved):
// MCU m = luminance blocks 2m, 2m+1, Cb block m, Cr block m -> scan index 4m + {0, 1, 2, 3}.
// ------------------------------------------------------------------------------------------------
struct BitSink {
  uint32_t* words;
  uint64_t acc;
  int nacc;
  int64_t word;
  __device__ __forceinline__ void put(uint32_t code, int len) {
    acc = (acc << len) | code;
    nacc += len;
    if (nacc >= 32) {
      atomicOr(words + word, static_cast<uint32_t>(acc >> (nacc - 32)));
      ++word;
      nacc -= 32;
    }
  }
  __device__ __forceinline__ void flush() {
    if (nacc > 0) atomicOr(words + word, static_cast<uint32_t>(acc << (32 - nacc)));
  }
};

template <bool WRITE>
__device__ __forceinline__ int encode_block(const int16_t* __restrict__ zz, int pred, int tbl, BitSink& sink) {
  int bits = 0;
  const uint4* src = reinterpret_cast<const uint4*>(zz);
  int run = 0;
#pragma unroll 1
  for (int g = 0; g < 8; ++g) {
    const uint4 raw = __ldg(src + g);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int v = static_cast<int16_t>((w[i >> 1] >> (16 * (i & 1))) & 0xFFFF);
      if (g == 0 && i == 0) {
        const int diff = v - pred;
        const int a = diff < 0 ? -diff : diff;
        const int nb = 32 - __clz(a);
        const uint32_t e = g_huff.dc[tbl][nb];
        const int len = static_cast<int>(e & 0xFF) + nb;
        bits += len;
        if (WRITE) {
          const uint32_t extra = static_cast<uint32_t>(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u);
          sink.put(((e >> 8) << nb) | extra, len);
        }
        continue;
      }
      if (v == 0) {
        ++run;
        continue;
      }
      while (run > 15) {
        const uint32_t e = g_huff.ac[tbl][0xF0];
        bits += static_cast<int>(e & 0xFF);
        if (WRITE) sink.put(e >> 8, static_cast<int>(e & 0xFF));
        run -= 16;
      }
      const int a = v < 0 ? -v : v;
      const int nb = 32 - __clz(a);
      const uint32_t e = g_huff.ac[tbl][(run << 4) + nb];
      const int len = static_cast<int>(e & 0xFF) + nb;
      bits += len;
      if (WRITE) {
        const uint32_t extra = static_cast<uint32_t>(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
        sink.put(((e >> 8) << nb) | extra, len);
      }
      run = 0;
    }
  }
  if (run > 0) {
    const uint32_t e = g_huff.ac[tbl][0];
    bits += static_cast<int>(e & 0xFF);
    if (WRITE) sink.put(e >> 8, static_cast<int>(e & 0xFF));
  }
  return bits;
}

// scan index -> (coefficient block of this image, table, coefficient block holding the DC predictor or -1)
__device__ __forceinline__ void scan_block(int s, int nY, int& blk, int& tbl, int& pred_blk) {
  const int m = s >> 2, j = s & 3;
  if (j < 2) {
    blk = 2 * m + j;
    tbl = 0;
    pred_blk = blk - 1;
  } else {
    blk = nY + (j == 3 ? (nY >> 1) : 0) + m;
    tbl = 1;
    pred_blk = m > 0 ? blk - 1 : -1;
  }
}

template <bool WRITE>
__global__ void jpeg_block_code_kernel(const int16_t* __restrict__ coef, int32_t* __restrict__ lens,
                                       const int64_t* __restrict__ offs, uint32_t* __restrict__ words,
                                       int64_t words_per_img, int B, int nY) {
  const int per_img = 2 * nY;
  const int64_t total = static_cast<int64_t>(per_img) * B;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = t / per_img;
    const int s = static_cast<int>(t - b * per_img);
    int blk, tbl, pb;
    scan_block(s, nY, blk, tbl, pb);
    const int16_t* cimg = coef + b * per_img * 64;
    const int pred = pb >= 0 ? cimg[static_cast<int64_t>(pb) * 64] : 0;
    BitSink sink;
    if (WRITE) {
      const int64_t start = offs[t];
      sink.words = words + b * words_per_img;
      sink.word = start >> 5;
      sink.nacc = static_cast<int>(start & 31);
      sink.acc = 0;
    }
    const int bits = encode_block<WRITE>(cimg + static_cast<int64_t>(blk) * 64, pred, tbl, sink);
    if (WRITE)
      sink.flush();
    else
      lens[t] = bits;
  }
}

// exclusive prefix sum of the block lengths of one image (one CTA per image), total into nbits[b]
__global__ void __launch_bounds__(1024) jpeg_scan_offsets_kernel(const int32_t* __restrict__ lens,
                                                                  int64_t* __restrict__ offs,
                                                                  int64_t* __restrict__ nbits, int per_img) {
  __shared__ int64_t part[1024];
  const int b = blockIdx.x;
  const int32_t* l = lens + static_cast<int64_t>(b) * per_img;
  int64_t* o = offs + static_cast<int64_t>(b) * per_img;
  const int chunk = (per_img + 1023) / 1024;
  const int lo = threadIdx.x * chunk, hi = min(lo + chunk, per_img);
  int64_t s = 0;
  for (int i = lo; i < hi; ++i) s += l[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {  // Hillis-Steele inclusive scan
    const int64_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int64_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (int i = lo; i < hi; ++i) {
    o[i] = run;
    run += l[i];
  }
  if (threadIdx.x == 1023) nbits[b] = part[1023];
}

// file size = header + scan bytes (final byte padded with one bits) + one stuffed 0x00 per 0xFF byte + EOI
__global__ void __launch_bounds__(1024) jpeg_file_size_kernel(const uint32_t* __restrict__ words, int64_t words_per_img,
                                                               const int64_t* __restrict__ nbits,
                                                               int64_t* __restrict__ sizes) {
  __shared__ int part[32];
  const int b = blockIdx.x;
  const uint32_t* w = words + b * words_per_img;
  const int64_t nb = nbits[b];
  const int64_t nbytes = (nb + 7) >> 3;
  const int64_t nwords = (nbytes + 3) >> 2;
  int ff = 0;
  for (int64_t i = threadIdx.x; i < nwords; i += blockDim.x) {
    uint32_t v = w[i];
    if (i == ((nb - 1) >> 5) && (nb & 7)) v |= ((1u << (8 - (nb & 7))) - 1u) << (24 - 8 * (((nb - 1) >> 3) & 3));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i * 4 + k < nbytes && ((v >> (24 - 8 * k)) & 0xFF) == 0xFF) ++ff;
  }
  for (int d = 16; d > 0; d >>= 1) ff += __shfl_xor_sync(0xffffffffu, ff, d);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ff;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < 32; ++i) tot += part[i];
    sizes[b] = kHeaderBytes + nbytes + tot + 2;
  }
}

// bits per pixel of the batch = 8 * sum(file sizes) / pixels, as the fp32 the reference's Python float becomes
__global__ void jpeg_bpp_kernel(const int64_t* __restrict__ sizes, int B, double pixels, float* __restrict__ bpp) {
  if (threadIdx.x != 0) return;
  int64_t tot = 0;
  for (int i = 0; i < B; ++i) tot += sizes[i];
  bpp[0] = static_cast<float>(static_cast<double>(tot) * 8.0 / pixels);
}

inline int64_t align256(int64_t v) { return (v + 255) & ~static_cast<int64_t>(255); }

struct Workspace {
  int64_t planes, recon, coef, lens, offs, end;
};
Workspace layout(int B, int H, int W) {
  Workspace w;
  const int64_t px = static_cast<int64_t>(B) * H * W;
  const int64_t nblk = px / 32;  // 2 * (H/8) * (W/8) blocks per image
  w.planes = 0;
  w.recon = align256(w.planes + 2 * px);
  w.coef = align256(w.recon + 2 * px);
  w.lens = align256(w.coef + nblk * 64 * 2);
  w.offs = align256(w.lens + nblk * 4);
  w.end = align256(w.offs + nblk * 8);
  return w;
}

inline int grid_for(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<int>(g);
}

}  // namespace

extern "C" {

int64_t hyres_jpeg_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 8 || W < 16 || (H % 8) || (W % 16)) return 0;
  return layout(B, H, W).end;
}

int64_t hyres_jpeg_scan_words(int H, int W) {
  if (H < 8 || W < 16 || (H % 8) || (W % 16)) return 0;
  const int64_t nblk = static_cast<int64_t>(H / 8) * (W / 8) * 2;
  return nblk * (kMaxBlockBits / 32) + 2;
}

int hyres_jpeg_header_bytes(void) { return kHeaderBytes; }

int hyres_jpeg_bpp(const int64_t* sizes, int B, int64_t pixels, float* bpp, void* stream_) {
  if (!sizes || !bpp || B < 1 || pixels < 1) return hy_fail(HYRES_ERR_ARG, "jpeg_bpp: bad argument");
  hy_count_launch();
  jpeg_bpp_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(sizes, B, static_cast<double>(pixels), bpp);
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_jpeg_forward(const float* x, int B, int H, int W, int quality, void* workspace, float* decoded,
                       int64_t* sizes, uint32_t* scan_words, int64_t* scan_bits, void* stream_) {
  if (!x || !workspace || B < 1) return hy_fail(HYRES_ERR_ARG, "jpeg_forward: null pointer or empty batch");
  if (H < 8 || W < 16 || (H % 8) || (W % 16))
    return hy_fail(HYRES_ERR_UNSUPPORTED, "jpeg_forward: H must be a multiple of 8 and W of 16 (no edge padding)");
  if ((sizes || scan_bits) && !scan_words) return hy_fail(HYRES_ERR_ARG, "jpeg_forward: sizes / scan_bits need scan_words");
  if (scan_words && !scan_bits) return hy_fail(HYRES_ERR_ARG, "jpeg_forward: scan_words needs scan_bits");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = upload_tables();
  if (rc != HYRES_OK) return rc;
  QuantTables qt;
  quant_tables_host(quality, &qt);
  const Workspace ws = layout(B, H, W);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  const int64_t px = static_cast<int64_t>(B) * H * W;
  uint8_t* planes = base + ws.planes;
  uint8_t* recon = base + ws.recon;
  int16_t* coef = reinterpret_cast<int16_t*>(base + ws.coef);
  int32_t* lens = reinterpret_cast<int32_t*>(base + ws.lens);
  int64_t* offs = reinterpret_cast<int64_t*>(base + ws.offs);
  const int nY = (H / 8) * (W / 8);
  const int64_t nblk = static_cast<int64_t>(2 * nY) * B;

  hy_count_launch();
  jpeg_color_fwd_kernel<<<grid_for(px / 2, 256), 256, 0, stream>>>(x, planes, planes + px, planes + px + px / 2, B, H, W);
  HY_CUDA(cudaGetLastError());
  hy_count_launch();
  jpeg_dct_kernel<<<grid_for(nblk, 32), 256, 0, stream>>>(planes, recon, coef, qt, B, H, W);
  HY_CUDA(cudaGetLastError());
  if (decoded) {
    hy_count_launch();
    jpeg_color_inv_kernel<<<grid_for(px / 2, 256), 256, 0, stream>>>(recon, recon + px, recon + px + px / 2, decoded, B, H, W);
    HY_CUDA(cudaGetLastError());
  }
  if (scan_words) {
    const int64_t wpi = hyres_jpeg_scan_words(H, W);
    HY_CUDA(cudaMemsetAsync(scan_words, 0, static_cast<size_t>(wpi) * B * 4, stream));
    hy_count_launch();
    jpeg_block_code_kernel<false><<<grid_for(nblk, 128), 128, 0, stream>>>(coef, lens, nullptr, nullptr, wpi, B, nY);
    HY_CUDA(cudaGetLastError());
    hy_count_launch();
    jpeg_scan_offsets_kernel<<<B, 1024, 0, stream>>>(lens, offs, scan_bits, 2 * nY);
    HY_CUDA(cudaGetLastError());
    hy_count_launch();
    jpeg_block_code_kernel<true><<<grid_for(nblk, 128), 128, 0, stream>>>(coef, nullptr, offs, scan_words, wpi, B, nY);
    HY_CUDA(cudaGetLastError());
    if (sizes) {
      hy_count_launch();
      jpeg_file_size_kernel<<<B, 1024, 0, stream>>>(scan_words, wpi, scan_bits, sizes);
      HY_CUDA(cudaGetLastError());
    }
  }
  return HYRES_OK;
}

// Host: the JPEG file of one image from its scan bits (big-endian words as written by hyres_jpeg_forward).
int hyres_jpeg_assemble(const uint32_t* words, int64_t nbits, int H, int W, int quality, uint8_t* out, int64_t cap,
                        int64_t* len) {
  if (!words || !out || !len || nbits < 0) return hy_fail(HYRES_ERR_ARG, "jpeg_assemble: bad argument");
  if (H < 1 || W < 1 || H > 65535 || W > 65535) return hy_fail(HYRES_ERR_ARG, "jpeg_assemble: image size out of range");
  QuantTables qt;
  quant_tables_host(quality, &qt);
  int64_t n = 0;
  auto put = [&](int v) {
    if (n < cap) out[n] = static_cast<uint8_t>(v);
    ++n;
  };
  auto put16 = [&](int v) { put(v >> 8); put(v & 0xFF); };
  put(0xFF); put(0xD8);
  put(0xFF); put(0xE0); put16(16);
  for (int c : std::initializer_list<int>{'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0}) put(c);
  for (int t = 0; t < 2; ++t) {
    put(0xFF); put(0xDB); put16(67); put(t);
    for (int i = 0; i < 64; ++i) put(qt.q[t][kZigzag[i]]);
  }
  put(0xFF); put(0xC0); put16(17); put(8); put16(H); put16(W); put(3);
  for (int c : std::initializer_list<int>{1, 0x21, 0, 2, 0x11, 1, 3, 0x11, 1}) put(c);
  struct { int id; const uint8_t* bits; const uint8_t* vals; int nvals; } dht[4] = {
      {0x00, kDcLumBits, kDcVals, 12}, {0x10, kAcLumBits, kAcLumVals, 162},
      {0x01, kDcChrBits, kDcVals, 12}, {0x11, kAcChrBits, kAcChrVals, 162}};
  for (auto& d : dht) {
    put(0xFF); put(0xC4); put16(3 + 16 + d.nvals); put(d.id);
    for (int i = 0; i < 16; ++i) put(d.bits[i]);
    for (int i = 0; i < d.nvals; ++i) put(d.vals[i]);
  }
  put(0xFF); put(0xDA); put16(12); put(3);
  for (int c : std::initializer_list<int>{1, 0x00, 2, 0x11, 3, 0x11, 0, 0x3F, 0}) put(c);
  const int64_t nbytes = (nbits + 7) >> 3;
  for (int64_t i = 0; i < nbytes; ++i) {
    int v = (words[i >> 2] >> (24 - 8 * (i & 3))) & 0xFF;
    if (i == nbytes - 1 && (nbits & 7)) v |= (1 << (8 - (nbits & 7))) - 1;
    put(v);
    if (v == 0xFF) put(0);
  }
  put(0xFF); put(0xD9);
  *len = n;
  if (n > cap) return hy_fail(HYRES_ERR_ARG, "jpeg_assemble: output buffer too small");
  return HYRES_OK;
}

}  // extern "C"
