// Private (library-internal) view of a packed convolution layer, shared by the generic
// implicit-GEMM kernel (conv_tc.cu) and the fused ResidualUnit kernel (ru_fused.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "hyres_b200.h"

constexpr int kMaxRows = 5;   // vertical taps (kernel rows) per patch
constexpr int kMaxTaps = 16;  // B tiles per patch: kernel rows x weight parts of a split-precision layer

struct TapGroup {
  int32_t src;      // 0: x0, 1: x1
  int32_t c_off;    // inner (channel) coordinate of the box
  int32_t dw;       // column coordinate offset relative to the tile origin
  int32_t dh;       // row coordinate offset of the patch start
  int32_t hpar;     // coordinate along the parity dimension (stride-2 view)
  int32_t ntaps;    // B tiles consumed against this patch
  int32_t kslot0;   // first 64-wide K slot in the packed weights
  int32_t tap_row[kMaxTaps];  // row shift (in patch rows) of each tap
  int32_t tap_acc[kMaxTaps];  // accumulator of each tap (split-precision layers keep the leading products apart)
};

struct hyres_conv {
  int kind, cin0, cin1, w_cin_total, cout, R, S, stride, pad, dil;
  int BN, cout_pad, ktot, nphase, extra_rows;
  // split-precision layers (fp32-equivalent arithmetic on the bf16 tensor cores): the input tensor carries
  // nsplit bf16 parts of every fp32 activation ([.., nsplit*cin], part p in channels [p*cin, (p+1)*cin)), the
  // weights are packed as nsplit bf16 parts, and part i of the activations meets parts 0 .. nsplit-1-i of the
  // weights (3 products for nsplit = 2, 6 for nsplit = 3); every product accumulates into the same fp32 TMEM tile.
  int nsplit = 1;
  // half-part format (2 | HYRES_SPLIT_F16): two IEEE half parts, the second scaled by 2^11; nsplit == 2 then, the
  // leading product p0 x w0 and the cross products p0 x w1 + p1 x w0 never share an accumulator, and the
  // epilogue adds the cross accumulator times 2^-11.
  bool split_f16 = false;
  // Accumulators per output tile of a split layer.  The tensor core truncates when it adds into an fp32
  // accumulator (about 0.05 ulp of bias per MMA, measured: tools/check_precise.py), so a chain of 1 200 MMAs
  // drifts by ~60 ulp.  The leading products (part 0 x part 0, K/16 MMAs) therefore get accumulators of their
  // own (round robin over nacc - 1 of them) and the 2^-8 / 2^-16 sized cross products share the last one, whose
  // ulp is 2^8 smaller; the epilogue adds the accumulators in fp32 (round to nearest).
  int nacc = 1;
  int bn_cap = 256;  // widest N block (TMEM: nacc * block width <= 512 columns)
  int lead_mmas = 0; // K / 16 of the longest leading-product chain
  int acc_mask[4] = {1, 1, 1, 1};  // per phase: accumulators that receive at least one MMA
  int ph_begin[4], ph_count[4];
  std::vector<TapGroup> groups;
  std::vector<uint8_t> tap_mask;
  // k-slot -> (src, chunk, r, s) for weight packing
  struct Slot { int src, chunk, r, s, wpart; };
  std::vector<Slot> slots;
  TapGroup* d_groups = nullptr;
  Slot* d_slots = nullptr;           // device copy of `slots` (hyres_conv_update_device packs on the GPU)
  void* d_wg_items = nullptr;        // weight-gradient work items of this layer's plan (wgrad.cu), uploaded once
  bool w_tap_valid = true;           // false once the weights were re-packed on the device (d_w_tap is stale)
  __nv_bfloat16* d_w = nullptr;
  __nv_bfloat16* d_w_tap = nullptr;  // tap-major packing for the three-output-channel layers (conv_sc.cu)
  int64_t w_tap_elems = 0;
  float* d_bias = nullptr;
  std::vector<float> h_bias;  // host copy (cout_pad entries): kernels that take the bias as launch parameters
  int64_t macs_per_pos = 0;
};


// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}


// [cout_pad][ktot] K-major bf16 weights -> 2-D map, box = 64 k-elements x `rows` output channels, SWIZZLE_128B.
int encode_w_map(CUtensorMap* m, const void* ptr, int ktot, int cout_pad, int rows);
int num_sms();
// conv_res.cu: persistent kernel with shared-memory-resident weights; *handled = 0 when the layer
// does not qualify (the caller then uses the streaming kernel of conv_tc.cu).
int conv_res_try_run(hyres_conv* c, const hyres_conv_io* io, cudaStream_t stream, int* handled);
// conv_sc.cu: layers with three output channels as one tap-major GEMM per tile plus a gather epilogue.
bool conv_sc_applicable(const hyres_conv* c);
void conv_sc_pack(const hyres_conv* c, const float* w, std::vector<__nv_bfloat16>& out);
int64_t conv_sc_packed_elems(const hyres_conv* c);
int conv_sc_try_run(hyres_conv* c, const hyres_conv_io* io, cudaStream_t stream, int* handled);
