// Device-resident entropy coder: the same rANS64 byte strings as csrc/rans.cpp (compressai 1.2.6 `ans` module,
// reached in the reference from models/checkerboard.py:159-165,172-173,206), produced and consumed on the GPU.
//
// Why: a rANS string is one dependency chain (state -> symbol -> state), so a string cannot be split across
// threads without changing its bytes.  The host coder runs that chain at ~5 ns per symbol and core, which is fine while
// one GPU has the box's cores to itself and becomes the limit of compress + decompress when eight GPUs share them
// (34 M state updates per 2048x1408 image).  Here every string is coded by ONE WARP: the chain runs redundantly in all
// 32 lanes (no divergence), while the lanes share everything that is not on the chain --
//   * 32 symbols' slots / codes are resolved at once, one per lane, and their table entries gathered a chunk ahead of the
//     chain; the chain reads them back as shared-memory broadcasts issued ahead of their use;
//   * the decoder first tries the row's MODE (value 0), whose start and frequency were fetched before the state was
//     known: one compare and one multiply-add; another symbol is found with a ballot over the 32 bins around the
//     centre (one bin per lane) and, rarely, a warp-parallel search of the row;
//   * renormalisation words are gathered in lane registers and stored / loaded 32 at a time.
// A warp runs the chain at ~60 ns per symbol -- ten times slower than a host core -- but the warps of a launch sit in a
// few blocks that hold whole SMs (see kHogBytes), dozens of strings (all images in flight) are coded beside the
// persistent convolution kernels of other streams, and the host cores are not involved at all: compress + decompress
// scale with the number of GPUs of a box (profiles/r02_scaling.md, profiles/r02_ncu_rans_dev.md).
//
// Table formats are the host coder's (hyres_rans_table_export): 16-byte encoder entries (64-bit reciprocal, bias,
// freq - 1, shift, valid), decoder words start | (freq - 1) << 16, rows = [4][n_rows] (first entry, offset, escape
// bin, usable).  Encoder SLOTS and decoder CODES are those of hyres_gc_symbols / hyres_gc_codes.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kRansL = 1ull << 31;
constexpr int32_t kKnownBit = 1 << 30;
enum { kStOk = 0, kStBad = 1, kStOverflow = 2 };

// ---------------------------------------------------------------------------------------------------------------
// Encoder
// ---------------------------------------------------------------------------------------------------------------
// 4 raw bits (escape path): renormalise against 2^59, then shift them in.  Rare: lane 0 stores directly.
__device__ __forceinline__ void enc_put4(unsigned long long& x, uint32_t* out, long long& p, uint32_t val, int lane) {
  if (x >= (1ull << 59)) {
    --p;
    if (lane == 0) out[p] = static_cast<uint32_t>(x);
    x >>= 32;
  }
  x = (x << 4) | val;
}

// One in-table symbol: renormalise (the word, if any, goes to lane `cnt`'s register), then
// x' = ((x / freq) << 16) + x % freq + start as x + bias + (x / freq) * (2^16 - freq), the division by reciprocal.
__device__ __forceinline__ void enc_symbol(unsigned long long& x, const uint4 ek, uint32_t& pend, int& cnt, int lane) {
  const uint32_t freq = (ek.w & 0xffffu) + 1u, shift = (ek.w >> 16) & 0xffu;
  const bool r = static_cast<uint32_t>(x >> 47) >= freq;  // x >= ((L >> 16) << 32) * freq
  pend = (r && lane == cnt) ? static_cast<uint32_t>(x) : pend;
  cnt += r ? 1 : 0;
  x = r ? (x >> 32) : x;
  const unsigned long long rcp = (static_cast<unsigned long long>(ek.y) << 32) | ek.x;
  const unsigned long long q = __umul64hi(x, rcp) >> shift;
  x = x + ek.z + q * static_cast<unsigned long long>(65536u - freq);
}

// One launch codes the strings of up to four groups (e.g. the z strings and both y passes of a batch): a group is
// `count` strings of n symbols with one table set.  One warp per string, all warps of a launch in as few blocks as
// possible, so that the coder occupies few SMs (the persistent convolution kernels of other streams want whole SMs).
struct EncGroup {
  const int32_t* symbols;
  const int32_t* index;
  const uint4* enc;
  const int32_t* rows;
  uint32_t* scratch;
  long long n, cap_words;
  uint32_t n_entries;
  int n_rows, count, slots;
};
struct EncParams {
  EncGroup g[4];
  int n_groups, total, meta_base;
  uint32_t* dst;
  long long dst_cap;
  int32_t* meta;  // [0] words used in dst, [1] status, [2 + 2 s] first word of string s in dst, [3 + 2 s] its word count
};
constexpr int kWarpsPerBlock = 8;
constexpr int kStage = 36;  // decoder: staged entries per warp (32 symbols + the entries that end a run)

__global__ void __launch_bounds__(32 * kWarpsPerBlock) rans_dev_encode_kernel(const __grid_constant__ EncParams P) {
  __shared__ uint4 se_all[kWarpsPerBlock][32];  // per warp: this chunk's entries, read back as broadcasts by the chain
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int string = blockIdx.x * (blockDim.x >> 5) + warp;
  if (string >= P.total) return;
  int gi = 0, s = string;
  while (gi + 1 < P.n_groups && s >= P.g[gi].count) s -= P.g[gi++].count;
  const EncGroup& G = P.g[gi];
  uint4* se = se_all[warp];
  const bool kSlots = G.slots != 0;
  const long long n = G.n, cap_words = G.cap_words;
  const int n_rows = G.n_rows;
  const uint32_t n_entries = G.n_entries;
  const int32_t* __restrict__ rows = G.rows;
  const uint4* __restrict__ enc = G.enc;
  const int32_t* __restrict__ symbols = G.symbols + static_cast<long long>(s) * n;
  const int32_t* __restrict__ index = G.index + static_cast<long long>(s) * n;
  uint32_t* __restrict__ out = G.scratch + static_cast<long long>(s) * cap_words;  // filled from the end, like the host coder's sink
  uint32_t* __restrict__ dst = P.dst;
  const long long dst_cap = P.dst_cap;
  int32_t* __restrict__ meta = P.meta;
  unsigned long long x = kRansL;
  long long p = cap_words;
  int status = kStOk;

  // -- off the chain, one chunk ahead of it: a chunk's 32 entries, one per lane (lane 31 = the symbol coded first); the
  // slots of the chunk after that are in flight too, so no load of this loop is waited for --
  struct Resolved {
    uint4 e;
    uint32_t raw;
    bool esc, bad, have;
  };
  auto load_index = [&](long long hi) -> int32_t {
    const long long i = hi - 32 + lane;
    return (hi > 0 && i >= 0) ? index[i] : 0;
  };
  auto resolve = [&](long long hi, int32_t ci) -> Resolved {
    Resolved r;
    const long long i = hi - 32 + lane;
    r.have = hi > 0 && i >= 0;
    r.raw = 0;
    r.esc = r.bad = false;
    uint32_t entry = 0;
    if (r.have) {
      if (kSlots && ci >= 0) {
        entry = static_cast<uint32_t>(ci);
      } else {
        const int row = kSlots ? -(ci + 1) : ci;
        if (static_cast<unsigned>(row) >= static_cast<unsigned>(n_rows) || !rows[3 * n_rows + row]) {
          r.bad = true;
        } else {
          const int32_t last = rows[2 * n_rows + row];
          int32_t value = symbols[i] - rows[n_rows + row];
          if (static_cast<uint32_t>(value) >= static_cast<uint32_t>(last)) {  // negative, or at / beyond the escape bin
            r.esc = true;
            r.raw = value < 0 ? static_cast<uint32_t>(-2 * value - 1) : static_cast<uint32_t>(2 * (value - last));
            value = last;
          }
          entry = static_cast<uint32_t>(rows[row]) + static_cast<uint32_t>(value);
        }
      }
      if (entry >= n_entries) r.bad = true;
    }
    // lanes before the start of the string code a symbol of frequency 2^16 and start 0: the state does not move
    r.e = (r.have && !r.bad) ? enc[entry] : make_uint4(0u, 0u, 0u, 0xffffu);
    return r;
  };
  Resolved nxt = resolve(n, load_index(n));
  int32_t ci_after = load_index(n - 32);
  for (long long hi = n; hi > 0; hi -= 32) {
    const Resolved cur = nxt;
    nxt = resolve(hi - 32, ci_after);
    ci_after = load_index(hi - 64);
    const uint4 e = cur.e;
    const uint32_t raw = cur.raw;
    const bool esc = cur.esc;
    const bool bad = cur.bad || (cur.have && !(e.w >> 24));  // (zero-width bin)
    if (__any_sync(kFull, bad)) { status = kStBad; break; }
    if (p < 160) { status = kStOverflow; break; }  // a chunk writes at most 32 x 3 words
    const unsigned esc_mask = __ballot_sync(kFull, esc);
    uint32_t pend = 0;  // lane j keeps the j-th word this chunk emitted
    int cnt = 0;
    // -- the chain --
    se[lane] = e;
    __syncwarp();
    if (esc_mask == 0u) {
      // straight-line code: the loads of the following symbols issue under the arithmetic of this one
#pragma unroll
      for (int k = 31; k >= 0; --k) enc_symbol(x, se[k], pend, cnt, lane);
    } else {
      // a chunk with out-of-table values: runs of in-table symbols as above (entries fetched two symbols ahead),
      // left at every escape
      const uint32_t se_addr = static_cast<uint32_t>(__cvta_generic_to_shared(se));
      int k = 31;
      while (k >= 0) {
        bool hit = false;
        uint4 e_next = se[k], e_after = se[(k - 1) & 31];
#pragma unroll 8
        for (; k >= 0; --k) {
          if ((esc_mask >> k) & 1u) {
            hit = true;
            break;
          }
          const uint4 ek = e_next;
          e_next = e_after;
          asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(e_after.x), "=r"(e_after.y), "=r"(e_after.z), "=r"(e_after.w)
                       : "r"(se_addr + 16u * ((k - 2) & 31))
                       : "memory");
          enc_symbol(x, ek, pend, cnt, lane);
        }
        if (!hit) break;
        // [main symbol] [count] [nibble 0 .. nibble nb-1] in decoding order, written in reverse
        if (lane < cnt) out[p - 1 - lane] = pend;
        p -= cnt;
        cnt = 0;
        const uint32_t r = __shfl_sync(kFull, raw, k);
        int nb = 0;
        while (nb < 8 && (r >> (nb * 4)) != 0) ++nb;
        for (int j = nb - 1; j >= 0; --j) enc_put4(x, out, p, (r >> (j * 4)) & 15u, lane);
        enc_put4(x, out, p, static_cast<uint32_t>(nb), lane);  // nb <= 8 < 15: one count chunk
        enc_symbol(x, se[k], pend, cnt, lane);
        --k;
      }
    }
    __syncwarp();
    if (lane < cnt) out[p - 1 - lane] = pend;
    p -= cnt;
  }
  if (status == kStOk) {
    if (p < 2) {
      status = kStOverflow;
    } else {
      p -= 2;
      if (lane == 0) {
        out[p] = static_cast<uint32_t>(x);
        out[p + 1] = static_cast<uint32_t>(x >> 32);
      }
    }
  }
  __syncwarp();
  // -- move the string into the shared output buffer (space taken with one atomic; order among strings is free) --
  const long long nwords = status == kStOk ? cap_words - p : 0;
  int off = 0;
  if (lane == 0 && status == kStOk) off = atomicAdd(&meta[0], static_cast<int>(nwords));
  off = __shfl_sync(kFull, off, 0);
  if (status == kStOk && off + nwords > dst_cap) status = kStOverflow;
  if (status == kStOk) {
#pragma unroll 4
    for (long long j = lane; j < nwords; j += 32) dst[off + j] = out[p + j];
  }
  if (lane == 0) {
    if (status != kStOk) atomicMax(&meta[1], status);
    meta[2 + 2 * (P.meta_base + string)] = off;
    meta[3 + 2 * (P.meta_base + string)] = status == kStOk ? static_cast<int>(nwords) : 0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Decoder
// ---------------------------------------------------------------------------------------------------------------
// The string's next 64 words, two per lane (0 past the end: a truncated stream decodes garbage, in bounds)
struct WordWindow {
  const uint32_t* w;
  long long nw, wp, wb;  // words in the string, next word, first word of the window
  uint32_t r0, r1;       // lane l: words wb + l and wb + 32 + l
  __device__ __forceinline__ uint32_t at(long long k, int lane) const { return k + lane < nw ? w[k + lane] : 0u; }
  __device__ __forceinline__ void load(int lane) {
    wb = wp;
    r0 = at(wb, lane);
    r1 = at(wb + 32, lane);
  }
  // after this the next 32 words are in the window
  __device__ __forceinline__ void refill(int lane) {
    if (wp - wb >= 64) {
      load(lane);
    } else if (wp - wb >= 32) {
      wb += 32;
      r0 = r1;
      r1 = at(wb + 32, lane);
    }
  }
  __device__ __forceinline__ uint32_t peek(int idx) const {  // word wb + idx, idx < 64
    const uint32_t a = __shfl_sync(kFull, r0, idx), b = __shfl_sync(kFull, r1, idx);
    return idx < 32 ? a : b;
  }
  __device__ __forceinline__ uint32_t next(int lane) {
    if (wp - wb >= 64) load(lane);
    const uint32_t v = peek(static_cast<int>(wp - wb));
    ++wp;
    return v;
  }
};

__device__ __forceinline__ void dec_renorm(unsigned long long& x, WordWindow& src, int lane) {
  if (x < kRansL) x = (x << 32) | src.next(lane);
}

__device__ __forceinline__ uint32_t dec_get4(unsigned long long& x, WordWindow& src, int lane) {
  const uint32_t v = static_cast<uint32_t>(x) & 15u;
  x >>= 4;
  dec_renorm(x, src, lane);
  return v;
}

// The chain per symbol, common case: the symbol is the MODE of its row (value 0, the centre bin) -- or a symbol the
// caller knows -- whose (start, frequency) every lane already holds: cum - start < freq, one multiply-add, done; the
// lane that owns the symbol has written the mode's value beforehand.  Only another symbol pays for a look at the row:
// a ballot over the 32 bins around the centre, then (rarely) a warp-parallel search of the whole row.
template <bool kCodes>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) rans_dev_decode_kernel(
    const uint32_t* __restrict__ words, const long long* __restrict__ str_off, const long long* __restrict__ str_len,
    const int32_t* __restrict__ codes, long long n, const uint32_t* __restrict__ sf, uint32_t n_entries,
    const int32_t* __restrict__ rows, int n_rows, int count, int32_t* __restrict__ out,
    int32_t* __restrict__ status_out) {
  extern __shared__ int4 smem[];
  int4* srow = smem;  // (first entry, offset, escape bin, usable) of every row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // this chunk's mode / known entries as (start, frequency), one per symbol, followed by entries no state passes
  // (start 2^32 - 1, frequency 0): the chain needs no end-of-chunk test
  uint2* stg = reinterpret_cast<uint2*>(smem + n_rows) + kStage * warp;
  if (lane < kStage - 32) stg[32 + lane] = make_uint2(0xffffffffu, 0u);
  for (int r = threadIdx.x; r < n_rows; r += blockDim.x)
    srow[r] = make_int4(rows[r], rows[n_rows + r], rows[2 * n_rows + r], rows[3 * n_rows + r]);
  __syncthreads();
  const int s = blockIdx.x * (blockDim.x >> 5) + warp;
  if (s >= count) return;
  codes += static_cast<long long>(s) * n;
  out += static_cast<long long>(s) * n;
  WordWindow src;
  src.w = words + str_off[s];
  src.nw = str_len[s];
  if (src.nw < 2) {
    if (lane == 0) atomicMax(status_out, kStBad);
    return;
  }
  unsigned long long x = static_cast<unsigned long long>(src.w[0]) | (static_cast<unsigned long long>(src.w[1]) << 32);
  src.wp = 2;
  src.load(lane);
  int status = kStOk;

  // -- off the chain, one chunk ahead of it: a chunk's codes resolved (one per lane) and its mode entries fetched; the
  // codes of the chunk after that are in flight too --
  struct Resolved {
    // a = first entry of the symbol's row, b = centre bin | escape bin << 16, c = row offset, mode = (start, freq - 1)
    // of the row's mode -- or of the symbol itself when the caller knows it
    uint32_t a, b, mode;
    int32_t c;
    bool known, bad;
  };
  auto load_code = [&](long long i0) -> int32_t { return i0 + lane < n ? codes[i0 + lane] : 0; };
  auto resolve = [&](long long i0, int32_t ci) -> Resolved {
    Resolved r;
    r.a = r.b = r.mode = 0;
    r.c = 0;
    r.known = r.bad = false;
    if (i0 + lane < n) {
      uint32_t m = 0;
      if (kCodes && (ci & kKnownBit)) {
        r.known = true;
        m = static_cast<uint32_t>(ci & (kKnownBit - 1));
      } else if (static_cast<unsigned>(ci) >= static_cast<unsigned>(n_rows)) {
        r.bad = true;
      } else {
        const int4 row = srow[ci];
        if (!row.w || row.z > 65535) {
          r.bad = true;
        } else {
          const int centre = min(max(-row.y, 0), row.z);
          r.a = static_cast<uint32_t>(row.x);
          r.b = static_cast<uint32_t>(centre) | (static_cast<uint32_t>(row.z) << 16);
          r.c = row.y;
          m = r.a + static_cast<uint32_t>(centre);
          if (r.a + static_cast<uint32_t>(row.z) >= n_entries) r.bad = true;
        }
      }
      if (m >= n_entries) r.bad = true;
      if (!r.bad) r.mode = sf[m];
    }
    return r;
  };
  Resolved nxt = resolve(0, load_code(0));
  int32_t ci_after = load_code(32);
  for (long long i0 = 0; i0 < n && status == kStOk; i0 += 32) {
    const long long i = i0 + lane;
    const bool have = i < n;
    const Resolved cur = nxt;
    nxt = resolve(i0 + 32, ci_after);
    ci_after = load_code(i0 + 64);
    const uint32_t a = cur.a, b = cur.b;
    const int32_t c = cur.c;
    const bool known = cur.known, bad = cur.bad;
    if (__any_sync(kFull, bad)) { status = kStBad; break; }
    const unsigned known_mask = __ballot_sync(kFull, known);
    const int count = static_cast<int>(min(static_cast<long long>(32), n - i0));
    stg[lane] = have ? make_uint2(cur.mode & 0xffffu, (cur.mode >> 16) + 1u) : make_uint2(0xffffffffu, 0u);
    int32_t myval = static_cast<int32_t>(b & 0xffffu) + c;  // the mode's value
    __syncwarp();

    // -- the chain --
    int k = 0;
    const uint32_t stg_addr = static_cast<uint32_t>(__cvta_generic_to_shared(stg));
    for (;;) {
      // a run of mode / known symbols that need no renormalisation: straight-line, one multiply-add per symbol,
      // falls through from one symbol to the next and leaves at the first symbol that needs anything else
      uint32_t d, f;
      unsigned long long xn;
      uint2 m0 = stg[k], m1 = stg[k + 1];
#pragma unroll 8
      for (;;) {
        const uint2 m = m0;
        m0 = m1;
        // the entry of the symbol after the next one: a whole iteration passes before it is used
        asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(m1.x), "=r"(m1.y) : "r"(stg_addr + 8u * (k + 2)) : "memory");
        f = m.y;
        d = (static_cast<uint32_t>(x) & 0xffffu) - m.x;
        xn = static_cast<unsigned long long>(f) * (x >> 16) + d;
        if (d >= f || xn < kRansL) break;
        x = xn;
        ++k;
      }
      if (k >= count) break;  // (the entry after the chunk's last symbol)
      const uint32_t cum = static_cast<uint32_t>(x) & 0xffffu;
      if (d < f) {  // the mode after all, with a renormalisation word
        x = (xn << 32) | src.next(lane);
        ++k;
        continue;
      }
      if ((known_mask >> k) & 1u) {  // the stream does not hold the symbol the caller says it holds
        status = kStBad;
        break;
      }
      const uint32_t row0 = __shfl_sync(kFull, a, k), bk = __shfl_sync(kFull, b, k);
      const int32_t off = __shfl_sync(kFull, c, k);
      const int centre = static_cast<int>(bk & 0xffffu), last = static_cast<int>(bk >> 16);
      // the 32 bins around the centre first, then the whole row (rows are strictly increasing CDFs starting at 0)
      const int ws = max(0, min(centre - 15, last - 31));
      int bin = -1;
      {
        const int idx = ws + lane;
        const uint32_t stt = idx <= last ? (sf[row0 + idx] & 0xffffu) : 0x10000u;
        const unsigned hb = __ballot_sync(kFull, stt <= cum);
        const int t = 31 - __clz(hb);
        if (hb != 0u && (t < 31 || ws + 31 >= last)) bin = ws + t;
      }
      if (bin < 0) {
        bin = 0;
        for (int w0 = 0; w0 <= last; w0 += 32) {
          const int idx = w0 + lane;
          const uint32_t stt = idx <= last ? (sf[row0 + idx] & 0xffffu) : 0x10000u;
          const unsigned hb = __ballot_sync(kFull, stt <= cum);
          if (hb == 0u) break;
          bin = w0 + 31 - __clz(hb);
          if (hb != kFull) break;
        }
      }
      const uint32_t e = sf[row0 + bin];
      x = static_cast<unsigned long long>((e >> 16) + 1u) * (x >> 16) + cum - (e & 0xffffu);
      dec_renorm(x, src, lane);
      int32_t value = bin;
      if (bin == last) {  // escape: nibble count in base-15 unary chunks, then the raw value, LSB nibble first
        int32_t val = static_cast<int32_t>(dec_get4(x, src, lane));
        int32_t nb = val;
        while (val == 15 && nb <= 8) {
          val = static_cast<int32_t>(dec_get4(x, src, lane));
          nb += val;
        }
        if (nb > 8) {
          status = kStBad;
          break;
        }
        uint32_t raw = 0;
        for (int j = 0; j < nb; ++j) raw |= dec_get4(x, src, lane) << (j * 4);
        value = static_cast<int32_t>(raw >> 1);
        value = (raw & 1u) ? -value - 1 : value + last;
      }
      if (lane == k) myval = value + off;
      ++k;
    }
    __syncwarp();
    if (status == kStOk && have && !known) out[i] = myval;
  }
  if (status != kStOk && lane == 0) atomicMax(status_out, status);
}

}  // namespace

// A coder block asks for most of an SM's shared memory although it uses little: no CTA of a convolution kernel then
// fits beside it.  A convolution CTA sharing its SM with a dozen coder warps runs late, and with a fixed share of the
// tiles per CTA the whole launch waits for it; the pipeline instead leaves whole SMs to the coder
// (hyres_set_reserved_sms) and this makes sure they are the ones the coder blocks sit on.
constexpr int kHogBytes = 192 * 1024;
static HyPerDevice g_hog_init;

static int coder_init() {
  if (g_hog_init.done()) return HYRES_OK;
  HY_CUDA(cudaFuncSetAttribute(rans_dev_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHogBytes));
  HY_CUDA(cudaFuncSetAttribute(rans_dev_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHogBytes));
  HY_CUDA(cudaFuncSetAttribute(rans_dev_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHogBytes));
  g_hog_init.mark();
  return HYRES_OK;
}

extern "C" {

int hyres_rans_dev_encode(int n_groups, const hyres_rans_group* groups, uint32_t* dst, int64_t dst_cap_words,
                          int32_t* meta, int meta_base, void* stream) {
  if (n_groups < 0 || n_groups > 4 || (n_groups > 0 && !groups) || !meta || meta_base < 0)
    return hy_fail(HYRES_ERR_ARG, "rans_dev_encode: bad argument (at most four groups per launch)");
  EncParams P = {};
  int total = 0;
  for (int k = 0; k < n_groups; ++k) {
    const hyres_rans_group& g = groups[k];
    if (g.count < 0 || g.n < 0) return hy_fail(HYRES_ERR_ARG, "rans_dev_encode: bad group");
    if (g.count == 0) continue;
    if (!g.index || (!g.slots && !g.symbols) || !g.enc || !g.rows || g.n_rows <= 0 || g.n_entries <= 0 ||
        g.n_entries > (1ll << 30) || !g.scratch || g.cap_words < 192)
      return hy_fail(HYRES_ERR_ARG, "rans_dev_encode: bad group");
    EncGroup& G = P.g[P.n_groups++];
    G.symbols = g.symbols;
    G.index = g.index;
    G.enc = static_cast<const uint4*>(g.enc);
    G.rows = g.rows;
    G.scratch = g.scratch;
    G.n = g.n;
    G.cap_words = g.cap_words;
    G.n_entries = static_cast<uint32_t>(g.n_entries);
    G.n_rows = g.n_rows;
    G.count = g.count;
    G.slots = g.slots;
    total += g.count;
  }
  if (total == 0) return HYRES_OK;
  if (!dst || dst_cap_words <= 0 || dst_cap_words > 0x7fffffffll) return hy_fail(HYRES_ERR_ARG, "rans_dev_encode: bad output buffer");
  P.total = total;
  P.meta_base = meta_base;
  P.dst = dst;
  P.dst_cap = dst_cap_words;
  P.meta = meta;
  const int blocks = (total + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int warps = (total + blocks - 1) / blocks;
  if (int rc = coder_init()) return rc;
  rans_dev_encode_kernel<<<blocks, 32 * warps, kHogBytes, static_cast<cudaStream_t>(stream)>>>(P);
  hy_count_launch();
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

int hyres_rans_dev_decode(const uint32_t* words, const int64_t* str_off, const int64_t* str_len, const int32_t* codes,
                          int count, int64_t n, int has_codes, const uint32_t* sf, const int32_t* rows, int n_rows,
                          int64_t n_entries, int32_t* symbols_out, int32_t* status, void* stream) {
  if (count < 0 || n < 0 || !status) return hy_fail(HYRES_ERR_ARG, "rans_dev_decode: bad argument");
  if (count == 0 || n == 0) return HYRES_OK;
  if (!words || !str_off || !str_len || !codes || !sf || !rows || n_rows <= 0 || n_rows > 2048 || n_entries <= 0 ||
      n_entries > (1ll << 26) || !symbols_out)
    return hy_fail(HYRES_ERR_ARG, "rans_dev_decode: bad argument");
  auto st = static_cast<cudaStream_t>(stream);
  const int blocks = (count + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int warps = (count + blocks - 1) / blocks;
  const size_t smem = std::max(static_cast<size_t>(kHogBytes),
                               static_cast<size_t>(n_rows) * sizeof(int4) + static_cast<size_t>(warps) * kStage * sizeof(uint2));
  if (int rc = coder_init()) return rc;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  const long long* so = reinterpret_cast<const long long*>(str_off);
  const long long* sl = reinterpret_cast<const long long*>(str_len);
  if (has_codes)
    rans_dev_decode_kernel<true><<<blocks, 32 * warps, smem, st>>>(words, so, sl, codes, n, sf, static_cast<uint32_t>(n_entries),
                                                                  rows, n_rows, count, symbols_out, status);
  else
    rans_dev_decode_kernel<false><<<blocks, 32 * warps, smem, st>>>(words, so, sl, codes, n, sf, static_cast<uint32_t>(n_entries),
                                                                   rows, n_rows, count, symbols_out, status);
  hy_count_launch();
  HY_CUDA(cudaGetLastError());
  return HYRES_OK;
}

}  // extern "C"
