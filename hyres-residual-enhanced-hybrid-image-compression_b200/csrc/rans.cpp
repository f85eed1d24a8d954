// Host entropy coder of the product path: rANS64 with CompressAI's bit-stream format
// (compressai 1.2.6 `ans` module: RansEncoder.encode_with_indexes /
// RansDecoder.decode_with_indexes) and CompressAI's pmf -> quantised CDF rule
// (compressai._CXX.pmf_to_quantized_cdf).  Reached in the reference from
// models/checkerboard.py:159-165,172-173,206 and :261-267.
//
// Format facts reproduced here:
//   * 64-bit rANS state, lower bound L = 2^31, 32-bit renormalisation words, 16-bit
//     probability precision;
//   * symbols are coded in reverse so the decoder pops them in order; the byte string is
//     the little-endian word buffer [state lo, state hi, word, word, ...];
//   * out-of-range values use the last CDF bin as an escape followed by a "bypass" code:
//     nibble count in base-15 unary chunks, then the raw value as 4-bit nibbles LSB first.
//
// Differences in *how*, not *what*: no intermediate symbol stack (the reverse traversal
// emits each symbol's escape nibbles directly), binary search instead of a linear CDF
// scan in the decoder (a per-row bucket look-up table plus a short forward scan), flat int32 tables instead of nested Python lists, and batched
// entry points that code independent strings on a small thread pool.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr uint64_t kRansL = 1ull << 31;
constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr int32_t kMaxBypass = (1 << kBypassBits) - 1;

// Backward-growing word buffer.
struct WordSink {
  std::vector<uint32_t> buf;
  size_t pos;  // index of the first used word
  explicit WordSink(size_t cap) : buf(std::max<size_t>(cap, 16)), pos(buf.size()) {}
  inline void push(uint32_t w) {
    if (pos == 0) grow();
    buf[--pos] = w;
  }
  void grow() {
    const size_t used = buf.size();
    std::vector<uint32_t> nb(used * 2);
    std::memcpy(nb.data() + used, buf.data(), used * sizeof(uint32_t));
    buf.swap(nb);
    pos = used;
  }
  size_t words() const { return buf.size() - pos; }
};

inline void enc_put(uint64_t& x, WordSink& out, uint32_t start, uint32_t freq) {
  const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
  if (x >= x_max) {
    out.push(static_cast<uint32_t>(x));
    x >>= 32;
  }
  x = ((x / freq) << kPrecision) + (x % freq) + start;
}

inline void enc_put_bits(uint64_t& x, WordSink& out, uint32_t val, uint32_t nbits) {
  const uint32_t freq = 1u << (16 - nbits);
  const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
  if (x >= x_max) {
    out.push(static_cast<uint32_t>(x));
    x >>= 32;
  }
  x = (x << nbits) | val;
}

int encode_one(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs, int n_cdfs,
               int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, std::vector<uint8_t>& out_bytes) {
  WordSink sink(static_cast<size_t>(n / 2 + 64));
  uint64_t x = kRansL;
  for (int64_t i = n - 1; i >= 0; --i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return HYRES_ERR_ARG;
    const int32_t* cdf = cdfs + static_cast<int64_t>(ci) * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    if (max_value < 0 || max_value + 1 >= cdf_stride) return HYRES_ERR_ARG;
    int32_t value = symbols[i] - offsets[ci];
    uint32_t raw = 0;
    bool escape = false;
    if (value < 0) {
      raw = static_cast<uint32_t>(-2 * value - 1);
      value = max_value;
      escape = true;
    } else if (value >= max_value) {
      raw = static_cast<uint32_t>(2 * (value - max_value));
      value = max_value;
      escape = true;
    }
    if (escape) {
      // reversed order of: [main] [count chunks...] [nibble 0 .. nibble n-1]
      int32_t n_bypass = 0;
      while ((raw >> (n_bypass * kBypassBits)) != 0) ++n_bypass;
      for (int32_t j = n_bypass - 1; j >= 0; --j)
        enc_put_bits(x, sink, (raw >> (j * kBypassBits)) & kMaxBypass, kBypassBits);
      int32_t full = n_bypass / kMaxBypass;  // number of saturated (15) chunks
      enc_put_bits(x, sink, static_cast<uint32_t>(n_bypass - full * kMaxBypass), kBypassBits);
      for (int32_t k = 0; k < full; ++k) enc_put_bits(x, sink, kMaxBypass, kBypassBits);
    }
    const uint32_t start = static_cast<uint16_t>(cdf[value]);
    const uint32_t freq = static_cast<uint16_t>(cdf[value + 1] - cdf[value]);
    if (freq == 0) return HYRES_ERR_ARG;
    enc_put(x, sink, start, freq);
  }
  sink.push(static_cast<uint32_t>(x >> 32));
  sink.push(static_cast<uint32_t>(x));
  const size_t nb = sink.words() * 4;
  out_bytes.resize(nb);
  std::memcpy(out_bytes.data(), sink.buf.data() + sink.pos, nb);
  return HYRES_OK;
}

struct WordSource {
  const uint8_t* p;
  const uint8_t* end;
  inline uint32_t next() {
    uint32_t w = 0;
    if (p + 4 <= end) {
      std::memcpy(&w, p, 4);
      p += 4;
    }
    return w;
  }
};

inline uint32_t dec_get_bits(uint64_t& x, WordSource& src, uint32_t nbits) {
  const uint32_t val = static_cast<uint32_t>(x & ((1u << nbits) - 1));
  x >>= nbits;
  if (x < kRansL) x = (x << 32) | src.next();
  return val;
}

// Decoder look-up: for every CDF row, the symbol that contains the start of each 64-wide bucket of the 16-bit
// cumulative range (1024 buckets).  A symbol is then found by a short forward scan from lut[cum >> 6] instead of
// a 12-step binary search over up to 3133 entries (strictly increasing CDF: same result as the reference's
// linear scan).  128 KB for the 64-row Gaussian table: L2-resident.
constexpr int kLutShift = 6;
constexpr int kLutSize = 1 << (kPrecision - kLutShift);

struct DecodeLut {
  std::vector<uint16_t> lut;    // [n_cdfs][kLutSize]
  std::vector<uint8_t> row_ok;  // rows with a malformed table fail only when a symbol refers to them
  DecodeLut(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes)
      : lut(static_cast<size_t>(std::max(n_cdfs, 0)) * kLutSize), row_ok(static_cast<size_t>(std::max(n_cdfs, 0)), 1) {
    for (int r = 0; r < n_cdfs; ++r) {
      const int32_t* cdf = cdfs + static_cast<int64_t>(r) * cdf_stride;
      const int size = cdf_sizes[r];
      if (size < 2 || size > cdf_stride || size > 65535 || cdf[0] != 0) { row_ok[r] = 0; continue; }
      int s = 0;
      for (int b = 0; b < kLutSize; ++b) {
        const int32_t c = b << kLutShift;
        while (s + 2 < size && cdf[s + 1] <= c) ++s;
        lut[static_cast<size_t>(r) * kLutSize + b] = static_cast<uint16_t>(s);
      }
    }
  }
};

int decode_one(const uint8_t* in, int64_t in_len, const int32_t* indexes, int64_t n, const int32_t* cdfs,
               int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out,
               const DecodeLut& dl) {
  if (in_len < 8) return HYRES_ERR_ARG;
  WordSource src{in, in + in_len};
  uint64_t x = src.next();
  x |= static_cast<uint64_t>(src.next()) << 32;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs || !dl.row_ok[ci]) return HYRES_ERR_ARG;
    const int32_t* cdf = cdfs + static_cast<int64_t>(ci) * cdf_stride;
    const int32_t size = cdf_sizes[ci];
    const int32_t max_value = size - 2;
    const uint32_t cum = static_cast<uint32_t>(x & ((1u << kPrecision) - 1));
    // last entry <= cum (strictly increasing CDF => same result as the reference's linear scan)
    int32_t s = dl.lut[static_cast<size_t>(ci) * kLutSize + (cum >> kLutShift)];
    while (s + 2 < size && static_cast<uint32_t>(cdf[s + 1]) <= cum) ++s;
    if (s + 1 >= cdf_stride) return HYRES_ERR_ARG;
    const uint32_t start = static_cast<uint32_t>(cdf[s]);
    const uint32_t freq = static_cast<uint32_t>(cdf[s + 1] - cdf[s]);
    x = static_cast<uint64_t>(freq) * (x >> kPrecision) + (x & ((1u << kPrecision) - 1)) - start;
    if (x < kRansL) x = (x << 32) | src.next();
    int32_t value = s;
    if (value == max_value) {
      int32_t val = static_cast<int32_t>(dec_get_bits(x, src, kBypassBits));
      int32_t n_bypass = val;
      while (val == kMaxBypass) {
        val = static_cast<int32_t>(dec_get_bits(x, src, kBypassBits));
        n_bypass += val;
      }
      if (n_bypass > 8) return HYRES_ERR_ARG;  // more than 32 raw bits: malformed stream
      uint32_t raw = 0;
      for (int32_t j = 0; j < n_bypass; ++j) {
        val = static_cast<int32_t>(dec_get_bits(x, src, kBypassBits));
        raw |= static_cast<uint32_t>(val) << (j * kBypassBits);
      }
      value = static_cast<int32_t>(raw >> 1);
      if (raw & 1) value = -value - 1;
      else value += max_value;
    }
    out[i] = value + offsets[ci];
  }
  return HYRES_OK;
}

template <typename F>
int run_pool(int count, int threads, F&& job) {
  if (count <= 0) return HYRES_OK;
  int nt = std::max(1, std::min(threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency()), count));
  std::atomic<int> next{0};
  std::atomic<int> status{HYRES_OK};
  auto worker = [&]() {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= count) return;
      const int rc = job(i);
      if (rc != HYRES_OK) status.store(rc);
    }
  };
  if (nt == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(nt);
    for (int t = 0; t < nt; ++t) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
  }
  return status.load();
}

}  // namespace

extern "C" {

int hyres_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* out) {
  if (!pmf || !out || n <= 0 || precision < 1 || precision > 16) return hy_fail(HYRES_ERR_ARG, "pmf_to_quantized_cdf: bad argument");
  for (int i = 0; i < n; ++i)
    if (!(pmf[i] >= 0.f) || !std::isfinite(pmf[i])) return hy_fail(HYRES_ERR_ARG, "pmf_to_quantized_cdf: invalid pmf entry");
  std::vector<uint32_t> cdf(static_cast<size_t>(n) + 1);
  cdf[0] = 0;
  const float scale = static_cast<float>(1 << precision);
  for (int i = 0; i < n; ++i) cdf[i + 1] = static_cast<uint32_t>(std::round(pmf[i] * scale));
  uint32_t total = 0;
  for (uint32_t v : cdf) total += v;
  if (total == 0) return hy_fail(HYRES_ERR_ARG, "pmf_to_quantized_cdf: pmf sums to zero");
  for (auto& v : cdf) v = static_cast<uint32_t>((static_cast<uint64_t>(1u << precision) * v) / total);
  for (size_t i = 1; i < cdf.size(); ++i) cdf[i] += cdf[i - 1];
  cdf.back() = 1u << precision;
  const int m = static_cast<int>(cdf.size());
  for (int i = 0; i < m - 1; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    // zero-width bin: take one count from the narrowest bin wider than 1
    uint32_t best_freq = ~0u;
    int best = -1;
    for (int j = 0; j < m - 1; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best_freq) { best_freq = f; best = j; }
    }
    if (best < 0) return hy_fail(HYRES_ERR_ARG, "pmf_to_quantized_cdf: cannot make every bin non-empty");
    if (best < i) {
      for (int j = best + 1; j <= i; ++j) --cdf[j];
    } else {
      for (int j = i + 1; j <= best; ++j) ++cdf[j];
    }
  }
  std::memcpy(out, cdf.data(), cdf.size() * sizeof(uint32_t));
  return HYRES_OK;
}

int64_t hyres_rans_encode_bound(int64_t n) { return n < 0 ? 8 : 2 * n + 1024; }

int hyres_rans_encode(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs, int n_cdfs,
                      int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, uint8_t* out, int64_t out_cap,
                      int64_t* out_len) {
  if (n < 0 || (n > 0 && (!symbols || !indexes)) || !cdfs || !cdf_sizes || !offsets || !out_len)
    return hy_fail(HYRES_ERR_ARG, "rans_encode: bad argument");
  std::vector<uint8_t> bytes;
  const int rc = encode_one(symbols, indexes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, bytes);
  if (rc != HYRES_OK) return hy_fail(rc, "rans_encode: index / cdf table out of range");
  *out_len = static_cast<int64_t>(bytes.size());
  if (!out || out_cap < *out_len) return hy_fail(HYRES_ERR_ARG, "rans_encode: output buffer too small (see *out_len)");
  std::memcpy(out, bytes.data(), bytes.size());
  return HYRES_OK;
}

int hyres_rans_decode(const uint8_t* in, int64_t in_len, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                      int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                      int32_t* symbols_out) {
  if (!in || n < 0 || (n > 0 && (!indexes || !symbols_out)) || !cdfs || !cdf_sizes || !offsets)
    return hy_fail(HYRES_ERR_ARG, "rans_decode: bad argument");
  const DecodeLut dl(cdfs, n_cdfs, cdf_stride, cdf_sizes);
  const int rc = decode_one(in, in_len, indexes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, symbols_out, dl);
  if (rc != HYRES_OK) return hy_fail(rc, "rans_decode: malformed stream or tables");
  return HYRES_OK;
}

int hyres_rans_encode_batch(int count, const int32_t* const* symbols, const int32_t* const* indexes, const int64_t* n,
                            const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, uint8_t* const* out, const int64_t* out_cap, int64_t* out_len,
                            int threads) {
  if (count < 0 || (count > 0 && (!symbols || !indexes || !n || !out || !out_cap || !out_len)))
    return hy_fail(HYRES_ERR_ARG, "rans_encode_batch: bad argument");
  const int rc = run_pool(count, threads, [&](int i) {
    std::vector<uint8_t> bytes;
    int r = encode_one(symbols[i], indexes[i], n[i], cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, bytes);
    if (r != HYRES_OK) return r;
    out_len[i] = static_cast<int64_t>(bytes.size());
    if (out_cap[i] < out_len[i]) return HYRES_ERR_ARG;
    std::memcpy(out[i], bytes.data(), bytes.size());
    return HYRES_OK;
  });
  if (rc != HYRES_OK) return hy_fail(rc, "rans_encode_batch: a string failed (bad tables or buffer too small)");
  return HYRES_OK;
}

int hyres_rans_decode_batch(int count, const uint8_t* const* in, const int64_t* in_len, const int32_t* const* indexes,
                            const int64_t* n, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, int32_t* const* symbols_out,
                            int threads) {
  if (count < 0 || (count > 0 && (!in || !in_len || !indexes || !n || !symbols_out)))
    return hy_fail(HYRES_ERR_ARG, "rans_decode_batch: bad argument");
  if (n_cdfs <= 0 || !cdfs || !cdf_sizes || !offsets) return hy_fail(HYRES_ERR_ARG, "rans_decode_batch: bad tables");
  const DecodeLut dl(cdfs, n_cdfs, cdf_stride, cdf_sizes);
  const int rc = run_pool(count, threads, [&](int i) {
    return decode_one(in[i], in_len[i], indexes[i], n[i], cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets,
                      symbols_out[i], dl);
  });
  if (rc != HYRES_OK) return hy_fail(rc, "rans_decode_batch: a string failed");
  return HYRES_OK;
}

}  // extern "C"
