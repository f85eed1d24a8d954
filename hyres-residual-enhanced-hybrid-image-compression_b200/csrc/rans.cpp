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
// Differences in *how*, not *what* (the coder is the throughput limit of compress + decompress once the
// convolutions run on the GPU: 2 x 1.08 M symbols per 704 x 512 tile, strictly sequential per string):
//   * no intermediate symbol stack (the reverse traversal emits each symbol's escape nibbles directly);
//   * tables are prepared once per CDF set and cached: the encoder divides by multiplying with a 64-bit
//     reciprocal (exact for every state below 2^63), the decoder finds a symbol through a per-row look-up table
//     over 16-wide buckets of the cumulative range plus a short forward scan, start and frequency packed in one word;
//   * renormalisation is branch-free (conditional moves) on both sides;
//   * TWO strings are coded in lock step by one thread: the per-symbol dependency chain (state -> slot -> start /
//     frequency -> state) is latency-bound, so two independent chains nearly double the symbols per core-second;
//   * batched entry points code independent strings on a small thread pool.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "host_util.h"
#include "hyres_b200.h"

namespace {

constexpr uint64_t kRansL = 1ull << 31;
constexpr uint32_t kPrecision = 16;
constexpr uint32_t kBypassBits = 4;
constexpr int32_t kMaxBypass = (1 << kBypassBits) - 1;

// ---------------------------------------------------------------------------------------------------------------
// Prepared tables (cached per CDF set)
// ---------------------------------------------------------------------------------------------------------------
// Encoder entry of one (row, value): x' = ((x / freq) << 16) + x % freq + start = x + start + q * (2^16 - freq) with
// q = x / freq = mulhi(x, rcp_freq) >> rcp_shift (Alverson's reciprocal division as used by ryg_rans' rans64.h:
// exact for all x < 2^63; freq == 1 uses rcp = 2^64 - 1, q = x - 1 and a bias that makes up for it).
struct EncSym {       // 16 bytes; rows are packed back to back (row_base), not at the CDF table's stride
  uint64_t rcp_freq;
  uint32_t bias;
  uint16_t freq_m1;   // freq - 1 (freq in 1 .. 65535)
  uint8_t rcp_shift;
  uint8_t valid;      // 0: zero-width bin -> the encoder fails like the division-based one did
};

constexpr int kLutShift = 4;
constexpr int kLutSize = 1 << (kPrecision - kLutShift);

// per-row facts both loops need, in one cache line read
struct RowInfo {
  int32_t last_bin;  // sizes - 2: the escape bin (max_value of the format)
  int32_t offset;
  uint32_t base;     // first entry of the row in the packed enc / sf arrays
  uint8_t enc_ok, dec_ok;
};

struct Prepared {
  uint64_t hash = 0;
  int n_cdfs = 0, stride = 0;
  std::vector<RowInfo> rows;
  std::vector<int32_t> sizes, offsets;
  std::vector<EncSym> enc;       // packed rows: entry row.base + value
  std::vector<uint32_t> sf;      // packed rows: start | (freq - 1) << 16 of every bin (decoder)
  std::vector<uint16_t> lut;     // [n_cdfs][kLutSize]: the bin that contains the first value of each 16-wide bucket
  std::vector<uint8_t> enc_ok;   // row usable by the encoder (sizes within the table)
  std::vector<uint8_t> dec_ok;   // row usable by the decoder (well-formed, strictly increasing CDF)
};

uint64_t table_hash(const int32_t* cdfs, int n_cdfs, int stride, const int32_t* sizes, const int32_t* offsets) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; };
  mix(static_cast<uint64_t>(n_cdfs));
  mix(static_cast<uint64_t>(stride));
  for (int r = 0; r < n_cdfs; ++r) {
    mix(static_cast<uint32_t>(sizes[r]));
    mix(static_cast<uint32_t>(offsets[r]));
    const int live = std::max(0, std::min(sizes[r], stride));
    const int32_t* row = cdfs + static_cast<int64_t>(r) * stride;
    for (int i = 0; i < live; ++i) mix(static_cast<uint32_t>(row[i]));
  }
  return h;
}

std::shared_ptr<const Prepared> prepare(const int32_t* cdfs, int n_cdfs, int stride, const int32_t* sizes,
                                        const int32_t* offsets) {
  static std::mutex mu;
  static std::vector<std::shared_ptr<const Prepared>> cache;
  const uint64_t h = table_hash(cdfs, n_cdfs, stride, sizes, offsets);
  {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& p : cache)
      if (p->hash == h && p->n_cdfs == n_cdfs && p->stride == stride) return p;
  }
  auto P = std::make_shared<Prepared>();
  P->hash = h; P->n_cdfs = n_cdfs; P->stride = stride;
  P->sizes.assign(sizes, sizes + n_cdfs);
  P->offsets.assign(offsets, offsets + n_cdfs);
  size_t total_bins = 0;
  for (int r = 0; r < n_cdfs; ++r) total_bins += static_cast<size_t>(std::max(1, std::min(sizes[r], stride)));
  P->enc.assign(total_bins, EncSym{0, 0, 0, 0, 0});
  P->sf.assign(total_bins, 0u);
  size_t base = 0;
  P->lut.assign(static_cast<size_t>(n_cdfs) * kLutSize, 0);
  P->enc_ok.assign(n_cdfs, 0);
  P->dec_ok.assign(n_cdfs, 0);
  for (int r = 0; r < n_cdfs; ++r) {
    const int32_t* cdf = cdfs + static_cast<int64_t>(r) * stride;
    const int size = sizes[r];
    const int max_value = size - 2;
    P->enc_ok[r] = (max_value >= 0 && max_value + 1 < stride) ? 1 : 0;
    bool dec_ok = size >= 2 && size <= stride && size <= 65535 && cdf[0] == 0;
    const int bins = std::max(0, std::min(size, stride) - 1);
    for (int v = 0; v < bins; ++v) {
      const uint32_t start = static_cast<uint16_t>(cdf[v]);
      const uint32_t freq = static_cast<uint16_t>(cdf[v + 1] - cdf[v]);  // as the 16-bit arithmetic of the format
      EncSym& e = P->enc[base + v];
      e.valid = freq != 0;
      e.freq_m1 = static_cast<uint16_t>(freq - 1);
      if (freq >= 2) {
        uint32_t shift = 0;
        while (freq > (1u << shift)) ++shift;
        const unsigned __int128 num = (static_cast<unsigned __int128>(1) << (shift + 63)) + freq - 1;
        e.rcp_freq = static_cast<uint64_t>(num / freq);
        e.rcp_shift = static_cast<uint8_t>(shift - 1);
        e.bias = start;
      } else if (freq == 1) {
        e.rcp_freq = ~0ull;
        e.rcp_shift = 0;
        e.bias = start + (1u << kPrecision) - 1;
      }
      const int64_t f = static_cast<int64_t>(cdf[v + 1]) - cdf[v];
      if (f < 1 || f > 65536 || cdf[v] < 0 || cdf[v] > 65535) dec_ok = false;
      else P->sf[base + v] = static_cast<uint32_t>(cdf[v]) | (static_cast<uint32_t>(f - 1) << 16);
    }
    if (dec_ok) {
      int s = 0;
      for (int b = 0; b < kLutSize; ++b) {
        const int32_t c = b << kLutShift;
        while (s + 2 < size && cdf[s + 1] <= c) ++s;
        P->lut[static_cast<size_t>(r) * kLutSize + b] = static_cast<uint16_t>(s);
      }
    }
    P->dec_ok[r] = dec_ok ? 1 : 0;
    P->rows.push_back(RowInfo{max_value, offsets[r], static_cast<uint32_t>(base), P->enc_ok[r], P->dec_ok[r]});
    base += static_cast<size_t>(std::max(1, std::min(size, stride)));
  }
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() >= 8) cache.erase(cache.begin());
  cache.push_back(P);
  return P;
}

// ---------------------------------------------------------------------------------------------------------------
// Encoder
// ---------------------------------------------------------------------------------------------------------------
// Backward-growing word buffer.  `room(k)` guarantees k free words, so the hot loop can store unconditionally.
struct WordSink {
  std::vector<uint32_t> buf;
  uint32_t* p;  // first used word
  explicit WordSink(size_t cap) : buf(std::max<size_t>(cap, 64)), p(buf.data() + buf.size()) {}
  inline void room(size_t k) {
    if (static_cast<size_t>(p - buf.data()) < k) grow(k);
  }
  void grow(size_t k) {
    const size_t used = buf.data() + buf.size() - p;
    std::vector<uint32_t> nb(std::max(buf.size() * 2, used + k + 64));
    std::memcpy(nb.data() + nb.size() - used, p, used * sizeof(uint32_t));
    buf.swap(nb);
    p = buf.data() + buf.size() - used;
  }
  size_t words() const { return buf.data() + buf.size() - p; }
};

// one renormalisation word at most, branch-free: the word is stored below the buffer head and the head only moves
// when the state had to shrink
inline void enc_renorm(uint64_t& x, uint32_t*& p, uint64_t x_max) {
  p[-1] = static_cast<uint32_t>(x);
  const bool r = x >= x_max;
  p -= r ? 1 : 0;
  x = r ? (x >> 32) : x;
}

inline void enc_put_bits(uint64_t& x, uint32_t*& p, uint32_t val, uint32_t nbits) {
  const uint64_t x_max = ((kRansL >> 16) << 32) << (16 - nbits);
  enc_renorm(x, p, x_max);
  x = (x << nbits) | val;
}

// f(0), f(1), ... f(N-1) with compile-time indices: the per-string states of a lock-step group stay in registers
// (no reliance on an unroll pragma, which nvcc's host pass does not forward)
template <typename F, int... K>
inline __attribute__((always_inline)) void for_each_k(F&& f, std::integer_sequence<int, K...>) {
  (f(std::integral_constant<int, K>{}), ...);
}
template <int N, typename F>
inline __attribute__((always_inline)) void unrolled(F&& f) {
  for_each_k(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}

inline uint64_t mulhi64(uint64_t a, uint64_t b) {
  return static_cast<uint64_t>((static_cast<unsigned __int128>(a) * b) >> 64);
}

// escape of an out-of-range value, in reversed order of [main] [count chunks...] [nibble 0 .. nibble n-1]
inline void enc_escape(uint64_t& x, WordSink& sink, uint32_t raw) {
  sink.room(64);
  int32_t n_bypass = 0;
  while (n_bypass < 8 && (raw >> (n_bypass * kBypassBits)) != 0) ++n_bypass;
  for (int32_t j = n_bypass - 1; j >= 0; --j)
    enc_put_bits(x, sink.p, (raw >> (j * kBypassBits)) & kMaxBypass, kBypassBits);
  const int32_t full = n_bypass / kMaxBypass;  // number of saturated (15) chunks
  enc_put_bits(x, sink.p, static_cast<uint32_t>(n_bypass - full * kMaxBypass), kBypassBits);
  for (int32_t k = 0; k < full; ++k) enc_put_bits(x, sink.p, kMaxBypass, kBypassBits);
}

// Device-side front-end (hyres_gc_symbols with coder rows): instead of a CDF row index the coder receives, per
// symbol, a SLOT -- slot >= 0: the entry row.base + value of the packed tables (the value is inside the table);
// slot < 0: row index -(slot + 1) of a value outside it, coded through the escape path from symbols[i].  The
// decoder's CODES: bit 30 set = the symbol is known to the caller (a structurally zero position of the checkerboard
// pass: round(-mean)), low bits = its packed entry, only the range-coder state is advanced and nothing is written;
// otherwise a plain row index.
constexpr int32_t kKnownBit = 1 << 30;

// NS strings of n symbols each, coded in lock step
template <int NS, bool kSlots = false>
int encode_n(const int32_t* const* symbols, const int32_t* const* indexes, int64_t n, const Prepared& T,
             std::vector<uint8_t>* const* out_bytes) {
  WordSink* sinks[NS];
  std::unique_ptr<WordSink> own[NS];
  uint64_t x[NS];
  for (int k = 0; k < NS; ++k) {
    own[k].reset(new WordSink(static_cast<size_t>(n / 2 + 256)));
    sinks[k] = own[k].get();
    x[k] = kRansL;
  }
  const int n_cdfs = T.n_cdfs;
  const RowInfo* rows = T.rows.data();
  const EncSym* enc = T.enc.data();
  const uint32_t n_entries = static_cast<uint32_t>(T.enc.size());
  constexpr int64_t kBlock = 256;  // symbols between two capacity checks (one word per symbol at most + escapes)
  for (int64_t hi = n; hi > 0; hi -= kBlock) {
    const int64_t lo = std::max<int64_t>(hi - kBlock, 0);
    for (int k = 0; k < NS; ++k) sinks[k]->room(kBlock + 64);
    for (int64_t i = hi - 1; i >= lo; --i) {
      int bad = 0;
      unrolled<NS>([&](auto kc) __attribute__((always_inline)) {
        constexpr int k = decltype(kc)::value;
        int32_t ci = indexes[k][i];
        uint32_t entry;
        if (kSlots && ci >= 0) {  // in-table value: the packed entry itself
          entry = static_cast<uint32_t>(ci);
          if (entry >= n_entries) { bad = 1; return; }
        } else {
          if (kSlots) ci = -(ci + 1);
          if (static_cast<uint32_t>(ci) >= static_cast<uint32_t>(n_cdfs)) { bad = 1; return; }
          const RowInfo ri = rows[ci];
          if (!ri.enc_ok) { bad = 1; return; }
          int32_t value = symbols[k][i] - ri.offset;
          if (static_cast<uint32_t>(value) >= static_cast<uint32_t>(ri.last_bin)) {  // negative, or at / beyond the escape bin
            const uint32_t raw = value < 0 ? static_cast<uint32_t>(-2 * value - 1) : static_cast<uint32_t>(2 * (value - ri.last_bin));
            enc_escape(x[k], *sinks[k], raw);
            value = ri.last_bin;
            sinks[k]->room(kBlock + 64);
          }
          entry = ri.base + static_cast<uint32_t>(value);
        }
        const EncSym e = enc[entry];
        if (!e.valid) { bad = 1; return; }
        const uint32_t freq = static_cast<uint32_t>(e.freq_m1) + 1u;
        enc_renorm(x[k], sinks[k]->p, static_cast<uint64_t>(freq) << 47);  // ((L >> 16) << 32) * freq
        const uint64_t q = mulhi64(x[k], e.rcp_freq) >> e.rcp_shift;
        x[k] = x[k] + e.bias + q * ((1u << kPrecision) - freq);
      });
      if (bad) return HYRES_ERR_ARG;
    }
  }
  for (int k = 0; k < NS; ++k) {
    WordSink& sink = *sinks[k];
    sink.room(2);
    *--sink.p = static_cast<uint32_t>(x[k] >> 32);
    *--sink.p = static_cast<uint32_t>(x[k]);
    const size_t nb = sink.words() * 4;
    out_bytes[k]->resize(nb);
    std::memcpy(out_bytes[k]->data(), sink.p, nb);
  }
  return HYRES_OK;
}

int encode_one(const int32_t* symbols, const int32_t* indexes, int64_t n, const Prepared& T, std::vector<uint8_t>& out) {
  std::vector<uint8_t>* o[1] = {&out};
  return encode_n<1>(&symbols, &indexes, n, T, o);
}

// ---------------------------------------------------------------------------------------------------------------
// Decoder
// ---------------------------------------------------------------------------------------------------------------
struct WordSource {
  const uint8_t* p;
  const uint8_t* last;  // address of the last whole word
  // the next word, or 0 past the end (a truncated stream decodes garbage, it never reads out of bounds)
  inline uint32_t peek() const {
    uint32_t w;
    std::memcpy(&w, p <= last ? p : last, 4);
    return p <= last ? w : 0u;
  }
};

// branch-free renormalisation: at most one word
inline void dec_renorm(uint64_t& x, WordSource& src) {
  const uint32_t w = src.peek();
  const bool r = x < kRansL;
  x = r ? ((x << 32) | w) : x;
  src.p += r ? 4 : 0;
}

inline uint32_t dec_get_bits(uint64_t& x, WordSource& src, uint32_t nbits) {
  const uint32_t val = static_cast<uint32_t>(x & ((1u << nbits) - 1));
  x >>= nbits;
  dec_renorm(x, src);
  return val;
}

// the escape that follows the last bin: nibble count in base-15 unary chunks, then the raw value, LSB nibble first
inline int dec_escape(uint64_t& x, WordSource& src, int32_t max_value, int32_t& value) {
  int32_t val = static_cast<int32_t>(dec_get_bits(x, src, kBypassBits));
  int32_t n_bypass = val;
  while (val == kMaxBypass) {
    val = static_cast<int32_t>(dec_get_bits(x, src, kBypassBits));
    n_bypass += val;
    if (n_bypass > 8) return HYRES_ERR_ARG;
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
  return HYRES_OK;
}

template <int NS, bool kCodes = false>
int decode_n(const uint8_t* const* in, const int64_t* in_len, const int32_t* const* indexes, int64_t n,
             const Prepared& T, int32_t* const* out) {
  WordSource src[NS];
  uint64_t x[NS];
  for (int k = 0; k < NS; ++k) {
    if (in_len[k] < 8) return HYRES_ERR_ARG;
    src[k] = WordSource{in[k], in[k] + in_len[k] - 4};
    uint32_t lo = src[k].peek();
    src[k].p += 4;
    uint32_t hi = src[k].peek();
    src[k].p += 4;
    x[k] = lo | (static_cast<uint64_t>(hi) << 32);
  }
  const int n_cdfs = T.n_cdfs;
  const RowInfo* rows = T.rows.data();
  const uint32_t* sf = T.sf.data();
  const uint16_t* lut = T.lut.data();
  const uint32_t n_entries = static_cast<uint32_t>(T.sf.size());
  for (int64_t i = 0; i < n; ++i) {
    int bad = 0;
    unrolled<NS>([&](auto kc) __attribute__((always_inline)) {
      constexpr int k = decltype(kc)::value;
      const int32_t ci = indexes[k][i];
      if (kCodes && (ci & kKnownBit)) {  // known symbol: advance the state by its (start, frequency), nothing to find
        const uint32_t entry = static_cast<uint32_t>(ci & (kKnownBit - 1));
        if (entry >= n_entries) { bad = 1; return; }
        const uint32_t e = sf[entry];
        x[k] = static_cast<uint64_t>((e >> 16) + 1) * (x[k] >> kPrecision) + (x[k] & ((1u << kPrecision) - 1)) - (e & 0xffffu);
        dec_renorm(x[k], src[k]);
        return;
      }
      if (static_cast<uint32_t>(ci) >= static_cast<uint32_t>(n_cdfs)) { bad = 1; return; }
      const RowInfo ri = rows[ci];
      if (!ri.dec_ok) { bad = 1; return; }
      const uint32_t* row = sf + ri.base;
      const int32_t last_bin = ri.last_bin;  // == max_value: the escape bin
      const uint32_t cum = static_cast<uint32_t>(x[k] & ((1u << kPrecision) - 1));
      // last bin whose start is <= cum (strictly increasing CDF => the reference's linear scan finds the same one)
      int32_t s = lut[static_cast<size_t>(ci) * kLutSize + (cum >> kLutShift)];
      while (s < last_bin && (row[s + 1] & 0xffffu) <= cum) ++s;
      const uint32_t e = row[s];
      const uint32_t start = e & 0xffffu, freq = (e >> 16) + 1;
      x[k] = static_cast<uint64_t>(freq) * (x[k] >> kPrecision) + cum - start;
      dec_renorm(x[k], src[k]);
      int32_t value = s;
      if (s == last_bin && dec_escape(x[k], src[k], last_bin, value) != HYRES_OK) { bad = 1; return; }
      out[k][i] = value + ri.offset;
    });
    if (bad) return HYRES_ERR_ARG;
  }
  return HYRES_OK;
}

inline int host_cores();

template <typename F>
int run_pool(int count, int threads, F&& job) {
  if (count <= 0) return HYRES_OK;
  int nt = std::max(1, std::min(std::min(threads > 0 ? threads : host_cores(), host_cores()), count));
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

// Jobs of a batch: strings of equal length are coded in lock step in groups of up to G (HYRES_RANS_GROUP overrides
// the caller's choice; 1 codes every string on its own).
constexpr int kMaxGroup = 4;
// Strings being coded right now by all callers of this process.  Lock-step decoding of two strings costs ~0.6 of
// the core time of decoding them one after the other (the decoder's chain state -> bucket -> bin -> state is
// latency-bound) but occupies half as many cores for 1.2x as long: it pays exactly when the cores are
// oversubscribed (several images in flight), not for a lone call.  The encoder gains less per symbol (5.9 -> 5.5 ns)
// but follows the same policy: half as many threads contend for the cores.
std::atomic<int> g_active_strings{0};
struct ActiveStrings {
  int n;
  explicit ActiveStrings(int k) : n(k) { g_active_strings.fetch_add(n); }
  ~ActiveStrings() { g_active_strings.fetch_sub(n); }
};
// Host cores this process may count on: all of them, divided by the number of processes torchrun started on this
// node (one per GPU share the box's cores); HYRES_HOST_CORES overrides.
inline int host_cores() {
  static const int n = [] {
    if (const char* e = getenv("HYRES_HOST_CORES")) return std::max(1, atoi(e));
    int hw = static_cast<int>(std::max(1u, std::thread::hardware_concurrency()));
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) hw = std::max(1, hw / std::max(1, atoi(e)));
    return hw;
  }();
  return n;
}
// lock-step group size for a batch of `count` strings: 1 while the cores are free, 2 when oversubscribed, 4 when
// oversubscribed four times over (fewer, longer jobs: 12.7 / 6.9 / 5.9 ns per decoded symbol and core)
inline int group_for(int count) {
  const int load = g_active_strings.load() + count, hw = host_cores();
  return load > 4 * hw ? 4 : load > hw ? 2 : 1;
}
struct Job { int id[kMaxGroup]; int count; };
std::vector<Job> make_jobs(int count, const int64_t* n, int dflt) {
  static const int env = [] { const char* e = getenv("HYRES_RANS_GROUP"); return e ? atoi(e) : 0; }();
  const int G = std::max(1, std::min(kMaxGroup, env > 0 ? env : dflt));
  std::vector<Job> jobs;
  std::vector<char> used(count, 0);
  for (int i = 0; i < count; ++i) {
    if (used[i]) continue;
    Job j{{i, -1, -1, -1}, 1};
    used[i] = 1;
    for (int t = i + 1; t < count && j.count < G; ++t)
      if (!used[t] && n[t] == n[i]) { j.id[j.count++] = t; used[t] = 1; }
    if (j.count == 3) { used[j.id[2]] = 0; j.id[2] = -1; j.count = 2; }  // groups of 1, 2 or 4
    jobs.push_back(j);
  }
  return jobs;
}

template <int NS, bool kSlots>
int encode_group(const Job& j, const int32_t* const* symbols, const int32_t* const* indexes, int64_t n, const Prepared& T,
                 std::vector<uint8_t>* bytes) {
  const int32_t* sy[NS];
  const int32_t* ix[NS];
  std::vector<uint8_t>* o[NS];
  for (int k = 0; k < NS; ++k) { sy[k] = symbols[j.id[k]]; ix[k] = indexes[j.id[k]]; o[k] = &bytes[k]; }
  return encode_n<NS, kSlots>(sy, ix, n, T, o);
}

template <int NS, bool kCodes>
int decode_group(const Job& j, const uint8_t* const* in, const int64_t* in_len, const int32_t* const* indexes, int64_t n,
                 const Prepared& T, int32_t* const* out) {
  const uint8_t* ins[NS];
  int64_t lens[NS];
  const int32_t* ix[NS];
  int32_t* o[NS];
  for (int k = 0; k < NS; ++k) { ins[k] = in[j.id[k]]; lens[k] = in_len[j.id[k]]; ix[k] = indexes[j.id[k]]; o[k] = out[j.id[k]]; }
  return decode_n<NS, kCodes>(ins, lens, ix, n, T, o);
}

}  // namespace

namespace {

template <bool kSlots>
int encode_batch_impl(int count, const int32_t* const* symbols, const int32_t* const* indexes, const int64_t* n,
                      const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                      uint8_t* const* out, const int64_t* out_cap, int64_t* out_len, int threads) {
  if (count < 0 || (count > 0 && (!symbols || !indexes || !n || !out || !out_cap || !out_len)))
    return hy_fail(HYRES_ERR_ARG, "rans_encode_batch: bad argument");
  if (count == 0) return HYRES_OK;
  if (n_cdfs <= 0 || cdf_stride <= 0 || !cdfs || !cdf_sizes || !offsets)
    return hy_fail(HYRES_ERR_ARG, "rans_encode_batch: bad tables");
  const auto T = prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets);
  const auto jobs = make_jobs(count, n, group_for(count));  // fewer threads when the cores are oversubscribed
  const ActiveStrings active(count);
  const int rc = run_pool(static_cast<int>(jobs.size()), threads, [&](int ji) {
    const Job& j = jobs[ji];
    std::vector<uint8_t> bytes[kMaxGroup];
    const int64_t len = n[j.id[0]];
    int r = j.count == 4   ? encode_group<4, kSlots>(j, symbols, indexes, len, *T, bytes)
            : j.count == 2 ? encode_group<2, kSlots>(j, symbols, indexes, len, *T, bytes)
                           : encode_group<1, kSlots>(j, symbols, indexes, len, *T, bytes);
    if (r != HYRES_OK) return r;
    for (int k = 0; k < j.count; ++k) {
      const int i = j.id[k];
      out_len[i] = static_cast<int64_t>(bytes[k].size());
      if (out_cap[i] < out_len[i]) { r = HYRES_ERR_ARG; continue; }
      std::memcpy(out[i], bytes[k].data(), bytes[k].size());
    }
    return r;
  });
  if (rc != HYRES_OK) return hy_fail(rc, "rans_encode_batch: a string failed (bad tables or buffer too small)");
  return HYRES_OK;
}

template <bool kCodes>
int decode_batch_impl(int count, const uint8_t* const* in, const int64_t* in_len, const int32_t* const* indexes,
                      const int64_t* n, const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                      const int32_t* offsets, int32_t* const* symbols_out, int threads) {
  if (count < 0 || (count > 0 && (!in || !in_len || !indexes || !n || !symbols_out)))
    return hy_fail(HYRES_ERR_ARG, "rans_decode_batch: bad argument");
  if (n_cdfs <= 0 || cdf_stride <= 0 || !cdfs || !cdf_sizes || !offsets) return hy_fail(HYRES_ERR_ARG, "rans_decode_batch: bad tables");
  if (count == 0) return HYRES_OK;
  const auto T = prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets);
  const auto jobs = make_jobs(count, n, group_for(count));
  const ActiveStrings active(count);
  const int rc = run_pool(static_cast<int>(jobs.size()), threads, [&](int ji) {
    const Job& j = jobs[ji];
    const int64_t len = n[j.id[0]];
    return j.count == 4   ? decode_group<4, kCodes>(j, in, in_len, indexes, len, *T, symbols_out)
           : j.count == 2 ? decode_group<2, kCodes>(j, in, in_len, indexes, len, *T, symbols_out)
                          : decode_group<1, kCodes>(j, in, in_len, indexes, len, *T, symbols_out);
  });
  if (rc != HYRES_OK) return hy_fail(rc, "rans_decode_batch: a string failed");
  return HYRES_OK;
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
  if (n < 0 || (n > 0 && (!symbols || !indexes)) || !cdfs || !cdf_sizes || !offsets || !out_len || n_cdfs <= 0 ||
      cdf_stride <= 0)
    return hy_fail(HYRES_ERR_ARG, "rans_encode: bad argument");
  const auto T = prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets);
  std::vector<uint8_t> bytes;
  const int rc = encode_one(symbols, indexes, n, *T, bytes);
  if (rc != HYRES_OK) return hy_fail(rc, "rans_encode: index / cdf table out of range");
  *out_len = static_cast<int64_t>(bytes.size());
  if (!out || out_cap < *out_len) return hy_fail(HYRES_ERR_ARG, "rans_encode: output buffer too small (see *out_len)");
  std::memcpy(out, bytes.data(), bytes.size());
  return HYRES_OK;
}

int hyres_rans_decode(const uint8_t* in, int64_t in_len, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                      int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                      int32_t* symbols_out) {
  if (!in || n < 0 || (n > 0 && (!indexes || !symbols_out)) || !cdfs || !cdf_sizes || !offsets || n_cdfs <= 0 ||
      cdf_stride <= 0)
    return hy_fail(HYRES_ERR_ARG, "rans_decode: bad argument");
  const auto T = prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets);
  const int rc = decode_n<1>(&in, &in_len, &indexes, n, *T, &symbols_out);
  if (rc != HYRES_OK) return hy_fail(rc, "rans_decode: malformed stream or tables");
  return HYRES_OK;
}

int hyres_rans_encode_batch(int count, const int32_t* const* symbols, const int32_t* const* indexes, const int64_t* n,
                            const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, uint8_t* const* out, const int64_t* out_cap, int64_t* out_len,
                            int threads) {
  return encode_batch_impl<false>(count, symbols, indexes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, out, out_cap,
                                  out_len, threads);
}

int hyres_rans_encode_slots_batch(int count, const int32_t* const* symbols, const int32_t* const* slots, const int64_t* n,
                                  const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                                  const int32_t* offsets, uint8_t* const* out, const int64_t* out_cap, int64_t* out_len,
                                  int threads) {
  return encode_batch_impl<true>(count, symbols, slots, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, out, out_cap,
                                 out_len, threads);
}

int hyres_rans_decode_batch(int count, const uint8_t* const* in, const int64_t* in_len, const int32_t* const* indexes,
                            const int64_t* n, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, int32_t* const* symbols_out,
                            int threads) {
  return decode_batch_impl<false>(count, in, in_len, indexes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, symbols_out,
                                  threads);
}

int hyres_rans_decode_codes_batch(int count, const uint8_t* const* in, const int64_t* in_len, const int32_t* const* codes,
                                  const int64_t* n, const int32_t* cdfs, int n_cdfs, int cdf_stride,
                                  const int32_t* cdf_sizes, const int32_t* offsets, int32_t* const* symbols_out,
                                  int threads) {
  return decode_batch_impl<true>(count, in, in_len, codes, n, cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets, symbols_out,
                                 threads);
}

int hyres_rans_table_layout(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, int32_t* rows_out) {
  if (!cdfs || !cdf_sizes || !offsets || !rows_out || n_cdfs <= 0 || cdf_stride <= 0)
    return hy_fail(HYRES_ERR_ARG, "rans_table_layout: bad argument");
  const auto T = prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets);
  for (int r = 0; r < n_cdfs; ++r) {
    const RowInfo& ri = T->rows[r];
    const bool ok = ri.enc_ok && ri.dec_ok;
    rows_out[r] = static_cast<int32_t>(ri.base);
    rows_out[n_cdfs + r] = ri.offset;
    rows_out[2 * n_cdfs + r] = ok ? ri.last_bin : 0;  // 0: no in-table value, every symbol takes the plain path
  }
  return HYRES_OK;
}

int64_t hyres_rans_table_entries(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                                 const int32_t* offsets) {
  if (!cdfs || !cdf_sizes || !offsets || n_cdfs <= 0 || cdf_stride <= 0) {
    hy_fail(HYRES_ERR_ARG, "rans_table_entries: bad argument");
    return -1;
  }
  return static_cast<int64_t>(prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets)->enc.size());
}

int hyres_rans_table_export(const int32_t* cdfs, int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, void* enc_out, uint32_t* sf_out, int32_t* rows_out) {
  if (!cdfs || !cdf_sizes || !offsets || !enc_out || !sf_out || !rows_out || n_cdfs <= 0 || cdf_stride <= 0)
    return hy_fail(HYRES_ERR_ARG, "rans_table_export: bad argument");
  static_assert(sizeof(EncSym) == 16, "the device coder reads encoder entries as 16-byte words");
  const auto T = prepare(cdfs, n_cdfs, cdf_stride, cdf_sizes, offsets);
  std::memcpy(enc_out, T->enc.data(), T->enc.size() * sizeof(EncSym));
  std::memcpy(sf_out, T->sf.data(), T->sf.size() * sizeof(uint32_t));
  for (int r = 0; r < n_cdfs; ++r) {
    const RowInfo& ri = T->rows[r];
    rows_out[r] = static_cast<int32_t>(ri.base);
    rows_out[n_cdfs + r] = ri.offset;
    rows_out[2 * n_cdfs + r] = ri.last_bin;
    rows_out[3 * n_cdfs + r] = (ri.enc_ok && ri.dec_ok) ? 1 : 0;
  }
  return HYRES_OK;
}

}  // extern "C"
