/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Parity unpinned (see oracle/README.md).
 *
 * Plain-C restatement of the entropy coder the reference reaches through
 * compressai 1.2.6 (pinned by /root/reference/requirements.txt:11, source not
 * vendored and not installed here):
 *   - compressai/cpp_exts/rans/rans_interface.cpp
 *       BufferedRansEncoder::encode_with_indexes / flush, RansDecoder::decode_with_indexes
 *   - compressai/cpp_exts/ops/ops.cpp  pmf_to_quantized_cdf
 *   - third_party/ryg_rans/rans64.h    Rans64Enc* / Rans64Dec*
 * called from /root/reference/models/checkerboard.py:159-165 (_compress_part /
 * _decompress_part), :172-173, :206 (EntropyBottleneck.compress/decompress) and
 * :261-267 (update -> _pmf_to_cdf).
 *
 * It follows the published algorithm literally (symbol stack, LIFO flush, linear CDF
 * scan) so that it is an independent check of csrc/rans.cpp, which is organised
 * differently.  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may
 * load this file's library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RANS64_L (1ull << 31)
#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)

typedef struct { uint16_t start, range; uint8_t bypass; } sym_t;

typedef struct { sym_t* v; size_t n, cap; } stack_t_;

static int push(stack_t_* s, uint16_t start, uint16_t range, uint8_t bypass) {
  if (s->n == s->cap) {
    size_t nc = s->cap ? s->cap * 2 : 1024;
    sym_t* nv = (sym_t*)realloc(s->v, nc * sizeof(sym_t));
    if (!nv) return -1;
    s->v = nv; s->cap = nc;
  }
  s->v[s->n].start = start; s->v[s->n].range = range; s->v[s->n].bypass = bypass;
  s->n++;
  return 0;
}

/* ryg_rans Rans64EncPut */
static void enc_put(uint64_t* r, uint32_t** pptr, uint32_t start, uint32_t freq, uint32_t scale_bits) {
  uint64_t x = *r;
  uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
  if (x >= x_max) { *pptr -= 1; **pptr = (uint32_t)x; x >>= 32; }
  *r = ((x / freq) << scale_bits) + (x % freq) + start;
}
/* rans_interface.cpp Rans64EncPutBits */
static void enc_put_bits(uint64_t* r, uint32_t** pptr, uint32_t val, uint32_t nbits) {
  uint64_t x = *r;
  uint32_t freq = 1u << (16 - nbits);
  uint64_t x_max = ((RANS64_L >> 16) << 32) * freq;
  if (x >= x_max) { *pptr -= 1; **pptr = (uint32_t)x; x >>= 32; }
  *r = (x << nbits) | val;
}

/* returns number of bytes written to out (<= cap) or -1 on error / -2 if cap too small (needed in *need) */
long long oracle_rans_encode(const int32_t* symbols, const int32_t* indexes, long long n, const int32_t* cdfs,
                             int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets,
                             uint8_t* out, long long cap, long long* need) {
  stack_t_ st = {0, 0, 0};
  for (long long i = 0; i < n; ++i) {
    int32_t cdf_idx = indexes[i];
    if (cdf_idx < 0 || cdf_idx >= n_cdfs) { free(st.v); return -1; }
    const int32_t* cdf = cdfs + (long long)cdf_idx * cdf_stride;
    int32_t max_value = cdf_sizes[cdf_idx] - 2;
    int32_t value = symbols[i] - offsets[cdf_idx];
    uint32_t raw_val = 0;
    if (value < 0) { raw_val = -2 * value - 1; value = max_value; }
    else if (value >= max_value) { raw_val = 2 * (value - max_value); value = max_value; }
    if (push(&st, (uint16_t)cdf[value], (uint16_t)(cdf[value + 1] - cdf[value]), 0)) { free(st.v); return -1; }
    if (value == max_value) {
      int32_t n_bypass = 0;
      while ((raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;
      int32_t val = n_bypass;
      while (val >= MAX_BYPASS_VAL) { push(&st, MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, 1); val -= MAX_BYPASS_VAL; }
      push(&st, (uint16_t)val, (uint16_t)(val + 1), 1);
      for (int32_t j = 0; j < n_bypass; ++j) {
        int32_t v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL;
        push(&st, (uint16_t)v, (uint16_t)(v + 1), 1);
      }
    }
  }
  /* flush(): pop LIFO into the tail of a word buffer */
  size_t words = st.n + 2;
  uint32_t* buf = (uint32_t*)malloc(words * sizeof(uint32_t));
  if (!buf) { free(st.v); return -1; }
  uint32_t* ptr = buf + words;
  uint64_t rans = RANS64_L;
  while (st.n) {
    sym_t s = st.v[--st.n];
    if (!s.bypass) enc_put(&rans, &ptr, s.start, s.range, PRECISION);
    else enc_put_bits(&rans, &ptr, s.start, BYPASS_PRECISION);
  }
  ptr -= 2; ptr[0] = (uint32_t)(rans >> 0); ptr[1] = (uint32_t)(rans >> 32);
  long long nbytes = (long long)((buf + words) - ptr) * 4;
  if (need) *need = nbytes;
  long long rc = nbytes;
  if (nbytes > cap || !out) rc = -2; else memcpy(out, ptr, (size_t)nbytes);
  free(buf); free(st.v);
  return rc;
}

static uint32_t dec_get_bits(uint64_t* r, const uint32_t** pptr, uint32_t n_bits) {
  uint64_t x = *r;
  uint32_t val = (uint32_t)(x & ((1u << n_bits) - 1));
  x = x >> n_bits;
  if (x < RANS64_L) { x = (x << 32) | **pptr; *pptr += 1; }
  *r = x;
  return val;
}

int oracle_rans_decode(const uint8_t* in, long long in_len, const int32_t* indexes, long long n, const int32_t* cdfs,
                       int n_cdfs, int cdf_stride, const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out) {
  if (in_len < 8) return -1;
  /* copy with slack so a truncated stream cannot read out of bounds */
  size_t words = (size_t)(in_len / 4) + 4;
  uint32_t* buf = (uint32_t*)calloc(words, 4);
  if (!buf) return -1;
  memcpy(buf, in, (size_t)in_len);
  const uint32_t* ptr = buf;
  const uint32_t* end = buf + words - 2;
  uint64_t rans = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32);
  ptr += 2;
  for (long long i = 0; i < n; ++i) {
    int32_t cdf_idx = indexes[i];
    if (cdf_idx < 0 || cdf_idx >= n_cdfs || ptr > end) { free(buf); return -1; }
    const int32_t* cdf = cdfs + (long long)cdf_idx * cdf_stride;
    int32_t max_value = cdf_sizes[cdf_idx] - 2;
    int32_t offset = offsets[cdf_idx];
    uint32_t cum_freq = (uint32_t)(rans & ((1u << PRECISION) - 1));
    int32_t k = 0;
    while (k < cdf_sizes[cdf_idx] && !((uint32_t)cdf[k] > cum_freq)) ++k;  /* find_if(v > cum) */
    uint32_t s = (uint32_t)(k - 1);
    /* Rans64DecAdvance */
    {
      uint64_t mask = (1ull << PRECISION) - 1;
      uint64_t x = rans;
      x = (uint64_t)(uint32_t)(cdf[s + 1] - cdf[s]) * (x >> PRECISION) + (x & mask) - (uint32_t)cdf[s];
      if (x < RANS64_L) { x = (x << 32) | *ptr; ptr += 1; }
      rans = x;
    }
    int32_t value = (int32_t)s;
    if (value == max_value) {
      int32_t val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
      int32_t n_bypass = val;
      while (val == MAX_BYPASS_VAL) {
        if (ptr > end) { free(buf); return -1; }
        val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
        n_bypass += val;
      }
      int32_t raw_val = 0;
      for (int j = 0; j < n_bypass; ++j) {
        if (ptr > end) { free(buf); return -1; }
        val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
        raw_val |= val << (j * BYPASS_PRECISION);
      }
      value = raw_val >> 1;
      if (raw_val & 1) value = -value - 1; else value += max_value;
    }
    out[i] = value + offset;
  }
  free(buf);
  return 0;
}

/* ops.cpp pmf_to_quantized_cdf; out has n+1 entries. returns 0, or -1 on invalid pmf */
int oracle_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf) {
  for (int i = 0; i < n; ++i) if (pmf[i] < 0 || !isfinite(pmf[i])) return -1;
  int m = n + 1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (1 << precision));
  uint32_t total = 0;
  for (int i = 0; i < m; ++i) total += cdf[i];
  if (total == 0) return -1;
  for (int i = 0; i < m; ++i) cdf[i] = (uint32_t)((((uint64_t)(1 << precision)) * cdf[i]) / total);
  for (int i = 1; i < m; ++i) cdf[i] += cdf[i - 1];
  cdf[m - 1] = 1u << precision;
  for (int i = 0; i < m - 1; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      uint32_t best_freq = ~0u;
      int best_steal = -1;
      for (int j = 0; j < m - 1; ++j) {
        uint32_t freq = cdf[j + 1] - cdf[j];
        if (freq > 1 && freq < best_freq) { best_freq = freq; best_steal = j; }
      }
      if (best_steal == -1) return -1;
      if (best_steal < i) { for (int j = best_steal + 1; j <= i; ++j) cdf[j]--; }
      else { for (int j = i + 1; j <= best_steal; ++j) cdf[j]++; }
    }
  }
  return 0;
}
