// sb_oracle.cpp -- CPU restatement of strawboat's page encode/decode path.
//
// TEST INFRASTRUCTURE ONLY (see sb_oracle.h).  Every function cites the reference
// file:line it restates (paths relative to the reference repo root).  Nothing here is
// copied: the reference is Rust on arrow2; this is a from-scratch C++17 restatement of
// the same byte-level algorithm, written to be line-traceable rather than fast.
//
// Deterministic stand-ins for the reference's non-deterministic inputs:
//   * rand::thread_rng() sample positions (src/compression/integer/mod.rs:316,332)
//       -> sbo_sample_draw(seed, codec, sample_i, range_end)
//   * std HashMap iteration order when picking Freq's top value
//     (src/compression/integer/freq.rs:50-55) -> max count, ties broken by earliest
//     first occurrence.
//   * debug-build env switches (src/util/env.rs) -> sbo_opts.force_codec.
#include "sb_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

// system liblz4 / libzstd (no headers in the image; prototypes declared by hand).
extern "C" {
int LZ4_compress_default(const char *src, char *dst, int srcSize, int dstCapacity);
int LZ4_decompress_safe(const char *src, char *dst, int compressedSize, int dstCapacity);
int LZ4_compressBound(int inputSize);
size_t ZSTD_compress(void *dst, size_t dstCapacity, const void *src, size_t srcSize, int level);
size_t ZSTD_decompress(void *dst, size_t dstCapacity, const void *src, size_t compressedSize);
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
}

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

using Bytes = std::vector<uint8_t>;

template <class T> inline T load_le(const uint8_t *p) {
  T v;
  std::memcpy(&v, p, sizeof(T));
  return v;
}
template <class T> inline void put_le(Bytes &b, T v) {
  const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
  b.insert(b.end(), p, p + sizeof(T));
}
inline void put_bytes(Bytes &b, const void *p, size_t n) {
  const uint8_t *q = static_cast<const uint8_t *>(p);
  b.insert(b.end(), q, q + n);
}
// The reference's decoders are generic over T and push whole values (`output.push(value)` after a reserve);
// the restatement is byte oriented, so the per-value loops dispatch once on the value width and then move
// W-byte values with constant-size copies (what the monomorphised Rust loop compiles to).
template <class F> inline int by_width(int W, F &&f) {
  switch (W) {
  case 1: return f(std::integral_constant<size_t, 1>{});
  case 2: return f(std::integral_constant<size_t, 2>{});
  case 4: return f(std::integral_constant<size_t, 4>{});
  case 8: return f(std::integral_constant<size_t, 8>{});
  case 16: return f(std::integral_constant<size_t, 16>{});
  default: return f(std::integral_constant<size_t, 32>{});
  }
}

struct BitView { // Option<&Bitmap> + slice offset
  const uint8_t *p = nullptr;
  int64_t off = 0;
  bool present() const { return p != nullptr; }
  bool get(int64_t i) const { // src/compression/mod.rs:111-117 is_valid
    if (!p) return true;
    int64_t k = off + i;
    return (p[k >> 3] >> (k & 7)) & 1;
  }
};

struct MutableBitmap {
  Bytes bytes;
  size_t len = 0;
  void push(bool v) {
    if ((len & 7) == 0) bytes.push_back(0);
    if (v) bytes[len >> 3] |= uint8_t(1u << (len & 7));
    ++len;
  }
  void extend_constant(size_t n, bool v) {
    for (size_t i = 0; i < n; ++i) push(v);
  }
};

inline uint32_t get_bits_needed(uint64_t x) { // src/compression/mod.rs:119-122
  return x == 0 ? 0u : 64u - uint32_t(__builtin_clzll(x));
}

// ---------------------------------------------------------------------------------
// ULEB128 + hybrid RLE (parquet2 0.17 encoding::{uleb128,hybrid_rle})
// ---------------------------------------------------------------------------------
void uleb_encode(uint64_t v, Bytes &out) {
  do {
    uint8_t b = v & 0x7f;
    v >>= 7;
    if (v) b |= 0x80;
    out.push_back(b);
  } while (v);
}
// returns bytes consumed, 0 on truncation
size_t uleb_decode(const uint8_t *p, size_t len, uint64_t &v) {
  v = 0;
  int shift = 0;
  for (size_t i = 0; i < len && i < 10; ++i) {
    v |= uint64_t(p[i] & 0x7f) << shift;
    if (!(p[i] & 0x80)) return i + 1;
    shift += 7;
  }
  return 0;
}

// parquet2 HybridRleDecoder::try_new(data, num_bits, num_values) iterated to the end
// (src/read/read_basic.rs:83-84).  Accepts bit-packed and RLE runs; a bit-packed run's
// byte length is min(groups*num_bits, remaining) (short tail tolerated).
int hybrid_rle_decode(const uint8_t *in, size_t len, uint32_t w, size_t n, std::vector<uint32_t> &out) {
  out.clear();
  out.reserve(n);
  if (w == 0) {
    out.assign(n, 0);
    return SBO_OK;
  }
  size_t pos = 0;
  while (out.size() < n) {
    if (pos >= len) return fail(SBO_OUT_OF_SPEC, "hybrid rle: stream exhausted");
    uint64_t header;
    size_t used = uleb_decode(in + pos, len - pos, header);
    if (!used) return fail(SBO_OUT_OF_SPEC, "hybrid rle: bad uleb");
    pos += used;
    if (header & 1) {
      size_t bytes = size_t(header >> 1) * w;
      bytes = std::min(bytes, len - pos);
      size_t avail = bytes * 8 / w;
      size_t take = std::min(avail, n - out.size());
      for (size_t i = 0; i < take; ++i) {
        size_t bit = i * w;
        uint64_t acc = 0;
        for (uint32_t b = 0; b < 8 && (bit >> 3) + b < bytes; ++b) acc |= uint64_t(in[pos + (bit >> 3) + b]) << (8 * b);
        out.push_back(uint32_t((acc >> (bit & 7)) & ((w >= 32) ? 0xffffffffull : ((1ull << w) - 1))));
      }
      pos += bytes;
      if (take == 0) return fail(SBO_OUT_OF_SPEC, "hybrid rle: empty bitpacked run");
    } else {
      size_t run = size_t(header >> 1);
      size_t vb = (w + 7) / 8;
      if (pos + vb > len) return fail(SBO_OUT_OF_SPEC, "hybrid rle: truncated rle value");
      uint32_t v = 0;
      for (size_t b = 0; b < vb; ++b) v |= uint32_t(in[pos + b]) << (8 * b);
      pos += vb;
      size_t take = std::min(run, n - out.size());
      out.insert(out.end(), take, v);
      if (run == 0) return fail(SBO_OUT_OF_SPEC, "hybrid rle: zero-length run");
    }
  }
  return SBO_OK;
}

// arrow2 write_def_levels(V2) -> parquet2 encode_bool: one bit-packed run
// (call site src/write/serialize.rs:209).  `[ULEB((ceil8(n)<<1)|1)][ceil8(n) bytes]`.
void encode_bool_levels(const BitView &validity, int64_t n, Bytes &out) {
  uint64_t header = (uint64_t((n + 7) / 8) << 1) | 1;
  uleb_encode(header, out);
  size_t start = out.size();
  out.resize(start + size_t((n + 7) / 8), 0);
  for (int64_t i = 0; i < n; ++i)
    if (validity.get(i)) out[start + (i >> 3)] |= uint8_t(1u << (i & 7));
}

// arrow2 write_rep_and_def(V2) -> parquet2 encode_u32 (call site serialize.rs:225):
// one bit-packed run, values LSB-first at width w, zero padded to ceil8(n)*w bytes.
void encode_u32_levels(const uint32_t *lv, size_t n, uint32_t w, Bytes &out) {
  uint64_t header = (uint64_t((n + 7) / 8) << 1) | 1;
  uleb_encode(header, out);
  size_t start = out.size();
  out.resize(start + ((n + 7) / 8) * w, 0);
  for (size_t i = 0; i < n; ++i) {
    size_t bit = i * w;
    for (uint32_t b = 0; b < w; ++b)
      if ((lv[i] >> b) & 1) out[start + ((bit + b) >> 3)] |= uint8_t(1u << ((bit + b) & 7));
  }
}

// ---------------------------------------------------------------------------------
// LZ4 block format decoder (own restatement of the public block spec).
// Cross-checked against liblz4 and pyarrow's lz4_raw in tests/test_oracle_thirdparty.py.
// ---------------------------------------------------------------------------------
int lz4_block_decode(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
  size_t ip = 0, op = 0;
  if (in_len == 0) return out_len == 0 ? SBO_OK : fail(SBO_EXTERNAL, "lz4: empty input");
  for (;;) {
    if (ip >= in_len) return fail(SBO_EXTERNAL, "lz4: truncated token");
    uint8_t token = in[ip++];
    size_t lit = token >> 4;
    if (lit == 15) {
      uint8_t b;
      do {
        if (ip >= in_len) return fail(SBO_EXTERNAL, "lz4: truncated literal length");
        b = in[ip++];
        lit += b;
      } while (b == 255);
    }
    if (ip + lit > in_len || op + lit > out_len) return fail(SBO_EXTERNAL, "lz4: literal overrun");
    std::memcpy(out + op, in + ip, lit);
    ip += lit;
    op += lit;
    if (ip == in_len) break; // last sequence: literals only
    if (ip + 2 > in_len) return fail(SBO_EXTERNAL, "lz4: truncated offset");
    size_t offset = size_t(in[ip]) | (size_t(in[ip + 1]) << 8);
    ip += 2;
    if (offset == 0 || offset > op) return fail(SBO_EXTERNAL, "lz4: bad offset");
    size_t ml = token & 15;
    if (ml == 15) {
      uint8_t b;
      do {
        if (ip >= in_len) return fail(SBO_EXTERNAL, "lz4: truncated match length");
        b = in[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (op + ml > out_len) return fail(SBO_EXTERNAL, "lz4: match overrun");
    for (size_t i = 0; i < ml; ++i) out[op + i] = out[op + i - offset]; // overlapping copy
    op += ml;
  }
  if (op != out_len) return fail(SBO_EXTERNAL, "lz4: decoded size mismatch");
  return SBO_OK;
}

// ---------------------------------------------------------------------------------
// CommonCompression (src/compression/basic.rs:62-152)
// ---------------------------------------------------------------------------------
// ---- Snappy raw format (snap = "1.1.0" raw::{Encoder,Decoder}; format_description.txt of google/snappy):
// [varint uncompressed length] then elements: tag & 3 == 0 literal (len-1 in the upper 6 bits, 60..63 => 1..4
// little-endian length bytes follow), 1 = copy with 11-bit offset and length 4..11, 2 = copy with 16-bit offset
// and length 1..64, 3 = copy with 32-bit offset.  The crate source is absent (FORMAT_ASSUMPTIONS #5): pinned
// against pyarrow's snappy codec (Google's C++ library) in tests/test_oracle_thirdparty.py.
int snappy_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
  size_t ip = 0;
  uint64_t n = 0;
  for (unsigned shift = 0;; shift += 7) {
    if (ip >= in_len || shift > 28) return fail(SBO_EXTERNAL, "decompress snappy faild: header");
    uint8_t b = in[ip++];
    n |= uint64_t(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
  }
  if (n != out_len) return fail(SBO_EXTERNAL, "decompress snappy faild: length mismatch"); // snap: BufferTooSmall when larger
  size_t op = 0;
  while (ip < in_len) {
    const uint8_t tag = in[ip++];
    size_t len, offset;
    switch (tag & 3) {
    case 0: {
      len = size_t(tag >> 2) + 1;
      if (len > 60) {
        const size_t nb = len - 60;
        if (in_len - ip < nb) return fail(SBO_EXTERNAL, "decompress snappy faild: literal length");
        len = 0;
        for (size_t k = 0; k < nb; ++k) len |= size_t(in[ip + k]) << (8 * k);
        len += 1;
        ip += nb;
      }
      if (len > in_len - ip || len > out_len - op) return fail(SBO_EXTERNAL, "decompress snappy faild: literal");
      std::memcpy(out + op, in + ip, len);
      ip += len;
      op += len;
      continue;
    }
    case 1:
      if (ip >= in_len) return fail(SBO_EXTERNAL, "decompress snappy faild: copy1");
      len = 4 + ((tag >> 2) & 7);
      offset = (size_t(tag >> 5) << 8) | in[ip++];
      break;
    case 2:
      if (in_len - ip < 2) return fail(SBO_EXTERNAL, "decompress snappy faild: copy2");
      len = size_t(tag >> 2) + 1;
      offset = size_t(in[ip]) | (size_t(in[ip + 1]) << 8);
      ip += 2;
      break;
    default:
      if (in_len - ip < 4) return fail(SBO_EXTERNAL, "decompress snappy faild: copy4");
      len = size_t(tag >> 2) + 1;
      offset = size_t(load_le<uint32_t>(in + ip));
      ip += 4;
      break;
    }
    if (offset == 0 || offset > op || len > out_len - op) return fail(SBO_EXTERNAL, "decompress snappy faild: copy");
    for (size_t k = 0; k < len; ++k) out[op + k] = out[op - offset + k]; // byte order matters: copies may overlap
    op += len;
  }
  if (op != out_len) return fail(SBO_EXTERNAL, "decompress snappy faild: short stream");
  return SBO_OK;
}
// greedy 4-byte-hash matcher; any valid stream decodes with the reference's snap::raw::Decoder
void snappy_compress(const uint8_t *in, size_t n, Bytes &out) {
  for (uint64_t v = n;;) {
    uint8_t b = v & 0x7f;
    v >>= 7;
    out.push_back(b | (v ? 0x80 : 0));
    if (!v) break;
  }
  auto emit_literal = [&](size_t from, size_t len) {
    if (!len) return;
    const size_t l1 = len - 1;
    if (l1 < 60) out.push_back(uint8_t(l1 << 2));
    else {
      int nb = l1 < (1u << 8) ? 1 : l1 < (1u << 16) ? 2 : l1 < (1u << 24) ? 3 : 4;
      out.push_back(uint8_t((59 + nb) << 2));
      for (int k = 0; k < nb; ++k) out.push_back(uint8_t(l1 >> (8 * k)));
    }
    put_bytes(out, in + from, len);
  };
  auto emit_copy = [&](size_t offset, size_t len) {
    while (len) {
      size_t l = len > 64 ? (len - 64 < 4 ? 60 : 64) : len; // never leave a tail shorter than 4 (it fits copy2 anyway)
      if (l >= 4 && l <= 11 && offset < 2048) {
        out.push_back(uint8_t(1 | ((l - 4) << 2) | ((offset >> 8) << 5)));
        out.push_back(uint8_t(offset));
      } else {
        out.push_back(uint8_t(2 | ((l - 1) << 2)));
        out.push_back(uint8_t(offset));
        out.push_back(uint8_t(offset >> 8));
      }
      len -= l;
    }
  };
  std::vector<int64_t> table(1 << 14, -1);
  size_t ip = 0, anchor = 0;
  while (n >= 8 && ip + 4 <= n) {
    const uint32_t seq = load_le<uint32_t>(in + ip);
    const uint32_t h = (seq * 0x1e35a7bdu) >> 18;
    const int64_t cand = table[h];
    table[h] = int64_t(ip);
    if (cand >= 0 && ip - size_t(cand) <= 65535 && load_le<uint32_t>(in + cand) == seq) {
      size_t ml = 4;
      while (ip + ml < n && in[cand + ml] == in[ip + ml]) ++ml;
      emit_literal(anchor, ip - anchor);
      emit_copy(ip - size_t(cand), ml);
      ip += ml;
      anchor = ip;
    } else {
      ++ip;
    }
  }
  emit_literal(anchor, n - anchor);
}

int common_decompress(int codec, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
  switch (codec) {
  case SBO_C_NONE: // basic.rs:67-70 copy_from_slice panics on length mismatch
    if (in_len != out_len) return fail(SBO_PANIC, "None: copy_from_slice length mismatch");
    std::memcpy(out, in, in_len);
    return SBO_OK;
  case SBO_C_LZ4: { // basic.rs:87-91 LZ4_decompress_safe with known output size
    if (in_len == 0 && out_len == 0) return SBO_OK;
    int r = LZ4_decompress_safe(reinterpret_cast<const char *>(in), reinterpret_cast<char *>(out), int(in_len),
                                int(out_len));
    if (r < 0) return fail(SBO_EXTERNAL, "lz4 decompress failed");
    return SBO_OK;
  }
  case SBO_C_ZSTD: { // basic.rs:93-97
    size_t r = ZSTD_decompress(out, out_len, in, in_len);
    if (ZSTD_isError(r)) return fail(SBO_EXTERNAL, "zstd decompress failed");
    return SBO_OK;
  }
  default:
    return snappy_decompress(in, in_len, out, out_len); // basic.rs:98-105
  }
}

int common_compress(int codec, const uint8_t *in, size_t in_len, Bytes &out, size_t &written) {
  switch (codec) {
  case SBO_C_NONE: // basic.rs:79-82
    put_bytes(out, in, in_len);
    written = in_len;
    return SBO_OK;
  case SBO_C_LZ4: { // basic.rs:107-120 (LZ4_compress_default, no size prefix)
    int bound = LZ4_compressBound(int(in_len));
    size_t start = out.size();
    out.resize(start + size_t(bound));
    int r = LZ4_compress_default(reinterpret_cast<const char *>(in), reinterpret_cast<char *>(out.data() + start),
                                 int(in_len), bound);
    if (r <= 0 && in_len != 0) return fail(SBO_EXTERNAL, "Compress lz4 faild");
    out.resize(start + size_t(r));
    written = size_t(r);
    return SBO_OK;
  }
  case SBO_C_ZSTD: { // basic.rs:122-136 level 0 (= default 3)
    size_t bound = ZSTD_compressBound(in_len);
    size_t start = out.size();
    out.resize(start + bound);
    size_t r = ZSTD_compress(out.data() + start, bound, in, in_len, 0);
    if (ZSTD_isError(r)) return fail(SBO_EXTERNAL, "Compress zstd faild");
    out.resize(start + r);
    written = r;
    return SBO_OK;
  }
  case SBO_C_SNAPPY: { // basic.rs:138-152
    const size_t start = out.size();
    snappy_compress(in, in_len, out);
    written = out.size() - start;
    return SBO_OK;
  }
  default: return fail(SBO_OUT_OF_SPEC, "not a common codec");
  }
}

// ---------------------------------------------------------------------------------
// bitpacking 0.8 BitPacker4x (crate source absent -> parity unpinned)
// Block = 128 u32.  Value i sits in lane l = i%4 at position k = i/4 of that lane's
// LSB-first bit stream; word w of lane l is stored at output u32 index 4*w + l.
// ---------------------------------------------------------------------------------
uint32_t bp4x_num_bits(const uint32_t *v) {
  uint32_t acc = 0;
  for (int i = 0; i < 128; ++i) acc |= v[i];
  return acc == 0 ? 0u : 32u - uint32_t(__builtin_clz(acc));
}
size_t bp4x_compress(const uint32_t *v, uint8_t *out, uint32_t b) {
  if (b == 0) return 0;
  std::vector<uint32_t> words(4 * b, 0);
  for (int i = 0; i < 128; ++i) {
    uint32_t lane = i & 3, k = i >> 2;
    uint64_t val = (b == 32) ? v[i] : (v[i] & ((1u << b) - 1));
    uint32_t bit = k * b, w = bit >> 5, s = bit & 31;
    words[4 * w + lane] |= uint32_t(val << s);
    if (s + b > 32) words[4 * (w + 1) + lane] |= uint32_t(val >> (32 - s));
  }
  std::memcpy(out, words.data(), 16 * b);
  return 16 * b;
}
size_t bp4x_decompress(const uint8_t *in, uint32_t *v, uint32_t b) {
  if (b == 0) {
    std::fill(v, v + 128, 0u);
    return 0;
  }
  std::vector<uint32_t> words(4 * b);
  std::memcpy(words.data(), in, 16 * b);
  for (int i = 0; i < 128; ++i) {
    uint32_t lane = i & 3, k = i >> 2;
    uint32_t bit = k * b, w = bit >> 5, s = bit & 31;
    uint64_t acc = words[4 * w + lane];
    if (s + b > 32) acc |= uint64_t(words[4 * (w + 1) + lane]) << 32;
    v[i] = uint32_t((acc >> s) & ((b == 32) ? 0xffffffffull : ((1ull << b) - 1)));
  }
  return 16 * b;
}
// compress_sorted: wrapping deltas d[i] = v[i] - v[i-1], v[-1] = initial (plain D1 delta)
size_t bp4x_compress_sorted(uint32_t initial, const uint32_t *v, uint8_t *out, uint32_t b) {
  uint32_t d[128];
  uint32_t prev = initial;
  for (int i = 0; i < 128; ++i) {
    d[i] = v[i] - prev;
    prev = v[i];
  }
  return bp4x_compress(d, out, b);
}
size_t bp4x_decompress_sorted(uint32_t initial, const uint8_t *in, uint32_t *v, uint32_t b) {
  size_t used = bp4x_decompress(in, v, b);
  uint32_t prev = initial;
  for (int i = 0; i < 128; ++i) {
    v[i] += prev;
    prev = v[i];
  }
  return used;
}

// ---------------------------------------------------------------------------------
// roaring 0.10.1 portable serialization (crate source absent -> unpinned)
// ---------------------------------------------------------------------------------
void roaring_serialize(const std::vector<uint32_t> &vals /* sorted, unique */, Bytes &out) {
  struct C {
    uint16_t key;
    size_t lo, hi;
  };
  std::vector<C> cs;
  for (size_t i = 0; i < vals.size();) {
    uint16_t key = uint16_t(vals[i] >> 16);
    size_t j = i;
    while (j < vals.size() && uint16_t(vals[j] >> 16) == key) ++j;
    cs.push_back({key, i, j});
    i = j;
  }
  put_le<uint32_t>(out, 12346u); // SERIAL_COOKIE_NO_RUNCONTAINER
  put_le<uint32_t>(out, uint32_t(cs.size()));
  for (auto &c : cs) {
    put_le<uint16_t>(out, c.key);
    put_le<uint16_t>(out, uint16_t(c.hi - c.lo - 1));
  }
  uint32_t offset = uint32_t(cs.size()) * 8 + 8;
  for (auto &c : cs) {
    put_le<uint32_t>(out, offset);
    size_t card = c.hi - c.lo;
    offset += (card > 4096) ? 8192u : uint32_t(card * 2);
  }
  for (auto &c : cs) {
    size_t card = c.hi - c.lo;
    if (card > 4096) {
      uint64_t bits[1024] = {0};
      for (size_t i = c.lo; i < c.hi; ++i) {
        uint32_t lo = vals[i] & 0xffff;
        bits[lo >> 6] |= 1ull << (lo & 63);
      }
      put_bytes(out, bits, sizeof(bits));
    } else {
      for (size_t i = c.lo; i < c.hi; ++i) put_le<uint16_t>(out, uint16_t(vals[i] & 0xffff));
    }
  }
}
size_t roaring_serialized_size(const std::vector<uint32_t> &vals) {
  Bytes tmp;
  roaring_serialize(vals, tmp);
  return tmp.size();
}
int roaring_deserialize(const uint8_t *in, size_t len, std::vector<uint32_t> &vals) {
  vals.clear();
  size_t pos = 0;
  auto need = [&](size_t n) { return pos + n <= len; };
  if (!need(4)) return fail(SBO_IO, "roaring: truncated cookie");
  uint32_t cookie = load_le<uint32_t>(in);
  pos = 4;
  size_t size;
  bool has_run = false;
  const uint8_t *run_bitmap = nullptr;
  if (cookie == 12346u) {
    if (!need(4)) return fail(SBO_IO, "roaring: truncated size");
    size = load_le<uint32_t>(in + pos);
    pos += 4;
  } else if ((cookie & 0xffff) == 12347u) {
    size = size_t(cookie >> 16) + 1;
    has_run = true;
    size_t rb = (size + 7) / 8;
    if (!need(rb)) return fail(SBO_IO, "roaring: truncated run bitmap");
    run_bitmap = in + pos;
    pos += rb;
  } else {
    return fail(SBO_IO, "roaring: unknown cookie");
  }
  if (size > 65536) return fail(SBO_IO, "roaring: size too large");
  if (!need(size * 4)) return fail(SBO_IO, "roaring: truncated descriptions");
  const uint8_t *desc = in + pos;
  pos += size * 4;
  bool has_offsets = !has_run || size >= 4;
  if (has_offsets) {
    if (!need(size * 4)) return fail(SBO_IO, "roaring: truncated offsets");
    pos += size * 4;
  }
  for (size_t c = 0; c < size; ++c) {
    uint32_t key = load_le<uint16_t>(desc + 4 * c);
    size_t card = size_t(load_le<uint16_t>(desc + 4 * c + 2)) + 1;
    bool is_run = has_run && ((run_bitmap[c >> 3] >> (c & 7)) & 1);
    if (is_run) {
      if (!need(2)) return fail(SBO_IO, "roaring: truncated run count");
      size_t nruns = load_le<uint16_t>(in + pos);
      pos += 2;
      if (!need(nruns * 4)) return fail(SBO_IO, "roaring: truncated runs");
      for (size_t r = 0; r < nruns; ++r) {
        uint32_t s = load_le<uint16_t>(in + pos + 4 * r), l = load_le<uint16_t>(in + pos + 4 * r + 2);
        for (uint32_t v = s; v <= s + l; ++v) vals.push_back((key << 16) | v);
      }
      pos += nruns * 4;
    } else if (card <= 4096) {
      if (!need(card * 2)) return fail(SBO_IO, "roaring: truncated array container");
      for (size_t i = 0; i < card; ++i) vals.push_back((key << 16) | load_le<uint16_t>(in + pos + 2 * i));
      pos += card * 2;
    } else {
      if (!need(8192)) return fail(SBO_IO, "roaring: truncated bitmap container");
      for (uint32_t w = 0; w < 1024; ++w) {
        uint64_t bits = load_le<uint64_t>(in + pos + 8 * w);
        while (bits) {
          int b = __builtin_ctzll(bits);
          vals.push_back((key << 16) | (w * 64 + uint32_t(b)));
          bits &= bits - 1;
        }
      }
      pos += 8192;
    }
  }
  return SBO_OK;
}

// ---------------------------------------------------------------------------------
// sampler stand-in (shared with the GPU chooser; see sb_oracle.h)
// ---------------------------------------------------------------------------------
uint64_t sample_draw(uint64_t seed, uint32_t codec, uint32_t sample_i, uint64_t range_end) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t(codec) * 16 + sample_i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return range_end ? z % range_end : 0;
}

inline bool forbidden(const sbo_opts &o, int codec) { return (o.forbidden_mask >> codec) & 1u; }
inline sbo_opts forbid(const sbo_opts &o, int codec) {
  sbo_opts r = o;
  r.forbidden_mask |= 1u << codec;
  return r;
}
inline sbo_opts default_opts_like(const sbo_opts &o) { // WriteOptions::default() + our stand-ins
  sbo_opts r{};
  r.default_compression = SBO_C_NONE;
  r.default_compress_ratio = -1.0;
  r.forbidden_mask = 0;
  r.force_codec = -1;
  r.seed = o.seed;
  r.float_bitwise = o.float_bitwise;
  return r;
}

// ---------------------------------------------------------------------------------
// Value traits: integers compare natively; floats are carried as bit patterns and compare
// as ordered_float::OrderedFloat (src/compression/double/traits.rs:51-53) unless
// float_bitwise is set (SURVEY App. C6 deviation used by the GPU encoder).
// ---------------------------------------------------------------------------------
// i128 / i256 (arrow2 `i128`, `i256`: src/util/mod.rs:77-78, src/compression/integer/traits.rs:28-39).  Little-endian
// two's complement; ordered as signed integers; as_i64 keeps the low 64 bits.
using i128 = __int128;
struct i256 {
  unsigned __int128 lo = 0;
  __int128 hi = 0;
  bool operator==(const i256 &o) const { return lo == o.lo && hi == o.hi; }
  bool operator!=(const i256 &o) const { return !(*this == o); }
  bool operator<(const i256 &o) const { return hi != o.hi ? hi < o.hi : lo < o.lo; }
  explicit operator int64_t() const { return int64_t(lo); }
};
static_assert(sizeof(i128) == 16 && sizeof(i256) == 32, "wide integer layout");
} // namespace
namespace std {
template <> struct hash<::i256> {
  size_t operator()(const ::i256 &v) const noexcept {
    uint64_t w[4];
    std::memcpy(w, &v, 32);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t x : w) h = (h ^ x) * 0xBF58476D1CE4E5B9ull, h ^= h >> 29;
    return size_t(h);
  }
};
} // namespace std
struct I128Hash {
  size_t operator()(const __int128 &v) const noexcept {
    uint64_t w[2];
    std::memcpy(w, &v, 16);
    uint64_t h = (w[0] ^ 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h = (h ^ (h >> 29) ^ w[1]) * 0x94D049BB133111EBull;
    return size_t(h ^ (h >> 31));
  }
};
namespace {
template <class K> struct KeyHash { using type = std::hash<K>; };
template <> struct KeyHash<__int128> { using type = I128Hash; };

template <class T> struct IntTr {
  using V = T;
  using K = T;
  static constexpr bool is_float = false;
  static K key(V v, bool) { return v; }
  static bool eq(V a, V b, bool) { return a == b; }
  static bool lt(V a, V b) { return a < b; }
  static int64_t as_i64(V v) { return int64_t(v); } // integer/traits.rs:19-39
};
template <class F, class B> struct FloatTr {
  using V = B; // bit pattern
  using K = B;
  static constexpr bool is_float = true;
  static F f(B b) {
    F x;
    std::memcpy(&x, &b, sizeof(F));
    return x;
  }
  static K key(V v, bool bitwise) {
    if (bitwise) return v;
    F x = f(v);
    if (x != x) { // canonical NaN
      F q = std::numeric_limits<F>::quiet_NaN();
      B r;
      std::memcpy(&r, &q, sizeof(F));
      return r;
    }
    if (x == F(0)) return B(0); // +0 == -0
    return v;
  }
  static bool eq(V a, V b, bool bitwise) { return key(a, bitwise) == key(b, bitwise); }
  static bool lt(V a, V b) { // OrderedFloat total order, NaN greatest
    F x = f(a), y = f(b);
    bool xn = x != x, yn = y != y;
    if (xn) return false;
    if (yn) return true;
    return x < y;
  }
  static int64_t as_i64(V) { return 0; }
};

// IntegerStats / DoubleStats (integer/mod.rs:165-229, double/mod.rs:165-229)
template <class Tr> struct Stats {
  using V = typename Tr::V;
  using K = typename Tr::K;
  const V *values = nullptr;
  BitView validity;
  size_t tuple_count = 0, total_bytes = 0, null_count = 0;
  bool is_sorted = true;
  V min{}, max{};
  struct Cnt {
    size_t count = 0, first = 0;
  };
  std::unordered_map<K, Cnt, typename KeyHash<K>::type> distinct;
  size_t unique_count = 0;
};

template <class Tr> Stats<Tr> gen_stats(const typename Tr::V *values, BitView validity, size_t n, bool bitwise) {
  using V = typename Tr::V;
  Stats<Tr> s;
  s.values = values;
  s.validity = validity;
  s.tuple_count = n;
  s.total_bytes = n * sizeof(V);
  bool init = false;
  V last{}; // T::default()
  for (size_t i = 0; i < n; ++i) {
    V cur = values[i];
    bool valid = validity.get(int64_t(i));
    if (!valid) ++s.null_count;
    if (valid) {
      if (Tr::lt(cur, last)) s.is_sorted = false;
      if (!Tr::eq(last, cur, bitwise)) last = cur; // run_count only feeds average_run_length (unused)
    }
    auto &c = s.distinct[Tr::key(cur, bitwise)];
    if (c.count == 0) c.first = i;
    ++c.count;
    if (!init) {
      init = true;
      s.min = cur;
      s.max = cur;
    }
    if (Tr::lt(s.max, cur)) s.max = cur;
    else if (Tr::lt(cur, s.min)) s.min = cur;
  }
  s.unique_count = s.distinct.size();
  return s;
}

// forward
template <class Tr>
int compress_fixed(const typename Tr::V *values, BitView validity, size_t n, const sbo_opts &opts, Bytes &out);
int decompress_fixed(const uint8_t *in, size_t in_len, size_t n, int W, bool is_float, Bytes &out, size_t &consumed);

// ---- RLE (integer/rle.rs:64-134, double/rle.rs:61-135) ----------------------------
template <class Tr>
void rle_compress(const typename Tr::V *values, BitView validity, size_t n, bool bitwise, Bytes &out) {
  using V = typename Tr::V;
  uint32_t seen = 0;
  V last{};
  bool all_null = true;
  for (size_t i = 0; i < n; ++i) {
    V item = values[i];
    if (validity.get(int64_t(i))) {
      if (all_null) {
        all_null = false;
        last = item;
        ++seen;
      } else if (!Tr::eq(last, item, bitwise)) {
        put_le<uint32_t>(out, seen);
        put_le<V>(out, last);
        last = item;
        seen = 1;
      } else {
        ++seen;
      }
    } else {
      ++seen; // NULL: extend current run (rle.rs:92-95)
    }
  }
  if (seen != 0) {
    put_le<uint32_t>(out, seen);
    put_le<V>(out, last);
  }
}
int rle_decompress(const uint8_t *in, size_t in_len, size_t n, int W, Bytes &out) {
  size_t pos = 0, num = 0;
  for (;;) { // rle.rs:115-132: loop { read; push; if num_values >= length break }
    if (pos + 4 + size_t(W) > in_len) return fail(SBO_IO, "rle: failed to fill whole buffer");
    uint32_t len = load_le<uint32_t>(in + pos);
    const uint8_t *v = in + pos + 4;
    pos += 4 + size_t(W);
    const size_t at = out.size();
    out.resize(at + size_t(len) * size_t(W));
    by_width(W, [&](auto w) {
      constexpr size_t WW = decltype(w)::value;
      uint8_t *d = out.data() + at;
      for (uint32_t k = 0; k < len; ++k) std::memcpy(d + size_t(k) * WW, v, WW);
      return 0;
    });
    num += len;
    if (num >= n) break;
  }
  return SBO_OK;
}

// ---- OneValue (integer/one_value.rs:31-95) ----------------------------------------
template <class Tr> void onevalue_compress(const typename Tr::V *values, BitView validity, size_t n, Bytes &out) {
  typename Tr::V val{};
  for (size_t i = 0; i < n; ++i)
    if (validity.get(int64_t(i))) {
      val = values[i];
      break;
    }
  put_le(out, val);
}
int onevalue_decompress(const uint8_t *in, size_t in_len, size_t n, int W, Bytes &out) {
  if (in_len < size_t(W)) return fail(SBO_IO, "one_value: failed to fill whole buffer");
  const size_t at = out.size();
  out.resize(at + n * size_t(W));
  by_width(W, [&](auto w) {
    constexpr size_t WW = decltype(w)::value;
    uint8_t *d = out.data() + at;
    for (size_t i = 0; i < n; ++i) std::memcpy(d + i * WW, in, WW);
    return 0;
  });
  return SBO_OK;
}

// ---- Dict (integer/dict.rs:33-121, double/dict.rs:38-124) --------------------------
// ids in first-occurrence order (dict.rs:212-214); null => repeat previous index, or
// T::default() if it is the first row (dict.rs:46-54).  Keys are compared by raw bytes.
template <class Tr>
int dict_compress(const typename Tr::V *values, BitView validity, size_t n, const sbo_opts &opts, Bytes &out) {
  using V = typename Tr::V;
  std::unordered_map<V, uint32_t, typename KeyHash<V>::type> interner;
  std::vector<V> sets;
  std::vector<uint32_t> indices;
  indices.reserve(n);
  auto push = [&](V v) {
    auto it = interner.find(v);
    if (it == interner.end()) {
      uint32_t k = uint32_t(sets.size());
      interner.emplace(v, k);
      sets.push_back(v);
      indices.push_back(k);
    } else {
      indices.push_back(it->second);
    }
  };
  for (size_t i = 0; i < n; ++i) {
    if (validity.get(int64_t(i))) push(values[i]);
    else if (indices.empty()) push(V{});
    else indices.push_back(indices.back());
  }
  sbo_opts sub = forbid(opts, SBO_C_DICT);
  int rc = compress_fixed<IntTr<uint32_t>>(indices.data(), BitView{}, indices.size(), sub, out);
  if (rc) return rc;
  put_le<uint32_t>(out, uint32_t(sets.size()));
  for (V v : sets) put_le<V>(out, v);
  return SBO_OK;
}
int dict_decompress(const uint8_t *in, size_t in_len, size_t n, int W, Bytes &out) {
  Bytes idx_bytes;
  size_t used = 0;
  int rc = decompress_fixed(in, in_len, n, 4, false, idx_bytes, used);
  if (rc) return rc;
  in += used;
  in_len -= used;
  if (in_len < 4) return fail(SBO_IO, "dict: failed to fill whole buffer");
  size_t data_size = size_t(load_le<uint32_t>(in)) * size_t(W);
  in += 4;
  in_len -= 4;
  if (in_len < data_size) return fail(SBO_OUT_OF_SPEC, "Invalid data size"); // dict.rs:80-86
  size_t k = data_size / size_t(W);
  size_t count = idx_bytes.size() / 4;
  const size_t at = out.size();
  out.resize(at + count * size_t(W));
  return by_width(W, [&](auto w) {
    constexpr size_t WW = decltype(w)::value;
    uint8_t *d = out.data() + at;
    const uint8_t *ib = idx_bytes.data();
    for (size_t i = 0; i < count; ++i) {
      uint32_t id = load_le<uint32_t>(ib + 4 * i);
      if (id >= k) {
        out.resize(at + i * WW);
        return fail(SBO_PANIC, "dict: index out of bounds"); // data[*i as usize]
      }
      std::memcpy(d + i * WW, in + size_t(id) * WW, WW);
    }
    return int(SBO_OK);
  });
}
template <class Tr> double dict_ratio(const Stats<Tr> &s) { // dict.rs:109-120
  if (s.unique_count * 3 >= s.tuple_count) return 0.0;
  size_t after = s.unique_count * sizeof(typename Tr::V) + s.tuple_count * size_t(get_bits_needed(s.unique_count) / 8);
  after += s.tuple_count * 2 / 128;
  return double(s.total_bytes) / double(after);
}

// ---- Freq (integer/freq.rs:33-152, double/freq.rs:33-152) --------------------------
template <class Tr> size_t freq_max_count(const Stats<Tr> &s, typename Tr::V *top) {
  size_t max_count = 0, best_first = 0;
  for (auto &kv : s.distinct) { // deterministic stand-in for HashMap order: ties -> earliest first occurrence
    if (kv.second.count > max_count || (kv.second.count == max_count && kv.second.first < best_first)) {
      max_count = kv.second.count;
      best_first = kv.second.first;
    }
  }
  if (top && max_count) *top = s.values[best_first];
  return max_count;
}
template <class Tr> double freq_ratio(const Stats<Tr> &s) { // freq.rs:129-151
  if (s.unique_count <= 1) return 0.0;
  if (double(s.null_count) / double(s.tuple_count) >= 0.9) return double(s.tuple_count - 1);
  size_t max_count = freq_max_count<Tr>(s, nullptr);
  bool big = double(max_count) / double(s.tuple_count) >= 0.9;
  if (!Tr::is_float) big = big && Tr::as_i64(s.max) >= (1 << 8); // integers only (freq.rs:146)
  return big ? double(s.tuple_count - 1) : 0.0;
}
template <class Tr> int freq_compress(const Stats<Tr> &s, const sbo_opts &opts, Bytes &out) {
  using V = typename Tr::V;
  bool top_is_null = false;
  V top{};
  if (double(s.null_count) / double(s.tuple_count) >= 0.9) top_is_null = true; // freq.rs:47-48
  else freq_max_count<Tr>(s, &top);
  std::vector<uint32_t> rows;
  std::vector<V> exceptions;
  for (size_t i = 0; i < s.tuple_count; ++i) {
    if (!s.validity.get(int64_t(i))) continue;
    if (top_is_null || !Tr::eq(s.values[i], top, opts.float_bitwise)) {
      rows.push_back(uint32_t(i));
      exceptions.push_back(s.values[i]);
    }
  }
  put_le<V>(out, top);
  put_le<uint32_t>(out, uint32_t(roaring_serialized_size(rows)));
  roaring_serialize(rows, out);
  sbo_opts sub = forbid(opts, SBO_C_FREQ);
  return compress_fixed<Tr>(exceptions.data(), BitView{}, exceptions.size(), sub, out);
}
int freq_decompress(const uint8_t *in, size_t in_len, size_t n, int W, bool is_float, Bytes &out) {
  if (in_len < size_t(W) + 4) return fail(SBO_IO, "freq: failed to fill whole buffer");
  size_t begin = out.size();
  out.resize(begin + n * size_t(W));
  by_width(W, [&](auto w) {
    constexpr size_t WW = decltype(w)::value;
    uint8_t *d = out.data() + begin;
    for (size_t i = 0; i < n; ++i) std::memcpy(d + i * WW, in, WW);
    return 0;
  });
  size_t bm = load_le<uint32_t>(in + W);
  in += W + 4;
  in_len -= size_t(W) + 4;
  if (bm > in_len) return fail(SBO_PANIC, "freq: bitmap slice out of range");
  std::vector<uint32_t> rows;
  int rc = roaring_deserialize(in, bm, rows);
  if (rc) return rc;
  in += bm;
  in_len -= bm;
  Bytes exc;
  size_t used = 0;
  rc = decompress_fixed(in, in_len, rows.size(), W, is_float, exc, used);
  if (rc) return rc;
  if (exc.size() != rows.size() * size_t(W)) return fail(SBO_PANIC, "freq: assert_eq exceptions len"); // freq.rs:116
  for (size_t i = 0; i < rows.size(); ++i) {
    if (rows[i] >= n) return fail(SBO_PANIC, "freq: exception row out of bounds");
    std::memcpy(out.data() + begin + size_t(rows[i]) * size_t(W), exc.data() + i * size_t(W), size_t(W));
  }
  return SBO_OK;
}

// ---- Bitpacking / DeltaBitpacking (integer/bp.rs:36-101, delta_bp.rs:36-110) -------
int bitpack_compress(const uint32_t *v, size_t n, bool delta, Bytes &out) {
  if (n % 128) return fail(SBO_PANIC, "bitpacking: block length must be 128"); // BitPacker asserts
  uint32_t initial = 0;
  for (size_t i = 0; i < n; i += 128) {
    uint32_t b = bp4x_num_bits(v + i); // width of RAW values, also for delta (delta_bp.rs:50)
    out.push_back(uint8_t(b));
    size_t start = out.size();
    out.resize(start + 16 * b);
    if (delta) bp4x_compress_sorted(initial, v + i, out.data() + start, b);
    else bp4x_compress(v + i, out.data() + start, b);
    initial = v[i + 127];
  }
  return SBO_OK;
}
int bitpack_decompress(const uint8_t *in, size_t in_len, size_t n, int W, bool delta, Bytes &out) {
  if (W != 4) return fail(SBO_PANIC, "bitpacking: only 4-byte types (bp.rs:70-79 writes u32 lanes)");
  size_t pos = 0;
  uint32_t initial = 0;
  for (size_t i = 0; i < n; i += 128) { // (0..length).step_by(128): always whole blocks
    if (pos >= in_len) return fail(SBO_IO, "bitpacking: failed to fill whole buffer");
    uint32_t b = in[pos++];
    if (b > 32 || pos + 16 * b > in_len) return fail(SBO_PANIC, "bitpacking: block out of range");
    uint32_t v[128];
    if (delta) bp4x_decompress_sorted(initial, in + pos, v, b);
    else bp4x_decompress(in + pos, v, b);
    pos += 16 * b;
    initial = v[127];
    put_bytes(out, v, sizeof(v));
  }
  return SBO_OK;
}

// ---- Patas (double/patas.rs:36-189) -------------------------------------------------
uint16_t patas_pack(uint8_t ref, uint8_t sig, uint8_t tz) { // patas.rs:141-146
  return uint16_t((uint16_t(ref) << 9) | (uint16_t(sig & 7) << 6) | uint16_t(tz));
}
void patas_unpack(uint16_t p, uint8_t &ref, uint8_t &sig, uint8_t &tz) { // patas.rs:148-161
  ref = uint8_t((p >> 9) & 0x7f);
  sig = uint8_t((p >> 6) & 7);
  tz = uint8_t(p & 0x3f);
  if (tz < 63 && sig == 0) sig = 8;
}
template <class B> void patas_compress(const B *values, size_t n, Bytes &out) {
  constexpr unsigned BITS = sizeof(B) * 8;
  std::unordered_map<B, size_t> indices; // value -> last index (patas.rs:45,103)
  for (size_t i = 0; i < n; ++i) {
    B val = values[i];
    if (i == 0) {
      put_le<B>(out, val);
    } else {
      auto it = indices.find(val);
      size_t ref = it == indices.end() ? 0 : it->second; // unwrap_or(0)  (patas.rs:59)
      if (ref > i || (i - ref) >= 128) ref = i - 1;
      size_t diff = i - ref;
      B x = val ^ values[i - diff]; // ring.get(-diff)
      unsigned tz = x == 0 ? BITS : unsigned(__builtin_ctzll(uint64_t(x)));
      unsigned lz = x == 0 ? BITS : unsigned(__builtin_clzll(uint64_t(x))) - (64 - BITS);
      unsigned is_equal = tz == BITS;
      unsigned sig_bits = is_equal ? 0 : BITS - tz - lz;
      unsigned sig_bytes = (sig_bits >> 3) + ((sig_bits & 7) != 0);
      put_le<uint16_t>(out, patas_pack(uint8_t(diff), uint8_t(sig_bytes), uint8_t(tz - is_equal)));
      B shifted = B(x >> (tz - is_equal));
      put_bytes(out, &shifted, sig_bytes); // write_value_bytes: first sig_bytes LE bytes
    }
    indices[val] = i;
  }
}
int patas_decompress(const uint8_t *in, size_t in_len, size_t n, int W, Bytes &out) {
  if (n == 0) return fail(SBO_PANIC, "patas: length - 1 underflow");
  if (in_len < size_t(W)) return fail(SBO_IO, "patas: failed to fill whole buffer");
  size_t begin = out.size();
  out.reserve(begin + n * size_t(W));
  put_bytes(out, in, size_t(W));
  size_t pos = size_t(W);
  for (size_t i = 1; i < n; ++i) {
    if (pos + 2 > in_len) return fail(SBO_IO, "patas: failed to fill whole buffer");
    uint8_t ref, sig, tz;
    patas_unpack(load_le<uint16_t>(in + pos), ref, sig, tz);
    pos += 2;
    // read_value_custom (patas.rs:164-189): copies `sig` bytes into a W-byte zeroed buffer
    if (sig > W) return fail(SBO_PANIC, "patas: significant bytes exceed value width (App. C4)");
    if (pos + sig > in_len) return fail(SBO_PANIC, "patas: read past input");
    uint64_t val = 0;
    std::memcpy(&val, in + pos, sig);
    pos += sig;
    if (ref == 0 || ref > i) return fail(SBO_PANIC, "patas: reference out of range");
    uint64_t prev = 0;
    std::memcpy(&prev, out.data() + begin + (i - ref) * size_t(W), size_t(W));
    uint64_t x = (tz >= 64 ? 0 : (val << tz)) ^ prev;
    put_bytes(out, &x, size_t(W));
  }
  return SBO_OK;
}

// ---- compress_sample_ratio (integer/mod.rs:310-347, double/mod.rs:309-347) ---------
template <class Tr>
double sample_ratio(const Stats<Tr> &s, int codec, const sbo_opts &opts) {
  using V = typename Tr::V;
  const size_t sample_count = 10, sample_size = 64; // compression/mod.rs:30,33
  std::vector<V> sv;
  Bytes sbits;
  const V *vals = s.values;
  BitView validity = s.validity;
  size_t n = s.tuple_count;
  if (!(n / sample_count <= sample_size)) {
    size_t sep = n / sample_count, rem = n % sample_count;
    MutableBitmap mb;
    for (size_t i = 0; i < sample_count; ++i) {
      size_t range_end = (i == sample_count - 1 ? sep + rem : sep) - sample_size;
      size_t begin = i * sep + size_t(sample_draw(opts.seed, uint32_t(codec), uint32_t(i), range_end));
      for (size_t k = 0; k < sample_size; ++k) {
        sv.push_back(s.values[begin + k]);
        mb.push(s.validity.get(int64_t(begin + k)));
      }
    }
    sbits = mb.bytes;
    vals = sv.data();
    n = sv.size();
    validity = s.validity.present() ? BitView{sbits.data(), 0} : BitView{};
  }
  Bytes tmp;
  size_t total = n * sizeof(V);
  int rc = SBO_OK;
  switch (codec) {
  case SBO_C_RLE: rle_compress<Tr>(vals, validity, n, opts.float_bitwise, tmp); break;
  case SBO_C_BITPACK:
    if constexpr (sizeof(V) == 4 && !Tr::is_float) rc = bitpack_compress(reinterpret_cast<const uint32_t *>(vals), n, false, tmp);
    break;
  case SBO_C_PATAS:
    if constexpr (Tr::is_float) patas_compress<V>(vals, n, tmp);
    break;
  default: break;
  }
  size_t size = rc ? total : tmp.size(); // unwrap_or(stats.total_bytes)
  return double(total) / double(size);
}

template <class Tr> bool bitpack_applicable(const Stats<Tr> &s) { // bp.rs:93-97
  if constexpr (Tr::is_float) return false;
  else return !(Tr::as_i64(s.min) < 0 || sizeof(typename Tr::V) != 4 || s.tuple_count % 128 != 0);
}

template <class Tr> double codec_ratio(int codec, const Stats<Tr> &s, const sbo_opts &opts) {
  switch (codec) {
  case SBO_C_ONEVALUE: return s.unique_count <= 1 ? double(s.tuple_count) : 0.0; // one_value.rs:53-59
  case SBO_C_FREQ: return freq_ratio<Tr>(s);
  case SBO_C_DICT: return dict_ratio<Tr>(s);
  case SBO_C_RLE: return sample_ratio<Tr>(s, SBO_C_RLE, opts);
  case SBO_C_PATAS: return sample_ratio<Tr>(s, SBO_C_PATAS, opts);
  case SBO_C_BITPACK: return bitpack_applicable<Tr>(s) ? sample_ratio<Tr>(s, SBO_C_BITPACK, opts) : 0.0;
  case SBO_C_DELTABP: // delta_bp.rs:97-109
    if (!bitpack_applicable<Tr>(s) || !s.is_sorted || s.null_count > 0) return 0.0;
    return sample_ratio<Tr>(s, SBO_C_BITPACK, opts) * 1.5;
  }
  return 0.0;
}

// Is a forced codec safe to emit for this input?  The reference's env switches apply
// unconditionally (and can corrupt, e.g. Bitpacking on 8-byte types); the test knob only
// fires where the codec is well defined.
template <class Tr> bool force_applicable(int codec, const Stats<Tr> &s) {
  switch (codec) {
  case SBO_C_FREQ:
  case SBO_C_DICT:
  case SBO_C_RLE: return true;
  case SBO_C_ONEVALUE: return s.unique_count <= 1;
  case SBO_C_BITPACK: return bitpack_applicable<Tr>(s);
  case SBO_C_DELTABP: return bitpack_applicable<Tr>(s) && s.is_sorted && s.null_count == 0;
  case SBO_C_PATAS: return Tr::is_float && sizeof(typename Tr::V) == 8 && s.tuple_count > 0;
  }
  return false;
}

// choose_compressor (integer/mod.rs:231-308, double/mod.rs:231-307); returns codec id
template <class Tr> int choose_compressor(const Stats<Tr> &s, const sbo_opts &opts) {
  if (opts.force_codec >= SBO_C_RLE && !forbidden(opts, opts.force_codec) && force_applicable<Tr>(opts.force_codec, s))
    return opts.force_codec;
  int result = opts.default_compression;
  if (opts.default_compress_ratio < 0) return result;
  double max_ratio = opts.default_compress_ratio;
  static const int int_order[] = {SBO_C_ONEVALUE, SBO_C_FREQ, SBO_C_DICT, SBO_C_RLE, SBO_C_BITPACK, SBO_C_DELTABP};
  static const int dbl_order[] = {SBO_C_ONEVALUE, SBO_C_FREQ, SBO_C_DICT, SBO_C_PATAS, SBO_C_RLE};
  const int *order = Tr::is_float ? dbl_order : int_order;
  int cnt = Tr::is_float ? 5 : 6;
  for (int i = 0; i < cnt; ++i) {
    int c = order[i];
    if (forbidden(opts, c)) continue;
    double r = codec_ratio<Tr>(c, s, opts);
    if (r > max_ratio) {
      max_ratio = r;
      result = c;
      if (r == double(s.tuple_count)) break;
    }
  }
  return result;
}

// compress_integer / compress_double (integer/mod.rs:35-70, double/mod.rs:32-67)
template <class Tr>
int compress_fixed(const typename Tr::V *values, BitView validity, size_t n, const sbo_opts &opts, Bytes &out) {
  using V = typename Tr::V;
  Stats<Tr> stats = gen_stats<Tr>(values, validity, n, opts.float_bitwise);
  int codec = choose_compressor<Tr>(stats, opts);
  out.push_back(uint8_t(codec));
  size_t pos = out.size();
  out.resize(pos + 8, 0);
  size_t body = out.size();
  int rc = SBO_OK;
  switch (codec) {
  case SBO_C_NONE:
  case SBO_C_LZ4:
  case SBO_C_ZSTD:
  case SBO_C_SNAPPY: {
    size_t w;
    rc = common_compress(codec, reinterpret_cast<const uint8_t *>(values), n * sizeof(V), out, w);
    break;
  }
  case SBO_C_RLE: rle_compress<Tr>(values, validity, n, opts.float_bitwise, out); break;
  case SBO_C_DICT: rc = dict_compress<Tr>(values, validity, n, opts, out); break;
  case SBO_C_ONEVALUE: onevalue_compress<Tr>(values, validity, n, out); break;
  case SBO_C_FREQ: rc = freq_compress<Tr>(stats, opts, out); break;
  case SBO_C_BITPACK:
  case SBO_C_DELTABP:
    if constexpr (sizeof(V) == 4 && !Tr::is_float)
      rc = bitpack_compress(reinterpret_cast<const uint32_t *>(values), n, codec == SBO_C_DELTABP, out);
    else rc = fail(SBO_PANIC, "bitpacking on non 4-byte type");
    break;
  case SBO_C_PATAS:
    if constexpr (Tr::is_float) patas_compress<V>(values, n, out);
    else rc = fail(SBO_OUT_OF_SPEC, "patas on integer");
    break;
  default: rc = fail(SBO_OUT_OF_SPEC, "Unknown compression codec");
  }
  if (rc) return rc;
  uint32_t compressed = uint32_t(out.size() - body), uncompressed = uint32_t(n * sizeof(V));
  std::memcpy(out.data() + pos, &compressed, 4);
  std::memcpy(out.data() + pos + 4, &uncompressed, 4);
  return SBO_OK;
}

// decompress_integer / decompress_double (integer/mod.rs:72-117, double/mod.rs:69-114).
// `in` is everything left in the reader; Extend codecs see the whole remainder
// (buffer_bytes(), integer/mod.rs:85-88,109) while Basic codecs see exactly
// compressed_size bytes (:107); `consumed` = 9 + compressed_size either way.
int decompress_fixed(const uint8_t *in, size_t in_len, size_t n, int W, bool is_float, Bytes &out, size_t &consumed) {
  if (in_len < 9) return fail(SBO_IO, "failed to fill whole buffer (compress header)");
  int codec = in[0];
  size_t compressed = load_le<uint32_t>(in + 1);
  const uint8_t *body = in + 9;
  size_t body_len = in_len - 9;
  if (body_len < compressed) return fail(SBO_IO, "failed to fill whole buffer (payload)");
  consumed = 9 + compressed;
  switch (codec) {
  case SBO_C_NONE:
  case SBO_C_LZ4:
  case SBO_C_ZSTD:
  case SBO_C_SNAPPY: {
    size_t start = out.size();
    out.resize(start + n * size_t(W));
    return common_decompress(codec, body, compressed, out.data() + start, n * size_t(W));
  }
  case SBO_C_RLE: return rle_decompress(body, body_len, n, W, out);
  case SBO_C_DICT: return dict_decompress(body, body_len, n, W, out);
  case SBO_C_ONEVALUE: return onevalue_decompress(body, body_len, n, W, out);
  case SBO_C_FREQ: return freq_decompress(body, body_len, n, W, is_float, out);
  case SBO_C_BITPACK:
    if (is_float) return fail(SBO_OUT_OF_SPEC, "Unknown compression codec Bitpacking"); // double/mod.rs:143-158
    return bitpack_decompress(body, body_len, n, W, false, out);
  case SBO_C_DELTABP:
    if (is_float) return fail(SBO_OUT_OF_SPEC, "Unknown compression codec DeltaBitpacking");
    return bitpack_decompress(body, body_len, n, W, true, out);
  case SBO_C_PATAS:
    if (!is_float) return fail(SBO_OUT_OF_SPEC, "Unknown compression codec Patas"); // integer/mod.rs:146-161
    return patas_decompress(body, body_len, n, W, out);
  default: return fail(SBO_OUT_OF_SPEC, "Unknown compression codec " + std::to_string(codec)); // mod.rs:78-80
  }
}

// ---------------------------------------------------------------------------------
// Binary (src/compression/binary/*)
// ---------------------------------------------------------------------------------
struct BinArr { // BinaryArray<O> slice
  const uint8_t *values;
  int64_t backing_len; // array.values().len()
  const void *offsets;
  bool large;
  BitView validity;
  size_t n;
  int64_t off(size_t i) const {
    return large ? static_cast<const int64_t *>(offsets)[i] : int64_t(static_cast<const int32_t *>(offsets)[i]);
  }
  std::string get(size_t i) const { return std::string(reinterpret_cast<const char *>(values) + off(i), size_t(off(i + 1) - off(i))); }
};
struct BinStats { // binary/mod.rs:253-291
  size_t tuple_count, total_bytes, unique_count, total_unique_size, null_count;
  struct Cnt {
    size_t count = 0, first = 0;
  };
  std::unordered_map<std::string, Cnt> distinct;
};
BinStats bin_gen_stats(const BinArr &a) {
  BinStats s;
  s.tuple_count = a.n;
  s.total_bytes = size_t(a.backing_len) + (a.n + 1) * (a.large ? 8 : 4);
  s.null_count = 0;
  if (a.validity.present())
    for (size_t i = 0; i < a.n; ++i) s.null_count += !a.validity.get(int64_t(i));
  for (size_t i = 0; i < a.n; ++i) {
    auto &c = s.distinct[a.get(i)];
    if (c.count == 0) c.first = i;
    ++c.count;
  }
  s.total_unique_size = 0;
  for (auto &kv : s.distinct) s.total_unique_size += kv.first.size() + 8;
  s.unique_count = s.distinct.size();
  return s;
}
size_t bin_max_count(const BinStats &s, size_t *first) {
  size_t max_count = 0, best_first = 0;
  for (auto &kv : s.distinct)
    if (kv.second.count > max_count || (kv.second.count == max_count && kv.second.first < best_first)) {
      max_count = kv.second.count;
      best_first = kv.second.first;
    }
  if (first) *first = best_first;
  return max_count;
}
double bin_ratio(int codec, const BinStats &s) {
  switch (codec) {
  case SBO_C_ONEVALUE: return s.unique_count <= 1 ? double(s.tuple_count) : 0.0; // binary/one_value.rs:43-49
  case SBO_C_FREQ: { // binary/freq.rs:147-169
    if (s.unique_count <= 1) return 0.0;
    if (double(s.null_count) / double(s.tuple_count) >= 0.9) return double(s.tuple_count - 1);
    size_t mc = bin_max_count(s, nullptr);
    return double(mc) / double(s.tuple_count) >= 0.9 ? double(s.tuple_count - 1) : 0.0;
  }
  case SBO_C_DICT: { // binary/dict.rs:43-53
    if (s.unique_count * 3 >= s.tuple_count) return 0.0;
    size_t after = s.total_unique_size + s.tuple_count * size_t(get_bits_needed(s.unique_count) / 8);
    after += s.tuple_count * 2 / 128;
    return double(s.total_bytes) / double(after);
  }
  }
  return 0.0;
}
int bin_choose(const BinStats &s, const sbo_opts &opts) { // binary/mod.rs:293-348
  if ((opts.force_codec == SBO_C_FREQ || opts.force_codec == SBO_C_DICT ||
       (opts.force_codec == SBO_C_ONEVALUE && s.unique_count <= 1)) &&
      !forbidden(opts, opts.force_codec))
    return opts.force_codec;
  int result = opts.default_compression;
  if (opts.default_compress_ratio < 0) return result;
  double max_ratio = opts.default_compress_ratio;
  static const int order[] = {SBO_C_ONEVALUE, SBO_C_FREQ, SBO_C_DICT};
  for (int c : order) {
    if (forbidden(opts, c)) continue;
    double r = bin_ratio(c, s);
    if (r > max_ratio) {
      max_ratio = r;
      result = c;
      if (r == double(s.tuple_count)) break;
    }
  }
  return result;
}
int compress_binary(const BinArr &a, const sbo_opts &opts, Bytes &out) { // binary/mod.rs:26-93
  BinStats stats = bin_gen_stats(a);
  int codec = bin_choose(stats, opts);
  const size_t OW = a.large ? 8 : 4;
  if (codec <= SBO_C_SNAPPY) {
    Bytes offs((a.n + 1) * OW);
    int64_t first = a.off(0);
    for (size_t i = 0; i <= a.n; ++i) { // rebased so first == 0 (:46-55)
      int64_t v = a.off(i) - first;
      if (a.large) std::memcpy(offs.data() + 8 * i, &v, 8);
      else {
        int32_t v32 = int32_t(v);
        std::memcpy(offs.data() + 4 * i, &v32, 4);
      }
    }
    for (int part = 0; part < 2; ++part) {
      const uint8_t *src = part == 0 ? offs.data() : a.values + first;
      size_t len = part == 0 ? offs.size() : size_t(a.off(a.n) - first);
      out.push_back(uint8_t(codec));
      size_t pos = out.size();
      out.resize(pos + 8, 0);
      size_t w = 0;
      int rc = common_compress(codec, src, len, out, w);
      if (rc) return rc;
      uint32_t c32 = uint32_t(w), u32 = uint32_t(len);
      std::memcpy(out.data() + pos, &c32, 4);
      std::memcpy(out.data() + pos + 4, &u32, 4);
    }
    return SBO_OK;
  }
  out.push_back(uint8_t(codec));
  size_t pos = out.size();
  out.resize(pos + 8, 0);
  size_t body = out.size();
  if (codec == SBO_C_ONEVALUE) { // binary/one_value.rs:51-69
    std::string val;
    for (size_t i = 0; i < a.n; ++i)
      if (a.validity.get(int64_t(i))) {
        val = a.get(i);
        break;
      }
    put_le<uint32_t>(out, uint32_t(val.size()));
    put_bytes(out, val.data(), val.size());
  } else if (codec == SBO_C_FREQ) { // binary/freq.rs:44-100
    bool top_is_null = double(stats.null_count) / double(stats.tuple_count) >= 0.9;
    std::string top;
    if (!top_is_null) {
      size_t first = 0;
      if (bin_max_count(stats, &first)) top = a.get(first);
    }
    std::vector<uint32_t> rows;
    for (size_t i = 0; i < a.n; ++i)
      if (a.validity.get(int64_t(i)) && (top_is_null || a.get(i) != top)) rows.push_back(uint32_t(i));
    put_le<uint64_t>(out, top.size());
    put_bytes(out, top.data(), top.size());
    put_le<uint32_t>(out, uint32_t(roaring_serialized_size(rows)));
    roaring_serialize(rows, out);
    for (uint32_t r : rows) { // exceptions inline, uncompressed (:94-98)
      std::string v = a.get(r);
      put_le<uint64_t>(out, v.size());
      put_bytes(out, v.data(), v.size());
    }
  } else if (codec == SBO_C_DICT) { // binary/dict.rs:55-93
    std::unordered_map<std::string, uint32_t> interner;
    std::vector<std::string> sets;
    std::vector<uint32_t> indices;
    for (size_t i = 0; i < a.n; ++i) {
      if (!a.validity.get(int64_t(i)) && !indices.empty()) {
        indices.push_back(indices.back());
      } else {
        std::string v = a.get(i);
        auto it = interner.find(v);
        if (it == interner.end()) {
          uint32_t k = uint32_t(sets.size());
          interner.emplace(v, k);
          sets.push_back(v);
          indices.push_back(k);
        } else indices.push_back(it->second);
      }
    }
    sbo_opts sub = forbid(opts, SBO_C_DICT);
    int rc = compress_fixed<IntTr<uint32_t>>(indices.data(), BitView{}, indices.size(), sub, out);
    if (rc) return rc;
    put_le<uint32_t>(out, uint32_t(sets.size()));
    for (auto &v : sets) {
      put_le<uint64_t>(out, v.size());
      put_bytes(out, v.data(), v.size());
    }
  } else {
    return fail(SBO_OUT_OF_SPEC, "Unknown compression codec");
  }
  uint32_t c32 = uint32_t(out.size() - body), u32 = uint32_t(a.backing_len); // binary/mod.rs:88
  std::memcpy(out.data() + pos, &c32, 4);
  std::memcpy(out.data() + pos + 4, &u32, 4);
  return SBO_OK;
}

struct BinOut {
  bool large = false;
  std::vector<int64_t> offsets; // kept as i64; narrowed on export
  Bytes values;
};
// decompress_binary (binary/mod.rs:95-183)
int decompress_binary(const uint8_t *in, size_t in_len, size_t n, BinOut &o, size_t &consumed) {
  if (in_len < 9) return fail(SBO_IO, "failed to fill whole buffer (compress header)");
  int codec = in[0];
  size_t compressed = load_le<uint32_t>(in + 1);
  const uint8_t *body = in + 9;
  size_t body_len = in_len - 9;
  if (body_len < compressed) return fail(SBO_IO, "failed to fill whole buffer (payload)");
  consumed = 9 + compressed;
  const size_t OW = o.large ? 8 : 4;
  if (codec <= SBO_C_SNAPPY) {
    Bytes raw((n + 1) * OW);
    int rc = common_decompress(codec, body, compressed, raw.data(), raw.size());
    if (rc) return rc;
    bool had = !o.offsets.empty();
    int64_t last = had ? o.offsets.back() : 0;
    // appended raw offsets are rebased by `last` and the leading 0 dropped when the
    // output already holds offsets (:136-144)
    for (size_t i = had ? 1 : 0; i <= n; ++i) {
      int64_t v = o.large ? load_le<int64_t>(raw.data() + 8 * i) : int64_t(load_le<int32_t>(raw.data() + 4 * i));
      o.offsets.push_back(last + v);
    }
    in += consumed;
    in_len -= consumed;
    if (in_len < 9) return fail(SBO_IO, "failed to fill whole buffer (values header)");
    size_t c2 = load_le<uint32_t>(in + 1), u2 = load_le<uint32_t>(in + 5);
    if (in_len - 9 < c2) return fail(SBO_IO, "failed to fill whole buffer (values payload)");
    size_t start = o.values.size();
    o.values.resize(start + u2);
    rc = common_decompress(codec, in + 9, c2, o.values.data() + start, u2); // same codec `c` (:166)
    if (rc) return rc;
    consumed += 9 + c2;
    return SBO_OK;
  }
  auto push_first = [&]() {
    if (o.offsets.empty()) o.offsets.push_back(0);
  };
  if (codec == SBO_C_ONEVALUE) { // binary/one_value.rs:71-99
    if (body_len < 4) return fail(SBO_IO, "failed to fill whole buffer");
    size_t len = load_le<uint32_t>(body);
    if (body_len - 4 < len) return fail(SBO_OUT_OF_SPEC, "data size is less than " + std::to_string(len));
    push_first();
    for (size_t i = 0; i < n; ++i) {
      put_bytes(o.values, body + 4, len);
      o.offsets.push_back(int64_t(o.values.size()));
    }
    return SBO_OK;
  }
  if (codec == SBO_C_FREQ) { // binary/freq.rs:102-145
    size_t pos = 0;
    if (body_len < 8) return fail(SBO_IO, "failed to fill whole buffer");
    size_t len = size_t(load_le<uint64_t>(body));
    pos = 8;
    if (body_len - pos < len) return fail(SBO_OUT_OF_SPEC, "data size is less than " + std::to_string(len));
    const uint8_t *top = body + pos;
    pos += len;
    if (body_len - pos < 4) return fail(SBO_IO, "failed to fill whole buffer");
    size_t bm = load_le<uint32_t>(body + pos);
    pos += 4;
    if (body_len - pos < bm) return fail(SBO_PANIC, "freq: bitmap slice out of range");
    std::vector<uint32_t> rows;
    int rc = roaring_deserialize(body + pos, bm, rows);
    if (rc) return rc;
    pos += bm;
    push_first();
    size_t r = 0;
    for (size_t i = 0; i < n; ++i) {
      while (r < rows.size() && rows[r] < i) ++r;
      if (r < rows.size() && rows[r] == i) {
        if (body_len - pos < 8) return fail(SBO_IO, "failed to fill whole buffer");
        size_t l = size_t(load_le<uint64_t>(body + pos));
        pos += 8;
        if (body_len - pos < l) return fail(SBO_OUT_OF_SPEC, "data size is less than " + std::to_string(l));
        put_bytes(o.values, body + pos, l);
        pos += l;
      } else {
        put_bytes(o.values, top, len);
      }
      o.offsets.push_back(int64_t(o.values.size()));
    }
    return SBO_OK;
  }
  if (codec == SBO_C_DICT) { // binary/dict.rs:95-141
    Bytes idx_bytes;
    size_t used = 0;
    int rc = decompress_fixed(body, body_len, n, 4, false, idx_bytes, used);
    if (rc) return rc;
    size_t pos = used;
    if (body_len - pos < 4) return fail(SBO_IO, "failed to fill whole buffer");
    size_t k = load_le<uint32_t>(body + pos);
    pos += 4;
    std::vector<size_t> data_off{0};
    Bytes data;
    for (size_t j = 0; j < k; ++j) {
      if (body_len - pos < 8) return fail(SBO_IO, "failed to fill whole buffer");
      size_t l = size_t(load_le<uint64_t>(body + pos));
      pos += 8;
      if (body_len - pos < l) return fail(SBO_OUT_OF_SPEC, "data size is less than " + std::to_string(l));
      put_bytes(data, body + pos, l);
      data_off.push_back(data.size());
      pos += l;
    }
    int64_t last = 0;
    if (o.offsets.empty()) o.offsets.push_back(0);
    else last = o.offsets.back();
    size_t count = idx_bytes.size() / 4;
    for (size_t i = 0; i < count; ++i) {
      uint32_t id = load_le<uint32_t>(idx_bytes.data() + 4 * i);
      if (size_t(id) + 1 >= data_off.size()) return fail(SBO_PANIC, "dict: index out of bounds");
      put_bytes(o.values, data.data() + data_off[id], data_off[id + 1] - data_off[id]);
      last += int64_t(data_off[id + 1] - data_off[id]);
      o.offsets.push_back(last);
    }
    return SBO_OK;
  }
  return fail(SBO_OUT_OF_SPEC, "Unknown compression codec " + std::to_string(codec));
}

// ---------------------------------------------------------------------------------
// Boolean (src/compression/boolean/*)
// ---------------------------------------------------------------------------------
struct BoolStats { // boolean/mod.rs:141-192
  size_t rows = 0, total_bytes = 0, null_count = 0, false_count = 0, true_count = 0;
};
BoolStats bool_gen_stats(const BitView &values, const BitView &validity, size_t n) {
  BoolStats s;
  s.rows = n;
  s.total_bytes = n / 8; // array.values().len() / 8
  for (size_t i = 0; i < n; ++i) {
    if (!validity.get(int64_t(i))) ++s.null_count;
    else if (values.get(int64_t(i))) ++s.true_count;
    else ++s.false_count;
  }
  return s;
}
void bool_rle_compress(const BitView &values, const BitView &validity, size_t n, Bytes &out) { // boolean/rle.rs:31-38
  std::vector<uint8_t> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = values.get(int64_t(i));
  rle_compress<IntTr<uint8_t>>(v.data(), validity, n, false, out);
}
double bool_sample_ratio(const BitView &values, const BitView &validity, size_t n, const sbo_opts &opts) { // boolean/mod.rs:241-278
  const size_t sample_count = 10, sample_size = 64;
  MutableBitmap sv, sm;
  BitView vv = values, vm = validity;
  size_t m = n;
  if (!(n / sample_count <= sample_size)) {
    size_t sep = n / sample_count, rem = n % sample_count;
    for (size_t i = 0; i < sample_count; ++i) {
      size_t range_end = (i == sample_count - 1 ? sep + rem : sep) - sample_size;
      size_t begin = i * sep + size_t(sample_draw(opts.seed, SBO_C_RLE, uint32_t(i), range_end));
      for (size_t k = 0; k < sample_size; ++k) {
        sv.push(values.get(int64_t(begin + k)));
        sm.push(validity.get(int64_t(begin + k)));
      }
    }
    m = sv.len;
    vv = BitView{sv.bytes.data(), 0};
    vm = validity.present() ? BitView{sm.bytes.data(), 0} : BitView{};
  }
  Bytes tmp;
  bool_rle_compress(vv, vm, m, tmp);
  return double(m / 8) / double(tmp.size());
}
int compress_boolean(const BitView &values, const BitView &validity, size_t n, const sbo_opts &opts, Bytes &out) { // boolean/mod.rs:23-61
  BoolStats s = bool_gen_stats(values, validity, n);
  int codec = opts.default_compression;
  bool one = s.true_count == 0 || s.false_count == 0; // boolean/one_value.rs:36-42
  if (opts.force_codec == SBO_C_RLE && !forbidden(opts, SBO_C_RLE)) codec = SBO_C_RLE;
  else if (opts.force_codec == SBO_C_ONEVALUE && one && !forbidden(opts, SBO_C_ONEVALUE)) codec = SBO_C_ONEVALUE;
  else if (opts.default_compress_ratio >= 0) { // boolean/mod.rs:194-239
    double max_ratio = opts.default_compress_ratio;
    static const int order[] = {SBO_C_ONEVALUE, SBO_C_RLE};
    for (int c : order) {
      if (forbidden(opts, c)) continue;
      double r = c == SBO_C_ONEVALUE ? (one ? double(s.rows) : 0.0) : bool_sample_ratio(values, validity, n, opts);
      if (r > max_ratio) {
        max_ratio = r;
        codec = c;
        if (r == double(s.rows)) break;
      }
    }
  }
  out.push_back(uint8_t(codec));
  size_t pos = out.size();
  out.resize(pos + 8, 0);
  size_t body = out.size();
  if (codec <= SBO_C_SNAPPY) { // re-packed to bit offset 0 (:47-52)
    Bytes packed((n + 7) / 8, 0);
    for (size_t i = 0; i < n; ++i)
      if (values.get(int64_t(i))) packed[i >> 3] |= uint8_t(1u << (i & 7));
    size_t w;
    int rc = common_compress(codec, packed.data(), packed.size(), out, w);
    if (rc) return rc;
  } else if (codec == SBO_C_RLE) {
    bool_rle_compress(values, validity, n, out);
  } else if (codec == SBO_C_ONEVALUE) { // boolean/one_value.rs:44-52
    uint8_t val = 0;
    for (size_t i = 0; i < n; ++i)
      if (validity.get(int64_t(i))) {
        val = values.get(int64_t(i));
        break;
      }
    out.push_back(val);
  } else return fail(SBO_OUT_OF_SPEC, "Unknown compression codec");
  uint32_t c32 = uint32_t(out.size() - body), u32 = uint32_t(n); // rows, not bytes (:59)
  std::memcpy(out.data() + pos, &c32, 4);
  std::memcpy(out.data() + pos + 4, &u32, 4);
  return SBO_OK;
}
int decompress_boolean(const uint8_t *in, size_t in_len, size_t n, MutableBitmap &out, size_t &consumed) { // boolean/mod.rs:63-102
  if (in_len < 9) return fail(SBO_IO, "failed to fill whole buffer (compress header)");
  int codec = in[0];
  size_t compressed = load_le<uint32_t>(in + 1);
  const uint8_t *body = in + 9;
  size_t body_len = in_len - 9;
  if (body_len < compressed) return fail(SBO_IO, "failed to fill whole buffer (payload)");
  consumed = 9 + compressed;
  if (codec <= SBO_C_SNAPPY) {
    Bytes buf((n + 7) / 8);
    int rc = common_decompress(codec, body, compressed, buf.data(), buf.size());
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) out.push((buf[i >> 3] >> (i & 7)) & 1);
    return SBO_OK;
  }
  if (codec == SBO_C_RLE) { // boolean/rle.rs:41-55 (pushes whole runs; no clamp to n)
    size_t pos = 0, num = 0;
    while (pos < body_len) {
      if (pos + 5 > body_len) return fail(SBO_IO, "failed to fill whole buffer");
      uint32_t len = load_le<uint32_t>(body + pos);
      bool t = body[pos + 4] != 0;
      pos += 5;
      out.extend_constant(len, t);
      num += len;
      if (num >= n) break;
    }
    return SBO_OK;
  }
  if (codec == SBO_C_ONEVALUE) { // boolean/one_value.rs:54-61
    if (body_len == 0) return fail(SBO_OUT_OF_SPEC, "data size is less than 1");
    out.extend_constant(n, body[0] > 0);
    return SBO_OK;
  }
  return fail(SBO_OUT_OF_SPEC, "Unknown compression codec " + std::to_string(codec));
}

// ---------------------------------------------------------------------------------
// page level: validity / nested levels + value block
// ---------------------------------------------------------------------------------
int type_width(int t) {
  switch (t) {
  case SBO_I8:
  case SBO_U8: return 1;
  case SBO_I16:
  case SBO_U16: return 2;
  case SBO_I32:
  case SBO_U32:
  case SBO_F32: return 4;
  case SBO_I64:
  case SBO_U64:
  case SBO_F64: return 8;
  case SBO_I128: return 16;
  case SBO_I256: return 32;
  }
  return 0;
}
bool type_is_float(int t) { return t == SBO_F32 || t == SBO_F64; }

// read_validity (src/read/read_basic.rs:36-63)
int read_validity(const uint8_t *in, size_t in_len, size_t n, MutableBitmap &out, size_t &consumed) {
  if (in_len < 4) return fail(SBO_IO, "failed to fill whole buffer (def levels len)");
  size_t L = load_le<uint32_t>(in);
  consumed = 4 + L;
  if (L == 0) return SBO_OK; // pushes nothing (:43-45)
  if (in_len - 4 < L) return fail(SBO_IO, "failed to fill whole buffer (def levels)");
  const uint8_t *p = in + 4;
  size_t pos = 0;
  while (pos < L) { // Decoder::new(def_levels, 1): every run pushes `length` bits
    uint64_t header;
    size_t used = uleb_decode(p + pos, L - pos, header);
    if (!used) return fail(SBO_PANIC, "read_validity: encoded.unwrap()");
    pos += used;
    if (!(header & 1)) return fail(SBO_PANIC, "read_validity: HybridEncoded::Rle => unreachable!()");
    size_t bytes = std::min(size_t(header >> 1), L - pos);
    if (bytes * 8 < n) return fail(SBO_PANIC, "read_validity: BitmapIter out of range");
    for (size_t i = 0; i < n; ++i) out.push((p[pos + (i >> 3)] >> (i & 7)) & 1);
    pos += bytes;
  }
  return SBO_OK;
}

struct Nested { // arrow2 NestedState entry
  int kind;
  bool nullable;
  std::vector<int64_t> offsets; // lists: start offsets (create_list appends the end)
  MutableBitmap validity;
  size_t len = 0;
  bool is_nullable() const { return nullable; }
  bool is_repeated() const { return kind == SBO_N_LIST; }
  bool is_required() const { return kind == SBO_N_STRUCT; } // SURVEY App. D.4 (unverified upstream)
  void push(int64_t length, bool is_valid) {
    if (kind == SBO_N_LIST) offsets.push_back(length);
    if (nullable) validity.push(is_valid);
    ++len;
  }
};

} // namespace

struct sbo_col {
  sbo_leaf leaf;
  Bytes values;          // primitives
  MutableBitmap bits;    // boolean values
  BinOut bin;            // binary
  Bytes offsets_export;  // narrowed offsets
  MutableBitmap validity;
  bool has_validity = false;
  int64_t len = 0;
  std::vector<Nested> nested; // accumulated over pages (all but the leaf entry)
};

namespace {

// read_validity_nested (src/read/read_basic.rs:65-173)
int read_validity_nested(sbo_col *c, const uint8_t *in, size_t in_len, size_t num_values, size_t &leaf_len,
                         size_t &consumed) {
  if (in_len < 12) return fail(SBO_IO, "failed to fill whole buffer (nested header)");
  size_t additional = load_le<uint32_t>(in), rep_len = load_le<uint32_t>(in + 4), def_len = load_le<uint32_t>(in + 8);
  if (in_len - 12 < rep_len + def_len) return fail(SBO_IO, "failed to fill whole buffer (levels)");
  consumed = 12 + rep_len + def_len;
  const sbo_leaf &lf = c->leaf;
  int depth_n = lf.n_nested;
  uint32_t max_rep = 0, max_def = 0;
  std::vector<uint32_t> cum_sum(depth_n + 1, 0), cum_rep(depth_n + 1, 0);
  for (int i = 0; i < depth_n; ++i) {
    bool rep = lf.nested_kind[i] == SBO_N_LIST, nul = lf.nested_nullable[i] != 0;
    cum_sum[i + 1] = cum_sum[i] + uint32_t(nul) + uint32_t(rep);
    cum_rep[i + 1] = cum_rep[i] + uint32_t(rep);
  }
  max_rep = cum_rep[depth_n];
  max_def = cum_sum[depth_n];
  std::vector<uint32_t> reps, defs;
  int rc = hybrid_rle_decode(in + 12, rep_len, get_bits_needed(max_rep), num_values, reps);
  if (rc) return rc;
  rc = hybrid_rle_decode(in + 12 + rep_len, def_len, get_bits_needed(max_def), num_values, defs);
  if (rc) return rc;
  if (c->nested.empty()) {
    c->nested.resize(depth_n);
    for (int i = 0; i < depth_n; ++i) {
      c->nested[i].kind = lf.nested_kind[i];
      c->nested[i].nullable = lf.nested_nullable[i] != 0;
    }
  }
  // per-page NestedState starts empty in the reference (init_nested); lengths pushed are
  // page-relative child lengths.  We accumulate across pages by rebasing on the running
  // child length at page start (what concatenating per-page arrays yields).
  std::vector<size_t> base(depth_n + 1, 0);
  for (int i = 0; i < depth_n; ++i) base[i] = c->nested[i].len;
  size_t rows = 0;
  size_t leaf_before = c->nested[depth_n - 1].len;
  for (size_t e = 0; e < num_values; ++e) {
    uint32_t rep = reps[e], def = defs[e];
    if (rep == 0) ++rows;
    bool is_required = false;
    for (int d = 0; d < depth_n; ++d) {
      bool right_level = rep <= cum_rep[d] && def >= cum_sum[d];
      if (is_required || right_level) {
        int64_t length = d + 1 < depth_n ? int64_t(c->nested[d + 1].len) : 1;
        Nested &nest = c->nested[d];
        bool is_valid = nest.is_nullable() && def > cum_sum[d];
        nest.push(length, is_valid);
        is_required = nest.is_required() && !is_valid;
        if (d == depth_n - 1 && nest.is_nullable()) {
          bool v = (def != cum_sum[d]) || !nest.is_nullable();
          c->validity.push(right_level && v);
          c->has_validity = true;
        }
      }
    }
    uint32_t next_rep = e + 1 < num_values ? reps[e + 1] : 0;
    if (next_rep == 0 && rows == additional) break;
  }
  leaf_len = c->nested[depth_n - 1].len - leaf_before;
  return SBO_OK;
}

template <class Tr> int compress_typed(const sbo_array *a, const sbo_opts *o, Bytes &out) {
  BitView v{a->validity, a->validity_offset};
  return compress_fixed<Tr>(static_cast<const typename Tr::V *>(a->values), v, size_t(a->n), *o, out);
}

int compress_values(int type, const sbo_array *a, const sbo_opts *o, Bytes &out) {
  switch (type) { // src/write/primitive.rs:30-96
  case SBO_NULL: return SBO_OK;
  case SBO_I8: return compress_typed<IntTr<int8_t>>(a, o, out);
  case SBO_I16: return compress_typed<IntTr<int16_t>>(a, o, out);
  case SBO_I32: return compress_typed<IntTr<int32_t>>(a, o, out);
  case SBO_I64: return compress_typed<IntTr<int64_t>>(a, o, out);
  case SBO_U8: return compress_typed<IntTr<uint8_t>>(a, o, out);
  case SBO_U16: return compress_typed<IntTr<uint16_t>>(a, o, out);
  case SBO_U32: return compress_typed<IntTr<uint32_t>>(a, o, out);
  case SBO_U64: return compress_typed<IntTr<uint64_t>>(a, o, out);
  case SBO_F32: return compress_typed<FloatTr<float, uint32_t>>(a, o, out);
  case SBO_F64: return compress_typed<FloatTr<double, uint64_t>>(a, o, out);
  case SBO_I128: return compress_typed<IntTr<i128>>(a, o, out);
  case SBO_I256: return compress_typed<IntTr<i256>>(a, o, out);
  case SBO_BOOL:
    return compress_boolean(BitView{static_cast<const uint8_t *>(a->values), a->values_bit_offset},
                            BitView{a->validity, a->validity_offset}, size_t(a->n), *o, out);
  case SBO_BINARY:
  case SBO_LARGE_BINARY: {
    BinArr b{static_cast<const uint8_t *>(a->values), a->values_backing_len, a->offsets, type == SBO_LARGE_BINARY,
             BitView{a->validity, a->validity_offset}, size_t(a->n)};
    return compress_binary(b, *o, out);
  }
  }
  return fail(SBO_NYI, "type not implemented");
}

void export_buf(const Bytes &b, sbo_buf *out) {
  size_t need = out->len + b.size();
  if (need > out->cap) {
    size_t cap = std::max(need, out->cap * 2 + 64);
    out->data = static_cast<uint8_t *>(std::realloc(out->data, cap));
    out->cap = cap;
  }
  std::memcpy(out->data + out->len, b.data(), b.size());
  out->len = need;
}

int stat_block(int type, const uint8_t *in, size_t len, std::string &s, size_t &consumed) { // src/stat.rs:86-152
  if (len < 9) return fail(SBO_IO, "stat: short header");
  int codec = in[0];
  size_t compressed = load_le<uint32_t>(in + 1);
  consumed = 9 + compressed;
  const uint8_t *body = in + 9;
  size_t body_len = len - 9;
  bool bin = type == SBO_BINARY || type == SBO_LARGE_BINARY;
  static const char *names[] = {"None", "Lz4", "Zstd", "Snappy", "", "", "", "", "", "", "Rle", "Dict", "OneValue", "Freq", "Bitpacking", "DeltaBitpacking", "Patas"};
  if (codec > 16 || names[codec][0] == 0) return fail(SBO_OUT_OF_SPEC, "Unknown compression codec");
  s += names[codec];
  if (codec == SBO_C_DICT) {
    std::string sub;
    size_t used;
    int rc = stat_block(SBO_U32, body, body_len, sub, used);
    if (rc) return rc;
    if (body_len < used + 4) return fail(SBO_IO, "stat: short dict");
    s += "(" + sub + ")[k=" + std::to_string(load_le<uint32_t>(body + used)) + "]";
  } else if (codec == SBO_C_FREQ && !bin) {
    int W = type_width(type);
    if (body_len < size_t(W) + 4) return fail(SBO_IO, "stat: short freq");
    size_t bm = load_le<uint32_t>(body + W);
    std::string sub;
    size_t used;
    if (body_len < size_t(W) + 4 + bm) return fail(SBO_IO, "stat: short freq bitmap");
    int rc = stat_block(type, body + W + 4 + bm, body_len - W - 4 - bm, sub, used);
    if (rc) return rc;
    s += "(" + sub + ")";
  }
  return SBO_OK;
}

} // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

const char *sbo_last_error(void) { return g_err.c_str(); }
void sbo_buf_free(sbo_buf *b) {
  std::free(b->data);
  b->data = nullptr;
  b->len = b->cap = 0;
}

int sbo_write_validity(const uint8_t *validity, int64_t bit_offset, int64_t n, sbo_buf *out) { // serialize.rs:200-215
  Bytes scratch, w;
  encode_bool_levels(BitView{validity, bit_offset}, n, scratch);
  put_le<uint32_t>(w, uint32_t(scratch.size()));
  put_bytes(w, scratch.data(), scratch.size());
  export_buf(w, out);
  return SBO_OK;
}

int sbo_compress_values(int32_t type, const sbo_array *arr, const sbo_opts *opts, sbo_buf *out) {
  Bytes b;
  int rc = compress_values(type, arr, opts, b);
  if (rc) return rc;
  export_buf(b, out);
  return SBO_OK;
}

int sbo_write_page(const sbo_leaf *leaf, const sbo_array *arr, const sbo_opts *opts, sbo_buf *out) { // write_simple
  if (leaf->n_nested > 1) return fail(SBO_NYI, "sbo_write_page: nested pages are written by the level helpers");
  if (leaf->type == SBO_NULL) return SBO_OK; // serialize.rs:63
  if (leaf->nullable) {
    int rc = sbo_write_validity(arr->validity, arr->validity_offset, arr->n, out);
    if (rc) return rc;
  }
  return sbo_compress_values(leaf->type, arr, opts, out);
}

sbo_col *sbo_col_new(const sbo_leaf *leaf) {
  sbo_col *c = new sbo_col();
  c->leaf = *leaf;
  c->bin.large = leaf->type == SBO_LARGE_BINARY;
  return c;
}
void sbo_col_free(sbo_col *c) { delete c; }

int sbo_col_read_page(sbo_col *c, const uint8_t *page, size_t len, uint64_t num_values) {
  const sbo_leaf &lf = c->leaf;
  size_t pos = 0, n = size_t(num_values);
  if (lf.type == SBO_NULL) { // read/array/null.rs: length only
    c->len += int64_t(n);
    return SBO_OK;
  }
  if (lf.n_nested > 1) {
    size_t used = 0, leaf_len = 0;
    int rc = read_validity_nested(c, page, len, n, leaf_len, used);
    if (rc) return rc;
    pos = used;
    n = leaf_len;
  } else if (lf.nullable) {
    size_t used = 0;
    int rc = read_validity(page, len, n, c->validity, used);
    if (rc) return rc;
    c->has_validity = true;
    pos = used;
  }
  size_t used = 0;
  int rc;
  if (lf.type == SBO_BOOL) rc = decompress_boolean(page + pos, len - pos, n, c->bits, used);
  else if (lf.type == SBO_BINARY || lf.type == SBO_LARGE_BINARY) rc = decompress_binary(page + pos, len - pos, n, c->bin, used);
  else rc = decompress_fixed(page + pos, len - pos, n, type_width(lf.type), type_is_float(lf.type), c->values, used);
  if (rc) return rc;
  c->len += int64_t(n);
  return SBO_OK;
}
// batch read of a whole column body (read_integer / read_double / read_binary / read_boolean page loops,
// src/read/array/integer.rs:210-238): pages back to back, `metas` = (length, num_values) pairs
int sbo_col_read_pages(sbo_col *c, const uint8_t *body, size_t nbytes, const uint64_t *metas, size_t n_pages) {
  size_t pos = 0;
  for (size_t p = 0; p < n_pages; ++p) {
    const size_t len = size_t(metas[2 * p]);
    if (len > nbytes - pos) return fail(SBO_IO, "column body shorter than its page lengths");
    int rc = sbo_col_read_page(c, body + pos, len, metas[2 * p + 1]);
    if (rc) return rc;
    pos += len;
  }
  return SBO_OK;
}

// page loop of NativeWriter::encode_chunk for one flat leaf (write/common.rs:71-115): pages of `page_rows`
// rows appended to `out`; metas_out receives (length, num_values) per page; page p samples with seed + p
int sbo_write_column(const sbo_leaf *leaf, const sbo_array *arr, const sbo_opts *opts, uint64_t page_rows, sbo_buf *out,
                     uint64_t *metas_out, size_t metas_cap, size_t *n_pages_out) {
  if (leaf->type == SBO_BINARY || leaf->type == SBO_LARGE_BINARY || leaf->type == SBO_BOOL || leaf->type == SBO_NULL)
    return fail(SBO_NYI, "sbo_write_column: fixed-width leaves only (the python loop covers the others)");
  const size_t n = size_t(arr->n), W = size_t(type_width(leaf->type));
  const size_t pr = page_rows ? size_t(std::min<uint64_t>(page_rows, n)) : n;
  Bytes all;
  size_t np = 0;
  for (size_t r = 0; r < n; r += std::max<size_t>(pr, 1), ++np) {
    if (np >= metas_cap) return fail(SBO_PANIC, "metas_out too small");
    sbo_array a = *arr;
    a.values = static_cast<const uint8_t *>(arr->values) + r * W;
    a.n = int64_t(std::min(pr, n - r));
    a.validity_offset = arr->validity_offset + int64_t(r);
    sbo_opts o = *opts;
    o.seed = opts->seed + np;
    sbo_buf b{nullptr, 0, 0};
    int rc = sbo_write_page(leaf, &a, &o, &b);
    if (rc) return rc;
    put_bytes(all, b.data, b.len);
    metas_out[2 * np] = b.len;
    metas_out[2 * np + 1] = uint64_t(a.n);
    sbo_buf_free(&b);
  }
  *n_pages_out = np;
  export_buf(all, out);
  return SBO_OK;
}

int64_t sbo_col_len(const sbo_col *c) { return c->len; }
const uint8_t *sbo_col_values(const sbo_col *c, size_t *nbytes) {
  if (c->leaf.type == SBO_BOOL) {
    *nbytes = c->bits.bytes.size();
    return c->bits.bytes.data();
  }
  if (c->leaf.type == SBO_BINARY || c->leaf.type == SBO_LARGE_BINARY) {
    *nbytes = c->bin.values.size();
    return c->bin.values.data();
  }
  *nbytes = c->values.size();
  return c->values.data();
}
const uint8_t *sbo_col_offsets(const sbo_col *cc, size_t *nbytes) {
  sbo_col *c = const_cast<sbo_col *>(cc);
  c->offsets_export.clear();
  for (int64_t v : c->bin.offsets) {
    if (c->bin.large) put_le<int64_t>(c->offsets_export, v);
    else put_le<int32_t>(c->offsets_export, int32_t(v));
  }
  *nbytes = c->offsets_export.size();
  return c->offsets_export.data();
}
const uint8_t *sbo_col_validity(const sbo_col *c, size_t *nbits) {
  *nbits = c->validity.len;
  return c->has_validity ? c->validity.bytes.data() : nullptr;
}
int32_t sbo_col_nested_depths(const sbo_col *c) { return int32_t(c->nested.size()); }
const int64_t *sbo_col_nested_offsets(const sbo_col *c, int32_t depth, size_t *n) {
  *n = c->nested[depth].offsets.size();
  return c->nested[depth].offsets.data();
}
const uint8_t *sbo_col_nested_validity(const sbo_col *c, int32_t depth, size_t *nbits) {
  *nbits = c->nested[depth].validity.len;
  return c->nested[depth].validity.bytes.data();
}

int sbo_stat_block(int32_t type, const uint8_t *block, size_t len, char *out, size_t out_cap) {
  std::string s;
  size_t used;
  int rc = stat_block(type, block, len, s, used);
  if (rc) return rc;
  std::snprintf(out, out_cap, "%s", s.c_str());
  return SBO_OK;
}
int64_t sbo_page_value_block_offset(const sbo_leaf *leaf, const uint8_t *page, size_t len) {
  if (leaf->n_nested > 1) {
    if (len < 12) return -1;
    return 12 + int64_t(load_le<uint32_t>(page + 4)) + int64_t(load_le<uint32_t>(page + 8));
  }
  if (!leaf->nullable) return 0;
  if (len < 4) return -1;
  return 4 + int64_t(load_le<uint32_t>(page));
}

uint32_t sbo_bp4x_num_bits(const uint32_t *b) { return bp4x_num_bits(b); }
size_t sbo_bp4x_compress(const uint32_t *in, uint8_t *out, uint32_t nb) { return bp4x_compress(in, out, nb); }
size_t sbo_bp4x_decompress(const uint8_t *in, uint32_t *out, uint32_t nb) { return bp4x_decompress(in, out, nb); }
size_t sbo_bp4x_compress_sorted(uint32_t i0, const uint32_t *in, uint8_t *out, uint32_t nb) { return bp4x_compress_sorted(i0, in, out, nb); }
size_t sbo_bp4x_decompress_sorted(uint32_t i0, const uint8_t *in, uint32_t *out, uint32_t nb) { return bp4x_decompress_sorted(i0, in, out, nb); }
int sbo_roaring_serialize(const uint32_t *v, size_t n, sbo_buf *out) {
  Bytes b;
  roaring_serialize(std::vector<uint32_t>(v, v + n), b);
  export_buf(b, out);
  return SBO_OK;
}
int sbo_roaring_deserialize(const uint8_t *in, size_t len, sbo_buf *out) {
  std::vector<uint32_t> v;
  int rc = roaring_deserialize(in, len, v);
  if (rc) return rc;
  Bytes b(v.size() * 4);
  std::memcpy(b.data(), v.data(), b.size());
  export_buf(b, out);
  return SBO_OK;
}
int sbo_lz4_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) { return lz4_block_decode(in, in_len, out, out_len); }
int sbo_lz4_decompress_lib(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) { return common_decompress(SBO_C_LZ4, in, in_len, out, out_len); }
int sbo_lz4_compress_lib(const uint8_t *in, size_t in_len, sbo_buf *out) {
  Bytes b;
  size_t w;
  int rc = common_compress(SBO_C_LZ4, in, in_len, b, w);
  if (rc) return rc;
  export_buf(b, out);
  return SBO_OK;
}
int sbo_common_compress(int32_t codec, const uint8_t *in, size_t in_len, sbo_buf *out) {
  Bytes b;
  size_t w = 0;
  int rc = common_compress(codec, in, in_len, b, w);
  if (rc) return rc;
  export_buf(b, out);
  return SBO_OK;
}
int sbo_common_decompress(int32_t codec, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
  return common_decompress(codec, in, in_len, out, out_len);
}
uint16_t sbo_patas_pack(uint8_t r, uint8_t s, uint8_t t) { return patas_pack(r, s, t); }
void sbo_patas_unpack(uint16_t p, uint8_t *o) { patas_unpack(p, o[0], o[1], o[2]); }
int sbo_hybrid_rle_decode(const uint8_t *in, size_t len, uint32_t w, size_t n, uint32_t *out) {
  std::vector<uint32_t> v;
  int rc = hybrid_rle_decode(in, len, w, n, v);
  if (rc) return rc;
  std::memcpy(out, v.data(), n * 4);
  return SBO_OK;
}
int sbo_levels_encode(const uint32_t *levels, size_t n, uint32_t w, sbo_buf *out) {
  Bytes b;
  if (w) encode_u32_levels(levels, n, w, b);
  export_buf(b, out);
  return SBO_OK;
}
uint64_t sbo_sample_draw(uint64_t seed, uint32_t codec, uint32_t i, uint64_t range_end) { return sample_draw(seed, codec, i, range_end); }

} // extern "C"
