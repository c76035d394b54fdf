// sb_encode.cuh -- CTA-cooperative page encoders (src/compression/*: gen_stats, choose_compressor,
// compress_sample_ratio and every codec's `compress`; src/write/serialize.rs write_validity).
//
// One CTA (SB_NT threads) encodes one page into its slab.  Every routine is executed by all
// threads with uniform arguments and returns a uniform result.  Layouts: SURVEY.md App. A;
// chooser: App. B.  Deliberate, documented deviations from the reference:
//   * the sampler is the seeded sbo/sb `sample_draw` (the reference uses thread_rng),
//   * Freq's top value on ties = the key whose first occurrence is earliest (the reference
//     iterates a randomly seeded HashMap),
//   * floats are compared by bit pattern everywhere (App. C6), Patas is never chosen for f32
//     (App. C4: the reference's own f32 Patas stream is corrupt for repeated values).
#pragma once
#include "sb_common.cuh"

namespace sb {

struct EOpts {
  int32_t def_codec;  // SB_C_NONE / SB_C_LZ4
  double ratio;       // < 0: adaptive off
  uint32_t forbidden; // bit c: codec c forbidden
  int32_t force;      // -1 or codec id
  uint64_t seed;
};
enum { TC_UINT = 0, TC_SINT = 1, TC_FLOAT = 2 };
constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kEncFail = 0xffffffffu;

__device__ __forceinline__ bool e_forbidden(const EOpts &o, int c) { return (o.forbidden >> c) & 1u; }

// the deterministic stand-in for thread_rng().gen_range (integer/mod.rs:332); shared definition
// with the test oracle's sampler (same mixing constants), so chooser parity is testable
__device__ __forceinline__ uint64_t sample_draw(uint64_t seed, uint32_t codec, uint32_t i, uint64_t range_end) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t(codec) * 16 + i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return range_end ? z % range_end : 0;
}

// ------------------------------------------------------------------------------------
// views
// ------------------------------------------------------------------------------------
struct Vals { // W-byte little-endian elements, W-aligned
  const uint8_t *p;
  int W;
  __device__ __forceinline__ uint64_t get(uint32_t i) const {
    switch (W) {
    case 1: return p[i];
    case 2: return reinterpret_cast<const uint16_t *>(p)[i];
    case 4: return reinterpret_cast<const uint32_t *>(p)[i];
    default: return reinterpret_cast<const uint64_t *>(p)[i];
    }
  }
};
struct Bits { // LSB-first bitmap at a bit offset; p == nullptr: all ones
  const uint8_t *p;
  uint64_t off;
  __device__ __forceinline__ bool get(uint32_t i) const {
    if (!p) return true;
    uint64_t b = off + i;
    return (p[b >> 3] >> (b & 7)) & 1;
  }
};
__device__ __forceinline__ void st_le(uint8_t *out, uint64_t v, int nbytes) {
  for (int b = 0; b < nbytes; ++b) out[b] = uint8_t(v >> (8 * b));
}
__device__ __forceinline__ void put_hdr9(uint8_t *out, int codec, uint32_t compressed, uint32_t uncompressed) {
  if (threadIdx.x == 0) {
    out[0] = uint8_t(codec);
    st_le(out + 1, compressed, 4);
    st_le(out + 5, uncompressed, 4);
  }
}

// ------------------------------------------------------------------------------------
// block primitives
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t block_max_u64(Dctx &cx, uint64_t v) {
  __shared__ unsigned long long s_m[SB_NWARP];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    uint64_t o = __shfl_xor_sync(0xffffffffu, v, d);
    v = o > v ? o : v;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = v;
  __syncthreads();
  uint64_t r = 0;
#pragma unroll
  for (int w = 0; w < SB_NWARP; ++w) r = s_m[w] > r ? s_m[w] : r;
  return r;
}
__device__ __forceinline__ uint64_t block_sum_u64e(Dctx &cx, uint64_t v) {
  __shared__ unsigned long long s_s[SB_NWARP];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_s[threadIdx.x >> 5] = v;
  __syncthreads();
  uint64_t r = 0;
#pragma unroll
  for (int w = 0; w < SB_NWARP; ++w) r += s_s[w];
  return r;
}
// exclusive max-scan over the threads of the CTA (values are "index + 1", 0 = none)
__device__ __forceinline__ uint32_t block_excl_scan_max(uint32_t v, uint32_t *ws, uint32_t *total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= uint32_t(d)) inc = max(inc, t);
  }
  uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) exc = 0;
  __syncthreads();
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SB_NWARP; ++w) {
    uint32_t x = ws[w];
    if (uint32_t(w) < warp) base = max(base, x);
    tot = max(tot, x);
  }
  *total = tot;
  return max(base, exc);
}

// ------------------------------------------------------------------------------------
// null handling: src[i] = last valid row <= i, kNone for leading nulls.  RLE, Dict and Freq
// all replace a null slot by a neighbour (integer/rle.rs:92-95, integer/dict.rs:46-54).
// `row0_valid`: binary Dict treats row 0 as pushed even when null (binary/dict.rs:66-74).
// ------------------------------------------------------------------------------------
__device__ void fill_forward(Dctx &cx, const Bits &valid, uint32_t n, bool row0_valid, uint32_t *src) {
  constexpr uint32_t EPT = 8, CH = SB_NT * EPT;
  uint32_t carry = 0; // last valid index + 1 over previous chunks
  for (uint32_t c0 = 0; c0 < n; c0 += CH) {
    const uint32_t e0 = c0 + threadIdx.x * EPT;
    uint32_t loc[EPT], last = 0;
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j) {
      uint32_t i = e0 + j;
      if (i < n && (valid.get(i) || (row0_valid && i == 0))) last = i + 1;
      loc[j] = last;
    }
    uint32_t total;
    uint32_t pre = max(carry, block_excl_scan_max(last, cx.ws, &total));
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j) {
      uint32_t i = e0 + j;
      if (i < n) {
        uint32_t s = loc[j] ? loc[j] : pre;
        src[i] = s ? s - 1 : kNone;
      }
    }
    carry = max(carry, total);
  }
  __syncthreads();
}

// effective fixed-width value of a row after null replacement
struct EffFixed {
  Vals v;
  const uint32_t *src; // nullptr: no nulls
  uint64_t lead;       // value of leading null rows
  __device__ __forceinline__ uint64_t key(uint32_t i) const {
    if (!src) return v.get(i);
    uint32_t s = src[i];
    return s == kNone ? lead : v.get(s);
  }
  __device__ __forceinline__ uint32_t hash(uint32_t i) const {
    // doubles that hold small integers differ only in their top 20-odd bits: fold the halves before the
    // multiply and keep the product's HIGH bits, the ones every input bit reaches
    uint64_t k = key(i);
    k = (k ^ (k >> 32)) * 0x9E3779B97F4A7C15ull;
    k = (k ^ (k >> 29)) * 0xBF58476D1CE4E5B9ull;
    return uint32_t(k >> 32);
  }
  __device__ __forceinline__ bool equal(uint32_t i, uint32_t j) const { return key(i) == key(j); }
  __device__ __forceinline__ void put(uint8_t *out, uint32_t i, int W) const { st_le(out, key(i), W); }
};

// ------------------------------------------------------------------------------------
// exact distinct counting: open-addressing table of (representative row, count, first row).
// Returns the number of distinct keys, or kNone when it exceeds `limit` (the callers only need
// exact counts below n/3, see Dict / Freq ratios).  slot_of (optional) receives each row's slot.
// ------------------------------------------------------------------------------------
struct HashTab {
  uint32_t *rep, *cnt; // rep: 1 + the FIRST row holding the slot's key (0 = empty); cnt: rows with that key
  uint32_t mask;
  uint32_t cnt16;      // counts packed two per word (pages below 65536 rows): half the shared memory
  __device__ __forceinline__ uint32_t first(uint32_t h) const { return rep[h] - 1; }
  __device__ __forceinline__ uint32_t count(uint32_t h) const { return cnt16 ? (cnt[h >> 1] >> (16 * (h & 1))) & 0xffffu : cnt[h]; }
  __device__ __forceinline__ void add(uint32_t h, uint32_t c) {
    if (cnt16) atomicAdd(cnt + (h >> 1), c << (16 * (h & 1)));
    else atomicAdd(cnt + h, c);
  }
};
__device__ __forceinline__ uint32_t hash_cap(uint32_t limit) {
  uint64_t want = 2ull * limit + 2 * SB_NT + 16;
  uint32_t cap = 64;
  while (cap < want && cap < 0x80000000u) cap <<= 1;
  return cap;
}
// The table lives in shared memory when it fits (8192 slots x 8 bytes for an 8192-row page): every row costs a
// shared-memory probe, one key compare and -- aggregated over the lanes of a warp that landed on the same slot --
// one shared-memory atomic.
template <class Acc>
__device__ uint32_t hash_distinct(Dctx &cx, const Acc &acc, uint32_t n, uint32_t limit, HashTab *t, uint32_t *slot_of, bool compact = false) {
  const uint32_t cap = hash_cap(limit);
  t->cnt16 = compact && n < 65536u; // callers that reuse cnt[] as a slot -> id map keep full words
  t->rep = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(cap) * 4));
  t->cnt = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(cap) * (t->cnt16 ? 2 : 4)));
  t->mask = cap - 1;
  if (!t->rep || !t->cnt) {
    cx.flag(SB_NYI);
    return kNone;
  }
  for (uint32_t i = threadIdx.x; i < cap; i += SB_NT) {
    t->rep[i] = 0;
    if (!t->cnt16 || i < cap / 2) t->cnt[i] = 0;
  }
  volatile int *ctr = cx.bcast; // [0] distinct keys, [1] overflow flag
  __syncthreads();
  if (threadIdx.x == 0) cx.bcast[0] = cx.bcast[1] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  uint32_t h_next = threadIdx.x < n ? acc.hash(threadIdx.x) : 0u;
  for (uint32_t i0 = 0; i0 < n; i0 += SB_NT) { // warp-uniform trip count: the lanes vote on their slots below
    const uint32_t i = i0 + threadIdx.x;
    uint32_t h = kNone;
    const uint32_t h_cur = h_next;
    if (i + SB_NT < n) h_next = acc.hash(i + SB_NT); // the next row's key load travels under this row's probes
    if (i < n) {
      h = h_cur & t->mask;
      for (;;) {
        if (ctr[1]) {
          h = kNone;
          break;
        }
        uint32_t r = *reinterpret_cast<volatile uint32_t *>(t->rep + h);
        if (r == 0) {
          const uint32_t old = atomicCAS(t->rep + h, 0u, i + 1);
          if (old == 0) {
            r = i + 1;
            if (uint32_t(atomicAdd(cx.bcast, 1)) >= limit) atomicExch(cx.bcast + 1, 1);
          } else {
            r = old;
          }
        }
        if (r - 1 == i || acc.equal(r - 1, i)) {
          if (i + 1 < r) atomicMin(t->rep + h, i + 1); // the representative is replaced by an equal key: probes stay valid
          if (slot_of) slot_of[i] = h;
          break;
        }
        h = (h + 1) & t->mask;
      }
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, h);
    if (h != kNone && lane == uint32_t(__ffs(peers) - 1)) t->add(h, uint32_t(__popc(peers)));
  }
  __syncthreads();
  uint32_t distinct = uint32_t(cx.bcast[0]);
  bool ovf = cx.bcast[1] != 0;
  __syncthreads();
  return ovf ? kNone : distinct;
}
// slot of every row in a table built over the same keys (read-only probes)
template <class Acc> __device__ void hash_lookup(const Acc &acc, uint32_t n, const HashTab &t, uint32_t *slot_of) {
  for (uint32_t i = threadIdx.x; i < n; i += SB_NT) {
    uint32_t h = acc.hash(i) & t.mask;
    for (;;) {
      const uint32_t r = t.rep[h];
      if (r == 0 || r - 1 == i || acc.equal(r - 1, i)) break; // r == 0 cannot happen for a key of the build set
      h = (h + 1) & t.mask;
    }
    slot_of[i] = h;
  }
  __syncthreads();
}
// (max count, earliest first row) over the table
__device__ void hash_top(Dctx &cx, const HashTab &t, uint32_t *max_count, uint32_t *first_row) {
  uint64_t best = 0;
  for (uint32_t h = threadIdx.x; h <= t.mask; h += SB_NT)
    if (t.rep[h]) {
      uint64_t k = (uint64_t(t.count(h)) << 32) | uint64_t(kNone - t.first(h));
      best = k > best ? k : best;
    }
  best = block_max_u64(cx, best);
  *max_count = uint32_t(best >> 32);
  *first_row = kNone - uint32_t(best);
}

// ------------------------------------------------------------------------------------
// LZ4 block compressor (basic.rs:107-120 -> LZ4_compress_default).  Compressed bytes are
// implementation defined; any valid block decodes with the reference's LZ4_decompress_safe.
// Warp 0 runs a greedy hash-chain-less matcher, 32 candidate positions per step.
// Returns the compressed size (uniform over the CTA).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld4(const uint8_t *p) {
  return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}
__device__ __forceinline__ uint32_t lz4_len_bytes(uint32_t len) { return len < 15 ? 0 : 1 + (len - 15) / 255; } // extension bytes
// literal runs: bytes below 64, 16-byte vectors (aligned on dst) above
__device__ __forceinline__ void warp_copy(uint8_t *dst, const uint8_t *src, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31;
  if (n < 64) {
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
    return;
  }
  const uint32_t head = uint32_t((16 - (uintptr_t(dst) & 15)) & 15);
  if (lane < head) dst[lane] = src[lane];
  const uint32_t nv = (n - head) >> 4;
  for (uint32_t v = lane; v < nv; v += 32) *reinterpret_cast<uint4 *>(dst + head + 16 * v) = ld_u128u(src + head + 16 * v);
  for (uint32_t i = head + 16 * nv + lane; i < n; i += 32) dst[i] = src[i];
}
// One sequence, written by a warp.  The returned position is a pure function of the arguments, so the join phase
// of lz_compress_cta calls it from every warp with `hdr` true on one of them and copies the literals block-wide.
__device__ __forceinline__ uint32_t lz4_emit_seq(uint8_t *out, uint32_t op, const uint8_t *lit, uint32_t nlit, uint32_t offset, uint32_t ml /*0: last*/,
                                 bool hdr = true, bool copy_lit = true) {
  const uint32_t lane = hdr ? threadIdx.x & 31 : 1u; // lane 0 of the writing warp stores the token / length / offset bytes
  uint32_t mlc = ml ? ml - 4 : 0;
  if (lane == 0) {
    out[op] = uint8_t((min(nlit, 15u) << 4) | min(mlc, 15u));
    uint32_t q = op + 1;
    if (nlit >= 15) {
      uint32_t r = nlit - 15;
      while (r >= 255) {
        out[q++] = 255;
        r -= 255;
      }
      out[q++] = uint8_t(r);
    }
  }
  uint32_t q = op + 1 + lz4_len_bytes(nlit);
  if (copy_lit) warp_copy(out + q, lit, nlit);
  q += nlit;
  if (ml) {
    if (lane == 0) {
      out[q] = uint8_t(offset);
      out[q + 1] = uint8_t(offset >> 8);
      uint32_t z = q + 2;
      if (mlc >= 15) {
        uint32_t r = mlc - 15;
        while (r >= 255) {
          out[z++] = 255;
          r -= 255;
        }
        out[z++] = uint8_t(r);
      }
    }
    q += 2 + lz4_len_bytes(mlc);
  }
  __syncwarp();
  return q;
}
// Snappy raw elements for one (literal run, match) pair (basic.rs:138-152 -> snap::raw::Encoder; compressed bytes
// are implementation defined, any valid stream decodes with snap::raw::Decoder): a literal element, then copy
// elements with a 16-bit offset (1..64 bytes each; 11-bit form for 4..11 bytes at offsets < 2048).
__device__ __forceinline__ uint32_t snappy_emit_seq(uint8_t *out, uint32_t op, const uint8_t *lit, uint32_t nlit, uint32_t offset, uint32_t ml /*0: last*/,
                                    bool hdr = true, bool copy_lit = true) {
  const uint32_t lane = hdr ? threadIdx.x & 31 : 1u;
  uint32_t q = op;
  if (nlit) {
    const uint32_t l1 = nlit - 1;
    const uint32_t nb = l1 < 60 ? 0u : l1 < (1u << 8) ? 1u : l1 < (1u << 16) ? 2u : l1 < (1u << 24) ? 3u : 4u;
    if (lane == 0) {
      out[q] = uint8_t(nb ? (59 + nb) << 2 : l1 << 2);
      for (uint32_t k = 0; k < nb; ++k) out[q + 1 + k] = uint8_t(l1 >> (8 * k));
    }
    q += 1 + nb;
    if (copy_lit) warp_copy(out + q, lit, nlit);
    q += nlit;
  }
  // copy elements: their sizes are a pure function of (offset, ml), so every lane tracks q; lane 0 writes
  for (uint32_t left = ml; left;) {
    const uint32_t l = left > 64 ? (left - 64 < 4 ? 60u : 64u) : left;
    if (l >= 4 && l <= 11 && offset < 2048) {
      if (lane == 0) {
        out[q] = uint8_t(1u | ((l - 4) << 2) | ((offset >> 8) << 5));
        out[q + 1] = uint8_t(offset);
      }
      q += 2;
    } else {
      if (lane == 0) {
        out[q] = uint8_t(2u | ((l - 1) << 2));
        out[q + 1] = uint8_t(offset);
        out[q + 2] = uint8_t(offset >> 8);
      }
      q += 3;
    }
    left -= l;
  }
  __syncwarp();
  return q;
}
// ------------------------------------------------------------------------------------
// Greedy LZ matcher shared by the LZ4 and the Snappy writer (the element syntax is the only difference).
// The input is cut into up to SB_NWARP chunks, one warp each with its own hash table (positions of its chunk
// only, so the result does not depend on warp timing).  Per step a warp probes 32 consecutive positions, every
// lane with a verified candidate extends its own match (8 bytes per compare, capped), and the warp then walks
// the hits left to right taking each one that starts at or after the end of the previous match -- several
// sequences per step on data with short matches (f64 / i64 with a few significant bytes), one cooperative
// 32-bytes-per-step extension on long ones.
// Chunk 0 writes its sequences in place.  A later chunk w does not know where its bytes land, nor where the
// literal run before its first match starts (the previous chunks' tail): it parks its first match in shared
// memory, writes the rest at a temporary offset at or beyond the final one, and the join phase emits the
// junction sequences and moves the bodies left.
// ------------------------------------------------------------------------------------
struct LzChunk {
  uint32_t has_first, mp1, mc1, ml1; // first match of the chunk (chunks 1..: emitted by the join phase)
  uint32_t body_off, body_len;       // sequences after the first one, where the warp wrote them
  uint32_t tail;                     // first byte after the chunk's last match (its trailing literals)
  uint32_t pad;
};
template <bool SNAPPY> __device__ __forceinline__ uint32_t lz_slack(uint32_t len) { return SNAPPY ? len / 32 + 32 : len / 255 + 16; }
template <bool SNAPPY>
__device__ __forceinline__ uint32_t lz_emit(uint8_t *out, uint32_t op, const uint8_t *lit, uint32_t nlit, uint32_t offset, uint32_t ml, bool hdr = true,
                                            bool copy_lit = true) {
  return SNAPPY ? snappy_emit_seq(out, op, lit, nlit, offset, ml, hdr, copy_lit) : lz4_emit_seq(out, op, lit, nlit, offset, ml, hdr, copy_lit);
}
// the same sequence written by the whole CTA: warp 0 stores the few header bytes, everyone copies the literals
template <bool SNAPPY>
__device__ uint32_t lz_emit_cta(uint8_t *out, uint32_t op, const uint8_t *lit, uint32_t nlit, uint32_t offset, uint32_t ml) {
  uint32_t qlit;
  if (SNAPPY) {
    const uint32_t l1 = nlit ? nlit - 1 : 0;
    qlit = op + (nlit ? 1 + (l1 < 60 ? 0u : l1 < (1u << 8) ? 1u : l1 < (1u << 16) ? 2u : l1 < (1u << 24) ? 3u : 4u) : 0);
  } else {
    qlit = op + 1 + lz4_len_bytes(nlit);
  }
  const uint32_t q = lz_emit<SNAPPY>(out, op, lit, nlit, offset, ml, threadIdx.x < 32, false);
  copy_bytes(out + qlit, lit, nlit);
  return q;
}
constexpr uint32_t kLzMinChunk = 2048;
constexpr uint32_t kStageMax = 64 * 1024; // a page of 8192 eight-byte values
constexpr uint32_t kStageKeep = 48 * 1024 + 512; // the distinct table of an 8192-row page (32 + 16 KiB)
constexpr uint32_t kLzLaneCap = 36; // per-lane match extension stops here; longer matches continue warp-wide

template <bool SNAPPY> __device__ uint32_t lz_compress_cta(Dctx &cx, const uint8_t *in, uint32_t n, uint8_t *out) {
  Arena mark = cx.ar;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t cl = ((n + SB_NWARP - 1) / SB_NWARP + 63) & ~63u;
  cl = cl < kLzMinChunk ? kLzMinChunk : cl;
  const uint32_t n_chunks = n ? (n + cl - 1) / cl : 1; // <= SB_NWARP
  uint32_t bits = 11;
  uint32_t *tabs = nullptr;
  while (bits >= 9 && !(tabs = static_cast<uint32_t *>(cx.ar.alloc_shared(uint64_t(n_chunks) << (bits + 2))))) --bits;
  if (!tabs) {
    bits = 11;
    tabs = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n_chunks) << (bits + 2)));
  }
  LzChunk *recs = static_cast<LzChunk *>(cx.ar.alloc(sizeof(LzChunk) * SB_NWARP));
  if (!tabs || !recs) {
    cx.flag(SB_NYI);
    return kEncFail;
  }
  uint32_t op0 = 0;
  if (SNAPPY) { // preamble: varint of the uncompressed length
    uint32_t v = n;
    do {
      if (threadIdx.x == 0) out[op0] = uint8_t((v & 0x7fu) | (v >> 7 ? 0x80u : 0u));
      ++op0;
      v >>= 7;
    } while (v);
  }
  const uint32_t mflimit = n >= 13 ? n - 12 : 0;        // last match start (inclusive) -- LZ4 block end rules
  const uint32_t match_end_limit = n >= 5 ? n - 5 : 0;  // matches end before the last 5 bytes
  if (warp < n_chunks) {
    uint32_t *tab = tabs + (size_t(warp) << bits);
    for (uint32_t i = lane; i < (1u << bits); i += 32) tab[i] = 0;
    __syncwarp();
    const uint32_t c0 = warp * cl, c1 = min(n, c0 + cl);
    const uint32_t start_end = n >= 13 ? min(c1, mflimit + 1) : c0; // match starts: [c0, start_end)
    const uint32_t mend = min(c1, match_end_limit);                  // match bytes stay below mend (and inside the chunk)
    uint32_t ip = c0, anchor = c0;
    bool first = warp != 0;
    uint32_t op = op0, body_off = op0;
    LzChunk rec{};
    while (ip < start_end) {
      const uint32_t p = ip + lane;
      const bool in_range = p < start_end && p + 4 <= mend;
      uint32_t seq = 0, cand = kNone;
      if (in_range) seq = ld_u32u(in + p);
      {
        // every lane reads its slot, then the slot takes the LAST position among the lanes with that hash (atomic max
        // of position + 1, 0 = empty): the table, hence the output, never depends on the order of a warp's stores
        const uint32_t h = (seq * 2654435761u) >> (32 - bits);
        if (in_range) cand = tab[h] - 1u; // empty -> kNone
        __syncwarp();
        if (in_range) atomicMax(tab + h, p + 1u);
        __syncwarp();
      }
      bool hit = in_range && cand != kNone && p - cand <= 65535u && ld_u32u(in + cand) == seq;
      uint32_t ml = 0;
      if (hit) { // own extension, 8 bytes per compare
        const uint32_t lim = min(mend - p, kLzLaneCap);
        ml = 4;
        bool open = true;
        while (open && ml + 8 <= lim) {
          const uint64_t x = ld_u64u(in + p + ml) ^ ld_u64u(in + cand + ml);
          if (x) {
            ml += uint32_t(__ffsll((long long)x) - 1) >> 3;
            open = false;
          } else {
            ml += 8;
          }
        }
        while (open && ml < lim && in[p + ml] == in[cand + ml]) ++ml;
      }
      const uint32_t hits = __ballot_sync(0xffffffffu, hit);
      // ---- take the hits left to right, each one that starts at or after the end of the previous match.  LZ4: the
      //      walk also hands every taken lane its output position (sequence sizes are uniform arithmetic), the
      //      lanes then write their sequences side by side.
      uint32_t cur = ip, sel = 0, my_anchor = 0, my_len = ml, my_q = 0;
      while (cur - ip < 32) {
        const uint32_t m = hits & (0xffffffffu << (cur - ip));
        if (!m) break;
        const uint32_t f = __ffs(m) - 1;
        const uint32_t mp = ip + f;
        uint32_t len = __shfl_sync(0xffffffffu, ml, f);
        if (len == kLzLaneCap) { // still open: 32 bytes per step
          const uint32_t mc = __shfl_sync(0xffffffffu, cand, f);
          for (;;) {
            const uint32_t a = mp + len + lane;
            const bool same = a < mend && in[a] == in[mc + len + lane];
            const uint32_t e = __ballot_sync(0xffffffffu, same);
            const uint32_t run = e == 0xffffffffu ? 32 : __ffs(~e) - 1;
            len += run;
            if (run < 32) break;
          }
        }
        if (first) { // the chunk's first match is emitted by the join phase, behind the previous chunks' tail
          first = false;
          rec.has_first = 1, rec.mp1 = mp, rec.mc1 = __shfl_sync(0xffffffffu, cand, f), rec.ml1 = len;
          const uint32_t lead = mp - c0;
          body_off = op0 + warp * (cl + lz_slack<SNAPPY>(cl) + 96) + lead + lz_slack<SNAPPY>(lead) + lz_slack<SNAPPY>(len) + 32;
          op = body_off;
        } else if constexpr (SNAPPY) { // element sizes depend on (offset, length): written by the warp, one at a time
          op = snappy_emit_seq(out, op, in + anchor, mp - anchor, mp - __shfl_sync(0xffffffffu, cand, f), len);
        } else {
          const uint32_t nl = mp - anchor;
          if (lane == f) my_anchor = anchor, my_len = len, my_q = op;
          op += 1 + lz4_len_bytes(nl) + nl + 2 + lz4_len_bytes(len - 4);
          sel |= 1u << f;
        }
        cur = mp + len;
        anchor = cur;
      }
      if (!SNAPPY && sel) {
        const bool mine = (sel >> lane) & 1u;
        const uint32_t nlit = mine ? p - my_anchor : 0u, mlc = my_len - 4;
        uint32_t q = my_q, lit_dst = 0;
        if (mine) {
          out[q++] = uint8_t((min(nlit, 15u) << 4) | min(mlc, 15u));
          if (nlit >= 15) {
            uint32_t r = nlit - 15;
            for (; r >= 255; r -= 255) out[q++] = 255;
            out[q++] = uint8_t(r);
          }
          lit_dst = q;
          if (nlit <= 16)
            for (uint32_t k = 0; k < nlit; ++k) out[q + k] = in[my_anchor + k];
          q += nlit;
          const uint32_t offset = p - cand;
          out[q] = uint8_t(offset), out[q + 1] = uint8_t(offset >> 8);
          q += 2;
          if (mlc >= 15) {
            uint32_t r = mlc - 15;
            for (; r >= 255; r -= 255) out[q++] = 255;
            out[q++] = uint8_t(r);
          }
        }
        for (uint32_t big = __ballot_sync(0xffffffffu, mine && nlit > 16); big; big &= big - 1) { // long literal runs: the warp copies
          const uint32_t f = __ffs(big) - 1;
          warp_copy(out + __shfl_sync(0xffffffffu, lit_dst, f), in + __shfl_sync(0xffffffffu, my_anchor, f), __shfl_sync(0xffffffffu, nlit, f));
        }
      }
      ip = max(ip + 32, cur);
    }
    if (lane == 0) {
      rec.body_off = body_off, rec.body_len = op - body_off;
      rec.tail = anchor; // chunk without a match (has_first == 0 on chunks 1..): tail is unused, the run carries on
      recs[warp] = rec;
    }
  }
  __syncthreads();
  // ---- join: junction sequences, bodies moved left, final literals
  uint32_t op = recs[0].body_off + recs[0].body_len, anchor = recs[0].tail;
  for (uint32_t w = 1; w < n_chunks; ++w) {
    const LzChunk r = recs[w];
    if (!r.has_first) continue;
    op = lz_emit_cta<SNAPPY>(out, op, in + anchor, r.mp1 - anchor, r.mp1 - r.mc1, r.ml1);
    if (op > r.body_off) cx.flag(SB_PANIC); // the temporary offset is an upper bound of the final one by construction
    else if (op < r.body_off) {
      uint8_t *dst = out + op;
      const uint8_t *src = out + r.body_off;
      const int64_t len = r.body_len;
      // 16-byte slots aligned on dst; the regions may overlap (dst < src): load, barrier, store
      for (int64_t o = -int64_t(uintptr_t(dst) & 15); o < len; o += SB_NT * 16) {
        const int64_t i = o + int64_t(threadIdx.x) * 16;
        const bool full = i >= 0 && i + 16 <= len;
        const int64_t lo = i < 0 ? 0 : i, hi = i + 16 < len ? i + 16 : len; // partial slot: [lo, hi)
        uint4 v = make_uint4(0, 0, 0, 0);
        uint8_t edge[16];
        if (full) v = ld_u128u(src + i);
        else
          for (int64_t k = lo; k < hi; ++k) edge[k - lo] = src[k];
        __syncthreads();
        if (full) *reinterpret_cast<uint4 *>(dst + i) = v;
        else
          for (int64_t k = lo; k < hi; ++k) dst[k] = edge[k - lo];
        __syncthreads();
      }
    }
    __syncthreads();
    op += r.body_len;
    anchor = r.tail;
  }
  const uint32_t total = lz_emit_cta<SNAPPY>(out, op, in + anchor, n - anchor, 0, 0);
  __syncthreads();
  cx.ar = mark;
  return total;
}
__host__ __device__ __forceinline__ uint64_t lz4_bound(uint64_t n) { return n + n / 255 + 16; }

// CommonCompression::compress (basic.rs:74-120) of `n` bytes -> out; returns written bytes
__device__ uint32_t enc_basic(Dctx &cx, int codec, const uint8_t *in, uint32_t n, uint8_t *out) {
  if (codec == SB_C_NONE) {
    copy_bytes(out, in, n);
    return n;
  }
  if (codec == SB_C_LZ4) return lz_compress_cta<false>(cx, in, n, out);
  if (codec == SB_C_SNAPPY) return lz_compress_cta<true>(cx, in, n, out);
  if (codec == SB_C_ZSTD) {
    // A valid Zstandard frame that any decoder (zstd::bulk::decompress_to_buffer, basic.rs:93-97) reads: single
    // segment, 4-byte frame content size, RAW blocks of at most 128 KiB (RFC 8878 3.1.1.2.2).  Stored, not
    // compressed: the entropy coder (FSE / Huffman writer) is not part of this build -- pages still shrink through
    // the adaptive codecs above the common codec; files are readable by the reference, only larger than libzstd's.
    constexpr uint32_t BLK = 128u << 10;
    const uint32_t nblk = n ? (n + BLK - 1) / BLK : 1;
    if (threadIdx.x == 0) {
      out[0] = 0x28, out[1] = 0xB5, out[2] = 0x2F, out[3] = 0xFD;
      out[4] = 0xA0; // FCS flag 2 (4 bytes), single segment, no checksum, no dictionary
      st_le(out + 5, n, 4);
      for (uint32_t b = 0; b < nblk; ++b) {
        const uint32_t sz = min(BLK, n - b * BLK);
        st_le(out + 9 + uint64_t(b) * (BLK + 3), (sz << 3) | (b + 1 == nblk ? 1u : 0u), 3);
      }
    }
    for (uint32_t b = 0; b < nblk; ++b) copy_bytes(out + 9 + uint64_t(b) * (BLK + 3) + 3, in + uint64_t(b) * BLK, min(BLK, n - b * BLK));
    return 9 + 3 * nblk + n;
  }
  cx.flag(SB_OUT_OF_SPEC);
  return kEncFail;
}

// ------------------------------------------------------------------------------------
// RLE (integer/rle.rs:64-104): runs over the effective values.  count_only: sample ratio.
// ------------------------------------------------------------------------------------
template <class Acc>
__device__ uint32_t enc_rle(Dctx &cx, const Acc &acc, uint32_t n, int W, uint8_t *out, bool count_only) {
  constexpr uint32_t EPT = 4, CH = SB_NT * EPT;
  Arena mark = cx.ar;
  uint32_t *starts = nullptr;
  if (!count_only) {
    starts = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n + 1) * 4));
    if (!starts) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
  }
  uint32_t nruns = 0;
  for (uint32_t c0 = 0; c0 < n; c0 += CH) {
    const uint32_t e0 = c0 + threadIdx.x * EPT;
    uint32_t flags = 0, c = 0;
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j) {
      uint32_t i = e0 + j;
      if (i < n && (i == 0 || !acc.equal(i, i - 1))) {
        flags |= 1u << j;
        ++c;
      }
    }
    uint32_t total;
    uint32_t pre = nruns + block_excl_scan(c, cx.ws, &total);
    if (!count_only) {
#pragma unroll
      for (uint32_t j = 0; j < EPT; ++j)
        if ((flags >> j) & 1u) starts[pre++] = e0 + j;
    }
    nruns += total;
  }
  if (!count_only) {
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < nruns; r += SB_NT) {
      uint32_t s = starts[r], e = r + 1 < nruns ? starts[r + 1] : n;
      uint8_t *o = out + uint64_t(r) * (4 + W);
      st_le(o, e - s, 4);
      acc.put(o + 4, s, W);
    }
    __syncthreads();
  }
  cx.ar = mark;
  return nruns * uint32_t(4 + W);
}

// ------------------------------------------------------------------------------------
// BitPacker4x (integer/bp.rs:36-65, delta_bp.rs:36-67; layout SURVEY App. D.1).  One warp
// per 128-value block; lane t owns positions t of the 4 BitPacker lanes = values 4t..4t+3.
// size_only: sample ratio (bp.rs:93-101).  The width byte is the width of the RAW values
// also for the delta variant (App. C2).
// ------------------------------------------------------------------------------------
__device__ uint32_t enc_bitpack(Dctx &cx, const uint32_t *v, uint32_t n, bool delta, uint8_t *out, bool size_only) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nblk = n >> 7;
  Arena mark = cx.ar;
  uint32_t *bpos = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(nblk + 1) * 4));
  uint32_t *stage = static_cast<uint32_t *>(cx.ar.alloc_shared(SB_NWARP * 128 * 4));
  if (!stage) stage = static_cast<uint32_t *>(cx.ar.alloc(SB_NWARP * 128 * 4));
  if (!bpos || !stage) {
    cx.flag(SB_NYI);
    return kEncFail;
  }
  for (uint32_t b = warp; b < nblk; b += SB_NWARP) {
    const uint32_t *vb = v + b * 128 + 4 * lane; // 4-byte aligned only (page starts are arbitrary rows)
    uint32_t acc = vb[0] | vb[1] | vb[2] | vb[3];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) bpos[b] = 1 + 16 * (acc ? 32 - __clz(acc) : 0);
  }
  __syncthreads();
  // exclusive scan of block sizes
  uint32_t carry = 0;
  for (uint32_t b0 = 0; b0 < nblk; b0 += SB_NT) {
    uint32_t b = b0 + threadIdx.x;
    uint32_t sz = b < nblk ? bpos[b] : 0, total;
    uint32_t ex = block_excl_scan(sz, cx.ws, &total);
    __syncthreads();
    if (b < nblk) bpos[b] = carry + ex;
    carry += total;
  }
  __syncthreads();
  if (!size_only) {
    uint32_t *st = stage + warp * 128;
    for (uint32_t b = warp; b < nblk; b += SB_NWARP) {
      uint32_t pos = bpos[b], end = b + 1 < nblk ? bpos[b + 1] : carry;
      uint32_t bits = (end - pos - 1) / 16;
      const uint32_t *vb = v + b * 128 + 4 * lane;
      uint4 q = make_uint4(vb[0], vb[1], vb[2], vb[3]);
      if (delta) { // wrapping deltas against the previous value (compress_sorted)
        uint32_t prev_w = __shfl_up_sync(0xffffffffu, q.w, 1);
        if (lane == 0) prev_w = b ? v[b * 128 - 1] : 0u;
        uint4 d;
        d.x = q.x - prev_w, d.y = q.y - q.x, d.z = q.z - q.y, d.w = q.w - q.z;
        q = d;
      }
      for (uint32_t i = lane; i < 128; i += 32) st[i] = 0;
      __syncwarp();
      if (bits) {
        uint32_t bit = lane * bits, w = bit >> 5, s = bit & 31;
        uint32_t m = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
        uint32_t vv[4] = {q.x & m, q.y & m, q.z & m, q.w & m};
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          atomicOr(st + 4 * w + l, vv[l] << s);
          if (s + bits > 32) atomicOr(st + 4 * (w + 1) + l, vv[l] >> (32 - s));
        }
      }
      __syncwarp();
      uint8_t *o = out + pos;
      if (lane == 0) o[0] = uint8_t(bits);
      for (uint32_t i = lane; i < 4 * bits; i += 32) st_le(o + 1 + 4 * i, st[i], 4);
      __syncwarp();
    }
    __syncthreads();
  }
  cx.ar = mark;
  return carry;
}

// ------------------------------------------------------------------------------------
// Patas (double/patas.rs:36-105), f64 only.  Reference index = last earlier identical value
// when it is < 128 back, else index 0 while i < 128, else i-1 (App. C5).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t patas_one(const uint64_t *v, uint32_t i, uint32_t *hdr, uint64_t *payload) {
  uint64_t val = v[i];
  uint32_t ref = kNone;
  uint32_t lo = i >= 127 ? i - 127 : 0;
  for (uint32_t j = i; j-- > lo;)
    if (v[j] == val) {
      ref = j;
      break;
    }
  if (ref == kNone) ref = i < 128 ? 0 : i - 1;
  uint32_t diff = i - ref;
  uint64_t x = val ^ v[ref];
  uint32_t tz = x ? uint32_t(__ffsll((long long)x) - 1) : 64, lz = x ? uint32_t(__clzll((long long)x)) : 64;
  uint32_t eq = tz == 64;
  uint32_t sig_bits = eq ? 0 : 64 - tz - lz;
  uint32_t sig_bytes = (sig_bits >> 3) + ((sig_bits & 7) != 0);
  *hdr = (diff << 9) | ((sig_bytes & 7) << 6) | (tz - eq);
  *payload = x >> (tz - eq);
  return sig_bytes;
}
__device__ uint32_t enc_patas(Dctx &cx, const uint64_t *v, uint32_t n, uint8_t *out, bool size_only) {
  constexpr uint32_t EPT = 4, CH = SB_NT * EPT;
  if (n == 0) return 0;
  if (!size_only && threadIdx.x == 0) st_le(out, v[0], 8);
  uint32_t pos = 8;
  for (uint32_t c0 = 1; c0 < n; c0 += CH) {
    const uint32_t e0 = c0 + threadIdx.x * EPT;
    uint32_t hdr[EPT], sb[EPT], sz = 0;
    uint64_t pay[EPT];
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j) {
      sb[j] = 0;
      if (e0 + j < n) {
        sb[j] = patas_one(v, e0 + j, &hdr[j], &pay[j]);
        sz += 2 + sb[j];
      }
    }
    uint32_t total;
    uint32_t p = pos + block_excl_scan(sz, cx.ws, &total);
    if (!size_only) {
#pragma unroll
      for (uint32_t j = 0; j < EPT; ++j)
        if (e0 + j < n) {
          st_le(out + p, hdr[j], 2);
          st_le(out + p + 2, pay[j], int(sb[j]));
          p += 2 + sb[j];
        }
    }
    pos += total;
  }
  __syncthreads();
  return pos;
}

// ------------------------------------------------------------------------------------
// Roaring portable serialisation of sorted row numbers: array containers
// up to 4096 entries, bitmap containers above.  Returns bytes written.
// ------------------------------------------------------------------------------------
__device__ uint32_t enc_roaring(Dctx &cx, const uint32_t *rows, uint32_t cnt, uint32_t n_rows, uint8_t *out) {
  const uint32_t tid = threadIdx.x;
  const uint32_t nkeys = cnt ? (rows[cnt - 1] >> 16) + 1 : 0;
  (void)n_rows;
  // uniform pass over keys: lower bounds by binary search (every thread computes the same)
  uint32_t ncont = 0;
  auto lower = [&](uint32_t key) {
    uint32_t lo = 0, hi = cnt, tgt = key << 16;
    if (key >= 65536) return cnt;
    while (lo < hi) {
      uint32_t mid = (lo + hi) >> 1;
      if (rows[mid] < tgt) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  };
  for (uint32_t k = 0; k < nkeys; ++k) ncont += lower(k + 1) > lower(k);
  uint32_t off = 8 + 8 * ncont, c = 0;
  if (tid == 0) {
    st_le(out, 12346u, 4);
    st_le(out + 4, ncont, 4);
  }
  for (uint32_t k = 0; k < nkeys; ++k) {
    uint32_t lo = lower(k), hi = lower(k + 1), card = hi - lo;
    if (!card) continue;
    if (tid == 0) {
      st_le(out + 8 + 4 * c, k, 2);
      st_le(out + 8 + 4 * c + 2, card - 1, 2);
      st_le(out + 8 + 4 * ncont + 4 * c, off, 4);
    }
    uint8_t *d = out + off;
    if (card > 4096) {
      for (uint32_t i = tid; i < 8192; i += SB_NT) d[i] = 0;
      __syncthreads();
      // bits of one byte may come from several threads: build 32-bit words through a
      // word-aligned view when possible, else byte-wise atomics are unavailable -> serialise per word
      for (uint32_t w = tid; w < 2048; w += SB_NT) { // 2048 32-bit words, each owned by one thread
        uint32_t first = (k << 16) | (w << 5);
        uint32_t a = lo, b = hi; // lower bound of `first` inside [lo, hi)
        while (a < b) {
          uint32_t mid = (a + b) >> 1;
          if (rows[mid] < first) a = mid + 1;
          else b = mid;
        }
        uint32_t word = 0;
        while (a < hi && rows[a] < first + 32) word |= 1u << (rows[a++] & 31);
        st_le(d + 4 * w, word, 4);
      }
      off += 8192;
    } else {
      for (uint32_t i = tid; i < card; i += SB_NT) st_le(d + 2 * i, rows[lo + i] & 0xffffu, 2);
      off += 2 * card;
    }
    ++c;
  }
  __syncthreads();
  return off;
}

// ------------------------------------------------------------------------------------
// sampling (compress_sample_ratio, integer/mod.rs:310-347): materialises the sample rows
// into arena buffers (values, validity bitmap at bit offset 0).  Returns the sample length.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool sample_whole(uint32_t n) { return n / 10 <= 64; }
__device__ __forceinline__ uint32_t sample_row(uint32_t n, const EOpts &o, int codec, uint32_t k) {
  uint32_t sep = n / 10, part = k / 64, rem = n % 10;
  uint64_t range_end = uint64_t(part == 9 ? sep + rem : sep) - 64;
  return part * sep + uint32_t(sample_draw(o.seed, uint32_t(codec), part, range_end)) + (k & 63);
}

// ------------------------------------------------------------------------------------
// compress_integer / compress_double (integer/mod.rs:35-70, double/mod.rs:32-67):
// hdr9 + payload into `out`; returns total bytes or kEncFail.
// ------------------------------------------------------------------------------------
struct FixedStats {
  uint32_t n, null_count, unique; // unique == kNone: more than n/3 distinct
  uint32_t max_count, top_first;
  bool sorted_nonnull, has_neg, max_ge_256;
};

template <int LEVEL>
__device__ uint32_t enc_fixed(Dctx &cx, Vals v, int tclass, Bits valid, uint32_t n, EOpts o, uint8_t *out);

__device__ double fixed_sample_ratio(Dctx &cx, const Vals &v, const Bits &valid, uint32_t n, int codec, const EOpts &o, uint32_t null_count) {
  // sample into arena: values (W bytes each) + validity
  Arena mark = cx.ar;
  const int W = v.W;
  uint32_t m = sample_whole(n) ? n : 640;
  uint8_t *sv = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(m) * W + 16));
  uint8_t *sb = static_cast<uint8_t *>(cx.ar.alloc((m + 7) / 8 + 16));
  uint32_t *src = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(m) * 4 + 16));
  if (!sv || !sb || !src) {
    cx.flag(SB_NYI);
    return 0.0;
  }
  for (uint32_t i = threadIdx.x; i < (m + 7) / 8; i += SB_NT) {
    uint32_t byte = 0;
    for (uint32_t b = 0; b < 8 && 8 * i + b < m; ++b) {
      uint32_t r = sample_whole(n) ? 8 * i + b : sample_row(n, o, codec, 8 * i + b);
      byte |= uint32_t(valid.get(r)) << b;
    }
    sb[i] = uint8_t(byte);
  }
  for (uint32_t k = threadIdx.x; k < m; k += SB_NT) {
    uint32_t r = sample_whole(n) ? k : sample_row(n, o, codec, k);
    uint64_t x = v.get(r);
    switch (W) {
    case 1: sv[k] = uint8_t(x); break;
    case 2: reinterpret_cast<uint16_t *>(sv)[k] = uint16_t(x); break;
    case 4: reinterpret_cast<uint32_t *>(sv)[k] = uint32_t(x); break;
    default: reinterpret_cast<uint64_t *>(sv)[k] = x; break;
    }
  }
  __syncthreads();
  Vals s{sv, W};
  uint32_t size = 0;
  if (codec == SB_C_RLE) {
    EffFixed acc{s, nullptr, 0};
    if (valid.p && null_count) {
      Bits sbits{sb, 0};
      fill_forward(cx, sbits, m, false, src);
      // leading nulls join the first valid run; all null: one run of T::default()
      __syncthreads();
      if (threadIdx.x == 0) cx.bcast[2] = int(kNone);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < m; i += SB_NT)
        if (sbits.get(i)) {
          atomicMin(reinterpret_cast<uint32_t *>(cx.bcast + 2), i);
          break;
        }
      __syncthreads();
      uint32_t fv = uint32_t(cx.bcast[2]);
      __syncthreads();
      acc.src = src;
      acc.lead = fv == kNone ? 0 : s.get(fv);
    }
    size = enc_rle(cx, acc, m, W, nullptr, true);
  } else if (codec == SB_C_BITPACK) {
    size = enc_bitpack(cx, reinterpret_cast<const uint32_t *>(sv), m, false, nullptr, true);
  } else if (codec == SB_C_PATAS) {
    size = enc_patas(cx, reinterpret_cast<const uint64_t *>(sv), m, nullptr, true);
  }
  cx.ar = mark;
  if (size == kEncFail) return 0.0;
  return double(uint64_t(m) * W) / double(size);
}

__device__ __forceinline__ uint32_t bits_needed(uint32_t x) { return x ? 32 - __clz(x) : 0; } // get_bits_needed

template <int LEVEL>
__device__ uint32_t enc_fixed(Dctx &cx, Vals v, int tclass, Bits valid, uint32_t n, EOpts o, uint8_t *out) {
  const int W = v.W;
  const uint32_t tid = threadIdx.x;
  const Arena entry = cx.ar;
  if constexpr (LEVEL == 0) {
    // The page is read by the statistics pass, the distinct table, the samples and the chosen codec (LZ4: by
    // every probe of the matcher): one copy into shared memory, every later read at shared-memory latency.
    const uint64_t bytes = uint64_t(n) * W;
    if (bytes && bytes <= kStageMax && bytes + 16 + kStageKeep <= uint64_t(cx.ar.s_end - cx.ar.s_cur)) { // room for the tables must remain
      uint8_t *stage = static_cast<uint8_t *>(cx.ar.alloc_shared(bytes + 16));
      if (stage) {
        copy_bytes(stage, v.p, bytes);
        __syncthreads();
        v.p = stage;
      }
    }
  }
  Arena mark = cx.ar;
  uint8_t *body = out + 9;

  // ---- gen_stats (integer/mod.rs:179-229): one pass + the distinct table over ALL slots
  FixedStats st{};
  st.n = n;
  // chooser off and nothing forced: the page goes to the default codec, the statistics have no reader
  const bool need_stats = (o.ratio >= 0 && n) || o.force >= SB_C_RLE;
  if (need_stats) {
    uint32_t nulls = 0;
    bool sorted = true, neg = false;
    uint64_t mx = 0;
    const uint64_t sign = 1ull << (8 * W - 1);
#pragma unroll 4
    for (uint32_t i = tid; i < n; i += SB_NT) {
      uint64_t x = v.get(i);
      nulls += !valid.get(i);
      if (tclass != TC_FLOAT) {
        uint64_t ord = tclass == TC_SINT ? x ^ sign : x; // order-preserving map to unsigned
        if (i) {
          uint64_t p = v.get(i - 1);
          if (ord < (tclass == TC_SINT ? p ^ sign : p)) sorted = false;
        }
        if (tclass == TC_SINT && (x & sign)) neg = true;
        mx = ord > mx ? ord : mx;
      }
    }
    st.null_count = uint32_t(block_sum_u64e(cx, nulls));
    st.sorted_nonnull = !__syncthreads_or(!sorted);
    st.has_neg = __syncthreads_or(neg);
    mx = block_max_u64(cx, mx);
    if (tclass == TC_SINT) mx ^= sign;
    // `max as i64 >= 256` (integer/freq.rs:146): u64 above i64::MAX wraps negative; narrower signed types sign-extend
    int64_t as_i64 = tclass == TC_SINT ? (int64_t(mx << (64 - 8 * W)) >> (64 - 8 * W)) : int64_t(mx);
    st.max_ge_256 = as_i64 >= 256;
  }
  EffFixed raw{v, nullptr, 0};
  HashTab tab{};
  const uint32_t limit = n / 3 + 1;
  st.unique = !need_stats ? kNone : n ? hash_distinct(cx, raw, n, limit, &tab, nullptr, true) : 0;
  if (*cx.err) return kEncFail;
  st.max_count = 0;
  st.top_first = 0;
  if (st.unique != kNone && n) hash_top(cx, tab, &st.max_count, &st.top_first);
  // the table stays for now: Dict over a page without nulls reuses it

  // ---- choose_compressor (integer/mod.rs:231-308)
  const bool bp_ok = tclass != TC_FLOAT && W == 4 && !st.has_neg && (n % 128 == 0); // bp.rs:93-97
  const bool exact_small = st.unique != kNone;                                      // unique <= n/3
  const bool one = n && exact_small && st.unique <= 1;
  int codec = o.def_codec;
  {
    auto applicable = [&](int c) {
      switch (c) {
      case SB_C_FREQ:
      case SB_C_DICT:
      case SB_C_RLE: return true;
      case SB_C_ONEVALUE: return n == 0 || one;
      case SB_C_BITPACK: return bp_ok;
      case SB_C_DELTABP: return bp_ok && st.sorted_nonnull && st.null_count == 0;
      case SB_C_PATAS: return tclass == TC_FLOAT && W == 8 && n > 0;
      }
      return false;
    };
    if (o.force >= SB_C_RLE && !e_forbidden(o, o.force) && applicable(o.force)) {
      codec = o.force;
    } else if (o.ratio >= 0 && n) {
      double max_ratio = o.ratio;
      const int order_i[6] = {SB_C_ONEVALUE, SB_C_FREQ, SB_C_DICT, SB_C_RLE, SB_C_BITPACK, SB_C_DELTABP};
      const int order_f[5] = {SB_C_ONEVALUE, SB_C_FREQ, SB_C_DICT, SB_C_PATAS, SB_C_RLE};
      const int cnt = tclass == TC_FLOAT ? 5 : 6;
      for (int k = 0; k < cnt; ++k) {
        int c = tclass == TC_FLOAT ? order_f[k] : order_i[k];
        if (e_forbidden(o, c)) continue;
        double r = 0.0;
        switch (c) {
        case SB_C_ONEVALUE: r = one ? double(n) : 0.0; break; // unique_count <= 1 (one_value.rs:53-59)
        case SB_C_FREQ: // freq.rs:129-151
          if (exact_small && st.unique <= 1) r = 0.0;
          else if (double(st.null_count) / double(n) >= 0.9) r = double(n - 1);
          else if (exact_small && double(st.max_count) / double(n) >= 0.9 && (tclass == TC_FLOAT || st.max_ge_256)) r = double(n - 1);
          break;
        case SB_C_DICT: // dict.rs:109-120 (unique * 3 >= n -> 0; integer division bits/8)
          if (exact_small && uint64_t(st.unique) * 3 < n) {
            uint64_t after = uint64_t(st.unique) * W + uint64_t(n) * (bits_needed(st.unique) / 8) + uint64_t(n) * 2 / 128;
            r = double(uint64_t(n) * W) / double(after);
          }
          break;
        case SB_C_RLE: r = fixed_sample_ratio(cx, v, valid, n, SB_C_RLE, o, st.null_count); break;
        case SB_C_PATAS: r = W == 8 ? fixed_sample_ratio(cx, v, valid, n, SB_C_PATAS, o, st.null_count) : 0.0; break;
        case SB_C_BITPACK: r = bp_ok ? fixed_sample_ratio(cx, v, valid, n, SB_C_BITPACK, o, st.null_count) : 0.0; break;
        case SB_C_DELTABP:
          r = (bp_ok && st.sorted_nonnull && st.null_count == 0) ? fixed_sample_ratio(cx, v, valid, n, SB_C_BITPACK, o, st.null_count) * 1.5 : 0.0;
          break;
        }
        if (r > max_ratio) {
          max_ratio = r;
          codec = c;
          if (r == double(n)) break;
        }
      }
    }
  }
  if (*cx.err) return kEncFail;

  // ---- null replacement sources, shared by RLE / Dict
  const bool has_nulls = valid.p && st.null_count;
  uint32_t *src = nullptr;
  uint32_t first_valid = kNone;
  const bool reuse_tab = codec == SB_C_DICT && !has_nulls && st.unique != kNone; // raw values == effective values
  if (!reuse_tab) cx.ar = mark;
  if (has_nulls && (codec == SB_C_RLE || codec == SB_C_DICT || codec == SB_C_ONEVALUE)) {
    src = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
    if (!src) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    fill_forward(cx, valid, n, false, src);
    if (tid == 0) cx.bcast[2] = int(kNone);
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SB_NT)
      if (valid.get(i)) {
        atomicMin(reinterpret_cast<uint32_t *>(cx.bcast + 2), i);
        break;
      }
    __syncthreads();
    first_valid = uint32_t(cx.bcast[2]);
    __syncthreads();
  } else if (n) {
    first_valid = 0;
  }

  uint32_t payload = 0;
  switch (codec) {
  case SB_C_NONE:
  case SB_C_LZ4:
  case SB_C_ZSTD:
  case SB_C_SNAPPY: payload = enc_basic(cx, codec, v.p, n * uint32_t(W), body); break;
  case SB_C_ONEVALUE: // first valid value, else default (one_value.rs:61-75)
    if (tid == 0) st_le(body, first_valid == kNone ? 0 : v.get(first_valid), W);
    payload = W;
    break;
  case SB_C_RLE: {
    EffFixed acc{v, src, first_valid == kNone ? 0 : v.get(first_valid)};
    payload = enc_rle(cx, acc, n, W, body, false);
    break;
  }
  case SB_C_BITPACK:
  case SB_C_DELTABP: payload = enc_bitpack(cx, reinterpret_cast<const uint32_t *>(v.p), n, codec == SB_C_DELTABP, body, false); break;
  case SB_C_PATAS: payload = enc_patas(cx, reinterpret_cast<const uint64_t *>(v.p), n, body, false); break;
  case SB_C_DICT: {
    if constexpr (LEVEL >= 2) {
      cx.flag(SB_OUT_OF_SPEC);
      return kEncFail;
    } else {
      // ids in first-occurrence order over the effective values; leading nulls intern T::default()
      EffFixed acc{v, src, 0};
      uint32_t *slot_of = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
      uint32_t *idx = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
      if (!slot_of || !idx) {
        cx.flag(SB_NYI);
        return kEncFail;
      }
      HashTab dt = tab;
      uint32_t k = st.unique;
      if (reuse_tab) hash_lookup(acc, n, dt, slot_of);
      // null replacement only repeats values of the page (or interns T::default()): at most unique + 1 keys
      else k = hash_distinct(cx, acc, n, st.unique != kNone ? st.unique + 2 : n + 1, &dt, slot_of);
      if (*cx.err || k == kNone) return kEncFail;
      // rank the first occurrences: idx[] is used as the flag / rank array over rows
      for (uint32_t i = tid; i < n; i += SB_NT) idx[i] = 0;
      __syncthreads();
      for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
        if (dt.rep[h]) idx[dt.first(h)] = 1;
      __syncthreads();
      {
        constexpr uint32_t EPT = 8, CH = SB_NT * EPT;
        uint32_t run = 0;
        for (uint32_t c0 = 0; c0 < n; c0 += CH) {
          const uint32_t e0 = c0 + tid * EPT;
          uint32_t f[EPT], c = 0;
#pragma unroll
          for (uint32_t j = 0; j < EPT; ++j) {
            f[j] = e0 + j < n ? idx[e0 + j] : 0;
            c += f[j];
          }
          uint32_t total;
          uint32_t pre = run + block_excl_scan(c, cx.ws, &total);
#pragma unroll
          for (uint32_t j = 0; j < EPT; ++j)
            if (e0 + j < n) {
              idx[e0 + j] = pre; // rank of the first occurrence at/after this row; exact at flagged rows
              pre += f[j];
            }
          run += total;
        }
      }
      __syncthreads();
      uint32_t *ids = slot_of; // rewrite in place: row -> slot -> first row of the key -> its rank
      for (uint32_t i = tid; i < n; i += SB_NT) ids[i] = idx[dt.first(slot_of[i])];
      __syncthreads();
      // the dictionary values are parked in scratch so that the table's shared memory is free for the index page
      uint8_t *dvals = static_cast<uint8_t *>(cx.ar.alloc_global(uint64_t(k) * W + 16));
      if (!dvals) {
        cx.flag(SB_NYI);
        return kEncFail;
      }
      for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
        if (dt.rep[h]) st_le(dvals + uint64_t(idx[dt.first(h)]) * W, acc.key(dt.first(h)), W);
      __syncthreads();
      cx.ar.s_cur = mark.s_cur;
      EOpts sub = o;
      sub.forbidden |= 1u << SB_C_DICT;
      uint32_t used = enc_fixed<LEVEL + 1>(cx, Vals{reinterpret_cast<const uint8_t *>(ids), 4}, TC_UINT, Bits{nullptr, 0}, n, sub, body);
      if (used == kEncFail) return kEncFail;
      if (tid == 0) st_le(body + used, k, 4);
      copy_bytes(body + used + 4, dvals, uint64_t(k) * W);
      __syncthreads();
      payload = used + 4 + k * uint32_t(W);
    }
    break;
  }
  case SB_C_FREQ: {
    if constexpr (LEVEL >= 2) {
      cx.flag(SB_OUT_OF_SPEC);
      return kEncFail;
    } else {
      // freq.rs:33-86: top = most frequent value over all slots (T::default() when >= 90 % null);
      // exceptions = valid rows != top, in row order; Freq forbidden below
      const bool top_is_null = n && double(st.null_count) / double(n) >= 0.9;
      uint64_t top = 0;
      if (!top_is_null && n) {
        if (st.unique == kNone) { // forced Freq on high-cardinality data: needs the full table
          HashTab ft{};
          uint32_t k = hash_distinct(cx, raw, n, n + 1, &ft, nullptr);
          if (*cx.err || k == kNone) return kEncFail;
          hash_top(cx, ft, &st.max_count, &st.top_first);
        }
        top = v.get(st.top_first);
      }
      uint32_t *rows = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
      uint8_t *exc = static_cast<uint8_t *>(cx.ar.alloc_global(uint64_t(n) * W + 16));
      if (!rows || !exc) {
        cx.flag(SB_NYI);
        return kEncFail;
      }
      uint32_t n_exc = 0;
      {
        constexpr uint32_t EPT = 4, CH = SB_NT * EPT;
        for (uint32_t c0 = 0; c0 < n; c0 += CH) {
          const uint32_t e0 = c0 + tid * EPT;
          uint32_t flags = 0, c = 0;
#pragma unroll
          for (uint32_t j = 0; j < EPT; ++j) {
            uint32_t i = e0 + j;
            if (i < n && valid.get(i) && (top_is_null || v.get(i) != top)) {
              flags |= 1u << j;
              ++c;
            }
          }
          uint32_t total;
          uint32_t pre = n_exc + block_excl_scan(c, cx.ws, &total);
#pragma unroll
          for (uint32_t j = 0; j < EPT; ++j)
            if ((flags >> j) & 1u) {
              rows[pre] = e0 + j;
              uint64_t x = v.get(e0 + j);
              switch (W) {
              case 1: exc[pre] = uint8_t(x); break;
              case 2: reinterpret_cast<uint16_t *>(exc)[pre] = uint16_t(x); break;
              case 4: reinterpret_cast<uint32_t *>(exc)[pre] = uint32_t(x); break;
              default: reinterpret_cast<uint64_t *>(exc)[pre] = x; break;
              }
              ++pre;
            }
          n_exc += total;
        }
      }
      __syncthreads();
      cx.ar.s_cur = mark.s_cur; // shared memory back to the exceptions' own page (rows / exc live in scratch)
      if (tid == 0) st_le(body, top, W);
      uint32_t bm = enc_roaring(cx, rows, n_exc, n, body + W + 4);
      if (tid == 0) st_le(body + W, bm, 4);
      EOpts sub = o;
      sub.forbidden |= 1u << SB_C_FREQ;
      uint32_t used = enc_fixed<LEVEL + 1>(cx, Vals{exc, W}, tclass, Bits{nullptr, 0}, n_exc, sub, body + W + 4 + bm);
      if (used == kEncFail) return kEncFail;
      payload = uint32_t(W) + 4 + bm + used;
    }
    break;
  }
  default: cx.flag(SB_OUT_OF_SPEC); return kEncFail;
  }
  if (payload == kEncFail || *cx.err) return kEncFail;
  put_hdr9(out, codec, payload, n * uint32_t(W)); // uncompressed = n * W (integer/mod.rs:62-63)
  __syncthreads();
  cx.ar = entry;
  return 9 + payload;
}

// ------------------------------------------------------------------------------------
// i128 / i256 leaves (Decimal128 / Decimal256 storage; compress_integer over `i128` / `i256`,
// src/write/primitive.rs:71-78, src/compression/integer/traits.rs:28-39).  Same statistics and chooser as
// enc_fixed; Bitpacking / DeltaBitpacking never apply (size_of::<T>() != 4, bp.rs:93-97), Patas is float only.
// Values are moved as 16-byte vectors and compared word by word; ordering (for `max.as_i64() >= 256`,
// freq.rs:146) is signed two's complement.
// ------------------------------------------------------------------------------------
struct EffWide {
  const uint8_t *p;    // 16-byte aligned values
  int W;               // 16 or 32
  const uint32_t *src; // null replacement rows (nullptr: identity); kNone = leading null
  uint32_t lead;       // row whose value leading nulls take, kNone = T::default() (zero)
  __device__ __forceinline__ uint32_t row(uint32_t i) const {
    uint32_t r = src ? src[i] : i;
    return r == kNone ? lead : r;
  }
  __device__ __forceinline__ uint4 vec(uint32_t r, int k) const {
    return r == kNone ? make_uint4(0, 0, 0, 0) : reinterpret_cast<const uint4 *>(p + uint64_t(r) * W)[k];
  }
  __device__ __forceinline__ uint32_t hash(uint32_t i) const {
    const uint32_t r = row(i);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int k = 0; k < W / 16; ++k) {
      const uint4 v = vec(r, k);
      h = (h ^ (uint64_t(v.x) | (uint64_t(v.y) << 32))) * 0xBF58476D1CE4E5B9ull;
      h = (h ^ (h >> 29) ^ (uint64_t(v.z) | (uint64_t(v.w) << 32))) * 0x94D049BB133111EBull;
    }
    return uint32_t(h >> 32) ^ uint32_t(h);
  }
  __device__ __forceinline__ bool equal(uint32_t i, uint32_t j) const {
    const uint32_t a = row(i), b = row(j);
    if (a == b) return true;
    for (int k = 0; k < W / 16; ++k) {
      const uint4 x = vec(a, k), y = vec(b, k);
      if (x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w) return false;
    }
    return true;
  }
  __device__ __forceinline__ void put(uint8_t *out, uint32_t i, int) const { // out: any alignment
    const uint32_t r = row(i);
    for (int k = 0; k < W / 16; ++k) {
      const uint4 v = vec(r, k);
      st_le(out + 16 * k, v.x, 4), st_le(out + 16 * k + 4, v.y, 4), st_le(out + 16 * k + 8, v.z, 4), st_le(out + 16 * k + 12, v.w, 4);
    }
  }
  // signed compare of the raw values at rows a, b
  __device__ __forceinline__ bool less(uint32_t a, uint32_t b) const {
    for (int k = W / 16 - 1; k >= 0; --k) {
      const uint4 x = vec(a, k), y = vec(b, k);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
      for (int w = 3; w >= 0; --w) {
        if (xs[w] == ys[w]) continue;
        if (k == W / 16 - 1 && w == 3) return int32_t(xs[w]) < int32_t(ys[w]);
        return xs[w] < ys[w];
      }
    }
    return false;
  }
};

// ids in first-occurrence order from a distinct table (shared by the wide Dict path; enc_fixed / enc_binary carry
// the same steps inline): on return slot_of[row] = id, dt.cnt[slot] = id of that slot.  idx is scratch (n entries).
__device__ void dict_assign_ids(Dctx &cx, HashTab &dt, uint32_t *slot_of, uint32_t *idx, uint32_t n) {
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < n; i += SB_NT) idx[i] = 0;
  __syncthreads();
  for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
    if (dt.rep[h]) idx[dt.first(h)] = 1;
  __syncthreads();
  constexpr uint32_t EPT = 8, CH = SB_NT * EPT;
  uint32_t run = 0;
  for (uint32_t c0 = 0; c0 < n; c0 += CH) {
    const uint32_t e0 = c0 + tid * EPT;
    uint32_t f[EPT], c = 0;
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j) {
      f[j] = e0 + j < n ? idx[e0 + j] : 0;
      c += f[j];
    }
    uint32_t total;
    uint32_t pre = run + block_excl_scan(c, cx.ws, &total);
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j)
      if (e0 + j < n) {
        idx[e0 + j] = pre;
        pre += f[j];
      }
    run += total;
  }
  __syncthreads();
  for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
    if (dt.rep[h]) dt.cnt[h] = idx[dt.first(h)];
  __syncthreads();
  for (uint32_t i = tid; i < n; i += SB_NT) slot_of[i] = dt.cnt[slot_of[i]];
  __syncthreads();
}

__device__ uint32_t first_valid_row(Dctx &cx, const Bits &valid, uint32_t n) {
  if (threadIdx.x == 0) cx.bcast[2] = int(kNone);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += SB_NT)
    if (valid.get(i)) {
      atomicMin(reinterpret_cast<uint32_t *>(cx.bcast + 2), i);
      break;
    }
  __syncthreads();
  const uint32_t fv = uint32_t(cx.bcast[2]);
  __syncthreads();
  return fv;
}

template <int LEVEL>
__device__ uint32_t enc_wide(Dctx &cx, const uint8_t *vals, int W, Bits valid, uint32_t n, EOpts o, uint8_t *out) {
  const uint32_t tid = threadIdx.x;
  Arena mark = cx.ar;
  uint8_t *body = out + 9;
  // ---- gen_stats: null count, max (signed, over all slots), distinct table over all slots
  EffWide raw{vals, W, nullptr, kNone};
  uint32_t nulls = 0, best = kNone;
  for (uint32_t i = tid; i < n; i += SB_NT) {
    nulls += !valid.get(i);
    if (best == kNone || raw.less(best, i)) best = i;
  }
  const uint32_t null_count = uint32_t(block_sum_u64e(cx, nulls));
  bool max_ge_256 = false;
  {
    uint32_t *cand = static_cast<uint32_t *>(cx.ar.alloc(SB_NT * 4));
    if (!cand) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    cand[tid] = best;
    __syncthreads();
    if (tid == 0) {
      uint32_t b = kNone;
      for (uint32_t t = 0; t < SB_NT; ++t)
        if (cand[t] != kNone && (b == kNone || raw.less(b, cand[t]))) b = cand[t];
      int64_t lo64 = 0;
      if (b != kNone) {
        const uint4 v = raw.vec(b, 0);
        lo64 = int64_t(uint64_t(v.x) | (uint64_t(v.y) << 32)); // as_i64: the low 64 bits (traits.rs:28-39)
      }
      cx.bcast[3] = lo64 >= 256;
    }
    __syncthreads();
    max_ge_256 = cx.bcast[3] != 0;
    __syncthreads();
    cx.ar = mark;
  }
  HashTab tab{};
  const uint32_t limit = n / 3 + 1;
  uint32_t unique = n ? hash_distinct(cx, raw, n, limit, &tab, nullptr) : 0;
  if (*cx.err) return kEncFail;
  uint32_t max_count = 0, top_first = 0;
  if (unique != kNone && n) hash_top(cx, tab, &max_count, &top_first);
  cx.ar = mark;
  const bool exact_small = unique != kNone, one = n && exact_small && unique <= 1;
  const bool has_nulls = valid.p && null_count;

  // null replacement rows (RLE / Dict), first valid row
  uint32_t *src = nullptr;
  auto need_src = [&]() -> bool {
    if (!has_nulls || src) return true;
    src = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!src) {
      cx.flag(SB_NYI);
      return false;
    }
    fill_forward(cx, valid, n, false, src);
    return true;
  };

  // ---- choose_compressor (integer/mod.rs:231-308): OneValue, Freq, Dict, RLE (Bitpacking / Delta: ratio 0 for this width)
  int codec = o.def_codec;
  auto rle_sample_ratio = [&]() -> double { // compress_sample_ratio on 10 x 64 rows (integer/mod.rs:310-347)
    Arena m2 = cx.ar;
    const uint32_t m = sample_whole(n) ? n : 640;
    uint8_t *sv = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(m) * W + 16));
    uint8_t *sb = static_cast<uint8_t *>(cx.ar.alloc((m + 7) / 8 + 16));
    uint32_t *ssrc = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(m) * 4 + 16));
    if (!sv || !sb || !ssrc) {
      cx.flag(SB_NYI);
      return 0.0;
    }
    for (uint32_t i = tid; i < (m + 7) / 8; i += SB_NT) {
      uint32_t byte = 0;
      for (uint32_t b = 0; b < 8 && 8 * i + b < m; ++b) {
        const uint32_t r = sample_whole(n) ? 8 * i + b : sample_row(n, o, SB_C_RLE, 8 * i + b);
        byte |= uint32_t(valid.get(r)) << b;
      }
      sb[i] = uint8_t(byte);
    }
    for (uint32_t k = tid; k < m; k += SB_NT) {
      const uint32_t r = sample_whole(n) ? k : sample_row(n, o, SB_C_RLE, k);
      for (int q = 0; q < W / 16; ++q) reinterpret_cast<uint4 *>(sv + uint64_t(k) * W)[q] = raw.vec(r, q);
    }
    __syncthreads();
    EffWide acc{sv, W, nullptr, kNone};
    if (has_nulls) {
      const Bits sbits{sb, 0};
      fill_forward(cx, sbits, m, false, ssrc);
      acc.src = ssrc;
      acc.lead = first_valid_row(cx, sbits, m);
    }
    const uint32_t size = enc_rle(cx, acc, m, W, nullptr, true);
    cx.ar = m2;
    return size == kEncFail || size == 0 ? 0.0 : double(uint64_t(m) * W) / double(size);
  };
  if (o.force >= SB_C_RLE && !e_forbidden(o, o.force) &&
      (o.force == SB_C_FREQ || o.force == SB_C_DICT || o.force == SB_C_RLE || (o.force == SB_C_ONEVALUE && (n == 0 || one)))) {
    codec = o.force;
  } else if (o.ratio >= 0 && n) {
    double max_ratio = o.ratio;
    const int order[4] = {SB_C_ONEVALUE, SB_C_FREQ, SB_C_DICT, SB_C_RLE};
    for (int k = 0; k < 4; ++k) {
      const int c = order[k];
      if (e_forbidden(o, c)) continue;
      double r = 0.0;
      switch (c) {
      case SB_C_ONEVALUE: r = one ? double(n) : 0.0; break;
      case SB_C_FREQ:
        if (exact_small && unique <= 1) r = 0.0;
        else if (double(null_count) / double(n) >= 0.9) r = double(n - 1);
        else if (exact_small && double(max_count) / double(n) >= 0.9 && max_ge_256) r = double(n - 1);
        break;
      case SB_C_DICT:
        if (exact_small && uint64_t(unique) * 3 < n) {
          const uint64_t after = uint64_t(unique) * W + uint64_t(n) * (bits_needed(unique) / 8) + uint64_t(n) * 2 / 128;
          r = double(uint64_t(n) * W) / double(after);
        }
        break;
      case SB_C_RLE: r = rle_sample_ratio(); break;
      }
      if (r > max_ratio) {
        max_ratio = r;
        codec = c;
        if (r == double(n)) break;
      }
    }
  }
  if (*cx.err) return kEncFail;

  uint32_t payload = 0;
  switch (codec) {
  case SB_C_NONE:
  case SB_C_LZ4:
  case SB_C_ZSTD:
  case SB_C_SNAPPY: payload = enc_basic(cx, codec, vals, n * uint32_t(W), body); break;
  case SB_C_ONEVALUE: { // first valid value, else default (one_value.rs:61-75)
    const uint32_t fv = n ? first_valid_row(cx, valid, n) : kNone;
    if (tid == 0) {
      if (fv != kNone) EffWide{vals, W, nullptr, kNone}.put(body, fv, W);
      else
        for (int b = 0; b < W; ++b) body[b] = 0;
    }
    payload = W;
    break;
  }
  case SB_C_RLE: {
    if (!need_src()) return kEncFail;
    EffWide acc{vals, W, src, has_nulls ? first_valid_row(cx, valid, n) : kNone};
    payload = enc_rle(cx, acc, n, W, body, false);
    break;
  }
  case SB_C_DICT: {
    if constexpr (LEVEL >= 2) {
      cx.flag(SB_OUT_OF_SPEC);
      return kEncFail;
    } else {
      if (!need_src()) return kEncFail;
      EffWide acc{vals, W, src, kNone}; // leading nulls intern T::default()
      uint32_t *slot_of = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
      uint32_t *idx = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
      if (!slot_of || !idx) {
        cx.flag(SB_NYI);
        return kEncFail;
      }
      HashTab dt{};
      const uint32_t k = hash_distinct(cx, acc, n, n + 1, &dt, slot_of);
      if (*cx.err || k == kNone) return kEncFail;
      dict_assign_ids(cx, dt, slot_of, idx, n);
      EOpts sub = o;
      sub.forbidden |= 1u << SB_C_DICT;
      const uint32_t used = enc_fixed<LEVEL + 1>(cx, Vals{reinterpret_cast<const uint8_t *>(slot_of), 4}, TC_UINT, Bits{nullptr, 0}, n, sub, body);
      if (used == kEncFail) return kEncFail;
      if (tid == 0) st_le(body + used, k, 4);
      uint8_t *tabo = body + used + 4;
      for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
        if (dt.rep[h]) acc.put(tabo + uint64_t(dt.cnt[h]) * W, dt.first(h), W);
      __syncthreads();
      payload = used + 4 + k * uint32_t(W);
    }
    break;
  }
  case SB_C_FREQ: {
    if constexpr (LEVEL >= 2) {
      cx.flag(SB_OUT_OF_SPEC);
      return kEncFail;
    } else {
      const bool top_is_null = n && double(null_count) / double(n) >= 0.9;
      uint32_t top_row = kNone; // kNone = T::default()
      if (!top_is_null && n) {
        if (unique == kNone) { // forced Freq on high-cardinality data: needs the full table
          HashTab ft{};
          const uint32_t k = hash_distinct(cx, raw, n, n + 1, &ft, nullptr);
          if (*cx.err || k == kNone) return kEncFail;
          hash_top(cx, ft, &max_count, &top_first);
        }
        top_row = top_first;
      }
      uint32_t *rows = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
      uint8_t *exc = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(n) * W + 16));
      if (!rows || !exc) {
        cx.flag(SB_NYI);
        return kEncFail;
      }
      const EffWide cmp{vals, W, nullptr, kNone};
      uint32_t n_exc = 0;
      constexpr uint32_t EPT = 4, CH = SB_NT * EPT;
      for (uint32_t c0 = 0; c0 < n; c0 += CH) {
        const uint32_t e0 = c0 + tid * EPT;
        uint32_t flags = 0, c = 0;
#pragma unroll
        for (uint32_t j = 0; j < EPT; ++j) {
          const uint32_t i = e0 + j;
          bool differs = false;
          if (i < n && valid.get(i)) {
            if (top_is_null) differs = true;
            else
              for (int q = 0; q < W / 16; ++q) {
                const uint4 x = cmp.vec(i, q), y = cmp.vec(top_row, q);
                differs |= x.x != y.x || x.y != y.y || x.z != y.z || x.w != y.w;
              }
          }
          if (differs) {
            flags |= 1u << j;
            ++c;
          }
        }
        uint32_t total;
        uint32_t pre = n_exc + block_excl_scan(c, cx.ws, &total);
#pragma unroll
        for (uint32_t j = 0; j < EPT; ++j)
          if ((flags >> j) & 1u) {
            rows[pre] = e0 + j;
            for (int q = 0; q < W / 16; ++q) reinterpret_cast<uint4 *>(exc + uint64_t(pre) * W)[q] = cmp.vec(e0 + j, q);
            ++pre;
          }
        n_exc += total;
      }
      __syncthreads();
      if (tid == 0) {
        if (top_row != kNone) cmp.put(body, top_row, W);
        else
          for (int b = 0; b < W; ++b) body[b] = 0;
      }
      const uint32_t bm = enc_roaring(cx, rows, n_exc, n, body + W + 4);
      if (tid == 0) st_le(body + W, bm, 4);
      EOpts sub = o;
      sub.forbidden |= 1u << SB_C_FREQ;
      const uint32_t used = enc_wide<LEVEL + 1>(cx, exc, W, Bits{nullptr, 0}, n_exc, sub, body + W + 4 + bm);
      if (used == kEncFail) return kEncFail;
      payload = uint32_t(W) + 4 + bm + used;
    }
    break;
  }
  default: cx.flag(SB_OUT_OF_SPEC); return kEncFail;
  }
  if (payload == kEncFail || *cx.err) return kEncFail;
  put_hdr9(out, codec, payload, n * uint32_t(W));
  __syncthreads();
  cx.ar = mark;
  return 9 + payload;
}

// ------------------------------------------------------------------------------------
// write_validity (write/serialize.rs:200-215): [u32 L][ULEB((ceil8(n) << 1) | 1)][bitmap],
// written for every nullable field, all ones when the array carries no validity.
// ------------------------------------------------------------------------------------
__device__ uint32_t enc_validity(const Bits &valid, uint32_t n, uint8_t *out) {
  const uint32_t nbytes = (n + 7) / 8;
  uint64_t header = (uint64_t(nbytes) << 1) | 1;
  uint32_t ul = 0;
  uint8_t ub[10];
  do {
    uint8_t b = header & 0x7f;
    header >>= 7;
    ub[ul++] = b | (header ? 0x80 : 0);
  } while (header);
  if (threadIdx.x == 0) {
    st_le(out, ul + nbytes, 4);
    for (uint32_t i = 0; i < ul; ++i) out[4 + i] = ub[i];
  }
  uint8_t *d = out + 4 + ul;
  for (uint32_t i = threadIdx.x; i < nbytes; i += SB_NT) {
    uint32_t byte = 0;
    for (uint32_t b = 0; b < 8 && 8 * i + b < n; ++b) byte |= uint32_t(valid.get(8 * i + b)) << b;
    d[i] = uint8_t(byte);
  }
  return 4 + ul + nbytes;
}

// ------------------------------------------------------------------------------------
// compress_boolean (boolean/mod.rs:23-61)
// ------------------------------------------------------------------------------------
struct EffBool {
  Bits v;
  const uint32_t *src;
  uint32_t lead;
  __device__ __forceinline__ uint64_t key(uint32_t i) const {
    if (!src) return v.get(i);
    uint32_t s = src[i];
    return s == kNone ? lead : uint32_t(v.get(s));
  }
  __device__ __forceinline__ bool equal(uint32_t i, uint32_t j) const { return key(i) == key(j); }
  __device__ __forceinline__ void put(uint8_t *out, uint32_t i, int W) const { st_le(out, key(i), W); }
};
__device__ uint32_t bool_rle(Dctx &cx, const Bits &vals, const Bits &valid, uint32_t n, bool has_nulls, uint8_t *out, bool count_only) {
  Arena mark = cx.ar;
  EffBool acc{vals, nullptr, 0};
  if (has_nulls) {
    uint32_t *src = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!src) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    fill_forward(cx, valid, n, false, src);
    if (threadIdx.x == 0) cx.bcast[2] = int(kNone);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += SB_NT)
      if (valid.get(i)) {
        atomicMin(reinterpret_cast<uint32_t *>(cx.bcast + 2), i);
        break;
      }
    __syncthreads();
    uint32_t fv = uint32_t(cx.bcast[2]);
    __syncthreads();
    acc.src = src;
    acc.lead = fv == kNone ? 0 : uint32_t(vals.get(fv));
  }
  uint32_t r = enc_rle(cx, acc, n, 1, out, count_only);
  cx.ar = mark;
  return r;
}
__device__ uint32_t enc_boolean(Dctx &cx, Bits vals, Bits valid, uint32_t n, EOpts o, uint8_t *out) {
  const uint32_t tid = threadIdx.x;
  uint8_t *body = out + 9;
  uint32_t nulls = 0, trues = 0, falses = 0;
  for (uint32_t i = tid; i < n; i += SB_NT) {
    if (!valid.get(i)) ++nulls;
    else if (vals.get(i)) ++trues;
    else ++falses;
  }
  nulls = uint32_t(block_sum_u64e(cx, nulls));
  trues = uint32_t(block_sum_u64e(cx, trues));
  falses = uint32_t(block_sum_u64e(cx, falses));
  const bool one = trues == 0 || falses == 0; // boolean/one_value.rs:36-42
  const bool has_nulls = valid.p && nulls;
  int codec = o.def_codec;
  if (o.force == SB_C_RLE && !e_forbidden(o, SB_C_RLE)) codec = SB_C_RLE;
  else if (o.force == SB_C_ONEVALUE && one && !e_forbidden(o, SB_C_ONEVALUE)) codec = SB_C_ONEVALUE;
  else if (o.ratio >= 0) { // boolean/mod.rs:194-239
    double max_ratio = o.ratio;
    const int order[2] = {SB_C_ONEVALUE, SB_C_RLE};
    for (int k = 0; k < 2; ++k) {
      int c = order[k];
      if (e_forbidden(o, c)) continue;
      double r;
      if (c == SB_C_ONEVALUE) r = one ? double(n) : 0.0;
      else { // boolean/mod.rs:241-278: sample, RLE it, (rows / 8) / compressed
        Arena mark = cx.ar;
        uint32_t m = sample_whole(n) ? n : 640;
        uint8_t *sv = static_cast<uint8_t *>(cx.ar.alloc((m + 7) / 8 + 16)), *sm = static_cast<uint8_t *>(cx.ar.alloc((m + 7) / 8 + 16));
        if (!sv || !sm) {
          cx.flag(SB_NYI);
          return kEncFail;
        }
        for (uint32_t i = tid; i < (m + 7) / 8; i += SB_NT) {
          uint32_t a = 0, b = 0;
          for (uint32_t q = 0; q < 8 && 8 * i + q < m; ++q) {
            uint32_t row = sample_whole(n) ? 8 * i + q : sample_row(n, o, SB_C_RLE, 8 * i + q);
            a |= uint32_t(vals.get(row)) << q;
            b |= uint32_t(valid.get(row)) << q;
          }
          sv[i] = uint8_t(a);
          sm[i] = uint8_t(b);
        }
        __syncthreads();
        uint32_t size = bool_rle(cx, Bits{sv, 0}, Bits{valid.p ? sm : nullptr, 0}, m, has_nulls, nullptr, true);
        cx.ar = mark;
        if (size == kEncFail) return kEncFail;
        r = double(m / 8) / double(size);
      }
      if (r > max_ratio) {
        max_ratio = r;
        codec = c;
        if (r == double(n)) break;
      }
    }
  }
  uint32_t payload;
  if (codec <= SB_C_SNAPPY) { // re-packed to bit offset 0 (boolean/mod.rs:47-52)
    Arena mark = cx.ar;
    const uint32_t nbytes = (n + 7) / 8;
    uint8_t *packed = codec == SB_C_NONE ? body : static_cast<uint8_t *>(cx.ar.alloc(uint64_t(nbytes) + 16));
    if (!packed) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    for (uint32_t i = tid; i < nbytes; i += SB_NT) {
      uint32_t byte = 0;
      for (uint32_t b = 0; b < 8 && 8 * i + b < n; ++b) byte |= uint32_t(vals.get(8 * i + b)) << b;
      packed[i] = uint8_t(byte);
    }
    __syncthreads();
    payload = codec == SB_C_NONE ? nbytes : enc_basic(cx, codec, packed, nbytes, body);
    cx.ar = mark;
  } else if (codec == SB_C_RLE) {
    payload = bool_rle(cx, vals, valid, n, has_nulls, body, false);
  } else if (codec == SB_C_ONEVALUE) { // first valid value, else false (boolean/one_value.rs:44-52)
    if (tid == 0) cx.bcast[2] = int(kNone);
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SB_NT)
      if (valid.get(i)) {
        atomicMin(reinterpret_cast<uint32_t *>(cx.bcast + 2), i);
        break;
      }
    __syncthreads();
    uint32_t fv = uint32_t(cx.bcast[2]);
    __syncthreads();
    if (tid == 0) body[0] = fv == kNone ? 0 : uint8_t(vals.get(fv));
    payload = 1;
  } else {
    cx.flag(SB_OUT_OF_SPEC);
    return kEncFail;
  }
  if (payload == kEncFail || *cx.err) return kEncFail;
  put_hdr9(out, codec, payload, n); // uncompressed = ROWS (boolean/mod.rs:59, App. C9)
  __syncthreads();
  return 9 + payload;
}

// ------------------------------------------------------------------------------------
// compress_binary (binary/mod.rs:26-93)
// ------------------------------------------------------------------------------------
struct BinView {
  const uint8_t *data;  // values buffer base
  const uint8_t *offs;  // offsets of this page: n + 1 entries, element 0 = first row of the page
  int OW;
  __device__ __forceinline__ int64_t off(uint32_t i) const {
    return OW == 4 ? int64_t(reinterpret_cast<const int32_t *>(offs)[i]) : reinterpret_cast<const int64_t *>(offs)[i];
  }
  __device__ __forceinline__ uint32_t len(uint32_t i) const { return uint32_t(off(i + 1) - off(i)); }
  __device__ __forceinline__ const uint8_t *ptr(uint32_t i) const { return data + off(i); }
};
struct EffBin {
  BinView b;
  const uint32_t *src; // nullptr: identity
  __device__ __forceinline__ uint32_t row(uint32_t i) const { return src ? src[i] : i; }
  __device__ __forceinline__ uint32_t hash(uint32_t i) const {
    uint32_t r = row(i), l = b.len(r);
    const uint8_t *p = b.ptr(r);
    uint32_t h = 2166136261u ^ l;
    for (uint32_t k = 0; k < l; ++k) h = (h ^ p[k]) * 16777619u;
    return h ^ (h >> 15);
  }
  __device__ __forceinline__ bool equal(uint32_t i, uint32_t j) const {
    uint32_t a = row(i), c = row(j);
    if (a == c) return true;
    uint32_t l = b.len(a);
    if (l != b.len(c)) return false;
    const uint8_t *p = b.ptr(a), *q = b.ptr(c);
    for (uint32_t k = 0; k < l; ++k)
      if (p[k] != q[k]) return false;
    return true;
  }
};
// copies `cnt` length-prefixed entries `[u64 len][bytes]` of the given rows; returns bytes
__device__ uint32_t put_entries(Dctx &cx, const BinView &b, const uint32_t *rows, uint32_t cnt, uint8_t *out) {
  constexpr uint32_t EPT = 4, CH = SB_NT * EPT;
  uint32_t pos = 0;
  for (uint32_t c0 = 0; c0 < cnt; c0 += CH) {
    const uint32_t e0 = c0 + threadIdx.x * EPT;
    uint32_t sz = 0;
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j)
      if (e0 + j < cnt) sz += 8 + b.len(rows[e0 + j]);
    uint32_t total;
    uint32_t p = pos + block_excl_scan(sz, cx.ws, &total);
#pragma unroll
    for (uint32_t j = 0; j < EPT; ++j)
      if (e0 + j < cnt) {
        uint32_t r = rows[e0 + j], l = b.len(r);
        st_le(out + p, l, 8);
        const uint8_t *s = b.ptr(r);
        for (uint32_t k = 0; k < l; ++k) out[p + 8 + k] = s[k];
        p += 8 + l;
      }
    pos += total;
  }
  __syncthreads();
  return pos;
}
__device__ uint32_t enc_binary(Dctx &cx, BinView b, Bits valid, uint32_t n, uint64_t backing_len, EOpts o, uint8_t *out) {
  const uint32_t tid = threadIdx.x;
  const int OW = b.OW;
  uint8_t *body = out + 9;
  Arena mark = cx.ar;
  // ---- stats (binary/mod.rs:253-291): distinct over ALL slots
  uint32_t nulls = 0;
  for (uint32_t i = tid; i < n; i += SB_NT) nulls += !valid.get(i);
  nulls = uint32_t(block_sum_u64e(cx, nulls));
  EffBin raw{b, nullptr};
  HashTab tab{};
  const uint32_t limit = n / 3 + 1;
  uint32_t unique = n ? hash_distinct(cx, raw, n, limit, &tab, nullptr, true) : 0;
  if (*cx.err) return kEncFail;
  uint32_t max_count = 0, top_first = 0;
  uint64_t total_unique_size = 0;
  if (unique != kNone && n) {
    hash_top(cx, tab, &max_count, &top_first);
    uint64_t s = 0;
    for (uint32_t h = tid; h <= tab.mask; h += SB_NT)
      if (tab.rep[h]) s += 8 + b.len(tab.rep[h] - 1);
    total_unique_size = block_sum_u64e(cx, s);
  }
  cx.ar = mark;
  const bool exact_small = unique != kNone;
  const bool one = exact_small && unique <= 1;
  int codec = o.def_codec;
  if ((o.force == SB_C_FREQ || o.force == SB_C_DICT || (o.force == SB_C_ONEVALUE && one)) && !e_forbidden(o, o.force)) {
    codec = o.force;
  } else if (o.ratio >= 0 && n) { // binary/mod.rs:293-348
    double max_ratio = o.ratio;
    const int order[3] = {SB_C_ONEVALUE, SB_C_FREQ, SB_C_DICT};
    const uint64_t total_bytes = backing_len + uint64_t(n + 1) * OW; // binary/mod.rs:268-270
    for (int k = 0; k < 3; ++k) {
      int c = order[k];
      if (e_forbidden(o, c)) continue;
      double r = 0.0;
      if (c == SB_C_ONEVALUE) r = one ? double(n) : 0.0;
      else if (c == SB_C_FREQ) {
        if (one) r = 0.0;
        else if (double(nulls) / double(n) >= 0.9) r = double(n - 1);
        else if (exact_small && double(max_count) / double(n) >= 0.9) r = double(n - 1);
      } else if (exact_small && uint64_t(unique) * 3 < n) {
        uint64_t after = total_unique_size + uint64_t(n) * (bits_needed(unique) / 8) + uint64_t(n) * 2 / 128;
        r = double(total_bytes) / double(after);
      }
      if (r > max_ratio) {
        max_ratio = r;
        codec = c;
        if (r == double(n)) break;
      }
    }
  }
  uint32_t payload = 0;
  if (codec <= SB_C_SNAPPY) {
    // Basic: hdr9 + rebased offsets, hdr9 + value bytes, same common codec (binary/mod.rs:43-81)
    const uint32_t obytes = (n + 1) * uint32_t(OW);
    uint8_t *tmp = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(obytes) + 16));
    if (!tmp) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    const int64_t first = b.off(0);
    for (uint32_t i = tid; i <= n; i += SB_NT) {
      int64_t r = b.off(i) - first;
      if (OW == 4) reinterpret_cast<int32_t *>(tmp)[i] = int32_t(r);
      else reinterpret_cast<int64_t *>(tmp)[i] = r;
    }
    __syncthreads();
    uint32_t w1 = enc_basic(cx, codec, tmp, obytes, body);
    if (w1 == kEncFail) return kEncFail;
    put_hdr9(out, codec, w1, obytes);
    const uint32_t vbytes = uint32_t(b.off(n) - first);
    uint8_t *h2 = body + w1;
    uint32_t w2 = enc_basic(cx, codec, b.data + first, vbytes, h2 + 9);
    if (w2 == kEncFail) return kEncFail;
    put_hdr9(h2, codec, w2, vbytes);
    __syncthreads();
    cx.ar = mark;
    return 9 + w1 + 9 + w2;
  }
  if (codec == SB_C_ONEVALUE) { // binary/one_value.rs:51-69: first valid row
    if (tid == 0) cx.bcast[2] = int(kNone);
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SB_NT)
      if (valid.get(i)) {
        atomicMin(reinterpret_cast<uint32_t *>(cx.bcast + 2), i);
        break;
      }
    __syncthreads();
    uint32_t fv = uint32_t(cx.bcast[2]);
    __syncthreads();
    uint32_t l = fv == kNone ? 0 : b.len(fv);
    if (tid == 0) st_le(body, l, 4);
    if (l) copy_bytes(body + 4, b.ptr(fv), l);
    payload = 4 + l;
  } else if (codec == SB_C_FREQ) { // binary/freq.rs:44-100
    const bool top_is_null = n && double(nulls) / double(n) >= 0.9;
    uint32_t top_row = kNone;
    if (!top_is_null && n) {
      if (unique == kNone) {
        HashTab ft{};
        uint32_t k = hash_distinct(cx, raw, n, n + 1, &ft, nullptr);
        if (*cx.err || k == kNone) return kEncFail;
        hash_top(cx, ft, &max_count, &top_first);
      }
      top_row = top_first;
    }
    uint32_t *rows = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!rows) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    uint32_t n_exc = 0;
    {
      constexpr uint32_t EPT = 4, CH = SB_NT * EPT;
      for (uint32_t c0 = 0; c0 < n; c0 += CH) {
        const uint32_t e0 = c0 + tid * EPT;
        uint32_t flags = 0, c = 0;
#pragma unroll
        for (uint32_t j = 0; j < EPT; ++j) {
          uint32_t i = e0 + j;
          if (i < n && valid.get(i) && (top_row == kNone || !raw.equal(i, top_row))) {
            flags |= 1u << j;
            ++c;
          }
        }
        uint32_t total;
        uint32_t pre = n_exc + block_excl_scan(c, cx.ws, &total);
#pragma unroll
        for (uint32_t j = 0; j < EPT; ++j)
          if ((flags >> j) & 1u) rows[pre++] = e0 + j;
        n_exc += total;
      }
    }
    __syncthreads();
    const uint32_t tl = top_row == kNone ? 0 : b.len(top_row);
    if (tid == 0) st_le(body, tl, 8);
    if (tl) copy_bytes(body + 8, b.ptr(top_row), tl);
    uint32_t bm = enc_roaring(cx, rows, n_exc, n, body + 8 + tl + 4);
    if (tid == 0) st_le(body + 8 + tl, bm, 4);
    uint32_t eb = put_entries(cx, b, rows, n_exc, body + 8 + tl + 4 + bm);
    payload = 8 + tl + 4 + bm + eb;
  } else if (codec == SB_C_DICT) { // binary/dict.rs:55-93
    // the per-row arrays live in scratch so that the table below gets the shared memory
    uint32_t *src = nullptr;
    if (valid.p && nulls) {
      src = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
      if (!src) {
        cx.flag(SB_NYI);
        return kEncFail;
      }
      fill_forward(cx, valid, n, true, src); // row 0 is interned even when null (:66-74)
    }
    EffBin acc{b, src};
    uint32_t *slot_of = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
    uint32_t *idx = static_cast<uint32_t *>(cx.ar.alloc_global(uint64_t(n) * 4 + 16));
    if (!slot_of || !idx) {
      cx.flag(SB_NYI);
      return kEncFail;
    }
    HashTab dt{};
    // null replacement only repeats slots of the page: no more keys than the statistics pass counted
    uint32_t k = hash_distinct(cx, acc, n, unique != kNone ? unique + 2 : n + 1, &dt, slot_of);
    if (*cx.err || k == kNone) return kEncFail;
    for (uint32_t i = tid; i < n; i += SB_NT) idx[i] = 0;
    __syncthreads();
    for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
      if (dt.rep[h]) idx[dt.first(h)] = 1;
    __syncthreads();
    {
      constexpr uint32_t EPT = 8, CH = SB_NT * EPT;
      uint32_t run = 0;
      for (uint32_t c0 = 0; c0 < n; c0 += CH) {
        const uint32_t e0 = c0 + tid * EPT;
        uint32_t f[EPT], c = 0;
#pragma unroll
        for (uint32_t j = 0; j < EPT; ++j) {
          f[j] = e0 + j < n ? idx[e0 + j] : 0;
          c += f[j];
        }
        uint32_t total;
        uint32_t pre = run + block_excl_scan(c, cx.ws, &total);
#pragma unroll
        for (uint32_t j = 0; j < EPT; ++j)
          if (e0 + j < n) {
            idx[e0 + j] = pre;
            pre += f[j];
          }
        run += total;
      }
    }
    __syncthreads();
    for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
      if (dt.rep[h]) dt.cnt[h] = idx[dt.first(h)];
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SB_NT) slot_of[i] = dt.cnt[slot_of[i]]; // row -> id
    __syncthreads();
    // id -> source row of the first occurrence, in id order
    uint32_t *id_row = idx; // idx is free now: rank array consumed
    for (uint32_t h = tid; h <= dt.mask; h += SB_NT)
      if (dt.rep[h]) id_row[dt.cnt[h]] = acc.row(dt.first(h));
    __syncthreads();
    cx.ar.s_cur = mark.s_cur; // the table is done: its shared memory goes to the index page's own statistics
    EOpts sub = o;
    sub.forbidden |= 1u << SB_C_DICT;
    uint32_t used = enc_fixed<1>(cx, Vals{reinterpret_cast<const uint8_t *>(slot_of), 4}, TC_UINT, Bits{nullptr, 0}, n, sub, body);
    if (used == kEncFail) return kEncFail;
    if (tid == 0) st_le(body + used, k, 4);
    uint32_t eb = put_entries(cx, b, id_row, k, body + used + 4);
    payload = used + 4 + eb;
  } else {
    cx.flag(SB_OUT_OF_SPEC);
    return kEncFail;
  }
  if (*cx.err) return kEncFail;
  put_hdr9(out, codec, payload, uint32_t(backing_len)); // array.values().len() (binary/mod.rs:88, App. C8)
  __syncthreads();
  cx.ar = mark;
  return 9 + payload;
}

} // namespace sb
