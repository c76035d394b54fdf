// sb_binary.cuh -- Binary / Utf8 value blocks (src/compression/binary/*).
//
// Output size of a binary page is data dependent, so binary columns are decoded in two
// passes: sb_size_kernel computes the value bytes every page appends (and,
// for Dict / Freq, records where each length-prefixed entry sits so the serial
// `[u64 len][bytes]` chain is walked only once), sb_scan_kernel turns those sizes into the
// page's first output byte, and the main decode kernel writes offsets and value bytes.
#pragma once
#include "sb_decode.cuh"

namespace sb {

struct __align__(8) BinEntry {
  uint32_t pos; // byte position of the entry's payload inside the page
  uint32_t len;
};

template <int OW> struct OffT;
template <> struct OffT<4> { using T = int32_t; };
template <> struct OffT<8> { using T = int64_t; };

// Walks k `[u64 len][bytes]` entries starting at page position `start` (thread 0 only;
// inherently serial: the next header sits after the previous payload).  On success
// bcast[0] = 0, bcast[1] = end position, and the 64-bit payload total is in bcast[2..3].
__device__ __forceinline__ bool walk_entries(Dctx &cx, const uint8_t *page, uint32_t page_len, uint32_t start, uint32_t k,
                                             BinEntry *tab, uint32_t *end_pos, uint64_t *total) {
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t pos = start;
    uint64_t sum = 0;
    int rc = 0;
    for (uint32_t e = 0; e < k; ++e) {
      if (page_len - pos < 8) {
        rc = SB_IO;
        break;
      }
      uint64_t len = ld_u64u(page + pos);
      pos += 8;
      if (len > uint64_t(page_len - pos)) { // general_err!("data size is less than {}")
        rc = SB_OUT_OF_SPEC;
        break;
      }
      tab[e].pos = pos;
      tab[e].len = uint32_t(len);
      pos += uint32_t(len);
      sum += len;
    }
    cx.bcast[0] = rc;
    cx.bcast[1] = int(pos);
    cx.bcast[2] = int(uint32_t(sum));
    cx.bcast[3] = int(uint32_t(sum >> 32));
  }
  __syncthreads();
  int rc = cx.bcast[0];
  *end_pos = uint32_t(cx.bcast[1]);
  *total = uint64_t(uint32_t(cx.bcast[2])) | (uint64_t(uint32_t(cx.bcast[3])) << 32);
  __syncthreads();
  if (rc) {
    cx.flag(rc);
    return false;
  }
  return true;
}

__device__ __forceinline__ uint64_t block_sum_u64(Dctx &cx, uint64_t v) {
  // two 32-bit block reductions through the scan workspace
  uint32_t lo = uint32_t(v), hi = uint32_t(v >> 32);
  __shared__ unsigned long long s_acc;
  __syncthreads();
  if (threadIdx.x == 0) s_acc = 0;
  __syncthreads();
  unsigned long long w = (unsigned long long)lo | ((unsigned long long)hi << 32);
  // warp reduce then one atomic per warp
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xffffffffu, w, d);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc, w);
  __syncthreads();
  uint64_t r = s_acc;
  __syncthreads();
  return r;
}

// Roaring header of Freq.  Uniform across the CTA.
struct RoaringHdr {
  const uint8_t *rb;
  uint32_t bm, ncont, n_exc;
};
__device__ __forceinline__ bool roaring_parse(Dctx &cx, const uint8_t *rb, uint32_t bm, RoaringHdr *h) {
  if (bm < 8) {
    cx.flag(SB_IO);
    return false;
  }
  uint32_t cookie = ld_u32u(rb);
  if (cookie != 12346u) {
    cx.flag((cookie & 0xffff) == 12347u ? SB_NYI : SB_IO);
    return false;
  }
  uint32_t ncont = ld_u32u(rb + 4);
  if (ncont > 65536 || uint64_t(ncont) * 8 + 8 > bm) {
    cx.flag(SB_IO);
    return false;
  }
  uint64_t off = 8 + uint64_t(ncont) * 8, tot = 0;
  for (uint32_t c = 0; c < ncont; ++c) {
    uint32_t card = ld_u16u(rb + 8 + 4 * c + 2) + 1;
    off += card > 4096 ? 8192 : card * 2;
    tot += card;
  }
  if (off > bm || tot > 0xffffffffull) {
    cx.flag(SB_IO);
    return false;
  }
  h->rb = rb;
  h->bm = bm;
  h->ncont = ncont;
  h->n_exc = uint32_t(tot);
  return true;
}
// rank[row] = 1 + exception index for exception rows (rank[] pre-zeroed, n entries)
__device__ __forceinline__ void roaring_scatter_ranks(Dctx &cx, const RoaringHdr &h, uint32_t n, uint32_t *rank) {
  const uint32_t tid = threadIdx.x;
  uint32_t rank_base = 0;
  uint64_t off = 8 + uint64_t(h.ncont) * 8;
  for (uint32_t c = 0; c < h.ncont; ++c) {
    uint32_t key = ld_u16u(h.rb + 8 + 4 * c), card = ld_u16u(h.rb + 8 + 4 * c + 2) + 1;
    const uint8_t *data = h.rb + off;
    if (card <= 4096) {
      for (uint32_t j = tid; j < card; j += SB_NT) {
        uint32_t row = (key << 16) | ld_u16u(data + 2 * j);
        if (row < n) rank[row] = rank_base + j + 1;
      }
      off += card * 2;
    } else {
      uint64_t wds[8];
      uint32_t cnt = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        wds[j] = ld_u64u(data + 8 * (tid * 8 + j));
        cnt += __popcll(wds[j]);
      }
      uint32_t total;
      uint32_t r = rank_base + block_excl_scan(cnt, cx.ws, &total);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint64_t bits = wds[j];
        while (bits) {
          uint32_t b = __ffsll((long long)bits) - 1;
          bits &= bits - 1;
          uint32_t row = (key << 16) | ((tid * 8 + j) * 64 + b);
          if (row < n) rank[row] = r + 1;
          ++r;
        }
      }
      off += 8192;
    }
    rank_base += card;
  }
}

// ---- parsed view of a binary value block (shared by the size and the decode pass) -------
struct BinBlock {
  int codec;
  const uint8_t *body;
  uint32_t body_avail, compressed;
};
__device__ __forceinline__ bool bin_header(Dctx &cx, const uint8_t *src, uint32_t avail, BinBlock *b) {
  if (avail < 9) {
    cx.flag(SB_IO);
    return false;
  }
  b->codec = src[0];
  b->compressed = ld_u32u(src + 1);
  b->body = src + 9;
  b->body_avail = avail - 9;
  if (b->compressed > b->body_avail) {
    cx.flag(SB_IO);
    return false;
  }
  return true;
}

// =========================================================================================
// pass 1: value bytes of one page.  `page`/`page_len` = whole page, `vb` = value block
// offset; `tab` = this page's slice of the entry table (global).
// =========================================================================================
__device__ bool binary_page_size(Dctx &cx, const uint8_t *page, uint32_t page_len, uint32_t vb, uint32_t n, BinEntry *tab,
                                 uint64_t *out_bytes, uint32_t *val_pos, uint32_t *n_ent, const PageAux *pre = nullptr) {
  *val_pos = 0;
  *n_ent = 0;
  BinBlock b;
  if (!bin_header(cx, page + vb, page_len - vb, &b)) return false;
  const uint32_t body_pos = vb + 9;
  switch (b.codec) {
  case SB_C_NONE:
  case SB_C_LZ4:
  case SB_C_ZSTD:
  case SB_C_SNAPPY: { // second hdr9 carries the value bytes (binary/mod.rs:147,159)
    if (b.body_avail - b.compressed < 9) {
      cx.flag(SB_IO);
      return false;
    }
    *out_bytes = ld_u32u(b.body + b.compressed + 5);
    // plain value bytes can be copied by several work items (tiles): position of the first one
    const uint32_t c2 = ld_u32u(b.body + b.compressed + 1);
    if (b.codec == SB_C_NONE && c2 == uint32_t(*out_bytes) && c2 <= b.body_avail - b.compressed - 9)
      *val_pos = body_pos + b.compressed + 9;
    return true;
  }
  case SB_C_ONEVALUE: { // binary/one_value.rs:71-99
    if (b.body_avail < 4) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t len = ld_u32u(b.body);
    if (len > b.body_avail - 4) {
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    *out_bytes = uint64_t(len) * n;
    return true;
  }
  case SB_C_DICT: { // binary/dict.rs:95-141
    Arena mark = cx.ar;
    uint32_t *idx = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!idx) {
      cx.flag(SB_NYI);
      return false;
    }
    uint32_t used = 0;
    if (!decode_fixed<1>(cx, b.body, b.body_avail, n, 4, false, reinterpret_cast<uint8_t *>(idx), &used)) return false;
    if (b.body_avail - used < 4) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t k = ld_u32u(b.body + used);
    if (uint64_t(k) * 8 > uint64_t(b.body_avail - used - 4)) { // every entry needs its 8-byte header
      cx.flag(SB_IO);
      return false;
    }
    uint32_t end;
    uint64_t tot;
    // sb_dict_walk_kernel may have walked this dictionary already (one warp per page, all pages side by side): its
    // record counts only if it describes the dictionary found here; `tab` then holds the k entries
    if (pre && pre->pad == 1 && pre->cnt[0] == body_pos + used + 4 && pre->cnt[1] == k) {
      end = pre->cnt[2];
      tot = pre->base[0];
      __syncthreads(); // idx[] complete before the sum below (walk_entries' barrier otherwise)
    } else if (!walk_entries(cx, page, page_len, body_pos + used + 4, k, tab, &end, &tot)) {
      return false;
    }
    (void)end;
    (void)tot;
    *n_ent = k;
    uint64_t sum = 0;
    for (uint32_t i = threadIdx.x; i < n; i += SB_NT) {
      uint32_t id = idx[i];
      if (id < k) sum += tab[id].len;
      else cx.flag(SB_PANIC);
    }
    *out_bytes = block_sum_u64(cx, sum);
    cx.ar = mark;
    return true;
  }
  case SB_C_FREQ: { // binary/freq.rs:102-145
    if (b.body_avail < 8) {
      cx.flag(SB_IO);
      return false;
    }
    uint64_t top_len = ld_u64u(b.body);
    if (top_len > uint64_t(b.body_avail - 8) || b.body_avail - 8 - uint32_t(top_len) < 4) {
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    uint32_t p = 8 + uint32_t(top_len);
    uint32_t bm = ld_u32u(b.body + p);
    p += 4;
    if (bm > b.body_avail - p) {
      cx.flag(SB_PANIC);
      return false;
    }
    RoaringHdr h;
    if (!roaring_parse(cx, b.body + p, bm, &h)) return false;
    p += bm;
    // exceptions with row >= n are never visited by the reference's `for i in 0..length`
    // loop; count the visited ones through the rank table
    Arena mark = cx.ar;
    uint32_t *rank = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!rank) {
      cx.flag(SB_NYI);
      return false;
    }
    for (uint32_t i = threadIdx.x; i < n; i += SB_NT) rank[i] = 0;
    __syncthreads();
    roaring_scatter_ranks(cx, h, n, rank);
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t i = threadIdx.x; i < n; i += SB_NT) cnt += rank[i] != 0;
    uint32_t n_vis = uint32_t(block_sum_u64(cx, cnt));
    uint32_t end;
    uint64_t tot;
    if (uint64_t(n_vis) * 8 > uint64_t(b.body_avail - p)) {
      cx.flag(SB_IO);
      return false;
    }
    if (!walk_entries(cx, page, page_len, body_pos + p, n_vis, tab, &end, &tot)) return false;
    *n_ent = n_vis;
    *out_bytes = tot + top_len * uint64_t(n - n_vis);
    cx.ar = mark;
    return true;
  }
  default: cx.flag(SB_OUT_OF_SPEC); return false;
  }
}

// =========================================================================================
// pass 2: offsets + value bytes
// =========================================================================================
// rows -> (len, src) provider; emits offsets and copies bytes.  Rows are processed in
// chunks of SB_NT * RPT with a block scan of the row lengths.
template <int OW, class RowSrc>
__device__ void emit_rows(Dctx &cx, uint32_t n, RowSrc &rs, typename OffT<OW>::T *out_off /* &offsets[elem] */,
                          uint8_t *out_val /* values + page base */, uint64_t base, bool first, uint64_t limit /* value bytes the plan pass sized */) {
  using O = typename OffT<OW>::T;
  constexpr uint32_t RPT = 4, CH = SB_NT * RPT;
  const uint32_t tid = threadIdx.x;
  if (first && tid == 0) out_off[0] = O(0);
  uint64_t run = 0; // bytes emitted by previous chunks
  for (uint32_t r0 = 0; r0 < n; r0 += CH) {
    uint32_t lens[RPT], sum = 0;
    const uint8_t *ptrs[RPT];
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j) { // one lookup per row: (payload pointer, length), all four in flight
      uint32_t r = r0 + tid * RPT + j;
      lens[j] = 0;
      ptrs[j] = nullptr;
      if (r < n) rs.get(r, &ptrs[j], &lens[j]);
      sum += lens[j];
    }
    uint32_t total;
    uint64_t pre = run + block_excl_scan(sum, cx.ws, &total);
    O offs[RPT];
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j) {
      const uint8_t *s = ptrs[j];
      uint8_t *d = out_val + pre;
      if (pre + lens[j] <= limit) {
        for (uint32_t i = 0; i < lens[j]; ++i) d[i] = s[i];
      } else {
        cx.flag(SB_PANIC); // the page changed between the two passes / inconsistent exception ranks: never write past the plan
      }
      pre += lens[j];
      offs[j] = O(base + pre);
    }
    const uint32_t r = r0 + tid * RPT;
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j)
      if (r + j < n) out_off[r + j + 1] = offs[j];
    run += total;
  }
}

struct RowsDict {
  const uint32_t *idx;
  const BinEntry *tab;
  const uint8_t *page;
  uint32_t k;
  __device__ __forceinline__ void get(uint32_t r, const uint8_t **p, uint32_t *l) const {
    uint32_t id = idx[r];
    uint2 e = id < k ? *reinterpret_cast<const uint2 *>(tab + id) : make_uint2(0u, 0u); // {pos, len}
    *p = page + e.x;
    *l = e.y;
  }
};
struct RowsFreq {
  const uint32_t *rank;
  const BinEntry *tab;
  const uint8_t *page;
  const uint8_t *top;
  uint32_t top_len;
  uint32_t n_ent; // entries the plan pass recorded (ranks beyond it: duplicate / unsorted bitmap values)
  int *err;
  __device__ __forceinline__ void get(uint32_t r, const uint8_t **p, uint32_t *l) const {
    uint32_t k = rank[r];
    if (k > n_ent) {
      atomicCAS(err, 0, int(SB_PANIC));
      *p = top;
      *l = 0;
    } else if (k) {
      uint2 e = *reinterpret_cast<const uint2 *>(tab + (k - 1));
      *p = page + e.x;
      *l = e.y;
    } else {
      *p = top;
      *l = top_len;
    }
  }
};
struct RowsConst {
  const uint8_t *val;
  uint32_t vlen;
  __device__ __forceinline__ void get(uint32_t, const uint8_t **p, uint32_t *l) const {
    *p = val;
    *l = vlen;
  }
};

// The plan pass left the (pos, len) table of the page's dictionary / exception entries in global
// memory: copy it next to the page in shared memory when it fits (8 bytes per entry), so the
// per-row lookups of emit_rows are shared-memory loads.  The caller synchronises.
__device__ __forceinline__ const BinEntry *stage_entries(Dctx &cx, const BinEntry *tab, uint32_t k) {
  BinEntry *s = static_cast<BinEntry *>(cx.ar.alloc_shared(uint64_t(k) * sizeof(BinEntry)));
  if (!s) return tab;
  for (uint32_t i = threadIdx.x; i < k; i += SB_NT) reinterpret_cast<uint2 *>(s)[i] = reinterpret_cast<const uint2 *>(tab)[i];
  return s;
}

// decompress_binary (binary/mod.rs:95-183).  out_off = &offsets[out_elem]; out_val = values
// + out_byte; `base` = out_byte (the last offset already in the column); `first` = the
// column's offsets are still empty (push the initial 0, binary/dict.rs:122-127).
template <int OW>
__device__ bool decode_binary(Dctx &cx, const uint8_t *page, uint32_t page_len, uint32_t vb, uint32_t n,
                              typename OffT<OW>::T *out_off, uint8_t *out_val, uint64_t base, bool first,
                              const BinEntry *tab, bool values_tiled, uint32_t n_ent, uint64_t limit) {
  using O = typename OffT<OW>::T;
  BinBlock b;
  if (!bin_header(cx, page + vb, page_len - vb, &b)) return false;
  const uint32_t tid = threadIdx.x;
  switch (b.codec) {
  case SB_C_NONE:
  case SB_C_LZ4:
  case SB_C_ZSTD:
  case SB_C_SNAPPY: { // Basic: hdr9(offsets) + hdr9(values) through the same common codec (mod.rs:120-173)
    const uint64_t obytes = uint64_t(n + 1) * OW;
    const uint8_t *raw = b.body;
    Arena mark = cx.ar;
    if (b.codec != SB_C_NONE) {
      uint8_t *tmp = static_cast<uint8_t *>(cx.ar.alloc(obytes + 16));
      if (!tmp) {
        cx.flag(SB_NYI);
        return false;
      }
      if (!dec_basic(cx, b.codec, b.body, b.compressed, tmp, obytes)) return false;
      raw = tmp;
    } else if (uint64_t(b.compressed) != obytes) {
      cx.flag(SB_PANIC);
      return false;
    }
    // offsets: first page keeps raw[0..n]; later pages drop raw[0] and rebase by `last`
    {
      // 16-byte vectors: one unaligned 16-byte read of the page (two aligned loads + funnel
      // shifts), add the rebase, one aligned 16-byte store; scalar head / tail
      constexpr uint32_t E = 16 / OW;
      const uint32_t j0 = first ? 0u : 1u;
      uint32_t head = uint32_t((16 - (uintptr_t(out_off + j0) & 15)) & 15) / OW; // elements before the first aligned vector
      const uint32_t total = n + 1 - j0;
      head = min(head, total);
      const uint32_t nvec = (total - head) / E;
      auto scalar = [&](uint32_t j) {
        uint64_t v = OW == 4 ? uint64_t(int64_t(int32_t(ld_u32u(raw + uint64_t(j) * 4)))) : ld_u64u(raw + uint64_t(j) * 8);
        out_off[j] = O(base + v);
      };
      for (uint32_t j = j0 + tid; j < j0 + head; j += SB_NT) scalar(j);
      auto rebase = [&](uint4 w) {
        if (OW == 4) {
          const uint32_t b32 = uint32_t(base);
          w.x += b32, w.y += b32, w.z += b32, w.w += b32;
        } else {
          uint64_t a = (uint64_t(w.y) << 32 | w.x) + base, c = (uint64_t(w.w) << 32 | w.z) + base;
          w.x = uint32_t(a), w.y = uint32_t(a >> 32), w.z = uint32_t(c), w.w = uint32_t(c >> 32);
        }
        return w;
      };
      const uint8_t *rv = raw + uint64_t(j0 + head) * OW;
      stream_vec(cx, reinterpret_cast<uint4 *>(out_off + j0 + head), rv, nvec, [&](uint4 w, uint64_t) { return rebase(w); });
      for (uint32_t j = j0 + head + nvec * E + tid; j <= n; j += SB_NT) scalar(j);
    }
    __syncthreads();
    cx.ar = mark;
    const uint8_t *h2 = b.body + b.compressed;
    uint32_t rest = b.body_avail - b.compressed;
    if (rest < 9) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t c2 = ld_u32u(h2 + 1), u2 = ld_u32u(h2 + 5);
    if (c2 > rest - 9) {
      cx.flag(SB_IO);
      return false;
    }
    if (uint64_t(u2) != limit) { // the size the plan pass reserved for this page
      cx.flag(SB_PANIC);
      return false;
    }
    if (values_tiled && b.codec == SB_C_NONE && c2 == u2) return true; // value bytes: tiles 1.. of this page
    return dec_basic(cx, b.codec, h2 + 9, c2, out_val, u2);
  }
  case SB_C_ONEVALUE: {
    if (b.body_avail < 4) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t len = ld_u32u(b.body);
    if (len > b.body_avail - 4) {
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    RowsConst rs{b.body + 4, len};
    emit_rows<OW>(cx, n, rs, out_off, out_val, base, first, limit);
    return true;
  }
  case SB_C_DICT: {
    Arena mark = cx.ar;
    uint32_t *idx = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!idx) {
      cx.flag(SB_NYI);
      return false;
    }
    uint32_t used = 0;
    if (!decode_fixed<1>(cx, b.body, b.body_avail, n, 4, false, reinterpret_cast<uint8_t *>(idx), &used)) return false;
    if (b.body_avail - used < 4) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t k = ld_u32u(b.body + used);
    if (k != n_ent) { // not the dictionary the plan pass walked
      cx.flag(SB_PANIC);
      return false;
    }
    tab = stage_entries(cx, tab, k);
    __syncthreads();
    RowsDict rs{idx, tab, page, k};
    emit_rows<OW>(cx, n, rs, out_off, out_val, base, first, limit);
    __syncthreads();
    cx.ar = mark;
    return true;
  }
  case SB_C_FREQ: {
    if (b.body_avail < 8) {
      cx.flag(SB_IO);
      return false;
    }
    uint64_t top_len = ld_u64u(b.body);
    if (top_len > uint64_t(b.body_avail - 8) || b.body_avail - 8 - uint32_t(top_len) < 4) {
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    uint32_t p = 8 + uint32_t(top_len);
    uint32_t bm = ld_u32u(b.body + p);
    p += 4;
    if (bm > b.body_avail - p) {
      cx.flag(SB_PANIC);
      return false;
    }
    RoaringHdr h;
    if (!roaring_parse(cx, b.body + p, bm, &h)) return false;
    Arena mark = cx.ar;
    uint32_t *rank = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!rank) {
      cx.flag(SB_NYI);
      return false;
    }
    for (uint32_t i = tid; i < n; i += SB_NT) rank[i] = 0;
    __syncthreads();
    roaring_scatter_ranks(cx, h, n, rank);
    __syncthreads();
    // exception e is the e-th VISITED exception row: ranks of rows < n are dense because
    // the bitmap is sorted, except for (malformed) rows >= n which the size pass ignored too
    tab = stage_entries(cx, tab, n_ent); // the entries the plan pass recorded (at most one per row)
    __syncthreads();
    RowsFreq rs{rank, tab, page, b.body + 8, uint32_t(top_len), n_ent, cx.err};
    emit_rows<OW>(cx, n, rs, out_off, out_val, base, first, limit);
    __syncthreads();
    cx.ar = mark;
    return true;
  }
  default: cx.flag(SB_OUT_OF_SPEC); return false;
  }
}

} // namespace sb
