// sb_decode.cuh -- CTA-cooperative decoders for strawboat value blocks.
//
// Every routine is executed by all SB_NT threads of the CTA with uniform arguments.
// `src` is a byte pointer of arbitrary alignment into the staged page (shared memory) or,
// for oversized pages, into global memory; `dst` is the element-aligned destination: the
// Arrow values buffer in HBM for the top-level block, or an arena buffer for nested
// blocks (Dict indices, Freq exceptions).  Routines return false on a uniform failure
// after flagging cx; data-dependent per-thread problems are flagged and clamped.
//
// Layouts follow SURVEY.md Appendix A; each decoder cites the reference function whose
// output it reproduces bit for bit.
#pragma once
#include "sb_common.cuh"
#include "sb_lz4.cuh"
#include "sb_zstd.cuh"

namespace sb {

// ------------------------------------------------------------------------------------
// emit: write elements [lo, hi) of an element array through a generator, using 16-byte
// vector stores for the aligned body.  Gen provides seek(i) and next() (consecutive i).
// ------------------------------------------------------------------------------------
template <int W, class Gen> __device__ __forceinline__ void emit(uint8_t *dst, uint32_t lo, uint32_t hi, Gen &g) {
  using T = typename Elem<W>::T;
  const uint32_t tid = threadIdx.x;
  if (hi <= lo) return;
  T *out = reinterpret_cast<T *>(dst);
  if constexpr (W >= 16) { // one element = one or two 16-byte vectors; dst is 16-byte aligned (column buffer / arena)
    for (uint32_t i = lo + tid; i < hi; i += SB_NT) {
      g.seek(i);
      out[i] = g.next();
    }
    return;
  }
  constexpr uint32_t E = W >= 16 ? 1 : 16 / W;
  uint32_t mis = uint32_t((16 - ((uintptr_t(dst) + uint64_t(lo) * W) & 15)) & 15) / W;
  uint32_t head_end = min(hi, lo + mis);
  for (uint32_t i = lo + tid; i < head_end; i += SB_NT) {
    g.seek(i);
    out[i] = g.next();
  }
  uint32_t nvec = (hi - head_end) / E;
  for (uint32_t v = tid; v < nvec; v += SB_NT) {
    uint32_t i = head_end + v * E;
    g.seek(i);
    union {
      uint4 q;
      T t[E];
    } u;
#pragma unroll
    for (uint32_t j = 0; j < E; ++j) u.t[j] = g.next();
    *reinterpret_cast<uint4 *>(dst + uint64_t(i) * W) = u.q;
  }
  for (uint32_t i = head_end + nvec * E + tid; i < hi; i += SB_NT) {
    g.seek(i);
    out[i] = g.next();
  }
}

// LZ4 value block into `dst` (one instance for every caller: the scanner / mover code is large).
__device__ __noinline__ bool dec_lz4_block(Dctx &cx, const uint8_t *src, uint32_t clen, uint8_t *dst, uint64_t out_bytes) {
    // Nested LZ4 blocks (Dict indices, Freq exceptions, boolean / binary buffers).  Blocks of
    // some size run on the scanner / mover pair of sb_lz4.cuh (warps 0 and 1 of this CTA; the
    // rings come out of the shared arena, the stream is read from the page's global copy, the
    // output must be global because the mover re-reads far match sources with ld.global).
    Arena mark = cx.ar;
    Lz4Shared *sh = nullptr;
    uint8_t *gdst = nullptr;
    if (clen >= 96 && out_bytes <= SB_LZ4_MAXPOS / 2 && cx.page_g != nullptr) {
      gdst = __isGlobal(dst) ? dst : static_cast<uint8_t *>(cx.ar.alloc_global(out_bytes + 16));
      if (gdst) sh = static_cast<Lz4Shared *>(cx.ar.alloc_shared(sizeof(Lz4Shared)));
    }
    __syncthreads();
    if (sh) {
      if (threadIdx.x == 0) {
        sh->produced = sh->in_ready = sh->consumed = sh->m_q = sh->abort = 0;
        cx.bcast[0] = 0;
      }
      __syncthreads();
      const uint8_t *gsrc = cx.page_g + (src - cx.page_s);
      int rc = 0;
      if (threadIdx.x < 32) rc = lz4_scan(gsrc, clen, sh);
      else if (threadIdx.x < 64) rc = lz4_move(gdst, uint32_t(out_bytes), uint32_t(uintptr_t(gsrc) & 15) + clen, sh);
      if (rc) cx.bcast[0] = rc;
      __syncthreads();
      rc = cx.bcast[0];
      if (rc == 0 && gdst != dst) {
        __threadfence_block();
        copy_bytes(dst, gdst, out_bytes);
      }
      cx.ar = mark;
      __syncthreads();
      if (rc) {
        cx.flag(rc);
        return false;
      }
      return true;
    }
    cx.ar = mark;
    if (threadIdx.x < 32) {
      FlatOut fo{dst};
      int rc = lz4_decode_warp2(src, clen, fo, uint32_t(out_bytes));
      if (threadIdx.x == 0) cx.bcast[0] = rc;
    }
    __syncthreads();
    int rc = cx.bcast[0];
    if (rc) {
      cx.flag(rc);
      return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------
// Snappy raw block (basic.rs:98-105 -> snap::raw::Decoder::decompress; format: google/snappy
// format_description.txt).  [varint n] then elements: literal (tag & 3 == 0), copy with 11-bit /
// 16-bit / 32-bit offset.  One warp walks the elements (the tag chain is serial) and moves the
// bytes of each element lane-parallel.  Returns 0 or SB_EXTERNAL, uniform over the warp.
// ------------------------------------------------------------------------------------
__device__ int snappy_decode_warp(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t dlen) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t ip = 0, n = 0;
  for (uint32_t shift = 0;; shift += 7) {
    if (ip >= clen || shift > 28) return SB_EXTERNAL;
    const uint32_t b = src[ip++];
    n |= (b & 0x7fu) << shift;
    if (!(b & 0x80u)) break;
  }
  if (n != dlen) return SB_EXTERNAL; // snap: BufferTooSmall when larger; a shorter block would leave rows undefined
  uint32_t op = 0;
  while (ip < clen) {
    const uint32_t tag = src[ip++];
    uint32_t len, offset;
    if ((tag & 3u) == 0) {
      len = (tag >> 2) + 1;
      if (len > 60) {
        const uint32_t nb = len - 60;
        if (clen - ip < nb) return SB_EXTERNAL;
        uint32_t l = 0;
        for (uint32_t k = 0; k < nb; ++k) l |= uint32_t(src[ip + k]) << (8 * k);
        ip += nb;
        if (l == 0xffffffffu) return SB_EXTERNAL;
        len = l + 1;
      }
      if (len > clen - ip || len > dlen - op) return SB_EXTERNAL;
      for (uint32_t i = lane; i < len; i += 32) dst[op + i] = src[ip + i];
      __syncwarp();
      ip += len;
      op += len;
      continue;
    }
    if ((tag & 3u) == 1) {
      if (ip >= clen) return SB_EXTERNAL;
      len = 4 + ((tag >> 2) & 7u);
      offset = ((tag >> 5) << 8) | src[ip++];
    } else if ((tag & 3u) == 2) {
      if (clen - ip < 2) return SB_EXTERNAL;
      len = (tag >> 2) + 1;
      offset = uint32_t(src[ip]) | (uint32_t(src[ip + 1]) << 8);
      ip += 2;
    } else {
      if (clen - ip < 4) return SB_EXTERNAL;
      len = (tag >> 2) + 1;
      offset = uint32_t(src[ip]) | (uint32_t(src[ip + 1]) << 8) | (uint32_t(src[ip + 2]) << 16) | (uint32_t(src[ip + 3]) << 24);
      ip += 4;
    }
    if (offset == 0 || offset > op || len > dlen - op) return SB_EXTERNAL;
    // len <= 64; everything before op is final: an overlapping copy replicates the `offset` bytes before op
    for (uint32_t i = lane; i < len; i += 32) {
      const uint32_t b = dst[op - offset + (offset < len ? i % offset : i)];
      dst[op + i] = uint8_t(b);
    }
    __syncwarp();
    op += len;
  }
  return op == dlen ? 0 : SB_EXTERNAL;
}
__device__ __noinline__ bool dec_snappy_block(Dctx &cx, const uint8_t *src, uint32_t clen, uint8_t *dst, uint64_t out_bytes) {
  __syncthreads();
  if (threadIdx.x < 32) {
    int rc = out_bytes > 0xffffffffull ? int(SB_EXTERNAL) : snappy_decode_warp(src, clen, dst, uint32_t(out_bytes));
    if (threadIdx.x == 0) cx.bcast[0] = rc;
  }
  __syncthreads();
  const int rc = cx.bcast[0];
  __syncthreads();
  if (rc) {
    cx.flag(rc);
    return false;
  }
  return true;
}

// Zstandard frame (basic.rs:93-97): warp 0 of the CTA, tables from the arena, literals in the global scratch
__device__ __noinline__ bool dec_zstd_block(Dctx &cx, const uint8_t *src, uint32_t clen, uint8_t *dst, uint64_t out_bytes) {
  Arena mark = cx.ar;
  ZstdTables *T = static_cast<ZstdTables *>(cx.ar.alloc(sizeof(ZstdTables)));
  uint8_t *lits = static_cast<uint8_t *>(cx.ar.alloc_global(min(out_bytes, uint64_t(128u << 10)) + 32));
  if (!T || !lits || out_bytes > 0xffffffffull) {
    cx.ar = mark;
    cx.flag(SB_NYI);
    return false;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    int rc = zstd_decode_warp(src, clen, dst, uint32_t(out_bytes), T, lits);
    if (threadIdx.x == 0) cx.bcast[0] = rc;
  }
  __syncthreads();
  const int rc = cx.bcast[0];
  __syncthreads();
  cx.ar = mark;
  if (rc) {
    cx.flag(rc);
    return false;
  }
  return true;
}

// Basic codecs (CommonCompression::decompress, basic.rs:62-72) into `dst`.
__device__ __forceinline__ bool dec_basic(Dctx &cx, int codec, const uint8_t *src, uint32_t clen, uint8_t *dst,
                                          uint64_t out_bytes) {
  if (codec == SB_C_NONE) {
    if (uint64_t(clen) != out_bytes) { // copy_from_slice length mismatch panics (basic.rs:68)
      cx.flag(SB_PANIC);
      return false;
    }
    stream_copy(cx, dst, src, out_bytes);
    return true;
  }
  if (codec == SB_C_LZ4) return dec_lz4_block(cx, src, clen, dst, out_bytes);
  if (codec == SB_C_SNAPPY) return dec_snappy_block(cx, src, clen, dst, out_bytes);
  if (codec == SB_C_ZSTD) return dec_zstd_block(cx, src, clen, dst, out_bytes);
  cx.flag(SB_OUT_OF_SPEC);
  return false;
}

// ------------------------------------------------------------------------------------
// OneValue (integer/one_value.rs:77-94): payload = one W-byte value
// ------------------------------------------------------------------------------------
template <int W> struct GenConst {
  typename Elem<W>::T v;
  __device__ __forceinline__ void seek(uint32_t) {}
  __device__ __forceinline__ typename Elem<W>::T next() { return v; }
};
template <int W> __device__ __forceinline__ bool dec_onevalue(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t lo,
                                                              uint32_t hi, uint8_t *dst) {
  if (avail < uint32_t(W)) {
    cx.flag(SB_IO);
    return false;
  }
  GenConst<W> g{ld_elem_u<W>(src)};
  emit<W>(dst, lo, hi, g);
  return true;
}

// ------------------------------------------------------------------------------------
// RLE (integer/rle.rs:106-134, double/rle.rs:105-135): (u32 run, W-byte value)*
// Runs are consumed in chunks of SB_NT*8; a block scan turns run lengths into start
// positions; each output vector binary-searches its first run and then walks forward.
// ------------------------------------------------------------------------------------
template <int W> struct GenRle {
  const uint32_t *starts; // chunk-local run start positions (absolute element index), CH+1 entries
  const uint8_t *runs;    // first run of the chunk
  uint32_t nr;            // runs in chunk
  uint32_t r, i;
  typename Elem<W>::T cur;
  __device__ __forceinline__ void seek(uint32_t idx) {
    uint32_t lo = 0, hi = nr; // largest r with starts[r] <= idx
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (starts[mid] <= idx) lo = mid;
      else hi = mid;
    }
    r = lo;
    i = idx;
    cur = ld_elem_u<W>(runs + uint64_t(r) * (4 + W) + 4);
  }
  __device__ __forceinline__ typename Elem<W>::T next() {
    if (i >= starts[r + 1]) {
      do ++r;
      while (r + 1 < nr && i >= starts[r + 1]);
      cur = ld_elem_u<W>(runs + uint64_t(r) * (4 + W) + 4);
    }
    ++i;
    return cur;
  }
};
template <int W> __device__ bool dec_rle(Dctx &cx, const uint8_t *src, uint32_t plen, uint32_t n, uint8_t *dst) {
  constexpr uint32_t STR = 4 + W, RPT = 8, CH = SB_NT * RPT;
  const uint32_t tid = threadIdx.x;
  const uint32_t nruns = plen / STR;
  Arena mark = cx.ar;
  uint32_t *starts = static_cast<uint32_t *>(cx.ar.alloc((CH + 1) * 4));
  if (!starts) {
    cx.flag(SB_NYI);
    return false;
  }
  uint32_t done = 0;
  for (uint32_t r0 = 0; r0 < nruns && done < n; r0 += CH) {
    uint32_t lens[RPT], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j) {
      uint32_t r = r0 + tid * RPT + j;
      lens[j] = r < nruns ? min(ld_u32u(src + uint64_t(r) * STR), n) : 0u;
      sum = sat_add(sum, lens[j], n);
    }
    uint32_t total;
    uint32_t pre = block_excl_scan_sat(sum, n, cx.ws, &total);
    pre = sat_add(pre, done, n);
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j) {
      starts[tid * RPT + j] = pre;
      pre = sat_add(pre, lens[j], n);
    }
    uint32_t hi = sat_add(done, total, n);
    if (tid == 0) starts[CH] = hi;
    __syncthreads();
    GenRle<W> g;
    g.starts = starts;
    g.runs = src + uint64_t(r0) * STR;
    g.nr = min(CH, nruns - r0);
    emit<W>(dst, done, hi, g);
    done = hi;
    __syncthreads();
  }
  cx.ar = mark;
  if (done < n) { // read past the runs: "failed to fill whole buffer"
    cx.flag(SB_IO);
    return false;
  }
  return true;
}

// ------------------------------------------------------------------------------------
// BitPacker4x blocks (integer/bp.rs:67-86, delta_bp.rs:69-91; layout SURVEY App. D.1):
// per 128 values `[u8 b][16*b bytes]`; value i = lane i%4, position i/4 of that lane's
// LSB-first stream; word w of lane l at u32 index 4w+l, i.e. 16-byte rows of 4 lanes.
// One warp decodes one block: thread t extracts position t of all 4 lanes = the 16-byte
// output vector out[4t..4t+3].  Every warp walks the (data-dependent) block chain itself.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 bp_unpack_lane(const uint8_t *blk, uint32_t bits, uint32_t t) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (bits == 0) return v;
  uint32_t bit = t * bits, w = bit >> 5, s = bit & 31;
  uint4 r0 = ld_u128u(blk + 16 * w);
  uint4 r1 = make_uint4(0, 0, 0, 0);
  if (s + bits > 32) r1 = ld_u128u(blk + 16 * (w + 1));
  uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
  v.x = __funnelshift_r(r0.x, r1.x, s) & mask;
  v.y = __funnelshift_r(r0.y, r1.y, s) & mask;
  v.z = __funnelshift_r(r0.z, r1.z, s) & mask;
  v.w = __funnelshift_r(r0.w, r1.w, s) & mask;
  return v;
}
__device__ __forceinline__ void bp_store(uint32_t *dst, uint32_t base, uint32_t n, uint4 v) {
  if (base + 4 <= n && ((uintptr_t(dst + base) & 15) == 0)) {
    *reinterpret_cast<uint4 *>(dst + base) = v;
  } else {
    if (base < n) dst[base] = v.x;
    if (base + 1 < n) dst[base + 1] = v.y;
    if (base + 2 < n) dst[base + 2] = v.z;
    if (base + 3 < n) dst[base + 3] = v.w;
  }
}
template <bool DELTA> __device__ bool dec_bitpack(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, uint8_t *dst8) {
  uint32_t *dst = reinterpret_cast<uint32_t *>(dst8);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nblk = (n + 127) >> 7;
  Arena mark = cx.ar;
  uint32_t *blk_pos = nullptr, *blk_sum = nullptr;
  if (DELTA) {
    blk_pos = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(nblk) * 4));
    blk_sum = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(nblk) * 4));
    if (!blk_pos || !blk_sum) {
      cx.flag(SB_NYI);
      return false;
    }
  }
  uint32_t pos = 0;
  // Uniform width (the rule for Dict indices, and for sorted values inside one page): every block header sits
  // at b * (1 + 16 * bits0).  All headers are checked at once (one per thread) and the blocks are handled without
  // the serial walk over the headers.
  bool uniform = false;
  if (nblk > 1 && avail > 0) {
    const uint32_t bits0 = src[0], stride = 1 + 16 * bits0;
    bool same = bits0 <= 32 && uint64_t(nblk) * stride <= avail;
    if (same)
      for (uint32_t b = tid; b < nblk; b += SB_NT) same &= src[b * stride] == bits0;
    uniform = __syncthreads_and(same) != 0;
    if (uniform) {
      for (uint32_t b = warp; b < nblk; b += SB_NWARP) {
        const uint4 v = bp_unpack_lane(src + b * stride + 1, bits0, lane);
        if (!DELTA) {
          bp_store(dst, b * 128 + lane * 4, n, v);
        } else {
          const uint32_t s = warp_sum(v.x + v.y + v.z + v.w);
          if (lane == 0) {
            blk_pos[b] = b * stride;
            blk_sum[b] = s;
          }
        }
      }
      if (!DELTA) {
        cx.ar = mark;
        return true;
      }
    }
  }
  if (!uniform)
  for (uint32_t b = 0; b < nblk; ++b) {
    if (pos >= avail) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t bits = src[pos];
    if (bits > 32 || 16 * bits > avail - pos - 1) {
      cx.flag(SB_PANIC);
      return false;
    }
    if ((b & (SB_NWARP - 1)) == warp) {
      uint4 v = bp_unpack_lane(src + pos + 1, bits, lane);
      if (!DELTA) {
        bp_store(dst, b * 128 + lane * 4, n, v);
      } else {
        uint32_t s = warp_sum(v.x + v.y + v.z + v.w);
        if (lane == 0) {
          blk_pos[b] = pos;
          blk_sum[b] = s;
        }
      }
    }
    pos += 1 + 16 * bits;
  }
  if (DELTA) {
    __syncthreads();
    // exclusive scan of block sums -> `initial` of every block (delta_bp.rs:73,87)
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += SB_NT) {
      uint32_t b = b0 + tid;
      uint32_t v = b < nblk ? blk_sum[b] : 0u, total;
      uint32_t ex = block_excl_scan(v, cx.ws, &total);
      if (b < nblk) blk_sum[b] = carry + ex;
      carry += total;
    }
    __syncthreads();
    for (uint32_t b = warp; b < nblk; b += SB_NWARP) {
      uint32_t p = blk_pos[b];
      uint32_t bits = src[p];
      uint4 v = bp_unpack_lane(src + p + 1, bits, lane);
      v.y += v.x;
      v.z += v.y;
      v.w += v.z;
      uint32_t inc = warp_incl_scan(v.w);
      uint32_t base = blk_sum[b] + inc - v.w;
      v.x += base, v.y += base, v.z += base, v.w += base;
      bp_store(dst, b * 128 + lane * 4, n, v);
    }
  }
  cx.ar = mark;
  return true;
}

// ------------------------------------------------------------------------------------
// Patas (double/patas.rs:107-132): [first value] then per value
//   [u16 = ref << 9 | (sig & 7) << 6 | tz][sig bytes of (xor >> tz)],  v[i] = (x << tz) ^ v[i - ref].
// Two chains: where value i starts (each header tells the length of its payload) and what
// value i is (it refers to one of the 127 values before it).
//   1. positions: warp 0 walks the stream like the LZ4 scanner -- every lane decodes the two bytes
//      at its window position as a header, the real chain is followed with one SHFL per value;
//      all format checks of the reference loop (short stream, sig > W, ref == 0, ref > i) happen
//      here, in stream order, so the first failure is the reference's failure;
//   2. values: 128 values per step, one per thread.  References into earlier steps are final;
//      chains inside the step are collapsed by pointer jumping on (xor accumulator, reference)
//      pairs in shared memory, at most 7 rounds.
// ------------------------------------------------------------------------------------
template <int W> __device__ bool dec_patas(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, uint8_t *dst) {
  using T = typename Elem<W>::T;
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  if (n == 0) { // `length - 1` underflows (patas.rs:117)
    cx.flag(SB_PANIC);
    return false;
  }
  if (avail < uint32_t(W)) {
    cx.flag(SB_IO);
    return false;
  }
  Arena mark = cx.ar;
  uint32_t *pos = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
  uint64_t *hist = static_cast<uint64_t *>(cx.ar.alloc(256 * 8));            // final values, index & 255
  uint64_t *acc_s = static_cast<uint64_t *>(cx.ar.alloc(2 * SB_NT * 8));     // double-buffered accumulators
  uint32_t *ptr_s = static_cast<uint32_t *>(cx.ar.alloc(2 * SB_NT * 4));     // double-buffered references
  if (!pos || !hist || !acc_s || !ptr_s) {
    cx.flag(SB_NYI);
    return false;
  }
  __syncthreads();
  // ---- 1. positions (warp 0)
  if (tid < 32) {
    int rc = 0;
    uint32_t q = W, i = 1; // stream position of value i
    while (i < n && rc == 0) {
      if (avail - q >= 34) {
        // window: lane L reads the header that would start at q + L
        const uint32_t h = uint32_t(src[q + lane]) | (uint32_t(src[q + lane + 1]) << 8);
        uint32_t sig = (h >> 6) & 7;
        const uint32_t tz = h & 63, ref = h >> 9;
        if (tz < 63 && sig == 0) sig = 8; // unpack(), patas.rs:158-160
        const uint32_t pack = (lane + 2 + sig) | (sig > uint32_t(W) ? 0x80u : 0u) | (ref << 8);
        uint32_t p = 0, cnt = 0, myp = 0;
#pragma unroll
        for (uint32_t hop = 0; hop < 16; ++hop) { // a value takes >= 2 stream bytes
          const uint32_t v = __shfl_sync(0xffffffffu, pack, p);
          const uint32_t r = v >> 8, nx = v & 0x7fu;
          if ((v & 0x80u) || r == 0 || r > i || avail - q < nx) { // sig > W, bad reference, payload past the end
            rc = SB_PANIC;
            break;
          }
          if (lane == hop) myp = q + p;
          p = nx;
          ++cnt;
          ++i;
          if (p >= 32 || i >= n) break;
        }
        if (lane < cnt) pos[i - cnt + lane] = myp;
        q += p;
      } else {
        // tail of the stream: one value at a time, the reference's checks in its order
        if (avail - q < 2) {
          rc = SB_IO;
          break;
        }
        const uint32_t h = ld_u16u(src + q);
        uint32_t sig = (h >> 6) & 7;
        const uint32_t tz = h & 63, ref = h >> 9;
        if (tz < 63 && sig == 0) sig = 8;
        if (sig > uint32_t(W) || avail - q - 2 < sig || ref == 0 || ref > i) {
          rc = SB_PANIC;
          break;
        }
        if (lane == 0) pos[i] = q;
        q += 2 + sig;
        ++i;
      }
    }
    if (lane == 0) cx.bcast[0] = rc;
  }
  __syncthreads();
  int rc = cx.bcast[0];
  if (rc) {
    cx.ar = mark;
    cx.flag(rc);
    return false;
  }
  // ---- 2. values, 128 per step
  T *out = reinterpret_cast<T *>(dst);
  const uint64_t first = uint64_t(ld_elem_u<W>(src));
  for (uint32_t c0 = 0; c0 < n; c0 += SB_NT) {
    const uint32_t i = c0 + tid;
    uint64_t acc = 0;
    uint32_t ptr = 0xffffffffu; // in-step index of the value this one still needs; 0xffffffff = final
    if (i == 0) {
      acc = first;
    } else if (i < n) {
      const uint32_t at = pos[i];
      const uint32_t h = ld_u16u(src + at);
      uint32_t sig = (h >> 6) & 7;
      const uint32_t tz = h & 63, ref = h >> 9;
      if (tz < 63 && sig == 0) sig = 8;
      uint64_t val = 0;
      for (uint32_t b = 0; b < sig; ++b) val |= uint64_t(src[at + 2 + b]) << (8 * b);
      acc = val << tz;
      const uint32_t P = i - ref;
      if (P < c0) acc ^= hist[P & 255]; // an earlier step: final
      else ptr = P - c0;
    }
    // pointer jumping inside the step: references point strictly backwards, <= 7 rounds
    for (uint32_t round = 0;; ++round) {
      uint64_t *ab = acc_s + (round & 1) * SB_NT;
      uint32_t *pb = ptr_s + (round & 1) * SB_NT;
      ab[tid] = acc;
      pb[tid] = ptr;
      if (!__syncthreads_or(ptr != 0xffffffffu)) break;
      if (ptr != 0xffffffffu) {
        acc ^= ab[ptr];
        ptr = pb[ptr];
      }
    }
    if (i < n) {
      if (W == 4) acc &= 0xffffffffull;
      out[i] = T(acc);
      hist[i & 255] = acc;
    }
    __syncthreads();
  }
  cx.ar = mark;
  return true;
}

// ------------------------------------------------------------------------------------
// value-block dispatcher: decompress_integer / decompress_double
// (integer/mod.rs:72-117, double/mod.rs:69-114).  LEVEL bounds the codec nesting
// (Dict -> indices, Freq -> exceptions; at most 3 stacked headers).
// ------------------------------------------------------------------------------------
// LIGHT: the instantiation of the "light" decode kernel (sb_lib.cu): flat fixed-width pages whose codec tree only
// holds None / OneValue / RLE / Bitpacking / DeltaBitpacking / Dict over those.  Everything serial or table heavy
// (LZ4 / Snappy / Zstd blocks, Freq, Patas) is compiled out, which is what lets that kernel run at 8 CTAs per SM.
template <int LEVEL, bool LIGHT = false>
__device__ bool decode_fixed(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, int W, bool is_float,
                             uint8_t *dst, uint32_t *consumed);

template <int W> struct GenDict {
  const uint32_t *idx;
  const uint8_t *tab; // k entries of W bytes (aligned copy or unaligned in-page)
  uint32_t k, i;
  int *err;
  bool aligned;
  __device__ __forceinline__ void seek(uint32_t x) { i = x; }
  __device__ __forceinline__ typename Elem<W>::T next() {
    uint32_t id = idx[i++];
    if (id >= k) { // data[*i as usize] out of bounds panics (integer/dict.rs:100)
      atomicCAS(err, 0, int(SB_PANIC));
      id = 0;
    }
    if (aligned) return reinterpret_cast<const typename Elem<W>::T *>(tab)[id];
    return ld_elem_u<W>(tab + uint64_t(id) * W);
  }
};

// Dict (integer/dict.rs:75-103): [VALUE_BLOCK<u32> indices][u32 k][k * W bytes]
template <int LEVEL, int W, bool LIGHT>
__device__ bool dec_dict(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, uint8_t *dst) {
  using T = typename Elem<W>::T;
  if constexpr (LEVEL >= 2) {
    cx.flag(SB_OUT_OF_SPEC);
    return false;
  } else {
    Arena mark = cx.ar;
    // ---- fused path: indices bit-packed at one width (what the chooser picks for 8192-row pages): every warp
    //      unpacks a 128-index block and gathers straight from the table -- no index buffer, no second pass
    if (avail >= 10 && src[0] == SB_C_BITPACK && (n & 127u) == 0 && n) {
      const uint32_t compressed = ld_u32u(src + 1), nblk = n >> 7, bits0 = src[9], stride = 1 + 16 * bits0;
      bool same = compressed <= avail - 9 && bits0 <= 32 && uint64_t(nblk) * stride <= compressed;
      if (same)
        for (uint32_t b = threadIdx.x; b < nblk; b += SB_NT) same &= src[9 + b * stride] == bits0;
      if (__syncthreads_and(same)) {
        const uint32_t used = 9 + compressed;
        if (avail - used < 4) {
          cx.flag(SB_IO);
          return false;
        }
        const uint32_t k = ld_u32u(src + used);
        const uint8_t *table = src + used + 4;
        if (uint64_t(k) * W > uint64_t(avail - used - 4)) { // dict.rs:80-86
          cx.flag(SB_OUT_OF_SPEC);
          return false;
        }
        const uint8_t *tab = table;
        bool aligned = (uintptr_t(table) & ((W > 16 ? 16 : W) - 1)) == 0;
        if (!aligned) {
          uint8_t *t2 = static_cast<uint8_t *>(cx.ar.alloc_shared(uint64_t(k) * W));
          if (t2) {
            copy_bytes(t2, table, uint64_t(k) * W);
            tab = t2;
            aligned = true;
          }
        }
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        T *out = reinterpret_cast<T *>(dst);
        for (uint32_t b = warp; b < nblk; b += SB_NWARP) {
          const uint4 v = bp_unpack_lane(src + 9 + b * stride + 1, bits0, lane);
          const uint32_t ids[4] = {v.x, v.y, v.z, v.w};
          T vals[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t id = ids[j];
            if (id >= k) { // data[*i as usize] out of bounds panics (integer/dict.rs:100)
              cx.flag(SB_PANIC);
              id = 0;
            }
            vals[j] = (k == 0) ? T{} : (aligned ? reinterpret_cast<const T *>(tab)[id] : ld_elem_u<W>(tab + uint64_t(id) * W));
          }
          T *o = out + b * 128 + lane * 4;
          if constexpr (W == 4) {
            if ((uintptr_t(o) & 15) == 0) {
              *reinterpret_cast<uint4 *>(o) = make_uint4(uint32_t(vals[0]), uint32_t(vals[1]), uint32_t(vals[2]), uint32_t(vals[3]));
              continue;
            }
          } else if constexpr (W == 8) {
            if ((uintptr_t(o) & 15) == 0) {
              reinterpret_cast<uint4 *>(o)[0] = make_uint4(uint32_t(vals[0]), uint32_t(vals[0] >> 32), uint32_t(vals[1]), uint32_t(vals[1] >> 32));
              reinterpret_cast<uint4 *>(o)[1] = make_uint4(uint32_t(vals[2]), uint32_t(vals[2] >> 32), uint32_t(vals[3]), uint32_t(vals[3] >> 32));
              continue;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = vals[j];
        }
        cx.ar = mark;
        return true;
      }
    }
    uint32_t *idx = static_cast<uint32_t *>(cx.ar.alloc(uint64_t(n) * 4 + 16));
    if (!idx) {
      cx.flag(SB_NYI);
      return false;
    }
    uint32_t used = 0;
    if (!decode_fixed<LEVEL + 1, LIGHT>(cx, src, avail, n, 4, false, reinterpret_cast<uint8_t *>(idx), &used)) return false;
    if (avail - used < 4) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t k = ld_u32u(src + used);
    const uint8_t *table = src + used + 4;
    if (uint64_t(k) * W > uint64_t(avail - used - 4)) { // dict.rs:80-86
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    GenDict<W> g;
    g.idx = idx;
    g.k = k;
    g.err = cx.err;
    g.tab = table;
    g.aligned = (uintptr_t(table) & ((W > 16 ? 16 : W) - 1)) == 0;
    if (!g.aligned) {
      uint8_t *tab = static_cast<uint8_t *>(cx.ar.alloc_shared(uint64_t(k) * W));
      if (tab) {
        copy_bytes(tab, table, uint64_t(k) * W);
        g.tab = tab;
        g.aligned = true;
      }
    }
    __syncthreads(); // indices + table visible
    emit<W>(dst, 0, n, g);
    cx.ar = mark;
    return true;
  }
}

// Freq (integer/freq.rs:88-123): [W top][u32 bm][roaring][VALUE_BLOCK<T> exceptions]
// Roaring portable format: array containers scatter directly; bitmap
// containers rank their bits with a block scan of word popcounts.
template <int LEVEL, int W>
__device__ bool dec_freq(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, bool is_float, uint8_t *dst) {
  using T = typename Elem<W>::T;
  if constexpr (LEVEL >= 2) {
    cx.flag(SB_OUT_OF_SPEC);
    return false;
  } else {
    const uint32_t tid = threadIdx.x;
    if (avail < uint32_t(W) + 4) {
      cx.flag(SB_IO);
      return false;
    }
    T top = ld_elem_u<W>(src);
    uint32_t bm = ld_u32u(src + W);
    const uint8_t *rb = src + W + 4;
    uint32_t rest = avail - W - 4;
    if (bm > rest) {
      cx.flag(SB_PANIC);
      return false;
    }
    // --- roaring header
    if (bm < 8) {
      cx.flag(SB_IO);
      return false;
    }
    uint32_t cookie = ld_u32u(rb);
    if (cookie != 12346u) { // run containers (cookie 12347) are never written by roaring 0.10 serialize_into
      cx.flag((cookie & 0xffff) == 12347u ? SB_NYI : SB_IO);
      return false;
    }
    uint32_t ncont = ld_u32u(rb + 4);
    if (ncont > 65536 || uint64_t(ncont) * 8 + 8 > bm) {
      cx.flag(SB_IO);
      return false;
    }
    // fill with the top value first (freq.rs:99-100)
    GenConst<W> gc{top};
    emit<W>(dst, 0, n, gc);
    // total cardinality and container data offsets (uniform serial walk; ncont = rows/65536)
    uint32_t n_exc = 0;
    {
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
      n_exc = uint32_t(tot);
    }
    Arena mark = cx.ar;
    T *exc = static_cast<T *>(cx.ar.alloc(uint64_t(n_exc) * W + 16));
    if (!exc) {
      cx.flag(SB_NYI);
      return false;
    }
    uint32_t used = 0;
    if (!decode_fixed<LEVEL + 1>(cx, rb + bm, rest - bm, n_exc, W, is_float, reinterpret_cast<uint8_t *>(exc), &used))
      return false;
    __syncthreads(); // fill + exceptions complete before the scatter
    T *out = reinterpret_cast<T *>(dst);
    uint32_t rank_base = 0;
    uint64_t off = 8 + uint64_t(ncont) * 8;
    for (uint32_t c = 0; c < ncont; ++c) {
      uint32_t key = ld_u16u(rb + 8 + 4 * c), card = ld_u16u(rb + 8 + 4 * c + 2) + 1;
      const uint8_t *data = rb + off;
      if (card <= 4096) {
        for (uint32_t j = tid; j < card; j += SB_NT) {
          uint32_t row = (key << 16) | ld_u16u(data + 2 * j);
          if (row < n) out[row] = exc[rank_base + j];
          else cx.flag(SB_PANIC); // output[begin + val] out of bounds
        }
        off += card * 2;
      } else {
        // 1024 u64 words; 8 consecutive words per thread
        uint64_t wds[8];
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          wds[j] = ld_u64u(data + 8 * (tid * 8 + j));
          cnt += __popcll(wds[j]);
        }
        uint32_t total;
        uint32_t pre = block_excl_scan(cnt, cx.ws, &total);
        uint32_t r = rank_base + pre;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint64_t bits = wds[j];
          while (bits) {
            uint32_t b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            uint32_t row = (key << 16) | ((tid * 8 + j) * 64 + b);
            if (row < n && r < n_exc) out[row] = exc[r];
            else cx.flag(SB_PANIC);
            ++r;
          }
        }
        off += 8192;
      }
      rank_base += card;
    }
    cx.ar = mark;
    return true;
  }
}

template <int LEVEL, int W, bool LIGHT>
__device__ __forceinline__ bool decode_fixed_w(Dctx &cx, int codec, const uint8_t *body, uint32_t compressed,
                                               uint32_t body_avail, uint32_t n, bool is_float, uint8_t *dst) {
  if constexpr (LIGHT) { // sb_classify_kernel only sends codec trees without these here; anything else is a logic error
    if (codec == SB_C_LZ4 || codec == SB_C_ZSTD || codec == SB_C_SNAPPY || codec == SB_C_FREQ || codec == SB_C_PATAS) {
      cx.flag(SB_PANIC);
      return false;
    }
  }
  switch (codec) {
  case SB_C_NONE:
    if constexpr (LIGHT) {
      if (uint64_t(compressed) != uint64_t(n) * W) { // copy_from_slice length mismatch panics (basic.rs:68)
        cx.flag(SB_PANIC);
        return false;
      }
      stream_copy(cx, dst, body, uint64_t(n) * W);
      return true;
    }
  case SB_C_LZ4:
  case SB_C_ZSTD:
  case SB_C_SNAPPY:
    if constexpr (!LIGHT) return dec_basic(cx, codec, body, compressed, dst, uint64_t(n) * W);
    return false;
  case SB_C_RLE: return dec_rle<W>(cx, body, min(compressed, body_avail), n, dst);
  case SB_C_ONEVALUE: return dec_onevalue<W>(cx, body, body_avail, 0, n, dst);
  case SB_C_DICT: return dec_dict<LEVEL, W, LIGHT>(cx, body, body_avail, n, dst);
  case SB_C_FREQ:
    if constexpr (!LIGHT) return dec_freq<LEVEL, W>(cx, body, body_avail, n, is_float, dst);
    return false;
  case SB_C_BITPACK:
  case SB_C_DELTABP:
    if (is_float) { // double/mod.rs:143-158
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    if constexpr (W == 4) {
      return codec == SB_C_BITPACK ? dec_bitpack<false>(cx, body, body_avail, n, dst)
                                   : dec_bitpack<true>(cx, body, body_avail, n, dst);
    } else {
      cx.flag(SB_PANIC); // bp.rs:70-79 writes u32 lanes whatever T is (App. C3)
      return false;
    }
  case SB_C_PATAS:
    if (!is_float) { // integer/mod.rs:146-161
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    if constexpr (!LIGHT && (W == 4 || W == 8)) return dec_patas<W>(cx, body, body_avail, n, dst);
    cx.flag(SB_OUT_OF_SPEC);
    return false;
  default: cx.flag(SB_OUT_OF_SPEC); return false; // compression/mod.rs:78-80
  }
}

template <int LEVEL, bool LIGHT>
__device__ bool decode_fixed(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, int W, bool is_float,
                             uint8_t *dst, uint32_t *consumed) {
  if (avail < 9) { // read_compress_header (read_basic.rs:181-189)
    cx.flag(SB_IO);
    return false;
  }
  int codec = src[0];
  uint32_t compressed = ld_u32u(src + 1);
  const uint8_t *body = src + 9;
  uint32_t body_avail = avail - 9;
  if (compressed > body_avail) {
    cx.flag(SB_IO);
    return false;
  }
  *consumed = 9 + compressed;
  switch (W) {
  case 1: return decode_fixed_w<LEVEL, 1, LIGHT>(cx, codec, body, compressed, body_avail, n, is_float, dst);
  case 2: return decode_fixed_w<LEVEL, 2, LIGHT>(cx, codec, body, compressed, body_avail, n, is_float, dst);
  case 4: return decode_fixed_w<LEVEL, 4, LIGHT>(cx, codec, body, compressed, body_avail, n, is_float, dst);
  case 16: return decode_fixed_w<LEVEL, 16, LIGHT>(cx, codec, body, compressed, body_avail, n, is_float, dst);
  case 32: return decode_fixed_w<LEVEL, 32, LIGHT>(cx, codec, body, compressed, body_avail, n, is_float, dst);
  default: return decode_fixed_w<LEVEL, 8, LIGHT>(cx, codec, body, compressed, body_avail, n, is_float, dst);
  }
}

// ------------------------------------------------------------------------------------
// bit streams: validity (read_validity, read_basic.rs:36-63) and boolean values
// (boolean/mod.rs:87-92).  Destination is an LSB-first Arrow bitmap at an arbitrary bit
// offset (pages are concatenated); it is zero-initialised by the host, whole words are
// stored, boundary words are OR-merged atomically with the neighbouring pages.
// ------------------------------------------------------------------------------------
template <class BitSrc>
__device__ __forceinline__ void emit_bits(uint8_t *dst_bitmap, uint64_t dst_bit, uint32_t nbits, BitSrc &bs) {
  if (nbits == 0) return;
  uint32_t *words = reinterpret_cast<uint32_t *>(dst_bitmap);
  uint64_t w_first = dst_bit >> 5, w_last = (dst_bit + nbits - 1) >> 5;
  for (uint64_t w = w_first + threadIdx.x; w <= w_last; w += SB_NT) {
    int64_t s = int64_t(w << 5) - int64_t(dst_bit); // source bit index of this word's bit 0
    uint32_t val = bs.word(s, nbits);
    bool partial = (s < 0) || (uint64_t(s) + 32 > nbits);
    if (partial) {
      if (val) atomicOr(words + w, val);
    } else {
      words[w] = val;
    }
  }
}
// 32 source bits starting at (possibly negative) bit s from a packed byte stream of nbits
struct BitsPacked {
  const uint8_t *p;
  __device__ __forceinline__ uint32_t word(int64_t s, uint32_t nbits) const {
    uint32_t lo_skip = s < 0 ? uint32_t(-s) : 0u; // bits below the stream start
    uint64_t s0 = s < 0 ? 0 : uint64_t(s);
    uint32_t want = 32 - lo_skip;                 // bits wanted from s0
    uint64_t avail = nbits - s0;
    uint32_t take = avail < want ? uint32_t(avail) : want;
    uint64_t byte = s0 >> 3;
    uint32_t sh = uint32_t(s0 & 7);
    uint32_t nbytes = (sh + take + 7) >> 3; // <= 5
    uint64_t acc = 0;
    for (uint32_t b = 0; b < nbytes; ++b) acc |= uint64_t(p[byte + b]) << (8 * b);
    uint32_t v = uint32_t(acc >> sh);
    if (take < 32) v &= (1u << take) - 1u;
    return v << lo_skip;
  }
};
struct BitsConst {
  bool one;
  __device__ __forceinline__ uint32_t word(int64_t s, uint32_t nbits) const {
    if (!one) return 0;
    uint32_t lo_skip = s < 0 ? uint32_t(-s) : 0u;
    uint64_t s0 = s < 0 ? 0 : uint64_t(s);
    uint32_t want = 32 - lo_skip;
    uint64_t avail = nbits - s0;
    uint32_t take = avail < want ? uint32_t(avail) : want;
    uint32_t v = take >= 32 ? 0xffffffffu : ((1u << take) - 1u);
    return v << lo_skip;
  }
};

// boolean RLE (boolean/rle.rs:41-55): (u32 run, u8 0/1)*; expanded per chunk of runs into
// a shared bit staging buffer, then emitted at the destination bit offset.
__device__ bool dec_bool_rle(Dctx &cx, const uint8_t *src, uint32_t plen, uint32_t n, uint8_t *dst_bitmap,
                             uint64_t dst_bit) {
  constexpr uint32_t STR = 5, RPT = 8, CH = SB_NT * RPT;
  const uint32_t tid = threadIdx.x;
  const uint32_t nruns = plen / STR;
  Arena mark = cx.ar;
  uint32_t *starts = static_cast<uint32_t *>(cx.ar.alloc((CH + 1) * 4));
  if (!starts) {
    cx.flag(SB_NYI);
    return false;
  }
  uint32_t *words = reinterpret_cast<uint32_t *>(dst_bitmap);
  uint32_t done = 0;
  for (uint32_t r0 = 0; r0 < nruns && done < n; r0 += CH) {
    uint32_t lens[RPT], sum = 0;
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j) {
      uint32_t r = r0 + tid * RPT + j;
      lens[j] = r < nruns ? min(ld_u32u(src + uint64_t(r) * STR), n) : 0u;
      sum = sat_add(sum, lens[j], n);
    }
    uint32_t total;
    uint32_t pre = sat_add(block_excl_scan_sat(sum, n, cx.ws, &total), done, n);
#pragma unroll
    for (uint32_t j = 0; j < RPT; ++j) {
      starts[tid * RPT + j] = pre;
      pre = sat_add(pre, lens[j], n);
    }
    uint32_t hi = sat_add(done, total, n);
    if (tid == 0) starts[CH] = hi;
    __syncthreads();
    uint32_t nr = min(CH, nruns - r0);
    if (hi > done) {
      // output words covering bits [done, hi)
      uint64_t b0 = dst_bit + done, b1 = dst_bit + hi; // absolute bit range
      for (uint64_t w = (b0 >> 5) + tid; w <= ((b1 - 1) >> 5); w += SB_NT) {
        uint64_t wb = w << 5;
        uint64_t lo_abs = wb < b0 ? b0 : wb, hi_abs = (wb + 32) < b1 ? (wb + 32) : b1;
        uint32_t e = uint32_t(lo_abs - dst_bit), e_end = uint32_t(hi_abs - dst_bit);
        uint32_t lo = 0, hh = nr; // largest r with starts[r] <= e
        while (hh - lo > 1) {
          uint32_t mid = (lo + hh) >> 1;
          if (starts[mid] <= e) lo = mid;
          else hh = mid;
        }
        uint32_t r = lo, val = 0;
        while (e < e_end) {
          while (r + 1 < nr && e >= starts[r + 1]) ++r;
          uint32_t run_end = min(starts[r + 1], e_end);
          if (run_end <= e) break;
          uint32_t len = run_end - e;
          if (src[uint64_t(r0 + r) * STR + 4] != 0) {
            uint32_t m = len >= 32 ? 0xffffffffu : ((1u << len) - 1u);
            val |= m << uint32_t(dst_bit + e - wb);
          }
          e = run_end;
        }
        bool partial = (wb < dst_bit) || (wb + 32 > dst_bit + n) || (lo_abs != wb) || (hi_abs != wb + 32);
        if (partial) {
          if (val) atomicOr(words + w, val);
        } else words[w] = val;
      }
    }
    done = hi;
    __syncthreads();
  }
  cx.ar = mark;
  return true; // the reference loop also ends quietly when the runs are exhausted (rle.rs:43)
}

// decompress_boolean (boolean/mod.rs:63-102)
__device__ bool decode_boolean(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, uint8_t *dst_bitmap,
                               uint64_t dst_bit) {
  if (avail < 9) {
    cx.flag(SB_IO);
    return false;
  }
  int codec = src[0];
  uint32_t compressed = ld_u32u(src + 1);
  const uint8_t *body = src + 9;
  uint32_t body_avail = avail - 9;
  if (compressed > body_avail) {
    cx.flag(SB_IO);
    return false;
  }
  uint32_t nbytes = (n + 7) >> 3;
  switch (codec) {
  case SB_C_NONE: {
    if (compressed != nbytes) {
      cx.flag(SB_PANIC);
      return false;
    }
    BitsPacked bs{body};
    emit_bits(dst_bitmap, dst_bit, n, bs);
    return true;
  }
  case SB_C_LZ4:
  case SB_C_ZSTD:
  case SB_C_SNAPPY: {
    Arena mark = cx.ar;
    uint8_t *tmp = static_cast<uint8_t *>(cx.ar.alloc(uint64_t(nbytes) + 16));
    if (!tmp) {
      cx.flag(SB_NYI);
      return false;
    }
    if (!dec_basic(cx, codec, body, compressed, tmp, nbytes)) return false;
    BitsPacked bs{tmp};
    emit_bits(dst_bitmap, dst_bit, n, bs);
    __syncthreads();
    cx.ar = mark;
    return true;
  }
  case SB_C_RLE: return dec_bool_rle(cx, body, min(compressed, body_avail), n, dst_bitmap, dst_bit);
  case SB_C_ONEVALUE: {
    if (body_avail == 0) { // boolean/one_value.rs:55-57
      cx.flag(SB_OUT_OF_SPEC);
      return false;
    }
    BitsConst bs{body[0] > 0};
    emit_bits(dst_bitmap, dst_bit, n, bs);
    return true;
  }
  default: cx.flag(SB_OUT_OF_SPEC); return false;
  }
}

// read_validity (read_basic.rs:36-63): [u32 L][ULEB((ceil8(n)<<1)|1)][bitmap]; L == 0
// pushes nothing.  Returns the byte offset of the value block, or 0xffffffff on failure.
__device__ uint32_t decode_validity(Dctx &cx, const uint8_t *src, uint32_t avail, uint32_t n, uint8_t *dst_bitmap,
                                    uint64_t dst_bit) {
  if (avail < 4) {
    cx.flag(SB_IO);
    return 0xffffffffu;
  }
  uint32_t L = ld_u32u(src);
  if (L > avail - 4) {
    cx.flag(SB_IO);
    return 0xffffffffu;
  }
  if (L == 0) {
    // the reference pushes nothing (read_basic.rs:43-45) and the array constructor then rejects the validity
    // length (PrimitiveArray::try_new -> OutOfSpec): a page with rows but no bitmap is an error, not all-null
    if (n) {
      cx.flag(SB_OUT_OF_SPEC);
      return 0xffffffffu;
    }
    return 4;
  }
  // one bit-packed hybrid-RLE run (parquet2 encode_bool); RLE runs are `unreachable!()` upstream
  const uint8_t *p = src + 4;
  uint64_t header = 0;
  uint32_t used = 0;
  for (uint32_t i = 0; i < L && i < 10; ++i) {
    uint32_t b = p[i];
    header |= uint64_t(b & 0x7f) << (7 * i);
    if (!(b & 0x80)) {
      used = i + 1;
      break;
    }
  }
  if (used == 0 || !(header & 1)) {
    cx.flag(SB_PANIC);
    return 0xffffffffu;
  }
  uint64_t bytes = header >> 1;
  if (bytes > L - used) bytes = L - used;
  if (bytes * 8 < n) {
    cx.flag(SB_PANIC);
    return 0xffffffffu;
  }
  BitsPacked bs{p + used};
  emit_bits(dst_bitmap, dst_bit, n, bs);
  return 4 + L;
}

} // namespace sb
