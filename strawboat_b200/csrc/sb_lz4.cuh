// sb_lz4.cuh -- LZ4 block decode by one warp (basic.rs:87-91 -> LZ4_decompress_safe with a
// known decoded size; block format: SURVEY App. D.5).
//
// The token stream is inherently serial, so the cost that matters is the latency of one
// sequence.  The fast path reads one 16-byte window at `ip` (token, <= 12 literals and
// the 2-byte offset all sit inside it), so the ip -> next ip chain is one load plus a few
// ALU ops; literal and match bytes are moved lane-parallel, one byte per lane.
//
// Output models:
//   FlatOut : dst is directly addressable (shared-memory arena or global) -- nested blocks.
//   RingOut : dedicated kernel; the warp keeps the last SB_LZ4_RING decoded bytes in a
//             shared-memory ring (match sources are read from it at shared-memory latency)
//             and writes them behind to HBM in 16-byte vectors; matches further back than
//             the ring are read from the already flushed global output.
#pragma once
#include "sb_common.cuh"

namespace sb {

constexpr uint32_t SB_LZ4_RING = 8192;   // per-warp output ring bytes
constexpr uint32_t SB_LZ4_IN = 2048;     // per-warp input ring bytes (two halves, cp.async prefetched)
constexpr uint32_t SB_LZ4_CHUNK = SB_LZ4_IN / 2;
constexpr uint32_t SB_LZ4_FLUSH = 2048;  // flush granularity
constexpr uint32_t SB_LZ4_PIECE = 1024;  // long literal / match runs are moved in pieces

struct FlatOut {
  uint8_t *dst;
  __device__ __forceinline__ void init() {}
  __device__ __forceinline__ void st(uint32_t pos, uint32_t b) { dst[pos] = uint8_t(b); }
  __device__ __forceinline__ uint32_t ld(uint32_t pos, uint32_t /*op*/) const { return dst[pos]; }
  __device__ __forceinline__ void advance(uint32_t /*op*/) {}
  __device__ __forceinline__ void finish(uint32_t /*op*/) {}
};

__device__ __forceinline__ uint32_t u4_byte(const uint4 &w, uint32_t idx) {
  uint32_t lo = (idx & 4) ? w.y : w.x, hi = (idx & 4) ? w.w : w.z;
  uint32_t word = (idx & 8) ? hi : lo;
  return (word >> ((idx & 3) * 8)) & 0xffu;
}

// Returns 0 or SB_EXTERNAL (uniform across the warp).  `src` may be shared or global.
template <class Out> __device__ int lz4_decode_warp2(const uint8_t *src, uint32_t clen, Out &out, uint32_t dlen) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t ip = 0, op = 0;
  out.init();
  if (clen == 0) return dlen == 0 ? 0 : SB_EXTERNAL;
  for (;;) {
    if (ip >= clen) return SB_EXTERNAL;
    uint32_t token, lit, mlc;
    bool fast = false;
    uint4 w;
    if (clen - ip >= 16) {
      w = ld_u128u(src + ip);
      token = w.x & 0xffu;
      lit = token >> 4;
      mlc = token & 15u;
      fast = lit <= 12 && mlc != 15;
    } else {
      token = src[ip];
      lit = token >> 4;
      mlc = token & 15u;
    }
    if (fast) {
      // ---- whole sequence inside the window: [token][lit bytes][offset lo][offset hi]
      uint32_t offset = u4_byte(w, 1 + lit) | (u4_byte(w, 2 + lit) << 8);
      uint32_t ml = mlc + 4;
      uint32_t mpos = op + lit; // first match byte
      if (lit + ml > dlen - op || offset == 0 || offset > mpos) return SB_EXTERNAL;
      if (lane < lit) out.st(op + lane, u4_byte(w, 1 + lane));
      __syncwarp();
      if (lane < ml) {
        uint32_t j = lane;
        if (offset < ml) j = lane % offset;
        uint32_t b = out.ld(mpos - offset + j, mpos);
        out.st(mpos + lane, b);
      }
      __syncwarp();
      ip += 3 + lit;
      op = mpos + ml;
      out.advance(op);
      continue;
    }
    // ---- general path: length extensions, long runs, stream tail
    ++ip;
    if (lit == 15) {
      uint32_t b;
      do {
        if (ip >= clen) return SB_EXTERNAL;
        b = src[ip++];
        lit += b;
      } while (b == 255);
    }
    if (lit > clen - ip || lit > dlen - op) return SB_EXTERNAL;
    for (uint32_t done = 0; done < lit;) {
      uint32_t p = min(lit - done, SB_LZ4_PIECE);
      for (uint32_t i = lane; i < p; i += 32) out.st(op + i, src[ip + i]);
      __syncwarp();
      ip += p;
      op += p;
      done += p;
      out.advance(op);
    }
    if (ip == clen) break; // last sequence carries literals only
    if (clen - ip < 2) return SB_EXTERNAL;
    uint32_t offset = uint32_t(src[ip]) | (uint32_t(src[ip + 1]) << 8);
    ip += 2;
    if (offset == 0 || offset > op) return SB_EXTERNAL;
    uint32_t ml = mlc;
    if (ml == 15) {
      uint32_t b;
      do {
        if (ip >= clen) return SB_EXTERNAL;
        b = src[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (ml > dlen - op) return SB_EXTERNAL;
    for (uint32_t done = 0; done < ml;) {
      uint32_t p = min(ml - done, SB_LZ4_PIECE);
      // piece sources all precede op: dst[op+i] = dst[op-offset + i % offset]
      if (offset >= p) {
        for (uint32_t i = lane; i < p; i += 32) {
          uint32_t b = out.ld(op - offset + i, op);
          out.st(op + i, b);
        }
      } else {
        for (uint32_t i = lane; i < p; i += 32) {
          uint32_t b = out.ld(op - offset + (i % offset), op);
          out.st(op + i, b);
        }
      }
      __syncwarp();
      op += p;
      done += p;
      out.advance(op);
    }
  }
  out.finish(op);
  return op == dlen ? 0 : SB_EXTERNAL;
}

// ------------------------------------------------------------------------------------
// Streaming decoder of the dedicated LZ4 kernel: one CTA of 2 warps per block (page).
//
// Measured on B200 (profiles/README.md): a warp that walks tokens AND moves bytes retires one
// dependent instruction every ~7-10 cycles and needs ~100 of them per sequence, whatever the
// lane count -- short sequences (2 literals + 6 match bytes are typical for numeric columns)
// leave 30 of 32 lanes idle.  So the work is split by what is serial and what is not:
//
//   scanner (warp 0): owns the input ring (global -> shared, cp.async).  Finds sequence
//     starts.  Every lane treats "its" byte of a 32-byte window as a token and computes
//     where the next token would be; the real chain is then followed with ONE SHFL per
//     sequence (no memory access on the chain).  Emits one queue entry per sequence = stream
//     position of its token.  Tokens with length-extension bytes and the block tail take a
//     scalar path.
//   mover (warp 1): owns the output ring (shared, written behind to HBM in 16-byte vectors).
//     Takes up to 32 entries at a time, ONE SEQUENCE PER LANE: every lane parses its token,
//     a warp scan of the sequence lengths gives every lane its output position, all literals
//     are copied at once, then the matches run lane-parallel in "independent prefix" rounds:
//     all leading matches whose source ends before the first pending match begins are copied
//     concurrently; a match that depends on a pending one starts the next round.  Long
//     literals / matches are moved by the whole warp.
//
// Shared state: entry queue (scanner -> mover), `produced` / `in_ready` (scanner-published),
// `consumed` / `m_q` (mover-published; m_q = first stream byte the mover still needs, so the
// scanner never refills ring slots under it), `abort`.
// A sequence whose literal or match length needs more than 4 extension bytes (>= 1035) is a
// BIG entry: it may not fit the rings, so both warps parse and move it in streaming fashion.
// ------------------------------------------------------------------------------------
constexpr uint32_t SB_LZ4_INR = 4096;       // input ring bytes
constexpr uint32_t SB_LZ4_INCH = 1024;      // refill granularity
constexpr uint32_t SB_LZ4_Q = 256;          // queue entries (u32 each)
#ifndef SB_LZ4_SLACK
#define SB_LZ4_SLACK 96                     // entries the mover retires before a blocked scanner resumes
#endif
constexpr uint32_t SB_LZ4_NEAR = SB_LZ4_RING - 2048 - 64; // match distance served from the ring
constexpr uint32_t SB_LZ4_FLUSHQ = 1024;    // write-behind granularity
constexpr uint32_t SB_LZ4_SMALL = 16;       // match lengths handled one sequence per lane
constexpr uint32_t SB_LZ4_SMALL_LIT = 32;   // literal runs handled one sequence per lane (two 16-byte passes)
constexpr uint32_t SB_LZ4_MAXPOS = 0x3fffffffu; // stream / output positions fit 30 bits
enum { LZ4_E_SEQ = 0u, LZ4_E_END = 1u, LZ4_E_BIG = 2u, LZ4_E_ERROR = 3u }; // entry >> 30

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_vol(uint32_t a) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_vol(uint32_t a, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

#ifdef SB_LZ4_PROF
__device__ unsigned long long g_lz4_prof[32];
#define LZ4_T(var) const uint32_t var = clock()
#define LZ4_ACC(slot, t0, t1) prof[slot] += (t1) - (t0)
#else
#define LZ4_T(var)
#define LZ4_ACC(slot, t0, t1)
#endif

struct __align__(16) Lz4Shared {
  uint8_t out[SB_LZ4_RING];
  uint8_t in[SB_LZ4_INR];
  uint32_t q[SB_LZ4_Q];
  uint32_t produced; // entries published by the scanner
  uint32_t in_ready; // stream bytes [.., in_ready) are in the input ring
  uint32_t consumed; // entries retired by the mover
  uint32_t m_q;      // the mover no longer needs stream bytes below m_q
  uint32_t abort;    // either side gave up (corrupt stream)
  uint32_t pad[3];
};

// ---- scanner -------------------------------------------------------------------------
struct Lz4Scan {
  uint32_t sh_b;      // shared address of Lz4Shared
  const uint8_t *gal; // 16-byte aligned global base of the stream
  uint32_t total;     // aligned stream bytes
  uint32_t issued, ready;
  uint32_t seq, pub;  // entries written / published
  uint32_t ready_pub; // in_ready as last published
  uint32_t own;       // lowest stream byte the scanner itself still reads
  uint32_t c_seen, mq_seen;

  __device__ __forceinline__ uint32_t a_in() const { return sh_b + offsetof(Lz4Shared, in); }
  __device__ __forceinline__ uint32_t ib(uint32_t q) const { return lds_u8(a_in() + (q & (SB_LZ4_INR - 1))); }
  __device__ __forceinline__ void publish() {
    fence_cta();
    if ((threadIdx.x & 31) == 0) {
      sts_vol(sh_b + offsetof(Lz4Shared, in_ready), ready);
      sts_vol(sh_b + offsetof(Lz4Shared, produced), seq);
    }
    pub = seq;
    ready_pub = ready;
  }
  __device__ __forceinline__ void refresh() {
    c_seen = lds_vol(sh_b + offsetof(Lz4Shared, consumed));
    mq_seen = lds_vol(sh_b + offsetof(Lz4Shared, m_q));
  }
  __device__ __forceinline__ bool aborted() const { return lds_vol(sh_b + offsetof(Lz4Shared, abort)) != 0; }
  // request every chunk whose ring slots the mover (and the scanner itself) has released
  __device__ __forceinline__ bool issue_allowed() {
    const uint32_t lane = threadIdx.x & 31;
    bool any = false;
    // the chunk at `issued` lands on the ring slots of stream bytes [issued - INR, issued - INR + INCH)
    while (issued < total && (issued < SB_LZ4_INR || min(mq_seen, own) + (SB_LZ4_INR - SB_LZ4_INCH) >= issued)) {
      uint32_t end = min(total, issued + SB_LZ4_INCH);
      for (uint32_t o = issued + lane * 16; o < end; o += 512) cp_async16(a_in() + (o & (SB_LZ4_INR - 1)), gal + o);
      issued = end;
      any = true;
    }
    if (any) cp_async_commit();
    return any;
  }
  // make stream bytes [.., upto) readable (upto <= total).  false = aborted by the mover.
  __device__ __forceinline__ bool need(uint32_t upto) {
    upto = min(upto, total);
    while (ready < upto) {
      if (issued > ready) {
        cp_async_wait_all();
        __syncwarp();
        ready = issued;
        publish(); // the mover may be streaming a long run: tell it right away
        if (ready >= upto) break;
      }
      refresh();
      if (issue_allowed()) continue;
      publish(); // blocked on the mover: everything found so far must be visible to it
      if (aborted()) return false;
      __nanosleep(64);
    }
    return true;
  }
  // room for `n` more entries.  Once the queue is full the scanner stays away until the mover has
  // retired `slack` more entries: it then scans that many sequences in one go instead of paying the
  // housekeeping (publish / refresh / poll) once per 32-byte window.
  __device__ __forceinline__ bool room(uint32_t n, uint32_t slack = 0) {
    if (seq + n - c_seen <= SB_LZ4_Q) return true;
    publish();
    while (seq + n + slack - c_seen > SB_LZ4_Q) {
      refresh();
      if (seq + n + slack - c_seen <= SB_LZ4_Q) break;
      if (aborted()) return false;
      __nanosleep(1000); // the queue holds 8 mover batches (~20 us of work): no need to poll fast
    }
    return true;
  }
  __device__ __forceinline__ void push(uint32_t q, uint32_t kind) { // room must exist
    if ((threadIdx.x & 31) == 0) sts_vol(sh_b + offsetof(Lz4Shared, q) + ((seq & (SB_LZ4_Q - 1)) << 2), q | (kind << 30));
    ++seq;
  }
};

// Returns 0 or SB_EXTERNAL.  The scanner validates the shape of the stream (every token, length
// byte and offset inside the block); the mover validates offsets and output sizes.
__device__ __forceinline__ int lz4_scan(const uint8_t *src, uint32_t clen, Lz4Shared *sh) {
  const uint32_t lane = threadIdx.x & 31;
  constexpr uint32_t IM = SB_LZ4_INR - 1;
  Lz4Scan s;
  s.sh_b = smem_u32(sh);
  const uint32_t mis = uint32_t(uintptr_t(src) & 15);
  s.gal = src - mis;
  s.total = (mis + clen + 15) & ~15u;
  s.issued = s.ready = s.ready_pub = 0;
  s.seq = s.pub = 0;
  s.c_seen = s.mq_seen = 0;
  s.own = mis;
  const uint32_t in_b = s.a_in(), q_b = s.sh_b + offsetof(Lz4Shared, q);
  const uint32_t end = mis + clen; // stream positions are offsets from gal
  uint32_t q = mis;
  uint32_t q_lim = 0; // windows run while q < q_lim
  int rc = 0;
  s.issue_allowed();
  for (;;) {
    if (q >= q_lim || s.seq + 12 - s.c_seen > SB_LZ4_Q) {
      // ---- housekeeping: completed prefetches, new prefetches, queue room
      if (s.issued > s.ready) {
        cp_async_wait_all();
        __syncwarp();
        s.ready = s.issued;
      }
      s.own = q;
      // publish BEFORE requesting new chunks: the release fence then has no copies in flight to wait for
      if (s.seq - s.pub >= 32 || s.ready != s.ready_pub) s.publish();
      s.refresh();
      s.issue_allowed();
      if (!s.room(12, SB_LZ4_SLACK) || !s.need(min(end, q + 64))) {
        rc = -1; // aborted by the mover (it reports the status)
        break;
      }
      // windows need 48 readable bytes and stay clear of the last 64 bytes of the block
      uint32_t lim_ready = s.ready >= 48 ? s.ready - 48 : 0, lim_end = end >= 64 ? end - 64 : 0;
      q_lim = min(min(lim_ready, lim_end), q + SB_LZ4_INCH);
    }
    if (q < q_lim) {
      // ---- window: lane i decodes byte q+i as a token; the chain hops with one SHFL per token
      const uint32_t b = lds_u8(in_b + ((q + lane) & IM));
      const uint32_t lit = b >> 4;
      const uint32_t pack = (lane + 3 + lit) | ((lit == 15 || (b & 15u) == 15) ? 0x100u : 0u);
      uint32_t p = 0, cnt = 0, myp = 0;
#pragma unroll
      for (uint32_t h = 0; h < 11; ++h) { // a sequence takes >= 3 stream bytes: <= 11 tokens in 32 bytes
        const uint32_t v = __shfl_sync(0xffffffffu, pack, p);
        if (v & 0x100u) break; // length bytes follow this token: scalar path
        if (lane == h) myp = p;
        p = v & 0xffu;
        cnt = h + 1;
        if (p >= 32) break;
      }
      if (lane < cnt) sts_vol(q_b + (((s.seq + lane) & (SB_LZ4_Q - 1)) << 2), q + myp);
      s.seq += cnt;
      q += p;
      if (cnt) continue;
    }
    // ---- scalar path: one token with all checks (length bytes, block tail)
    s.own = q;
    if (!s.room(1) || !s.need(min(end, q + 32))) {
      rc = -1;
      break;
    }
    const uint32_t q0 = q;
    const uint32_t tok = s.ib(q), mlc = tok & 15u;
    uint32_t lit = tok >> 4, r = q + 1;
    bool big = false;
    // length-extension bytes at r; after 4 of them the sequence becomes a BIG entry, which the
    // mover parses concurrently (it releases ring space as it goes, we keep feeding input)
    auto ext = [&](uint32_t &len) -> int {
      uint32_t x = 255, nx = 0;
      while (x == 255) {
        if (r >= end) return SB_EXTERNAL;
        if ((r & 15) == 0 || nx == 0) {
          if (big) s.own = r;
          if (!s.need(min(end, r + 16))) return -1;
        }
        x = s.ib(r++);
        len += x;
        if (len > SB_LZ4_MAXPOS) return SB_EXTERNAL;
        if (++nx == 4 && x == 255 && !big) {
          s.push(q0, LZ4_E_BIG);
          s.publish();
          big = true;
        }
      }
      return 0;
    };
    if (lit == 15 && (rc = ext(lit)) != 0) break;
    if (lit > end - r) {
      rc = SB_EXTERNAL;
      break;
    }
    if (r + lit == end) { // last sequence: literals only
      if (!big) {
        if (!s.need(end)) {
          rc = -1;
          break;
        }
        s.push(q0, LZ4_E_END);
      }
      s.publish();
      s.own = 0xffffffffu; // nothing left to scan
      if (!s.need(end)) rc = -1; // keep feeding the mover until everything is in the ring
      break;
    }
    r += lit;
    if (big) s.own = r; // the mover streams the literal run; we only need what follows it
    if (end - r < 2) {
      rc = SB_EXTERNAL;
      break;
    }
    if (!s.need(min(end, r + 32))) {
      rc = -1;
      break;
    }
    r += 2;
    uint32_t ml = mlc;
    if (mlc == 15 && (rc = ext(ml)) != 0) break;
    if (!big) s.push(q0, LZ4_E_SEQ);
    q = r;
    if (q >= end) { // a block must end with a literal-only sequence
      rc = SB_EXTERNAL;
      break;
    }
  }
  if (rc > 0) {
    if (s.room(1)) s.push(q, LZ4_E_ERROR);
    s.publish();
    if (lane == 0) sts_vol(s.sh_b + offsetof(Lz4Shared, abort), 1u);
  } else if (rc == 0) {
    s.publish();
  }
  return rc > 0 ? rc : 0;
}

// ---- mover ---------------------------------------------------------------------------
struct Lz4Out {
  uint32_t out_b; // shared address of the output ring
  uint8_t *ring;  // generic pointer to the same ring
  uint8_t *dst;   // global output
  uint32_t fl;    // bytes [0, fl) written to dst
  bool vec;
  // write ring bytes [fl, upto) behind to HBM.  Non-final flushes move whole 16-byte vectors
  // only (fl stays 16-byte aligned); the final flush also writes the byte tail.
  __device__ __forceinline__ void flush_to(uint32_t upto, bool final) {
    const uint32_t lane = threadIdx.x & 31;
    constexpr uint32_t OM = SB_LZ4_RING - 1;
    __syncwarp();
    if (vec) {
      uint32_t a = min(upto, (fl + 15) & ~15u);
      for (uint32_t pos = fl + lane; pos < a; pos += 32) dst[pos] = ring[pos & OM];
      fl = a;
      uint32_t vend = upto & ~15u;
      for (uint32_t pos = fl + lane * 16; pos + 16 <= vend; pos += 512)
        *reinterpret_cast<uint4 *>(dst + pos) = *reinterpret_cast<const uint4 *>(ring + (pos & OM));
      if (vend > fl) fl = vend;
    }
    if (final || !vec) {
      for (uint32_t pos = fl + lane; pos < upto; pos += 32) dst[pos] = ring[pos & OM];
      if (fl < upto) fl = upto;
    }
    __syncwarp();
  }
  // one earlier output byte for a match at `mpos` (ring when near or not flushed yet)
  __device__ __forceinline__ uint32_t src_byte(uint32_t sp, uint32_t mpos) const {
    if (mpos - sp <= SB_LZ4_NEAR || sp >= fl) return lds_u8(out_b + (sp & (SB_LZ4_RING - 1)));
    return __ldcg(dst + sp);
  }
  // whole-warp match copy of `ml` bytes at `mpos` (any length, any overlap)
  __device__ __forceinline__ void match_coop(uint32_t mpos, uint32_t offset, uint32_t ml) {
    const uint32_t lane = threadIdx.x & 31;
    constexpr uint32_t OM = SB_LZ4_RING - 1;
    for (uint32_t done = 0; done < ml;) {
      uint32_t p = min(ml - done, 1024u), at = mpos + done;
      if (offset < 32) { // pattern replication: lane byte = pattern[lane % offset]
        uint32_t v = lds_u8(out_b + ((at - offset + lane % offset) & OM));
        uint32_t step = 32 - 32 % offset;
        __syncwarp();
        for (uint32_t i = 0; i < p; i += step)
          if (lane < step && i + lane < p) sts_u8(out_b + ((at + i + lane) & OM), v);
        __syncwarp();
      } else {
        for (uint32_t i = 0; i < p; i += 32) { // a 32-byte step's sources precede it (offset >= 32)
          if (i + lane < p) sts_u8(out_b + ((at + i + lane) & OM), src_byte(at - offset + i + lane, at));
          __syncwarp();
        }
      }
      done += p;
      if (at + p - fl >= SB_LZ4_FLUSHQ) flush_to(at + p, false);
    }
  }
};

// Per-lane copy of bytes [K, K+4) of a short run, predicated on k < n.  Register + immediate
// addressing, no per-byte address arithmetic: 4 SETP + 4 loads + 4 stores.  The caller
// guarantees that neither range wraps around its ring and that the source bytes of one group
// are not written by the same group (distance >= 4 or disjoint buffers).
template <int K> __device__ __forceinline__ void lz4_copy4_ss(uint32_t s, uint32_t d, uint32_t n) {
  asm volatile(
      "{\n\t"
      ".reg .pred p<4>;\n\t"
      ".reg .u32 v<4>;\n\t"
      "setp.gt.u32 p0, %2, %3;\n\t"
      "setp.gt.u32 p1, %2, %4;\n\t"
      "setp.gt.u32 p2, %2, %5;\n\t"
      "setp.gt.u32 p3, %2, %6;\n\t"
      "@p0 ld.shared.u8 v0, [%0+%3];\n\t"
      "@p1 ld.shared.u8 v1, [%0+%4];\n\t"
      "@p2 ld.shared.u8 v2, [%0+%5];\n\t"
      "@p3 ld.shared.u8 v3, [%0+%6];\n\t"
      "@p0 st.shared.u8 [%1+%3], v0;\n\t"
      "@p1 st.shared.u8 [%1+%4], v1;\n\t"
      "@p2 st.shared.u8 [%1+%5], v2;\n\t"
      "@p3 st.shared.u8 [%1+%6], v3;\n\t"
      "}" ::"r"(s),
      "r"(d), "r"(n), "n"(K), "n"(K + 1), "n"(K + 2), "n"(K + 3)
      : "memory");
}
template <int K> __device__ __forceinline__ void lz4_copy4_gs(const uint8_t *s, uint32_t d, uint32_t n) {
  asm volatile(
      "{\n\t"
      ".reg .pred p<4>;\n\t"
      ".reg .u32 v<4>;\n\t"
      "setp.gt.u32 p0, %2, %3;\n\t"
      "setp.gt.u32 p1, %2, %4;\n\t"
      "setp.gt.u32 p2, %2, %5;\n\t"
      "setp.gt.u32 p3, %2, %6;\n\t"
      "@p0 ld.global.cg.u8 v0, [%0+%3];\n\t"
      "@p1 ld.global.cg.u8 v1, [%0+%4];\n\t"
      "@p2 ld.global.cg.u8 v2, [%0+%5];\n\t"
      "@p3 ld.global.cg.u8 v3, [%0+%6];\n\t"
      "@p0 st.shared.u8 [%1+%3], v0;\n\t"
      "@p1 st.shared.u8 [%1+%4], v1;\n\t"
      "@p2 st.shared.u8 [%1+%5], v2;\n\t"
      "@p3 st.shared.u8 [%1+%6], v3;\n\t"
      "}" ::"l"(s),
      "r"(d), "r"(n), "n"(K), "n"(K + 1), "n"(K + 2), "n"(K + 3)
      : "memory");
}
// Bytes [K, K+8) of a short run: 8 loads in flight before the first store (one shared-memory
// latency per 8 bytes); predicates are recomputed for the stores (only 7 predicate registers).
template <int K> __device__ __forceinline__ void lz4_copy8_ss(uint32_t s, uint32_t d, uint32_t n) {
  asm volatile(
      "{\n\t"
      ".reg .pred p<4>;\n\t"
      ".reg .u32 v<8>;\n\t"
      "setp.gt.u32 p0, %2, %3;\n\t"
      "setp.gt.u32 p1, %2, %4;\n\t"
      "setp.gt.u32 p2, %2, %5;\n\t"
      "setp.gt.u32 p3, %2, %6;\n\t"
      "@p0 ld.shared.u8 v0, [%0+%3];\n\t"
      "@p1 ld.shared.u8 v1, [%0+%4];\n\t"
      "@p2 ld.shared.u8 v2, [%0+%5];\n\t"
      "@p3 ld.shared.u8 v3, [%0+%6];\n\t"
      "setp.gt.u32 p0, %2, %7;\n\t"
      "setp.gt.u32 p1, %2, %8;\n\t"
      "setp.gt.u32 p2, %2, %9;\n\t"
      "setp.gt.u32 p3, %2, %10;\n\t"
      "@p0 ld.shared.u8 v4, [%0+%7];\n\t"
      "@p1 ld.shared.u8 v5, [%0+%8];\n\t"
      "@p2 ld.shared.u8 v6, [%0+%9];\n\t"
      "@p3 ld.shared.u8 v7, [%0+%10];\n\t"
      "@p0 st.shared.u8 [%1+%7], v4;\n\t"
      "@p1 st.shared.u8 [%1+%8], v5;\n\t"
      "@p2 st.shared.u8 [%1+%9], v6;\n\t"
      "@p3 st.shared.u8 [%1+%10], v7;\n\t"
      "setp.gt.u32 p0, %2, %3;\n\t"
      "setp.gt.u32 p1, %2, %4;\n\t"
      "setp.gt.u32 p2, %2, %5;\n\t"
      "setp.gt.u32 p3, %2, %6;\n\t"
      "@p0 st.shared.u8 [%1+%3], v0;\n\t"
      "@p1 st.shared.u8 [%1+%4], v1;\n\t"
      "@p2 st.shared.u8 [%1+%5], v2;\n\t"
      "@p3 st.shared.u8 [%1+%6], v3;\n\t"
      "}" ::"r"(s),
      "r"(d), "r"(n), "n"(K), "n"(K + 1), "n"(K + 2), "n"(K + 3), "n"(K + 4), "n"(K + 5), "n"(K + 6), "n"(K + 7)
      : "memory");
}
// n <= 16 bytes per lane (n = 0 for lanes that do not take part); whole-warp call: the upper
// groups are skipped when no lane needs them (`gt4` / `gt8`: warp-uniform "some lane has n > 4 / 8")
// `wide`: sources are at least 8 bytes away from their destinations (or in another buffer)
__device__ __forceinline__ void lz4_copy16_ss(uint32_t s, uint32_t d, uint32_t n, bool gt4, bool gt8, bool wide) {
  if (wide && gt4) {
    lz4_copy8_ss<0>(s, d, n);
    if (gt8) lz4_copy8_ss<8>(s, d, n);
    return;
  }
  lz4_copy4_ss<0>(s, d, n);
  if (gt4) {
    lz4_copy4_ss<4>(s, d, n);
    if (gt8) {
      lz4_copy4_ss<8>(s, d, n);
      lz4_copy4_ss<12>(s, d, n);
    }
  }
}
__device__ __forceinline__ void lz4_copy16_gs(const uint8_t *s, uint32_t d, uint32_t n) {
  lz4_copy4_gs<0>(s, d, n);
  lz4_copy4_gs<4>(s, d, n);
  if (__any_sync(0xffffffffu, n > 8)) {
    lz4_copy4_gs<8>(s, d, n);
    lz4_copy4_gs<12>(s, d, n);
  }
}

__device__ __forceinline__ int lz4_move(uint8_t *dst, uint32_t dlen, uint32_t stream_end, Lz4Shared *sh) {
  const uint32_t lane = threadIdx.x & 31;
  constexpr uint32_t OM = SB_LZ4_RING - 1, IM = SB_LZ4_INR - 1;
  const uint32_t sh_b = smem_u32(sh);
  const uint32_t in_b = sh_b + offsetof(Lz4Shared, in);
  Lz4Out o;
  o.out_b = sh_b + offsetof(Lz4Shared, out);
  o.ring = sh->out;
  o.dst = dst;
  o.fl = 0;
  o.vec = (uintptr_t(dst) & 15) == 0;
  const uint32_t out_b = o.out_b;
  uint32_t cons = 0, op_base = 0;
  int rc = 0;
#ifdef SB_LZ4_PROF
  uint32_t prof[12] = {0};
  struct ProfDump {
    uint32_t *p;
    __device__ ~ProfDump() {
      if ((threadIdx.x & 31) == 0)
        for (int i = 0; i < 12; ++i) atomicAdd(&g_lz4_prof[i], (unsigned long long)p[i]);
    }
  } prof_dump{prof};
#endif
  auto wait_in = [&](uint32_t upto) -> bool { // stream bytes [.., upto) in the ring
    for (;;) {
      if (lds_vol(sh_b + offsetof(Lz4Shared, in_ready)) >= upto) break;
      if (lds_vol(sh_b + offsetof(Lz4Shared, abort))) return false;
      __nanosleep(64);
    }
    return true;
  };
  // No fences on the mover side: its reads of the rings and its publications are shared-memory
  // accesses of ONE warp, which the LSU performs in program order; a MEMBAR here would also wait
  // for the write-behind global stores in flight (~1 us), on the critical path of every batch.
  auto release_in = [&](uint32_t q) {
    __syncwarp();
    if (lane == 0) sts_vol(sh_b + offsetof(Lz4Shared, m_q), q);
  };
  // whole-warp literal copy of stream bytes [ls, ls+lit) to output position op, streaming
  auto lit_coop = [&](uint32_t ls, uint32_t lit, uint32_t op, bool stream) -> bool {
    for (uint32_t done = 0; done < lit;) {
      uint32_t p = min(lit - done, 512u);
      if (stream) {
        release_in(ls + done); // BEFORE waiting: the scanner may need the ring space to deliver
        if (!wait_in(min(stream_end, ls + done + p))) return false;
      }
      for (uint32_t i = lane; i < p; i += 32) sts_u8(out_b + ((op + done + i) & OM), lds_u8(in_b + ((ls + done + i) & IM)));
      __syncwarp();
      done += p;
      if (op + done - o.fl >= SB_LZ4_FLUSHQ) o.flush_to(op + done, false);
    }
    if (stream) release_in(ls + lit);
    return true;
  };
  // length-extension bytes at stream position r, read as they arrive (BIG entries)
  auto ext_stream = [&](uint32_t &r, uint32_t &len) -> int {
    uint32_t x = 255, avail = 0;
    while (x == 255) {
      if (r >= stream_end) return SB_EXTERNAL;
      if (r >= avail) {
        release_in(r); // before waiting: the scanner may need the ring space to get further
        if (!wait_in(r + 1)) return -1;
        avail = lds_vol(sh_b + offsetof(Lz4Shared, in_ready));
      }
      x = lds_u8(in_b + (r++ & IM));
      len += x;
      if (len > SB_LZ4_MAXPOS) return SB_EXTERNAL;
    }
    return 0;
  };
  for (;;) {
    LZ4_T(t0);
    uint32_t prod;
    while ((prod = lds_vol(sh_b + offsetof(Lz4Shared, produced))) == cons) {
      if (lds_vol(sh_b + offsetof(Lz4Shared, abort))) return 0; // the scanner reports
      __nanosleep(32);
    }
    const uint32_t nb = min(prod - cons, 32u);
    LZ4_T(t1);
    LZ4_ACC(0, t0, t1);
#ifdef SB_LZ4_PROF
    prof[10] += 1;
    prof[11] += nb;
#endif
    // ---- one sequence per lane: parse
    uint32_t q = 0, kind = LZ4_E_SEQ, lit = 0, ml = 0, offset = 1, ls = 0, qn = 0;
    if (lane < nb) {
      const uint32_t e = lds_vol(sh_b + offsetof(Lz4Shared, q) + (((cons + lane) & (SB_LZ4_Q - 1)) << 2));
      q = e & SB_LZ4_MAXPOS;
      kind = e >> 30;
      const uint32_t tok = lds_u8(in_b + (q & IM));
      lit = tok >> 4;
      uint32_t mlc = tok & 15u, r = q + 1;
      if (kind == LZ4_E_BIG || kind == LZ4_E_ERROR) lit = 0; // BIG: parsed by the whole warp, streaming
      if (lit == 15) { // at most 4 length bytes (the scanner made longer ones BIG)
        uint32_t x;
        do {
          x = lds_u8(in_b + (r++ & IM));
          lit += x;
        } while (x == 255 && r < stream_end);
      }
      ls = r;
      r += lit;
      if (kind == LZ4_E_SEQ) {
        offset = lds_u8(in_b + (r & IM)) | (lds_u8(in_b + ((r + 1) & IM)) << 8);
        r += 2;
        ml = mlc;
        if (mlc == 15) {
          uint32_t x;
          do {
            x = lds_u8(in_b + (r++ & IM));
            ml += x;
          } while (x == 255 && r < stream_end);
        }
        ml += 4;
      }
      qn = kind == LZ4_E_SEQ ? r : q;
    }
    const uint32_t spec_mask = __ballot_sync(0xffffffffu, lane < nb && kind != LZ4_E_SEQ);
    LZ4_T(t2);
    LZ4_ACC(1, t1, t2);
    uint32_t i = 0;
    while (i < nb) {
      const uint32_t spec = spec_mask >> i;
      const uint32_t e = spec ? i + uint32_t(__ffs(int(spec))) - 1u : nb; // next special entry
      if (e == i) {
        // ---- END / BIG / ERROR entry: the whole warp handles it, streaming
        const uint32_t k_ = __shfl_sync(0xffffffffu, kind, i), q_ = __shfl_sync(0xffffffffu, q, i);
        uint32_t lit_ = __shfl_sync(0xffffffffu, lit, i), ls_ = __shfl_sync(0xffffffffu, ls, i);
        if (k_ == LZ4_E_ERROR) {
          rc = SB_EXTERNAL;
          break;
        }
        const bool big = k_ == LZ4_E_BIG;
        const uint32_t tok_ = lds_u8(in_b + (q_ & IM)); // before its ring slot is released
        release_in(q_);                                 // every earlier entry of the batch is done
        if (big) { // literal length, streaming
          uint32_t r = q_ + 1;
          lit_ = tok_ >> 4;
          if (lit_ == 15 && (rc = ext_stream(r, lit_)) != 0) break;
          ls_ = r;
        }
        if (lit_ > dlen - op_base) {
          rc = SB_EXTERNAL;
          break;
        }
        if (!lit_coop(ls_, lit_, op_base, true)) {
          rc = -1;
          break;
        }
        op_base += lit_;
        if (k_ == LZ4_E_END || ls_ + lit_ == stream_end) { // last sequence of the block
          if (op_base != dlen) rc = SB_EXTERNAL;
          else o.flush_to(dlen, true);
          if (rc && lane == 0) sts_vol(sh_b + offsetof(Lz4Shared, abort), 1u);
          return rc;
        }
        // offset and match length follow the literal run
        uint32_t r = ls_ + lit_;
        if (stream_end - r < 2) {
          rc = SB_EXTERNAL;
          break;
        }
        if (!wait_in(min(stream_end, r + 2))) {
          rc = -1;
          break;
        }
        const uint32_t off_ = lds_u8(in_b + (r & IM)) | (lds_u8(in_b + ((r + 1) & IM)) << 8);
        r += 2;
        uint32_t ml_ = tok_ & 15u;
        if (ml_ == 15 && (rc = ext_stream(r, ml_)) != 0) break;
        ml_ += 4;
        release_in(r);
        if (off_ == 0 || off_ > op_base || ml_ > dlen - op_base) {
          rc = SB_EXTERNAL;
          break;
        }
        o.match_coop(op_base, off_, ml_);
        op_base += ml_;
        ++i;
        continue;
      }
      // ---- segment [i, e) of ordinary sequences: output positions by a warp scan
      LZ4_T(t3);
      const bool in_seg = lane >= i && lane < e;
      const uint32_t len = in_seg ? lit + ml : 0u;
      const uint32_t incl = warp_incl_scan(len);
      const uint32_t op = op_base + incl - len, mpos = op + lit;
      const uint32_t seg_total = __shfl_sync(0xffffffffu, incl, 31);
      // lengths are < 2^12 each (longer ones are BIG), so the sums cannot wrap
      const bool bad = in_seg && (offset == 0 || offset > mpos || mpos + ml > dlen);
      if (__ballot_sync(0xffffffffu, bad)) {
        rc = SB_EXTERNAL;
        break;
      }
      const uint32_t small_mask = __ballot_sync(0xffffffffu, in_seg && lit <= SB_LZ4_SMALL_LIT && ml <= SB_LZ4_SMALL);
      LZ4_T(t4);
      LZ4_ACC(2, t3, t4);
      uint32_t a = i;
      while (a < e) {
        if (!(small_mask & (1u << a))) {
          // a long sequence: the whole warp moves it
          const uint32_t op_ = __shfl_sync(0xffffffffu, op, a), lit_ = __shfl_sync(0xffffffffu, lit, a);
          const uint32_t ls_ = __shfl_sync(0xffffffffu, ls, a), off_ = __shfl_sync(0xffffffffu, offset, a);
          const uint32_t ml_ = __shfl_sync(0xffffffffu, ml, a);
          lit_coop(ls_, lit_, op_, false);
          o.match_coop(op_ + lit_, off_, ml_);
          ++a;
          continue;
        }
        // run [a, j) of short sequences, one per lane
        LZ4_T(t5);
        const uint32_t rest = small_mask >> a;
        const uint32_t j = a + (~rest ? uint32_t(__ffs(int(~rest))) - 1u : 32u - a);
        const bool mine = lane >= a && lane < j;
        const uint32_t sp = mpos - offset;
        const uint32_t d_lit = op & OM, s_lit = ls & IM, d_m = mpos & OM;
        // literals of the whole run at once
        {
          const bool lean = mine && s_lit + lit <= SB_LZ4_INR && d_lit + lit <= SB_LZ4_RING;
          lz4_copy16_ss(in_b + s_lit, out_b + d_lit, lean ? min(lit, 16u) : 0u, __any_sync(0xffffffffu, mine && lit > 4),
                        __any_sync(0xffffffffu, mine && lit > 8), true);
          // literal runs of 17-32 bytes (sorted / slowly varying integer columns): a second per-lane pass instead of
          // moving those sequences one at a time with the whole warp
          if (__any_sync(0xffffffffu, lean && lit > 16)) {
            if (lean)
              for (uint32_t t = 16; t < lit; ++t) sts_u8(out_b + d_lit + t, lds_u8(in_b + s_lit + t)); // no extra live registers
          }
          if (__any_sync(0xffffffffu, mine && !lean) && mine && !lean) // a ring boundary inside: masked bytes
            for (uint32_t t = 0; t < lit; ++t) sts_u8(out_b + ((op + t) & OM), lds_u8(in_b + ((ls + t) & IM)));
        }
        LZ4_T(t6);
        LZ4_ACC(3, t5, t6);
        // Chains: fixed-width values make a match read the output of the previous match (offset ==
        // value width).  A match whose source lies entirely inside the destination of an earlier
        // match of this run reads that match's own source instead (pointer jumping, distances
        // 1,2,4,..: a chain of any length inside the run collapses in 5 steps), so the copies
        // below stay parallel.  Only exact containment in a non-overlapping match is redirected.
        uint32_t src = sp;
        {
          const uint32_t run_first = __shfl_sync(0xffffffffu, mpos, a);
          if (__any_sync(0xffffffffu, mine && lane > a && sp + min(ml, offset) > run_first)) {
            const bool plain = mine && offset >= ml;
#pragma unroll
            for (uint32_t d = 1; d < 32; d <<= 1) {
              const uint32_t k_mpos = __shfl_up_sync(0xffffffffu, mpos, d), k_ml = __shfl_up_sync(0xffffffffu, ml, d);
              const uint32_t k_src = __shfl_up_sync(0xffffffffu, src, d);
              const bool k_plain = __shfl_up_sync(0xffffffffu, int(plain), d) != 0;
              if (mine && lane >= a + d && k_plain && src >= k_mpos && src + ml <= k_mpos + k_ml) src = k_src + (src - k_mpos);
            }
          }
        }
        LZ4_T(t7);
        LZ4_ACC(4, t6, t7);
        const uint32_t off2 = mpos - src; // distance to the (possibly redirected) source
        const uint32_t s_m2 = src & OM;
        // matches whose source was flushed long ago never depend on anything pending: L2 -> ring
        const bool far = mine && off2 > SB_LZ4_NEAR + 32 && src + ml <= o.fl && d_m + ml <= SB_LZ4_RING;
        if (__any_sync(0xffffffffu, far)) lz4_copy16_gs(dst + src, out_b + d_m, far ? ml : 0u);
        __syncwarp();
        // the others run in independent-prefix rounds: everything before the first pending match is final
        LZ4_T(t8);
        LZ4_ACC(5, t7, t8);
        const bool nearl = mine && !far;
        const bool lean = nearl && off2 <= SB_LZ4_NEAR && (off2 >= 4 || off2 >= ml) && s_m2 + ml <= SB_LZ4_RING &&
                          d_m + ml <= SB_LZ4_RING;
        const bool gt4 = __any_sync(0xffffffffu, nearl && ml > 4), gt8 = __any_sync(0xffffffffu, nearl && ml > 8);
        const bool any_slow = __any_sync(0xffffffffu, nearl && !lean);
        const bool wide = !__any_sync(0xffffffffu, lean && off2 < 8 && off2 < ml);
        // K = last lane whose (pending) match this lane may read: mpos is increasing, so it is the
        // number of run lanes with mpos < source end, found by a 5-step binary search over the warp.
        // A lane can go as soon as the round starts beyond K.
        uint32_t K = 0;
        {
          const uint32_t s_end = src + min(ml, off2);
          uint32_t lo = a, hi = lane; // invariant: mpos[lo..] candidates; count lanes in [a, lane) with mpos < s_end
#pragma unroll
          for (int it = 0; it < 5; ++it) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t m_mid = __shfl_sync(0xffffffffu, mpos, mid & 31);
            if (lo < hi) {
              if (m_mid < s_end) lo = mid + 1;
              else hi = mid;
            }
          }
          K = lo; // lanes [a, lo) have mpos < s_end: this lane waits until the round start f >= lo
        }
        // L = first run lane whose match ends after this lane's source begins (mpos + ml is increasing too): the
        // pending matches this lane may read are exactly the lanes [L, K).  A lane copies as soon as none of them
        // is pending -- with liblz4-written blocks most near matches read literals or matches that are final, so a
        // batch takes as many rounds as its dependency DEPTH (2-3), not one per dependent match.
        uint32_t rmask = 0;
        {
          uint32_t lo = a, hi = lane;
          const uint32_t m_end = mpos + ml;
#pragma unroll
          for (int it = 0; it < 5; ++it) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t e_mid = __shfl_sync(0xffffffffu, m_end, mid & 31);
            if (lo < hi) {
              if (e_mid <= src) lo = mid + 1;
              else hi = mid;
            }
          }
          if (K > lo) rmask = ((1u << K) - 1u) & ~((1u << lo) - 1u); // K <= lane <= 31
        }
        bool pending = nearl;
        for (;;) {
          const uint32_t pend = __ballot_sync(0xffffffffu, pending);
          if (!pend) break;
          const bool inr = pending && (pend & rmask) == 0; // the lowest pending lane always qualifies
          lz4_copy16_ss(out_b + s_m2, out_b + d_m, inr && lean ? ml : 0u, gt4, gt8, wide);
          if (any_slow && inr && !lean) // overlap < 4, ring boundary, odd distances
            for (uint32_t t = 0; t < ml; ++t) sts_u8(out_b + ((mpos + t) & OM), o.src_byte(src + t, mpos));
          __syncwarp();
          if (inr) pending = false;
        }
        LZ4_T(t9);
        LZ4_ACC(6, t8, t9);
        const uint32_t op_end = __shfl_sync(0xffffffffu, mpos + ml, j - 1);
        if (op_end - o.fl >= SB_LZ4_FLUSHQ) o.flush_to(op_end, false);
        LZ4_T(t10);
        LZ4_ACC(7, t9, t10);
        a = j;
      }
      op_base += seg_total;
      i = e;
    }
    if (rc) break;
    LZ4_T(t11);
    cons += nb;
    const uint32_t q_next = __shfl_sync(0xffffffffu, qn, nb - 1);
    if (lane == 0) {
      sts_vol(sh_b + offsetof(Lz4Shared, m_q), q_next);
      sts_vol(sh_b + offsetof(Lz4Shared, consumed), cons);
    }
    LZ4_T(t12);
    LZ4_ACC(8, t11, t12);
  }
  if (lane == 0) sts_vol(sh_b + offsetof(Lz4Shared, abort), 1u);
  return rc > 0 ? rc : 0;
}

} // namespace sb
