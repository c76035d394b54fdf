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
// Streaming warp decoder of the dedicated LZ4 kernel.
//   input : global -> shared input ring, 2 KiB halves prefetched with cp.async (LDGSTS), so
//           token / literal reads are shared-memory loads that never wait on HBM
//   output: shared output ring (last SB_LZ4_RING bytes) written behind to HBM in 16-byte
//           vectors; match sources older than the ring come from the flushed global output
// Per fast sequence: one broadcast token load + one byte load per lane (lane i holds
// stream byte ip+1+i, i.e. its own literal), the offset arrives by two shuffles.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct Lz4Stream {
  // input side
  const uint8_t *gal; // 16-byte aligned global base of the compressed stream
  uint32_t total;     // aligned stream bytes (multiple of 16)
  uint8_t *in;        // shared input ring
  uint32_t issued;    // stream bytes requested so far (multiple of 2048, or total)
  uint32_t ready;     // stream bytes known complete
  // output side
  uint8_t *ring;      // shared output ring
  uint8_t *dst;       // global output
  uint32_t fl;        // bytes [0, fl) flushed to dst
  uint32_t ring_from; // output bytes >= ring_from are (or will be) present in the ring
  bool vec;

  __device__ __forceinline__ void issue_chunk() { // request the next 2 KiB (or the tail)
    const uint32_t lane = threadIdx.x & 31;
    uint32_t end = min(total, issued + SB_LZ4_CHUNK);
    for (uint32_t o = issued + lane * 16; o < end; o += 512) cp_async16(in + (o & (SB_LZ4_IN - 1)), gal + o);
    cp_async_commit();
    issued = end;
  }
  // make stream bytes [0, q_end) readable; keeps one chunk of prefetch in flight
  __device__ __forceinline__ void ensure(uint32_t q, uint32_t q_end) {
    if (issued < total && q + SB_LZ4_CHUNK >= issued) { // the half before `q`'s half is free again
      __syncwarp();
      issue_chunk();
    }
    if (q_end > ready) {
      cp_async_wait_all();
      __syncwarp();
      ready = issued;
    }
  }
  __device__ __forceinline__ void in_reset(uint32_t q) { // restart streaming at stream position q
    cp_async_wait_all();
    __syncwarp();
    issued = q & ~(SB_LZ4_CHUNK - 1);
    ready = issued;
    issue_chunk();
    if (issued < total) issue_chunk();
  }
  __device__ __forceinline__ uint32_t ib(uint32_t q) const { return in[q & (SB_LZ4_IN - 1)]; }
  __device__ __forceinline__ void st(uint32_t pos, uint32_t b) { ring[pos & (SB_LZ4_RING - 1)] = uint8_t(b); }
  __device__ __forceinline__ uint32_t ld(uint32_t pos, uint32_t op) const {
    if (op - pos <= SB_LZ4_RING - SB_LZ4_PIECE - 64 && pos >= ring_from) return ring[pos & (SB_LZ4_RING - 1)];
    return __ldcg(dst + pos);
  }
  // write ring bytes [fl, upto) behind to HBM.  Non-final flushes move whole 16-byte vectors
  // only (fl stays 16-byte aligned); the final flush also writes the byte tail.
  __device__ __forceinline__ void flush_to(uint32_t upto, bool final) {
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
    if (vec) {
      uint32_t a = min(upto, (fl + 15) & ~15u);
      for (uint32_t pos = fl + lane; pos < a; pos += 32) dst[pos] = ring[pos & (SB_LZ4_RING - 1)];
      fl = a;
      uint32_t vend = upto & ~15u;
      for (uint32_t pos = fl + lane * 16; pos + 16 <= vend; pos += 512)
        *reinterpret_cast<uint4 *>(dst + pos) = *reinterpret_cast<const uint4 *>(ring + (pos & (SB_LZ4_RING - 1)));
      if (vend > fl) fl = vend;
    }
    if (final || !vec) {
      for (uint32_t pos = fl + lane; pos < upto; pos += 32) dst[pos] = ring[pos & (SB_LZ4_RING - 1)];
      if (fl < upto) fl = upto;
    }
    __syncwarp();
  }
  __device__ __forceinline__ void advance(uint32_t op) {
    if (op - fl >= SB_LZ4_FLUSH) flush_to(op, false);
  }
};

__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// The fast path is written for a short dependent chain and few branches (a single warp
// retires one dependent instruction every ~6 cycles, so instruction count == latency):
//   * token/literal bytes of the NEXT sequence are loaded before the current copy,
//   * input refill + output flush are folded into one rarely taken housekeeping branch,
//   * all validity / near-source conditions are uniform and folded into one predicate.
__device__ int lz4_decode_stream(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t dlen, uint8_t *in_ring,
                                 uint8_t *out_ring) {
  const uint32_t lane = threadIdx.x & 31;
  if (clen == 0) return dlen == 0 ? 0 : SB_EXTERNAL;
  constexpr uint32_t IM = SB_LZ4_IN - 1, OM = SB_LZ4_RING - 1;
  constexpr uint32_t NEAR = SB_LZ4_RING - SB_LZ4_PIECE - 64;
  Lz4Stream s;
  const uint32_t mis = uint32_t(uintptr_t(src) & 15);
  s.gal = src - mis;
  s.total = (mis + clen + 15) & ~15u;
  s.in = in_ring;
  s.ring = out_ring;
  s.dst = dst;
  s.fl = 0;
  s.ring_from = 0;
  s.vec = (uintptr_t(dst) & 15) == 0;
  s.issued = 0;
  s.ready = 0;
  s.issue_chunk();
  if (s.issued < s.total) s.issue_chunk();
  const uint32_t in_b = smem_u32(in_ring), out_b = smem_u32(out_ring);
  uint32_t ip = 0, op = 0;
  uint32_t ip_lim = 0, op_lim = 0; // housekeeping is due when ip >= ip_lim or op >= op_lim
  uint32_t tok = 0, b = 0;
  bool reload = true;
  for (;;) {
    if (ip >= ip_lim || op >= op_lim || reload) {
      // ---- housekeeping: input prefetch / completion, output write-behind, limits
      if (ip >= clen) return SB_EXTERNAL;
      uint32_t q = mis + ip;
      s.ensure(q, min(s.total, q + 48));
      s.advance(op);
      ip_lim = (s.ready >= s.total) ? 0xffffffffu : min(s.ready - 48, s.issued - SB_LZ4_CHUNK) - mis;
      op_lim = s.fl + SB_LZ4_FLUSH;
      tok = lds_u8(in_b + (q & IM));
      b = lds_u8(in_b + ((q + 1 + lane) & IM));
      reload = false;
    }
    uint32_t lit = tok >> 4, mlc = tok & 15u;
    uint32_t nip = ip + 3 + lit;
    if (lit <= 12 && mlc != 15 && nip <= clen) {
      // speculative loads for the next sequence (ring reads are always in bounds)
      uint32_t nq = mis + nip;
      uint32_t tok_n = lds_u8(in_b + (nq & IM));
      uint32_t b_n = lds_u8(in_b + ((nq + 1 + lane) & IM));
      uint32_t offset = __shfl_sync(0xffffffffu, b, lit) | (__shfl_sync(0xffffffffu, b, lit + 1) << 8);
      uint32_t ml = mlc + 4;
      uint32_t mpos = op + lit, nop = mpos + ml;
      if (nop > dlen || nop < op) return SB_EXTERNAL;
      if (lane < lit) sts_u8(out_b + ((op + lane) & OM), b);
      __syncwarp();
      if (offset >= ml && offset <= NEAR && offset <= mpos - s.ring_from) {
        // common: source entirely inside the ring, no overlap with the bytes being written
        if (lane < ml) sts_u8(out_b + ((mpos + lane) & OM), lds_u8(out_b + ((mpos - offset + lane) & OM)));
      } else {
        if (offset == 0 || offset > mpos) return SB_EXTERNAL;
        if (lane < ml) {
          uint32_t j = offset < ml ? lane % offset : lane;
          s.st(mpos + lane, s.ld(mpos - offset + j, mpos));
        }
      }
      __syncwarp();
      ip = nip;
      op = nop;
      tok = tok_n;
      b = b_n;
      continue;
    }
    // ---- general path: length extensions, long runs, stream tail
    reload = true;
    ++ip;
    if (lit == 15) {
      uint32_t x;
      do {
        if (ip >= clen) return SB_EXTERNAL;
        s.ensure(mis + ip, mis + ip + 1);
        x = s.ib(mis + ip);
        ++ip;
        lit += x;
      } while (x == 255);
    }
    if (lit > clen - ip || lit > dlen - op) return SB_EXTERNAL;
    if (lit >= 1024) {
      // long literal run: bypass both rings, global -> global in 16-byte vectors
      s.flush_to(op, true);
      const uint8_t *sp = src + ip;
      uint8_t *dp = dst + op;
      uint32_t head = min(lit, uint32_t((16 - (uintptr_t(dp) & 15)) & 15));
      for (uint32_t i = lane; i < head; i += 32) dp[i] = sp[i];
      uint32_t nvec = (lit - head) >> 4;
      for (uint32_t v = lane; v < nvec; v += 32)
        *reinterpret_cast<uint4 *>(dp + head + (v << 4)) = ld_u128u(sp + head + (v << 4));
      for (uint32_t i = head + (nvec << 4) + lane; i < lit; i += 32) dp[i] = sp[i];
      __syncwarp();
      ip += lit;
      op += lit;
      s.fl = op;
      s.ring_from = op; // those bytes are not in the ring: later matches read them from global
      if (ip < clen) s.in_reset(mis + ip);
    } else {
      for (uint32_t done = 0; done < lit;) {
        uint32_t p = min(lit - done, 512u);
        s.ensure(mis + ip, mis + ip + p);
        for (uint32_t i = lane; i < p; i += 32) s.st(op + i, s.ib(mis + ip + i));
        __syncwarp();
        ip += p;
        op += p;
        done += p;
        s.advance(op);
      }
    }
    if (ip == clen) break; // last sequence carries literals only
    if (clen - ip < 2) return SB_EXTERNAL;
    s.ensure(mis + ip, mis + ip + 2);
    uint32_t offset = s.ib(mis + ip) | (s.ib(mis + ip + 1) << 8);
    ip += 2;
    if (offset == 0 || offset > op) return SB_EXTERNAL;
    uint32_t ml = mlc;
    if (ml == 15) {
      uint32_t x;
      do {
        if (ip >= clen) return SB_EXTERNAL;
        s.ensure(mis + ip, mis + ip + 1);
        x = s.ib(mis + ip);
        ++ip;
        ml += x;
      } while (x == 255);
    }
    ml += 4;
    if (ml > dlen - op) return SB_EXTERNAL;
    for (uint32_t done = 0; done < ml;) {
      uint32_t p = min(ml - done, SB_LZ4_PIECE);
      if (offset >= p) {
        for (uint32_t i = lane; i < p; i += 32) s.st(op + i, s.ld(op - offset + i, op));
      } else {
        for (uint32_t i = lane; i < p; i += 32) s.st(op + i, s.ld(op - offset + (i % offset), op));
      }
      __syncwarp();
      op += p;
      done += p;
      s.advance(op);
    }
  }
  s.flush_to(op, true);
  return op == dlen ? 0 : SB_EXTERNAL;
}

} // namespace sb
