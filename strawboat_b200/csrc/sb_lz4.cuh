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
// Warp-PAIR streaming decoder of the dedicated LZ4 kernel (one CTA of 64 threads per page).
//
// An LZ4 block is two dependent chains: the token chain (where does the next sequence
// start) and the match chain (a match may read what the previous match wrote).  One warp
// walking both pays ~100 dependent instructions per sequence.  Here they run on two warps:
//
//   producer (warp 0): input ring (global -> shared, cp.async prefetched), token walk,
//                      literal bytes straight into the shared output ring, one 8-byte
//                      descriptor {mpos, offset | ml << 16} per match into a shared queue;
//                      validates the stream (the consumer trusts descriptors).
//   consumer (warp 1): pops descriptors 32 at a time (one coalesced load, then shuffles),
//                      executes the match copies inside the ring and writes the ring behind
//                      to HBM in 16-byte vectors.
//
// Per sequence the producer's chain is token LDS -> 3 ALU -> next token LDS, the consumer's
// is source LDS -> STS; everything else is off the critical path.
//
// Flow control (shared counters, producer-published `produced`, consumer-published
// `consumed` / `flushed`):
//   * queue: the producer writes slot seq only while seq - consumed < SB_LZ4_Q;
//   * ring : the producer keeps every byte it (or a match it described) writes below
//            flushed + SB_LZ4_AHEAD, so ring bytes at distance <= SB_LZ4_NEAR behind any
//            match stay intact; older match sources come from the flushed global output.
// ------------------------------------------------------------------------------------
constexpr uint32_t SB_LZ4_Q = 128;                             // descriptor queue entries
constexpr uint32_t SB_LZ4_AHEAD = 4096;                        // producer lead over `flushed`
constexpr uint32_t SB_LZ4_NEAR = SB_LZ4_RING - SB_LZ4_AHEAD;   // ring-resident match distance
constexpr uint32_t SB_LZ4_FLUSHQ = 1024;                       // consumer write-behind granularity
enum { LZ4_D_END = 1, LZ4_D_ADVANCE = 2, LZ4_D_ERROR = 3 };    // control descriptors (offset == 0)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
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
__device__ __forceinline__ uint32_t lds_vol(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_vol(uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

struct __align__(16) Lz4PairShared {
  uint8_t out[SB_LZ4_RING];
  uint8_t in[SB_LZ4_IN];
  uint2 desc[SB_LZ4_Q];
  uint32_t produced; // descriptors published by the producer
  uint32_t consumed; // descriptors retired by the consumer
  uint32_t flushed;  // output bytes written behind to HBM
  uint32_t pad;
};

// input half: global -> shared ring, 1 KiB chunks, one chunk of prefetch in flight
struct Lz4In {
  const uint8_t *gal; // 16-byte aligned global base of the compressed stream
  uint32_t total;     // aligned stream bytes (multiple of 16)
  uint8_t *in;        // shared input ring
  uint32_t issued;    // stream bytes requested so far
  uint32_t ready;     // stream bytes known complete
  __device__ __forceinline__ void issue_chunk() {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t end = min(total, issued + SB_LZ4_CHUNK);
    for (uint32_t o = issued + lane * 16; o < end; o += 512) cp_async16(in + (o & (SB_LZ4_IN - 1)), gal + o);
    cp_async_commit();
    issued = end;
  }
  // make stream bytes [.., q_end) readable; keeps one chunk of prefetch in flight
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
  __device__ __forceinline__ uint32_t ib(uint32_t q) const { return in[q & (SB_LZ4_IN - 1)]; }
};

// ---- producer ------------------------------------------------------------------------
struct Lz4Producer {
  Lz4PairShared *sh;
  uint32_t seq;      // descriptors written
  uint32_t c_seen;   // last `consumed` read
  uint32_t f_seen;   // last `flushed` read
  __device__ __forceinline__ void publish() {
    __threadfence_block();
    if ((threadIdx.x & 31) == 0) sts_vol(&sh->produced, seq);
  }
  // Wait until the queue has `slots` free entries and output bytes below `op_end` may be
  // written.  When it has to wait it waits for real room (half the queue, 1 KiB of ring) so
  // the two warps exchange work in large batches instead of ping-ponging per sequence; the
  // idle consumer always satisfies it: consumed == seq and flushed > op - FLUSHQ - 16.
  __device__ __forceinline__ void wait_room(uint32_t slots, uint32_t op_end) {
    if (seq + slots - c_seen <= SB_LZ4_Q && op_end <= f_seen + SB_LZ4_AHEAD) return;
    publish();
    for (;;) {
      c_seen = lds_vol(&sh->consumed);
      f_seen = lds_vol(&sh->flushed);
      if (seq + slots - c_seen <= SB_LZ4_Q / 2 && op_end + 1024 <= f_seen + SB_LZ4_AHEAD) break;
      __nanosleep(32);
    }
  }
  __device__ __forceinline__ void push(uint32_t x, uint32_t y) { // room must have been waited for
    if ((threadIdx.x & 31) == 0) sh->desc[seq & (SB_LZ4_Q - 1)] = make_uint2(x, y);
    ++seq;
  }
};

__device__ int lz4_pair_produce(const uint8_t *src, uint32_t clen, uint32_t dlen, Lz4PairShared *sh) {
  const uint32_t lane = threadIdx.x & 31;
  constexpr uint32_t IM = SB_LZ4_IN - 1, OM = SB_LZ4_RING - 1;
  Lz4In s;
  const uint32_t mis = uint32_t(uintptr_t(src) & 15);
  s.gal = src - mis;
  s.total = (mis + clen + 15) & ~15u;
  s.in = sh->in;
  s.issued = 0;
  s.ready = 0;
  s.issue_chunk();
  if (s.issued < s.total) s.issue_chunk();
  Lz4Producer pr{sh, 0, 0, 0};
  const uint32_t in_b = smem_u32(sh->in), out_b = smem_u32(sh->out), desc_b = smem_u32(sh->desc);
  uint32_t ip = 0, op = 0;
  uint32_t ip_lim = 0, op_lim = 0, seq_lim = 0; // housekeeping is due when a limit is reached
  uint32_t tok = 0, b = 0;
  bool reload = true;
  int rc = 0;
  for (;;) {
    if (ip >= ip_lim || op >= op_lim || pr.seq >= seq_lim || reload) {
      // ---- housekeeping: input prefetch / completion, flow control against the consumer
      if (ip >= clen) {
        rc = SB_EXTERNAL;
        break;
      }
      uint32_t q = mis + ip;
      s.ensure(q, min(s.total, q + 48));
      pr.wait_room(1, op + 48);
      ip_lim = (s.ready >= s.total) ? 0xffffffffu : min(s.ready - 48, s.issued - SB_LZ4_CHUNK) - mis;
      op_lim = pr.f_seen + SB_LZ4_AHEAD - 47; // fast sequences write at most 33 bytes
      seq_lim = pr.c_seen + SB_LZ4_Q;
      tok = lds_u8(in_b + (q & IM));
      b = lds_u8(in_b + ((q + 1 + lane) & IM));
      reload = false;
    }
    uint32_t lit = tok >> 4, mlc = tok & 15u;
    uint32_t nip = ip + 3 + lit;
    if (lit != 15 && mlc != 15 && nip <= clen) {
      // ---- fast sequence: token, <= 14 literals and the offset sit in the lane window
      uint32_t nq = mis + nip;
      uint32_t tok_n = lds_u8(in_b + (nq & IM));
      uint32_t b_n = lds_u8(in_b + ((nq + 1 + lane) & IM));
      uint32_t offset = __shfl_sync(0xffffffffu, b, lit) | (__shfl_sync(0xffffffffu, b, lit + 1) << 8);
      uint32_t ml = mlc + 4;
      uint32_t mpos = op + lit, nop = mpos + ml;
      if (lane < lit) sts_u8(out_b + ((op + lane) & OM), b);
      if (nop > dlen || offset == 0 || offset > mpos) {
        rc = SB_EXTERNAL;
        break;
      }
      if (lane == 0) {
        uint32_t y = offset | (ml << 16);
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(desc_b + ((pr.seq & (SB_LZ4_Q - 1)) << 3)), "r"(mpos), "r"(y)
                     : "memory");
      }
      ++pr.seq;
      if ((pr.seq & 15u) == 0) pr.publish();
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
      uint32_t x = 0;
      do {
        if (ip >= clen) {
          rc = SB_EXTERNAL;
          break;
        }
        s.ensure(mis + ip, mis + ip + 1);
        x = s.ib(mis + ip);
        ++ip;
        lit += x;
      } while (x == 255);
      if (rc) break;
    }
    if (ip > clen || lit > clen - ip || lit > dlen - op) {
      rc = SB_EXTERNAL;
      break;
    }
    for (uint32_t done = 0; done < lit;) {
      uint32_t p = min(lit - done, 512u);
      s.ensure(mis + ip, mis + ip + p);
      pr.wait_room(1, op + p);
      for (uint32_t i = lane; i < p; i += 32) sts_u8(out_b + ((op + i) & OM), s.ib(mis + ip + i));
      __syncwarp();
      ip += p;
      op += p;
      done += p;
      pr.push(op, LZ4_D_ADVANCE << 16);
    }
    if (ip == clen) break; // last sequence carries literals only
    if (clen - ip < 2) {
      rc = SB_EXTERNAL;
      break;
    }
    s.ensure(mis + ip, mis + ip + 2);
    uint32_t offset = s.ib(mis + ip) | (s.ib(mis + ip + 1) << 8);
    ip += 2;
    if (offset == 0 || offset > op) {
      rc = SB_EXTERNAL;
      break;
    }
    uint32_t ml = mlc;
    if (ml == 15) {
      uint32_t x = 0;
      do {
        if (ip >= clen) {
          rc = SB_EXTERNAL;
          break;
        }
        s.ensure(mis + ip, mis + ip + 1);
        x = s.ib(mis + ip);
        ++ip;
        ml += x;
      } while (x == 255);
      if (rc) break;
    }
    ml += 4;
    if (ml > dlen - op) {
      rc = SB_EXTERNAL;
      break;
    }
    for (uint32_t done = 0; done < ml;) { // long matches travel as pieces (sources precede each piece)
      uint32_t p = min(ml - done, 1024u);
      pr.wait_room(1, op + p);
      pr.push(op, offset | (p << 16));
      op += p;
      done += p;
    }
  }
  if (rc == 0 && op != dlen) rc = SB_EXTERNAL;
  pr.wait_room(1, 0);
  pr.push(op, uint32_t(rc ? LZ4_D_ERROR : LZ4_D_END) << 16);
  pr.publish();
  return rc;
}

// ---- consumer ------------------------------------------------------------------------
struct Lz4Out {
  uint8_t *ring; // shared output ring
  uint8_t *dst;  // global output
  uint32_t fl;   // bytes [0, fl) written to dst
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
};

__device__ void lz4_pair_consume(uint8_t *dst, Lz4PairShared *sh) {
  const uint32_t lane = threadIdx.x & 31;
  constexpr uint32_t OM = SB_LZ4_RING - 1;
  Lz4Out o{sh->out, dst, 0, (uintptr_t(dst) & 15) == 0};
  const uint32_t out_b = smem_u32(sh->out);
  uint32_t cons = 0;
  for (;;) {
    uint32_t prod;
    while ((prod = lds_vol(&sh->produced)) == cons) __nanosleep(20);
    __threadfence_block();
    uint32_t nb = min(prod - cons, 32u);
    uint2 d = make_uint2(0, 0);
    if (lane < nb) d = sh->desc[(cons + lane) & (SB_LZ4_Q - 1)];
    for (uint32_t k = 0; k < nb; ++k) {
      uint32_t mpos = __shfl_sync(0xffffffffu, d.x, k), y = __shfl_sync(0xffffffffu, d.y, k);
      uint32_t offset = y & 0xffffu, ml = y >> 16;
      if (offset == 0) { // control descriptor
        if (ml == LZ4_D_ADVANCE) {
          if (mpos - o.fl >= SB_LZ4_FLUSHQ) o.flush_to(mpos, false);
          continue;
        }
        if (ml == LZ4_D_END) o.flush_to(mpos, true);
        return;
      }
      if (ml <= 32 && offset >= ml && offset <= SB_LZ4_NEAR) {
        // common: source inside the ring, no overlap with the bytes being written
        if (lane < ml) sts_u8(out_b + ((mpos + lane) & OM), lds_u8(out_b + ((mpos - offset + lane) & OM)));
        __syncwarp();
      } else if (offset <= SB_LZ4_NEAR) {
        // ring-resident, long or self-overlapping: 32-byte steps (a step's sources precede it
        // when offset >= 32; shorter offsets replicate a pattern: index modulo offset)
        if (offset >= 32) {
          for (uint32_t i = 0; i < ml; i += 32) {
            if (i + lane < ml) sts_u8(out_b + ((mpos + i + lane) & OM), lds_u8(out_b + ((mpos - offset + i + lane) & OM)));
            __syncwarp();
          }
        } else {
          uint32_t v = lds_u8(out_b + ((mpos - offset + lane % offset) & OM)); // pattern byte of lane
          uint32_t step = 32 - 32 % offset;                                   // multiple of offset
          for (uint32_t i = 0; i < ml; i += step)
            if (lane < step && i + lane < ml) sts_u8(out_b + ((mpos + i + lane) & OM), v);
          __syncwarp();
        }
      } else {
        // far source: read it from the flushed global output
        if (mpos - offset + min(ml, offset) > o.fl) o.flush_to(mpos, true);
        for (uint32_t i = 0; i < ml; i += 32) {
          uint32_t n_i = min(32u, ml - i);
          // offset > NEAR >= 32: sources of a 32-byte step never overlap its destination,
          // but may not be flushed yet when offset < ml: those bytes are still in the ring
          if (lane < n_i) {
            uint32_t sp = mpos - offset + i + lane;
            uint32_t v = sp < o.fl ? uint32_t(__ldcg(dst + sp)) : lds_u8(out_b + (sp & OM));
            sts_u8(out_b + ((mpos + i + lane) & OM), v);
          }
          __syncwarp();
        }
      }
      if (mpos + ml - o.fl >= SB_LZ4_FLUSHQ) o.flush_to(mpos + ml, false);
    }
    cons += nb;
    __threadfence_block();
    if (lane == 0) {
      sts_vol(&sh->flushed, o.fl);
      sts_vol(&sh->consumed, cons);
    }
  }
}

} // namespace sb
