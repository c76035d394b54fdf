// sb_lz4x.cuh -- LZ4 block decode for the dedicated kernel (basic.rs:87-91 -> LZ4_decompress_safe with a known
// decoded size), second generation: one CTA of 3 warps per block.
//
//   scanner (warp 0)      owns the input ring (global -> shared, cp.async) and walks the token chain: every lane
//                         treats its byte of a 32-byte window as a token, the real chain is followed with ONE SHFL
//                         per sequence; tokens with length bytes and the block tail take a scalar path with all the
//                         format checks.  It fills BATCHES of up to 64 (token position, token byte) entries.
//   workers (warps 1-2)   take a batch, ONE SEQUENCE PER THREAD: parse (lengths, offset) straight from the stream
//                         in L2, prefix-sum the lengths over the 64 threads -> output positions, copy every literal
//                         run and every match whose source precedes the batch concurrently, then the (few, with
//                         liblz4-written blocks) matches that read this batch's own output in order.  Output goes
//                         through an 8 KiB shared-memory ring that is written behind to HBM in 16-byte vectors;
//                         sources further back than the ring are read back from L2.
//
// Hand-off is by NAMED BARRIERS only (bar.arrive / bar.sync, two batch buffers): FULL[b] scanner -> workers,
// EMPTY[b] workers -> scanner, TEAM between the two worker warps.  No polled flags: compute-sanitizer racecheck
// sees every shared-memory dependency ordered by a barrier.
//
// Measured reason for this shape (profiles/README.md, round 2): liblz4-written blocks of numeric columns hold
// ~8000 sequences of ~8 bytes whose match offsets are spread over the whole 64 KiB window, and only ~2 % of the
// matches read the previous sequence's output -- so sequences are almost all independent, and the cost is parsing
// and moving them, not ordering them.  The first-generation mover handled 32 sequences per ~4500 cycles on one
// warp; here 64 threads share a batch and nothing in the common path is serial.
#pragma once
#include "sb_lz4.cuh"

namespace sb {

constexpr uint32_t LZX_RING = 8192;      // output ring bytes
constexpr uint32_t LZX_INR = 4096;       // scanner's input ring
constexpr uint32_t LZX_INCH = 1024;      // refill granularity
constexpr uint32_t LZX_B = 64;           // sequences per batch = worker threads
constexpr uint32_t LZX_FAST_MAX = 4096;  // output bytes of a batch handled thread-per-sequence
constexpr uint32_t LZX_PER_THREAD = 32;  // literal / match bytes a thread moves by itself
constexpr uint32_t LZX_FLUSHQ = 2048;    // write-behind granularity
constexpr uint32_t LZX_THREADS = 96;

enum { LZX_BAR_FULL0 = 1, LZX_BAR_FULL1 = 2, LZX_BAR_EMPTY0 = 3, LZX_BAR_EMPTY1 = 4, LZX_BAR_TEAM = 5 };

struct __align__(16) Lz4xShared {
  uint8_t out[LZX_RING];
  uint8_t in[LZX_INR];
  uint32_t ent[2][LZX_B]; // stream position of the token | kind << 30
  uint32_t tok[2][LZX_B]; // the token byte
  uint32_t nb[2];         // entries in the batch
  uint32_t last[2];       // 1 = no batch follows
  uint32_t rc[2];         // scanner's verdict travelling with the last batch (0 / SB_EXTERNAL)
  // worker scratch
  uint32_t wsum[2];       // per-warp totals of the length scan
  uint32_t dep_mask[2];   // per-warp ballots: matches that read this batch's output
  uint32_t long_lit[2], long_m[2];
  uint32_t p_mpos[LZX_B], p_off[LZX_B], p_ml[LZX_B], p_ls[LZX_B], p_lit[LZX_B];
  uint32_t pend[LZX_FAST_MAX / 32 + 2]; // bit i: output byte op_base + i is still owed by a dependent match
  uint32_t ready_w[2], left_w[2];
  uint32_t bad;           // worker-side validation failure of the current batch
  uint32_t job;
  uint32_t pad;
};

// barrier ids are immediates (ptxas then reserves 6 barriers per CTA, not all 16)
template <int ID, int N> __device__ __forceinline__ void lzx_bar_sync_i() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
template <int ID, int N> __device__ __forceinline__ void lzx_bar_arrive_i() { asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
__device__ __forceinline__ void lzx_bar_sync(uint32_t id, uint32_t) {
  switch (id) {
  case 1: lzx_bar_sync_i<1, 96>(); break;
  case 2: lzx_bar_sync_i<2, 96>(); break;
  case 3: lzx_bar_sync_i<3, 96>(); break;
  case 4: lzx_bar_sync_i<4, 96>(); break;
  default: lzx_bar_sync_i<5, 64>(); break;
  }
}
__device__ __forceinline__ void lzx_bar_arrive(uint32_t id, uint32_t) {
  switch (id) {
  case 1: lzx_bar_arrive_i<1, 96>(); break;
  case 2: lzx_bar_arrive_i<2, 96>(); break;
  case 3: lzx_bar_arrive_i<3, 96>(); break;
  default: lzx_bar_arrive_i<4, 96>(); break;
  }
}

// ---- scanner ---------------------------------------------------------------------------------------------
// Ring sharing: the workers parse a sequence from the scanner's input ring when its entry carries LZX_IN_RING
// (token, length bytes, literals, offset: the whole stream extent [q, next token) is at most LZX_EXTENT bytes and
// resident).  The scanner keeps those bytes until the batch has been consumed (EMPTY): requests never overwrite
// ring slots at or above the first token of an unconsumed batch, and a batch is closed once it spans LZX_SPAN
// stream bytes, so two batches plus the look-ahead always fit.  Longer sequences are read from L2 by the workers.
constexpr uint32_t LZX_IN_RING = 0x100u;
constexpr uint32_t LZX_EXTENT = 160;
constexpr uint32_t LZX_SPAN = 768;

struct Lz4xScan {
  Lz4xShared *sh;
  uint32_t in_b;      // shared address of the input ring
  const uint8_t *gal; // 16-byte aligned global base of the stream
  uint32_t total;     // aligned stream bytes
  uint32_t issued, ready;
  uint32_t own;       // lowest stream byte still read by the scanner itself
  uint32_t n;         // entries in the batch being filled
  uint32_t batch;     // index of that batch (buffer = batch & 1)
  uint32_t consumed;  // batches whose EMPTY arrival has been waited for
  uint32_t qf0, qf1;  // first token of the batch in buffer 0 / 1

  __device__ __forceinline__ uint32_t ib(uint32_t q) const { return lds_u8(in_b + (q & (LZX_INR - 1))); }
  __device__ __forceinline__ uint32_t qf(uint32_t b) const { return b ? qf1 : qf0; }
  // lowest stream byte that must stay in the ring
  __device__ __forceinline__ uint32_t floor_pos() const {
    uint32_t f = own;
    if (consumed < batch) f = min(f, qf((batch - 1) & 1)); // published, maybe not parsed yet
    if (n) f = min(f, qf(batch & 1));
    return f;
  }
  __device__ __forceinline__ bool issue_allowed() {
    const uint32_t lane = threadIdx.x & 31;
    bool any = false;
    const uint32_t keep = floor_pos();
    if (keep >= issued + 2 * LZX_INCH && issued == ready) // a long literal run was skipped: its bytes are not staged at all
      issued = ready = (keep / LZX_INCH - 1) * LZX_INCH;
    // the chunk at `issued` lands on the ring slots of stream bytes [issued - INR, issued - INR + INCH);
    // no two copies in flight onto the same slots: the chunks in [ready, issued + INCH) must fit the ring
    while (issued < total && (issued < LZX_INR || keep + (LZX_INR - LZX_INCH) >= issued) && issued + LZX_INCH - ready <= LZX_INR) {
      const uint32_t end = min(total, issued + LZX_INCH);
      for (uint32_t o = issued + lane * 16; o < end; o += 512) cp_async16(in_b + (o & (LZX_INR - 1)), gal + o);
      issued = end;
      any = true;
    }
    if (any) cp_async_commit();
    return any;
  }
  __device__ __forceinline__ void wait_consumed(uint32_t upto) {
    while (consumed < upto) {
      lzx_bar_sync(LZX_BAR_EMPTY0 + (consumed & 1), LZX_THREADS);
      ++consumed;
    }
  }
  // hand the current batch to the workers and open the next buffer
  __device__ __forceinline__ void publish(bool last, uint32_t rc) {
    const uint32_t b = batch & 1;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
      sh->nb[b] = n;
      sh->last[b] = last ? 1u : 0u;
      sh->rc[b] = rc;
    }
    __syncwarp();
    lzx_bar_arrive(LZX_BAR_FULL0 + b, LZX_THREADS);
    ++batch;
    n = 0;
    if (!last && batch >= 2) wait_consumed(batch - 1); // the buffer of batch - 2 is free again
  }
  // make stream bytes [.., upto) readable.  When the ring is held by unconsumed batches, wait for them (closing
  // the current batch if need be).  false = the request cannot fit (a logic error; reported, never spun on).
  __device__ __forceinline__ bool need(uint32_t upto) {
    upto = min(upto, total);
    while (ready < upto) {
      if (issued > ready) {
        cp_async_wait_all();
        __syncwarp();
        ready = issued;
        continue;
      }
      if (issue_allowed()) continue;
      if (consumed < batch) wait_consumed(batch);
      else if (n) publish(false, 0);
      else return false;
    }
    return true;
  }
  __device__ __forceinline__ void mark_first(uint32_t q) {
    if (n == 0) {
      if (batch & 1) qf1 = q;
      else qf0 = q;
    }
  }
  __device__ __forceinline__ void push(uint32_t q, uint32_t kind, uint32_t token) { // room must exist
    mark_first(q);
    if ((threadIdx.x & 31) == 0) {
      sh->ent[batch & 1][n] = q | (kind << 30);
      sh->tok[batch & 1][n] = token;
    }
    ++n;
  }
};

// Walks the whole block; returns after the LAST batch has been published and every EMPTY arrival of the workers
// has been consumed (so the barriers are balanced for the next job).
__device__ __forceinline__ void lz4x_scan(const uint8_t *src, uint32_t clen, Lz4xShared *sh) {
  const uint32_t lane = threadIdx.x & 31;
  constexpr uint32_t IM = LZX_INR - 1;
  Lz4xScan s;
  s.sh = sh;
  s.in_b = smem_u32(sh->in);
  const uint32_t mis = uint32_t(uintptr_t(src) & 15);
  s.gal = src - mis;
  s.total = (mis + clen + 15) & ~15u;
  s.issued = s.ready = 0;
  s.own = mis;
  s.n = 0;
  s.batch = s.consumed = 0;
  s.qf0 = s.qf1 = 0;
  const uint32_t end = mis + clen; // stream positions are offsets from gal
  uint32_t q = mis, q_lim = 0;
  int rc = 0;
  bool done = false;
  s.issue_allowed();
  while (!done && rc == 0) {
    if (s.n && q - s.qf(s.batch & 1) >= LZX_SPAN) s.publish(false, 0); // bounded stream span per batch
    if (q >= q_lim) {
      // ---- housekeeping: completed prefetches, new prefetches
      s.own = q;
      if (s.issued > s.ready) {
        cp_async_wait_all();
        __syncwarp();
        s.ready = s.issued;
      }
      s.issue_allowed();
      if (!s.need(min(end, q + 64))) {
        rc = SB_EXTERNAL;
        break;
      }
      // windows need 48 readable bytes and stay clear of the last 64 bytes of the block
      const uint32_t lim_ready = s.ready >= 48 ? s.ready - 48 : 0, lim_end = end >= 64 ? end - 64 : 0;
      q_lim = min(min(lim_ready, lim_end), q + LZX_SPAN);
    }
    if (q < q_lim) {
      if (s.n + 11 > LZX_B) s.publish(false, 0);
      // ---- window: lane i decodes byte q+i as a token; the chain hops with one SHFL per token
      const uint32_t b = lds_u8(s.in_b + ((q + lane) & IM));
      const uint32_t lit = b >> 4;
      const uint32_t pack = (lane + 3 + lit) | ((lit == 15 || (b & 15u) == 15) ? 0x100u : 0u) | (b << 16);
      uint32_t p = 0, cnt = 0, myp = 0, mytok = 0;
#pragma unroll
      for (uint32_t h = 0; h < 11; ++h) { // a sequence takes >= 3 stream bytes: <= 11 tokens in 32 bytes
        const uint32_t v = __shfl_sync(0xffffffffu, pack, p);
        if (v & 0x100u) break; // length bytes follow this token: scalar path
        if (lane == h) {
          myp = p;
          mytok = v >> 16;
        }
        p = v & 0xffu;
        cnt = h + 1;
        if (p >= 32) break;
      }
      if (cnt) {
        s.mark_first(q);
        if (lane < cnt) {
          sh->ent[s.batch & 1][s.n + lane] = (q + myp) | (uint32_t(LZ4_E_SEQ) << 30);
          sh->tok[s.batch & 1][s.n + lane] = mytok | LZX_IN_RING; // at most 17 stream bytes, all inside the prepared region
        }
        s.n += cnt;
        q += p;
        continue;
      }
    }
    // ---- scalar path: one token with all checks (length bytes, block tail)
    if (s.n + 1 > LZX_B) s.publish(false, 0);
    s.own = q;
    if (!s.need(min(end, q + 32))) {
      rc = SB_EXTERNAL;
      break;
    }
    const uint32_t q0 = q;
    const uint32_t tok = s.ib(q), mlc = tok & 15u;
    uint32_t lit = tok >> 4, r = q + 1;
    // the ring keeps this sequence from q0 on while its extent is small (the workers will read it there)
    auto follow = [&]() { s.own = (r - q0 > LZX_EXTENT) ? r : q0; };
    // length-extension bytes at r
    auto ext = [&](uint32_t &len) -> int {
      uint32_t x = 255;
      while (x == 255) {
        if (r >= end) return SB_EXTERNAL;
        if ((r & 15) == 0 || r >= s.ready) {
          follow();
          if (!s.need(min(end, r + 16))) return SB_EXTERNAL;
        }
        x = s.ib(r++);
        len += x;
        if (len > SB_LZ4_MAXPOS) return SB_EXTERNAL;
      }
      return 0;
    };
    if (lit == 15 && (rc = ext(lit)) != 0) break;
    if (lit > end - r) {
      rc = SB_EXTERNAL;
      break;
    }
    if (r + lit == end) { // last sequence: literals only
      const bool in_ring = r + lit - q0 <= LZX_EXTENT;
      if (in_ring && !s.need(end)) {
        rc = SB_EXTERNAL;
        break;
      }
      s.push(q0, LZ4_E_END, tok | (in_ring ? LZX_IN_RING : 0u));
      done = true;
      break;
    }
    r += lit;
    follow(); // a long literal run is read by the workers from L2, not from this ring
    if (end - r < 2) {
      rc = SB_EXTERNAL;
      break;
    }
    if (!s.need(min(end, r + 32))) {
      rc = SB_EXTERNAL;
      break;
    }
    r += 2;
    uint32_t ml = mlc;
    if (mlc == 15 && (rc = ext(ml)) != 0) break;
    s.push(q0, LZ4_E_SEQ, tok | ((r - q0 <= LZX_EXTENT) ? LZX_IN_RING : 0u));
    q = r;
    q_lim = min(q_lim, q); // force housekeeping when the scalar path ran past the prepared region
    if (q >= end) { // a block must end with a literal-only sequence
      rc = SB_EXTERNAL;
      break;
    }
  }
  if (!done && rc == 0) rc = SB_EXTERNAL; // the chain left the block
  // the last batch carries the verdict; then drain the EMPTY arrivals of every batch not waited for yet
  s.publish(true, uint32_t(rc));
  s.wait_consumed(s.batch);
}

// ---- workers ---------------------------------------------------------------------------------------------
struct Lz4xOut {
  Lz4xShared *sh;
  uint32_t out_b; // shared address of the ring
  uint8_t *dst;   // global output
  uint32_t fl;    // bytes [0, fl) written to dst (16-byte aligned while streaming)
  bool vec;
  // write ring bytes [fl, upto) behind to HBM (team call, 64 threads; caller synchronises the team before and after)
  __device__ __forceinline__ void flush_to(uint32_t upto, bool final, uint32_t wt) {
    constexpr uint32_t OM = LZX_RING - 1;
    const uint8_t *ring = sh->out;
    if (vec) {
      const uint32_t a = min(upto, (fl + 15) & ~15u);
      for (uint32_t pos = fl + wt; pos < a; pos += LZX_B) dst[pos] = ring[pos & OM];
      uint32_t f = a;
      const uint32_t vend = upto & ~15u;
      for (uint32_t pos = f + wt * 16; pos + 16 <= vend; pos += LZX_B * 16)
        *reinterpret_cast<uint4 *>(dst + pos) = *reinterpret_cast<const uint4 *>(ring + (pos & OM));
      if (vend > f) f = vend;
      fl = f;
    }
    if (final || !vec) {
      for (uint32_t pos = fl + wt; pos < upto; pos += LZX_B) dst[pos] = ring[pos & OM];
      if (fl < upto) fl = upto;
    }
  }
};

// n bytes ring -> ring (positions masked), 8 loads in flight before the first store; the ranges do not overlap
__device__ __forceinline__ void lzx_copy_ss(uint32_t s_base, uint32_t s_pos, uint32_t s_mask, uint32_t d_base, uint32_t d_pos, uint32_t d_mask,
                                            uint32_t n) {
  for (uint32_t t0 = 0; t0 < n; t0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
      if (t0 + j < n) v[j] = lds_u8(s_base + ((s_pos + t0 + j) & s_mask));
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
      if (t0 + j < n) sts_u8(d_base + ((d_pos + t0 + j) & d_mask), v[j]);
  }
}
// n bytes global -> ring, 8 loads in flight (one L2 round trip per 8 bytes)
__device__ __forceinline__ void lzx_copy_gs(const uint8_t *gsrc, uint32_t d_base, uint32_t d_pos, uint32_t d_mask, uint32_t n) {
  for (uint32_t t0 = 0; t0 < n; t0 += 8) {
    uint32_t v[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
      if (t0 + j < n) v[j] = __ldcg(gsrc + t0 + j);
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j)
      if (t0 + j < n) sts_u8(d_base + ((d_pos + t0 + j) & d_mask), v[j]);
  }
}

// Returns 0 or SB_EXTERNAL (uniform over the 64 worker threads).
__device__ __forceinline__ int lz4x_work(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t dlen, Lz4xShared *sh) {
  const uint32_t wt = threadIdx.x - 32, lane = threadIdx.x & 31, ww = wt >> 5; // worker thread / warp index
  constexpr uint32_t OM = LZX_RING - 1, IM = LZX_INR - 1;
  const uint32_t mis = uint32_t(uintptr_t(src) & 15);
  const uint8_t *g = src - mis; // stream positions are offsets from here
  const uint32_t end = mis + clen;
  Lz4xOut o;
  o.sh = sh;
  o.out_b = smem_u32(sh->out);
  o.dst = dst;
  o.fl = 0;
  o.vec = (uintptr_t(dst) & 15) == 0;
  const uint32_t out_b = o.out_b, in_b = smem_u32(sh->in);
  uint32_t op_base = 0;
  int rc = 0;
  bool finished = false; // END sequence seen and written
  auto team = [&]() { lzx_bar_sync(LZX_BAR_TEAM, LZX_B); };
  // one earlier output byte (position sp) while the batch that ends at `bend` is being written
  auto src_byte = [&](uint32_t sp, uint32_t bend) -> uint32_t {
    if (bend - sp <= LZX_RING) return lds_u8(out_b + (sp & OM));
    return __ldcg(dst + sp);
  };
  // team-cooperative literal run: stream bytes [ls, ls+lit) (L2) -> output position op, in pieces that fit the ring
  auto lit_coop = [&](uint32_t ls, uint32_t lit, uint32_t op) {
    for (uint32_t done = 0; done < lit;) {
      const uint32_t p = min(lit - done, 2048u);
      for (uint32_t i = wt; i < p; i += LZX_B) sts_u8(out_b + ((op + done + i) & OM), g[ls + done + i]);
      done += p;
      team();
      if (op + done - o.fl >= LZX_FLUSHQ) {
        o.flush_to((op + done) & ~15u, false, wt);
        team();
      }
    }
  };
  // team-cooperative match of any length / overlap at output position mpos (everything before mpos is final)
  auto match_coop = [&](uint32_t mpos, uint32_t offset, uint32_t ml) {
    for (uint32_t done = 0; done < ml;) {
      const uint32_t at = mpos + done;
      uint32_t p = min(ml - done, 2048u);
      if (offset < LZX_B) { // pattern replication: byte i = pattern[i % offset], pattern = the `offset` bytes before mpos
        const uint32_t pat = mpos - offset;
        for (uint32_t i = wt; i < p; i += LZX_B) {
          const uint32_t v = src_byte(pat + (done + i) % offset, at + p);
          sts_u8(out_b + ((at + i) & OM), v);
        }
      } else { // a piece only reads bytes written before it started
        p = min(p, offset);
        for (uint32_t i = wt; i < p; i += LZX_B) sts_u8(out_b + ((at + i) & OM), src_byte(at - offset + i, at + p));
      }
      done += p;
      team();
      if (at + p - o.fl >= LZX_FLUSHQ) {
        o.flush_to((at + p) & ~15u, false, wt);
        team();
      }
    }
  };

#ifdef SB_LZ4_PROF
  uint32_t prof[12] = {0};
  struct ProfDump {
    uint32_t *p;
    uint32_t wt;
    __device__ ~ProfDump() {
      if (wt == 0)
        for (int i = 0; i < 12; ++i) atomicAdd(&g_lz4_prof[i], (unsigned long long)p[i]);
    }
  } prof_dump{prof, wt};
#endif
  for (uint32_t batch = 0;; ++batch) {
    const uint32_t b = batch & 1;
    LZ4_T(t0);
    lzx_bar_sync(LZX_BAR_FULL0 + b, LZX_THREADS);
    LZ4_T(t1);
    LZ4_ACC(0, t0, t1);
    const uint32_t nb = sh->nb[b];
#ifdef SB_LZ4_PROF
    prof[10] += 1;
    prof[11] += nb;
#endif
    const bool last = sh->last[b] != 0;
    const uint32_t scan_rc = sh->rc[b];
    // ---- one sequence per thread: parse, from the scanner's ring when the entry says so, else from L2
    uint32_t kind = LZ4_E_SEQ, lit = 0, ml = 0, offset = 1, ls = 0;
    bool in_ring = false;
    const bool active = wt < nb && rc == 0 && !finished;
    if (active) {
      const uint32_t e = sh->ent[b][wt];
      const uint32_t q = e & SB_LZ4_MAXPOS;
      kind = e >> 30;
      const uint32_t tk = sh->tok[b][wt];
      in_ring = (tk & LZX_IN_RING) != 0;
      lit = (tk >> 4) & 15u;
      const uint32_t mlc = tk & 15u;
      uint32_t r = q + 1;
      auto rd = [&](uint32_t pos) -> uint32_t { return in_ring ? lds_u8(in_b + (pos & IM)) : uint32_t(g[pos]); };
      if (lit == 15) {
        uint32_t x;
        do {
          x = rd(r++);
          lit += x;
        } while (x == 255 && r < end);
      }
      ls = r;
      r += lit;
      if (kind == LZ4_E_SEQ) { // the scanner checked that the offset and the length bytes lie inside the block
        offset = rd(r) | (rd(r + 1) << 8);
        r += 2;
        ml = mlc;
        if (mlc == 15) {
          uint32_t x;
          do {
            x = rd(r++);
            ml += x;
          } while (x == 255 && r < end);
        }
        ml += 4;
      }
    }
    LZ4_T(t2);
    LZ4_ACC(1, t1, t2);
    // The batch buffer and its stream bytes go back to the scanner (EMPTY) once the short literal runs have left
    // the input ring: in the fast path below, else right after this block.
    bool released = false;
    auto release = [&]() {
      __syncwarp();
      lzx_bar_arrive(LZX_BAR_EMPTY0 + b, LZX_THREADS);
      released = true;
    };

    if (nb && rc == 0 && !finished) {
      // ---- output positions: saturating scan of the lengths over the 64 threads (cap: dlen + 1 = out of bounds)
      const uint32_t cap = dlen + 1;
      const uint32_t len = active ? sat_add(min(lit, cap), min(ml, cap), cap) : 0u;
      const uint32_t incl = warp_incl_scan_sat(len, cap);
      uint32_t excl_w = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl_w = 0;
      if (lane == 31) sh->wsum[ww] = incl;
      if (wt == 0) sh->bad = 0;
      team();
      const uint32_t w0 = sh->wsum[0], btotal = sat_add(w0, sh->wsum[1], cap);
      const uint32_t excl = sat_add(ww ? w0 : 0u, excl_w, cap);
      const uint32_t op = sat_add(op_base, excl, cap), mpos = sat_add(op, lit, cap);
      const uint32_t bend = sat_add(op_base, btotal, cap);
      // LZ4_decompress_safe: offset 0 / before the start of the output, output overrun
      bool bad = active && (mpos > dlen || (kind == LZ4_E_SEQ && (offset == 0 || offset > mpos || ml > dlen - mpos)));
      // the END entry closes the block: it must be the last entry and end exactly at dlen
      if (active && kind == LZ4_E_END && (wt + 1 != nb || !last || mpos != dlen)) bad = true;
      if (active && kind == LZ4_E_ERROR) bad = true;
      if (bad) sh->bad = 1;
      const bool own_lit = active && lit <= LZX_PER_THREAD;
      const bool is_long_lit = active && !own_lit, is_long_m = active && kind == LZ4_E_SEQ && ml > LZX_PER_THREAD;
      // a match that reads bytes this batch writes must wait for them: handled in order after the others
      // (an overlapping match, offset < ml, replicates a pattern into its own output: same ordered path)
      const bool dep = active && kind == LZ4_E_SEQ && !bad && ((mpos - offset) + min(ml, offset) > op_base || offset < ml);
      const uint32_t m_dep = __ballot_sync(0xffffffffu, dep), m_ll = __ballot_sync(0xffffffffu, is_long_lit && !bad);
      const uint32_t m_lm = __ballot_sync(0xffffffffu, is_long_m && !dep && !bad);
      if (lane == 0) {
        sh->dep_mask[ww] = m_dep;
        sh->long_lit[ww] = m_ll;
        sh->long_m[ww] = m_lm;
      }
      if (active) {
        sh->p_mpos[wt] = mpos;
        sh->p_off[wt] = offset;
        sh->p_ml[wt] = ml;
        sh->p_ls[wt] = ls;
        sh->p_lit[wt] = lit;
      }
      team();
      LZ4_T(t3);
      LZ4_ACC(2, t2, t3);
      if (sh->bad) {
        rc = SB_EXTERNAL;
      } else if (bend > dlen) {
        rc = SB_EXTERNAL;
      } else if (btotal <= LZX_FAST_MAX) {
        // ---- fast path.  Ring room: everything from fl to the end of the batch must fit.
        if (bend - o.fl > LZX_RING - 64) {
          o.flush_to(op_base & ~15u, false, wt);
          team();
        }
        LZ4_T(t4);
        LZ4_ACC(7, t3, t4);
        // matches whose source precedes the batch, short ones by their thread.  Sources behind the ring come from L2:
        // those loads are issued first and land under the literal copies.
        const bool own_m = active && kind == LZ4_E_SEQ && !dep && ml <= LZX_PER_THREAD;
        const uint32_t sp = mpos - offset;
        const bool m_far = own_m && bend - (sp + ml - 1) > LZX_RING, m_near = own_m && bend - sp <= LZX_RING;
        uint32_t far_v[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) far_v[j] = (m_far && j < ml) ? uint32_t(__ldcg(dst + sp + j)) : 0u;
        // literals: short runs by their thread (from the input ring for ring sequences, L2 otherwise), long runs by the team
        if (own_lit) {
          if (in_ring) lzx_copy_ss(in_b, ls, IM, out_b, op, OM, lit);
          else lzx_copy_gs(g + ls, out_b, op, OM, lit);
        }
        release();
        if (m_far) {
#pragma unroll
          for (uint32_t j = 0; j < 8; ++j)
            if (j < ml) sts_u8(out_b + ((mpos + j) & OM), far_v[j]);
          if (ml > 8) lzx_copy_gs(dst + sp + 8, out_b, mpos + 8, OM, ml - 8);
        } else if (m_near) {
          lzx_copy_ss(out_b, sp, OM, out_b, mpos, OM, ml); // independent: source and destination do not overlap
        } else if (own_m) { // straddles the ring boundary
          for (uint32_t t = 0; t < ml; ++t) sts_u8(out_b + ((mpos + t) & OM), src_byte(sp + t, bend));
        }
        LZ4_T(t5);
        LZ4_ACC(3, t4, t5);
        for (uint32_t w = 0; w < 2; ++w) { // long literal runs / long independent matches: the team, one at a time
          uint32_t m = sh->long_lit[w];
          while (m) {
            const uint32_t k = w * 32 + uint32_t(__ffs(int(m))) - 1u;
            m &= m - 1;
            const uint32_t l_ = sh->p_lit[k], s_ = sh->p_ls[k], o_ = sh->p_mpos[k] - l_;
            for (uint32_t i = wt; i < l_; i += LZX_B) sts_u8(out_b + ((o_ + i) & OM), g[s_ + i]);
          }
          m = sh->long_m[w];
          while (m) {
            const uint32_t k = w * 32 + uint32_t(__ffs(int(m))) - 1u;
            m &= m - 1;
            const uint32_t mp = sh->p_mpos[k], of = sh->p_off[k], l_ = sh->p_ml[k];
            // independent: the whole source precedes the batch, so no overlap with the destination
            for (uint32_t i = wt; i < l_; i += LZX_B) sts_u8(out_b + ((mp + i) & OM), src_byte(mp - of + i, bend));
          }
        }
        team();
        LZ4_T(t6);
        LZ4_ACC(4, t5, t6);
#ifdef SB_LZ4_PROF
        prof[9] += __popc(sh->dep_mask[0]) + __popc(sh->dep_mask[1]);
#endif
        // Matches that read this batch's own output.  With liblz4-written blocks most of them only need literals or
        // independent matches (final by now), so they resolve in parallel: a bitmap of the bytes still owed by a
        // pending match tells every thread whether its source is final; ready ones copy, clear their bits, repeat.
        // What is left after a few rounds (true chains, long dependent matches) runs in stream order on one warp.
        if (sh->dep_mask[0] | sh->dep_mask[1]) {
          const uint32_t nw = ((btotal + 31) >> 5) + 1;
          for (uint32_t i = wt; i < nw; i += LZX_B) sh->pend[i] = 0;
          team();
          const uint32_t d0 = mpos - op_base;
          if (dep) {
            for (uint32_t t = 0; t < ml;) { // set bits [d0, d0 + ml)
              const uint32_t bit = d0 + t, k = min(32u - (bit & 31u), ml - t);
              atomicOr(&sh->pend[bit >> 5], (k == 32 ? 0xffffffffu : ((1u << k) - 1u)) << (bit & 31u));
              t += k;
            }
          }
          bool mine = dep && ml <= LZX_PER_THREAD, owed = dep;
          team();
          const uint32_t need = min(ml, offset); // source bytes that must be final (an overlapping match only needs its pattern)
#pragma unroll 1
          for (uint32_t round = 0; round < 4; ++round) {
            bool ready = false;
            if (mine) {
              const uint32_t s_end = sp + need; // > op_base for every dependent match except a pure overlap at the batch start
              if (s_end <= op_base) {
                ready = true;
              } else {
                const uint32_t s_lo = sp > op_base ? sp - op_base : 0u, len = s_end - op_base - s_lo; // <= 32
                const uint32_t a_ = s_lo >> 5, sh0 = s_lo & 31u;
                const uint64_t w64 = uint64_t(sh->pend[a_]) | (uint64_t(sh->pend[a_ + 1]) << 32);
                ready = ((w64 >> sh0) & ((1ull << len) - 1ull)) == 0;
              }
            }
            const uint32_t rb = __ballot_sync(0xffffffffu, ready);
            if (lane == 0) sh->ready_w[ww] = rb;
            if (ready) {
              if (offset >= ml) {
                for (uint32_t t = 0; t < ml; ++t) sts_u8(out_b + ((mpos + t) & OM), src_byte(sp + t, bend));
              } else {
                for (uint32_t t = 0; t < ml; ++t) sts_u8(out_b + ((mpos + t) & OM), src_byte(sp + t % offset, bend));
              }
            }
            team(); // the copies of this round are in the ring
            if (ready) {
              for (uint32_t t = 0; t < ml;) {
                const uint32_t bit = d0 + t, k = min(32u - (bit & 31u), ml - t);
                atomicAnd(&sh->pend[bit >> 5], ~((k == 32 ? 0xffffffffu : ((1u << k) - 1u)) << (bit & 31u)));
                t += k;
              }
              mine = false;
              owed = false;
            }
            const bool more = (sh->ready_w[0] | sh->ready_w[1]) != 0;
            team(); // bits cleared, flags read
            if (!more) break;
          }
          const uint32_t lb = __ballot_sync(0xffffffffu, owed);
          if (lane == 0) sh->left_w[ww] = lb;
          team();
          if ((sh->left_w[0] | sh->left_w[1]) != 0) {
            if (ww == 0) {
              for (uint32_t w = 0; w < 2; ++w) {
                uint32_t m = sh->left_w[w];
                while (m) {
                  const uint32_t k = w * 32 + uint32_t(__ffs(int(m))) - 1u;
                  m &= m - 1;
                  const uint32_t mp = sh->p_mpos[k], of = sh->p_off[k], l_ = sh->p_ml[k];
                  const uint32_t sp_ = mp - of;
                  if (of >= 32 || of >= l_) { // a 32-byte step never reads what it writes
                    for (uint32_t i = 0; i < l_; i += 32) {
                      if (i + lane < l_) sts_u8(out_b + ((mp + i + lane) & OM), src_byte(sp_ + i + lane, bend));
                      __syncwarp();
                    }
                  } else { // pattern replication: byte i = pattern[i % offset]
                    for (uint32_t i = lane; i < l_; i += 32) {
                      const uint32_t v = src_byte(sp_ + i % of, bend);
                      sts_u8(out_b + ((mp + i) & OM), v);
                    }
                    __syncwarp();
                  }
                }
              }
            }
            team();
          }
        }
        LZ4_T(t7);
        LZ4_ACC(6, t6, t7);
        op_base = bend;
      } else {
        // ---- slow path (long runs: highly compressible or incompressible data): one sequence at a time, the team
        for (uint32_t k = 0; k < nb; ++k) {
          const uint32_t mp = sh->p_mpos[k], of = sh->p_off[k], l_ = sh->p_ml[k], li = sh->p_lit[k], s_ = sh->p_ls[k];
          lit_coop(s_, li, mp - li);
          if (l_) match_coop(mp, of, l_);
        }
        op_base = bend;
      }
      if (rc == 0) {
        LZ4_T(t8);
        if (last && scan_rc == 0) finished = op_base == dlen; // the END entry was checked to end exactly at dlen
        if (op_base - o.fl >= LZX_FLUSHQ && !finished) {
          team();
          o.flush_to(op_base & ~15u, false, wt);
        }
        LZ4_T(t9);
        LZ4_ACC(5, t8, t9);
      }
    }
    if (!released) release();
    if (last) {
      if (rc == 0 && scan_rc != 0) rc = int(scan_rc);
      if (rc == 0 && !finished) rc = SB_EXTERNAL; // the stream ended before the output was complete
      break;
    }
  }
  team();
  if (rc == 0) o.flush_to(dlen, true, wt);
  return rc;
}

} // namespace sb
