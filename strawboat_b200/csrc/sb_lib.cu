// sb_lib.cu -- kernels + C ABI of libstrawboat_b200.so (see include/strawboat_b200.h).
//
// Decode pipeline of one sb_decode_columns call:
//   host : page table (PageDesc per page, WorkItem per CTA job) -> pinned -> H2D
//   D0/1 : sb_size_kernel   binary columns only: value bytes per page      (two-pass sizing)
//          sb_scan_kernel   per-column exclusive scan of those sizes
//   D*   : sb_decode_kernel persistent grid, one CTA (128 thr) per page at a time: TMA bulk
//          load of the page into shared memory, codec dispatch, coalesced 16-byte stores
// There is no CPU fallback anywhere: without a CUDA device every entry point returns SB_CUDA.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "sb_common.cuh"
#include "sb_decode.cuh"
#include "sb_binary.cuh"
#include "sb_nested.cuh"

namespace sb {

__device__ __forceinline__ bool is_fixed_type(int t) { return (t >= SB_I8 && t <= SB_F64) || t == SB_I128 || t == SB_I256; }

constexpr uint32_t kLightSmem = 26 * 1024;   // light kernel: 8 CTAs / SM
constexpr uint32_t kSmemMax = 74 * 1024;     // dynamic shared memory per CTA: 3 CTAs / SM
constexpr uint32_t kSmemMin = 40 * 1024;
#ifndef SB_NEED_EXTRA
#define SB_NEED_EXTRA (7 * 1024)
#endif
// arena wanted next to a compact page and its index buffer (run-start arrays, scan workspaces).  7 KiB: a bit-packed
// i32 page of 16 KiB + the 32 KiB index allowance of 8192 rows stay at 55 KiB = 4 CTAs per SM (16 KiB made config 2 run
// at 3: 1256 -> 1202 us, tools/ab_cols.py)
constexpr uint32_t kNeedExtra = SB_NEED_EXTRA;
constexpr uint32_t kArenaMin = 6 * 1024;     // arena left after the largest staged page
constexpr uint32_t kTileBytes = 64 * 1024;   // output bytes per work item of an unstaged page
constexpr uint32_t kTmaChunk = 32 * 1024;
// plain value bytes per work item of an unstaged binary page.  Measured on plain utf8 pages (16 x 1 M rows, ~78 KB of
// value bytes per page; tools/ab_utf8.py): 16 KiB 167 us, 32 KiB 118, 64 KiB 108, 128 KiB 94, 256 KiB 95 -- an item costs
// ~4 us of dependent latency (ticket, descriptors, first TMA round trip) whatever its size.
#ifndef SB_BIN_TILE
#define SB_BIN_TILE (128 * 1024)
#endif
constexpr uint32_t kBinTile = SB_BIN_TILE;



// ------------------------------------------------------------------------------------
// LZ4 side path.  Top-level LZ4 value blocks of fixed-width columns are inherently serial
// per page, so they run in their own kernel (one WARP per page, shared-memory ring, many
// pages in flight per SM) concurrently with the main kernel instead of pinning a whole CTA.
// sb_classify_kernel finds those pages; both kernels apply the same predicate.
// ------------------------------------------------------------------------------------
struct Lz4Job {
  const uint8_t *src;
  uint8_t *dst;
  uint32_t clen, dlen, page, pad;
};

// Parses [validity section][hdr9] of a flat fixed-width page straight from global memory.
// Returns true when the value block is a top-level LZ4 block with an in-bounds payload.
__device__ __forceinline__ bool lz4_side_page(const uint8_t *p, uint32_t len, bool nullable, uint32_t *vb_out,
                                              uint32_t *clen_out) {
  uint32_t vb = 0;
  if (nullable) {
    if (len < 4) return false;
    uint32_t L = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
    if (L > len - 4) return false;
    vb = 4 + L;
  }
  if (len - vb < 9 || p[vb] != SB_C_LZ4) return false;
  const uint8_t *h = p + vb;
  uint32_t clen = uint32_t(h[1]) | (uint32_t(h[2]) << 8) | (uint32_t(h[3]) << 16) | (uint32_t(h[4]) << 24);
  if (clen > len - vb - 9) return false;
  *vb_out = vb;
  *clen_out = clen;
  return true;
}

// [validity section][hdr9 codec None][n * W value bytes]: a page that is one plain copy
__device__ __forceinline__ bool plain_page(const uint8_t *p, uint32_t len, bool nullable, uint64_t out_bytes) {
  uint32_t vb = 0;
  if (nullable) {
    if (len < 4) return false;
    uint32_t L = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
    if (L > len - 4) return false;
    vb = 4 + L;
  }
  if (len - vb < 9 || p[vb] != SB_C_NONE) return false;
  const uint8_t *h = p + vb;
  uint32_t clen = uint32_t(h[1]) | (uint32_t(h[2]) << 8) | (uint32_t(h[3]) << 16) | (uint32_t(h[4]) << 24);
  return clen <= len - vb - 9 && uint64_t(clen) == out_bytes && out_bytes >= 4 * SB_RING_CHUNK;
}

// side_flags: 0 = decoded by the main kernel, 3 = plain page (codec None): streamed through the TMA ring
// instead of being staged whole, 1 = top-level LZ4 block -> sb_lz4_kernel,
// 2 = "stored" LZ4 block (one literal run covering the whole output: what LZ4 emits for
// incompressible data) -> plain copy in the main kernel.
// Jobs are binned by compressed size: long streams from the front of `jobs`, short ones from
// the back, so the (latency-bound) long pages start first and the short ones fill the tail.
// One WARP per page (the length bytes of a stored block are checked 32 at a time).
__device__ __forceinline__ bool lz4_stored_block(const uint8_t *s, uint32_t clen, uint32_t dlen) {
  const uint32_t lane = threadIdx.x & 31;
  if (dlen < 15 || clen < 2 || s[0] != 0xF0) return false;
  uint32_t ne = (dlen - 15) / 255 + 1; // length-extension bytes: (ne-1) x 255, then the rest
  if (ne > 4096 || uint64_t(clen) != 1ull + ne + dlen) return false;
  bool ok = s[ne] == (dlen - 15) % 255;
  for (uint32_t i = 1 + lane; i < ne; i += 32) ok &= s[i] == 255;
  return __all_sync(0xffffffffu, ok);
}

__device__ __forceinline__ uint32_t ldg_u32u(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

constexpr uint8_t SB_FLAG_HEAVY = 0x80; // side_flags bit: the page belongs to the full decode kernel
// codec trees the light decode kernel handles (flat fixed-width page; `h` = hdr9 of the value block)
__device__ __forceinline__ bool light_codec_tree(const uint8_t *h, uint32_t avail) {
  if (avail < 9) return false;
  const uint32_t c = h[0];
  if (c == SB_C_NONE || c == SB_C_ONEVALUE || c == SB_C_RLE || c == SB_C_BITPACK || c == SB_C_DELTABP) return true;
  if (c == SB_C_DICT && avail >= 18) { // the index sub-page
    const uint32_t s = h[9];
    return s == SB_C_NONE || s == SB_C_ONEVALUE || s == SB_C_RLE || s == SB_C_BITPACK;
  }
  return false;
}

// Returns the page's flags: low bits = side (0 decode kernel, 1 sb_lz4_kernel, 2 stored LZ4 block, 3 plain page),
// SB_FLAG_HEAVY = not for the light decode kernel.  One warp per page (uniform control flow: every lane loads the
// same bytes); the family test only runs for pages that are neither plain nor LZ4.
__device__ __forceinline__ uint8_t classify_page(const PageDesc &pg, const ColDesc &col, uint32_t i, uint32_t n_pages, Lz4Job *jobs, uint32_t *n_jobs,
                                                 const PageAux *__restrict__ aux, uint32_t light_cap, uint32_t lane) {
  const bool nested = col.n_nested > 1;
  if (!is_fixed_type(col.type)) return SB_FLAG_HEAVY; // binary, boolean: the full kernel
  uint32_t vb, clen;
  uint64_t n_vals = pg.num_values;
  const uint8_t *body = pg.src;
  uint32_t body_len = pg.len;
  if (nested) {
    // [u32 rows][u32 rep_len][u32 def_len][rep][def][VALUE_BLOCK over the leaf slots]: the plan pass counted the slots
    if (pg.len < 12 || pg.aux == 0xffffffffu) return SB_FLAG_HEAVY;
    const uint64_t lv = 12ull + ldg_u32u(pg.src + 4) + ldg_u32u(pg.src + 8);
    if (lv > pg.len) return SB_FLAG_HEAVY;
    body += lv;
    body_len -= uint32_t(lv);
    n_vals = aux[pg.aux].cnt[col.n_nested - 1];
  } else if (plain_page(pg.src, pg.len, col.nullable != 0, uint64_t(pg.num_values) * uint32_t(col.W))) {
    // plain pages stream through the TMA ring whatever their size: light unless the validity section needs a large stage
    const uint32_t stage = pg.len - pg.num_values * uint32_t(col.W);
    return 3 | ((!col.nullable || stage + 32 <= light_cap) ? 0 : SB_FLAG_HEAVY);
  }
  if (!lz4_side_page(body, body_len, !nested && col.nullable != 0, &vb, &clen)) {
    if (nested) return SB_FLAG_HEAVY;
    // flat fixed-width page: family by codec tree and staging size (pages the light kernel cannot stage stay heavy)
    uint32_t vb0 = 0;
    if (col.nullable) {
      if (pg.len < 4) return SB_FLAG_HEAVY;
      const uint32_t L = ldg_u32u(pg.src);
      if (L > pg.len - 4) return SB_FLAG_HEAVY;
      vb0 = 4 + L;
    }
    return (pg.len + 32 <= light_cap && light_codec_tree(pg.src + vb0, pg.len - vb0)) ? 0 : SB_FLAG_HEAVY;
  }
  const uint64_t dlen64 = n_vals * uint32_t(col.W);
  if (dlen64 > SB_LZ4_MAXPOS / 2 || clen > SB_LZ4_MAXPOS / 2) return SB_FLAG_HEAVY; // positions are 30-bit in sb_lz4_kernel
  const uint32_t dlen = uint32_t(dlen64);
  // an LZ4 page leaves the validity section (if any) to the decode kernel: light when that stages small
  const uint8_t lz_fam = nested ? SB_FLAG_HEAVY : ((!col.nullable || vb + 32 <= light_cap) ? 0 : SB_FLAG_HEAVY);
  if (lz4_stored_block(body + vb + 9, clen, dlen)) return 2 | (nested || pg.len + 32 > light_cap ? SB_FLAG_HEAVY : lz_fam);
  if (lane == 0) {
    Lz4Job j;
    j.src = body + vb + 9;
    j.dst = col.values + pg.out_elem * uint64_t(col.W);
    j.clen = clen;
    j.dlen = dlen;
    j.page = i;
    j.pad = 0;
    // n_jobs[0] = long jobs (front), n_jobs[1] = short jobs (back)
    const bool big = clen >= 8192;
    const uint32_t slot = big ? atomicAdd(n_jobs, 1u) : n_pages - 1 - atomicAdd(n_jobs + 1, 1u);
    jobs[slot] = j;
  }
  return 1 | lz_fam;
}

// ------------------------------------------------------------------------------------
// Entry walk of binary Dict pages, ahead of the plan pass.  The `[u64 len][bytes]` entries of a dictionary are one
// serial chain (binary/dict.rs:102-120): ~1000 dependent steps per page of configs[2], on one thread, which made the
// plan pass rounds x the latency of one page at 3 CTAs per SM.  Here one WARP takes one page (12 per SM), stages the
// dictionary region in shared memory and walks it, so the chains of a whole column run side by side.  Purely an
// accelerator: it records (start, k, end, payload total) in the page's PageAux and fills the page's BinEntry slice only
// when the whole chain parsed cleanly; binary_page_size uses the record only if it describes the dictionary it finds
// itself, and walks (and reports errors) as before otherwise.
// ------------------------------------------------------------------------------------
constexpr uint32_t kWalkWin = 12 * 1024;                               // bytes of the page staged per window
constexpr uint32_t kWalkSmem = kWalkWin + 48 + (kWalkWin / 8 + 1) * 4 + 12; // + the header offsets of its entries
__device__ __forceinline__ uint32_t rd32_bytes(const uint8_t *p) {
  return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
}
__global__ void __launch_bounds__(32)
    sb_dict_walk_kernel(const PageDesc *__restrict__ pages, const ColDesc *__restrict__ cols, const WorkItem *__restrict__ items,
                        uint32_t n_items, PageAux *aux, BinEntry *entries) {
  extern __shared__ __align__(16) uint8_t walk_win[]; // kWalkSmem
  const uint32_t lane = threadIdx.x;
  for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
    const PageDesc pg = pages[items[it].page];
    const ColDesc &col = cols[pg.col];
    if (!(col.type == SB_BINARY || col.type == SB_LARGE_BINARY) || col.n_nested > 1 || pg.aux == 0xffffffffu) continue;
    const uint8_t *p = pg.src;
    const uint32_t L = pg.len;
    uint32_t vb = 0;
    if (col.nullable) { // validity section: [u32 L][...]
      if (L < 4) continue;
      const uint32_t vl = rd32_bytes(p);
      if (vl > L - 4) continue;
      vb = 4 + vl;
    }
    if (L - vb < 9 + 9 + 4 || p[vb] != SB_C_DICT) continue;
    if (rd32_bytes(p + vb + 1) > L - vb - 9) continue;
    const uint32_t body = vb + 9; // index sub-page: hdr9 + its compressed bytes, then [u32 k][entries]
    const uint64_t used = 9 + uint64_t(rd32_bytes(p + body + 1));
    if (used + 4 > uint64_t(L - body)) continue;
    const uint32_t k = rd32_bytes(p + body + uint32_t(used));
    const uint32_t start = body + uint32_t(used) + 4;
    if (k < 32 || uint64_t(k) * 8 > uint64_t(L - start)) continue;
    BinEntry *tab = entries + pg.tab_off;
    uint32_t pos = start, e = 0;
    bool good = true;
    uint32_t *s_off = reinterpret_cast<uint32_t *>(walk_win + kWalkWin + 48); // header offsets of the window's entries (+ end)
    while (e < k && good) {
      // window [pos, pos + wl) of the page, staged at the same 16-byte phase as in global memory
      const uint32_t wl = min(kWalkWin, L - pos);
      const uint8_t *g = p + pos;
      const uint32_t mis = uint32_t(uintptr_t(g) & 15);
      const uint32_t nvec = (mis + wl + 15) >> 4;
      const uint4 *gv = reinterpret_cast<const uint4 *>(g - mis);
      uint4 *sv = reinterpret_cast<uint4 *>(walk_win);
      for (uint32_t v0 = lane; v0 < nvec; v0 += 32 * 8) { // eight independent 16-byte loads per lane in flight
        uint4 r[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
          if (v0 + 32 * j < nvec) r[j] = gv[v0 + 32 * j];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j)
          if (v0 + 32 * j < nvec) sv[v0 + 32 * j] = r[j];
      }
      __syncwarp();
      // lane 0 follows the chain through the window and only notes where each header starts (the loop is one
      // dependent chain on one lane: every instruction in it costs its full latency); the records are written
      // by all lanes afterwards.  An entry takes at least 8 bytes, so a window holds at most kWalkWin / 8 of them.
      uint32_t off = 0, c = 0; // off = position - window start, c = entries found in this window
      if (lane == 0) {
        const uint32_t cmax = k - e, room = L - pos - 8; // lo <= room - off  <=>  the payload ends inside the page
        // one loop-carried chain (off -> LDS -> funnel shift -> off) and ONE branch per entry: the exit reasons are
        // folded into `stop` with selects (a taken-or-not branch per check cost more than the loads)
        uint32_t stop = 0; // 1 = fewer than 8 bytes of window left, 2 = not a length that fits the page
        do {
          const uint32_t so = mis + off;
          const uint32_t *q = reinterpret_cast<const uint32_t *>(walk_win + (so & ~3u)); // may run into the 48 spare bytes
          const uint32_t sh = (so & 3) * 8;
          const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
          const uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
          const bool short_hdr = wl - off < 8;
          const bool bad = hi != 0 || lo > room - off;
          stop = short_hdr ? 1u : (bad ? 2u : 0u);
          s_off[c] = off;
          off += stop ? 0u : 8 + lo;
          c += stop ? 0u : 1u;
        } while (stop == 0 && c != cmax && off < wl);
        // stop == 1 with entries found: the header straddles the window end, stage again from here;
        // without any: the page ends inside a header -- the plan pass reports it
        if (stop == 2 || (stop == 1 && c == 0)) good = false;
        s_off[c] = off;
      }
      off = __shfl_sync(0xffffffffu, off, 0);
      c = __shfl_sync(0xffffffffu, c, 0);
      good = __shfl_sync(0xffffffffu, int(good), 0) != 0;
      __syncwarp(); // s_off[] written by lane 0 is read by every lane
      if (good) {
        for (uint32_t i = lane; i < c; i += 32) {
          const uint32_t o = s_off[i];
          tab[e + i] = BinEntry{pos + o + 8, s_off[i + 1] - o - 8};
        }
        e += c;
        pos += off;
      }
      __syncwarp();
    }
    if (good && e == k && lane == 0) {
      PageAux &ax = aux[pg.aux];
      ax.cnt[0] = start;
      ax.cnt[1] = k;
      ax.cnt[2] = pos;
      ax.base[0] = uint64_t(pos - start) - 8ull * k; // payload bytes = the chain's span minus its headers
      ax.pad = 1;
    }
    __syncwarp();
  }
}

__global__ void sb_classify_kernel(const PageDesc *__restrict__ pages, const ColDesc *__restrict__ cols, uint32_t n_pages,
                                   Lz4Job *jobs, uint32_t *n_jobs, uint8_t *side_flags, const PageAux *__restrict__ aux, uint32_t light_cap,
                                   uint32_t *n_heavy) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n_pages) return;
  const PageDesc pg = pages[i];
  const ColDesc col = cols[pg.col];
  const uint8_t f = classify_page(pg, col, i, n_pages, jobs, n_jobs, aux, light_cap, lane);
  if (lane == 0) {
    if (f) side_flags[i] = f; // the array was zeroed with the tables
    if (n_heavy && (f & SB_FLAG_HEAVY)) atomicAdd(n_heavy, 1u);
  }
}

// One CTA = SB_LZ4_PAIRS scanner warps (warps 0..PAIRS-1, one warpgroup) + as many mover warps (the next
// warpgroup); pair p = warps p and PAIRS + p works on one block at a time.  The mover needs ~120 registers
// to keep its per-sequence state out of local memory (spills on its critical path cost ~35 % of the block
// time), the scanner a third of that: the two warpgroups re-split the CTA's register file with setmaxnreg,
// so 12 blocks stay resident per SM (3 CTAs x 4 pairs) instead of the 8 a uniform 124-register kernel allows.
constexpr uint32_t SB_LZ4_PAIRS = 4;
constexpr uint32_t SB_LZ4_CTA = SB_LZ4_PAIRS * 64;
__device__ __forceinline__ void pair_sync(uint32_t pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

constexpr uint32_t SB_LZ4_SMEM = uint32_t(sizeof(Lz4Shared)) * SB_LZ4_PAIRS + 64;

// job loop of one warp of a pair (ROLE 0 = scanner, 1 = mover).  The two roles are separate code paths
// from the setmaxnreg on, so ptxas allocates each under its own register limit.
template <int ROLE>
__device__ __forceinline__ void lz4_pair_loop(const Lz4Job *__restrict__ jobs, uint32_t n_big, uint32_t n_small, uint32_t n_pages,
                                              uint32_t *counter, int32_t *status, unsigned long long *bytes_done, Lz4Shared &sh,
                                              uint32_t *s_job, uint32_t pair) {
  const uint32_t lane = threadIdx.x & 31;
  // First job of a pair = its global pair index: the block scheduler deals consecutive CTAs to different
  // SMs, so the resident blocks spread evenly over the SMs; later jobs come from the ticket counter.
  bool first = true;
  for (;;) {
    if (ROLE == 0 && lane == 0) {
      *s_job = first ? blockIdx.x * SB_LZ4_PAIRS + pair : gridDim.x * SB_LZ4_PAIRS + atomicAdd(counter, 1u);
      sh.produced = 0;
      sh.in_ready = 0;
      sh.consumed = 0;
      sh.m_q = 0;
      sh.abort = 0;
    }
    pair_sync(pair);
    const uint32_t j = *s_job;
    if (j >= n_big + n_small) break;
    const Lz4Job job = jobs[j < n_big ? j : n_pages - 1 - (j - n_big)];
    int rc = 0;
    if (job.clen == 0 || job.dlen == 0) {
      // an empty block decodes to nothing; LZ4_decompress_safe rejects everything else here
      if (!(job.dlen == 0 && job.clen == 1 && job.src[0] == 0) && !(job.clen == 0 && job.dlen == 0)) rc = SB_EXTERNAL;
    } else if (ROLE == 0) {
      rc = lz4_scan(job.src, job.clen, &sh);
    } else {
      rc = lz4_move(job.dst, job.dlen, uint32_t(uintptr_t(job.src) & 15) + job.clen, &sh);
    }
    if (rc && lane == 0) atomicCAS(status + job.page, 0, rc);
    if (ROLE == 0 && lane == 0) atomicAdd(bytes_done, (unsigned long long)(job.clen) + job.dlen);
    first = false;
    pair_sync(pair); // both warps are done with the shared state before it is reset
  }
}

__global__ void __launch_bounds__(SB_LZ4_CTA, 3)
    sb_lz4_kernel(const Lz4Job *__restrict__ jobs, const uint32_t *__restrict__ n_jobs_p, uint32_t n_pages, uint32_t *counter,
                  int32_t *status, unsigned long long *bytes_done) {
  extern __shared__ __align__(16) uint8_t lz4_dsm[];
  __shared__ uint32_t s_jobs[SB_LZ4_PAIRS];
  Lz4Shared *shs = reinterpret_cast<Lz4Shared *>(lz4_dsm);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t pair = warp % SB_LZ4_PAIRS;
  const uint32_t n_big = n_jobs_p[0], n_small = n_jobs_p[1];
  if (warp < SB_LZ4_PAIRS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    lz4_pair_loop<0>(jobs, n_big, n_small, n_pages, counter, status, bytes_done, shs[pair], &s_jobs[pair], pair);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    lz4_pair_loop<1>(jobs, n_big, n_small, n_pages, counter, status, bytes_done, shs[pair], &s_jobs[pair], pair);
  }
}

// ------------------------------------------------------------------------------------
// main decode kernel.  pass 0 = plan pass over the pages whose output size is data dependent
// (binary value bytes, nested per-depth entry counts); pass 1 = decode.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_binary_type(int t) { return t == SB_BINARY || t == SB_LARGE_BINARY; }

// Two instantiations share this body.  LIGHT = flat fixed-width pages whose codec tree is None / OneValue / RLE /
// Bitpacking / DeltaBitpacking / Dict over those (sb_classify_kernel sets SB_FLAG_HEAVY on everything else): no LZ4
// scanner / mover, no Freq, Patas, binary, boolean or nested code, so it compiles to half the registers and runs at
// 8 CTAs per SM -- the dependent chain of a small page (ticket -> descriptors -> TMA -> header -> gather -> store)
// is hidden by twice as many pages in flight.  Both kernels walk the same item list; each takes the items of its
// family (thread 0 skips the others while drawing tickets).
template <bool LIGHT>
__device__ __forceinline__ void decode_body(const PageDesc *__restrict__ pages, const ColDesc *__restrict__ cols,
                     const WorkItem *__restrict__ items, uint32_t n_items, uint32_t *counter, uint8_t *scratch,
                     uint64_t scratch_per_cta, int32_t *status, uint32_t stage_cap, uint32_t smem_bytes,
                     const uint8_t *__restrict__ side_flags, PageAux *aux, BinEntry *entries, uint32_t *codec_hist, int pass,
                     bool split) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ __align__(8) uint64_t s_rbar[SB_RING_STAGES];
  __shared__ int s_err;
  __shared__ uint32_t s_ws[SB_NWARP + 1];
  __shared__ int s_bcast[4];
  __shared__ uint32_t s_item;

  const uint32_t tid = threadIdx.x;
  // next item of this kernel's family (thread 0)
  auto draw = [&]() -> uint32_t {
    for (;;) {
      const uint32_t t = atomicAdd(counter, 1u);
      if (t >= n_items || !split) return t;
      const bool heavy = (side_flags[items[t].page] & SB_FLAG_HEAVY) != 0;
      if (heavy != LIGHT) return t;
    }
  };
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    for (int i = 0; i < SB_RING_STAGES; ++i) mbar_init(&s_rbar[i], 1);
    fence_mbar_init();
    s_item = draw();
    s_err = 0;
  }
  __syncthreads();
  uint32_t phase = 0, ring_phase = 0;

  for (;;) {
    const uint32_t it = s_item;
    if (it >= n_items) break;
    // the next item's ticket is drawn now and travels under this item's work
    uint32_t next_it = 0;
    if (tid == 0) next_it = draw();
    const WorkItem wi = items[it];
    const PageDesc pg = pages[wi.page];
    const ColDesc &col = cols[pg.col];
    const uint32_t side = (pass == 1 && side_flags != nullptr) ? (side_flags[wi.page] & 3u) : 0u;
    const bool lz4_side = side == 1; // value block decoded by sb_lz4_kernel
    const bool stored = side == 2;   // LZ4 block that is one literal run: plain copy here
    if (lz4_side && !col.nullable && col.n_nested <= 1) { // value block handled by sb_lz4_kernel, nothing else in the page
      if (tid == 0 && (wi.tile == 0 || wi.tile == 0xffffffffu)) atomicAdd(codec_hist + SB_C_LZ4, 1u);
      __syncthreads();
      if (tid == 0) s_item = next_it;
      __syncthreads();
      continue;
    }
    const bool plain = side == 3;    // flat page, codec None, header validated by sb_classify_kernel
    // plain pages stage only what precedes the value bytes (validity section + hdr9; nothing when not nullable)
    const uint32_t stage_len = plain ? pg.len - pg.num_values * uint32_t(col.W) : pg.len;
    const bool staged = stage_len + 32 <= stage_cap && !(plain && !col.nullable);

    Dctx cx;
    cx.err = &s_err;
    cx.ws = s_ws;
    cx.bcast = s_bcast;
    cx.ar.g_cur = scratch + uint64_t(blockIdx.x) * scratch_per_cta;
    cx.ar.g_end = cx.ar.g_cur + scratch_per_cta;
    cx.ar.s_end = dsm + smem_bytes;
    cx.rbar = s_rbar;
    cx.rphase = ring_phase;

    const uint8_t *p;
    if (staged) {
      const uint32_t mis = uint32_t(uintptr_t(pg.src) & 15);
      const uint32_t bytes = (mis + stage_len + 15) & ~15u;
      if (tid == 0 && bytes) {
        fence_proxy_async();
        mbar_expect_tx(&s_bar, bytes);
        const uint8_t *g = pg.src - mis;
        for (uint32_t o = 0; o < bytes; o += kTmaChunk) tma_load_1d(dsm + o, g + o, min(kTmaChunk, bytes - o), &s_bar);
      }
      if (bytes) {
        mbar_wait(&s_bar, phase);
        phase ^= 1;
      }
      p = dsm + mis;
      cx.ar.s_cur = dsm + ((bytes + 15) & ~15u) + 16; // 16 bytes of slack after the page
    } else {
      p = pg.src;
      cx.ar.s_cur = dsm;
    }

    cx.page_s = p;
    cx.page_g = pg.src;
    const uint32_t avail = pg.len;
    uint32_t n = pg.num_values;
    bool ok = true;
    const bool nested = !LIGHT && col.n_nested > 1;
    const bool flat_fixed = is_fixed_type(col.type) && !nested;

    if (col.type == SB_NULL) {
      // null.rs: length only, nothing to decode
    } else if (plain && !col.nullable && wi.tile == 0xffffffffu) {
      if (tid == 0) atomicAdd(codec_hist + SB_C_NONE, 1u);
      stream_copy(cx, col.values + pg.out_elem * uint64_t(col.W), p + 9, uint64_t(n) * uint32_t(col.W), true);
    } else if (pass == 1 && !staged && wi.tile != 0xffffffffu && flat_fixed && !col.nullable) {
      // ---- oversized page (e.g. max_page_size = None): None / OneValue are split into
      //      tiles that stream straight from global memory; anything else runs on tile 0.
      int codec = avail >= 9 ? int(p[0]) : -1;
      uint32_t compressed = avail >= 9 ? ld_u32u(p + 1) : 0;
      const uint32_t W = uint32_t(col.W);
      const uint32_t tile_elems = kTileBytes / W;
      uint32_t lo = min(n, wi.tile * tile_elems), hi = min(n, lo + tile_elems);
      uint8_t *dst = col.values + pg.out_elem * W;
      if (wi.tile == 0 && tid == 0 && codec >= 0 && codec < 32) atomicAdd(codec_hist + codec, 1u);
      if (codec == SB_C_NONE && avail >= 9 && compressed <= avail - 9 && uint64_t(compressed) == uint64_t(n) * W) {
        stream_copy(cx, dst + uint64_t(lo) * W, p + 9 + uint64_t(lo) * W, uint64_t(hi - lo) * W, true);
      } else if (stored) { // validated by sb_classify_kernel: [token 0xF0][ne length bytes][n*W literals]
        const uint32_t lit0 = 9 + 1 + ((n * W - 15) / 255 + 1);
        stream_copy(cx, dst + uint64_t(lo) * W, p + lit0 + uint64_t(lo) * W, uint64_t(hi - lo) * W, true);
      } else if (codec == SB_C_ONEVALUE && avail >= 9 + W) {
        switch (W) {
        case 1: dec_onevalue<1>(cx, p + 9, avail - 9, lo, hi, dst); break;
        case 2: dec_onevalue<2>(cx, p + 9, avail - 9, lo, hi, dst); break;
        case 4: dec_onevalue<4>(cx, p + 9, avail - 9, lo, hi, dst); break;
        case 16: dec_onevalue<16>(cx, p + 9, avail - 9, lo, hi, dst); break;
        case 32: dec_onevalue<32>(cx, p + 9, avail - 9, lo, hi, dst); break;
        default: dec_onevalue<8>(cx, p + 9, avail - 9, lo, hi, dst); break;
        }
      } else if (wi.tile == 0) {
        uint32_t used = 0;
        if (!lz4_side) ok = decode_fixed<0, LIGHT>(cx, p, avail, n, col.W, col.is_float != 0, dst, &used);
      }
    } else if (!LIGHT && pass == 1 && !staged && wi.tile != 0xffffffffu && wi.tile != 0 && is_binary_type(col.type) && !nested) {
      // ---- oversized binary page, tiles 1..: a slice of the plain value bytes found by the plan pass
      const PageAux &ax = aux[pg.aux];
      const uint64_t lo = uint64_t(wi.tile - 1) * kBinTile;
      if (ax.val_pos != 0 && lo < ax.value_bytes)
        stream_copy(cx, col.values + pg.out_byte + lo, p + ax.val_pos + lo, min(uint64_t(kBinTile), ax.value_bytes - lo), true);
    } else if (wi.tile == 0 || wi.tile == 0xffffffffu) {
      PageAux *ax = pg.aux != 0xffffffffu ? aux + pg.aux : nullptr;
      uint32_t vb = 0;
      uint64_t out_elem = pg.out_elem;
      if (nested) {
        if constexpr (!LIGHT) {
          // levels section: NestedState entries + leaf validity; n becomes the leaf slot count
          uint32_t leaf_len = 0;
          vb = decode_levels(cx, p, avail, n, col, pass, ax, pg.last != 0, &leaf_len);
          if (vb == 0xffffffffu) ok = false;
          n = leaf_len;
        }
      } else if (col.nullable) {
        if (pass == 1) {
          vb = decode_validity(cx, p, avail, n, col.validity, pg.out_elem);
        } else { // plan pass: only skip the section
          uint32_t L = avail >= 4 ? ld_u32u(p) : 0xffffffffu;
          vb = (avail >= 4 && L <= avail - 4) ? 4 + L : 0xffffffffu;
          if (vb == 0xffffffffu) cx.flag(SB_IO);
        }
        if (vb == 0xffffffffu) ok = false;
      }
      if (ok && pass == 1 && tid == 0 && avail > vb && p[vb] < 32) atomicAdd(codec_hist + p[vb], 1u);
      // plain value bytes of an oversized flat binary page travel as tiles 1.. (same predicate as the host's item list)
      const bool vtiled = pass == 1 && !staged && wi.tile == 0 && is_binary_type(col.type) && !nested && ax && ax->val_pos != 0;
      if (ok && pass == 0) {
        if constexpr (!LIGHT) {
          if (is_binary_type(col.type)) {
            uint64_t vbytes = 0;
            uint32_t val_pos = 0, n_ent = 0;
            ok = binary_page_size(cx, p, avail, vb, n, entries + pg.tab_off, &vbytes, &val_pos, &n_ent, nested ? nullptr : ax);
            ok = !__syncthreads_or(!ok || *cx.err != 0); // per-thread flags (bad dictionary index, ...) count too
            if (tid == 0) {
              ax->value_bytes = ok ? vbytes : 0;
              ax->val_pos = ok ? val_pos : 0;
              ax->n_ent = ok ? n_ent : 0;
              ax->failed = ok ? 0u : 1u;
            }
          } else if (tid == 0) {
            ax->value_bytes = 0;
          }
        }
      } else if (ok && pass == 1 && ax && ax->failed) {
        // rejected by the plan pass (status already holds the reason): its outputs were never sized
      } else if (ok) {
        if (!LIGHT && col.type == SB_BOOL) {
          if constexpr (!LIGHT) ok = decode_boolean(cx, p + vb, avail - vb, n, col.values, out_elem);
        } else if (is_fixed_type(col.type)) {
          uint32_t used = 0;
          // top-level LZ4 blocks are decoded by sb_lz4_kernel (same predicate as sb_classify_kernel)
          if (stored) {
            const uint32_t dlen = n * uint32_t(col.W);
            stream_copy(cx, col.values + out_elem * uint64_t(col.W), p + vb + 9 + 1 + ((dlen - 15) / 255 + 1), dlen);
          } else if (plain) { // value bytes stream from global memory behind the staged validity section
            stream_copy(cx, col.values + out_elem * uint64_t(col.W), pg.src + vb + 9, uint64_t(n) * uint32_t(col.W));
          } else if (!lz4_side)
            ok = decode_fixed<0, LIGHT>(cx, p + vb, avail - vb, n, col.W, col.is_float != 0,
                                        col.values + out_elem * uint64_t(col.W), &used);
        } else if (!LIGHT && col.type == SB_BINARY) {
          if constexpr (!LIGHT)
            ok = decode_binary<4>(cx, p, avail, vb, n, reinterpret_cast<int32_t *>(col.offsets) + out_elem,
                                  col.values + pg.out_byte, pg.out_byte, pg.ordinal == 0, entries + pg.tab_off, vtiled,
                                  ax->n_ent, ax->value_bytes);
        } else if (!LIGHT && col.type == SB_LARGE_BINARY) {
          if constexpr (!LIGHT)
            ok = decode_binary<8>(cx, p, avail, vb, n, reinterpret_cast<int64_t *>(col.offsets) + out_elem,
                                  col.values + pg.out_byte, pg.out_byte, pg.ordinal == 0, entries + pg.tab_off, vtiled,
                                  ax->n_ent, ax->value_bytes);
        } else {
          cx.flag(LIGHT ? SB_PANIC : SB_NYI);
        }
      }
    }
    (void)ok;
    ring_phase = cx.rphase;
    __syncthreads(); // all reads of the staged page / arena done before the next TMA lands
    if (tid == 0) {
      if (s_err) atomicCAS(status + wi.page, 0, s_err);
      s_err = 0;
      s_item = next_it;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SB_NT, 4)
    sb_decode_kernel(const PageDesc *__restrict__ pages, const ColDesc *__restrict__ cols, const WorkItem *__restrict__ items, uint32_t n_items,
                     uint32_t *counter, uint8_t *scratch, uint64_t scratch_per_cta, int32_t *status, uint32_t stage_cap, uint32_t smem_bytes,
                     const uint8_t *__restrict__ side_flags, PageAux *aux, BinEntry *entries, uint32_t *codec_hist, int pass, int split,
                     const uint32_t *__restrict__ n_heavy) {
  if (split && n_heavy && *n_heavy == 0) return; // every page went to the light kernel
  decode_body<false>(pages, cols, items, n_items, counter, scratch, scratch_per_cta, status, stage_cap, smem_bytes, side_flags, aux, entries,
                     codec_hist, pass, split != 0);
}
__global__ void __launch_bounds__(SB_NT, 8)
    sb_decode_light_kernel(const PageDesc *__restrict__ pages, const ColDesc *__restrict__ cols, const WorkItem *__restrict__ items,
                           uint32_t n_items, uint32_t *counter, uint8_t *scratch, uint64_t scratch_per_cta, int32_t *status, uint32_t stage_cap,
                           uint32_t smem_bytes, const uint8_t *__restrict__ side_flags, uint32_t *codec_hist) {
  decode_body<true>(pages, cols, items, n_items, counter, scratch, scratch_per_cta, status, stage_cap, smem_bytes, side_flags, nullptr, nullptr,
                    codec_hist, 1, true);
}

} // namespace sb

// =====================================================================================
// host side
// =====================================================================================
using namespace sb;

#include "sb_host.h"

extern "C" {

const char *sb_version(void) { return "strawboat_b200 0.1 (sm_100a)"; }

#ifdef SB_LZ4_PROF
// diagnostic build only (make PROF=1): cumulative cycles per mover phase, see tools/lz4_prof.py
int32_t sb_debug_lz4_prof(unsigned long long *out32, int32_t reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out32, sb::g_lz4_prof, sizeof(unsigned long long) * 32) != cudaSuccess) return SB_CUDA;
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(sb::g_lz4_prof, z, sizeof(z));
  }
  return SB_OK;
}
#endif

int32_t sb_ctx_create(int32_t device, sb_ctx **out) {
  if (!out) return SB_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return SB_CUDA; // no CPU fallback
  sb_ctx *ctx = new (std::nothrow) sb_ctx();
  if (!ctx) return SB_CUDA;
  ctx->device = device;
  // the side stream carries sb_lz4_kernel, whose blocks must all be resident from the start (each is one
  // long serial chain): highest priority, so that its CTAs are placed before the main kernel's
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
    delete ctx;
    return SB_CUDA;
  }
  ctx->own_stream = true;
  cudaStreamCreateWithFlags(&ctx->aux2, cudaStreamNonBlocking);
  for (cudaEvent_t *ev : {&ctx->ev_l0, &ctx->ev_l1}) cudaEventCreate(ev);
  cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming);
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  for (cudaEvent_t *ev : {&ctx->ev_m0, &ctx->ev_m1, &ctx->ev_lz0, &ctx->ev_lz1}) cudaEventCreate(ev);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_cls, cudaEventDisableTiming);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // keep freed blocks cached in the stream-ordered pool: steady-state calls do not hit the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~uint64_t(0);
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaFuncSetAttribute(sb_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemMax));
  cudaFuncSetAttribute(sb_lz4_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(sb_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  cudaFuncSetAttribute(sb_dict_walk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100); // 12 one-warp CTAs of 18.5 KB per SM
  *out = ctx;
  return SB_OK;
}

void sb_ctx_destroy(sb_ctx *ctx) {
  if (!ctx) return;
  if (ctx->pending.active) { // never collected: give its buffers back
    sb_decode_finish_pending(ctx);
  }
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->aux);
  cudaStreamSynchronize(ctx->aux2);
  for (DevBuf *b : {&ctx->d_tables, &ctx->d_scratch, &ctx->d_entries})
    if (b->p) cudaFreeAsync(b->p, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->h_tables) cudaFreeHost(ctx->h_tables);
  for (auto &b : ctx->pinned_free) cudaFreeHost(b.p);
  for (cudaEvent_t ev : {ctx->ev0, ctx->ev1, ctx->ev_fork, ctx->ev_join, ctx->ev_cls, ctx->ev_m0, ctx->ev_m1, ctx->ev_lz0, ctx->ev_lz1})
    if (ev) cudaEventDestroy(ev);
  for (cudaEvent_t ev : {ctx->ev_l0, ctx->ev_l1, ctx->ev_join2})
    if (ev) cudaEventDestroy(ev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->aux);
  cudaStreamDestroy(ctx->aux2);
  delete ctx;
}

int32_t sb_ctx_set_stream(sb_ctx *ctx, void *cuda_stream) {
  if (!ctx) return SB_INVALID_ARG;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->own_stream = false;
  return SB_OK;
}

const char *sb_last_error(const sb_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context (CUDA device unavailable)"; }

// ---- page inspector (src/stat.rs:63-152), host only ------------------------------------
namespace {
uint32_t rd_u32(const uint8_t *p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
const char *codec_name(int c) {
  switch (c) {
  case SB_C_NONE: return "None";
  case SB_C_LZ4: return "Lz4";
  case SB_C_ZSTD: return "Zstd";
  case SB_C_SNAPPY: return "Snappy";
  case SB_C_RLE: return "Rle";
  case SB_C_DICT: return "Dict";
  case SB_C_ONEVALUE: return "OneValue";
  case SB_C_FREQ: return "Freq";
  case SB_C_BITPACK: return "Bitpacking";
  case SB_C_DELTABP: return "DeltaBitpacking";
  case SB_C_PATAS: return "Patas";
  default: return nullptr;
  }
}
// stat_body: one hdr9 + its sub-pages.  `lvl` = position on the path.
int stat_block(int type, const uint8_t *in, uint64_t len, sb_page_info *info, int lvl, std::string &tree) {
  if (len < 9) return SB_IO;
  const int codec = in[0];
  const char *name = codec_name(codec);
  if (!name) return SB_OUT_OF_SPEC; // Compression::from_codec (compression/mod.rs:78)
  const uint32_t compressed = rd_u32(in + 1);
  if (lvl == 0) {
    info->codec = codec;
    info->compressed_size = compressed;
    info->uncompressed_size = rd_u32(in + 5);
  }
  if (lvl < 4) info->path[lvl] = codec;
  info->depth = lvl + 1;
  tree += name;
  const uint8_t *body = in + 9;
  const uint64_t avail = len - 9;
  const bool binary = type == SB_BINARY || type == SB_LARGE_BINARY;
  // same nesting cap as the device decoders (dec_dict / dec_freq: at most 3 stacked headers), so a crafted
  // chain of Dict / Freq headers cannot recurse once per 13 bytes of page
  if ((codec == SB_C_DICT || codec == SB_C_FREQ) && lvl >= 2) return SB_OUT_OF_SPEC;
  if (codec == SB_C_DICT) { // stat_dict_body: [index page (u32)][u32 unique_num]
    if (avail < 9) return SB_IO;
    const uint64_t sub_len = 9 + uint64_t(rd_u32(body + 1));
    if (sub_len + 4 > avail) return SB_IO;
    std::string sub;
    int rc = stat_block(SB_U32, body, sub_len, info, lvl + 1, sub);
    if (rc) return rc;
    const uint32_t k = rd_u32(body + sub_len);
    if (lvl == 0) info->unique_num = k;
    tree += "(" + sub + ")[k=" + std::to_string(k) + "]";
  } else if (codec == SB_C_FREQ) { // stat_freq_body
    const uint64_t top = binary ? 8 : uint64_t(type_width(type));
    if (avail < top + 4) return SB_IO;
    uint64_t p = top;
    if (binary) {
      uint64_t l = 0;
      for (int b = 0; b < 8; ++b) l |= uint64_t(body[b]) << (8 * b);
      if (l > avail - 12) return SB_IO;
      p += l;
    }
    const uint32_t bm = rd_u32(body + p);
    if (lvl == 0) info->exceptions_bitmap_size = bm;
    if (!binary) { // exceptions are a nested page of the same type
      if (uint64_t(bm) + p + 4 > avail) return SB_IO;
      std::string sub;
      int rc = stat_block(type, body + p + 4 + bm, avail - p - 4 - bm, info, lvl + 1, sub);
      if (rc) return rc;
      tree += "(" + sub + ")";
    }
  }
  return SB_OK;
}
} // namespace

int32_t sb_stat_page(const sb_leaf *leaf, const uint8_t *page, uint64_t len, sb_page_info *info, char *tree, uint64_t tree_cap) {
  if (!leaf || !info || (len && !page)) return SB_INVALID_ARG;
  std::memset(info, 0, sizeof(*info));
  info->validity_size = 0xffffffffu;
  info->codec = -1;
  for (int i = 0; i < 4; ++i) info->path[i] = -1;
  if (tree && tree_cap) tree[0] = 0;
  if (leaf->type == SB_NULL) return SB_OK; // empty page
  uint64_t vb = 0;
  if (leaf->n_nested > 1) { // [u32 rows][u32 rep_len][u32 def_len][rep][def]
    if (len < 12) return SB_IO;
    vb = 12 + uint64_t(rd_u32(page + 4)) + rd_u32(page + 8);
    if (vb > len) return SB_IO;
    info->levels_size = uint32_t(vb);
  } else if (leaf->nullable) {
    if (len < 4) return SB_IO;
    const uint32_t L = rd_u32(page);
    if (L > len - 4) return SB_IO;
    info->validity_size = L;
    vb = 4 + uint64_t(L);
  }
  std::string t;
  int type = leaf->type;
  if (type == SB_BOOL) type = SB_U8; // boolean blocks: None / Lz4 / Rle / OneValue, no sub-pages
  int rc = stat_block(type, page + vb, len - vb, info, 0, t);
  if (rc) return rc;
  if (tree && tree_cap) {
    const size_t n = std::min<size_t>(t.size(), size_t(tree_cap) - 1);
    std::memcpy(tree, t.data(), n);
    tree[n] = 0;
  }
  return SB_OK;
}

int32_t sb_last_stats(const sb_ctx *ctx, sb_stats *out) {
  if (!ctx || !out) return SB_INVALID_ARG;
  *out = ctx->stats;
  return SB_OK;
}

void sb_release_columns(sb_ctx *ctx, sb_column_out *outs, uint64_t n) {
  if (!ctx || !outs) return;
  cudaSetDevice(ctx->device);
  for (uint64_t i = 0; i < n; ++i) {
    Owner *o = static_cast<Owner *>(outs[i]._owner);
    if (!o) continue;
    for (void *p : o->dev) cudaFreeAsync(p, ctx->stream);
    for (auto &b : o->host_pinned) pinned_put(ctx, b);
    for (void *p : o->host_malloc) std::free(p);
    delete o;
    std::memset(&outs[i], 0, sizeof(outs[i]));
  }
}

// value-byte tiles of an unstaged binary page: a plain (codec None) page holds (n + 1) offsets before its
// value bytes, so at most length - (n + 1) * OW of them
static uint64_t bin_value_tiles(const sb_page_meta &m, uint64_t OW) {
  const uint64_t off_bytes = (m.num_values + 1) * OW;
  const uint64_t vmax = m.length > off_bytes ? m.length - off_bytes : 0;
  return (vmax + kBinTile - 1) / kBinTile;
}

} // extern "C"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Submits every copy and kernel of one decode call on the context's streams and records what sb_decode_finish_pending
// needs to hand the results over.  `bufs` (optional): caller-owned output buffers per column.  `sizes` (optional):
// plan only -- run the size pass, report the buffer sizes, decode nothing.
static int32_t decode_submit(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem, sb_column_out *outs,
                             const sb_out_buffers *bufs, sb_column_sizes *sizes) {
  if (!ctx) return SB_CUDA;
  if (!cols || (!outs && !sizes) || (out_mem != SB_MEM_HOST && out_mem != SB_MEM_DEVICE)) return fail(ctx, SB_INVALID_ARG, "bad arguments");
  if (ctx->pending.active) sb_decode_finish_pending(ctx); // one outstanding call per context: the previous one is collected first
  std::vector<sb_column_out> plan_outs;
  if (sizes) { // plan only: outputs are never materialised for the caller
    plan_outs.resize(n_cols);
    outs = plan_outs.data();
  }
  SB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::memset(outs, 0, sizeof(sb_column_out) * n_cols);
  ctx->stats = sb_stats{};
  const double t_host0 = now_ms();

  // ---- host pass: validate, count pages / work items
  uint64_t n_pages_total = 0, n_items = 0, n_plan = 0, n_entries = 0, max_elems_bytes = 0;
  uint32_t max_stage = 0;
  uint64_t max_need = 0; // staged page + a u32 index / rank buffer for its rows (Dict, Freq) + small tables
  const uint32_t stage_cap = kSmemMax - kArenaMin;
  auto col_binary = [](const sb_column_in &ci) { return ci.leaf.type == SB_BINARY || ci.leaf.type == SB_LARGE_BINARY; };
  auto col_nested = [](const sb_column_in &ci) { return ci.leaf.n_nested > 1; };
  for (uint64_t c = 0; c < n_cols; ++c) {
    const sb_column_in &ci = cols[c];
    if (ci.leaf.type < SB_NULL || ci.leaf.type > SB_I256) return fail(ctx, SB_NYI, "unsupported physical type");
    if (ci.leaf.n_nested > SB_MAX_NESTED) return fail(ctx, SB_NYI, "nesting deeper than SB_MAX_NESTED");
    if (col_nested(ci)) {
      if (ci.leaf.nested_kind[ci.leaf.n_nested - 1] != SB_N_PRIMITIVE) return fail(ctx, SB_INVALID_ARG, "the last nested entry must be the primitive leaf");
      if (ci.leaf.type == SB_NULL) return fail(ctx, SB_NYI, "nested Null leaves");
    }
    if (ci.n_pages && !ci.metas) return fail(ctx, SB_INVALID_ARG, "metas is NULL");
    if (ci.nbytes && !ci.bytes) return fail(ctx, SB_INVALID_ARG, "bytes is NULL");
    n_pages_total += ci.n_pages;
    const bool plan = col_binary(ci) || col_nested(ci);
    const uint64_t W = std::max(1, type_width(ci.leaf.type));
    for (uint64_t p = 0; p < ci.n_pages; ++p) {
      const sb_page_meta &m = ci.metas[p];
      if (m.length > 0xffffffffull || m.num_values > 0xffffffffull) return fail(ctx, SB_OUT_OF_SPEC, "page larger than 4 GiB (u32 size fields)");
      if (m.length + 32 <= stage_cap) {
        n_items += 1;
        // a fixed-width page at least as long as its raw values is a plain page or a stored LZ4 block: its value
        // bytes stream through the ring from global memory, only what precedes them is ever staged
        const bool raw_sized = fixed_type(ci.leaf.type) && !col_nested(ci) && m.length >= m.num_values * W + 9;
        max_stage = std::max<uint32_t>(max_stage, uint32_t(raw_sized ? m.length - m.num_values * W : m.length));
        // a page less than half its decoded size is Dict / Freq / RLE coded (binary: any page): give its
        // index buffer (4 bytes per row) room in shared memory instead of the L2 scratch
        if (col_binary(ci) || 2 * m.length < m.num_values * W)
          max_need = std::max<uint64_t>(max_need, m.length + 4 * m.num_values + (col_binary(ci) ? 16 * 1024 : kNeedExtra)); // binary: + the entry table
      } else {
        uint64_t out_bytes = ci.leaf.type == SB_BOOL ? (m.num_values + 7) / 8 : m.num_values * W;
        bool tiled = !ci.leaf.nullable && fixed_type(ci.leaf.type) && !col_nested(ci);
        if (col_binary(ci) && !col_nested(ci)) n_items += 1 + bin_value_tiles(m, W); // tile 0 + value slices
        else n_items += tiled ? std::max<uint64_t>(1, (out_bytes + kTileBytes - 1) / kTileBytes) : 1;
      }
      if (plan) n_plan += 1;
      if (col_binary(ci)) n_entries += m.length / 8 + 1;
      max_elems_bytes = std::max<uint64_t>(max_elems_bytes, m.num_values * std::max<uint64_t>(W, 4));
    }
  }
  std::vector<Owner *> owners(n_cols, nullptr);
  std::vector<void *> d_inputs; // device copies of host inputs, freed at the end of the call
  auto cleanup = [&]() {
    // kernels already launched on either stream may still write into the outputs: join before freeing
    cudaStreamSynchronize(ctx->aux);
    cudaStreamSynchronize(ctx->aux2);
    cudaStreamSynchronize(st);
    for (void *d : d_inputs) cudaFreeAsync(d, st);
    d_inputs.clear();
    for (uint64_t c = 0; c < n_cols; ++c) outs[c]._owner = owners[c];
    sb_release_columns(ctx, outs, n_cols);
  };
#define SB_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != SB_OK) {        \
      cleanup();                \
      return rc__;              \
    }                           \
  } while (0)
#define SB_TRY_CUDA(call)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      cleanup();                                                                               \
      return fail(ctx, SB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    }                                                                                          \
  } while (0)

  // one pinned staging buffer, mirrored on the device:
  //   uploaded : [ColDesc * n_cols][PageDesc * P][WorkItem * I][WorkItem * plan][PageAux * plan]
  //   zeroed   : [status * P][counters * 48][side_flags * P]      device only: [Lz4Job * P]
  size_t off_cols = 0;
  size_t off_pages = align_up(off_cols + sizeof(ColDesc) * n_cols, 16);
  size_t off_items = align_up(off_pages + sizeof(PageDesc) * n_pages_total, 16);
  size_t off_items0 = align_up(off_items + sizeof(WorkItem) * n_items, 16);
  size_t off_aux = align_up(off_items0 + sizeof(WorkItem) * n_plan, 16);
  size_t tables_bytes = align_up(off_aux + sizeof(PageAux) * n_plan, 16);
  size_t off_status = tables_bytes;
  size_t off_counters = align_up(off_status + sizeof(int32_t) * n_pages_total, 16);
  size_t off_flags = off_counters + 48 * 4;
  size_t zero_end = align_up(off_flags + n_pages_total, 16);
  size_t off_jobs = zero_end;
  size_t dev_bytes = off_jobs + sizeof(Lz4Job) * n_pages_total;
  int rc;
  if ((rc = host_tables_reserve(ctx, zero_end))) return rc;
  if ((rc = dev_reserve(ctx, ctx->d_tables, dev_bytes))) return rc;
  if (n_entries && (rc = dev_reserve(ctx, ctx->d_entries, n_entries * sizeof(BinEntry)))) return rc;
  uint8_t *hT = static_cast<uint8_t *>(ctx->h_tables);
  uint8_t *dT = static_cast<uint8_t *>(ctx->d_tables.p);
  ColDesc *h_cols = reinterpret_cast<ColDesc *>(hT + off_cols);
  PageDesc *h_pages = reinterpret_cast<PageDesc *>(hT + off_pages);
  WorkItem *h_items = reinterpret_cast<WorkItem *>(hT + off_items);
  WorkItem *h_items0 = reinterpret_cast<WorkItem *>(hT + off_items0);
  PageAux *h_aux = reinterpret_cast<PageAux *>(hT + off_aux);
  std::memset(h_aux, 0, sizeof(PageAux) * n_plan);

  // device allocation of an output buffer, owned by the column
  // `user` / `cap`: a caller-owned DEVICE buffer for this output (sb_out_buffers, out_mem = SB_MEM_DEVICE): used in place
  auto dev_out = [&](Owner *ow, uint64_t bytes, bool zero, uint8_t **out, void *user = nullptr, uint64_t cap = 0) -> int {
    if (sizes) { // plan only: nothing is written, a token buffer keeps the bookkeeping uniform
      bytes = 0;
      user = nullptr;
    }
    if (user && out_mem == SB_MEM_DEVICE) {
      if (cap < bytes) return fail(ctx, SB_INVALID_ARG, "caller-provided output buffer is too small (see sb_plan_columns)");
      if (uintptr_t(user) & 15) return fail(ctx, SB_INVALID_ARG, "caller-provided output buffers must be 16-byte aligned");
      if (zero && bytes) SB_CUDA_CHECK(ctx, cudaMemsetAsync(user, 0, std::min<uint64_t>(cap, align_up(bytes, 4)), st));
      *out = static_cast<uint8_t *>(user);
      return SB_OK;
    }
    void *d = nullptr;
    SB_CUDA_CHECK(ctx, cudaMallocAsync(&d, bytes + 16, st));
    ow->dev.push_back(d);
    if (zero) SB_CUDA_CHECK(ctx, cudaMemsetAsync(d, 0, bytes + 16, st));
    *out = static_cast<uint8_t *>(d);
    return SB_OK;
  };
  auto ubuf = [&](uint64_t c) -> const sb_out_buffers * { return bufs ? &bufs[c] : nullptr; };

  uint64_t pi = 0, ii = 0, i0 = 0, ent = 0, bytes_in = 0, bytes_out = 0;
  bool any_fixed = false;
  for (uint64_t c = 0; c < n_cols; ++c) {
    const sb_column_in &ci = cols[c];
    Owner *ow = new Owner();
    owners[c] = ow;
    const int W = type_width(ci.leaf.type);
    const bool nested = col_nested(ci), binary = col_binary(ci);
    any_fixed |= fixed_type(ci.leaf.type);
    uint64_t rows = 0, total_len = 0;
    for (uint64_t p = 0; p < ci.n_pages; ++p) {
      rows += ci.metas[p].num_values;
      total_len += ci.metas[p].length;
    }
    if (total_len > ci.nbytes) {
      cleanup();
      return fail(ctx, SB_IO, "column bytes shorter than the sum of its page lengths");
    }
    // input
    const uint8_t *d_in = ci.bytes;
    if (ci.mem == SB_MEM_HOST && total_len) {
      void *d = nullptr;
      SB_TRY_CUDA(cudaMallocAsync(&d, align_up(total_len + 32, 256), st));
      d_inputs.push_back(d);
      SB_TRY_CUDA(cudaMemcpyAsync(d, ci.bytes, total_len, cudaMemcpyHostToDevice, st));
      d_in = static_cast<const uint8_t *>(d);
    }
    // outputs whose size is known up front (flat columns; binary value bytes come later)
    sb_column_out &o = outs[c];
    o.mem = out_mem;
    ColDesc &cd = h_cols[c];
    std::memset(&cd, 0, sizeof(cd));
    cd.type = ci.leaf.type;
    cd.nullable = !nested && ci.leaf.nullable != 0 && ci.leaf.type != SB_NULL;
    cd.W = W;
    cd.is_float = ci.leaf.type == SB_F32 || ci.leaf.type == SB_F64;
    cd.n_nested = nested ? ci.leaf.n_nested : 0;
    if (nested) {
      for (int d = 0; d < ci.leaf.n_nested; ++d) {
        cd.kind[d] = uint8_t(ci.leaf.nested_kind[d]);
        cd.nnull[d] = ci.leaf.nested_nullable[d] != 0;
        cd.cum_sum[d + 1] = uint8_t(cd.cum_sum[d] + cd.nnull[d] + (cd.kind[d] == SB_N_LIST));
        cd.cum_rep[d + 1] = uint8_t(cd.cum_rep[d] + (cd.kind[d] == SB_N_LIST));
      }
    } else {
      o.length = rows;
      cd.length = rows;
      uint64_t bitmap_bytes = align_up((rows + 7) / 8, 4);
      const sb_out_buffers *ub = ubuf(c);
      if (ci.leaf.type == SB_BOOL) {
        o.values_bytes = (rows + 7) / 8;
        SB_TRY(dev_out(ow, bitmap_bytes, true, &cd.values, ub ? ub->values : nullptr, ub ? ub->values_cap : 0));
      } else if (binary) {
        o.offsets_bytes = (rows + 1) * uint64_t(W);
        SB_TRY(dev_out(ow, o.offsets_bytes, false, &cd.offsets, ub ? ub->offsets : nullptr, ub ? ub->offsets_cap : 0));
        if (!sizes) SB_TRY_CUDA(cudaMemsetAsync(cd.offsets, 0, size_t(W), st)); // offsets[0] = 0 even for a column without pages
      } else if (W && ci.leaf.type != SB_NULL) {
        o.values_bytes = rows * uint64_t(W);
        SB_TRY(dev_out(ow, o.values_bytes, false, &cd.values, ub ? ub->values : nullptr, ub ? ub->values_cap : 0));
      }
      if (cd.nullable) {
        o.validity_bytes = (rows + 7) / 8;
        SB_TRY(dev_out(ow, bitmap_bytes, true, &cd.validity, ub ? ub->validity : nullptr, ub ? ub->validity_cap : 0));
      }
    }
    // pages
    uint64_t src_off = 0, elem = 0;
    for (uint64_t p = 0; p < ci.n_pages; ++p) {
      const sb_page_meta &m = ci.metas[p];
      PageDesc &pd = h_pages[pi];
      pd.src = d_in + src_off;
      pd.len = uint32_t(m.length);
      pd.num_values = uint32_t(m.num_values);
      pd.col = uint32_t(c);
      pd.ordinal = uint32_t(p);
      pd.out_elem = elem;
      pd.out_byte = 0;
      pd.aux = 0xffffffffu;
      pd.last = p + 1 == ci.n_pages;
      pd.tab_off = 0;
      if (binary || nested) {
        pd.aux = uint32_t(i0);
        h_items0[i0++] = WorkItem{uint32_t(pi), 0xffffffffu};
      }
      if (binary) {
        pd.tab_off = ent;
        ent += m.length / 8 + 1;
      }
      if (m.length + 32 <= stage_cap) {
        h_items[ii++] = WorkItem{uint32_t(pi), 0xffffffffu};
      } else {
        uint64_t out_b = ci.leaf.type == SB_BOOL ? (m.num_values + 7) / 8 : m.num_values * uint64_t(std::max(1, W));
        bool tiled = !ci.leaf.nullable && fixed_type(ci.leaf.type) && !nested;
        uint64_t nt = tiled ? std::max<uint64_t>(1, (out_b + kTileBytes - 1) / kTileBytes) : 1;
        if (binary && !nested) nt = 1 + bin_value_tiles(m, uint64_t(std::max(1, W)));
        for (uint64_t t = 0; t < nt; ++t) h_items[ii++] = WorkItem{uint32_t(pi), uint32_t(t)};
      }
      src_off += m.length;
      elem += m.num_values;
      bytes_in += m.length;
      ++pi;
    }
  }

  // ---- launch configuration
  uint32_t smem = uint32_t(align_up(std::min<uint64_t>(uint64_t(max_stage) + 48, stage_cap) + kArenaMin, 1024));
  smem = std::max<uint32_t>(smem, uint32_t(std::min<uint64_t>(align_up(max_need, 1024), kSmemMax)));
  smem = std::min(std::max(smem, kSmemMin), kSmemMax);
  if (ctx->occ_smem != smem) { // occupancy queries cost tens of microseconds: once per shared-memory size
    int q = 1;
    SB_TRY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, sb_decode_kernel, SB_NT, smem));
    ctx->occ_smem = smem;
    ctx->occ_val = std::max(1, q);
  }
  const int occ = ctx->occ_val;
  uint32_t grid = uint32_t(std::min<uint64_t>(std::max<uint64_t>(n_items, 1), uint64_t(ctx->sm_count) * occ));
  uint64_t scratch_per_cta = align_up(3 * (max_elems_bytes + 64) + 16 * 1024, 256);
  const PageDesc *d_pages = reinterpret_cast<const PageDesc *>(dT + off_pages);
  const ColDesc *d_cols = reinterpret_cast<const ColDesc *>(dT + off_cols);
  int32_t *d_status = reinterpret_cast<int32_t *>(dT + off_status);
  // counters: [0] main queue, [1] lz4 queue, [2] long lz4 jobs, [3] short lz4 jobs, [4] plan queue,
  //           [5..36] pages per top-level codec, [37] light queue, [38..39] u64 bytes handled by sb_lz4_kernel,
  //           [40] pages of the full kernel's family (sb_classify_kernel)
  uint32_t *d_counters = reinterpret_cast<uint32_t *>(dT + off_counters);
  uint8_t *d_flags = dT + off_flags;
  Lz4Job *d_jobs = reinterpret_cast<Lz4Job *>(dT + off_jobs);
  PageAux *d_aux = reinterpret_cast<PageAux *>(dT + off_aux);
  BinEntry *d_entries = static_cast<BinEntry *>(ctx->d_entries.p);
  const auto t_planned = std::chrono::steady_clock::now();
  // Light / full split of the decode kernel (SB_SPLIT=1; OFF by default).  Measured on B200 (profiles/README.md, round 2):
  // the light instantiation (64 registers) makes Dict / RLE pages of a column decoded alone ~20 % faster at 8 CTAs per
  // SM, but next to the LZ4 kernel (161 KiB of shared memory per SM) and the full kernel (74 KiB per CTA) the three do
  // not co-reside; at the 5 + 2 CTAs per SM that do fit, plain pages lose their 32 KiB TMA ring and the step gets slower
  // (config 2: 362 -> 297 GB/s).  Kept selectable for experiments; the default is one kernel at its own occupancy.
  static const bool want_split = std::getenv("SB_SPLIT") != nullptr;
  bool fixed_only = any_fixed && want_split;
  for (uint64_t c = 0; c < n_cols; ++c) fixed_only = fixed_only && fixed_type(cols[c].leaf.type) && !col_nested(cols[c]);
  const bool split = fixed_only && n_items > 0;
  uint32_t grid_light = 0, smem_heavy = smem;
  if (split) {
    if (ctx->light_occ == 0) {
      int q = 1;
      SB_TRY_CUDA(cudaFuncSetAttribute(sb_decode_light_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      SB_TRY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, sb_decode_light_kernel, SB_NT, kLightSmem));
      ctx->light_occ = std::max(1, q);
    }
    grid_light = uint32_t(std::min<uint64_t>(n_items, uint64_t(ctx->sm_count) * std::min(ctx->light_occ, 5)));
    // 5 x 26 KiB + 2 x 40 KiB per SM: the full kernel stages pages up to 34 KiB here and reads larger ones in place
    smem_heavy = std::min<uint32_t>(smem, 40 * 1024);
    grid = uint32_t(std::min<uint64_t>(grid, uint64_t(ctx->sm_count) * 2));
  }
  const uint32_t stage_cap_heavy = smem_heavy - kArenaMin; // what the kernel may stage: its ACTUAL shared memory, not the host's upper bound
  if (n_items) {
    SB_TRY(dev_reserve(ctx, ctx->d_scratch, scratch_per_cta * (uint64_t(grid) + grid_light)));
    SB_TRY_CUDA(cudaMemcpyAsync(dT, hT, tables_bytes, cudaMemcpyHostToDevice, st));
    SB_TRY_CUDA(cudaMemsetAsync(dT + off_status, 0, zero_end - off_status, st));
    SB_TRY_CUDA(cudaEventRecord(ctx->ev0, st));
  }

  // ---- pass 0 (plan): sizes of binary / nested pages, then the host lays out their outputs
  if (n_plan) {
    uint32_t grid0 = uint32_t(std::min<uint64_t>(n_plan, uint64_t(ctx->sm_count) * occ));
    if (n_entries) { // flat or nested binary columns present: walk the dictionaries of flat Dict pages, one warp per page
      const uint32_t gridw = uint32_t(std::min<uint64_t>(n_plan, uint64_t(ctx->sm_count) * 12));
      sb_dict_walk_kernel<<<gridw, 32, kWalkSmem, st>>>(d_pages, d_cols, reinterpret_cast<const WorkItem *>(dT + off_items0), uint32_t(n_plan), d_aux,
                                                           d_entries);
      SB_TRY_CUDA(cudaGetLastError());
      ctx->stats.kernel_launches += 1;
    }
    sb_decode_kernel<<<grid0, SB_NT, smem, st>>>(d_pages, d_cols, reinterpret_cast<const WorkItem *>(dT + off_items0), uint32_t(n_plan),
                                                 d_counters + 4, static_cast<uint8_t *>(ctx->d_scratch.p), scratch_per_cta, d_status,
                                                 smem - kArenaMin, smem, nullptr, d_aux, d_entries, d_counters + 5, 0, 0, nullptr);
    SB_TRY_CUDA(cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    SB_TRY_CUDA(cudaMemcpyAsync(h_aux, d_aux, sizeof(PageAux) * n_plan, cudaMemcpyDeviceToHost, st));
    SB_TRY_CUDA(cudaStreamSynchronize(st));
    pi = 0;
    for (uint64_t c = 0; c < n_cols; ++c) {
      const sb_column_in &ci = cols[c];
      const bool nested = col_nested(ci), binary = col_binary(ci);
      if (!nested && !binary) {
        pi += ci.n_pages;
        continue;
      }
      Owner *ow = owners[c];
      sb_column_out &o = outs[c];
      ColDesc &cd = h_cols[c];
      const int D = nested ? ci.leaf.n_nested : 0;
      uint64_t base[SB_MAX_NESTED] = {0}, vbytes = 0;
      for (uint64_t p = 0; p < ci.n_pages; ++p, ++pi) {
        PageDesc &pd = h_pages[pi];
        PageAux &ax = h_aux[pd.aux];
        pd.out_byte = vbytes;
        vbytes += ax.value_bytes;
        if (nested) {
          pd.out_elem = base[D - 1];
          for (int d = 0; d < D; ++d) {
            ax.base[d] = base[d];
            base[d] += ax.cnt[d];
          }
        }
      }
      if (nested) {
        const uint64_t rows = base[D - 1]; // leaf slots
        const int W = type_width(ci.leaf.type);
        o.length = rows;
        cd.length = rows;
        const uint64_t bitmap_bytes = align_up((rows + 7) / 8, 4);
        if (ci.leaf.type == SB_BOOL) {
          o.values_bytes = (rows + 7) / 8;
          SB_TRY(dev_out(ow, bitmap_bytes, true, &cd.values));
        } else if (binary) {
          o.offsets_bytes = (rows + 1) * uint64_t(W);
          SB_TRY(dev_out(ow, o.offsets_bytes, false, &cd.offsets));
          if (!sizes) SB_TRY_CUDA(cudaMemsetAsync(cd.offsets, 0, size_t(W), st));
        } else {
          o.values_bytes = rows * uint64_t(W);
          SB_TRY(dev_out(ow, o.values_bytes, false, &cd.values));
        }
        if (cd.nnull[D - 1]) {
          o.validity_bytes = (rows + 7) / 8;
          SB_TRY(dev_out(ow, bitmap_bytes, true, &cd.validity));
        }
        for (int d = 0; d + 1 < D; ++d) {
          o.nested_len[d] = base[d];
          if (cd.kind[d] == SB_N_LIST) {
            uint8_t *q = nullptr;
            SB_TRY(dev_out(ow, (base[d] + 1) * 8, ci.n_pages == 0, &q));
            cd.nest_off[d] = reinterpret_cast<int64_t *>(q);
          }
          if (cd.nnull[d]) SB_TRY(dev_out(ow, align_up((base[d] + 7) / 8, 4), true, &cd.nest_val[d]));
        }
      }
      if (binary) {
        if (ci.leaf.type == SB_BINARY && vbytes > 0x7fffffffull) { // Offsets<i32>::try_push overflows (arrow2): no silent wrap
          cleanup();
          return fail(ctx, SB_OUT_OF_SPEC, "column " + std::to_string(c) + ": more than 2 GiB of value bytes do not fit i32 offsets");
        }
        o.values_bytes = vbytes;
        const sb_out_buffers *ub = nested ? nullptr : ubuf(c);
        SB_TRY(dev_out(ow, vbytes, false, &cd.values, ub ? ub->values : nullptr, ub ? ub->values_cap : 0));
      }
    }
    SB_TRY_CUDA(cudaMemcpyAsync(dT, hT, tables_bytes, cudaMemcpyHostToDevice, st));
  }
  if (sizes) { // sb_plan_columns: the sizes are known now; nothing else runs
    for (uint64_t c = 0; c < n_cols; ++c) {
      sizes[c].length = outs[c].length;
      sizes[c].values_bytes = outs[c].values_bytes;
      sizes[c].offsets_bytes = outs[c].offsets_bytes;
      sizes[c].validity_bytes = outs[c].validity_bytes;
    }
    cleanup();
    return SB_OK;
  }

  // ---- pass 1 (decode)
  if (n_items) {
    if (any_fixed) {
      // D0: find top-level LZ4 blocks; run them next to the main kernel.  The two kernels cannot share an SM
      // at full occupancy (registers), and the main kernel is persistent: if its CTAs get there first the LZ4
      // blocks start only when it has finished.  So the side stream (high priority) runs classify and the LZ4
      // kernel back to back, and the main kernel is released by an event recorded between the two: the LZ4
      // CTAs are placed first, the main kernel's CTAs fill the SMs as LZ4 blocks retire.
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_fork, st));
      SB_TRY_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
      sb_classify_kernel<<<uint32_t((n_pages_total + 7) / 8), 256, 0, ctx->aux>>>(d_pages, d_cols, uint32_t(n_pages_total), d_jobs,
                                                                                    d_counters + 2, d_flags, d_aux, kLightSmem - kArenaMin,
                                                                                    split ? d_counters + 40 : nullptr);
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_cls, ctx->aux));
      if (ctx->lz4_occ == 0) {
        int q = 1;
        SB_TRY_CUDA(cudaFuncSetAttribute(sb_lz4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SB_LZ4_SMEM)));
        SB_TRY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, sb_lz4_kernel, SB_LZ4_CTA, SB_LZ4_SMEM));
        ctx->lz4_occ = std::max(1, q);
      }
      const int lz4_occ = ctx->lz4_occ;
      uint32_t lz4_grid = uint32_t(std::min<uint64_t>((n_pages_total + SB_LZ4_PAIRS - 1) / SB_LZ4_PAIRS, uint64_t(ctx->sm_count) * std::max(1, lz4_occ)));
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_lz0, ctx->aux));
      sb_lz4_kernel<<<lz4_grid, SB_LZ4_CTA, SB_LZ4_SMEM, ctx->aux>>>(d_jobs, d_counters + 2, uint32_t(n_pages_total), d_counters + 1, d_status,
                                                   reinterpret_cast<unsigned long long *>(d_counters + 38));
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_lz1, ctx->aux));
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux));
      ctx->stats.kernel_launches += 2;
    }
    if (split) {
      // the light kernel (flat fixed-width pages with light codec trees) on its own stream, next to the full one
      SB_TRY_CUDA(cudaStreamWaitEvent(ctx->aux2, ctx->ev_cls, 0));
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_l0, ctx->aux2));
      sb_decode_light_kernel<<<grid_light, SB_NT, kLightSmem, ctx->aux2>>>(
          d_pages, d_cols, reinterpret_cast<const WorkItem *>(dT + off_items), uint32_t(n_items), d_counters + 37,
          static_cast<uint8_t *>(ctx->d_scratch.p) + scratch_per_cta * grid, scratch_per_cta, d_status, kLightSmem - kArenaMin, kLightSmem, d_flags,
          d_counters + 5);
      SB_TRY_CUDA(cudaGetLastError());
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_l1, ctx->aux2));
      SB_TRY_CUDA(cudaEventRecord(ctx->ev_join2, ctx->aux2));
      ctx->stats.kernel_launches += 1;
    }
    if (any_fixed) SB_TRY_CUDA(cudaStreamWaitEvent(st, ctx->ev_cls, 0));
    SB_TRY_CUDA(cudaEventRecord(ctx->ev_m0, st));
    sb_decode_kernel<<<grid, SB_NT, smem_heavy, st>>>(d_pages, d_cols, reinterpret_cast<const WorkItem *>(dT + off_items),
                                                uint32_t(n_items), d_counters, static_cast<uint8_t *>(ctx->d_scratch.p),
                                                scratch_per_cta, d_status, stage_cap_heavy, smem_heavy, any_fixed ? d_flags : nullptr, d_aux,
                                                d_entries, d_counters + 5, 1, split ? 1 : 0, d_counters + 40);
    SB_TRY_CUDA(cudaGetLastError());
    SB_TRY_CUDA(cudaEventRecord(ctx->ev_m1, st));
    if (any_fixed) SB_TRY_CUDA(cudaStreamWaitEvent(st, ctx->ev_join, 0));
    if (split) SB_TRY_CUDA(cudaStreamWaitEvent(st, ctx->ev_join2, 0));
    SB_TRY_CUDA(cudaEventRecord(ctx->ev1, st));
    ctx->stats.kernel_launches += 1;
    // statuses + codec histogram back (pinned), reuse the tail of the host table buffer
    SB_TRY_CUDA(cudaMemcpyAsync(hT + off_status, dT + off_status, off_flags - off_status, cudaMemcpyDeviceToHost, st));
  }

  // ---- results to the caller
  auto to_caller = [&](Owner *ow, const void *dev, uint64_t bytes, void **out, void *user = nullptr, uint64_t cap = 0) -> int {
    if (!dev) return SB_OK;
    if (out_mem == SB_MEM_DEVICE) {
      *out = const_cast<void *>(dev);
      return SB_OK;
    }
    if (user) { // caller-owned host buffer (pinned memory keeps the copy asynchronous)
      if (cap < bytes) return fail(ctx, SB_INVALID_ARG, "caller-provided output buffer is too small (see sb_plan_columns)");
      if (bytes) SB_CUDA_CHECK(ctx, cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, st));
      *out = user;
      return SB_OK;
    }
    PinnedBlock b;
    int r = pinned_get(ctx, bytes, &b);
    if (r) return r;
    ow->host_pinned.push_back(b);
    if (bytes) SB_CUDA_CHECK(ctx, cudaMemcpyAsync(b.p, dev, bytes, cudaMemcpyDeviceToHost, st));
    *out = b.p;
    return SB_OK;
  };
  for (uint64_t c = 0; c < n_cols; ++c) {
    sb_column_out &o = outs[c];
    Owner *ow = owners[c];
    const ColDesc &cd = h_cols[c];
    const sb_out_buffers *ub = cd.n_nested > 1 ? nullptr : ubuf(c);
    SB_TRY(to_caller(ow, cd.values, o.values_bytes, &o.values, ub ? ub->values : nullptr, ub ? ub->values_cap : 0));
    SB_TRY(to_caller(ow, cd.offsets, o.offsets_bytes, &o.offsets, ub ? ub->offsets : nullptr, ub ? ub->offsets_cap : 0));
    SB_TRY(to_caller(ow, cd.validity, o.validity_bytes, reinterpret_cast<void **>(&o.validity), ub ? ub->validity : nullptr, ub ? ub->validity_cap : 0));
    for (int d = 0; d + 1 < cd.n_nested; ++d) {
      SB_TRY(to_caller(ow, cd.nest_off[d], (o.nested_len[d] + 1) * 8, reinterpret_cast<void **>(&o.nested_offsets[d])));
      SB_TRY(to_caller(ow, cd.nest_val[d], (o.nested_len[d] + 7) / 8, reinterpret_cast<void **>(&o.nested_validity[d])));
    }
    bytes_out += o.values_bytes + o.offsets_bytes + o.validity_bytes;
    for (int d = 0; d + 1 < cd.n_nested; ++d)
      bytes_out += (cd.nest_off[d] ? (o.nested_len[d] + 1) * 8 : 0) + (cd.nest_val[d] ? (o.nested_len[d] + 7) / 8 : 0);
  }
  for (void *d : d_inputs) cudaFreeAsync(d, st);
  // ---- everything is submitted: remember what the collection step needs
  DecodePending &pd = ctx->pending;
  pd.active = true;
  pd.n_cols = n_cols;
  pd.n_pages_total = n_pages_total;
  pd.n_items = n_items;
  pd.out_mem = out_mem;
  pd.outs = outs;
  pd.owners = owners;
  pd.col_pages.resize(n_cols);
  for (uint64_t c = 0; c < n_cols; ++c) pd.col_pages[c] = cols[c].n_pages;
  pd.off_status = off_status;
  pd.off_counters = off_counters;
  pd.any_fixed = any_fixed;
  pd.split = split;
  pd.bytes_in = bytes_in;
  pd.bytes_out = bytes_out;
  pd.t_host0 = t_host0;
  pd.t_planned = std::chrono::duration<double, std::milli>(t_planned.time_since_epoch()).count();
  pd.t_submitted = now_ms();
  return SB_OK;
#undef SB_TRY
#undef SB_TRY_CUDA
}

// Waits for a submitted decode call and hands its results over: per-page statuses, statistics, ownership.
int32_t sb_decode_finish_pending(sb_ctx *ctx) {
  DecodePending &pd = ctx->pending;
  if (!pd.active) return SB_OK;
  pd.active = false;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  sb_column_out *outs = pd.outs;
  const uint64_t n_cols = pd.n_cols;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    for (uint64_t c = 0; c < n_cols; ++c) outs[c]._owner = pd.owners[c];
    sb_release_columns(ctx, outs, n_cols);
    return fail(ctx, SB_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
  }
  const double t_synced = now_ms();
  if (pd.out_mem == SB_MEM_HOST) { // device copies no longer needed
    for (uint64_t c = 0; c < n_cols; ++c) {
      for (void *p : pd.owners[c]->dev) cudaFreeAsync(p, st);
      pd.owners[c]->dev.clear();
    }
  }
  if (pd.n_items) {
    cudaEventElapsedTime(&ctx->stats.device_ms, ctx->ev0, ctx->ev1);
    cudaEventElapsedTime(&ctx->stats.main_kernel_ms, ctx->ev_m0, ctx->ev_m1);
    if (pd.any_fixed) cudaEventElapsedTime(&ctx->stats.lz4_kernel_ms, ctx->ev_lz0, ctx->ev_lz1);
    if (pd.split) cudaEventElapsedTime(&ctx->stats.light_kernel_ms, ctx->ev_l0, ctx->ev_l1);
  }
  int32_t first_err = SB_OK;
  const uint8_t *hT = static_cast<const uint8_t *>(ctx->h_tables);
  const int32_t *h_status = reinterpret_cast<const int32_t *>(hT + pd.off_status);
  const uint32_t *h_counters = reinterpret_cast<const uint32_t *>(hT + pd.off_counters);
  if (pd.n_items) {
    for (int i = 0; i < 32; ++i) ctx->stats.codec_pages[i] = h_counters[5 + i];
    std::memcpy(&ctx->stats.lz4_bytes, h_counters + 38, 8);
  }
  uint64_t pi = 0;
  for (uint64_t c = 0; c < n_cols; ++c) {
    sb_column_out &o = outs[c];
    o._owner = pd.owners[c];
    int32_t *ps = static_cast<int32_t *>(std::malloc(sizeof(int32_t) * std::max<uint64_t>(1, pd.col_pages[c])));
    pd.owners[c]->host_malloc.push_back(ps);
    o.page_status = ps;
    for (uint64_t p = 0; p < pd.col_pages[c]; ++p, ++pi) {
      ps[p] = h_status[pi];
      if (ps[p] != SB_OK && first_err == SB_OK) {
        first_err = ps[p];
        ctx->err = "page " + std::to_string(p) + " of column " + std::to_string(c) + " failed with status " + std::to_string(ps[p]);
      }
    }
  }
  ctx->stats.host_ms = float(now_ms() - pd.t_host0);
  static const bool dbg_timing = std::getenv("SB_TIMING") != nullptr; // diagnostic: host phases of the call on stderr
  if (dbg_timing)
    std::fprintf(stderr, "[sb timing] plan %.3f  submit %.3f  wait %.3f  finish %.3f  (device %.3f) ms\n", pd.t_planned - pd.t_host0,
                 pd.t_submitted - pd.t_planned, t_synced - pd.t_submitted, now_ms() - t_synced, ctx->stats.device_ms);
  ctx->stats.pages = pd.n_pages_total;
  ctx->stats.bytes_in = pd.bytes_in;
  ctx->stats.bytes_out = pd.bytes_out;
  pd.owners.clear();
  return first_err;
}

extern "C" {

int32_t sb_decode_columns(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem, sb_column_out *outs) {
  const int32_t rc = decode_submit(ctx, cols, n_cols, out_mem, outs, nullptr, nullptr);
  return rc != SB_OK ? rc : sb_decode_finish_pending(ctx);
}

int32_t sb_decode_columns_into(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem, const sb_out_buffers *bufs,
                               sb_column_out *outs) {
  const int32_t rc = decode_submit(ctx, cols, n_cols, out_mem, outs, bufs, nullptr);
  return rc != SB_OK ? rc : sb_decode_finish_pending(ctx);
}

int32_t sb_decode_columns_async(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem, const sb_out_buffers *bufs,
                                sb_column_out *outs) {
  return decode_submit(ctx, cols, n_cols, out_mem, outs, bufs, nullptr);
}

int32_t sb_decode_wait(sb_ctx *ctx) {
  if (!ctx) return SB_CUDA;
  return sb_decode_finish_pending(ctx);
}

int32_t sb_decode_ready(sb_ctx *ctx) {
  if (!ctx) return SB_CUDA;
  if (!ctx->pending.active) return 1;
  cudaSetDevice(ctx->device);
  return cudaStreamQuery(ctx->stream) == cudaSuccess ? 1 : 0;
}

int32_t sb_plan_columns(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, sb_column_sizes *sizes) {
  if (!sizes) return ctx ? fail(ctx, SB_INVALID_ARG, "sizes is NULL") : SB_CUDA;
  return decode_submit(ctx, cols, n_cols, SB_MEM_DEVICE, nullptr, nullptr, sizes);
}

int32_t sb_decode_pages(sb_ctx *ctx, const sb_column_in *pages, uint64_t n_pages, int32_t out_mem, sb_column_out *outs) {
  if (!ctx) return SB_CUDA;
  for (uint64_t i = 0; i < n_pages; ++i)
    if (pages[i].n_pages != 1) return fail(ctx, SB_INVALID_ARG, "sb_decode_pages: every entry must hold exactly one page");
  return sb_decode_columns(ctx, pages, n_pages, out_mem, outs);
}

} // extern "C"
