// sb_common.cuh -- device-side building blocks shared by the decode and encode kernels.
//
// Execution model: one CTA of SB_NT = 128 threads (one warp-group) works on
// one page at a time.  Page bytes arrive in shared memory through the TMA engine
// (cp.async.bulk, 1-D, mbarrier completion); everything irregular (byte-unaligned fields,
// bit unpacking, scans, gathers) happens in shared memory / registers; HBM only sees
// 16-byte aligned bulk reads and coalesced 16-byte vector stores.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/strawboat_b200.h"

#define SB_NT 128
#define SB_NWARP (SB_NT / 32)

namespace sb {

// ------------------------------------------------------------------------------------
// descriptors shared between host and device
// ------------------------------------------------------------------------------------
struct PageDesc {
  const uint8_t *src;  // device pointer to the first byte of the page
  uint32_t len;        // PageMeta.length
  uint32_t num_values; // PageMeta.num_values (rows, or level entries for nested leaves)
  uint32_t col;        // column index
  uint32_t ordinal;    // page index inside its column
  uint64_t out_elem;   // first output element (row / leaf slot) of this page in its column
  uint64_t out_byte;   // binary: first output value byte of this page in its column
  uint32_t aux;        // index into the PageAux array (binary / nested pages), else 0xffffffff
  uint32_t last;       // 1 for the last page of its column
  uint64_t tab_off;    // binary: first BinEntry of this page in the entry table
};

// Plan-pass results of pages whose output size is data dependent (binary value bytes, nested
// per-depth entry counts).  Pass 0 writes value_bytes / cnt; the host scans them per column
// into PageDesc.out_elem / out_byte and base[] before pass 1.
struct PageAux {
  uint64_t value_bytes;
  uint32_t val_pos; // binary Basic / None pages: page position of the plain value bytes (0 = not tileable)
  uint32_t n_ent;   // binary Dict / Freq pages: BinEntry records the plan pass wrote for this page
  uint32_t failed;  // plan pass rejected the page (status holds the reason): pass 1 must not touch it
  uint32_t pad;
  uint32_t cnt[SB_MAX_NESTED];
  uint64_t base[SB_MAX_NESTED];
};

struct ColDesc {
  int32_t type;
  int32_t nullable; // flat: the leaf field is nullable (validity section present)
  int32_t W;        // value width in bytes (primitives), offset width (binary)
  int32_t is_float;
  uint8_t *values;
  uint8_t *offsets;
  uint8_t *validity;
  uint64_t length;  // total elements
  // nested leaves (n_nested > 1): InitNested root -> leaf and the derived level thresholds
  int32_t n_nested;
  uint8_t kind[SB_MAX_NESTED], nnull[SB_MAX_NESTED];
  uint8_t cum_sum[SB_MAX_NESTED + 1], cum_rep[SB_MAX_NESTED + 1];
  int64_t *nest_off[SB_MAX_NESTED];
  uint8_t *nest_val[SB_MAX_NESTED];
};

struct WorkItem {
  uint32_t page;
  uint32_t tile;
};

// ------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + TMA 1-D bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order prior generic-proxy accesses to shared memory before subsequent async-proxy (TMA) ones
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "SB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra SB_DONE;\n\t"
      "bra SB_WAIT;\n\t"
      "SB_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; gsrc and smem_dst 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------
// unaligned little-endian loads through aligned accesses (generic address space: the
// source is either the TMA-staged page in shared memory or, for oversized pages, global)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return *p; }
__device__ __forceinline__ uint32_t ld_u16u(const uint8_t *p) {
  if (((uintptr_t)p & 1) == 0) return *reinterpret_cast<const uint16_t *>(p);
  return uint32_t(p[0]) | (uint32_t(p[1]) << 8);
}
__device__ __forceinline__ uint32_t ld_u32u(const uint8_t *p) {
  uint32_t a = uint32_t((uintptr_t)p & 3);
  const uint32_t *q = reinterpret_cast<const uint32_t *>(p - a);
  uint32_t lo = q[0];
  if (a == 0) return lo;
  return __funnelshift_r(lo, q[1], a * 8);
}
__device__ __forceinline__ uint64_t ld_u64u(const uint8_t *p) {
  uint32_t a = uint32_t((uintptr_t)p & 3);
  const uint32_t *q = reinterpret_cast<const uint32_t *>(p - a);
  uint32_t w0 = q[0], w1 = q[1];
  if (a == 0) return uint64_t(w0) | (uint64_t(w1) << 32);
  uint32_t w2 = q[2];
  return uint64_t(__funnelshift_r(w0, w1, a * 8)) | (uint64_t(__funnelshift_r(w1, w2, a * 8)) << 32);
}
// 16 bytes from an arbitrary byte address: two aligned 16-byte loads + funnel shifts.
// Only touches 16-byte granules that contain at least one requested byte.
__device__ __forceinline__ uint4 ld_u128u(const uint8_t *p) {
  uint32_t a = uint32_t((uintptr_t)p & 15);
  const uint4 *q = reinterpret_cast<const uint4 *>(p - a);
  uint4 q0 = q[0];
  if (a == 0) return q0;
  uint4 q1 = q[1];
  uint32_t sh = (a & 3) * 8;
  uint4 r;
  switch (a >> 2) {
  case 0:
    r.x = __funnelshift_r(q0.x, q0.y, sh), r.y = __funnelshift_r(q0.y, q0.z, sh);
    r.z = __funnelshift_r(q0.z, q0.w, sh), r.w = __funnelshift_r(q0.w, q1.x, sh);
    break;
  case 1:
    r.x = __funnelshift_r(q0.y, q0.z, sh), r.y = __funnelshift_r(q0.z, q0.w, sh);
    r.z = __funnelshift_r(q0.w, q1.x, sh), r.w = __funnelshift_r(q1.x, q1.y, sh);
    break;
  case 2:
    r.x = __funnelshift_r(q0.z, q0.w, sh), r.y = __funnelshift_r(q0.w, q1.x, sh);
    r.z = __funnelshift_r(q1.x, q1.y, sh), r.w = __funnelshift_r(q1.y, q1.z, sh);
    break;
  default:
    r.x = __funnelshift_r(q0.w, q1.x, sh), r.y = __funnelshift_r(q1.x, q1.y, sh);
    r.z = __funnelshift_r(q1.y, q1.z, sh), r.w = __funnelshift_r(q1.z, q1.w, sh);
    break;
  }
  return r;
}

template <int W> struct Elem;
template <> struct Elem<1> { using T = uint8_t; };
template <> struct Elem<2> { using T = uint16_t; };
template <> struct Elem<4> { using T = uint32_t; };
template <> struct Elem<8> { using T = uint64_t; };
// i128 / i256 (Decimal128 / Decimal256 storage, src/util/mod.rs:77-78): moved as 16-byte vectors, never interpreted
struct __align__(16) U128 {
  uint4 v;
};
struct __align__(16) U256 {
  uint4 lo, hi;
};
template <> struct Elem<16> { using T = U128; };
template <> struct Elem<32> { using T = U256; };

template <int W> __device__ __forceinline__ typename Elem<W>::T ld_elem_u(const uint8_t *p) {
  if constexpr (W == 1) return uint8_t(*p);
  else if constexpr (W == 2) return uint16_t(ld_u16u(p));
  else if constexpr (W == 4) return ld_u32u(p);
  else if constexpr (W == 8) return ld_u64u(p);
  else if constexpr (W == 16) return U128{ld_u128u(p)};
  else return U256{ld_u128u(p), ld_u128u(p + 16)};
}

// ------------------------------------------------------------------------------------
// warp / block scans
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= uint32_t(d)) v += t;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// saturating add: sums of run lengths are clamped at `cap` (monotone, associative on
// non-negative inputs) so hostile run lengths cannot wrap the 32-bit positions.
__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b, uint32_t cap) {
  uint64_t s = uint64_t(a) + uint64_t(b);
  return s > cap ? cap : uint32_t(s);
}
__device__ __forceinline__ uint32_t warp_incl_scan_sat(uint32_t v, uint32_t cap) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= uint32_t(d)) v = sat_add(v, t, cap);
  }
  return v;
}

// Block-wide exclusive scan over SB_NT threads (one value per thread).  `ws` is a
// shared array of SB_NWARP+1 words.  Returns the exclusive prefix; *total = block sum.
// Contains two __syncthreads().
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *ws, uint32_t *total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v);
  __syncthreads(); // ws may still be read from a previous call
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SB_NWARP; ++w) {
    uint32_t x = ws[w];
    if (uint32_t(w) < warp) base += x;
    tot += x;
  }
  *total = tot;
  return base + inc - v;
}
__device__ __forceinline__ uint32_t block_excl_scan_sat(uint32_t v, uint32_t cap, uint32_t *ws, uint32_t *total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan_sat(v, cap);
  uint32_t exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) exc = 0;
  __syncthreads();
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SB_NWARP; ++w) {
    uint32_t x = ws[w];
    if (uint32_t(w) < warp) base = sat_add(base, x, cap);
    tot = sat_add(tot, x, cap);
  }
  *total = tot;
  return sat_add(base, exc, cap);
}

// ------------------------------------------------------------------------------------
// scratch arena: a bump allocator over the unused part of the CTA's shared memory, falling
// back to the CTA slot's global scratch (L2-resident) when an intermediate does not fit.
// All threads of the CTA hold identical copies (allocation is a uniform computation).
// ------------------------------------------------------------------------------------
struct Arena {
  uint8_t *s_cur, *s_end; // shared
  uint8_t *g_cur, *g_end; // global
  __device__ __forceinline__ void *alloc(uint64_t bytes) {
    bytes = (bytes + 15) & ~uint64_t(15);
    if (uint64_t(s_end - s_cur) >= bytes) {
      void *p = s_cur;
      s_cur += bytes;
      return p;
    }
    if (uint64_t(g_end - g_cur) >= bytes) {
      void *p = g_cur;
      g_cur += bytes;
      return p;
    }
    return nullptr;
  }
  // global-only allocation (buffers that other memory paths read with ld.global)
  __device__ __forceinline__ void *alloc_global(uint64_t bytes) {
    bytes = (bytes + 15) & ~uint64_t(15);
    if (uint64_t(g_end - g_cur) >= bytes) {
      void *p = g_cur;
      g_cur += bytes;
      return p;
    }
    return nullptr;
  }
  // shared-only allocation (small hot tables)
  __device__ __forceinline__ void *alloc_shared(uint64_t bytes) {
    bytes = (bytes + 15) & ~uint64_t(15);
    if (uint64_t(s_end - s_cur) >= bytes) {
      void *p = s_cur;
      s_cur += bytes;
      return p;
    }
    return nullptr;
  }
};

struct Dctx {
  Arena ar;
  int *err;          // shared: first error of the page (0 = ok)
  uint32_t *ws;      // shared: SB_NWARP+1 words of scan workspace
  int *bcast;        // shared: 4 ints for CTA-wide broadcasts
  const uint8_t *page_s = nullptr; // first byte of the page as the decoders see it (shared when staged)
  const uint8_t *page_g = nullptr; // the same byte in global memory (source of cp.async streams)
  uint64_t *rbar = nullptr;        // shared: SB_RING_STAGES mbarriers of the streaming ring (nullptr = no ring)
  uint32_t rphase = 0;             // their phase bits (CTA-uniform, carried across pages)
  __device__ __forceinline__ void flag(int code) { atomicCAS(err, 0, code); }
};

// ------------------------------------------------------------------------------------
// byte copy, arbitrary source alignment -> destination (basic.rs:67-70 `None`)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_bytes(uint8_t *dst, const uint8_t *src, uint64_t nbytes) {
  const uint32_t tid = threadIdx.x;
  uint64_t head = min(nbytes, uint64_t((16 - (uintptr_t(dst) & 15)) & 15));
  for (uint64_t i = tid; i < head; i += SB_NT) dst[i] = src[i];
  dst += head;
  src += head;
  nbytes -= head;
  uint64_t nvec = nbytes >> 4;
  if ((uintptr_t(src) & 15) == 0) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    uint64_t v = tid;
    for (; v + 3 * SB_NT < nvec; v += 4 * SB_NT) { // 4 independent 16-byte loads in flight
      uint4 a = s[v], b = s[v + SB_NT], c = s[v + 2 * SB_NT], e = s[v + 3 * SB_NT];
      d[v] = a, d[v + SB_NT] = b, d[v + 2 * SB_NT] = c, d[v + 3 * SB_NT] = e;
    }
    for (; v < nvec; v += SB_NT) d[v] = s[v];
  } else {
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    uint64_t v = tid;
    for (; v + 3 * SB_NT < nvec; v += 4 * SB_NT) { // 8 aligned 16-byte loads in flight per thread
      uint4 a = ld_u128u(src + (v << 4)), b = ld_u128u(src + ((v + SB_NT) << 4));
      uint4 c = ld_u128u(src + ((v + 2 * SB_NT) << 4)), e = ld_u128u(src + ((v + 3 * SB_NT) << 4));
      d[v] = a, d[v + SB_NT] = b, d[v + 2 * SB_NT] = c, d[v + 3 * SB_NT] = e;
    }
    for (; v < nvec; v += SB_NT) d[v] = ld_u128u(src + (v << 4));
  }
  for (uint64_t i = (nvec << 4) + tid; i < nbytes; i += SB_NT) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------
// Streaming ring: pages that do not fit the staging buffer are pulled through shared memory
// in SB_RING_CHUNK pieces by the TMA engine (SB_RING_STAGES bulk loads in flight per CTA),
// realigned with funnel shifts and written with aligned 16-byte stores.  `f(vec, index)`
// transforms every 16-byte vector (identity for copies, offset rebase for binary pages).
// `d` is 16-byte aligned, `src` any byte address; nvec vectors are produced.
// ------------------------------------------------------------------------------------
#define SB_RING_STAGES 4
#define SB_RING_CHUNK 8192

template <class F>
__device__ __forceinline__ void stream_vec(Dctx &cx, uint4 *d, const uint8_t *src, uint64_t nvec, F f, bool fresh = false) {
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t VPC = SB_RING_CHUNK / 16, SLOT = SB_RING_CHUNK + 16;
  Arena mark = cx.ar;
  uint8_t *ring = nullptr;
  if (cx.rbar != nullptr && nvec >= 2 * VPC && __isGlobal(src))
    ring = static_cast<uint8_t *>(cx.ar.alloc_shared(SB_RING_STAGES * SLOT));
  if (!ring) { // small, or the source already sits in shared memory
    uint64_t v = tid;
    for (; v + 3 * SB_NT < nvec; v += 4 * SB_NT) {
      uint4 a = ld_u128u(src + (v << 4)), b = ld_u128u(src + ((v + SB_NT) << 4));
      uint4 c = ld_u128u(src + ((v + 2 * SB_NT) << 4)), e = ld_u128u(src + ((v + 3 * SB_NT) << 4));
      d[v] = f(a, v), d[v + SB_NT] = f(b, v + SB_NT), d[v + 2 * SB_NT] = f(c, v + 2 * SB_NT), d[v + 3 * SB_NT] = f(e, v + 3 * SB_NT);
    }
    for (; v < nvec; v += SB_NT) d[v] = f(ld_u128u(src + (v << 4)), v);
    return;
  }
  const uint32_t a = uint32_t(uintptr_t(src) & 15);
  const uint8_t *g = src - a;
  const uint64_t nch = (nvec + VPC - 1) / VPC;
  auto issue = [&](uint64_t c) { // thread 0: chunk c -> slot c % STAGES
    const uint32_t s = uint32_t(c % SB_RING_STAGES);
    const uint64_t v0 = c * VPC;
    const uint32_t nv = uint32_t(min(uint64_t(VPC), nvec - v0));
    const uint32_t bytes = nv * 16 + (a ? 16u : 0u); // only granules that hold requested bytes
    mbar_expect_tx(cx.rbar + s, bytes);
    tma_load_1d(ring + s * SLOT, g + (v0 << 4), bytes, cx.rbar + s);
  };
  // earlier generic accesses to this part of the arena are done (`fresh`: the work item has not touched
  // the arena yet, and the barrier at the end of the previous item covers everything before it)
  if (!fresh) __syncthreads();
  if (tid == 0) {
    fence_proxy_async();
    for (uint64_t c = 0; c < nch && c < SB_RING_STAGES; ++c) issue(c);
  }
  for (uint64_t c = 0; c < nch; ++c) {
    const uint32_t s = uint32_t(c % SB_RING_STAGES);
    mbar_wait(cx.rbar + s, (cx.rphase >> s) & 1u);
    cx.rphase ^= 1u << s;
    const uint8_t *sp = ring + s * SLOT + a;
    const uint64_t v0 = c * VPC;
    const uint32_t nv = uint32_t(min(uint64_t(VPC), nvec - v0));
    if (nv == VPC) {
      static_assert(VPC == 4 * SB_NT, "one chunk = four vectors per thread");
      uint4 x0 = ld_u128u(sp + (tid << 4)), x1 = ld_u128u(sp + ((tid + SB_NT) << 4));
      uint4 x2 = ld_u128u(sp + ((tid + 2 * SB_NT) << 4)), x3 = ld_u128u(sp + ((tid + 3 * SB_NT) << 4));
      uint4 *dd = d + v0 + tid;
      dd[0] = f(x0, v0 + tid), dd[SB_NT] = f(x1, v0 + tid + SB_NT);
      dd[2 * SB_NT] = f(x2, v0 + tid + 2 * SB_NT), dd[3 * SB_NT] = f(x3, v0 + tid + 3 * SB_NT);
    } else {
      for (uint32_t i = tid; i < nv; i += SB_NT) d[v0 + i] = f(ld_u128u(sp + (i << 4)), v0 + i);
    }
    __syncthreads(); // slot s fully read
    if (tid == 0 && c + SB_RING_STAGES < nch) {
      fence_proxy_async();
      issue(c + SB_RING_STAGES);
    }
  }
  cx.ar = mark;
}

// copy_bytes for sources that may be large and in global memory (unstaged pages)
__device__ __forceinline__ void stream_copy(Dctx &cx, uint8_t *dst, const uint8_t *src, uint64_t nbytes, bool fresh = false) {
  const uint32_t tid = threadIdx.x;
  uint64_t head = min(nbytes, uint64_t((16 - (uintptr_t(dst) & 15)) & 15));
  for (uint64_t i = tid; i < head; i += SB_NT) dst[i] = src[i];
  dst += head, src += head, nbytes -= head;
  const uint64_t nvec = nbytes >> 4;
  stream_vec(cx, reinterpret_cast<uint4 *>(dst), src, nvec, [](uint4 v, uint64_t) { return v; }, fresh);
  for (uint64_t i = (nvec << 4) + tid; i < nbytes; i += SB_NT) dst[i] = src[i];
}

} // namespace sb
