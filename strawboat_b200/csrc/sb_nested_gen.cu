// sb_nested_gen.cu -- device-side generation of Dremel levels from Arrow nested buffers: the write half of SURVEY §8 f2.
//
// The reference gets the (rep, def) streams of a nested leaf from arrow2: to_nested() turns the array into one
// `Nested` descriptor per depth and write_rep_and_def iterates them (src/write/common.rs:66-68,
// src/write/serialize.rs:217-232).  Here the same levels are produced on the device, depth by depth, root -> leaf:
//
//   entries start as one per top-level row: (elem = row, rep = 0, def = 0, alive, slot)
//   Struct / Primitive depth : elementwise -- a valid element raises def by 1 when the depth is nullable, a null one
//                              stops def from rising any further (alive = 0); the entry keeps its child slot
//   List depth               : a null list (def stays) or an empty list (def + nullable) ends the entry: no slot
//                              below; a list of k items raises def by nullable + 1 and EXPANDS into k entries, the
//                              first keeping the parent's rep, the others rep = this list's repetition level
//                              (counts -> exclusive scan -> each new entry binary-searches its parent)
//
// Level semantics are the ones pinned against pyarrow's Parquet writer (tests/test_dremel_pin.py).  Lists must be
// compact (a null list has no children; children are laid out in order), which is what to_leaves hands over.
#include <cstdio>
#include <new>

#include "sb_common.cuh"
#include "sb_host.h"

namespace sb {

struct GenEntry { // structure of arrays, one element per level entry
  uint32_t *elem; // element index at the current depth
  uint32_t *rep, *def;
  uint8_t *flags; // bit 0 alive (def still rising), bit 1 slot (the entry reaches an element of the next depth)
};

__device__ __forceinline__ bool bit_at(const uint8_t *bm, uint64_t i) { return !bm || ((bm[i >> 3] >> (i & 7)) & 1); }
__device__ __forceinline__ int64_t off_at(const void *o, int w, uint64_t i) {
  return w == 4 ? int64_t(static_cast<const int32_t *>(o)[i]) : static_cast<const int64_t *>(o)[i];
}

__global__ void gen_init_kernel(GenEntry e, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  e.elem[i] = i;
  e.rep[i] = 0;
  e.def[i] = 0;
  e.flags[i] = 3;
}
// Struct / Primitive depth
__global__ void gen_plain_kernel(GenEntry e, uint32_t n, const uint8_t *validity, int nullable) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t f = e.flags[i];
  if (!(f & 1) || !nullable) return;
  if ((f & 2) && bit_at(validity, e.elem[i])) e.def[i] += 1;
  else e.flags[i] = f & ~1;
}
// List depth, step 1: entries each element expands to
__global__ void gen_list_count_kernel(GenEntry e, uint32_t n, const void *offsets, int ow, const uint8_t *validity, int nullable, uint32_t *cnt) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t f = e.flags[i];
  uint32_t c = 1;
  if (f & 2) {
    const uint32_t el = e.elem[i];
    const bool valid = !nullable || bit_at(validity, el);
    const int64_t len = off_at(offsets, ow, el + 1) - off_at(offsets, ow, el);
    if (f & 1) {
      if (!valid) f &= ~1;            // null list: def stays
      else if (nullable) e.def[i] += 1;
    }
    if (!valid || len <= 0) {
      f &= ~3;                         // the entry ends here: no slot below, def frozen (an empty list keeps def + nullable)
    } else {
      if (f & 1) e.def[i] += 1;        // the repeated level itself
      c = uint32_t(len);
    }
    e.flags[i] = f;
  }
  cnt[i] = c;
}
// exclusive scan of u32 counts into u64 positions: per-block sums, scan of the sums by one CTA, per-block scan
constexpr uint32_t kScanBlock = 1024;
__global__ void __launch_bounds__(256) scan_sums_kernel(const uint32_t *__restrict__ v, uint32_t n, uint64_t *sums) {
  __shared__ unsigned long long s;
  if (threadIdx.x == 0) s = 0;
  __syncthreads();
  const uint32_t b0 = blockIdx.x * kScanBlock;
  uint32_t t = 0;
  for (uint32_t i = threadIdx.x; i < kScanBlock && b0 + i < n; i += 256) t += v[b0 + i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s, (unsigned long long)t);
  __syncthreads();
  if (threadIdx.x == 0) sums[blockIdx.x] = s;
}
__global__ void scan_top_kernel(uint64_t *sums, uint32_t nb, uint64_t *total) { // one thread: nb = entries / 1024
  uint64_t run = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    const uint64_t x = sums[b];
    sums[b] = run;
    run += x;
  }
  *total = run;
}
__global__ void __launch_bounds__(256) scan_final_kernel(const uint32_t *__restrict__ v, uint32_t n, const uint64_t *__restrict__ sums, uint64_t *pos) {
  __shared__ uint32_t ws[9];
  const uint32_t b0 = blockIdx.x * kScanBlock, i0 = b0 + threadIdx.x * 4;
  uint32_t x[4], t = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    x[j] = i0 + j < n ? v[i0 + j] : 0u;
    t += x[j];
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(t);
  if (lane == 31) ws[warp] = inc;
  __syncthreads();
  uint32_t base = 0;
  for (uint32_t w = 0; w < warp; ++w) base += ws[w];
  uint64_t p = sums[blockIdx.x] + base + inc - t;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (i0 + j < n) pos[i0 + j] = p;
    p += x[j];
  }
}
// List depth, step 2: new entry j belongs to the parent p with pos[p] <= j < pos[p + 1]
__global__ void gen_list_expand_kernel(GenEntry src, uint32_t n_src, const uint64_t *__restrict__ pos, GenEntry dst, uint64_t n_dst,
                                       const void *offsets, int ow, uint32_t list_rep) {
  const uint64_t j = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= n_dst) return;
  uint32_t lo = 0, hi = n_src; // largest p with pos[p] <= j
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (pos[mid] <= j) lo = mid;
    else hi = mid;
  }
  const uint32_t p = lo;
  const uint8_t f = src.flags[p];
  const uint32_t k = uint32_t(j - pos[p]);
  dst.flags[j] = f;
  dst.def[j] = src.def[p];
  dst.rep[j] = k == 0 ? src.rep[p] : list_rep;
  dst.elem[j] = (f & 2) ? uint32_t(off_at(offsets, ow, src.elem[p]) + k) : 0u;
}
__global__ void gen_slots_kernel(GenEntry e, uint64_t n, unsigned long long *slots) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  uint32_t c = (i < n && (e.flags[i] & 2)) ? 1u : 0u;
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(slots, (unsigned long long)c);
}

} // namespace sb

using namespace sb;

extern "C" {

void sb_free_device(sb_ctx *ctx, void *p) {
  if (!ctx || !p) return;
  cudaSetDevice(ctx->device);
  cudaFreeAsync(p, ctx->stream);
}

int32_t sb_nested_levels(sb_ctx *ctx, const sb_nested_level *path, int32_t n_depths, int32_t mem, uint32_t **rep_out, uint32_t **def_out,
                         uint64_t *n_levels, uint64_t *n_slots) {
  if (!ctx) return SB_CUDA;
  if (!path || n_depths < 1 || n_depths > SB_MAX_NESTED || !rep_out || !def_out || !n_levels || !n_slots)
    return fail(ctx, SB_INVALID_ARG, "bad arguments");
  if (path[n_depths - 1].kind != SB_N_PRIMITIVE) return fail(ctx, SB_INVALID_ARG, "the last nested entry must be the primitive leaf");
  SB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  *rep_out = *def_out = nullptr;
  *n_levels = *n_slots = 0;
  const uint64_t rows = path[0].length;
  if (rows > 0xfffffff0ull) return fail(ctx, SB_OUT_OF_SPEC, "more than 2^32 top-level rows");
  std::vector<void *> tmp;
  auto cleanup = [&]() {
    for (void *p : tmp) cudaFreeAsync(p, st);
    tmp.clear();
  };
#define SB_NTRY(call)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      cleanup();                                                                      \
      return fail(ctx, SB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                 \
  } while (0)
  auto upload = [&](const void *src, uint64_t bytes, const void **dst) -> cudaError_t {
    *dst = src;
    if (!src || mem == SB_MEM_DEVICE || bytes == 0) return cudaSuccess;
    void *d = nullptr;
    cudaError_t e = cudaMallocAsync(&d, bytes + 16, st);
    if (e != cudaSuccess) return e;
    tmp.push_back(d);
    *dst = d;
    return cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st);
  };
  auto alloc_entries = [&](uint64_t n, GenEntry *e) -> cudaError_t {
    uint8_t *base = nullptr;
    cudaError_t er = cudaMallocAsync(reinterpret_cast<void **>(&base), 13 * n + 64, st);
    if (er != cudaSuccess) return er;
    e->elem = reinterpret_cast<uint32_t *>(base);
    e->rep = e->elem + n;
    e->def = e->rep + n;
    e->flags = reinterpret_cast<uint8_t *>(e->def + n);
    return cudaSuccess;
  };
  uint64_t n = rows;
  GenEntry cur{};
  SB_NTRY(alloc_entries(std::max<uint64_t>(n, 1), &cur));
  if (n) gen_init_kernel<<<uint32_t((n + 255) / 256), 256, 0, st>>>(cur, uint32_t(n));
  uint32_t rep_level = 0;
  for (int d = 0; d < n_depths; ++d) {
    const sb_nested_level &L = path[d];
    const void *d_val = nullptr, *d_off = nullptr;
    SB_NTRY(upload(L.validity, (L.length + 7) / 8, &d_val));
    if (L.kind == SB_N_LIST) {
      if (L.offset_width != 4 && L.offset_width != 8) {
        cleanup();
        cudaFreeAsync(cur.elem, st);
        return fail(ctx, SB_INVALID_ARG, "list offsets must be 4 or 8 bytes wide");
      }
      ++rep_level;
      SB_NTRY(upload(L.offsets, (L.length + 1) * uint64_t(L.offset_width), &d_off));
      if (n == 0) continue;
      uint32_t *cnt = nullptr;
      uint64_t *pos = nullptr, *sums = nullptr, *d_total = nullptr;
      const uint32_t nb = uint32_t((n + kScanBlock - 1) / kScanBlock);
      SB_NTRY(cudaMallocAsync(reinterpret_cast<void **>(&cnt), 4 * n + 16, st));
      tmp.push_back(cnt);
      SB_NTRY(cudaMallocAsync(reinterpret_cast<void **>(&pos), 8 * (n + 1) + 8 * (nb + 2), st));
      tmp.push_back(pos);
      sums = pos + n + 1;
      d_total = sums + nb;
      gen_list_count_kernel<<<uint32_t((n + 255) / 256), 256, 0, st>>>(cur, uint32_t(n), d_off, L.offset_width, static_cast<const uint8_t *>(d_val),
                                                                       L.nullable, cnt);
      scan_sums_kernel<<<nb, 256, 0, st>>>(cnt, uint32_t(n), sums);
      scan_top_kernel<<<1, 1, 0, st>>>(sums, nb, d_total);
      scan_final_kernel<<<nb, 256, 0, st>>>(cnt, uint32_t(n), sums, pos);
      uint64_t total = 0;
      SB_NTRY(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
      SB_NTRY(cudaStreamSynchronize(st));
      if (total > 0xfffffff0ull) {
        cleanup();
        cudaFreeAsync(cur.elem, st);
        return fail(ctx, SB_OUT_OF_SPEC, "more than 2^32 level entries");
      }
      GenEntry nxt{};
      SB_NTRY(alloc_entries(std::max<uint64_t>(total, 1), &nxt));
      if (total)
        gen_list_expand_kernel<<<uint32_t((total + 255) / 256), 256, 0, st>>>(cur, uint32_t(n), pos, nxt, total, d_off, L.offset_width, rep_level);
      tmp.push_back(cur.elem);
      cur = nxt;
      n = total;
    } else if (n) {
      gen_plain_kernel<<<uint32_t((n + 255) / 256), 256, 0, st>>>(cur, uint32_t(n), static_cast<const uint8_t *>(d_val), L.nullable);
    }
  }
  // leaf slots = entries that reach the leaf depth; rep / def leave as two separate allocations the caller frees
  unsigned long long *d_slots = nullptr;
  SB_NTRY(cudaMallocAsync(reinterpret_cast<void **>(&d_slots), 8, st));
  tmp.push_back(d_slots);
  SB_NTRY(cudaMemsetAsync(d_slots, 0, 8, st));
  if (n) gen_slots_kernel<<<uint32_t((n + 255) / 256), 256, 0, st>>>(cur, n, d_slots);
  uint32_t *rep = nullptr, *def = nullptr;
  SB_NTRY(cudaMallocAsync(reinterpret_cast<void **>(&rep), 4 * std::max<uint64_t>(n, 1), st));
  if (cudaMallocAsync(reinterpret_cast<void **>(&def), 4 * std::max<uint64_t>(n, 1), st) != cudaSuccess) {
    cudaFreeAsync(rep, st);
    cleanup();
    return fail(ctx, SB_CUDA, "cudaMallocAsync(def levels)");
  }
  if (n) {
    cudaMemcpyAsync(rep, cur.rep, 4 * n, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(def, cur.def, 4 * n, cudaMemcpyDeviceToDevice, st);
  }
  unsigned long long slots = 0;
  cudaMemcpyAsync(&slots, d_slots, 8, cudaMemcpyDeviceToHost, st);
  tmp.push_back(cur.elem);
  cudaError_t e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
    cudaFreeAsync(rep, st);
    cudaFreeAsync(def, st);
    return fail(ctx, SB_CUDA, "level generation kernels failed");
  }
  *rep_out = rep;
  *def_out = def;
  *n_levels = n;
  *n_slots = slots;
  return SB_OK;
#undef SB_NTRY
}

} // extern "C"
