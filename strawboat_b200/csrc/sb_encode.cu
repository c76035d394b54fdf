// sb_encode.cu -- encode half of libstrawboat_b200.so: sb_encode_columns replaces the page loop of
// NativeWriter::encode_chunk for flat leaves (src/write/common.rs:71-115 -> write::write,
// src/write/serialize.rs:36-132 -> compress_integer / compress_double / compress_binary /
// compress_boolean).
//
//   host : split every leaf into pages of max_page_size rows, give each page a slab sized by
//          its worst-case encoding
//   E*   : sb_encode_kernel   persistent grid, one CTA per page: [validity section] + stats +
//          chooser + codec into the slab, page length out
//   host : exclusive scan of the page lengths per column (PageMeta.length, column body size)
//   E12  : sb_gather_kernel   slabs -> contiguous column bodies (what the file holds between
//          ColumnMeta.offset and the next column)
// No CPU fallback: without a CUDA device every entry point returns SB_CUDA.
#include <cstdio>
#include <new>

#include "sb_common.cuh"
#include "sb_encode.cuh"
#include "sb_host.h"

namespace sb {

struct EncCol {
  int32_t type, nullable, W, tclass;
  const uint8_t *values, *offsets, *validity;
  uint64_t length, values_bytes;
  // nested leaves: the column's Dremel levels and the level thresholds of every depth
  const uint32_t *rep, *def;
  uint64_t n_levels, rows;
  int32_t n_nested, w_rep, w_def, pad;
  uint8_t kind[SB_MAX_NESTED], nnull[SB_MAX_NESTED];
  uint8_t cum_sum[SB_MAX_NESTED + 1], cum_rep[SB_MAX_NESTED + 1];
};
struct EncPage {
  uint64_t row0, slab_off, dst_off; // row0 = first row (flat) / first leaf slot (nested)
  uint32_t col, n, ordinal, rows;   // n = rows (flat) / leaf slots (nested); rows = top-level rows of a nested page
  uint64_t lv0;                     // nested: first level entry of the page
  uint32_t n_lv, pad;               // nested: level entries of the page (PageMeta.num_values)
};

// ------------------------------------------------------------------------------------
// Nested leaves: where do pages start?  Pages are cut by top-level rows (write/common.rs:79-86,
// slice_parquet_array), a row starts at every level entry with rep == 0, and an entry owns a
// leaf slot when the leaf depth pushes (same rule as the reader, read_basic.rs:119-151).
//   N1 sb_level_count_kernel  : (rows, leaf slots) of every block of kLvBlock entries
//   N2 sb_level_scan_kernel   : exclusive scan of the block counts, one CTA per column
//   N3 sb_level_bounds_kernel : page p -> first entry / first leaf slot of row p * page_rows
// ------------------------------------------------------------------------------------
constexpr uint32_t kLvBlock = 4096;
struct LvCol {       // one nested column of the call
  uint32_t col;      // index into EncCol
  uint32_t blk0;     // first block of this column in the block-count arrays
  uint32_t n_blk;
  uint32_t page0;    // first page of this column in the bounds arrays
  uint32_t n_pages;
  uint32_t page_rows;
};
__device__ __forceinline__ bool lv_slot(const EncCol &c, uint32_t rep, uint32_t def) {
  // leaf depth pushes <=> the entry carries a value slot (nest_entry of sb_nested.cuh, leaf bit only)
  bool is_required = false, pushed = false;
  for (int d = 0; d < c.n_nested; ++d) {
    const bool right = rep <= c.cum_rep[d] && def >= c.cum_sum[d];
    pushed = is_required || right;
    if (pushed) {
      const bool v = c.nnull[d] && def > c.cum_sum[d];
      is_required = c.kind[d] == SB_N_STRUCT && !v;
    }
  }
  return pushed;
}
__global__ void __launch_bounds__(256)
    sb_level_count_kernel(const EncCol *__restrict__ cols, const LvCol *__restrict__ lcols, uint32_t n_lcols, uint2 *blk) {
  // blockIdx.x = global block index; find its column (few nested columns per call)
  uint32_t lc = 0;
  while (lc + 1 < n_lcols && blockIdx.x >= lcols[lc + 1].blk0) ++lc;
  const LvCol L = lcols[lc];
  const EncCol &c = cols[L.col];
  const uint64_t e0 = uint64_t(blockIdx.x - L.blk0) * kLvBlock;
  uint32_t rows = 0, slots = 0;
  for (uint32_t i = threadIdx.x; i < kLvBlock && e0 + i < c.n_levels; i += 256) {
    const uint32_t rep = c.rep ? c.rep[e0 + i] : 0u, def = c.def ? c.def[e0 + i] : 0u;
    rows += rep == 0;
    slots += lv_slot(c, rep, def);
  }
  __shared__ uint32_t s_r, s_s;
  if (threadIdx.x == 0) s_r = s_s = 0;
  __syncthreads();
  rows = warp_sum(rows), slots = warp_sum(slots);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_r, rows);
    atomicAdd(&s_s, slots);
  }
  __syncthreads();
  if (threadIdx.x == 0) blk[blockIdx.x] = make_uint2(s_r, s_s);
}
// in place: blk[b] = counts before block b (64-bit running sums kept in two arrays)
__global__ void __launch_bounds__(SB_NT)
    sb_level_scan_kernel(const LvCol *__restrict__ lcols, const uint2 *__restrict__ blk, uint64_t *rows_before, uint64_t *slots_before,
                         uint64_t *totals /* 2 per nested column */) {
  const LvCol L = lcols[blockIdx.x];
  __shared__ uint32_t ws[SB_NWARP + 1];
  uint64_t run_r = 0, run_s = 0;
  for (uint32_t b0 = 0; b0 < L.n_blk; b0 += SB_NT) {
    const uint32_t b = b0 + threadIdx.x;
    const uint2 v = b < L.n_blk ? blk[L.blk0 + b] : make_uint2(0u, 0u);
    uint32_t tr, ts;
    const uint32_t pr = block_excl_scan(v.x, ws, &tr);
    const uint32_t ps = block_excl_scan(v.y, ws, &ts);
    if (b < L.n_blk) {
      rows_before[L.blk0 + b] = run_r + pr;
      slots_before[L.blk0 + b] = run_s + ps;
    }
    run_r += tr, run_s += ts;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    totals[2 * blockIdx.x] = run_r;
    totals[2 * blockIdx.x + 1] = run_s;
  }
}
// one CTA per page: first level entry and first leaf slot of top-level row p * page_rows
__global__ void __launch_bounds__(SB_NT)
    sb_level_bounds_kernel(const EncCol *__restrict__ cols, const LvCol *__restrict__ lcols, uint32_t n_lcols,
                           const uint64_t *__restrict__ rows_before, const uint64_t *__restrict__ slots_before,
                           uint64_t *bounds /* 2 per page: entry, slot */) {
  uint32_t lc = 0;
  while (lc + 1 < n_lcols && blockIdx.x >= lcols[lc + 1].page0) ++lc;
  const LvCol L = lcols[lc];
  const EncCol &c = cols[L.col];
  const uint64_t target = uint64_t(blockIdx.x - L.page0) * L.page_rows; // rows before the page
  // last block with rows_before <= target (uniform binary search)
  uint32_t lo = 0, hi = L.n_blk;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (rows_before[L.blk0 + mid] <= target) lo = mid;
    else hi = mid;
  }
  const uint64_t e0 = uint64_t(lo) * kLvBlock;
  const uint32_t need = uint32_t(target - rows_before[L.blk0 + lo]); // row starts to skip inside the block
  __shared__ uint32_t ws[SB_NWARP + 1];
  __shared__ unsigned long long s_entry, s_slot;
  if (threadIdx.x == 0) s_entry = ~0ull, s_slot = 0;
  __syncthreads();
  uint32_t run_r = 0, run_s = 0;
  constexpr uint32_t PER = kLvBlock / SB_NT; // consecutive entries per thread
  uint32_t my_r = 0, my_s = 0;
  uint8_t flags[PER];
  for (uint32_t j = 0; j < PER; ++j) {
    const uint64_t e = e0 + threadIdx.x * PER + j;
    uint32_t f = 0;
    if (e < c.n_levels) {
      const uint32_t rep = c.rep ? c.rep[e] : 0u, def = c.def ? c.def[e] : 0u;
      f = (rep == 0 ? 1u : 0u) | (lv_slot(c, rep, def) ? 2u : 0u);
    }
    flags[j] = uint8_t(f);
    my_r += f & 1u, my_s += f >> 1;
  }
  uint32_t tr, ts;
  run_r = block_excl_scan(my_r, ws, &tr);
  run_s = block_excl_scan(my_s, ws, &ts);
  for (uint32_t j = 0; j < PER; ++j) {
    if ((flags[j] & 1u) && run_r == need) { // the (need+1)-th row start of the block: exactly one thread
      s_entry = e0 + threadIdx.x * PER + j;
      s_slot = slots_before[L.blk0 + lo] + run_s;
    }
    run_r += flags[j] & 1u, run_s += flags[j] >> 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bounds[2 * blockIdx.x] = s_entry; // ~0 = the levels hold fewer rows than the caller said
    bounds[2 * blockIdx.x + 1] = s_slot;
  }
}

// [ULEB((ceil8(n) << 1) | 1)][ceil8(n) * w bytes]: one bit-packed run, values LSB-first at width w
// (arrow2 write_rep_and_def V2 -> parquet2 encode_u32, call site src/write/serialize.rs:225)
__device__ uint32_t enc_levels(const uint32_t *lv, uint32_t n, uint32_t w, uint8_t *out) {
  if (w == 0) return 0;
  const uint32_t groups = (n + 7) / 8;
  uint64_t header = (uint64_t(groups) << 1) | 1;
  uint32_t ul = 0;
  uint8_t ub[10];
  do {
    uint8_t b = header & 0x7f;
    header >>= 7;
    ub[ul++] = b | (header ? 0x80 : 0);
  } while (header);
  if (threadIdx.x == 0)
    for (uint32_t i = 0; i < ul; ++i) out[i] = ub[i];
  uint8_t *d = out + ul;
  const uint64_t mask = (1ull << w) - 1;
  for (uint32_t g = threadIdx.x; g < groups; g += SB_NT) {
    unsigned __int128 acc = 0; // 8 values of up to 16 bits
    for (uint32_t j = 0; j < 8; ++j) {
      const uint32_t e = 8 * g + j;
      const uint64_t v = (e < n && lv) ? (uint64_t(lv[e]) & mask) : 0ull;
      acc |= (unsigned __int128)v << (j * w);
    }
    for (uint32_t b = 0; b < w; ++b) d[uint64_t(g) * w + b] = uint8_t(acc >> (8 * b));
  }
  return ul + groups * w;
}

constexpr uint32_t kEncSmem = 56 * 1024; // 4 CTAs per SM: the 8192-slot distinct table of a page (32 + 16 KiB), or the LZ tables

__global__ void __launch_bounds__(SB_NT, 4)
    sb_encode_kernel(const EncPage *__restrict__ pages, const EncCol *__restrict__ cols, uint32_t n_pages, uint32_t *counter,
                     uint8_t *slab, uint8_t *scratch, uint64_t scratch_per_cta, uint32_t *page_len, int32_t *status, EOpts base,
                     uint32_t *codec_hist) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ int s_err;
  __shared__ uint32_t s_ws[SB_NWARP + 1];
  __shared__ int s_bcast[4];
  __shared__ uint32_t s_item;
  const uint32_t tid = threadIdx.x;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      s_item = atomicAdd(counter, 1u);
      s_err = 0;
    }
    __syncthreads();
    const uint32_t it = s_item;
    if (it >= n_pages) break;
    const EncPage pg = pages[it];
    const EncCol &col = cols[pg.col];
    Dctx cx;
    cx.err = &s_err;
    cx.ws = s_ws;
    cx.bcast = s_bcast;
    cx.ar.s_cur = dsm;
    cx.ar.s_end = dsm + kEncSmem;
    cx.ar.g_cur = scratch + uint64_t(blockIdx.x) * scratch_per_cta;
    cx.ar.g_end = cx.ar.g_cur + scratch_per_cta;
    EOpts o = base;
    o.seed = base.seed + pg.ordinal; // page p of a column samples with seed + p
    uint8_t *out = slab + pg.slab_off;
    const uint32_t n = pg.n;
    uint32_t pos = 0;
    const bool nested = col.n_nested > 1;
    if (nested) { // write_nested_validity (serialize.rs:217-232): [rows][rep_len][def_len][rep][def]
      const uint32_t rl = enc_levels(col.rep ? col.rep + pg.lv0 : nullptr, pg.n_lv, uint32_t(col.w_rep), out + 12);
      const uint32_t dl = enc_levels(col.def ? col.def + pg.lv0 : nullptr, pg.n_lv, uint32_t(col.w_def), out + 12 + rl);
      if (tid == 0) {
        st_le(out, pg.rows, 4);
        st_le(out + 4, rl, 4);
        st_le(out + 8, dl, 4);
      }
      pos = 12 + rl + dl;
    }
    if (col.type != SB_NULL) { // Null-typed column: empty page (serialize.rs:63)
      const Bits valid{col.validity, pg.row0};
      if (col.nullable && !nested) pos += enc_validity(valid, n, out);
      uint32_t used;
      if (col.type == SB_BOOL) {
        used = enc_boolean(cx, Bits{col.values, pg.row0}, valid, n, o, out + pos);
      } else if (col.type == SB_BINARY || col.type == SB_LARGE_BINARY) {
        used = enc_binary(cx, BinView{col.values, col.offsets + pg.row0 * uint64_t(col.W), col.W}, valid, n, col.values_bytes, o,
                          out + pos);
      } else if (col.W >= 16) { // i128 / i256
        used = enc_wide<0>(cx, col.values + pg.row0 * uint64_t(col.W), col.W, valid, n, o, out + pos);
      } else {
        used = enc_fixed<0>(cx, Vals{col.values + pg.row0 * uint64_t(col.W), col.W}, col.tclass, valid, n, o, out + pos);
      }
      if (used == kEncFail) {
        if (!s_err) cx.flag(SB_EXTERNAL);
        used = 0;
      } else if (tid == 0) {
        atomicAdd(codec_hist + (out[pos] & 31), 1u);
      }
      pos += used;
    }
    __syncthreads();
    if (tid == 0) {
      page_len[it] = s_err ? 0u : pos;
      if (s_err) status[it] = s_err;
    }
  }
}

// slabs -> contiguous column bodies
__global__ void __launch_bounds__(SB_NT)
    sb_gather_kernel(const EncPage *__restrict__ pages, uint32_t n_pages, const uint32_t *__restrict__ page_len,
                     const uint8_t *__restrict__ slab, uint8_t *const *__restrict__ col_dst, uint32_t *counter) {
  __shared__ uint32_t s_item;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1u);
    __syncthreads();
    const uint32_t it = s_item;
    if (it >= n_pages) break;
    const EncPage pg = pages[it];
    copy_bytes(col_dst[pg.col] + pg.dst_off, slab + pg.slab_off, page_len[it]);
  }
}

// offsets[row0 of every page] (and the final offset) of binary columns, for slab sizing
__global__ void sb_page_offsets_kernel(const EncPage *__restrict__ pages, const EncCol *__restrict__ cols, uint32_t n_pages,
                                       int64_t *out /* 2 per page: first, last */) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pages) return;
  const EncPage pg = pages[i];
  const EncCol &c = cols[pg.col];
  if (c.type != SB_BINARY && c.type != SB_LARGE_BINARY) {
    out[2 * i] = out[2 * i + 1] = 0;
    return;
  }
  if (c.W == 4) {
    const int32_t *o = reinterpret_cast<const int32_t *>(c.offsets);
    out[2 * i] = o[pg.row0];
    out[2 * i + 1] = o[pg.row0 + pg.n];
  } else {
    const int64_t *o = reinterpret_cast<const int64_t *>(c.offsets);
    out[2 * i] = o[pg.row0];
    out[2 * i + 1] = o[pg.row0 + pg.n];
  }
}

} // namespace sb

using namespace sb;


extern "C" {

void sb_release_encoded(sb_ctx *ctx, sb_encoded_column *outs, uint64_t n) {
  if (!ctx || !outs) return;
  cudaSetDevice(ctx->device);
  for (uint64_t i = 0; i < n; ++i) {
    EncOwner *o = static_cast<EncOwner *>(outs[i]._owner);
    if (!o) continue;
    if (o->dev) cudaFreeAsync(o->dev, ctx->stream);
    if (o->pinned.p) pinned_put(ctx, o->pinned);
    std::free(o->metas);
    delete o;
    std::memset(&outs[i], 0, sizeof(outs[i]));
  }
}

int32_t sb_encode_columns(sb_ctx *ctx, const sb_leaf_array *cols, uint64_t n_cols, const sb_write_options *opts, int32_t out_mem,
                          sb_encoded_column *outs) {
  if (!ctx) return SB_CUDA;
  if (!cols || !outs || !opts || (out_mem != SB_MEM_HOST && out_mem != SB_MEM_DEVICE)) return fail(ctx, SB_INVALID_ARG, "bad arguments");
  if (opts->default_compression != SB_C_NONE && opts->default_compression != SB_C_LZ4 && opts->default_compression != SB_C_SNAPPY &&
      opts->default_compression != SB_C_ZSTD)
    return fail(ctx, SB_OUT_OF_SPEC, "default_compression must be a common codec (None / LZ4 / Zstd / Snappy)");
  if (ctx->pending.active) sb_decode_finish_pending(ctx); // the table buffers are shared with an in-flight decode call
  SB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  std::memset(outs, 0, sizeof(sb_encoded_column) * n_cols);
  ctx->stats = sb_stats{};

  // ---- page split (NativeWriter::encode_chunk, write/common.rs:54-58,79-86)
  std::vector<EncCol> h_cols(n_cols);
  std::vector<EncPage> h_pages;
  std::vector<uint64_t> col_first_page(n_cols + 1, 0);
  std::vector<void *> d_inputs;
  auto free_inputs = [&]() {
    for (void *d : d_inputs) cudaFreeAsync(d, st);
    d_inputs.clear();
  };
  auto upload = [&](const void *src, uint64_t bytes, int mem, const uint8_t **dst) -> int {
    *dst = static_cast<const uint8_t *>(src);
    if (!src || mem == SB_MEM_DEVICE || bytes == 0) return SB_OK;
    void *d = nullptr;
    SB_CUDA_CHECK(ctx, cudaMallocAsync(&d, align_up(bytes + 32, 256), st));
    d_inputs.push_back(d);
    SB_CUDA_CHECK(ctx, cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st));
    *dst = static_cast<const uint8_t *>(d);
    return SB_OK;
  };
  uint64_t bytes_in = 0, max_rows = 0;
  bool any_binary = false;
  std::vector<LvCol> lv_cols; // nested columns of the call
  uint64_t n_lv_blocks = 0, n_lv_pages = 0;
  for (uint64_t c = 0; c < n_cols; ++c) {
    const sb_leaf_array &a = cols[c];
    if (a.leaf.type < SB_NULL || a.leaf.type > SB_I256) return fail(ctx, SB_NYI, "unsupported physical type");
    const bool nested = a.leaf.n_nested > 1;
    if (a.leaf.n_nested > SB_MAX_NESTED) return fail(ctx, SB_NYI, "nesting deeper than SB_MAX_NESTED");
    EncCol &ec = h_cols[c];
    std::memset(&ec, 0, sizeof(ec));
    if (nested) { // level thresholds per depth (read_basic.rs:91-117; same derivation as the decoder's ColDesc)
      if (a.leaf.nested_kind[a.leaf.n_nested - 1] != SB_N_PRIMITIVE) return fail(ctx, SB_INVALID_ARG, "the last nested entry must be the primitive leaf");
      if (a.leaf.type == SB_NULL) return fail(ctx, SB_NYI, "nested Null leaves");
      ec.n_nested = a.leaf.n_nested;
      for (int d = 0; d < a.leaf.n_nested; ++d) {
        ec.kind[d] = uint8_t(a.leaf.nested_kind[d]);
        ec.nnull[d] = a.leaf.nested_nullable[d] != 0;
        ec.cum_sum[d + 1] = uint8_t(ec.cum_sum[d] + ec.nnull[d] + (ec.kind[d] == SB_N_LIST));
        ec.cum_rep[d + 1] = uint8_t(ec.cum_rep[d] + (ec.kind[d] == SB_N_LIST));
      }
      auto bit_width = [](uint32_t v) { int w = 0; while (v) ++w, v >>= 1; return w; };
      ec.w_rep = bit_width(ec.cum_rep[a.leaf.n_nested]);
      ec.w_def = bit_width(ec.cum_sum[a.leaf.n_nested]);
      ec.n_levels = a.n_levels;
      ec.rows = a.rows;
      if (a.n_levels > 0xfffffff0ull * 16) return fail(ctx, SB_OUT_OF_SPEC, "too many level entries");
      if (a.rows > a.n_levels) return fail(ctx, SB_INVALID_ARG, "every top-level row owns at least one level entry");
      if (a.n_levels && ((ec.w_rep && !a.rep_levels) || (ec.w_def && !a.def_levels))) return fail(ctx, SB_INVALID_ARG, "rep_levels / def_levels is NULL");
    }
    ec.type = a.leaf.type;
    ec.nullable = a.leaf.nullable != 0;
    ec.W = type_width(a.leaf.type);
    ec.tclass = (a.leaf.type == SB_F32 || a.leaf.type == SB_F64) ? TC_FLOAT : ((a.leaf.type >= SB_I8 && a.leaf.type <= SB_I64) || a.leaf.type >= SB_I128) ? TC_SINT : TC_UINT;
    ec.length = a.length;
    ec.values_bytes = a.values_bytes;
    const bool binary = a.leaf.type == SB_BINARY || a.leaf.type == SB_LARGE_BINARY;
    any_binary |= binary;
    if (a.length && a.leaf.type != SB_NULL && !a.values && !(binary && a.values_bytes == 0)) return fail(ctx, SB_INVALID_ARG, "values is NULL");
    if (binary && !a.offsets) return fail(ctx, SB_INVALID_ARG, "offsets is NULL");
    uint64_t vbytes = a.leaf.type == SB_BOOL ? (a.length + 7) / 8 : binary ? a.values_bytes : a.length * uint64_t(ec.W);
    int rc;
    if ((rc = upload(a.values, vbytes, a.mem, &ec.values)) || (rc = upload(binary ? a.offsets : nullptr, (a.length + 1) * uint64_t(ec.W), a.mem, &ec.offsets)) ||
        (rc = upload(a.validity, (a.length + 7) / 8, a.mem, &ec.validity))) {
      free_inputs();
      return rc;
    }
    if (fixed_type(a.leaf.type) && (uintptr_t(ec.values) % uintptr_t(std::min(ec.W, 16))) != 0) {
      free_inputs();
      return fail(ctx, SB_INVALID_ARG, "values must be aligned to the element width");
    }
    if (nested) {
      const uint8_t *dr = nullptr, *dd = nullptr;
      if ((rc = upload(ec.w_rep ? a.rep_levels : nullptr, a.n_levels * 4, a.mem, &dr)) || (rc = upload(ec.w_def ? a.def_levels : nullptr, a.n_levels * 4, a.mem, &dd))) {
        free_inputs();
        return rc;
      }
      ec.rep = reinterpret_cast<const uint32_t *>(dr);
      ec.def = reinterpret_cast<const uint32_t *>(dd);
      bytes_in += (ec.w_rep ? a.n_levels * 4 : 0) + (ec.w_def ? a.n_levels * 4 : 0);
    }
    bytes_in += vbytes + (binary ? (a.length + 1) * uint64_t(ec.W) : 0) + (a.validity ? (a.length + 7) / 8 : 0);
    col_first_page[c] = h_pages.size();
    // pages are cut by top-level rows: the array length of a flat leaf, `rows` of a nested one
    const uint64_t n_split = nested ? a.rows : a.length;
    const uint64_t page_rows = opts->max_page_size ? std::min<uint64_t>(opts->max_page_size, n_split) : n_split;
    if (nested && n_split) {
      LvCol L{};
      L.col = uint32_t(c);
      L.blk0 = uint32_t(n_lv_blocks);
      L.n_blk = uint32_t((a.n_levels + kLvBlock - 1) / kLvBlock);
      L.page0 = uint32_t(n_lv_pages);
      L.n_pages = uint32_t((n_split + page_rows - 1) / page_rows);
      L.page_rows = uint32_t(std::min<uint64_t>(page_rows, 0xffffffffu));
      n_lv_blocks += L.n_blk;
      n_lv_pages += L.n_pages;
      lv_cols.push_back(L);
    }
    for (uint64_t r = 0, p = 0; r < n_split; r += page_rows, ++p) {
      EncPage pg{};
      pg.row0 = r; // nested: replaced by the first leaf slot once the level kernels have run
      pg.n = uint32_t(std::min<uint64_t>(page_rows, n_split - r));
      pg.rows = nested ? pg.n : 0;
      pg.col = uint32_t(c);
      pg.ordinal = uint32_t(p);
      if (page_rows > 0xfffffff0ull) {
        free_inputs();
        return fail(ctx, SB_OUT_OF_SPEC, "page larger than 2^32 rows (u32 size fields)");
      }
      h_pages.push_back(pg);
      max_rows = std::max<uint64_t>(max_rows, pg.n);
    }
  }
  col_first_page[n_cols] = h_pages.size();
  const uint64_t n_pages = h_pages.size();

  // device tables: [EncCol][EncPage][page_len][status][counters 2 + hist 32][col_dst ptrs][page offsets 2 x i64]
  size_t off_cols = 0;
  size_t off_pages = align_up(off_cols + sizeof(EncCol) * n_cols, 16);
  size_t off_len = align_up(off_pages + sizeof(EncPage) * n_pages, 16);
  size_t off_status = align_up(off_len + 4 * n_pages, 16);
  size_t off_ctr = align_up(off_status + 4 * n_pages, 16);
  size_t off_dst = align_up(off_ctr + 4 * 36, 16);
  size_t off_po = align_up(off_dst + 8 * n_cols, 16);
  size_t tbytes = align_up(off_po + 16 * n_pages, 16);
  int rc;
  if ((rc = host_tables_reserve(ctx, tbytes)) || (rc = dev_reserve(ctx, ctx->d_tables, tbytes))) {
    free_inputs();
    return rc;
  }
  uint8_t *hT = static_cast<uint8_t *>(ctx->h_tables), *dT = static_cast<uint8_t *>(ctx->d_tables.p);
#define SB_ETRY(call)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      free_inputs();                                                                  \
      sb_release_encoded(ctx, outs, n_cols);                                          \
      return fail(ctx, SB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    }                                                                                 \
  } while (0)
  std::memcpy(hT + off_cols, h_cols.data(), sizeof(EncCol) * n_cols);
  EncPage *hp = reinterpret_cast<EncPage *>(hT + off_pages);
  if (n_pages) std::memcpy(hp, h_pages.data(), sizeof(EncPage) * n_pages);
  const EncCol *d_cols = reinterpret_cast<const EncCol *>(dT + off_cols);
  const EncPage *d_pages = reinterpret_cast<const EncPage *>(dT + off_pages);

  // ---- nested leaves: first level entry / first leaf slot of every page (kernels N1-N3)
  if (!lv_cols.empty()) {
    const size_t lc_bytes = align_up(sizeof(LvCol) * lv_cols.size(), 16);
    const size_t blk_bytes = align_up(8 * n_lv_blocks, 16);
    const size_t tot_bytes = align_up(16 * lv_cols.size(), 16), bnd_bytes = 16 * n_lv_pages;
    uint8_t *d_lv = nullptr;
    SB_ETRY(cudaMallocAsync(reinterpret_cast<void **>(&d_lv), lc_bytes + 3 * blk_bytes + tot_bytes + bnd_bytes + 64, st));
    d_inputs.push_back(d_lv);
    LvCol *d_lcols = reinterpret_cast<LvCol *>(d_lv);
    uint2 *d_blk = reinterpret_cast<uint2 *>(d_lv + lc_bytes);
    uint64_t *d_rb = reinterpret_cast<uint64_t *>(d_lv + lc_bytes + blk_bytes), *d_sb = reinterpret_cast<uint64_t *>(d_lv + lc_bytes + 2 * blk_bytes);
    uint64_t *d_tot = reinterpret_cast<uint64_t *>(d_lv + lc_bytes + 3 * blk_bytes), *d_bnd = reinterpret_cast<uint64_t *>(d_lv + lc_bytes + 3 * blk_bytes + tot_bytes);
    SB_ETRY(cudaMemcpyAsync(dT, hT, off_pages, cudaMemcpyHostToDevice, st));
    SB_ETRY(cudaMemcpyAsync(d_lcols, lv_cols.data(), sizeof(LvCol) * lv_cols.size(), cudaMemcpyHostToDevice, st));
    SB_ETRY(cudaEventRecord(ctx->ev0, st)); // device_ms of a call with nested leaves includes the level kernels (and their host round trip)
    sb_level_count_kernel<<<uint32_t(n_lv_blocks), 256, 0, st>>>(d_cols, d_lcols, uint32_t(lv_cols.size()), d_blk);
    sb_level_scan_kernel<<<uint32_t(lv_cols.size()), SB_NT, 0, st>>>(d_lcols, d_blk, d_rb, d_sb, d_tot);
    sb_level_bounds_kernel<<<uint32_t(n_lv_pages), SB_NT, 0, st>>>(d_cols, d_lcols, uint32_t(lv_cols.size()), d_rb, d_sb, d_bnd);
    SB_ETRY(cudaGetLastError());
    ctx->stats.kernel_launches += 3;
    std::vector<uint64_t> h_tot(2 * lv_cols.size()), h_bnd(2 * n_lv_pages);
    SB_ETRY(cudaMemcpyAsync(h_tot.data(), d_tot, 16 * lv_cols.size(), cudaMemcpyDeviceToHost, st));
    SB_ETRY(cudaMemcpyAsync(h_bnd.data(), d_bnd, bnd_bytes, cudaMemcpyDeviceToHost, st));
    SB_ETRY(cudaStreamSynchronize(st));
    for (size_t k = 0; k < lv_cols.size(); ++k) {
      const LvCol &L = lv_cols[k];
      const sb_leaf_array &a = cols[L.col];
      if (h_tot[2 * k] != a.rows || h_tot[2 * k + 1] != a.length) {
        free_inputs();
        return fail(ctx, SB_INVALID_ARG, "column " + std::to_string(L.col) + ": the levels describe " + std::to_string(h_tot[2 * k]) + " rows / " +
                                             std::to_string(h_tot[2 * k + 1]) + " leaf slots, the caller passed " + std::to_string(a.rows) + " / " +
                                             std::to_string(a.length));
      }
      for (uint32_t p = 0; p < L.n_pages; ++p) {
        EncPage &pg = hp[col_first_page[L.col] + p];
        const uint64_t e0 = h_bnd[2 * (L.page0 + p)], s0 = h_bnd[2 * (L.page0 + p) + 1];
        const uint64_t e1 = p + 1 < L.n_pages ? h_bnd[2 * (L.page0 + p + 1)] : a.n_levels;
        const uint64_t s1 = p + 1 < L.n_pages ? h_bnd[2 * (L.page0 + p + 1) + 1] : a.length;
        if (e0 == ~0ull || e1 == ~0ull || e1 < e0 || s1 < s0 || e1 - e0 > 0xfffffff0ull || s1 - s0 > 0xfffffff0ull || (p == 0 && e0 != 0)) {
          free_inputs();
          return fail(ctx, SB_INVALID_ARG, "column " + std::to_string(L.col) + ": rep_levels do not start a row where a page begins");
        }
        pg.row0 = s0;
        pg.n = uint32_t(s1 - s0);
        pg.lv0 = e0;
        pg.n_lv = uint32_t(e1 - e0);
        max_rows = std::max<uint64_t>(max_rows, std::max<uint64_t>(pg.n, pg.n_lv));
      }
    }
  }

  // ---- slab sizing (worst-case page encodings)
  std::vector<int64_t> page_bytes(n_pages, 0);
  if (any_binary && n_pages) {
    SB_ETRY(cudaMemcpyAsync(dT, hT, off_len, cudaMemcpyHostToDevice, st));
    sb_page_offsets_kernel<<<uint32_t((n_pages + 255) / 256), 256, 0, st>>>(d_pages, d_cols, uint32_t(n_pages), reinterpret_cast<int64_t *>(dT + off_po));
    SB_ETRY(cudaMemcpyAsync(hT + off_po, dT + off_po, 16 * n_pages, cudaMemcpyDeviceToHost, st));
    SB_ETRY(cudaStreamSynchronize(st));
    const int64_t *po = reinterpret_cast<const int64_t *>(hT + off_po);
    for (uint64_t i = 0; i < n_pages; ++i) {
      page_bytes[i] = po[2 * i + 1] - po[2 * i];
      if (page_bytes[i] < 0 || page_bytes[i] > 0xfffffff0ll) {
        free_inputs();
        return fail(ctx, SB_OUT_OF_SPEC, "binary offsets are not monotone, or a page holds more than 4 GiB of values");
      }
    }
  }
  uint64_t slab_total = 0;
  for (uint64_t i = 0; i < n_pages; ++i) {
    EncPage &pg = hp[i];
    const EncCol &ec = h_cols[pg.col];
    const uint64_t n = pg.n;
    uint64_t cap = 64 + (ec.nullable ? 16 + n / 8 : 0);
    if (ec.n_nested > 1) cap += 12 + 2 * 16 + (uint64_t(pg.n_lv) / 8 + 1) * uint64_t(ec.w_rep + ec.w_def);
    if (ec.type == SB_BOOL) cap += 5 * n + 64;
    else if (ec.type == SB_BINARY || ec.type == SB_LARGE_BINARY) cap += 2 * uint64_t(page_bytes[i]) + 25 * n + 9000 * (n / 65536 + 1) + 2048;
    else if (ec.type != SB_NULL) cap += n * uint64_t(ec.W + 13) + 9000 * (n / 65536 + 1) + 2048;
    if (cap > 0xffffffffull) {
      free_inputs();
      return fail(ctx, SB_OUT_OF_SPEC, "page larger than 4 GiB (u32 size fields)");
    }
    pg.slab_off = slab_total;
    slab_total += align_up(cap, 16);
  }
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  int occ = 1;
  SB_ETRY(cudaFuncSetAttribute(sb_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kEncSmem)));
  SB_ETRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sb_encode_kernel, SB_NT, kEncSmem));
  occ = std::max(1, occ);
  const uint64_t scratch_per_cta = align_up(176 * max_rows + 128 * 1024, 256);
  uint64_t grid = std::min<uint64_t>(std::max<uint64_t>(n_pages, 1), uint64_t(ctx->sm_count) * occ);
  grid = std::max<uint64_t>(1, std::min<uint64_t>(grid, (uint64_t(8) << 30) / scratch_per_cta));
  void *d_slab = nullptr;
  if (n_pages) {
    if (slab_total + scratch_per_cta * grid > free_b + ctx->d_scratch.cap) {
      free_inputs();
      return fail(ctx, SB_NYI, "encode batch does not fit in device memory: split the call into fewer columns");
    }
    SB_ETRY(cudaMallocAsync(&d_slab, slab_total + 64, st));
    d_inputs.push_back(d_slab);
    if ((rc = dev_reserve(ctx, ctx->d_scratch, scratch_per_cta * grid))) {
      free_inputs();
      return rc;
    }
    SB_ETRY(cudaMemcpyAsync(dT, hT, off_len, cudaMemcpyHostToDevice, st));
    SB_ETRY(cudaMemsetAsync(dT + off_len, 0, off_dst - off_len, st));
    EOpts eo;
    eo.def_codec = opts->default_compression;
    eo.ratio = opts->default_compress_ratio;
    eo.forbidden = opts->forbidden_mask;
    eo.force = opts->force_codec;
    eo.seed = opts->seed;
    uint32_t *d_ctr = reinterpret_cast<uint32_t *>(dT + off_ctr);
    if (lv_cols.empty()) SB_ETRY(cudaEventRecord(ctx->ev0, st));
    sb_encode_kernel<<<uint32_t(grid), SB_NT, kEncSmem, st>>>(d_pages, d_cols, uint32_t(n_pages), d_ctr, static_cast<uint8_t *>(d_slab),
                                                             static_cast<uint8_t *>(ctx->d_scratch.p), scratch_per_cta,
                                                             reinterpret_cast<uint32_t *>(dT + off_len), reinterpret_cast<int32_t *>(dT + off_status), eo,
                                                             d_ctr + 4);
    SB_ETRY(cudaGetLastError());
    SB_ETRY(cudaEventRecord(ctx->ev1, st));
    ctx->stats.kernel_launches += 1;
    SB_ETRY(cudaMemcpyAsync(hT + off_len, dT + off_len, off_dst - off_len, cudaMemcpyDeviceToHost, st));
    SB_ETRY(cudaStreamSynchronize(st));
  }

  // ---- PageMeta + column bodies
  const uint32_t *h_len = reinterpret_cast<const uint32_t *>(hT + off_len);
  const int32_t *h_status = reinterpret_cast<const int32_t *>(hT + off_status);
  void **h_dst = reinterpret_cast<void **>(hT + off_dst);
  int32_t first_err = SB_OK;
  uint64_t bytes_out = 0;
  for (uint64_t c = 0; c < n_cols; ++c) {
    EncOwner *ow = new EncOwner();
    outs[c]._owner = ow;
    outs[c].mem = out_mem;
    const uint64_t p0 = col_first_page[c], p1 = col_first_page[c + 1];
    outs[c].n_pages = p1 - p0;
    sb_page_meta *metas = static_cast<sb_page_meta *>(std::malloc(sizeof(sb_page_meta) * std::max<uint64_t>(1, p1 - p0)));
    ow->metas = metas;
    outs[c].metas = metas;
    uint64_t pos = 0;
    for (uint64_t p = p0; p < p1; ++p) {
      if (h_status[p] != SB_OK && first_err == SB_OK) {
        first_err = h_status[p];
        ctx->err = "page " + std::to_string(p - p0) + " of column " + std::to_string(c) + " failed to encode with status " + std::to_string(h_status[p]);
      }
      metas[p - p0].length = h_len[p];
      metas[p - p0].num_values = h_cols[c].n_nested > 1 ? hp[p].n_lv : hp[p].n; // rows (flat) / level entries (nested), common.rs:103
      hp[p].dst_off = pos;
      pos += h_len[p];
    }
    outs[c].nbytes = pos;
    bytes_out += pos;
    h_dst[c] = nullptr;
    if (pos) {
      SB_ETRY(cudaMallocAsync(&ow->dev, pos + 16, st));
      h_dst[c] = ow->dev;
    }
  }
  if (n_pages) {
    SB_ETRY(cudaMemcpyAsync(dT + off_pages, hT + off_pages, sizeof(EncPage) * n_pages, cudaMemcpyHostToDevice, st));
    SB_ETRY(cudaMemcpyAsync(dT + off_dst, hT + off_dst, 8 * n_cols, cudaMemcpyHostToDevice, st));
    uint32_t *d_ctr = reinterpret_cast<uint32_t *>(dT + off_ctr);
    uint32_t ggrid = uint32_t(std::min<uint64_t>(n_pages, uint64_t(ctx->sm_count) * 8));
    sb_gather_kernel<<<ggrid, SB_NT, 0, st>>>(d_pages, uint32_t(n_pages), reinterpret_cast<const uint32_t *>(dT + off_len),
                                              static_cast<const uint8_t *>(d_slab), reinterpret_cast<uint8_t *const *>(dT + off_dst), d_ctr + 1);
    SB_ETRY(cudaGetLastError());
    ctx->stats.kernel_launches += 1;
  }
  for (uint64_t c = 0; c < n_cols; ++c) {
    EncOwner *ow = static_cast<EncOwner *>(outs[c]._owner);
    if (out_mem == SB_MEM_DEVICE) {
      outs[c].bytes = static_cast<uint8_t *>(ow->dev);
    } else if (outs[c].nbytes) {
      if ((rc = pinned_get(ctx, outs[c].nbytes, &ow->pinned))) {
        free_inputs();
        sb_release_encoded(ctx, outs, n_cols);
        return rc;
      }
      SB_ETRY(cudaMemcpyAsync(ow->pinned.p, ow->dev, outs[c].nbytes, cudaMemcpyDeviceToHost, st));
      outs[c].bytes = static_cast<uint8_t *>(ow->pinned.p);
    }
  }
  free_inputs();
  SB_ETRY(cudaStreamSynchronize(st));
  if (out_mem == SB_MEM_HOST)
    for (uint64_t c = 0; c < n_cols; ++c) {
      EncOwner *ow = static_cast<EncOwner *>(outs[c]._owner);
      if (ow->dev) cudaFreeAsync(ow->dev, st);
      ow->dev = nullptr;
    }
  if (n_pages) cudaEventElapsedTime(&ctx->stats.device_ms, ctx->ev0, ctx->ev1);
  const uint32_t *h_ctr = reinterpret_cast<const uint32_t *>(hT + off_ctr);
  if (n_pages)
    for (int i = 0; i < 32; ++i) ctx->stats.codec_pages[i] = h_ctr[4 + i];
  ctx->stats.pages = n_pages;
  ctx->stats.bytes_in = bytes_in;
  ctx->stats.bytes_out = bytes_out;
  return first_err;
#undef SB_ETRY
}

} // extern "C"
