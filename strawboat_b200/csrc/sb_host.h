// sb_host.h -- host-side state shared by the decode (sb_lib.cu) and encode (sb_encode.cu)
// translation units of libstrawboat_b200.so.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/strawboat_b200.h"

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};
struct PinnedBlock {
  void *p;
  size_t cap;
};
struct Owner { // allocations handed to the caller through sb_column_out / sb_encoded_column
  std::vector<void *> dev;
  std::vector<PinnedBlock> host_pinned;
  std::vector<void *> host_malloc;
};

struct EncOwner { // what an sb_encoded_column owns (sb_encode_columns, sb_gather_encoded)
  void *dev = nullptr;
  PinnedBlock pinned{nullptr, 0};
  void *metas = nullptr;
};

struct DecodePending { // a submitted decode call whose results have not been collected yet (sb_decode_wait)
  bool active = false;
  uint64_t n_cols = 0, n_pages_total = 0, n_items = 0;
  int32_t out_mem = 0;
  sb_column_out *outs = nullptr;
  std::vector<Owner *> owners;
  std::vector<uint64_t> col_pages;
  size_t off_status = 0, off_counters = 0;
  bool any_fixed = false, split = false;
  uint64_t bytes_in = 0, bytes_out = 0;
  double t_host0 = 0, t_planned = 0, t_submitted = 0;
};

struct sb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, aux = nullptr, aux2 = nullptr;
  bool own_stream = false;
  std::string err;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_cls = nullptr;
  cudaEvent_t ev_m0 = nullptr, ev_m1 = nullptr, ev_lz0 = nullptr, ev_lz1 = nullptr; // per-kernel timing
  cudaEvent_t ev_l0 = nullptr, ev_l1 = nullptr, ev_join2 = nullptr;                 // light decode kernel (third stream)
  int sm_count = 0;
  int max_smem_optin = 0;
  int lz4_occ = 0, occ_val = 0, light_occ = 0; // cached occupancy queries
  uint32_t occ_smem = 0;
  DevBuf d_tables, d_scratch, d_entries;
  void *h_tables = nullptr;
  size_t h_tables_cap = 0;
  std::vector<PinnedBlock> pinned_free; // pinned host blocks are expensive to create: recycled
  sb_stats stats{};
  DecodePending pending;
};
int32_t sb_decode_finish_pending(sb_ctx *ctx); // sb_lib.cu: collect an outstanding asynchronous decode (no-op when none)

inline int fail(sb_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  return code;
}
#define SB_CUDA_CHECK(ctx, call)                                                                          \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess) return fail(ctx, SB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

inline int dev_reserve(sb_ctx *ctx, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return SB_OK;
  if (b.p) SB_CUDA_CHECK(ctx, cudaFreeAsync(b.p, ctx->stream));
  b.p = nullptr;
  b.cap = 0;
  size_t cap = std::max(bytes + bytes / 4, size_t(1) << 20);
  SB_CUDA_CHECK(ctx, cudaMallocAsync(&b.p, cap, ctx->stream));
  b.cap = cap;
  return SB_OK;
}
inline int host_tables_reserve(sb_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->h_tables_cap) return SB_OK;
  if (ctx->h_tables) cudaFreeHost(ctx->h_tables);
  ctx->h_tables = nullptr;
  size_t cap = std::max(bytes + bytes / 4, size_t(1) << 20);
  SB_CUDA_CHECK(ctx, cudaMallocHost(&ctx->h_tables, cap));
  ctx->h_tables_cap = cap;
  return SB_OK;
}
// best-fit from the recycle list, else a new pinned allocation
inline int pinned_get(sb_ctx *ctx, size_t bytes, PinnedBlock *out) {
  bytes = std::max<size_t>(bytes, 64);
  int best = -1;
  for (size_t i = 0; i < ctx->pinned_free.size(); ++i)
    if (ctx->pinned_free[i].cap >= bytes && ctx->pinned_free[i].cap <= 2 * bytes + 4096 &&
        (best < 0 || ctx->pinned_free[i].cap < ctx->pinned_free[size_t(best)].cap))
      best = int(i);
  if (best >= 0) {
    *out = ctx->pinned_free[size_t(best)];
    ctx->pinned_free.erase(ctx->pinned_free.begin() + best);
    return SB_OK;
  }
  void *h = nullptr;
  SB_CUDA_CHECK(ctx, cudaMallocHost(&h, bytes));
  *out = PinnedBlock{h, bytes};
  return SB_OK;
}
inline void pinned_put(sb_ctx *ctx, PinnedBlock b) {
  if (ctx->pinned_free.size() >= 64) {
    cudaFreeHost(b.p);
    return;
  }
  ctx->pinned_free.push_back(b);
}

inline int type_width(int t) {
  switch (t) {
  case SB_I8:
  case SB_U8: return 1;
  case SB_I16:
  case SB_U16: return 2;
  case SB_I32:
  case SB_U32:
  case SB_F32: return 4;
  case SB_I64:
  case SB_U64:
  case SB_F64: return 8;
  case SB_I128: return 16;
  case SB_I256: return 32;
  case SB_BINARY: return 4;
  case SB_LARGE_BINARY: return 8;
  }
  return 0;
}
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline bool fixed_type(int t) { return (t >= SB_I8 && t <= SB_F64) || t == SB_I128 || t == SB_I256; }

