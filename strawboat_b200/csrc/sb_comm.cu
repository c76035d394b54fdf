// sb_comm.cu -- the one exchange step of the multi-GPU encode path: the file is column major with
// absolute ColumnMeta offsets and a single footer (src/write/common.rs:76,111-114, src/write/writer.rs:128-167), so
// the encoded column bodies of all ranks are gathered on the writer rank before the footer is written.
//
//   1. ncclAllGather of (encoded bytes, page count) per leaf column        -> every rank knows the file layout
//   2. grouped ncclSend / ncclRecv of the bodies, each straight to its final position in the writer's staging
//      buffer (= the file's body region, leaf order), plus one message of PageMeta records per rank
//   3. the writer frames header, body region and footer (host side)
//
// Leaf c lives on rank c mod world; decode needs no collective at all (pages are independent).
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when there is one), so the
// library keeps loading on a host without NCCL or without a GPU.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <new>

#include "sb_host.h"

namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
  bool ok = false;
};
NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the copy already in the process (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
#define SB_SYM(field, name)                                          \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); \
  if (!api.field) return nullptr;
  SB_SYM(GetUniqueId, "ncclGetUniqueId")
  SB_SYM(CommInitRank, "ncclCommInitRank")
  SB_SYM(CommDestroy, "ncclCommDestroy")
  SB_SYM(AllGather, "ncclAllGather")
  SB_SYM(Send, "ncclSend")
  SB_SYM(Recv, "ncclRecv")
  SB_SYM(GroupStart, "ncclGroupStart")
  SB_SYM(GroupEnd, "ncclGroupEnd")
  SB_SYM(GetErrorString, "ncclGetErrorString")
#undef SB_SYM
  api.ok = true;
  return &api;
}
} // namespace

struct sb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

#define SB_NCCL(ctx, api, call)                                                                              \
  do {                                                                                                       \
    ncclResult_t r__ = (call);                                                                               \
    if (r__ != ncclSuccess) return fail(ctx, SB_EXTERNAL, std::string(#call) + ": " + (api)->GetErrorString(r__)); \
  } while (0)

extern "C" {

int32_t sb_comm_unique_id(uint8_t *id128) {
  NcclApi *api = nccl_api();
  if (!api || !id128) return SB_NYI;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return SB_EXTERNAL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(id128, &id, 128);
  return SB_OK;
}

int32_t sb_comm_create(sb_ctx *ctx, int32_t rank, int32_t world, const uint8_t *id128, sb_comm **out) {
  if (!ctx) return SB_CUDA;
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, SB_INVALID_ARG, "bad arguments");
  *out = nullptr;
  NcclApi *api = nccl_api();
  if (!api) return fail(ctx, SB_NYI, "libnccl.so.2 could not be loaded");
  SB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  sb_comm *c = new (std::nothrow) sb_comm();
  if (!c) return fail(ctx, SB_CUDA, "out of memory");
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    delete c;
    return fail(ctx, SB_EXTERNAL, std::string("ncclCommInitRank: ") + api->GetErrorString(r));
  }
  *out = c;
  return SB_OK;
}

void sb_comm_destroy(sb_comm *c) {
  if (!c) return;
  NcclApi *api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
}

int32_t sb_gather_encoded(sb_ctx *ctx, sb_comm *comm, const sb_encoded_column *local, uint64_t n_local, uint64_t n_total, int32_t writer,
                          sb_encoded_column *outs, sb_gather_stats *stats) {
  if (!ctx) return SB_CUDA;
  NcclApi *api = nccl_api();
  if (!api) return fail(ctx, SB_NYI, "libnccl.so.2 could not be loaded");
  if (!comm || (n_local && !local) || writer < 0 || writer >= comm->world) return fail(ctx, SB_INVALID_ARG, "bad arguments");
  const int rank = comm->rank, world = comm->world;
  const uint64_t expect_local = n_total > uint64_t(rank) ? (n_total - rank + world - 1) / world : 0;
  if (n_local != expect_local) return fail(ctx, SB_INVALID_ARG, "leaf c lives on rank c mod world: wrong number of local columns");
  if (rank == writer && !outs) return fail(ctx, SB_INVALID_ARG, "outs is NULL on the writer rank");
  for (uint64_t k = 0; k < n_local; ++k)
    if (local[k].nbytes && local[k].mem != SB_MEM_DEVICE) return fail(ctx, SB_INVALID_ARG, "encoded columns must be device resident (out_mem = SB_MEM_DEVICE)");
  SB_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (stats) *stats = sb_gather_stats{};

  // ---- 1. layout: (bytes, pages) of every leaf column, from its owner
  std::vector<uint64_t> mine(2 * n_total, 0), all(size_t(2) * n_total * world, 0);
  for (uint64_t k = 0; k < n_local; ++k) {
    mine[2 * (rank + k * world)] = local[k].nbytes;
    mine[2 * (rank + k * world) + 1] = local[k].n_pages;
  }
  uint64_t *d_mine = nullptr, *d_all = nullptr;
  std::vector<void *> tmp;
  auto cleanup = [&]() {
    for (void *p : tmp) cudaFreeAsync(p, st);
    tmp.clear();
  };
#define SB_GTRY(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      cleanup();                                                                               \
      return fail(ctx, SB_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    }                                                                                          \
  } while (0)
#define SB_GNCCL(call)                                                                          \
  do {                                                                                          \
    ncclResult_t r__ = (call);                                                                  \
    if (r__ != ncclSuccess) {                                                                   \
      cleanup();                                                                                \
      return fail(ctx, SB_EXTERNAL, std::string(#call) + ": " + api->GetErrorString(r__));     \
    }                                                                                           \
  } while (0)
  SB_GTRY(cudaMallocAsync(reinterpret_cast<void **>(&d_mine), 16 * n_total + 16, st));
  tmp.push_back(d_mine);
  SB_GTRY(cudaMallocAsync(reinterpret_cast<void **>(&d_all), 16 * n_total * world + 16, st));
  tmp.push_back(d_all);
  SB_GTRY(cudaMemcpyAsync(d_mine, mine.data(), 16 * n_total, cudaMemcpyHostToDevice, st));
  SB_GNCCL(api->AllGather(d_mine, d_all, 2 * n_total, ncclUint64, comm->comm, st));
  SB_GTRY(cudaMemcpyAsync(all.data(), d_all, 16 * n_total * world, cudaMemcpyDeviceToHost, st));
  SB_GTRY(cudaStreamSynchronize(st));
  std::vector<uint64_t> col_bytes(n_total), col_pages(n_total), body_off(n_total + 1, 0);
  std::vector<uint64_t> rank_pages(world, 0);
  for (uint64_t c = 0; c < n_total; ++c) {
    const uint64_t *e = all.data() + (size_t(c % world) * n_total + c) * 2;
    col_bytes[c] = e[0];
    col_pages[c] = e[1];
    body_off[c + 1] = body_off[c] + e[0];
    rank_pages[c % world] += e[1];
  }

  // ---- 2. bodies to their final position on the writer; PageMeta records, one message per rank
  uint64_t my_pages = 0;
  for (uint64_t k = 0; k < n_local; ++k) my_pages += local[k].n_pages;
  sb_page_meta *d_my_metas = nullptr;
  if (rank != writer && my_pages) {
    std::vector<sb_page_meta> flat;
    flat.reserve(my_pages);
    for (uint64_t k = 0; k < n_local; ++k) flat.insert(flat.end(), local[k].metas, local[k].metas + local[k].n_pages);
    SB_GTRY(cudaMallocAsync(reinterpret_cast<void **>(&d_my_metas), sizeof(sb_page_meta) * my_pages, st));
    tmp.push_back(d_my_metas);
    SB_GTRY(cudaMemcpyAsync(d_my_metas, flat.data(), sizeof(sb_page_meta) * my_pages, cudaMemcpyHostToDevice, st));
    SB_GTRY(cudaStreamSynchronize(st)); // `flat` goes out of scope
  }
  uint8_t *staging = nullptr;
  std::vector<sb_page_meta *> d_rank_metas(world, nullptr);
  if (rank == writer) {
    SB_GTRY(cudaMallocAsync(reinterpret_cast<void **>(&staging), body_off[n_total] + 16, st));
    for (int r = 0; r < world; ++r)
      if (r != writer && rank_pages[r]) {
        SB_GTRY(cudaMallocAsync(reinterpret_cast<void **>(&d_rank_metas[r]), sizeof(sb_page_meta) * rank_pages[r], st));
        tmp.push_back(d_rank_metas[r]);
      }
  }
  SB_GTRY(cudaEventRecord(ctx->ev0, st));
  uint64_t moved = 0;
  SB_GNCCL(api->GroupStart());
  if (rank == writer) {
    for (uint64_t c = 0; c < n_total; ++c) {
      const int owner = int(c % world);
      if (owner == writer || col_bytes[c] == 0) continue;
      SB_GNCCL(api->Recv(staging + body_off[c], col_bytes[c], ncclUint8, owner, comm->comm, st));
      moved += col_bytes[c];
    }
    for (int r = 0; r < world; ++r)
      if (d_rank_metas[r]) {
        SB_GNCCL(api->Recv(d_rank_metas[r], 2 * rank_pages[r], ncclUint64, r, comm->comm, st));
        moved += sizeof(sb_page_meta) * rank_pages[r];
      }
  } else {
    for (uint64_t k = 0; k < n_local; ++k)
      if (local[k].nbytes) {
        SB_GNCCL(api->Send(local[k].bytes, local[k].nbytes, ncclUint8, writer, comm->comm, st));
        moved += local[k].nbytes;
      }
    if (my_pages) {
      SB_GNCCL(api->Send(d_my_metas, 2 * my_pages, ncclUint64, writer, comm->comm, st));
      moved += sizeof(sb_page_meta) * my_pages;
    }
  }
  SB_GNCCL(api->GroupEnd());
  SB_GTRY(cudaEventRecord(ctx->ev1, st));
  if (rank == writer) // the writer's own columns: device-to-device into the staging buffer
    for (uint64_t k = 0; k < n_local; ++k)
      if (local[k].nbytes)
        SB_GTRY(cudaMemcpyAsync(staging + body_off[rank + k * world], local[k].bytes, local[k].nbytes, cudaMemcpyDeviceToDevice, st));

  // ---- 3. hand the gathered columns to the caller (writer only)
  if (rank == writer) {
    std::vector<std::vector<sb_page_meta>> h_rank_metas(world);
    for (int r = 0; r < world; ++r)
      if (d_rank_metas[r]) {
        h_rank_metas[r].resize(rank_pages[r]);
        SB_GTRY(cudaMemcpyAsync(h_rank_metas[r].data(), d_rank_metas[r], sizeof(sb_page_meta) * rank_pages[r], cudaMemcpyDeviceToHost, st));
      }
    SB_GTRY(cudaStreamSynchronize(st));
    std::vector<uint64_t> rank_pos(world, 0);
    std::memset(outs, 0, sizeof(sb_encoded_column) * n_total);
    for (uint64_t c = 0; c < n_total; ++c) {
      const int owner = int(c % world);
      EncOwner *ow = new EncOwner();
      if (c == 0) ow->dev = staging; // the first column owns the staging buffer (released with it)
      sb_page_meta *m = static_cast<sb_page_meta *>(std::malloc(sizeof(sb_page_meta) * std::max<uint64_t>(1, col_pages[c])));
      ow->metas = m;
      if (owner == writer) {
        const sb_encoded_column &lc = local[(c - rank) / world];
        std::memcpy(m, lc.metas, sizeof(sb_page_meta) * col_pages[c]);
      } else {
        std::memcpy(m, h_rank_metas[owner].data() + rank_pos[owner], sizeof(sb_page_meta) * col_pages[c]);
        rank_pos[owner] += col_pages[c];
      }
      outs[c].bytes = staging + body_off[c];
      outs[c].nbytes = col_bytes[c];
      outs[c].metas = m;
      outs[c].n_pages = col_pages[c];
      outs[c].mem = SB_MEM_DEVICE;
      outs[c]._owner = ow;
    }
    if (n_total == 0 && staging) cudaFreeAsync(staging, st);
  } else {
    SB_GTRY(cudaStreamSynchronize(st));
  }
  cleanup();
  if (stats) {
    stats->bytes_moved = moved;
    stats->total_bytes = body_off[n_total];
    cudaEventElapsedTime(&stats->gather_ms, ctx->ev0, ctx->ev1);
  }
  return SB_OK;
#undef SB_GTRY
#undef SB_GNCCL
}

} // extern "C"
