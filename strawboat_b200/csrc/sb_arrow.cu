// sb_arrow.cu -- nested assembly + Arrow C Data Interface export: the read half of SURVEY §8 f2 and the ownership
// half of the boundary (§8b).
//
// The reference turns the decoded leaves and their NestedState into arrays with create_list / create_struct
// (src/read/array/{list,struct_}.rs, src/read/batch_read.rs:66-187; arrow2 io::parquet::read::create_list): list
// offsets + validity and struct validity come from the FIRST leaf below the node, the children are the leaves.
// sb_export_arrow does the same over the buffers sb_decode_columns produced, without copying them, and hands the
// result out as ArrowArray / ArrowSchema (the C Data Interface): the consumer -- arrow-rs / arrow2 FFI on the Rust
// side, pyarrow in the tests -- gets ordinary ListArray / StructArray / PrimitiveArray / Utf8Array values whose
// release callback gives the buffers back to the context.
#include <cstdio>
#include <cstring>
#include <new>

#include "sb_common.cuh"
#include "sb_host.h"

namespace {

__global__ void narrow_offsets_kernel(const int64_t *__restrict__ in, int32_t *out, uint64_t n) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = int32_t(in[i]);
}

struct NodePriv {
  sb_ctx *ctx = nullptr;
  std::vector<const void *> buffers;
  std::vector<ArrowArray *> children;
  std::vector<void *> host_allocs;  // converted list offsets (host)
  std::vector<void *> dev_allocs;   // converted list offsets (device)
  std::vector<sb_column_out> owned; // root only: the decoded leaves whose buffers the tree points into
};

void release_array(ArrowArray *a) {
  if (!a || !a->release) return;
  NodePriv *p = static_cast<NodePriv *>(a->private_data);
  for (ArrowArray *c : p->children) {
    if (c->release) c->release(c);
    delete c;
  }
  for (void *h : p->host_allocs) std::free(h);
  if (!p->dev_allocs.empty() || !p->owned.empty()) {
    cudaSetDevice(p->ctx->device);
    for (void *d : p->dev_allocs) cudaFreeAsync(d, p->ctx->stream);
    if (!p->owned.empty()) sb_release_columns(p->ctx, p->owned.data(), p->owned.size());
  }
  delete p;
  a->release = nullptr;
  a->private_data = nullptr;
}

struct SchemaPriv {
  std::string format, name;
  std::vector<ArrowSchema *> children;
};
void release_schema(ArrowSchema *s) {
  if (!s || !s->release) return;
  SchemaPriv *p = static_cast<SchemaPriv *>(s->private_data);
  for (ArrowSchema *c : p->children) {
    if (c->release) c->release(c);
    delete c;
  }
  delete p;
  s->release = nullptr;
  s->private_data = nullptr;
}

const char *leaf_format(int type, int utf8) {
  switch (type) {
  case SB_NULL: return "n";
  case SB_BOOL: return "b";
  case SB_I8: return "c";
  case SB_I16: return "s";
  case SB_I32: return "i";
  case SB_I64: return "l";
  case SB_U8: return "C";
  case SB_U16: return "S";
  case SB_U32: return "I";
  case SB_U64: return "L";
  case SB_F32: return "f";
  case SB_F64: return "g";
  case SB_BINARY: return utf8 ? "u" : "z";
  case SB_LARGE_BINARY: return utf8 ? "U" : "Z";
  case SB_I128: return "d:38,0";
  case SB_I256: return "d:76,0,256";
  }
  return nullptr;
}

int build_schema(const sb_field *f, ArrowSchema *out) {
  SchemaPriv *p = new SchemaPriv();
  std::memset(out, 0, sizeof(*out));
  p->name = f->name ? f->name : "";
  if (f->kind == SB_N_PRIMITIVE) {
    const char *fmt = leaf_format(f->type, f->utf8);
    if (!fmt) {
      delete p;
      return SB_NYI;
    }
    p->format = fmt;
  } else if (f->kind == SB_N_LIST) {
    p->format = f->large ? "+L" : "+l";
  } else {
    p->format = "+s";
  }
  out->private_data = p;
  out->release = release_schema;
  out->flags = f->nullable ? 2 /* ARROW_FLAG_NULLABLE */ : 0;
  for (int32_t c = 0; c < f->n_children; ++c) {
    ArrowSchema *cs = new ArrowSchema();
    int rc = build_schema(&f->children[c], cs);
    if (rc) {
      delete cs;
      release_schema(out);
      return rc;
    }
    p->children.push_back(cs);
  }
  out->format = p->format.c_str();
  out->name = p->name.c_str();
  out->n_children = int64_t(p->children.size());
  out->children = p->children.empty() ? nullptr : p->children.data();
  return SB_OK;
}

uint64_t count_leaves(const sb_field *f) {
  if (f->kind == SB_N_PRIMITIVE) return 1;
  uint64_t n = 0;
  for (int32_t c = 0; c < f->n_children; ++c) n += count_leaves(&f->children[c]);
  return n;
}

// depth = number of ancestors of this node = index into the leaves' nested_* arrays
int build_array(sb_ctx *ctx, const sb_field *f, const sb_column_out *leaves, uint64_t *cursor, int depth, ArrowArray *out) {
  NodePriv *p = new NodePriv();
  p->ctx = ctx;
  std::memset(out, 0, sizeof(*out));
  out->private_data = p;
  out->release = release_array;
  out->null_count = -1;
  const sb_column_out &first = leaves[*cursor]; // create_list / create_struct read the first leaf's NestedState
  if (f->kind == SB_N_PRIMITIVE) {
    const sb_column_out &lf = leaves[(*cursor)++];
    out->length = int64_t(lf.length);
    if (f->type == SB_NULL) {
      out->null_count = out->length;
    } else {
      p->buffers.push_back(lf.validity);
      if (f->type == SB_BINARY || f->type == SB_LARGE_BINARY) p->buffers.push_back(lf.offsets);
      p->buffers.push_back(lf.values);
      if (!lf.validity) out->null_count = 0;
    }
  } else if (f->kind == SB_N_LIST) {
    if (f->n_children != 1) return SB_INVALID_ARG;
    const uint64_t len = first.nested_len[depth];
    const int64_t *off64 = first.nested_offsets[depth];
    if (!off64 && len == 0) { // a column without pages: no NestedState was ever built -- an empty list array has offsets [0]
      if (first.mem == SB_MEM_HOST) {
        void *z = std::calloc(1, 16);
        if (!z) return SB_CUDA;
        p->host_allocs.push_back(z);
        off64 = static_cast<const int64_t *>(z);
      } else {
        void *z = nullptr;
        cudaSetDevice(ctx->device);
        if (cudaMallocAsync(&z, 16, ctx->stream) != cudaSuccess || cudaMemsetAsync(z, 0, 16, ctx->stream) != cudaSuccess) return SB_CUDA;
        p->dev_allocs.push_back(z);
        off64 = static_cast<const int64_t *>(z);
      }
    }
    if (!off64) return SB_INVALID_ARG;
    out->length = int64_t(len);
    p->buffers.push_back(first.nested_validity[depth]);
    if (!first.nested_validity[depth]) out->null_count = 0;
    if (f->large) {
      p->buffers.push_back(off64);
    } else if (first.mem == SB_MEM_HOST) { // ListArray<i32>: create_list narrows the offsets (try_from)
      int32_t *o32 = static_cast<int32_t *>(std::malloc(4 * (len + 1)));
      if (!o32) return SB_CUDA;
      p->host_allocs.push_back(o32);
      for (uint64_t i = 0; i <= len; ++i) {
        if (off64[i] > 0x7fffffffll) return SB_OUT_OF_SPEC;
        o32[i] = int32_t(off64[i]);
      }
      p->buffers.push_back(o32);
    } else {
      int32_t *o32 = nullptr;
      cudaSetDevice(ctx->device);
      if (cudaMallocAsync(reinterpret_cast<void **>(&o32), 4 * (len + 1) + 16, ctx->stream) != cudaSuccess) return SB_CUDA;
      p->dev_allocs.push_back(o32);
      narrow_offsets_kernel<<<uint32_t((len + 1 + 255) / 256), 256, 0, ctx->stream>>>(off64, o32, len + 1);
      p->buffers.push_back(o32);
    }
    ArrowArray *child = new ArrowArray();
    p->children.push_back(child);
    int rc = build_array(ctx, &f->children[0], leaves, cursor, depth + 1, child);
    if (rc) return rc;
  } else { // struct
    if (f->n_children < 1) return SB_INVALID_ARG;
    bool top = depth == 0 && first.nested_len[0] == 0 && first.nested_offsets[0] == nullptr && first.nested_validity[0] == nullptr;
    // a struct's length: its NestedState entry; a top-level struct of flat children has none -- take the child length
    out->length = top ? int64_t(first.length) : int64_t(first.nested_len[depth]);
    p->buffers.push_back(top ? nullptr : first.nested_validity[depth]);
    if (!p->buffers[0]) out->null_count = 0;
    for (int32_t c = 0; c < f->n_children; ++c) {
      ArrowArray *child = new ArrowArray();
      p->children.push_back(child);
      int rc = build_array(ctx, &f->children[c], leaves, cursor, depth + 1, child);
      if (rc) return rc;
    }
  }
  out->n_buffers = int64_t(p->buffers.size());
  out->buffers = p->buffers.empty() ? nullptr : p->buffers.data();
  out->n_children = int64_t(p->children.size());
  out->children = p->children.empty() ? nullptr : p->children.data();
  return SB_OK;
}

} // namespace

extern "C" {

int32_t sb_export_arrow(sb_ctx *ctx, const sb_field *field, sb_column_out *leaves, uint64_t n_leaves, struct ArrowArray *out_array,
                        struct ArrowSchema *out_schema) {
  if (!ctx) return SB_CUDA;
  if (!field || !leaves || !out_array) return fail(ctx, SB_INVALID_ARG, "bad arguments");
  if (count_leaves(field) != n_leaves) return fail(ctx, SB_INVALID_ARG, "the field tree and the decoded leaves disagree on the leaf count");
  for (uint64_t i = 1; i < n_leaves; ++i)
    if (leaves[i].mem != leaves[0].mem) return fail(ctx, SB_INVALID_ARG, "all leaves must live in the same memory space");
  if (out_schema) {
    int rc = build_schema(field, out_schema);
    if (rc) return fail(ctx, rc, "unsupported type in the field tree");
  }
  uint64_t cursor = 0;
  int rc = build_array(ctx, field, leaves, &cursor, 0, out_array);
  if (rc != SB_OK) {
    release_array(out_array);
    if (out_schema) release_schema(out_schema);
    return fail(ctx, rc, "nested assembly failed (NestedState of the first leaf does not match the field tree)");
  }
  if (leaves[0].mem == SB_MEM_DEVICE) cudaStreamSynchronize(ctx->stream); // narrowed offsets are ready when the call returns
  // ownership of the decoded buffers moves into the root array: its release callback gives them back to the context
  NodePriv *p = static_cast<NodePriv *>(out_array->private_data);
  p->owned.assign(leaves, leaves + n_leaves);
  for (uint64_t i = 0; i < n_leaves; ++i) std::memset(&leaves[i], 0, sizeof(leaves[i]));
  return SB_OK;
}

} // extern "C"
