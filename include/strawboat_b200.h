/*
 * strawboat_b200.h -- C ABI of the B200-native strawboat page encode/decode backend.
 *
 * This is the drop-in boundary for ONE hot path of sundy-li/strawboat: page bytes <-> Arrow
 * buffers (src/compression + the decompress-into-Arrow half of src/read + the per-page half
 * of src/write).  Every entry point names the reference interface it replaces (paths are
 * relative to the reference repo root).  Plain pointers and sizes only; no torch, no C++
 * types.  The shared library is strawboat_b200/csrc/libstrawboat_b200.so (sm_100a SASS).
 *
 * There is NO CPU fallback: every compute entry point fails with SB_CUDA when no usable
 * CUDA device is present.
 */
#ifndef STRAWBOAT_B200_H
#define STRAWBOAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status: arrow::error::Error variants used on the path (src/errors.rs:19-31,
 *      src/compression/mod.rs:78-80, src/compression/basic.rs:104,116) ---------------- */
enum {
  SB_OK = 0,
  SB_OUT_OF_SPEC = 1, /* Error::OutOfSpec */
  SB_IO = 2,          /* Error::Io (short read: "failed to fill whole buffer") */
  SB_EXTERNAL = 3,    /* Error::External (LZ4 block corrupt) */
  SB_NYI = 4,         /* Error::NotYetImplemented (zstd / snappy pages, f16, ...) */
  SB_CUDA = 5,        /* CUDA runtime failure or no device */
  SB_INVALID_ARG = 6,
  SB_PANIC = 7,       /* input on which the reference panics (unwrap / assert / OOB slice) */
};

/* ---- physical leaf types: arrow2 PhysicalType as dispatched by
 *      src/write/serialize.rs:52-132 and src/read/deserialize.rs:100-138.
 *      Utf8 is handled as Binary, LargeUtf8 as LargeBinary (serialize.rs:99-105). -------- */
enum {
  SB_NULL = 0,
  SB_BOOL = 1,
  SB_I8 = 2,
  SB_I16 = 3,
  SB_I32 = 4,
  SB_I64 = 5,
  SB_U8 = 6,
  SB_U16 = 7,
  SB_U32 = 8,
  SB_U64 = 9,
  SB_F32 = 10,
  SB_F64 = 11,
  SB_BINARY = 12,       /* i32 offsets */
  SB_LARGE_BINARY = 13, /* i64 offsets */
  SB_I128 = 14,         /* PrimitiveType::Int128 (Decimal128 storage): 16-byte little-endian two's complement */
  SB_I256 = 15,         /* PrimitiveType::Int256 (Decimal256 storage): 32 bytes */
};

/* ---- codec ids: src/compression/mod.rs:37-108 --------------------------------------- */
enum {
  SB_C_NONE = 0,
  SB_C_LZ4 = 1,
  SB_C_ZSTD = 2,
  SB_C_SNAPPY = 3,
  SB_C_RLE = 10,
  SB_C_DICT = 11,
  SB_C_ONEVALUE = 12,
  SB_C_FREQ = 13,
  SB_C_BITPACK = 14,
  SB_C_DELTABP = 15,
  SB_C_PATAS = 16,
};

enum { SB_MEM_HOST = 0, SB_MEM_DEVICE = 1 };

/* arrow2 InitNested, root -> leaf (src/read/deserialize.rs:154-221) */
enum { SB_N_PRIMITIVE = 0, SB_N_LIST = 1, SB_N_STRUCT = 2 };
#define SB_MAX_NESTED 8

/* PageMeta (src/lib.rs:71-80): length = encoded bytes of the page, num_values = rows
 * (flat) or level entries (nested, src/write/common.rs:103). */
typedef struct {
  uint64_t length;
  uint64_t num_values;
} sb_page_meta;

/* What the reference derives from (Field, ColumnDescriptor, Vec<InitNested>) for one leaf
 * (src/read/deserialize.rs:100-234). n_nested <= 1 means a flat column. */
typedef struct {
  int32_t type;
  int32_t nullable;
  int32_t n_nested;
  int32_t nested_kind[SB_MAX_NESTED];
  int32_t nested_nullable[SB_MAX_NESTED];
} sb_leaf;

typedef struct sb_ctx sb_ctx;

/* One context per host thread (the reference's readers/writers are single-threaded objects
 * with Send+Sync bounds, src/read/deserialize.rs:28).  Owns device scratch pools and a
 * stream; all work of a call is ordered on that stream and complete on return. */
int32_t sb_ctx_create(int32_t device, sb_ctx **out);
void sb_ctx_destroy(sb_ctx *ctx);
/* Use the caller's CUDA stream (cudaStream_t) instead of the context's own. */
int32_t sb_ctx_set_stream(sb_ctx *ctx, void *cuda_stream);
const char *sb_last_error(const sb_ctx *ctx);
const char *sb_version(void);

/* ------------------------------------------------------------------------------------
 * DECODE
 * ------------------------------------------------------------------------------------ */

/* All pages of one leaf column, back to back, exactly as they sit in the file between
 * ColumnMeta.offset and ColumnMeta.offset + total_len() (src/lib.rs:40-68) -- i.e. what a
 * NativeReadBuf positioned at the column yields (src/read/mod.rs:26-28). */
typedef struct {
  sb_leaf leaf;
  const uint8_t *bytes;
  uint64_t nbytes;
  int32_t mem; /* SB_MEM_HOST or SB_MEM_DEVICE: where `bytes` lives */
  const sb_page_meta *metas; /* host */
  uint64_t n_pages;
} sb_column_in;

/* One decoded Arrow array (the buffers of PrimitiveArray / BooleanArray / BinaryArray /
 * Utf8Array that src/read/array/{integer,double,boolean,binary}.rs build), plus -- for
 * nested leaves -- the NestedState of read_validity_nested (src/read/read_basic.rs:65-173).
 * Pointers are device pointers (out_mem == SB_MEM_DEVICE) or host pointers owned by the
 * context (SB_MEM_HOST); release with sb_release_columns. */
typedef struct {
  uint64_t length;         /* rows (flat) / leaf slots (nested) */
  void *values;            /* primitives: length*W bytes; bool: bitmap; binary: value bytes */
  uint64_t values_bytes;
  void *offsets;           /* binary: (length+1) i32 / i64 */
  uint64_t offsets_bytes;
  uint8_t *validity;       /* LSB-first bitmap, NULL when the leaf is not nullable */
  uint64_t validity_bytes;
  /* nested leaves only: per depth d < n_nested-1, the entries pushed into NestedState */
  int64_t *nested_offsets[SB_MAX_NESTED];  /* list start offsets (+ final end offset) */
  uint8_t *nested_validity[SB_MAX_NESTED]; /* bitmap, NULL if depth not nullable */
  uint64_t nested_len[SB_MAX_NESTED];      /* entries at that depth */
  int32_t *page_status;    /* host, n_pages entries: SB_OK or the page's error */
  int32_t mem;
  void *_owner;
} sb_column_out;

/* Replaces read::batch_read::batch_read_array for leaf columns
 * (src/read/batch_read.rs:190-209 -> read_integer / read_double / read_binary /
 * read_boolean / read_null, src/read/array/integer.rs:210-238 etc.): decodes every page of
 * every given column in one batched launch sequence and returns ONE concatenated array per
 * column.  Returns SB_OK if every page decoded; otherwise the first failing page's status
 * (per-page detail in page_status; the other pages are still decoded). */
int32_t sb_decode_columns(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem,
                          sb_column_out *outs);

/* Replaces read::deserialize::column_iter_to_arrays(...).next() / IntegerIter::deserialize
 * (src/read/deserialize.rs:237-253, src/read/array/integer.rs:68-88): one array PER PAGE.
 * `pages[i]` is a one-page column (n_pages == 1). Same batching on the device. */
int32_t sb_decode_pages(sb_ctx *ctx, const sb_column_in *pages, uint64_t n_pages, int32_t out_mem,
                        sb_column_out *outs);

void sb_release_columns(sb_ctx *ctx, sb_column_out *outs, uint64_t n);

/* ---- ownership and overlap: the rest of the reader boundary (SURVEY §8b "Ownership") ----------------------
 * The reference moves freshly allocated Vec<T> / MutableBitmap into arrow2 Buffers (src/read/array/integer.rs:86)
 * and recycles page buffers through PageIterator::swap_buffer (src/read/mod.rs:55-57).  Here the caller may own
 * the output memory instead (plan -> allocate -> run), and a call may be left in flight while the host prepares
 * the next one. */

/* Caller-owned output buffers of one FLAT column (nested leaves stay library-allocated: their NestedState sizes
 * are only known after the level pass).  A NULL pointer = let the library allocate that buffer.  With
 * out_mem = SB_MEM_DEVICE the pointers are device memory, 16-byte aligned, and the decoders write into them
 * directly; with SB_MEM_HOST they are host memory (pinned, to keep the copy asynchronous) filled by the D2H copy.
 * Buffers handed in are never freed by sb_release_columns. */
typedef struct {
  void *values;
  uint64_t values_cap;
  void *offsets;
  uint64_t offsets_cap;
  uint8_t *validity;
  uint64_t validity_cap; /* bitmaps are written in 4-byte words: round the capacity up to a multiple of 4 */
} sb_out_buffers;

/* Buffer sizes of every column: exact, including the data-dependent value bytes of binary / utf8 columns (runs
 * the size pass over their pages) and the leaf slots of nested leaves. */
typedef struct {
  uint64_t length;
  uint64_t values_bytes;
  uint64_t offsets_bytes;
  uint64_t validity_bytes;
} sb_column_sizes;
int32_t sb_plan_columns(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, sb_column_sizes *sizes);

/* sb_decode_columns into caller-owned buffers (`bufs`: n_cols entries, or NULL). */
int32_t sb_decode_columns_into(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem, const sb_out_buffers *bufs,
                               sb_column_out *outs);

/* Asynchronous form: submits every copy and kernel of the call on the context's stream and returns.  `cols`, the
 * page bytes, `bufs` and `outs` must stay valid until sb_decode_wait(ctx), which blocks until the work is done,
 * fills page_status / statistics and returns what sb_decode_columns would have returned.  One call may be in
 * flight per context (submitting another one, or encoding, collects it first); several contexts driven by one host
 * thread overlap their H2D copies, kernels and D2H copies.  Binary / nested columns still block inside the submit
 * for their size pass.  sb_decode_ready: 1 when sb_decode_wait would not block. */
int32_t sb_decode_columns_async(sb_ctx *ctx, const sb_column_in *cols, uint64_t n_cols, int32_t out_mem, const sb_out_buffers *bufs,
                                sb_column_out *outs);
int32_t sb_decode_wait(sb_ctx *ctx);
int32_t sb_decode_ready(sb_ctx *ctx);

/* Counters of the last decode/encode call on this context. */
typedef struct {
  uint64_t pages;
  uint64_t bytes_in;       /* sum of PageMeta.length */
  uint64_t bytes_out;      /* Arrow bytes produced (values + offsets + validity) */
  uint64_t kernel_launches;
  float device_ms;         /* CUDA-event time of the kernels of the last call */
  uint64_t codec_pages[32]; /* pages per top-level codec id */
  float main_kernel_ms;    /* decode: sb_decode_kernel (pass 1: binary, boolean, nested, LZ4-in-tree, Freq, Patas pages) alone */
  float lz4_kernel_ms;     /* decode: sb_lz4_kernel alone (runs concurrently with the main kernel) */
  uint64_t lz4_bytes;      /* decode: compressed + decoded bytes of the blocks sb_lz4_kernel handled */
  float host_ms;           /* wall time of the whole call on the host */
  float light_kernel_ms;   /* decode: sb_decode_light_kernel alone (flat fixed-width pages with light codec trees; concurrent) */
} sb_stats;
int32_t sb_last_stats(const sb_ctx *ctx, sb_stats *out);

/* ------------------------------------------------------------------------------------
 * PAGE INSPECTOR (host only: walks headers, never touches the device)
 * ------------------------------------------------------------------------------------ */

/* What stat_body collects for one page (src/stat.rs:36-61,86-152): PageInfo + the codec tree
 * of the nested sub-pages (Dict -> index page, Freq -> exceptions page of primitive columns). */
typedef struct {
  int32_t codec;                   /* top-level Compression id of the value block */
  uint32_t validity_size;          /* flat nullable leaves: bytes of the validity section after its u32
                                      length; 0xffffffff otherwise.  (The reference prints the first 4
                                      bytes of the value block here, src/stat.rs:76 -- not replicated.) */
  uint32_t levels_size;            /* nested leaves: 12 + rep_len + def_len; 0 otherwise */
  uint32_t compressed_size;        /* hdr9 */
  uint32_t uncompressed_size;      /* hdr9 */
  uint32_t unique_num;             /* Dict: entries in the dictionary; else 0 */
  uint32_t exceptions_bitmap_size; /* Freq: bytes of the Roaring bitmap; else 0 */
  int32_t depth;                   /* codecs on the path top -> innermost sub-page */
  int32_t path[4];                 /* e.g. {Dict, Bitpacking}, {Freq, Lz4}, {Dict, Freq, Bitpacking} */
} sb_page_info;

/* Replaces stat::stat_simple's per-page body (src/stat.rs:63-152) for one page of a leaf column.
 * `tree` (optional, NUL-terminated, at most tree_cap bytes) receives the codec tree as text, e.g.
 * "Dict(Bitpacking)[k=8]".  Errors: SB_IO (truncated page), SB_OUT_OF_SPEC (unknown codec id). */
int32_t sb_stat_page(const sb_leaf *leaf, const uint8_t *page, uint64_t len, sb_page_info *info,
                     char *tree, uint64_t tree_cap);

/* ------------------------------------------------------------------------------------
 * ENCODE
 * ------------------------------------------------------------------------------------ */

/* WriteOptions (src/write/common.rs:37-45) + the explicit stand-ins for the reference's
 * hidden inputs: the sampler seed (thread_rng, src/compression/integer/mod.rs:316) and the
 * force-codec switch (debug env vars, src/util/env.rs:20-24). */
typedef struct {
  int32_t default_compression;   /* SB_C_NONE / SB_C_LZ4 (zstd, snappy: SB_NYI) */
  double default_compress_ratio; /* < 0 => None: adaptive compression off */
  uint64_t max_page_size;        /* rows per page; 0 => None: one page per column */
  uint32_t forbidden_mask;       /* bit c => Compression id c in forbidden_compressions */
  int32_t force_codec;           /* -1 or a codec id, honoured where applicable */
  uint64_t seed;                 /* page p of a column samples with seed + p */
} sb_write_options;

/* One leaf array (what to_leaves yields, src/write/common.rs:68). */
typedef struct {
  sb_leaf leaf;
  uint64_t length;
  const void *values;          /* primitives / bool bitmap (bit offset 0) / binary value bytes */
  uint64_t values_bytes;
  const void *offsets;         /* binary: length+1 offsets */
  const uint8_t *validity;     /* may be NULL */
  int32_t mem;
  /* Nested leaves (leaf.n_nested > 1) only; all zero for flat leaves.  The Dremel levels of
   * the whole column in row order -- what arrow2's write_rep_and_def iterates over the
   * `Nested` descriptors (src/write/serialize.rs:217-232).  `length` is then the number of
   * LEAF SLOTS, `rows` the number of top-level rows pages are cut by (write/common.rs:79-86);
   * rep_levels may be NULL when the column has no list depth (max_rep == 0). */
  const uint32_t *rep_levels;
  const uint32_t *def_levels;
  uint64_t n_levels;
  uint64_t rows;
} sb_leaf_array;

typedef struct {
  uint8_t *bytes;              /* encoded pages back to back (the column body in the file) */
  uint64_t nbytes;
  sb_page_meta *metas;         /* host, n_pages */
  uint64_t n_pages;
  int32_t mem;
  void *_owner;
} sb_encoded_column;

/* Replaces the page loop of NativeWriter::encode_chunk
 * (src/write/common.rs:71-115 -> write::write, src/write/serialize.rs:36-132 ->
 * compress_integer / compress_double / compress_binary / compress_boolean): slices each
 * leaf into pages of max_page_size rows, chooses a codec per page as choose_compressor
 * does, and emits the page bytes plus the PageMeta{length,num_values} the footer needs.
 * Nested leaves (write_nested + write_nested_validity, src/write/serialize.rs:135-198,
 * 217-232): pages are cut by top-level rows; each page is
 * [u32 rows][u32 rep_len][u32 def_len][rep stream][def stream][VALUE_BLOCK over the page's
 * leaf slots], PageMeta.num_values = level entries of the page. */
int32_t sb_encode_columns(sb_ctx *ctx, const sb_leaf_array *cols, uint64_t n_cols,
                          const sb_write_options *opts, int32_t out_mem, sb_encoded_column *outs);
void sb_release_encoded(sb_ctx *ctx, sb_encoded_column *outs, uint64_t n);

/* ------------------------------------------------------------------------------------
 * NESTED ASSEMBLY + ARROW C DATA INTERFACE EXPORT (read side of nested columns; ownership)
 * ------------------------------------------------------------------------------------ */

#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
/* The Arrow C Data Interface (arrow.apache.org/docs/format/CDataInterface.html), ABI-stable definitions */
struct ArrowSchema {
  const char *format;
  const char *name;
  const char *metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema **children;
  struct ArrowSchema *dictionary;
  void (*release)(struct ArrowSchema *);
  void *private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void **buffers;
  struct ArrowArray **children;
  struct ArrowArray *dictionary;
  void (*release)(struct ArrowArray *);
  void *private_data;
};
#endif

/* The Field / DataType tree of one column (what column_iter_to_arrays / batch_read_array receive as `field`,
 * src/read/deserialize.rs:237-245, src/read/batch_read.rs:190-196). */
typedef struct sb_field {
  int32_t kind;      /* SB_N_PRIMITIVE (leaf) / SB_N_LIST / SB_N_STRUCT */
  int32_t type;      /* leaf: SB_* physical type */
  int32_t utf8;      /* leaf SB_BINARY / SB_LARGE_BINARY: export as Utf8 / LargeUtf8 */
  int32_t nullable;
  int32_t large;     /* SB_N_LIST: LargeList (i64 offsets) */
  int32_t n_children;
  const struct sb_field *children;
  const char *name;
} sb_field;

/* Replaces the array construction of the readers: create_list / create_struct over the NestedState of the first
 * leaf (src/read/array/list.rs, struct_.rs; src/read/batch_read.rs:66-187) and PrimitiveArray / BinaryArray /
 * BooleanArray::try_new (src/read/array/integer.rs:86 ...).  `leaves`: the decoded leaves of the field, in leaf
 * order, all from one sb_decode_columns / sb_decode_pages call.  Nothing is copied (List<i32> offsets are narrowed
 * from the i64 NestedState offsets, as create_list does).  OWNERSHIP MOVES: the leaves are zeroed and their buffers
 * live until out_array->release is called (which returns them to `ctx`, so the context must outlive the array).
 * Buffers are host pointers for out_mem = SB_MEM_HOST; for SB_MEM_DEVICE wrap the result in an ArrowDeviceArray
 * {array, device_id, ARROW_DEVICE_CUDA}. */
int32_t sb_export_arrow(sb_ctx *ctx, const sb_field *field, sb_column_out *leaves, uint64_t n_leaves, struct ArrowArray *out_array,
                        struct ArrowSchema *out_schema);

/* ------------------------------------------------------------------------------------
 * NESTED LEVELS from Arrow buffers (write side of nested columns)
 * ------------------------------------------------------------------------------------ */

/* One depth of a nested leaf, root -> leaf: what arrow2's to_nested() yields per leaf (Nested::{List, LargeList,
 * Struct, Primitive}; call site src/write/common.rs:66-68).  Lists must be compact (children in order, a null
 * list has none). */
typedef struct {
  int32_t kind;            /* SB_N_LIST / SB_N_STRUCT / SB_N_PRIMITIVE */
  int32_t nullable;
  const void *offsets;     /* SB_N_LIST: length + 1 offsets */
  int32_t offset_width;    /* 4 (List) or 8 (LargeList) */
  const uint8_t *validity; /* LSB-first bitmap, NULL = all valid */
  uint64_t length;         /* elements at this depth */
} sb_nested_level;

/* Replaces arrow2's write_rep_and_def iteration (src/write/serialize.rs:217-232): the Dremel (rep, def) levels of
 * one leaf, generated on the device from the Arrow offsets / validity buffers of its ancestors.  `mem` says where
 * the input buffers live; the level arrays are DEVICE memory owned by the caller (sb_free_device) and can be
 * passed straight to sb_encode_columns (sb_leaf_array.rep_levels / def_levels with mem = SB_MEM_DEVICE).
 * n_slots = leaf slots the levels describe (sb_leaf_array.length), n_levels = level entries. */
int32_t sb_nested_levels(sb_ctx *ctx, const sb_nested_level *path, int32_t n_depths, int32_t mem, uint32_t **rep_levels, uint32_t **def_levels,
                         uint64_t *n_levels, uint64_t *n_slots);
void sb_free_device(sb_ctx *ctx, void *p);

/* ------------------------------------------------------------------------------------
 * MULTI-GPU ENCODE: gather of the encoded column bodies on the writer rank
 * ------------------------------------------------------------------------------------ */

/* The reference writes a file from one process: NativeWriter::encode_chunk appends every column body at its
 * absolute offset and finish() writes one footer (src/write/common.rs:71-119, src/write/writer.rs:128-167).  When
 * the leaf columns of a chunk are encoded on several GPUs (leaf c on rank c mod world), this is the one exchange
 * step: every rank learns the layout (ncclAllGather of sizes), the bodies travel once over NVLink (grouped
 * ncclSend / ncclRecv) straight to their final position in the writer's staging buffer.  One process per GPU. */
typedef struct sb_comm sb_comm;
/* rank 0 makes the id (ncclGetUniqueId, 128 bytes) and hands it to the other ranks out of band */
int32_t sb_comm_unique_id(uint8_t *id128);
int32_t sb_comm_create(sb_ctx *ctx, int32_t rank, int32_t world, const uint8_t *id128, sb_comm **out);
void sb_comm_destroy(sb_comm *comm);

typedef struct {
  uint64_t bytes_moved; /* bytes this rank sent (or, on the writer, received) over the interconnect */
  uint64_t total_bytes; /* size of the gathered body region (all columns) */
  float gather_ms;      /* CUDA-event time of the grouped send / recv on this rank */
} sb_gather_stats;

/* `local`: this rank's encoded columns (device resident), leaf order: local[k] is leaf rank + k * world.
 * On the writer rank `outs` receives all n_total columns in leaf order: `bytes` point into one device buffer
 * that holds the bodies back to back -- the file's body region -- and `metas` are the PageMeta records for the
 * footer.  Release with sb_release_encoded(ctx, outs, n_total).  Collective: every rank must call it. */
int32_t sb_gather_encoded(sb_ctx *ctx, sb_comm *comm, const sb_encoded_column *local, uint64_t n_local, uint64_t n_total, int32_t writer,
                          sb_encoded_column *outs, sb_gather_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
