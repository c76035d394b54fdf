/*
 * sb_oracle.h -- C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a CPU restatement of the page
 * encode/decode path of sundy-li/strawboat (src/compression, src/read/read_basic.rs,
 * src/read/array, src/write/{serialize,primitive,binary,boolean}.rs).  It exists so
 * that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs can check and time the reference algorithm.  The product (strawboat_b200/) never
 * links, imports or calls it.
 *
 * PARITY STATUS: the Rust reference cannot be built in this image (no cargo/rustc), and
 * the reference repo ships no golden byte fixtures.  The restatement is pinned against
 *   - the one byte-level known-answer test upstream has (patas pack/unpack,
 *     src/compression/double/patas.rs:191-202),
 *   - the hand-derived vectors K1..K8 of SURVEY.md Appendix A.6,
 *   - independent implementations present in the image for the third-party byte
 *     layouts: liblz4 1.9.4 / pyarrow lz4_raw (LZ4 block), pyarrow's parquet writer
 *     (hybrid-RLE level streams), see tests/test_oracle_*.py.
 * For bitpacking::BitPacker4x and roaring::RoaringBitmap::serialize_into (crates whose
 * source is absent) parity is UNPINNED: the layouts are restated from their published
 * format descriptions (oracle/FORMAT_ASSUMPTIONS.md).
 */
#ifndef SB_ORACLE_H
#define SB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Physical leaf types (arrow2 PhysicalType / PrimitiveType as dispatched by
 * src/write/serialize.rs:52-132 and src/write/primitive.rs:30-96). Utf8 == Binary,
 * LargeUtf8 == LargeBinary on this path (serialize.rs:99-105). */
enum {
  SBO_NULL = 0,
  SBO_BOOL = 1,
  SBO_I8 = 2,
  SBO_I16 = 3,
  SBO_I32 = 4,
  SBO_I64 = 5,
  SBO_U8 = 6,
  SBO_U16 = 7,
  SBO_U32 = 8,
  SBO_U64 = 9,
  SBO_F32 = 10,
  SBO_F64 = 11,
  SBO_BINARY = 12,       /* i32 offsets */
  SBO_LARGE_BINARY = 13, /* i64 offsets */
  SBO_I128 = 14,         /* Decimal128 storage */
  SBO_I256 = 15,         /* Decimal256 storage */
};

/* Codec ids, src/compression/mod.rs:37-108 */
enum {
  SBO_C_NONE = 0,
  SBO_C_LZ4 = 1,
  SBO_C_ZSTD = 2,
  SBO_C_SNAPPY = 3,
  SBO_C_RLE = 10,
  SBO_C_DICT = 11,
  SBO_C_ONEVALUE = 12,
  SBO_C_FREQ = 13,
  SBO_C_BITPACK = 14,
  SBO_C_DELTABP = 15,
  SBO_C_PATAS = 16,
};

/* Status codes (map of arrow::error::Error variants, src/errors.rs) */
enum {
  SBO_OK = 0,
  SBO_OUT_OF_SPEC = 1,
  SBO_IO = 2, /* short read */
  SBO_EXTERNAL = 3,
  SBO_NYI = 4,
  SBO_PANIC = 5, /* the reference would panic here (unwrap / assert / slice OOB) */
};

/* Nested descriptor, root -> leaf (arrow2 InitNested, src/read/deserialize.rs:154-221) */
enum { SBO_N_PRIMITIVE = 0, SBO_N_LIST = 1, SBO_N_STRUCT = 2 };
#define SBO_MAX_NESTED 8

typedef struct {
  int32_t type;     /* SBO_* physical type */
  int32_t nullable; /* Field.is_nullable of the leaf (flat) */
  int32_t n_nested; /* 0 or 1 => flat column; else number of InitNested entries */
  int32_t nested_kind[SBO_MAX_NESTED];
  int32_t nested_nullable[SBO_MAX_NESTED];
} sbo_leaf;

/* WriteOptions, src/write/common.rs:37-45, plus the deterministic stand-ins for the
 * reference's non-deterministic inputs (thread_rng, HashMap order, debug env switches). */
typedef struct {
  int32_t default_compression;   /* SBO_C_NONE..SBO_C_SNAPPY */
  double default_compress_ratio; /* < 0  => None (adaptive off) */
  uint32_t forbidden_mask;       /* bit c set => codec id c forbidden */
  int32_t force_codec;           /* -1, or codec id honoured like util/env.rs switches */
  uint64_t seed;                 /* sampler seed (stands in for thread_rng) */
  int32_t float_bitwise;         /* 0: OrderedFloat equality (reference); 1: bit equality */
} sbo_opts;

/* One leaf array slice handed to the page writer. */
typedef struct {
  const void *values;         /* primitives: n*W bytes; bool: bitmap; binary: value bytes base */
  int64_t values_bit_offset;  /* bool only */
  const void *offsets;        /* binary: n+1 offsets (i32 / i64), absolute into values */
  int64_t values_backing_len; /* binary: array.values().len() (whole backing buffer) */
  const uint8_t *validity;    /* may be NULL */
  int64_t validity_offset;    /* bit offset into validity */
  int64_t n;                  /* rows */
} sbo_array;

typedef struct {
  uint8_t *data;
  size_t len;
  size_t cap;
} sbo_buf;
void sbo_buf_free(sbo_buf *b);

const char *sbo_last_error(void);

/* ---- encode ------------------------------------------------------------------ */
/* write::write for a flat leaf (write_simple, src/write/serialize.rs:52-132): optional
 * validity section then the value block; appends to out. */
int sbo_write_page(const sbo_leaf *leaf, const sbo_array *arr, const sbo_opts *opts, sbo_buf *out);
/* compress_{integer,double,binary,boolean}: value block only; appends to out. */
int sbo_compress_values(int32_t type, const sbo_array *arr, const sbo_opts *opts, sbo_buf *out);
/* write_validity (serialize.rs:200-215). */
int sbo_write_validity(const uint8_t *validity, int64_t bit_offset, int64_t n, sbo_buf *out);

/* ---- decode ------------------------------------------------------------------ */
/* Column builder = the Vec<T> / MutableBitmap / offsets+values that read_integer,
 * read_double, read_binary, read_boolean append into (src/read/array batch forms). */
typedef struct sbo_col sbo_col;
sbo_col *sbo_col_new(const sbo_leaf *leaf);
void sbo_col_free(sbo_col *c);
/* One iteration of the page loop of read_* (read_validity + decompress_*). */
int sbo_col_read_page(sbo_col *c, const uint8_t *page, size_t len, uint64_t num_values);
/* whole column body, page loop inside the library (metas = (length, num_values) pairs) */
int sbo_col_read_pages(sbo_col *c, const uint8_t *body, size_t nbytes, const uint64_t *metas, size_t n_pages);
/* page loop of encode_chunk for one fixed-width flat leaf */
int sbo_write_column(const sbo_leaf *leaf, const sbo_array *arr, const sbo_opts *opts, uint64_t page_rows, sbo_buf *out,
                     uint64_t *metas_out, size_t metas_cap, size_t *n_pages_out);
int64_t sbo_col_len(const sbo_col *c);
const uint8_t *sbo_col_values(const sbo_col *c, size_t *nbytes);
const uint8_t *sbo_col_offsets(const sbo_col *c, size_t *nbytes);
const uint8_t *sbo_col_validity(const sbo_col *c, size_t *nbits);
/* nested: per depth offsets / validity of the NestedState accumulated over pages */
int32_t sbo_col_nested_depths(const sbo_col *c);
const int64_t *sbo_col_nested_offsets(const sbo_col *c, int32_t depth, size_t *n);
const uint8_t *sbo_col_nested_validity(const sbo_col *c, int32_t depth, size_t *nbits);

/* stat.rs-style codec tree of a value block, e.g. "Dict(Bitpacking)[k=8]". */
int sbo_stat_block(int32_t type, const uint8_t *block, size_t len, char *out, size_t out_cap);
/* offset of the value block inside a flat page (skips the validity section). */
int64_t sbo_page_value_block_offset(const sbo_leaf *leaf, const uint8_t *page, size_t len);

/* ---- third-party layouts exposed for pinning tests ----------------------------- */
uint32_t sbo_bp4x_num_bits(const uint32_t *block128);
size_t sbo_bp4x_compress(const uint32_t *in128, uint8_t *out, uint32_t num_bits);
size_t sbo_bp4x_decompress(const uint8_t *in, uint32_t *out128, uint32_t num_bits);
size_t sbo_bp4x_compress_sorted(uint32_t initial, const uint32_t *in128, uint8_t *out, uint32_t num_bits);
size_t sbo_bp4x_decompress_sorted(uint32_t initial, const uint8_t *in, uint32_t *out128, uint32_t num_bits);
int sbo_roaring_serialize(const uint32_t *sorted_vals, size_t n, sbo_buf *out);
int sbo_roaring_deserialize(const uint8_t *in, size_t len, sbo_buf *out_u32);
/* CommonCompression::{compress,decompress} (basic.rs:62-152) for one buffer */
int sbo_common_compress(int32_t codec, const uint8_t *in, size_t in_len, sbo_buf *out);
int sbo_common_decompress(int32_t codec, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len);
int sbo_lz4_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len); /* own block decoder */
int sbo_lz4_decompress_lib(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len); /* liblz4 */
int sbo_lz4_compress_lib(const uint8_t *in, size_t in_len, sbo_buf *out);
uint16_t sbo_patas_pack(uint8_t reference_index, uint8_t significant_bytes, uint8_t trailing_zeros);
void sbo_patas_unpack(uint16_t packed, uint8_t *out3);
/* hybrid-RLE level stream decode (parquet2 HybridRleDecoder) */
int sbo_hybrid_rle_decode(const uint8_t *in, size_t len, uint32_t bit_width, size_t n, uint32_t *out);
/* nested level encode: rep/def streams for a page (arrow2 write_rep_and_def V2) */
int sbo_levels_encode(const uint32_t *levels, size_t n, uint32_t bit_width, sbo_buf *out);

/* sampler: the deterministic stand-in for rand::thread_rng().gen_range(0..range_end) at
 * src/compression/integer/mod.rs:332.  Shared definition with the GPU chooser. */
uint64_t sbo_sample_draw(uint64_t seed, uint32_t codec, uint32_t sample_i, uint64_t range_end);

/* ---- file framing (src/write/writer.rs, src/read/reader.rs) --------------------- */
typedef struct {
  uint64_t length;
  uint64_t num_values;
} sbo_page_meta;

#ifdef __cplusplus
}
#endif
#endif
