"""ctypes declarations for libstrawboat_b200.so (mirror of include/strawboat_b200.h).

The product path has no CPU fallback: if the CUDA library is missing this module raises at
import time, and every compute entry point raises when no CUDA device is usable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libstrawboat_b200.so")

SB_OK, SB_OUT_OF_SPEC, SB_IO, SB_EXTERNAL, SB_NYI, SB_CUDA, SB_INVALID_ARG, SB_PANIC = range(8)
STATUS_NAMES = ["SB_OK", "SB_OUT_OF_SPEC", "SB_IO", "SB_EXTERNAL", "SB_NYI", "SB_CUDA", "SB_INVALID_ARG", "SB_PANIC"]

NULL, BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, BINARY, LARGE_BINARY, I128, I256 = range(16)
C_NONE, C_LZ4, C_ZSTD, C_SNAPPY = 0, 1, 2, 3
C_RLE, C_DICT, C_ONEVALUE, C_FREQ, C_BITPACK, C_DELTABP, C_PATAS = 10, 11, 12, 13, 14, 15, 16
MEM_HOST, MEM_DEVICE = 0, 1
N_PRIMITIVE, N_LIST, N_STRUCT = 0, 1, 2
MAX_NESTED = 8


class PageMeta(C.Structure):
    _fields_ = [("length", C.c_uint64), ("num_values", C.c_uint64)]


class Leaf(C.Structure):
    _fields_ = [("type", C.c_int32), ("nullable", C.c_int32), ("n_nested", C.c_int32),
                ("nested_kind", C.c_int32 * MAX_NESTED), ("nested_nullable", C.c_int32 * MAX_NESTED)]


class ColumnIn(C.Structure):
    _fields_ = [("leaf", Leaf), ("bytes", C.c_void_p), ("nbytes", C.c_uint64), ("mem", C.c_int32),
                ("metas", C.POINTER(PageMeta)), ("n_pages", C.c_uint64)]


class ColumnOut(C.Structure):
    _fields_ = [("length", C.c_uint64), ("values", C.c_void_p), ("values_bytes", C.c_uint64),
                ("offsets", C.c_void_p), ("offsets_bytes", C.c_uint64),
                ("validity", C.c_void_p), ("validity_bytes", C.c_uint64),
                ("nested_offsets", C.c_void_p * MAX_NESTED), ("nested_validity", C.c_void_p * MAX_NESTED),
                ("nested_len", C.c_uint64 * MAX_NESTED),
                ("page_status", C.POINTER(C.c_int32)), ("mem", C.c_int32), ("_owner", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("pages", C.c_uint64), ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_float), ("codec_pages", C.c_uint64 * 32),
                ("main_kernel_ms", C.c_float), ("lz4_kernel_ms", C.c_float),
                ("lz4_bytes", C.c_uint64), ("host_ms", C.c_float), ("light_kernel_ms", C.c_float)]


class WriteOptions(C.Structure):
    _fields_ = [("default_compression", C.c_int32), ("default_compress_ratio", C.c_double),
                ("max_page_size", C.c_uint64), ("forbidden_mask", C.c_uint32), ("force_codec", C.c_int32),
                ("seed", C.c_uint64)]


class LeafArray(C.Structure):
    _fields_ = [("leaf", Leaf), ("length", C.c_uint64), ("values", C.c_void_p), ("values_bytes", C.c_uint64),
                ("offsets", C.c_void_p), ("validity", C.c_void_p), ("mem", C.c_int32),
                ("rep_levels", C.c_void_p), ("def_levels", C.c_void_p), ("n_levels", C.c_uint64), ("rows", C.c_uint64)]


class PageInfo(C.Structure):
    _fields_ = [("codec", C.c_int32), ("validity_size", C.c_uint32), ("levels_size", C.c_uint32),
                ("compressed_size", C.c_uint32), ("uncompressed_size", C.c_uint32), ("unique_num", C.c_uint32),
                ("exceptions_bitmap_size", C.c_uint32), ("depth", C.c_int32), ("path", C.c_int32 * 4)]


class EncodedColumn(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("nbytes", C.c_uint64), ("metas", C.POINTER(PageMeta)),
                ("n_pages", C.c_uint64), ("mem", C.c_int32), ("_owner", C.c_void_p)]


class OutBuffers(C.Structure):
    _fields_ = [("values", C.c_void_p), ("values_cap", C.c_uint64), ("offsets", C.c_void_p), ("offsets_cap", C.c_uint64),
                ("validity", C.c_void_p), ("validity_cap", C.c_uint64)]


class ColumnSizes(C.Structure):
    _fields_ = [("length", C.c_uint64), ("values_bytes", C.c_uint64), ("offsets_bytes", C.c_uint64), ("validity_bytes", C.c_uint64)]


class GatherStats(C.Structure):
    _fields_ = [("bytes_moved", C.c_uint64), ("total_bytes", C.c_uint64), ("gather_ms", C.c_float)]


class NestedLevel(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nullable", C.c_int32), ("offsets", C.c_void_p), ("offset_width", C.c_int32),
                ("validity", C.c_void_p), ("length", C.c_uint64)]


class Field(C.Structure):
    pass


Field._fields_ = [("kind", C.c_int32), ("type", C.c_int32), ("utf8", C.c_int32), ("nullable", C.c_int32), ("large", C.c_int32),
                  ("n_children", C.c_int32), ("children", C.POINTER(Field)), ("name", C.c_char_p)]


class ArrowSchema(C.Structure):
    _fields_ = [("format", C.c_char_p), ("name", C.c_char_p), ("metadata", C.c_char_p), ("flags", C.c_int64), ("n_children", C.c_int64),
                ("children", C.c_void_p), ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


class ArrowArray(C.Structure):
    _fields_ = [("length", C.c_int64), ("null_count", C.c_int64), ("offset", C.c_int64), ("n_buffers", C.c_int64), ("n_children", C.c_int64),
                ("buffers", C.c_void_p), ("children", C.c_void_p), ("dictionary", C.c_void_p), ("release", C.c_void_p), ("private_data", C.c_void_p)]


# every symbol include/strawboat_b200.h declares
EXPORTS = ["sb_ctx_create", "sb_ctx_destroy", "sb_ctx_set_stream", "sb_last_error", "sb_version",
           "sb_decode_columns", "sb_decode_pages", "sb_release_columns", "sb_last_stats",
           "sb_encode_columns", "sb_release_encoded", "sb_stat_page",
           "sb_comm_unique_id", "sb_comm_create", "sb_comm_destroy", "sb_gather_encoded",
           "sb_nested_levels", "sb_free_device", "sb_export_arrow",
           "sb_plan_columns", "sb_decode_columns_into", "sb_decode_columns_async", "sb_decode_wait", "sb_decode_ready"]


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(strawboat_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.sb_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    L.sb_ctx_create.restype = C.c_int32
    L.sb_ctx_destroy.argtypes = [C.c_void_p]
    L.sb_ctx_destroy.restype = None
    L.sb_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.sb_ctx_set_stream.restype = C.c_int32
    L.sb_last_error.argtypes = [C.c_void_p]
    L.sb_last_error.restype = C.c_char_p
    L.sb_version.restype = C.c_char_p
    L.sb_decode_columns.argtypes = [C.c_void_p, C.POINTER(ColumnIn), C.c_uint64, C.c_int32, C.POINTER(ColumnOut)]
    L.sb_decode_columns.restype = C.c_int32
    L.sb_decode_pages.argtypes = [C.c_void_p, C.POINTER(ColumnIn), C.c_uint64, C.c_int32, C.POINTER(ColumnOut)]
    L.sb_decode_pages.restype = C.c_int32
    L.sb_release_columns.argtypes = [C.c_void_p, C.POINTER(ColumnOut), C.c_uint64]
    L.sb_release_columns.restype = None
    L.sb_last_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.sb_last_stats.restype = C.c_int32
    L.sb_encode_columns.argtypes = [C.c_void_p, C.POINTER(LeafArray), C.c_uint64, C.POINTER(WriteOptions), C.c_int32,
                                    C.POINTER(EncodedColumn)]
    L.sb_encode_columns.restype = C.c_int32
    L.sb_release_encoded.argtypes = [C.c_void_p, C.POINTER(EncodedColumn), C.c_uint64]
    L.sb_release_encoded.restype = None
    L.sb_stat_page.argtypes = [C.POINTER(Leaf), C.c_char_p, C.c_uint64, C.POINTER(PageInfo), C.c_char_p, C.c_uint64]
    L.sb_stat_page.restype = C.c_int32
    L.sb_comm_unique_id.argtypes = [C.c_char_p]
    L.sb_comm_unique_id.restype = C.c_int32
    L.sb_comm_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.POINTER(C.c_void_p)]
    L.sb_comm_create.restype = C.c_int32
    L.sb_comm_destroy.argtypes = [C.c_void_p]
    L.sb_comm_destroy.restype = None
    L.sb_gather_encoded.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(EncodedColumn), C.c_uint64, C.c_uint64, C.c_int32,
                                    C.POINTER(EncodedColumn), C.POINTER(GatherStats)]
    L.sb_gather_encoded.restype = C.c_int32
    L.sb_nested_levels.argtypes = [C.c_void_p, C.POINTER(NestedLevel), C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.sb_nested_levels.restype = C.c_int32
    L.sb_free_device.argtypes = [C.c_void_p, C.c_void_p]
    L.sb_free_device.restype = None
    L.sb_export_arrow.argtypes = [C.c_void_p, C.POINTER(Field), C.POINTER(ColumnOut), C.c_uint64, C.POINTER(ArrowArray), C.POINTER(ArrowSchema)]
    L.sb_export_arrow.restype = C.c_int32
    L.sb_plan_columns.argtypes = [C.c_void_p, C.POINTER(ColumnIn), C.c_uint64, C.POINTER(ColumnSizes)]
    L.sb_plan_columns.restype = C.c_int32
    for f in (L.sb_decode_columns_into, L.sb_decode_columns_async):
        f.argtypes = [C.c_void_p, C.POINTER(ColumnIn), C.c_uint64, C.c_int32, C.POINTER(OutBuffers), C.POINTER(ColumnOut)]
        f.restype = C.c_int32
    L.sb_decode_wait.argtypes = [C.c_void_p]
    L.sb_decode_wait.restype = C.c_int32
    L.sb_decode_ready.argtypes = [C.c_void_p]
    L.sb_decode_ready.restype = C.c_int32
    return L
