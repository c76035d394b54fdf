"""ctypes binding of the CPU oracle (oracle/libsb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (strawboat_b200/) never imports
this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsb_oracle.so")

# physical types (same numbering as include/strawboat_b200.h)
NULL, BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, BINARY, LARGE_BINARY, I128, I256 = range(16)
# codecs
C_NONE, C_LZ4, C_ZSTD, C_SNAPPY = 0, 1, 2, 3
C_RLE, C_DICT, C_ONEVALUE, C_FREQ, C_BITPACK, C_DELTABP, C_PATAS = 10, 11, 12, 13, 14, 15, 16
N_PRIMITIVE, N_LIST, N_STRUCT = 0, 1, 2

NP_OF = {I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, U8: np.uint8, U16: np.uint16,
         U32: np.uint32, U64: np.uint64, F32: np.float32, F64: np.float64,
         I128: np.dtype("V16"), I256: np.dtype("V32")}
WIDTH = {t: np.dtype(d).itemsize for t, d in NP_OF.items()}


class Leaf(C.Structure):
    _fields_ = [("type", C.c_int32), ("nullable", C.c_int32), ("n_nested", C.c_int32),
                ("nested_kind", C.c_int32 * 8), ("nested_nullable", C.c_int32 * 8)]


class Opts(C.Structure):
    _fields_ = [("default_compression", C.c_int32), ("default_compress_ratio", C.c_double),
                ("forbidden_mask", C.c_uint32), ("force_codec", C.c_int32), ("seed", C.c_uint64),
                ("float_bitwise", C.c_int32)]


class Array(C.Structure):
    _fields_ = [("values", C.c_void_p), ("values_bit_offset", C.c_int64), ("offsets", C.c_void_p),
                ("values_backing_len", C.c_int64), ("validity", C.c_void_p), ("validity_offset", C.c_int64),
                ("n", C.c_int64)]


class Buf(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("cap", C.c_size_t)]


def _host_arch():
    """identity of this host's CPU (model + ISA flags): the library is built -march=native"""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        keep = [ln for ln in txt.splitlines() if ln.startswith(("model name", "flags"))][:2]
        return hashlib.sha256("\n".join(keep).encode()).hexdigest()[:16]
    except OSError:
        return "unknown"


_STAMP = os.path.join(_HERE, ".build_arch")


def build(force=False):
    stale = not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "sb_oracle.cpp"))
    try:
        other_cpu = open(_STAMP).read().strip() != _host_arch()
    except OSError:
        other_cpu = True
    if force or stale or other_cpu:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
        with open(_STAMP, "w") as f:
            f.write(_host_arch())
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()  # no-op when the library is current and was built on this CPU model
        L = C.CDLL(_SO)
        L.sbo_last_error.restype = C.c_char_p
        L.sbo_col_new.restype = C.c_void_p
        L.sbo_col_free.argtypes = [C.c_void_p]
        L.sbo_col_read_page.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64]
        L.sbo_col_read_pages.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.sbo_write_column.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.sbo_col_len.argtypes = [C.c_void_p]
        L.sbo_col_len.restype = C.c_int64
        for f in ("sbo_col_values", "sbo_col_offsets", "sbo_col_validity"):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
            getattr(L, f).restype = C.c_void_p
        L.sbo_col_nested_depths.argtypes = [C.c_void_p]
        L.sbo_col_nested_offsets.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]
        L.sbo_col_nested_offsets.restype = C.c_void_p
        L.sbo_col_nested_validity.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_size_t)]
        L.sbo_col_nested_validity.restype = C.c_void_p
        L.sbo_page_value_block_offset.restype = C.c_int64
        L.sbo_page_value_block_offset.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sbo_bp4x_compress.restype = C.c_size_t
        L.sbo_bp4x_decompress.restype = C.c_size_t
        L.sbo_bp4x_compress_sorted.restype = C.c_size_t
        L.sbo_bp4x_decompress_sorted.restype = C.c_size_t
        L.sbo_patas_pack.restype = C.c_uint16
        L.sbo_sample_draw.restype = C.c_uint64
        L.sbo_sample_draw.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"oracle status {code}: {msg}")
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().sbo_last_error().decode())


def _take(buf):
    out = C.string_at(buf.data, buf.len) if buf.len else b""
    lib().sbo_buf_free(C.byref(buf))
    return out


def make_leaf(type_, nullable=False, nested=None):
    """nested: list of (kind, nullable) root->leaf (InitNested), or None for a flat column."""
    lf = Leaf()
    lf.type = type_
    lf.nullable = int(bool(nullable))
    if nested:
        lf.n_nested = len(nested)
        for i, (k, nu) in enumerate(nested):
            lf.nested_kind[i] = k
            lf.nested_nullable[i] = int(bool(nu))
    return lf


def make_opts(default_compression=C_NONE, ratio=None, forbidden=(), force=-1, seed=0, float_bitwise=0):
    o = Opts()
    o.default_compression = default_compression
    o.default_compress_ratio = -1.0 if ratio is None else float(ratio)
    m = 0
    for c in forbidden:
        m |= 1 << c
    o.forbidden_mask = m
    o.force_codec = force
    o.seed = seed
    o.float_bitwise = float_bitwise
    return o


def pack_bits(bools):
    """bool ndarray -> LSB-first bitmap (np.uint8)."""
    return np.packbits(np.asarray(bools, dtype=bool), bitorder="little")


def unpack_bits(bitmap, n, offset=0):
    return np.unpackbits(np.asarray(bitmap, dtype=np.uint8), bitorder="little")[offset:offset + n].astype(bool)


def _make_array(type_, values, validity, keep):
    a = Array()
    if type_ == BOOL:
        bits = pack_bits(values)
        keep.append(bits)
        a.values = bits.ctypes.data
        a.n = len(values)
    elif type_ in (BINARY, LARGE_BINARY):
        offsets, data = values[0], values[1]
        odt = np.int64 if type_ == LARGE_BINARY else np.int32
        offsets = np.ascontiguousarray(offsets, dtype=odt)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if len(data) == 0:
            data = np.zeros(1, dtype=np.uint8)
            backing = 0
        else:
            backing = len(data)
        keep += [offsets, data]
        a.offsets = offsets.ctypes.data
        a.values = data.ctypes.data
        a.values_backing_len = values[2] if len(values) > 2 else backing
        a.n = len(offsets) - 1
    elif type_ == NULL:
        a.n = int(values)
    else:
        v = np.ascontiguousarray(values, dtype=NP_OF[type_])
        keep.append(v)
        a.values = v.ctypes.data if len(v) else None
        a.n = len(v)
    if validity is not None:
        vb = pack_bits(validity)
        keep.append(vb)
        a.validity = vb.ctypes.data
    return a


def write_page(type_, values, validity=None, nullable=None, opts=None):
    """write::write for a flat leaf -> page bytes.  `values`: ndarray (primitives), bool
    ndarray (BOOL), (offsets, data[, backing_len]) (BINARY/LARGE_BINARY)."""
    if nullable is None:
        nullable = validity is not None
    opts = opts or make_opts()
    keep = []
    a = _make_array(type_, values, validity, keep)
    lf = make_leaf(type_, nullable)
    buf = Buf()
    _check(lib().sbo_write_page(C.byref(lf), C.byref(a), C.byref(opts), C.byref(buf)))
    return _take(buf)


def compress_values(type_, values, validity=None, opts=None):
    opts = opts or make_opts()
    keep = []
    a = _make_array(type_, values, validity, keep)
    buf = Buf()
    _check(lib().sbo_compress_values(type_, C.byref(a), C.byref(opts), C.byref(buf)))
    return _take(buf)


def read_column(leaf, pages):
    """batch read (read_integer / read_double / read_binary / read_boolean page loop).
    pages: iterable of (bytes, num_values).  Returns dict(values, offsets, validity, length)."""
    L = lib()
    if not isinstance(leaf, Leaf):
        leaf = make_leaf(*leaf)
    col = L.sbo_col_new(C.byref(leaf))
    try:
        for data, nv in pages:
            b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
            _check(L.sbo_col_read_page(col, b, len(data), nv))
        n = C.c_size_t()
        out = {"length": L.sbo_col_len(col)}
        p = L.sbo_col_values(col, C.byref(n))
        raw = np.frombuffer(C.string_at(p, n.value), dtype=np.uint8).copy() if n.value else np.zeros(0, np.uint8)
        t = leaf.type
        out["values"] = raw.view(NP_OF[t]) if t in NP_OF else raw
        if t in (BINARY, LARGE_BINARY):
            p = L.sbo_col_offsets(col, C.byref(n))
            o = np.frombuffer(C.string_at(p, n.value), dtype=np.uint8).copy()
            out["offsets"] = o.view(np.int64 if t == LARGE_BINARY else np.int32)
        p = L.sbo_col_validity(col, C.byref(n))
        out["validity"] = None
        if p:
            out["validity"] = np.frombuffer(C.string_at(p, (n.value + 7) // 8), dtype=np.uint8).copy()
            out["validity_len"] = n.value
        if leaf.n_nested > 1:
            nd = L.sbo_col_nested_depths(col)
            out["nested"] = []
            for d in range(nd):
                p = L.sbo_col_nested_offsets(col, d, C.byref(n))
                offs = np.frombuffer(C.string_at(p, n.value * 8), dtype=np.int64).copy() if n.value else np.zeros(0, np.int64)
                p = L.sbo_col_nested_validity(col, d, C.byref(n))
                val = np.frombuffer(C.string_at(p, (n.value + 7) // 8), dtype=np.uint8).copy() if n.value else np.zeros(0, np.uint8)
                out["nested"].append({"offsets": offs, "validity": val, "validity_len": n.value})
        return out
    finally:
        L.sbo_col_free(col)


def read_column_body(leaf, body, metas, fetch=True):
    """batch read of a whole column body with the page loop inside the library (no per-page Python work: this
    is the form the CPU baseline times).  body: bytes / uint8 ndarray; metas: [(length, num_values)].
    fetch=False returns only the length (the decoded buffers stay inside the library and are freed)."""
    L = lib()
    if not isinstance(leaf, Leaf):
        leaf = make_leaf(*leaf)
    buf = np.frombuffer(body, dtype=np.uint8) if isinstance(body, (bytes, bytearray, memoryview)) else np.ascontiguousarray(body, dtype=np.uint8)
    m = np.ascontiguousarray(np.array(metas, dtype=np.uint64).reshape(-1, 2))
    col = L.sbo_col_new(C.byref(leaf))
    try:
        _check(L.sbo_col_read_pages(col, buf.ctypes.data if buf.size else None, buf.size, m.ctypes.data, len(m)))
        out = {"length": L.sbo_col_len(col)}
        if not fetch:
            return out
        n = C.c_size_t()
        p = L.sbo_col_values(col, C.byref(n))
        raw = np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(p)).copy() if n.value else np.zeros(0, np.uint8)
        t = leaf.type
        out["values"] = raw.view(NP_OF[t]) if t in NP_OF else raw
        if t in (BINARY, LARGE_BINARY):
            p = L.sbo_col_offsets(col, C.byref(n))
            o = np.ctypeslib.as_array((C.c_uint8 * n.value).from_address(p)).copy()
            out["offsets"] = o.view(np.int64 if t == LARGE_BINARY else np.int32)
        p = L.sbo_col_validity(col, C.byref(n))
        out["validity"] = None
        if p:
            out["validity"] = np.ctypeslib.as_array((C.c_uint8 * ((n.value + 7) // 8)).from_address(p)).copy() if n.value else np.zeros(0, np.uint8)
            out["validity_len"] = n.value
        if leaf.n_nested > 1:
            nd = L.sbo_col_nested_depths(col)
            out["nested"] = []
            for d in range(nd):
                p = L.sbo_col_nested_offsets(col, d, C.byref(n))
                offs = np.frombuffer(C.string_at(p, n.value * 8), dtype=np.int64).copy() if n.value else np.zeros(0, np.int64)
                p = L.sbo_col_nested_validity(col, d, C.byref(n))
                val = np.frombuffer(C.string_at(p, (n.value + 7) // 8), dtype=np.uint8).copy() if n.value else np.zeros(0, np.uint8)
                out["nested"].append({"offsets": offs, "validity": val, "validity_len": n.value})
        return out
    finally:
        L.sbo_col_free(col)


def write_column(type_, values, validity=None, nullable=None, opts=None, page_rows=8192):
    """encode_chunk page loop of one fixed-width flat leaf inside the library: (column body, [(length, num_values)])"""
    if nullable is None:
        nullable = validity is not None
    opts = opts or make_opts()
    keep = []
    a = _make_array(type_, values, validity, keep)
    lf = make_leaf(type_, nullable)
    n = int(a.n)
    cap = (n + max(1, page_rows) - 1) // max(1, page_rows) + 1 if page_rows else 2
    metas = np.zeros((cap, 2), dtype=np.uint64)
    npg = C.c_size_t()
    buf = Buf()
    _check(lib().sbo_write_column(C.byref(lf), C.byref(a), C.byref(opts), page_rows or 0, C.byref(buf), metas.ctypes.data, cap, C.byref(npg)))
    body = _take(buf)
    return body, [(int(metas[i, 0]), int(metas[i, 1])) for i in range(npg.value)]


def stat_block(type_, block):
    out = C.create_string_buffer(512)
    b = (C.c_uint8 * len(block)).from_buffer_copy(block)
    _check(lib().sbo_stat_block(type_, b, len(block), out, 512))
    return out.value.decode()


def value_block_offset(leaf, page):
    if not isinstance(leaf, Leaf):
        leaf = make_leaf(*leaf)
    b = (C.c_uint8 * max(1, len(page))).from_buffer_copy(page if len(page) else b"\0")
    return lib().sbo_page_value_block_offset(C.byref(leaf), b, len(page))


def stat_page(type_, nullable, page):
    off = value_block_offset((type_, nullable), page)
    return stat_block(type_, page[off:])


# ---- third-party layouts -------------------------------------------------------------
def bp4x_compress(block, num_bits=None, initial=None):
    block = np.ascontiguousarray(block, dtype=np.uint32)
    assert len(block) == 128
    L = lib()
    if num_bits is None:
        num_bits = L.sbo_bp4x_num_bits(block.ctypes.data_as(C.c_void_p))
    out = np.zeros(512, dtype=np.uint8)
    if initial is None:
        n = L.sbo_bp4x_compress(block.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), num_bits)
    else:
        n = L.sbo_bp4x_compress_sorted(C.c_uint32(initial), block.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), num_bits)
    return num_bits, out[:n].tobytes()


def bp4x_decompress(data, num_bits, initial=None):
    buf = np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()
    out = np.zeros(128, dtype=np.uint32)
    L = lib()
    if initial is None:
        L.sbo_bp4x_decompress(buf.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), num_bits)
    else:
        L.sbo_bp4x_decompress_sorted(C.c_uint32(initial), buf.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), num_bits)
    return out


def roaring_serialize(vals):
    v = np.ascontiguousarray(vals, dtype=np.uint32)
    buf = Buf()
    _check(lib().sbo_roaring_serialize(v.ctypes.data_as(C.c_void_p), C.c_size_t(len(v)), C.byref(buf)))
    return _take(buf)


def roaring_deserialize(data):
    b = (C.c_uint8 * len(data)).from_buffer_copy(data)
    buf = Buf()
    _check(lib().sbo_roaring_deserialize(b, C.c_size_t(len(data)), C.byref(buf)))
    return np.frombuffer(_take(buf), dtype=np.uint32)


def lz4_decompress(data, out_len, use_lib=False):
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    out = np.zeros(max(1, out_len), dtype=np.uint8)
    f = lib().sbo_lz4_decompress_lib if use_lib else lib().sbo_lz4_decompress
    _check(f(b, C.c_size_t(len(data)), out.ctypes.data_as(C.c_void_p), C.c_size_t(out_len)))
    return out[:out_len].tobytes()


def lz4_compress(data):
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    buf = Buf()
    _check(lib().sbo_lz4_compress_lib(b, C.c_size_t(len(data)), C.byref(buf)))
    return _take(buf)


def common_compress(codec, data):
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    buf = Buf()
    _check(lib().sbo_common_compress(codec, b, C.c_size_t(len(data)), C.byref(buf)))
    return _take(buf)


def common_decompress(codec, data, out_len):
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    out = np.zeros(max(1, out_len), dtype=np.uint8)
    _check(lib().sbo_common_decompress(codec, b, C.c_size_t(len(data)), out.ctypes.data_as(C.c_void_p), C.c_size_t(out_len)))
    return out[:out_len].tobytes()


def patas_pack(r, s, t):
    return lib().sbo_patas_pack(C.c_uint8(r), C.c_uint8(s), C.c_uint8(t))


def patas_unpack(p):
    o = (C.c_uint8 * 3)()
    lib().sbo_patas_unpack(C.c_uint16(p), o)
    return tuple(o)


def hybrid_rle_decode(data, bit_width, n):
    b = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data if len(data) else b"\0")
    out = np.zeros(max(1, n), dtype=np.uint32)
    _check(lib().sbo_hybrid_rle_decode(b, C.c_size_t(len(data)), C.c_uint32(bit_width), C.c_size_t(n), out.ctypes.data_as(C.c_void_p)))
    return out[:n]


def levels_encode(levels, bit_width):
    v = np.ascontiguousarray(levels, dtype=np.uint32)
    buf = Buf()
    _check(lib().sbo_levels_encode(v.ctypes.data_as(C.c_void_p), C.c_size_t(len(v)), C.c_uint32(bit_width), C.byref(buf)))
    return _take(buf)


def sample_draw(seed, codec, i, range_end):
    return lib().sbo_sample_draw(seed, codec, i, range_end)
