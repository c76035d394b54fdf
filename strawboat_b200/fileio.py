"""File framing of the strawboat format (host-side glue).

    "ARROW2" 00 00 | column bodies (leaf order) | schema bytes | meta | u32 schema_size |
    u32 meta_size | FF FF FF FF 00 00 00 00

mirrors NativeWriter::{start, finish} (src/write/writer.rs:91-167), read_meta / infer_schema
(src/read/reader.rs:168-241) and ColumnMeta / PageMeta (src/lib.rs:40-80).  No codec work happens
here: column bodies come from Context.encode_columns and go to Context.decode_columns.
The footer schema is the raw Arrow IPC Schema message flatbuffer (arrow2 `schema_to_bytes`);
it is produced / parsed with pyarrow when a schema object is wanted, and passed through as
opaque bytes otherwise.
"""
import struct

MAGIC = b"ARROW2"            # src/lib.rs:34
CONTINUATION = b"\xff\xff\xff\xff"  # src/lib.rs:35


def schema_to_bytes(schema):
    """pyarrow.Schema -> raw IPC Message flatbuffer (no continuation marker / length prefix)."""
    buf = schema.serialize().to_pybytes()
    assert buf[:4] == CONTINUATION
    (n,) = struct.unpack_from("<i", buf, 4)
    return buf[8:8 + n]


def schema_from_bytes(raw):
    import pyarrow as pa
    pad = (-len(raw)) % 8
    return pa.ipc.read_schema(pa.py_buffer(CONTINUATION + struct.pack("<i", len(raw) + pad) + raw + b"\0" * pad))


def write_file(columns, schema_bytes=b""):
    """columns: list of (body bytes, [(length, num_values)]) in leaf order.  Returns the file bytes
    and the ColumnMeta list [(offset, pages)] (NativeWriter.metas)."""
    out = bytearray(MAGIC + b"\0\0")                       # writer.rs:98-100: data starts at byte 8
    metas = []
    for body, pages in columns:
        assert sum(p[0] for p in pages) == len(body)
        metas.append((len(out), list(pages)))               # ColumnMeta.offset is absolute (common.rs:76)
        out += body
    if isinstance(schema_bytes, (bytes, bytearray)):
        sb_ = bytes(schema_bytes)
    else:
        sb_ = schema_to_bytes(schema_bytes)
    out += sb_
    meta = bytearray(struct.pack("<Q", len(metas)))         # writer.rs:143-150
    for off, pages in metas:
        meta += struct.pack("<QQ", off, len(pages))
        for length, num_values in pages:
            meta += struct.pack("<QQ", length, num_values)
    out += meta
    out += struct.pack("<II", len(sb_), len(meta))         # writer.rs:157-161
    out += CONTINUATION + b"\0\0\0\0"                       # writer.rs:163
    return bytes(out), metas


def read_meta(data):
    """read_meta (src/read/reader.rs:168-225): [(offset, [(length, num_values)])]."""
    if data[:6] != MAGIC:
        raise ValueError("not a strawboat file (magic)")
    end = len(data)
    (meta_size,) = struct.unpack_from("<I", data, end - 12)
    pos = end - 16 - meta_size
    (n_cols,) = struct.unpack_from("<Q", data, pos)
    pos += 8
    metas = []
    for _ in range(n_cols):
        off, n_pages = struct.unpack_from("<QQ", data, pos)
        pos += 16
        pages = []
        for _ in range(n_pages):
            pages.append(struct.unpack_from("<QQ", data, pos))
            pos += 16
        metas.append((off, pages))
    return metas


def infer_schema_bytes(data):
    """infer_schema (src/read/reader.rs:227-241): the raw schema flatbuffer."""
    end = len(data)
    schema_size, meta_size = struct.unpack_from("<II", data, end - 16)
    start = end - 16 - meta_size - schema_size
    return bytes(data[start:start + schema_size])


def column_body(data, column_meta):
    off, pages = column_meta
    return data[off:off + sum(p[0] for p in pages)]
