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


class NativeReader:
    """Page iterator over one column body: read::reader::NativeReader (src/read/reader.rs:45-145).
    `source` is the whole file (bytes-like) or any seekable binary reader positioned anywhere; `column_meta` =
    (offset, [(length, num_values)]) from read_meta.  next() / nth(n) yield (num_values, page bytes) -- the items
    column_iter_to_arrays consumes (src/read/deserialize.rs:237-245); skip_page() and nth() seek over the skipped
    pages without reading them (reader.rs:91-116,136-145)."""

    def __init__(self, source, column_meta):
        self.offset, self.page_metas = column_meta[0], list(column_meta[1])
        self.current_page = 0
        self._pos = self.offset
        self._src = source
        self._file = hasattr(source, "read") and hasattr(source, "seek")
        self._scratch = None

    def has_next(self):
        return self.current_page < len(self.page_metas)

    def swap_buffer(self, scratch):
        """PageIterator::swap_buffer (src/read/mod.rs:55-57): hand a page buffer back for reuse"""
        self._scratch, scratch = scratch, self._scratch
        return scratch

    def _read(self, n):
        if self._file:
            self._src.seek(self._pos)
            buf = self._src.read(n)
        else:
            buf = bytes(self._src[self._pos:self._pos + n])
        if len(buf) != n:
            raise EOFError("failed to fill whole buffer")
        self._pos += n
        return buf

    def next(self):
        if not self.has_next():
            return None
        length, num_values = self.page_metas[self.current_page]
        buf = self._read(length)
        self.current_page += 1
        return num_values, buf

    def nth(self, n):
        """the next n-th page, skipping the pages in between (Iterator::nth)"""
        skipped, length = 0, 0
        while skipped < n:
            if self.current_page == len(self.page_metas):
                return None
            length += self.page_metas[self.current_page][0]
            self.current_page += 1
            skipped += 1
        self._pos += length
        return self.next()

    def skip_page(self):
        if self.has_next():
            self._pos += self.page_metas[self.current_page][0]
            self.current_page += 1

    def __iter__(self):
        return self

    def __next__(self):
        item = self.next()
        if item is None:
            raise StopIteration
        return item


class NativeWriter:
    """write::NativeWriter (src/write/writer.rs:42-173) over the GPU page encoder: start() writes the header,
    write(chunk) encodes every leaf of a chunk page by page (encode_chunk, src/write/common.rs:49-119 ->
    Context.encode_columns) and appends the column bodies, finish() writes schema + ColumnMeta footer.
    `chunk` = list of strawboat_b200.LeafArray in leaf order (what to_leaves / to_nested yield for the chunk)."""

    def __init__(self, ctx, sink, schema_bytes=b"", options=None):
        self.ctx, self.sink, self.options = ctx, sink, options
        self.schema_bytes = schema_bytes if isinstance(schema_bytes, (bytes, bytearray)) else schema_to_bytes(schema_bytes)
        self.metas = []            # pub metas: Vec<ColumnMeta> (writer.rs:56)
        self.offset = 0
        self.started = self.finished = False

    def start(self):
        self.sink.write(MAGIC + b"\0\0")
        self.offset = 8
        self.started = True

    def write(self, chunk):
        if not self.started:
            raise RuntimeError("start must be called before written")  # writer.rs:110-114
        for enc in self.ctx.encode_columns(list(chunk), self.options):
            self.write_encoded(enc.data, enc.metas)

    def write_encoded(self, body, pages):
        """append one already encoded column body (the multi-GPU writer rank feeds gathered bodies through here)"""
        self.metas.append((self.offset, list(pages)))
        self.sink.write(body)
        self.offset += len(body)

    def finish(self):
        if not self.started:
            raise RuntimeError("start must be called before finish")
        meta = bytearray(struct.pack("<Q", len(self.metas)))
        for off, pages in self.metas:
            meta += struct.pack("<QQ", off, len(pages))
            for length, num_values in pages:
                meta += struct.pack("<QQ", length, num_values)
        self.sink.write(self.schema_bytes)
        self.sink.write(bytes(meta))
        self.sink.write(struct.pack("<II", len(self.schema_bytes), len(meta)))
        self.sink.write(CONTINUATION + b"\0\0\0\0")
        self.finished = True
        return self.offset + len(self.schema_bytes) + len(meta) + 16


def read_columns(ctx, data, leaves):
    """batch_read_array over every column of a file: `leaves` = [(type, nullable[, nested])] per ColumnMeta.
    Returns the decoded columns (host)."""
    import strawboat_b200 as sb
    cols = []
    for cm, lf in zip(read_meta(data), leaves):
        cols.append(sb.Column(lf[0], lf[1], column_body(data, cm), cm[1], lf[2] if len(lf) > 2 else None))
    return ctx.decode_columns(cols)
