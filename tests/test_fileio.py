"""CPU: file framing (NativeWriter::{start,finish}, read_meta, infer_schema; NativeReader::{next,nth,skip_page}):
src/write/writer.rs:91-167, src/read/reader.rs:45-241.  GPU: NativeWriter.write (encode_chunk) -> file ->
NativeReader pages -> decode, against the oracle."""
import io

import numpy as np
import pytest
import sbo
from helpers import oracle_encode_column

from strawboat_b200 import fileio


def make_file(rng, ncols=3, rows=5000):
    cols, srcs = [], []
    for c in range(ncols):
        v = rng.integers(0, 1000, rows).astype(np.int64)
        body, metas = oracle_encode_column(sbo.I64, v, page_size=1024, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
        cols.append((body, metas))
        srcs.append(v)
    return cols, srcs


def test_file_roundtrip_and_schema():
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(0)
    cols, _ = make_file(rng)
    schema = pa.schema([pa.field(f"c{i}", pa.int64(), nullable=False) for i in range(len(cols))])
    data, metas = fileio.write_file(cols, schema)
    assert data[:8] == b"ARROW2\0\0" and data[-8:] == b"\xff\xff\xff\xff\0\0\0\0"
    got = fileio.read_meta(data)
    assert [(o, [tuple(p) for p in pg]) for o, pg in got] == [(o, [tuple(p) for p in pg]) for o, pg in metas]
    assert fileio.schema_from_bytes(fileio.infer_schema_bytes(data)).equals(schema)
    for (body, _), cm in zip(cols, got):
        assert fileio.column_body(data, cm) == body


def test_native_reader_next_nth_skip():
    """NativeReader::{next, nth, skip_page, has_next, current_page} (src/read/reader.rs:75-145) on bytes and on a file object"""
    rng = np.random.default_rng(1)
    cols, srcs = make_file(rng, ncols=2)
    data, _ = fileio.write_file(cols)
    metas = fileio.read_meta(data)
    for source in (data, io.BytesIO(data)):
        r = fileio.NativeReader(source, metas[1])
        pages = list(fileio.NativeReader(source, metas[1]))
        assert len(pages) == len(metas[1][1]) == 5 and b"".join(p for _, p in pages) == cols[1][0]
        assert r.has_next() and r.current_page == 0
        assert r.nth(2) == pages[2] and r.current_page == 3          # skips pages 0, 1
        r.skip_page()
        assert r.current_page == 4 and r.next() == pages[4]
        assert not r.has_next() and r.next() is None and r.nth(0) is None
        r.skip_page()                                                  # no-op at the end
        r2 = fileio.NativeReader(source, metas[0])
        assert r2.nth(5) is None and r2.current_page == 5             # ran off the end: None (reader.rs:104-106)
        # every page decodes on its own with the oracle (one array per page, like column_iter_to_arrays)
        got = np.concatenate([sbo.read_column((sbo.I64, False), [(p, nv)])["values"] for nv, p in fileio.NativeReader(source, metas[0])])
        assert np.array_equal(got, srcs[0])
    short = fileio.NativeReader(data[:metas[1][0] + 10], metas[1])
    with pytest.raises(EOFError):
        short.next()


@pytest.mark.gpu
def test_native_writer_chunks_on_gpu(ctx):
    """NativeWriter: start / write(chunk) x 2 / finish; the file reads back page by page (NativeReader ->
    decode_pages) and in batch (read_columns), and the oracle reads every page too"""
    import strawboat_b200 as sb
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(2)
    schema = pa.schema([pa.field("a", pa.int64(), nullable=False), pa.field("b", pa.float64()), pa.field("s", pa.string())])
    sink = io.BytesIO()
    w = fileio.NativeWriter(ctx, sink, schema, sb.write_options(sb.C_LZ4, 2.0, 1000, seed=5))
    with pytest.raises(RuntimeError):
        w.write([])
    w.start()
    chunks = []
    for rows in (3500, 1200):
        a = rng.integers(0, 1 << 40, rows).astype(np.int64)
        b, bv = rng.integers(0, 8, rows).astype(np.float64), rng.random(rows) > 0.2
        words = [b"x%d" % i for i in rng.integers(0, 30, rows)]
        off = np.concatenate([[0], np.cumsum([len(x) for x in words])]).astype(np.int32)
        dat = np.frombuffer(b"".join(words), np.uint8)
        chunks.append((a, b, bv, off, dat))
        w.write([sb.LeafArray(sb.I64, a), sb.LeafArray(sb.F64, b, validity=bv), sb.LeafArray(sb.BINARY, (off, dat), nullable=True)])
    size = w.finish()
    data = sink.getvalue()
    assert size == len(data) and len(w.metas) == 6
    assert fileio.schema_from_bytes(fileio.infer_schema_bytes(data)).equals(schema)
    metas = fileio.read_meta(data)
    assert [m[0] for m in metas] == [m[0] for m in w.metas]
    leaves = [(sb.I64, False), (sb.F64, True), (sb.BINARY, True)] * 2
    dec = fileio.read_columns(ctx, data, leaves)
    for k, (a, b, bv, off, dat) in enumerate(chunks):
        assert np.array_equal(dec[3 * k].values, a)
        assert np.array_equal(dec[3 * k + 1].values[bv], b[bv]) and np.array_equal(sbo.unpack_bits(dec[3 * k + 1].validity, len(b)), bv)
        assert np.array_equal(dec[3 * k + 2].offsets, off) and np.array_equal(dec[3 * k + 2].values, dat)
    # streaming form: one array per page, with a page skipped in the middle
    r = fileio.NativeReader(io.BytesIO(data), metas[0])
    first = r.next()
    r.skip_page()
    third = r.next()
    pages = [sb.Column(sb.I64, False, p, [(len(p), nv)]) for nv, p in (first, third)]
    out = ctx.decode_pages(pages)
    assert np.array_equal(out[0].values, chunks[0][0][:1000]) and np.array_equal(out[1].values, chunks[0][0][2000:3000])
    ref = np.concatenate([sbo.read_column((sbo.I64, False), [(p, nv)])["values"] for nv, p in fileio.NativeReader(data, metas[0])])
    assert np.array_equal(ref, chunks[0][0])
