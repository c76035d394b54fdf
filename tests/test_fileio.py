"""CPU: file framing (src/write/writer.rs:91-167, src/read/reader.rs:168-241) -- the analogue of
tests/it/read_meta.rs:60-97 (`read_meta == writer.metas`) plus a schema round trip via pyarrow."""
import numpy as np
import pytest
import sbo
from helpers import oracle_decode_column, oracle_encode_column

from strawboat_b200 import fileio


def test_write_read_meta_and_bodies():
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(0)
    schema = pa.schema([pa.field("a", pa.int64(), nullable=False), pa.field("b", pa.float64(), nullable=True),
                        pa.field("c", pa.int32(), nullable=False)])
    cols_in = [(sbo.I64, rng.integers(0, 100, 5000), None), (sbo.F64, rng.standard_normal(5000), rng.random(5000) > 0.2),
               (sbo.I32, np.arange(5000, dtype=np.int32), None)]
    columns = [oracle_encode_column(t, v, val, page_size=2048, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0)) for t, v, val in cols_in]
    data, metas = fileio.write_file(columns, schema)
    assert data[:8] == b"ARROW2\0\0" and data[-8:] == b"\xff\xff\xff\xff\0\0\0\0"
    assert metas[0][0] == 8  # first column starts right after the 8-byte header
    got = fileio.read_meta(data)
    assert [(o, [tuple(p) for p in pg]) for o, pg in got] == [(o, [tuple(p) for p in pg]) for o, pg in metas]
    assert fileio.schema_from_bytes(fileio.infer_schema_bytes(data)).equals(schema)
    for (t, v, val), cm in zip(cols_in, got):
        ref = oracle_decode_column(t, val is not None, fileio.column_body(data, cm), cm[1])
        m = np.ones(len(v), bool) if val is None else val
        assert np.array_equal(ref["values"][m], np.asarray(v, sbo.NP_OF[t])[m])


def test_config1_sizes():
    """BASELINE config 1: 1 M x i64, codec None, 8192 rows/page -> 123 pages, 8 001 107 bytes."""
    from strawboat_b200.workloads import config1
    _, t, v, _ = config1()[0]
    body, pages = oracle_encode_column(t, v, None, page_size=8192, opts=sbo.make_opts())
    assert len(pages) == 123 and len(body) == 8_001_107
    ref = oracle_decode_column(t, False, body, pages)
    assert np.array_equal(ref["values"], v)
