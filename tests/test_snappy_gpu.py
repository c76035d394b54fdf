"""Snappy as the common codec (CommonCompression::Snappy, src/compression/basic.rs:98-105,138-152; the reference's
test matrix runs it: tests/it/io.rs:420-425).  GPU decode of oracle-written and of Google-snappy-written blocks,
oracle decode of GPU-written pages, corrupt streams."""
import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb
from strawboat_b200.workloads import random_strings

pytestmark = pytest.mark.gpu


def columns(rng, n):
    yield sbo.I32, np.cumsum(rng.integers(0, 4, n)).astype(np.int32), None
    yield sbo.I64, rng.integers(0, 50, n), rng.random(n) > 0.2
    yield sbo.F64, rng.integers(0, 65536, n).astype(np.float64), None
    yield sbo.U8, rng.integers(0, 256, n).astype(np.uint8), None          # incompressible: literal-only elements
    yield sbo.I64, np.full(n, 7), None                                       # one long overlapping copy chain
    yield sbo.BOOL, rng.random(n) < 0.3, rng.random(n) > 0.1
    o, d, v = random_strings(rng, n, 40, 0.2)
    yield sbo.BINARY, (o, d), v
    o, d, v = random_strings(rng, n, 5000, 0.0, large=True)
    yield sbo.LARGE_BINARY, (o, d), None


@pytest.mark.parametrize("ratio", [None, 2.0])
def test_gpu_reads_oracle_snappy_pages(ctx, ratio):
    rng = np.random.default_rng(1)
    for n in (1, 100, 5000, 20000):
        for t, v, val in columns(rng, n):
            data, metas = oracle_encode_column(t, v, val, page_size=8192, opts=sbo.make_opts(sbo.C_SNAPPY, ratio=ratio))
            ref = oracle_decode_column(t, val is not None, data, metas)
            dec = ctx.batch_read_array(sb.Column(t, val is not None, data, metas))
            assert_same(dec, ref, t, val is not None)


def test_gpu_reads_google_snappy_blocks(ctx):
    """value blocks compressed by pyarrow's snappy codec (Google's C++ library: other element choices than ours)"""
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    rng = np.random.default_rng(2)
    for v in (rng.integers(0, 65536, 8192).astype(np.float64), np.cumsum(rng.integers(0, 4, 8192)).astype(np.int64),
              np.repeat(rng.integers(0, 1 << 40, 100), 80).astype(np.int64), rng.integers(-2**62, 2**62, 3000)):
        raw = np.ascontiguousarray(v).tobytes()
        comp = codec.compress(raw).to_pybytes()
        page = bytes([sb.C_SNAPPY]) + len(comp).to_bytes(4, "little") + len(raw).to_bytes(4, "little") + comp
        t = sbo.F64 if v.dtype == np.float64 else sbo.I64
        dec = ctx.batch_read_array(sb.Column(t, False, page, [(len(page), len(v))]))
        assert np.array_equal(dec.values.view(np.uint8), np.ascontiguousarray(v).view(np.uint8))


@pytest.mark.parametrize("ratio", [None, 2.0])
def test_oracle_reads_gpu_snappy_pages(ctx, ratio):
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(3)
    for n in (1, 100, 5000, 20000):
        for t, v, val in columns(rng, n):
            enc = ctx.encode_columns([sb.LeafArray(t, v, validity=val)], sb.write_options(sb.C_SNAPPY, ratio, 8192, seed=1))[0]
            ref = oracle_decode_column(t, val is not None, enc.data, enc.metas)
            dec = ctx.batch_read_array(sb.Column(t, val is not None, enc.data, enc.metas))
            assert_same(dec, ref, t, val is not None)
            if t in sbo.NP_OF and val is None:
                assert np.array_equal(ref["values"].view(np.uint8), np.ascontiguousarray(v, dtype=sbo.NP_OF[t]).view(np.uint8))
    # a GPU-written top-level snappy block is a valid stream for Google's decoder too
    v = rng.integers(0, 65536, 8192).astype(np.float64)
    enc = ctx.encode_columns([sb.LeafArray(sb.F64, v)], sb.write_options(sb.C_SNAPPY, None, 8192))[0]
    assert enc.data[0] == sb.C_SNAPPY
    clen = int.from_bytes(enc.data[1:5], "little")
    assert pa.Codec("snappy").decompress(enc.data[9:9 + clen], decompressed_size=v.nbytes).to_pybytes() == v.tobytes()


def test_corrupt_snappy_streams(ctx):
    rng = np.random.default_rng(4)
    v = rng.integers(0, 1000, 4096).astype(np.int64)
    data, metas = oracle_encode_column(sbo.I64, v, page_size=2048, opts=sbo.make_opts(sbo.C_SNAPPY))
    l0 = metas[0][0]
    for mutate in ("truncate", "length", "offset"):
        bad = bytearray(data)
        if mutate == "truncate":      # payload shorter than the elements need
            bad[1:5] = (int.from_bytes(bad[1:5], "little") - 7).to_bytes(4, "little")
        elif mutate == "length":      # preamble disagrees with the rows
            bad[9] ^= 0x01
        else:                          # the first copy element points before the start of the output
            i = 9
            while bad[i] & 0x80:       # preamble
                i += 1
            i += 1
            while True:                # walk the elements
                tag = bad[i]
                if tag & 3 == 0:
                    ln = (tag >> 2) + 1
                    i += 1
                    if ln > 60:
                        nb = ln - 60
                        ln = int.from_bytes(bad[i:i + nb], "little") + 1
                        i += nb
                    i += ln
                elif tag & 3 == 1:
                    bad[i] |= 0xe0
                    bad[i + 1] = 0xff
                    break
                else:
                    bad[i + 1], bad[i + 2] = 0xff, 0xff
                    break
        res = ctx.decode_columns([sb.Column(sb.I64, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] in (sb._capi.SB_EXTERNAL, sb._capi.SB_IO) and res.page_status[1] == 0
        assert np.array_equal(res.values[2048:], v[2048:])
