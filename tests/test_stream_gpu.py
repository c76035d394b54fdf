"""GPU parity for the pages that go through the TMA streaming ring instead of being staged whole
(DESIGN.md §3): plain (codec None) pages of 32 KiB and more, "stored" LZ4 blocks, tiles of oversized
pages, and the offsets / value-byte slices of large binary Basic pages.  Decode through the C ABI ==
oracle decode, bit for bit, at page sizes and alignments that put every chunk boundary somewhere else."""
import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb

pytestmark = pytest.mark.gpu


def rand_values(rng, type_, n):
    dt = sbo.NP_OF[type_]
    if type_ in (sbo.F32, sbo.F64):
        return rng.standard_normal(n).astype(dt)
    info = np.iinfo(dt)
    return rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)


def column(type_, values, validity, page_size, opts=None):
    nullable = validity is not None
    data, metas = oracle_encode_column(type_, values, validity, nullable, page_size, opts)
    return sb.Column(type_, nullable, data, metas), oracle_decode_column(type_, nullable, data, metas), data, metas


@pytest.mark.parametrize("type_", [sbo.I8, sbo.I16, sbo.I32, sbo.I64, sbo.F64])
@pytest.mark.parametrize("nullable", [False, True])
def test_plain_pages(ctx, type_, nullable):
    rng = np.random.default_rng(11)
    W = np.dtype(sbo.NP_OF[type_]).itemsize
    for n, page in ((70_001, 8192), (70_001, 8191), (150_000, 20_000), (40_000 // W * 8 + 5, 32768 // W), (300_001, 100_003)):
        v = rand_values(rng, type_, n)
        val = (rng.random(n) >= 0.2) if nullable else None
        col, ref, data, metas = column(type_, v, val, page)
        assert sbo.stat_page(type_, nullable, data[:metas[0][0]]) == "None"
        assert_same(ctx.batch_read_array(col), ref, type_, nullable)


def test_stored_lz4_blocks(ctx):
    """incompressible data under default LZ4: liblz4 emits one literal run per block (copied through the ring)"""
    rng = np.random.default_rng(12)
    for t, n, page in ((sbo.I32, 100_000, 8192), (sbo.I64, 100_000, 8192), (sbo.I64, 200_000, None), (sbo.F64, 33_333, 11_111)):
        v = rand_values(rng, t, n)
        for val in (None, rng.random(n) >= 0.5):
            col, ref, data, metas = column(t, v, val, page, sbo.make_opts(sbo.C_LZ4))
            assert sbo.stat_page(t, val is not None, data[:metas[0][0]]) == "Lz4"
            assert_same(ctx.batch_read_array(col), ref, t, val is not None)


def test_many_plain_columns_one_call_and_per_page(ctx):
    rng = np.random.default_rng(13)
    cols, refs, types = [], [], []
    for k, t in enumerate((sbo.I64, sbo.I32, sbo.F64, sbo.I16, sbo.I64, sbo.I8)):
        n = 50_000 + 1237 * k
        val = (rng.random(n) >= 0.1) if k % 2 else None
        col, ref, _, _ = column(t, rand_values(rng, t, n), val, 8192 - k)
        cols.append(col), refs.append(ref), types.append((t, val is not None))
    for dec, ref, (t, nu) in zip(ctx.decode_columns(cols), refs, types):
        assert_same(dec, ref, t, nu)
    # streaming entry: one array per page (sb_decode_pages)
    t, n = sbo.I64, 30_000
    v = rand_values(rng, t, n)
    data, metas = oracle_encode_column(t, v, None, False, 8192)
    pages, pos = [], 0
    for ln, nv in metas:
        pages.append(sb.Column(t, False, data[pos:pos + ln], [(ln, nv)]))
        pos += ln
    outs = ctx.decode_pages(pages)
    assert np.array_equal(np.concatenate([o.values for o in outs]), v)


def strings(rng, n, maxlen, large, nulls=None, minlen=0):
    lens = rng.integers(minlen, maxlen + 1, n)
    off = np.zeros(n + 1, dtype=np.int64 if large else np.int32)
    np.cumsum(lens, out=off[1:])
    data = rng.integers(0, 256, int(off[-1])).astype(np.uint8)
    return (off, data), ((rng.random(n) >= nulls) if nulls else None)


@pytest.mark.parametrize("type_", [sbo.BINARY, sbo.LARGE_BINARY])
@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
def test_large_binary_basic_pages(ctx, type_, default):
    """pages above the staging buffer: tile 0 = validity + offsets, value bytes in 128 KiB slices"""
    rng = np.random.default_rng(14)
    large = type_ == sbo.LARGE_BINARY
    for n, maxlen, page, nulls in ((60_000, 40, 8192, None), (60_000, 40, 8191, 0.3), (200_000, 9, None, None), (50_000, 300, 5000, 0.1),
                                   (40_000, 0, None, None), (70_000, 3, 30_000, 0.5)):
        vals, v = strings(rng, n, maxlen, large, nulls)
        col, ref, data, metas = column(type_, vals, v, page, sbo.make_opts(default))
        assert default != sbo.C_NONE or max(m[0] for m in metas) > 70 * 1024
        assert_same(ctx.batch_read_array(col), ref, type_, v is not None)


def test_binary_next_to_fixed_columns(ctx):
    rng = np.random.default_rng(15)
    vals, v = strings(rng, 80_000, 30, False, 0.2, minlen=4)
    c1, r1, _, _ = column(sbo.BINARY, vals, v, 8192)
    c2, r2, _, _ = column(sbo.I64, rand_values(rng, sbo.I64, 80_000), None, 8192)
    vals3, _ = strings(rng, 80_000, 12, True)
    c3, r3, _, _ = column(sbo.LARGE_BINARY, vals3, None, 8192)
    d1, d2, d3 = ctx.decode_columns([c1, c2, c3])
    assert_same(d1, r1, sbo.BINARY, True)
    assert_same(d2, r2, sbo.I64, False)
    assert_same(d3, r3, sbo.LARGE_BINARY, False)


def test_truncated_plain_page_reports_status(ctx):
    """a plain page whose header promises more bytes than the page holds is not classified as plain:
    it takes the checked path and fails with a per-page status"""
    rng = np.random.default_rng(16)
    v = rand_values(rng, sbo.I64, 8192)
    data, metas = oracle_encode_column(sbo.I64, v, None, False, 8192)
    bad = data[:-100]
    with pytest.raises(sb.StrawboatError):
        ctx.batch_read_array(sb.Column(sb.I64, False, bad, [(len(bad), 8192)]))
