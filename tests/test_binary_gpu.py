"""GPU parity for Binary / Utf8 / LargeBinary value blocks (src/compression/binary/*,
src/read/array/binary.rs): strawboat_b200 decode == oracle decode, bit for bit."""
import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb
from strawboat_b200.workloads import random_strings

pytestmark = pytest.mark.gpu


def roundtrip(ctx, type_, values, validity=None, nullable=None, page_size=2048, opts=None, expect_codec=None):
    if nullable is None:
        nullable = validity is not None
    data, metas = oracle_encode_column(type_, values, validity, nullable, page_size, opts)
    if expect_codec is not None:
        tree = sbo.stat_page(type_, nullable, data[:metas[0][0]])
        assert tree.startswith(expect_codec), tree
    ref = oracle_decode_column(type_, nullable, data, metas)
    dec = ctx.batch_read_array(sb.Column(type_, nullable, data, metas))
    assert_same(dec, ref, type_, nullable)
    return dec


def strings(rng, n, uniq, nulls=0.0, large=False, maxlen=None):
    if maxlen is None:
        o, d, v = random_strings(rng, n, uniq, nulls, large=large)
        return (o, d), v
    lens = rng.integers(0, maxlen + 1, n)
    offsets = np.zeros(n + 1, dtype=np.int64 if large else np.int32)
    np.cumsum(lens, out=offsets[1:])
    data = rng.integers(0, 256, int(offsets[-1])).astype(np.uint8)
    v = (rng.random(n) >= nulls) if nulls else None
    return (offsets, data), v


@pytest.mark.parametrize("type_", [sbo.BINARY, sbo.LARGE_BINARY])
@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
def test_basic(ctx, type_, default):
    """new_test_chunk utf8 columns + random binary (io.rs:48-102)."""
    large = type_ == sbo.LARGE_BINARY
    odt = np.int64 if large else np.int32
    words = [b"a", b"bbbbb", b"", b"cc", b"dddddddd", b"e"]
    off = np.cumsum([0] + [len(w) for w in words]).astype(odt)
    dat = np.frombuffer(b"".join(words), np.uint8)
    roundtrip(ctx, type_, (off, dat), opts=sbo.make_opts(default))
    rng = np.random.default_rng(1)
    for n in (1, 100, 5000):
        vals, _ = strings(rng, n, None, large=large, maxlen=40)
        roundtrip(ctx, type_, vals, opts=sbo.make_opts(default))
        vals, v = strings(rng, n, None, nulls=0.3, large=large, maxlen=40)
        roundtrip(ctx, type_, vals, validity=v, opts=sbo.make_opts(default))
        roundtrip(ctx, type_, vals, validity=v, page_size=777, opts=sbo.make_opts(default))


@pytest.mark.parametrize("type_", [sbo.BINARY, sbo.LARGE_BINARY])
@pytest.mark.parametrize("force", [sbo.C_DICT, sbo.C_FREQ, sbo.C_ONEVALUE])
def test_forced(ctx, type_, force):
    large = type_ == sbo.LARGE_BINARY
    rng = np.random.default_rng(2)
    for uniq in (1, 8, 1000):
        for nulls in (0.0, 0.4):
            if force == sbo.C_ONEVALUE and uniq != 1:
                continue
            vals, v = strings(rng, 6000, uniq, nulls=nulls, large=large)
            for default in (sbo.C_NONE, sbo.C_LZ4):
                roundtrip(ctx, type_, vals, validity=v, opts=sbo.make_opts(default, ratio=2.0, force=force))


@pytest.mark.parametrize("type_", [sbo.BINARY, sbo.LARGE_BINARY])
def test_adaptive(ctx, type_):
    """config 3 shapes: low-cardinality decimal strings, 40 % nulls, adaptive on."""
    large = type_ == sbo.LARGE_BINARY
    rng = np.random.default_rng(3)
    opts = sbo.make_opts(sbo.C_LZ4, ratio=2.0)
    vals, v = strings(rng, 8192 * 3 + 55, 1000, nulls=0.4, large=large)
    roundtrip(ctx, type_, vals, validity=v, page_size=8192, opts=opts, expect_codec="Dict")
    vals, v = strings(rng, 8192 * 2, 1, large=large)
    roundtrip(ctx, type_, vals, page_size=8192, opts=opts, expect_codec="OneValue")
    # sorted within the page: the Dict index sub-page picks RLE ("dict+RLE")
    o, d, v = random_strings(rng, 8192 * 2, 100, 0.0, large=large, sort_within=8192)
    roundtrip(ctx, type_, (o, d), page_size=8192, opts=opts, expect_codec="Dict(Rle")
    # 95 % one value: Freq
    n = 8192
    ids = np.where(rng.random(n) < 0.95, 0, rng.integers(1, 500, n))
    table = [b"the-frequent-value"] + [b"x%d" % i for i in range(1, 500)]
    lens = np.array([len(table[i]) for i in ids])
    off = np.zeros(n + 1, np.int64 if large else np.int32)
    np.cumsum(lens, out=off[1:])
    dat = np.frombuffer(b"".join(table[i] for i in ids), np.uint8)
    roundtrip(ctx, type_, (off, dat), page_size=8192, opts=opts, expect_codec="Freq")
    # high-cardinality random bytes: falls back to the default codec
    vals, v = strings(rng, 5000, None, nulls=0.1, large=large, maxlen=24)
    roundtrip(ctx, type_, vals, validity=v, page_size=2048, opts=opts)


def test_empty_and_long_values(ctx):
    rng = np.random.default_rng(4)
    n = 300
    off = np.zeros(n + 1, np.int32)
    roundtrip(ctx, sbo.BINARY, (off, np.zeros(0, np.uint8)))  # all empty strings
    lens = np.where(rng.random(n) < 0.05, rng.integers(1000, 20000, n), rng.integers(0, 5, n))
    off = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    dat = rng.integers(0, 256, int(off[-1])).astype(np.uint8)
    for opts in (sbo.make_opts(), sbo.make_opts(sbo.C_LZ4), sbo.make_opts(force=sbo.C_DICT), sbo.make_opts(force=sbo.C_FREQ)):
        roundtrip(ctx, sbo.BINARY, (off, dat), page_size=100, opts=opts)
        roundtrip(ctx, sbo.BINARY, (off, dat), page_size=None, opts=opts)


def test_binary_with_other_columns(ctx):
    rng = np.random.default_rng(5)
    cols, refs = [], []
    for i, t in enumerate([sbo.BINARY, sbo.I64, sbo.LARGE_BINARY, sbo.BOOL, sbo.BINARY]):
        n = 4000 + 13 * i
        if t in (sbo.BINARY, sbo.LARGE_BINARY):
            vals, v = strings(rng, n, 50, nulls=0.2 if i else 0.0, large=t == sbo.LARGE_BINARY)
        elif t == sbo.BOOL:
            vals, v = rng.random(n) < 0.5, None
        else:
            vals, v = rng.integers(0, 100, n).astype(np.int64), rng.random(n) > 0.1
        data, metas = oracle_encode_column(t, vals, v, page_size=1000, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
        cols.append(sb.Column(t, v is not None, data, metas))
        refs.append((oracle_decode_column(t, v is not None, data, metas), t, v is not None))
    for d, (r, t, nu) in zip(ctx.decode_columns(cols), refs):
        assert_same(d, r, t, nu)


def test_corrupt_binary(ctx):
    rng = np.random.default_rng(6)
    vals, v = strings(rng, 4096, 20)
    for opts in (sbo.make_opts(), sbo.make_opts(force=sbo.C_DICT), sbo.make_opts(force=sbo.C_FREQ)):
        data, metas = oracle_encode_column(sbo.BINARY, vals, page_size=2048, opts=opts)
        bad = bytearray(data)
        bad[0] = 77
        res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] == sb._capi.SB_OUT_OF_SPEC and res.page_status[1] == 0
        bad = bytearray(data)
        bad[1:5] = (0x7fffffff).to_bytes(4, "little")
        res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] != 0 and res.page_status[1] == 0


def _u32(b, p):
    return int.from_bytes(b[p:p + 4], "little")


def _second_page_intact(ctx, res, data, metas):
    """page 1 of a two-page column decodes to what the oracle reads from it, whatever happened to page 0"""
    assert res.page_status[1] == 0
    l0, n0 = metas[0]
    ref = oracle_decode_column(sbo.BINARY, False, data[l0:], metas[1:])
    # page 0's own offsets are undefined when it was rejected (and it may or may not have been given value bytes,
    # depending on which pass rejected it): page 1 is checked relative to its own first offset
    base = int(res.offsets[n0 + 1]) - int(ref["offsets"][1])
    assert base >= 0 and np.array_equal(res.offsets[n0 + 1:] - base, ref["offsets"][1:])
    assert np.array_equal(res.values[base:base + len(ref["values"])], ref["values"])


def test_corrupt_dict_entries(ctx):
    """a binary Dict page whose `[u64 len][bytes]` chain is broken fails in the plan pass;
    the decode pass must not touch it (no reads through a stale entry table, no writes past the sized output)."""
    rng = np.random.default_rng(16)
    vals, _ = strings(rng, 4096, 20)
    data, metas = oracle_encode_column(sbo.BINARY, vals, page_size=2048, opts=sbo.make_opts(force=sbo.C_DICT))
    # a healthy call first: the context's entry-table pool now holds stale records for the next call
    ctx.decode_columns([sb.Column(sb.BINARY, False, data, metas)])
    sub_len = 9 + _u32(data, 9 + 1)  # index sub-page inside the Dict body
    ent0 = 9 + sub_len + 4           # first `[u64 len]`
    for new_len in (1 << 40, 0xfffffff0, 3, 0):
        bad = bytearray(data)
        bad[ent0:ent0 + 8] = int(new_len).to_bytes(8, "little")
        res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        if new_len > 0xffff:
            assert res.page_status[0] in (sb._capi.SB_OUT_OF_SPEC, sb._capi.SB_IO)
        _second_page_intact(ctx, res, data, metas)
    # the dictionary size itself: larger / smaller than what the index page refers to
    for k in (1, 5, 1 << 20):
        bad = bytearray(data)
        bad[ent0 - 4:ent0] = int(k).to_bytes(4, "little")
        res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] != 0
        _second_page_intact(ctx, res, data, metas)


def test_corrupt_freq_roaring(ctx):
    """a binary Freq page whose Roaring array container holds duplicate / unsorted rows: exception ranks no
    longer match the entries the plan pass walked -- flagged, never read or written out of bounds"""
    rng = np.random.default_rng(17)
    n = 4096
    ids = np.where(rng.random(n) < 0.95, 0, rng.integers(1, 300, n))
    table = [b"top-value"] + [b"exception-%d" % i for i in range(1, 300)]
    lens = np.array([len(table[i]) for i in ids])
    off = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    dat = np.frombuffer(b"".join(table[i] for i in ids), np.uint8)
    data, metas = oracle_encode_column(sbo.BINARY, (off, dat), page_size=2048, opts=sbo.make_opts(force=sbo.C_FREQ))
    top_len = int.from_bytes(data[9:17], "little")
    rb = 9 + 8 + top_len + 4                      # roaring bytes
    assert _u32(data, rb) == 12346 and _u32(data, rb + 4) == 1
    card = int.from_bytes(data[rb + 10:rb + 12], "little") + 1
    v0 = rb + 8 + 4 + 4                           # first u16 of the array container
    assert card >= 8
    cases = []
    bad = bytearray(data)                         # duplicates
    bad[v0 + 2:v0 + 4] = bad[v0:v0 + 2]
    bad[v0 + 6:v0 + 8] = bad[v0 + 4:v0 + 6]
    cases.append(bad)
    bad = bytearray(data)                         # unsorted
    bad[v0:v0 + 2], bad[v0 + 2 * (card - 1):v0 + 2 * card] = bad[v0 + 2 * (card - 1):v0 + 2 * card], bad[v0:v0 + 2]
    cases.append(bad)
    bad = bytearray(data)                         # rows beyond the page
    bad[v0 + 2 * (card - 1):v0 + 2 * card] = (60000).to_bytes(2, "little")
    cases.append(bad)
    bad = bytearray(data)                         # cardinality larger than the exception entries that follow
    bad[rb + 10:rb + 12] = (card + 40 - 1).to_bytes(2, "little")
    cases.append(bad)
    for bad in cases:
        res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        _second_page_intact(ctx, res, data, metas)


def test_validity_section_without_bitmap(ctx):
    """`L == 0` pushes no validity (read_basic.rs:43-45) and the array constructor then rejects the
    page; it must not come back as all-null data with status OK"""
    v = np.arange(100, dtype=np.int64)
    page = sbo.write_page(sbo.I64, v, None, nullable=True, opts=sbo.make_opts())
    L = _u32(page, 0)
    bad = (0).to_bytes(4, "little") + page[4 + L:]
    res = ctx.decode_columns([sb.Column(sb.I64, True, bad, [(len(bad), 100)])], raise_on_page_error=False)[0]
    assert res.page_status[0] == sb._capi.SB_OUT_OF_SPEC


def _table_column(rng, n, table, nulls=0.0, large=False):
    """n rows drawn from `table` (list of bytes); returns ((offsets, data), validity)"""
    ids = rng.integers(0, len(table), n)
    v = (rng.random(n) >= nulls) if nulls else None
    lens = np.array([len(t) for t in table], dtype=np.int64)[ids]
    if v is not None:
        lens = np.where(v, lens, 0)
    off = np.zeros(n + 1, dtype=np.int64 if large else np.int32)
    np.cumsum(lens, out=off[1:])
    parts = [table[i] if (v is None or v[r]) else b"" for r, i in enumerate(ids)]
    return (off, np.frombuffer(b"".join(parts), np.uint8)), v


@pytest.mark.parametrize("type_", [sbo.BINARY, sbo.LARGE_BINARY])
def test_dict_walk_kernel_windows(ctx, type_):
    """sb_dict_walk_kernel (one warp per Dict page, dictionary staged in 20 KB windows): dictionaries longer than one
    window, headers that straddle a window end, one entry longer than a window, empty entries; binary/dict.rs:95-141."""
    large = type_ == sbo.LARGE_BINARY
    rng = np.random.default_rng(23)
    force = sbo.make_opts(force=sbo.C_DICT)
    # 3000 entries of 0..60 random bytes: ~115 KB of dictionary = 6 windows, arbitrary header phases
    table = [rng.integers(0, 256, int(l)).astype(np.uint8).tobytes() for l in rng.integers(0, 61, 3000)]
    table = list(dict.fromkeys(table))
    vals, _ = _table_column(rng, 6000, table, large=large)
    roundtrip(ctx, type_, vals, page_size=6000, opts=force, expect_codec="Dict")
    vals, v = _table_column(rng, 6000, table, nulls=0.3, large=large)
    roundtrip(ctx, type_, vals, validity=v, page_size=4096, opts=force, expect_codec="Dict")
    # one 50 KB entry in the middle of 100 short ones (its payload spans three windows), and a 70000-byte one (> u16)
    table = [b"k%d" % i for i in range(50)] + [bytes(rng.integers(0, 256, 50_000).astype(np.uint8))] + [b"v%03d" % i for i in range(50)] \
        + [bytes(rng.integers(0, 256, 70_000).astype(np.uint8))]
    vals, _ = _table_column(rng, 400, table, large=large)
    roundtrip(ctx, type_, vals, page_size=400, opts=force, expect_codec="Dict")
    # entry lengths chosen so that headers land on every phase of the 20 KB window end
    for pad in range(0, 24, 3):
        table = [b"x" * (997 + pad)] + [b"%05d" % i + b"y" * (i % 7) for i in range(2500)]
        vals, _ = _table_column(rng, 5000, table, large=large)
        roundtrip(ctx, type_, vals, page_size=5000, opts=force, expect_codec="Dict")


def test_dict_walk_kernel_corrupt_chain(ctx):
    """a broken `[u64 len]` chain in a dictionary large enough for sb_dict_walk_kernel (k >= 32): the pre-walk leaves no
    record, the plan pass walks itself and reports the page; the next page is intact"""
    rng = np.random.default_rng(24)
    vals, _ = strings(rng, 4096, 300)
    data, metas = oracle_encode_column(sbo.BINARY, vals, page_size=2048, opts=sbo.make_opts(force=sbo.C_DICT))
    ctx.decode_columns([sb.Column(sb.BINARY, False, data, metas)])  # healthy call first: stale records in the pools
    sub_len = 9 + _u32(data, 9 + 1)
    k = _u32(data, 9 + sub_len)
    assert k >= 32
    pos = 9 + sub_len + 4
    starts = []
    for _ in range(k):
        starts.append(pos)
        pos += 8 + int.from_bytes(data[pos:pos + 8], "little")
    for e in (0, 40, k - 1):
        for new_len in (1 << 40, 0xfffffff0, metas[0][0]):
            bad = bytearray(data)
            bad[starts[e]:starts[e] + 8] = int(new_len).to_bytes(8, "little")
            res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
            assert res.page_status[0] in (sb._capi.SB_OUT_OF_SPEC, sb._capi.SB_IO)
            _second_page_intact(ctx, res, data, metas)
    # a shifted but in-bounds length: the chain still parses somewhere; no crash, page 1 untouched
    bad = bytearray(data)
    bad[starts[40]:starts[40] + 8] = (1).to_bytes(8, "little") if data[starts[40]] != 1 else (2).to_bytes(8, "little")
    res = ctx.decode_columns([sb.Column(sb.BINARY, False, bytes(bad), metas)], raise_on_page_error=False)[0]
    _second_page_intact(ctx, res, data, metas)
    # the healthy column again: same result as the oracle
    ref = oracle_decode_column(sbo.BINARY, False, data, metas)
    dec = ctx.batch_read_array(sb.Column(sb.BINARY, False, data, metas))
    assert_same(dec, ref, sbo.BINARY, False)
