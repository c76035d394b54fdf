"""GPU parity: strawboat_b200 decode (through the C ABI) == oracle decode, bit for bit,
on pages written by the oracle writer.  Mirrors tests/it/io.rs of the reference: every case
runs under the default codecs and with the forced-codec knobs of the CI matrix
(.github/workflows/rust.yml:23-25)."""
import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb

pytestmark = pytest.mark.gpu

INT_TYPES = [sbo.I8, sbo.I16, sbo.I32, sbo.I64, sbo.U8, sbo.U16, sbo.U32, sbo.U64]
FLT_TYPES = [sbo.F32, sbo.F64]
WRITE_PAGE = 2048  # tests/it/io.rs:46


def rand_values(rng, type_, n, card=None):
    dt = sbo.NP_OF[type_]
    if type_ in FLT_TYPES:
        if card:
            return rng.integers(0, card, n).astype(dt)
        return rng.standard_normal(n).astype(dt)
    info = np.iinfo(dt)
    if card:
        return rng.integers(0, min(card, int(info.max)), n).astype(dt)
    return rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)


def roundtrip(ctx, type_, values, validity=None, nullable=None, page_size=WRITE_PAGE, opts=None, expect_codec=None):
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


def test_basic_chunk(ctx):
    """new_test_chunk (io.rs:48-75): 6-row columns of every primitive type."""
    for t in INT_TYPES:
        roundtrip(ctx, t, np.arange(1, 7).astype(sbo.NP_OF[t]))
    for t in FLT_TYPES:
        roundtrip(ctx, t, np.array([1.1, 2.2, 3.3, 4.4, 5.5, 6.6], dtype=sbo.NP_OF[t]))
    roundtrip(ctx, sbo.BOOL, np.array([1, 1, 1, 0, 0, 0], bool))


@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
@pytest.mark.parametrize("type_", INT_TYPES + FLT_TYPES)
def test_random_plain(ctx, type_, default):
    rng = np.random.default_rng(1)
    v = rand_values(rng, type_, 10000)
    roundtrip(ctx, type_, v, opts=sbo.make_opts(default))
    val = rng.random(10000) >= 0.3
    roundtrip(ctx, type_, v, validity=val, opts=sbo.make_opts(default))


@pytest.mark.parametrize("force", [sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ])
@pytest.mark.parametrize("type_", INT_TYPES + FLT_TYPES)
def test_forced_codecs(ctx, type_, force):
    """the forced-codec CI matrix, for low-cardinality and random data, with/without nulls."""
    rng = np.random.default_rng(2)
    for card in (8, None):
        v = rand_values(rng, type_, 10000, card)
        for nulls in (0.0, 0.4):
            val = (rng.random(10000) >= nulls) if nulls else None
            for default in (sbo.C_NONE, sbo.C_LZ4):
                roundtrip(ctx, type_, v, validity=val, opts=sbo.make_opts(default, ratio=2.0, force=force))


@pytest.mark.parametrize("type_", INT_TYPES + FLT_TYPES)
def test_adaptive(ctx, type_):
    """test_write_read's options (io.rs:430-434): LZ4 default, ratio 2.0, adaptive on."""
    rng = np.random.default_rng(3)
    n = 10240
    dt = sbo.NP_OF[type_]
    cases = {
        "lowcard": rand_values(rng, type_, n, 8),
        "random": rand_values(rng, type_, n),
        "const": np.full(n, 3, dtype=dt),
        "runs": np.repeat(rand_values(rng, type_, n // 64, 100), 64),
    }
    for name, v in cases.items():
        for nulls in (0.0, 0.2, 0.95):
            val = (rng.random(n) >= nulls) if nulls else None
            roundtrip(ctx, type_, v, validity=val, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))


def test_freq_pattern(ctx):
    """test_freq (io.rs:119-132): 2045 x 20 then 3 x 10000 per page."""
    v = np.tile(np.array([20] * (WRITE_PAGE - 3) + [10000] * 3, dtype=np.uint32), 5)
    roundtrip(ctx, sbo.U32, v, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="Freq")
    roundtrip(ctx, sbo.I64, v.astype(np.int64), opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="Freq")


def test_freq_bitmap_container(ctx):
    """forced Freq with > 4096 exceptions in one 65536-row chunk: Roaring bitmap container."""
    rng = np.random.default_rng(4)
    n = 70000
    v = np.where(rng.random(n) < 0.5, 7, rng.integers(100, 1000, n)).astype(np.int32)
    roundtrip(ctx, sbo.I32, v, page_size=None, opts=sbo.make_opts(force=sbo.C_FREQ))


def test_bitpacking(ctx):
    """test_bitpcking / test_deletabitpacking (io.rs:134-152)."""
    rng = np.random.default_rng(5)
    n = 10240
    for t in (sbo.U32, sbo.I32):
        dt = sbo.NP_OF[t]
        roundtrip(ctx, t, rng.integers(0, 1000, n).astype(dt), opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="Bitpacking")
        roundtrip(ctx, t, np.arange(n).astype(dt), opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="DeltaBitpacking")
        for bits in (1, 7, 16, 31, 32 if t == sbo.U32 else 31):
            hi = (1 << bits) - 1
            v = rng.integers(0, hi, n, endpoint=True).astype(dt)
            roundtrip(ctx, t, v, opts=sbo.make_opts(force=sbo.C_BITPACK), expect_codec="Bitpacking")
            vs = np.sort(v)
            roundtrip(ctx, t, vs, opts=sbo.make_opts(force=sbo.C_DELTABP), expect_codec="DeltaBitpacking")
        roundtrip(ctx, t, np.zeros(n, dtype=dt), opts=sbo.make_opts(force=sbo.C_BITPACK), expect_codec="Bitpacking")


def test_onevalue(ctx):
    """test_onevalue (io.rs:154-165)."""
    n = 10000
    roundtrip(ctx, sbo.BOOL, np.ones(n, bool), opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="OneValue")
    roundtrip(ctx, sbo.U32, np.full(n, 3, np.uint32), opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="OneValue")
    roundtrip(ctx, sbo.F64, np.full(n, 3.5), opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0), expect_codec="OneValue")


def test_patas(ctx):
    rng = np.random.default_rng(6)
    n = 10000
    v = np.cumsum(rng.integers(-3, 4, n)).astype(np.float64) * 0.5
    roundtrip(ctx, sbo.F64, v, opts=sbo.make_opts(force=sbo.C_PATAS), expect_codec="Patas")
    v[::7] = v[3]
    roundtrip(ctx, sbo.F64, v, opts=sbo.make_opts(force=sbo.C_PATAS), expect_codec="Patas")
    roundtrip(ctx, sbo.F64, rng.standard_normal(n), opts=sbo.make_opts(force=sbo.C_PATAS), expect_codec="Patas")


@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
def test_boolean(ctx, default):
    rng = np.random.default_rng(7)
    for n in (1, 7, 8, 9, 2048, 10000, 10007):
        v = rng.random(n) < 0.5
        roundtrip(ctx, sbo.BOOL, v, opts=sbo.make_opts(default))
        val = rng.random(n) >= 0.1
        roundtrip(ctx, sbo.BOOL, v, validity=val, opts=sbo.make_opts(default))
        roundtrip(ctx, sbo.BOOL, v, validity=val, page_size=1001, opts=sbo.make_opts(default))
    runs = np.repeat(rng.random(200) < 0.5, 50)
    roundtrip(ctx, sbo.BOOL, np.repeat(rng.random(40) < 0.5, 1000), page_size=8192, opts=sbo.make_opts(default, ratio=2.0), expect_codec="Rle")
    roundtrip(ctx, sbo.BOOL, runs, page_size=777, opts=sbo.make_opts(default, force=sbo.C_RLE), expect_codec="Rle")
    roundtrip(ctx, sbo.BOOL, rng.random(5000) < 0.5, page_size=777, validity=rng.random(5000) > 0.3, opts=sbo.make_opts(default, force=sbo.C_RLE))


def test_ragged_pages(ctx):
    """page sizes that leave every page start unaligned (bits and 16-byte vectors)."""
    rng = np.random.default_rng(8)
    for t in (sbo.I8, sbo.I16, sbo.I32, sbo.I64, sbo.F64):
        v = rand_values(rng, t, 5003, 16)
        val = rng.random(5003) >= 0.25
        for ps in (1, 3, 129, 1001, None):
            for opts in (sbo.make_opts(), sbo.make_opts(sbo.C_LZ4, ratio=1.5), sbo.make_opts(force=sbo.C_RLE), sbo.make_opts(force=sbo.C_DICT)):
                if ps == 1 and t != sbo.I32:
                    continue
                roundtrip(ctx, t, v, validity=val, page_size=ps, opts=opts)


def test_single_large_page(ctx):
    """max_page_size = None: one page for the whole column (oversized, tiled path)."""
    rng = np.random.default_rng(9)
    n = 300_000
    roundtrip(ctx, sbo.I64, rand_values(rng, sbo.I64, n), page_size=None)
    roundtrip(ctx, sbo.I64, np.full(n, 5, np.int64), page_size=None, opts=sbo.make_opts(ratio=2.0), expect_codec="OneValue")
    roundtrip(ctx, sbo.I32, rand_values(rng, sbo.I32, n, 8), page_size=None, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
    roundtrip(ctx, sbo.I32, np.sort(rand_values(rng, sbo.I32, 128 * 2000, 1 << 30)), page_size=None, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
    val = rng.random(n) > 0.5
    roundtrip(ctx, sbo.F64, rand_values(rng, sbo.F64, n), validity=val, page_size=None)
    roundtrip(ctx, sbo.I64, np.repeat(rand_values(rng, sbo.I64, n // 100), 100), page_size=None, opts=sbo.make_opts(force=sbo.C_RLE))


def test_null_and_empty(ctx):
    data, metas = oracle_encode_column(sbo.NULL, 100, page_size=30, n=100)
    dec = ctx.batch_read_array(sb.Column(sb.NULL, False, data, metas))
    assert dec.length == 100
    dec = ctx.batch_read_array(sb.Column(sb.I32, False, b"", []))
    assert dec.length == 0


def test_many_columns_one_call(ctx):
    rng = np.random.default_rng(10)
    cols, refs = [], []
    for i, t in enumerate(INT_TYPES + FLT_TYPES + [sbo.BOOL]):
        n = 5000 + 17 * i
        v = (rng.random(n) < 0.5) if t == sbo.BOOL else rand_values(rng, t, n, 50)
        val = rng.random(n) > 0.2 if i % 2 else None
        data, metas = oracle_encode_column(t, v, val, page_size=1000, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
        cols.append(sb.Column(t, val is not None, data, metas))
        refs.append((oracle_decode_column(t, val is not None, data, metas), t, val is not None))
    decs = ctx.decode_columns(cols)
    for d, (r, t, nu) in zip(decs, refs):
        assert_same(d, r, t, nu)


def test_corrupt_pages_report_status(ctx):
    """malformed input yields a per-page status, never a crash or an OOB read."""
    rng = np.random.default_rng(11)
    v = rand_values(rng, sbo.I32, 4096, 8)
    for opts in (sbo.make_opts(), sbo.make_opts(force=sbo.C_DICT), sbo.make_opts(force=sbo.C_RLE), sbo.make_opts(force=sbo.C_FREQ), sbo.make_opts(force=sbo.C_BITPACK), sbo.make_opts(sbo.C_LZ4)):
        data, metas = oracle_encode_column(sbo.I32, v, page_size=2048, opts=opts)
        bad = bytearray(data)
        bad[0] = 99  # unknown codec id in page 0
        res = ctx.decode_columns([sb.Column(sb.I32, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] == sb._capi.SB_OUT_OF_SPEC and res.page_status[1] == 0
        # truncated page: claim the payload is longer than the page
        bad = bytearray(data)
        bad[1:5] = (0x7fffffff).to_bytes(4, "little")
        res = ctx.decode_columns([sb.Column(sb.I32, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] != 0 and res.page_status[1] == 0
        with pytest.raises(sb.StrawboatError):
            ctx.decode_columns([sb.Column(sb.I32, False, bytes(bad), metas)])


def test_concurrent_contexts(ctx):
    """one context per host thread (the reference's readers are per-task objects, deserialize.rs:28):
    calls on different contexts overlap on the device and must not disturb each other"""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(12)
    jobs = []
    for i, t in enumerate([sbo.I32, sbo.I64, sbo.F64, sbo.I16, sbo.U8, sbo.F32]):
        v = rand_values(rng, t, 40_000 + 13 * i, 200 if i % 2 else None)
        val = rng.random(len(v)) > 0.3 if i % 3 == 0 else None
        data, metas = oracle_encode_column(t, v, val, page_size=2048, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
        jobs.append((t, val is not None, data, metas, oracle_decode_column(t, val is not None, data, metas)))
    ctxs = [sb.Context(0) for _ in range(3)]

    def work(k):
        out = []
        for rep in range(4):
            for j in range(k, len(jobs), len(ctxs)):
                t, nu, data, metas, _ = jobs[j]
                out.append((j, ctxs[k].batch_read_array(sb.Column(t, nu, data, metas))))
        return out

    with ThreadPoolExecutor(len(ctxs)) as ex:
        results = list(ex.map(work, range(len(ctxs))))
    for res in results:
        for j, dec in res:
            assert_same(dec, jobs[j][4], jobs[j][0], jobs[j][1])
    for c in ctxs:
        c.close()
