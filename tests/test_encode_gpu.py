"""GPU parity for the encode path (compress_integer / compress_double / compress_binary /
compress_boolean + write_validity, through sb_encode_columns):
  (i)   pages written on the GPU decode through the ORACLE reader (the restatement of the
        reference decoder, LZ4 blocks through liblz4) to the input, bit for bit on valid slots;
  (ii)  the chooser picks the same codec tree as the oracle's chooser for the same sampler seed;
  (iii) with default_compression = None the page bytes are identical to the oracle writer's;
  (iv)  GPU decode(GPU encode(x)) == x."""
import numpy as np
import pytest
import sbo
from helpers import oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb
from strawboat_b200.workloads import random_strings

pytestmark = pytest.mark.gpu

INT_TYPES = [sbo.I8, sbo.I16, sbo.I32, sbo.I64, sbo.U8, sbo.U16, sbo.U32, sbo.U64]
FLT_TYPES = [sbo.F32, sbo.F64]
BIN = (sbo.BINARY, sbo.LARGE_BINARY)


def split(data, metas):
    out, pos = [], 0
    for ln, nv in metas:
        out.append(data[pos:pos + ln])
        pos += ln
    return out


def check(ctx, type_, values, validity=None, nullable=None, page_size=2048, default=sbo.C_NONE, ratio=None, force=-1, seed=7,
          same_tree=True, expect=None):
    if nullable is None:
        nullable = validity is not None
    arr = sb.LeafArray(type_, values, validity=validity, nullable=nullable)
    wo = sb.write_options(default, ratio, page_size, force_codec=force, seed=seed)
    enc = ctx.encode_columns([arr], wo)[0]
    n = arr.length
    assert sum(m[1] for m in enc.metas) == n
    assert sum(m[0] for m in enc.metas) == len(enc.data)
    # (i) oracle decode
    ref = oracle_decode_column(type_, nullable, enc.data, enc.metas)
    assert ref["length"] == n
    vmask = np.ones(n, bool) if validity is None else np.asarray(validity, bool)
    if type_ == sbo.BOOL:
        got = sbo.unpack_bits(ref["values"], n)
        assert np.array_equal(got[vmask], np.asarray(values, bool)[vmask])
    elif type_ in BIN:
        off, dat = np.asarray(values[0]).astype(np.int64), np.asarray(values[1], np.uint8)
        ro = ref["offsets"].astype(np.int64)
        lens_in, lens_out = np.diff(off), np.diff(ro)
        assert np.array_equal(lens_in[vmask], lens_out[vmask])
        if vmask.all():
            assert np.array_equal(ref["values"], dat[off[0]:off[-1]])
        else:
            for i in np.flatnonzero(vmask)[:: max(1, n // 500)]:
                assert np.array_equal(ref["values"][ro[i]:ro[i + 1]], dat[off[i]:off[i + 1]])
    elif type_ != sbo.NULL and n:
        v = np.ascontiguousarray(values, dtype=sbo.NP_OF[type_])
        assert np.array_equal(ref["values"].view(np.uint8).reshape(n, -1)[vmask], v.view(np.uint8).reshape(n, -1)[vmask])
    if nullable and n:
        assert np.array_equal(sbo.unpack_bits(ref["validity"], n), vmask)
    # (ii) / (iii) oracle writer with the same options and seeds
    opts = sbo.make_opts(default, ratio=ratio, force=force, float_bitwise=1)
    odata, ometas = oracle_encode_column(type_, values, validity, nullable, page_size, opts, seed=seed)
    gp, op = split(enc.data, enc.metas), split(odata, ometas)
    assert len(gp) == len(op)
    if type_ != sbo.NULL and same_tree:
        for i, (g, o) in enumerate(zip(gp, op)):
            tg, to = sbo.stat_page(type_, nullable, g), sbo.stat_page(type_, nullable, o)
            assert tg == to, (i, tg, to)
            if default == sbo.C_NONE and "Lz4" not in to:
                assert g == o, (i, tg)
        if expect is not None:
            assert sbo.stat_page(type_, nullable, gp[0]).startswith(expect), sbo.stat_page(type_, nullable, gp[0])
    # (iv) GPU decode of the GPU pages == oracle decode of the same pages
    dec = ctx.batch_read_array(sb.Column(type_, nullable, enc.data, enc.metas))
    if type_ == sbo.BOOL:
        assert np.array_equal(sbo.unpack_bits(dec.values, n), sbo.unpack_bits(ref["values"], n))
    elif type_ in BIN:
        assert np.array_equal(dec.offsets, ref["offsets"]) and np.array_equal(dec.values, ref["values"])
    elif type_ != sbo.NULL:
        assert np.array_equal(dec.values.view(np.uint8), ref["values"].view(np.uint8))
    return enc


def rand_values(rng, type_, n, card=None):
    dt = sbo.NP_OF[type_]
    if type_ in FLT_TYPES:
        return rng.integers(0, card, n).astype(dt) if card else rng.standard_normal(n).astype(dt)
    info = np.iinfo(dt)
    if card:
        return rng.integers(0, min(card, int(info.max)), n).astype(dt)
    return rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)


def test_basic_chunk(ctx):
    for t in INT_TYPES:
        check(ctx, t, np.arange(1, 7).astype(sbo.NP_OF[t]))
    for t in FLT_TYPES:
        check(ctx, t, np.array([1.1, 2.2, 3.3, 4.4, 5.5, 6.6], dtype=sbo.NP_OF[t]))
    check(ctx, sbo.BOOL, np.array([1, 1, 1, 0, 0, 0], bool))
    check(ctx, sbo.NULL, 100, page_size=30)
    check(ctx, sbo.I32, np.zeros(0, np.int32))


@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
@pytest.mark.parametrize("type_", INT_TYPES + FLT_TYPES)
def test_plain(ctx, type_, default):
    rng = np.random.default_rng(1)
    v = rand_values(rng, type_, 10000)
    check(ctx, type_, v, default=default)
    check(ctx, type_, v, validity=rng.random(10000) >= 0.3, default=default, page_size=1001)
    check(ctx, type_, rand_values(rng, type_, 10000, 4), default=default)  # compressible: real LZ4 matches


@pytest.mark.parametrize("force", [sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ])
@pytest.mark.parametrize("type_", INT_TYPES + FLT_TYPES)
def test_forced(ctx, type_, force):
    rng = np.random.default_rng(2)
    for card in (8, None):
        v = rand_values(rng, type_, 6000, card)
        for nulls in (0.0, 0.4):
            val = (rng.random(6000) >= nulls) if nulls else None
            for default in (sbo.C_NONE, sbo.C_LZ4):
                check(ctx, type_, v, validity=val, default=default, ratio=2.0, force=force)


@pytest.mark.parametrize("type_", INT_TYPES + FLT_TYPES)
def test_adaptive(ctx, type_):
    rng = np.random.default_rng(3)
    n = 10240
    dt = sbo.NP_OF[type_]
    cases = {
        "lowcard": rand_values(rng, type_, n, 8),
        "random": rand_values(rng, type_, n),
        "const": np.full(n, 3, dtype=dt),
        "runs": np.repeat(rand_values(rng, type_, n // 64, 100), 64),
        "freq": np.where(rng.random(n) < 0.95, 3, rand_values(rng, type_, n, 120)).astype(dt),
    }
    for name, v in cases.items():
        for nulls in (0.0, 0.2, 0.95):
            val = (rng.random(n) >= nulls) if nulls else None
            for default in (sbo.C_NONE, sbo.C_LZ4):
                # f32 Patas is never chosen on the GPU: trees may differ there
                check(ctx, type_, v, validity=val, default=default, ratio=2.0, same_tree=type_ != sbo.F32)


def test_expected_codecs(ctx):
    rng = np.random.default_rng(4)
    n = 8192 * 2
    o = dict(default=sbo.C_LZ4, ratio=2.0, page_size=8192)
    check(ctx, sbo.U32, np.tile(np.array([20] * 2045 + [10000] * 3, np.uint32), 8), expect="Freq", **o)
    check(ctx, sbo.U32, rng.integers(0, 1000, n).astype(np.uint32), expect="Bitpacking", **o)
    check(ctx, sbo.I32, np.arange(n).astype(np.int32), expect="DeltaBitpacking", **o)
    check(ctx, sbo.I32, rng.integers(0, 8, n).astype(np.int32), expect="Dict(Bitpacking", **o)
    check(ctx, sbo.I64, np.full(n, 7, np.int64), expect="OneValue", **o)
    check(ctx, sbo.I64, np.repeat(rng.integers(0, 1 << 40, n // 64), 64), expect="Rle", **o)
    check(ctx, sbo.F64, rng.integers(0, 8, n).astype(np.float64), expect="Dict", **o)
    check(ctx, sbo.F64, rng.integers(0, 65536, n).astype(np.float64), expect="Lz4", **o)
    check(ctx, sbo.F64, np.cumsum(rng.integers(-3, 4, n)).astype(np.float64) * 0.5, force=sbo.C_PATAS, expect="Patas", default=sbo.C_LZ4, page_size=8192)
    check(ctx, sbo.F64, np.cumsum(rng.integers(-3, 4, n)).astype(np.float64) * 0.5, default=sbo.C_LZ4, ratio=1.2, page_size=8192)
    check(ctx, sbo.I64, rng.integers(-2**62, 2**62, n), expect="Lz4", **o)
    for bits in (1, 7, 16, 31):
        v = rng.integers(0, (1 << bits) - 1, n, endpoint=True).astype(np.int32)
        check(ctx, sbo.I32, v, force=sbo.C_BITPACK, expect="Bitpacking", page_size=8192)
        check(ctx, sbo.I32, np.sort(v), force=sbo.C_DELTABP, expect="DeltaBitpacking", page_size=8192)
    check(ctx, sbo.U32, rng.integers(0, 2**32 - 1, n, endpoint=True).astype(np.uint32), force=sbo.C_BITPACK, expect="Bitpacking", page_size=8192)


def test_freq_bitmap_container_and_big_pages(ctx):
    rng = np.random.default_rng(5)
    n = 150_000
    v = np.where(rng.random(n) < 0.5, 7, rng.integers(100, 1000, n)).astype(np.int32)
    check(ctx, sbo.I32, v, page_size=None, force=sbo.C_FREQ)
    check(ctx, sbo.I64, rng.integers(0, 50, n), page_size=None, default=sbo.C_LZ4, ratio=2.0)
    check(ctx, sbo.I64, rng.integers(-2**62, 2**62, n), page_size=None, default=sbo.C_LZ4)
    check(ctx, sbo.I64, rng.integers(-2**62, 2**62, n), validity=rng.random(n) > 0.5, page_size=None)


@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
def test_boolean(ctx, default):
    rng = np.random.default_rng(6)
    for n in (1, 7, 8, 9, 2048, 10007):
        v = rng.random(n) < 0.5
        check(ctx, sbo.BOOL, v, default=default)
        val = rng.random(n) >= 0.1
        check(ctx, sbo.BOOL, v, validity=val, default=default, page_size=1001)
    check(ctx, sbo.BOOL, np.ones(10000, bool), default=default, ratio=2.0, expect="OneValue")
    check(ctx, sbo.BOOL, np.repeat(rng.random(40) < 0.5, 1000), default=default, ratio=2.0, page_size=8192, expect="Rle")
    check(ctx, sbo.BOOL, rng.random(5000) < 0.5, validity=rng.random(5000) > 0.3, default=default, force=sbo.C_RLE, page_size=777)


def strings(rng, n, uniq, nulls=0.0, large=False, maxlen=None):
    if maxlen is None:
        o, d, v = random_strings(rng, n, uniq, nulls, large=large)
        return (o, d), v
    lens = rng.integers(0, maxlen + 1, n)
    offsets = np.zeros(n + 1, dtype=np.int64 if large else np.int32)
    np.cumsum(lens, out=offsets[1:])
    data = rng.integers(0, 256, int(offsets[-1])).astype(np.uint8)
    return (offsets, data), ((rng.random(n) >= nulls) if nulls else None)


@pytest.mark.parametrize("type_", list(BIN))
def test_binary(ctx, type_):
    large = type_ == sbo.LARGE_BINARY
    rng = np.random.default_rng(7)
    for default in (sbo.C_NONE, sbo.C_LZ4):
        vals, _ = strings(rng, 5000, None, large=large, maxlen=30)
        check(ctx, type_, vals, default=default)
        vals, v = strings(rng, 5000, None, nulls=0.3, large=large, maxlen=30)
        check(ctx, type_, vals, validity=v, default=default, page_size=777)
        for uniq in (1, 8, 1000):
            for nulls in (0.0, 0.4):
                vals, v = strings(rng, 8192 + 100, uniq, nulls=nulls, large=large)
                check(ctx, type_, vals, validity=v, default=default, ratio=2.0, page_size=4096)
                for force in (sbo.C_DICT, sbo.C_FREQ):
                    check(ctx, type_, vals, validity=v, default=default, ratio=2.0, force=force, page_size=4096)
    o, d, _ = random_strings(rng, 8192 * 2, 100, 0.0, large=large, sort_within=8192)
    check(ctx, type_, (o, d), default=sbo.C_LZ4, ratio=2.0, page_size=8192, expect="Dict(Rle")


def test_many_columns_one_call(ctx):
    rng = np.random.default_rng(8)
    arrays, inputs = [], []
    for i, t in enumerate(INT_TYPES + FLT_TYPES + [sbo.BOOL, sbo.BINARY]):
        n = 5000 + 17 * i
        if t == sbo.BOOL:
            v = rng.random(n) < 0.5
        elif t == sbo.BINARY:
            v = strings(rng, n, 20)[0]
        else:
            v = rand_values(rng, t, n, 50)
        val = rng.random(n) > 0.2 if i % 2 else None
        arrays.append(sb.LeafArray(t, v, validity=val))
        inputs.append((t, v, val, n))
    encs = ctx.encode_columns(arrays, sb.write_options(sbo.C_LZ4, 2.0, 1000, seed=3))
    for e, (t, v, val, n) in zip(encs, inputs):
        ref = oracle_decode_column(t, val is not None, e.data, e.metas)
        assert ref["length"] == n
        if t not in (sbo.BOOL, sbo.BINARY):
            m = np.ones(n, bool) if val is None else val
            assert np.array_equal(ref["values"][m], np.asarray(v, sbo.NP_OF[t])[m])
