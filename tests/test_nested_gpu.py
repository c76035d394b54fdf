"""GPU parity for nested pages: rep/def level decode (read_validity_nested,
src/read/read_basic.rs:65-173) + the leaf value block, against the oracle."""
import numpy as np
import pytest
import sbo
from helpers import assert_same_nested, gen_levels, oracle_decode_column, write_nested_page

import strawboat_b200 as sb

pytestmark = pytest.mark.gpu

L, S, P = sbo.N_LIST, sbo.N_STRUCT, sbo.N_PRIMITIVE

SHAPES = {
    "list<i64?>?": [(L, True), (P, True)],
    "list<i64>": [(L, False), (P, False)],
    "list<struct<a?>?>?  (config 4)": [(L, True), (S, True), (P, True)],
    "struct<a?>": [(S, False), (P, True)],
    "struct?<a?>": [(S, True), (P, True)],
    "list<list<a?>?>?": [(L, True), (L, True), (P, True)],
    "list<struct<list<a>>>": [(L, False), (S, False), (L, False), (P, False)],
    "struct<struct?<list?<a?>>>": [(S, False), (S, True), (L, True), (P, True)],
}


def leaf_fn(type_, rng, card=None):
    def f(n, validity):
        if type_ in (sbo.BINARY, sbo.LARGE_BINARY):
            lens = rng.integers(0, 6, n)
            off = np.zeros(n + 1, np.int64 if type_ == sbo.LARGE_BINARY else np.int32)
            np.cumsum(lens, out=off[1:])
            return (off, rng.integers(97, 123, int(off[-1])).astype(np.uint8))
        if type_ == sbo.BOOL:
            return rng.random(n) < 0.5
        dt = sbo.NP_OF[type_]
        return rng.integers(0, card or 1000, n).astype(dt)
    return f


def build_column(type_, nested, rng, rows, page_rows, opts=None, card=None, p_null=0.15):
    rows_entries = gen_levels(nested, rng, rows, p_null=p_null)
    pages, metas = [], []
    for o in range(0, rows, page_rows):
        page, nv = write_nested_page(type_, nested, rows_entries[o:o + page_rows], leaf_fn(type_, rng, card), opts)
        pages.append(page)
        metas.append((len(page), nv))
    return b"".join(pages), metas


@pytest.mark.parametrize("shape", list(SHAPES))
def test_shapes_i64(ctx, shape):
    nested = SHAPES[shape]
    rng = np.random.default_rng(list(SHAPES).index(shape))
    for rows, page_rows in ((1, 1), (37, 10), (3000, 1024), (9000, 8192)):
        data, metas = build_column(sbo.I64, nested, rng, rows, page_rows)
        ref = oracle_decode_column(sbo.I64, nested[-1][1], data, metas, nested)
        dec = ctx.batch_read_array(sb.Column(sb.I64, nested[-1][1], data, metas, nested))
        assert_same_nested(dec, ref, sbo.I64, nested)


@pytest.mark.parametrize("type_", [sbo.I32, sbo.F64, sbo.BOOL, sbo.BINARY, sbo.LARGE_BINARY])
def test_config4_leaf_types(ctx, type_):
    """List<Struct<i64, f64, utf8>>-style leaves with adaptive codecs on the leaf values."""
    nested = SHAPES["list<struct<a?>?>?  (config 4)"]
    rng = np.random.default_rng(7)
    for opts in (sbo.make_opts(), sbo.make_opts(sbo.C_LZ4, ratio=2.0)):
        data, metas = build_column(type_, nested, rng, 5000, 2048, opts=opts, card=8)
        ref = oracle_decode_column(type_, True, data, metas, nested)
        dec = ctx.batch_read_array(sb.Column(type_, True, data, metas, nested))
        assert_same_nested(dec, ref, type_, nested)


def test_all_null_and_all_empty(ctx):
    nested = SHAPES["list<struct<a?>?>?  (config 4)"]
    rng = np.random.default_rng(8)
    for p_null in (1.0, 0.0):
        data, metas = build_column(sbo.I64, nested, rng, 500, 200, p_null=p_null)
        ref = oracle_decode_column(sbo.I64, True, data, metas, nested)
        dec = ctx.batch_read_array(sb.Column(sb.I64, True, data, metas, nested))
        assert_same_nested(dec, ref, sbo.I64, nested)


def hybrid_mixed(levels, w, rng):
    """level stream mixing RLE runs and bit-packed runs (a parquet-style writer)."""
    out = bytearray()
    i, n = 0, len(levels)

    def uleb(v):
        b = bytearray()
        while True:
            if v < 0x80:
                b.append(v)
                return bytes(b)
            b.append((v & 0x7f) | 0x80)
            v >>= 7
    while i < n:
        j = i
        while j < n and levels[j] == levels[i]:
            j += 1
        if j - i >= 8 and rng.random() < 0.7:
            out += uleb((j - i) << 1) + bytes([levels[i]])
            i = j
        else:
            groups = int(rng.integers(1, 5))
            take = min(n - i, groups * 8)
            groups = (take + 7) // 8
            vals = list(levels[i:i + take]) + [0] * (groups * 8 - take)
            acc, nb = 0, 0
            body = bytearray()
            for v in vals:
                acc |= v << nb
                nb += w
                while nb >= 8:
                    body.append(acc & 0xff)
                    acc >>= 8
                    nb -= 8
            out += uleb((groups << 1) | 1) + bytes(body)
            i += take
    return bytes(out)


def test_mixed_rle_and_bitpacked_level_runs(ctx):
    """HybridRleDecoder accepts RLE runs too (read_basic.rs:83-84); arrow2 only writes bit-packed."""
    from helpers import leaf_slots, nested_thresholds
    nested = SHAPES["list<struct<a?>?>?  (config 4)"]
    rng = np.random.default_rng(9)
    rows_entries = gen_levels(nested, rng, 4000, p_null=0.5)
    cum_sum, cum_rep = nested_thresholds(nested)
    pages, metas = [], []
    for o in range(0, 4000, 1000):
        ents = [e for r in rows_entries[o:o + 1000] for e in r]
        reps, defs = [e[0] for e in ents], [e[1] for e in ents]
        rb = hybrid_mixed(reps, cum_rep[-1].bit_length(), rng)
        db = hybrid_mixed(defs, cum_sum[-1].bit_length(), rng)
        assert list(sbo.hybrid_rle_decode(rb, cum_rep[-1].bit_length(), len(reps))) == reps
        n_slots, lval = leaf_slots(nested, reps, defs)
        block = sbo.compress_values(sbo.I64, rng.integers(0, 50, n_slots).astype(np.int64), validity=lval)
        page = np.array([1000, len(rb), len(db)], "<u4").tobytes() + rb + db + block
        pages.append(page)
        metas.append((len(page), len(reps)))
    data = b"".join(pages)
    ref = oracle_decode_column(sbo.I64, True, data, metas, nested)
    dec = ctx.batch_read_array(sb.Column(sb.I64, True, data, metas, nested))
    assert_same_nested(dec, ref, sbo.I64, nested)
