"""GPU parity for nested page ENCODE (write_nested + write_nested_validity,
src/write/serialize.rs:135-198,217-232, page loop src/write/common.rs:71-115):
  (i)   pages are cut by top-level rows and carry PageMeta.num_values = level entries;
  (ii)  without LZ4 the page bytes are identical to the oracle writer's (level streams included);
  (iii) the oracle reader and the GPU reader rebuild the same NestedState + leaf buffers from them."""
import numpy as np
import pytest
import sbo
from helpers import assert_same_nested, gen_levels, leaf_slots, oracle_decode_column, write_nested_page

import strawboat_b200 as sb
from test_nested_gpu import SHAPES

pytestmark = pytest.mark.gpu

BIN = (sbo.BINARY, sbo.LARGE_BINARY)


def column_values(type_, rng, n, card):
    if type_ in BIN:
        lens = rng.integers(0, 6, n)
        off = np.zeros(n + 1, np.int64 if type_ == sbo.LARGE_BINARY else np.int32)
        np.cumsum(lens, out=off[1:])
        return (off, rng.integers(97, 100, int(off[-1])).astype(np.uint8))
    if type_ == sbo.BOOL:
        return rng.random(n) < 0.5
    return rng.integers(0, card, n).astype(sbo.NP_OF[type_])


def slice_leaf(type_, values, o, l):
    if type_ in BIN:
        return (values[0][o:o + l + 1], values[1], len(values[1]))  # a sliced array keeps the whole values buffer
    return values[o:o + l]


def check(ctx, type_, nested, rows, page_rows, seed, default=sbo.C_NONE, ratio=None, p_null=0.15, card=1000):
    rng = np.random.default_rng(seed)
    rows_entries = gen_levels(nested, rng, rows, p_null=p_null)
    reps = np.array([e[0] for r in rows_entries for e in r], np.uint32)
    defs = np.array([e[1] for r in rows_entries for e in r], np.uint32)
    n_slots, lval = leaf_slots(nested, reps, defs)
    leaf_nullable = bool(nested[-1][1])
    values = column_values(type_, rng, n_slots, card)
    arr = sb.LeafArray(type_, values, validity=lval if leaf_nullable else None, nullable=leaf_nullable, nested=nested,
                       rep_levels=reps, def_levels=defs, rows=rows)
    assert arr.length == n_slots
    enc = ctx.encode_columns([arr], sb.write_options(default, ratio, page_rows, seed=7))[0]

    # the oracle writer, page by page, on the same slices
    pages, metas, cur = [], [], [0]
    for pi, o in enumerate(range(0, rows, page_rows)):
        opts = sbo.make_opts(default, ratio=ratio, float_bitwise=1)
        opts.seed = 7 + pi

        def leaf_fn(n, validity):
            v = slice_leaf(type_, values, cur[0], n)
            assert np.array_equal(validity, lval[cur[0]:cur[0] + n])
            cur[0] += n
            return v
        page, nv = write_nested_page(type_, nested, rows_entries[o:o + page_rows], leaf_fn, opts)
        pages.append(page)
        metas.append((len(page), nv))
    # (i)
    assert [m[1] for m in enc.metas] == [m[1] for m in metas]
    assert sum(m[0] for m in enc.metas) == len(enc.data)
    # (ii)
    if default == sbo.C_NONE:
        pos = 0
        for i, (page, (ln, _)) in enumerate(zip(pages, enc.metas)):
            got = enc.data[pos:pos + ln]
            pos += ln
            if "Lz4" not in sbo.stat_block(type_, page[sbo_value_block(page):]):
                assert got == page, (i, len(got), len(page))
    # (iii)
    ref = oracle_decode_column(type_, leaf_nullable, b"".join(pages), metas, nested)
    mine_by_oracle = oracle_decode_column(type_, leaf_nullable, enc.data, enc.metas, nested)
    assert mine_by_oracle["length"] == ref["length"] == n_slots
    dec = ctx.batch_read_array(sb.Column(type_, leaf_nullable, enc.data, enc.metas, nested))
    assert_same_nested(dec, mine_by_oracle, type_, nested)
    for d in range(len(nested) - 1):
        a, b = mine_by_oracle["nested"][d], ref["nested"][d]
        assert np.array_equal(a["offsets"], b["offsets"]) and a["validity_len"] == b["validity_len"]
    # leaf values on valid slots
    if type_ not in BIN and type_ != sbo.BOOL:
        v = np.ascontiguousarray(values, dtype=sbo.NP_OF[type_])
        m = lval if leaf_nullable else np.ones(n_slots, bool)
        assert np.array_equal(dec.values[m], v[m])
    return enc


def sbo_value_block(page):
    """offset of the VALUE_BLOCK of a nested page: 12 + rep_len + def_len"""
    h = np.frombuffer(page[:12], "<u4")
    return 12 + int(h[1]) + int(h[2])


@pytest.mark.parametrize("shape", list(SHAPES))
def test_shapes_i64(ctx, shape):
    nested = SHAPES[shape]
    for k, (rows, page_rows) in enumerate(((1, 1), (37, 10), (3000, 1024), (9000, 4096))):
        check(ctx, sbo.I64, nested, rows, page_rows, seed=100 * list(SHAPES).index(shape) + k)


@pytest.mark.parametrize("type_", [sbo.I32, sbo.F64, sbo.BOOL, sbo.BINARY, sbo.LARGE_BINARY])
@pytest.mark.parametrize("default,ratio", [(sbo.C_NONE, None), (sbo.C_NONE, 1.2), (sbo.C_LZ4, 2.0)])
def test_config4_leaf_types(ctx, type_, default, ratio):
    nested = SHAPES["list<struct<a?>?>?  (config 4)"]
    check(ctx, type_, nested, 5000, 2048, seed=5, default=default, ratio=ratio, card=8)


def test_all_null_and_all_valid(ctx):
    nested = SHAPES["list<struct<a?>?>?  (config 4)"]
    for p_null in (1.0, 0.0):
        check(ctx, sbo.I64, nested, 500, 200, seed=8, p_null=p_null)


def test_page_boundaries_across_level_blocks(ctx):
    """page starts that fall on / next to the 4096-entry blocks of the boundary kernels"""
    nested = SHAPES["list<i64>"]
    for rows, page_rows in ((20000, 4096), (20000, 1), (8192, 8192), (12289, 4097)):
        if page_rows == 1:
            rows = 700
        check(ctx, sbo.I64, nested, rows, page_rows, seed=rows + page_rows)


def test_inconsistent_levels_are_rejected(ctx):
    nested = SHAPES["list<i64>"]
    reps = np.array([0, 1, 0, 1, 1], np.uint32)
    defs = np.array([1, 1, 1, 1, 1], np.uint32)
    vals = np.arange(5, dtype=np.int64)
    ok = sb.LeafArray(sbo.I64, vals, nested=nested, rep_levels=reps, def_levels=defs, rows=2)
    assert [m[1] for m in ctx.encode_columns([ok], sb.write_options(sbo.C_NONE, None, 1))[0].metas] == [2, 3]
    for rows, n in ((3, 5), (2, 4)):
        bad = sb.LeafArray(sbo.I64, vals[:n], nested=nested, rep_levels=reps, def_levels=defs, rows=rows)
        with pytest.raises(sb.StrawboatError) as e:
            ctx.encode_columns([bad], sb.write_options(sbo.C_NONE, None, 1))
        assert e.value.code == sb._capi.SB_INVALID_ARG
