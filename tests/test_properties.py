"""Property tests (hypothesis) of the page path, modelled on the reference's only test oracle --
round-trip equality over random chunks and the whole option matrix (tests/it/io.rs:167-278, 417-497):

  * CPU : oracle write -> oracle read == input, for every codec that applies (forced), any page size, any
          null pattern, NaN / -0.0 payloads, ragged last pages;
  * GPU : GPU read of oracle-written pages == input, and oracle read of GPU-written pages == input.

Size-independent properties only; the byte-level comparisons live in the parity tests."""
import numpy as np
import pytest
import sbo
from helpers import oracle_decode_column, oracle_encode_column
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

INT_TYPES = [sbo.I8, sbo.I16, sbo.I32, sbo.I64, sbo.U8, sbo.U16, sbo.U32, sbo.U64]
FIXED = INT_TYPES + [sbo.F32, sbo.F64]
def common(n):
    return settings(max_examples=n, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True, database=None)


@st.composite
def fixed_column(draw):
    t = draw(st.sampled_from(FIXED))
    n = draw(st.sampled_from([0, 1, 2, 127, 128, 129, 640, 641, 1000, 2048, 3000]))
    seed = draw(st.integers(0, 2**31))
    shape = draw(st.sampled_from(["random", "lowcard", "const", "sorted", "runs", "freq", "special"]))
    rng = np.random.default_rng(seed)
    dt = np.dtype(sbo.NP_OF[t])
    if dt.kind == "f":
        base = {"random": rng.standard_normal(n) * 1e6, "lowcard": rng.integers(0, 7, n).astype(np.float64),
                "const": np.full(n, 3.25), "sorted": np.sort(rng.standard_normal(n)), "runs": np.repeat(rng.standard_normal(n // 17 + 1), 17)[:n],
                "freq": np.where(rng.random(n) < 0.95, 1.5, rng.standard_normal(n)),
                "special": rng.choice(np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1.0, 5e-324]), n)}[shape]
        v = base.astype(dt)
    else:
        info = np.iinfo(dt)
        hi = min(int(info.max), 1 << 40)
        base = {"random": rng.integers(info.min, info.max, n, dtype=dt, endpoint=True),
                "lowcard": rng.integers(0, 7, n), "const": np.full(n, min(77, int(info.max))),
                "sorted": np.sort(rng.integers(0, hi, n)), "runs": np.repeat(rng.integers(0, hi, n // 17 + 1), 17)[:n],
                "freq": np.where(rng.random(n) < 0.95, min(20, hi), rng.integers(0, hi, n)),
                "special": rng.choice(np.array([info.min, info.max, 0, 1], dtype=dt), n)}[shape]
        v = np.asarray(base).astype(dt)
    nulls = draw(st.sampled_from([None, 0.0, 0.3, 0.95, 1.0]))
    validity = None if nulls is None else (rng.random(n) >= nulls)
    page = draw(st.sampled_from([None, 128, 640, 1000]))
    default = draw(st.sampled_from([sbo.C_NONE, sbo.C_LZ4]))
    force = draw(st.sampled_from([-1, -1, sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ, sbo.C_ONEVALUE, sbo.C_BITPACK, sbo.C_DELTABP, sbo.C_PATAS]))
    ratio = draw(st.sampled_from([None, 1.2, 2.0]))
    return t, v, validity, page, sbo.make_opts(default, ratio=ratio, force=force, seed=seed & 0xffff, float_bitwise=1)


def same_valid(t, got, v, validity):
    a, b = np.ascontiguousarray(got), np.ascontiguousarray(v)
    if validity is None:
        return np.array_equal(a.view(np.uint8), b.view(np.uint8))
    return np.array_equal(a[validity].view(np.uint8), b[validity].view(np.uint8))


@common(120)
@given(fixed_column())
def test_oracle_roundtrip_fixed(case):
    t, v, validity, page, opts = case
    data, metas = oracle_encode_column(t, v, validity, page_size=page, opts=opts)
    ref = oracle_decode_column(t, validity is not None, data, metas)
    assert ref["length"] == len(v)
    assert same_valid(t, ref["values"], v, validity)
    if validity is not None and len(v):
        assert np.array_equal(sbo.unpack_bits(ref["validity"], len(v)), validity)


@st.composite
def binary_column(draw):
    t = draw(st.sampled_from([sbo.BINARY, sbo.LARGE_BINARY]))
    n = draw(st.sampled_from([0, 1, 50, 777, 2048]))
    seed = draw(st.integers(0, 2**31))
    rng = np.random.default_rng(seed)
    uniq = draw(st.sampled_from([1, 5, 300, None]))
    if uniq is None:
        lens = rng.integers(0, 30, n)
    else:
        tl = rng.integers(0, 12, uniq)
        ids = rng.integers(0, uniq, n)
        lens = tl[ids]
    nulls = draw(st.sampled_from([None, 0.4, 1.0]))
    validity = None if nulls is None else (rng.random(n) >= nulls)
    if validity is not None:
        lens = np.where(validity, lens, 0)
    off = np.zeros(n + 1, np.int64 if t == sbo.LARGE_BINARY else np.int32)
    np.cumsum(lens, out=off[1:])
    if uniq is None:
        dat = rng.integers(0, 256, int(off[-1])).astype(np.uint8)
    else:  # equal ids -> equal bytes
        tab = [rng.integers(97, 123, int(x)).astype(np.uint8) for x in tl]
        parts = [tab[i] for i, ok in zip(ids, np.ones(n, bool) if validity is None else validity) if ok]
        dat = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    page = draw(st.sampled_from([None, 100, 1000]))
    force = draw(st.sampled_from([-1, sbo.C_DICT, sbo.C_FREQ, sbo.C_ONEVALUE]))
    default = draw(st.sampled_from([sbo.C_NONE, sbo.C_LZ4]))
    return t, (off, dat), validity, page, sbo.make_opts(default, ratio=draw(st.sampled_from([None, 2.0])), force=force, seed=seed & 0xffff)


def valid_rows_bytes(off, dat, validity):
    """(lengths, concatenated bytes) of the valid rows"""
    off = np.asarray(off, dtype=np.int64)
    lens = np.diff(off)
    keep = np.ones(len(lens), bool) if validity is None else np.asarray(validity, bool)
    starts, l = off[:-1][keep], lens[keep]
    total = int(l.sum())
    if total == 0:
        return l, np.zeros(0, np.uint8)
    row = np.repeat(np.arange(len(l)), l)
    pos = np.arange(total) - np.repeat(np.cumsum(l) - l, l)
    return l, np.asarray(dat)[starts[row] + pos]


def check_binary(ref, off, dat, validity):
    """valid rows carry their bytes; what a null slot decodes to is codec dependent in the reference (Dict / Freq /
    OneValue substitute a neighbour's value: binary/dict.rs:66-74), so only the valid rows are compared"""
    n = len(off) - 1
    assert ref["length"] == n
    if n == 0:  # a column without pages: nothing was ever pushed
        return
    assert len(ref["offsets"]) == n + 1 and ref["offsets"][0] == 0
    assert np.all(np.diff(ref["offsets"]) >= 0) and ref["offsets"][-1] == len(ref["values"])
    gl, gb = valid_rows_bytes(ref["offsets"], ref["values"], validity)
    el, eb = valid_rows_bytes(off, dat, validity)
    assert np.array_equal(gl, el) and np.array_equal(gb, eb)
    if validity is not None:
        assert np.array_equal(sbo.unpack_bits(ref["validity"], n), validity)


@common(60)
@given(binary_column())
def test_oracle_roundtrip_binary(case):
    t, (off, dat), validity, page, opts = case
    data, metas = oracle_encode_column(t, (off, dat), validity, page_size=page, opts=opts)
    check_binary(oracle_decode_column(t, validity is not None, data, metas), off, dat, validity)


@common(40)
@given(st.integers(0, 3000), st.integers(0, 2**31), st.sampled_from([None, 0.3]), st.sampled_from([-1, sbo.C_RLE, sbo.C_ONEVALUE]),
       st.sampled_from([0.5, 0.02, 1.0]), st.sampled_from([None, 333]))
def test_oracle_roundtrip_boolean(n, seed, nulls, force, p_true, page):
    rng = np.random.default_rng(seed)
    v = rng.random(n) < p_true
    validity = None if nulls is None else rng.random(n) >= nulls
    data, metas = oracle_encode_column(sbo.BOOL, v, validity, page_size=page, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0, force=force))
    ref = oracle_decode_column(sbo.BOOL, validity is not None, data, metas)
    got = sbo.unpack_bits(ref["values"], n)
    assert np.array_equal(got if validity is None else got[validity], v if validity is None else v[validity])


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@common(60)
@given(fixed_column())
def test_gpu_reads_oracle_pages(ctx, case):
    import strawboat_b200 as sb
    t, v, validity, page, opts = case
    data, metas = oracle_encode_column(t, v, validity, page_size=page, opts=opts)
    dec = ctx.batch_read_array(sb.Column(t, validity is not None, data, metas))
    assert dec.length == len(v) and same_valid(t, dec.values, v, validity)
    if validity is not None and len(v):
        assert np.array_equal(sbo.unpack_bits(dec.validity, len(v)), validity)


@pytest.mark.gpu
@common(60)
@given(fixed_column())
def test_oracle_reads_gpu_pages(ctx, case):
    import strawboat_b200 as sb
    t, v, validity, page, opts = case
    wo = sb.write_options(opts.default_compression, None if opts.default_compress_ratio < 0 else opts.default_compress_ratio, page,
                          force_codec=opts.force_codec, seed=opts.seed)
    enc = ctx.encode_columns([sb.LeafArray(t, v, validity=validity)], wo)[0]
    assert [m[1] for m in enc.metas] == [l for _, l in sb.workloads.split_pages(len(v), page)]
    ref = oracle_decode_column(t, validity is not None, enc.data, enc.metas)
    assert ref["length"] == len(v) and same_valid(t, ref["values"], v, validity)
    dec = ctx.batch_read_array(sb.Column(t, validity is not None, enc.data, enc.metas))
    assert same_valid(t, dec.values, v, validity)


@pytest.mark.gpu
@common(40)
@given(binary_column())
def test_gpu_binary_both_ways(ctx, case):
    import strawboat_b200 as sb
    t, (off, dat), validity, page, opts = case
    data, metas = oracle_encode_column(t, (off, dat), validity, page_size=page, opts=opts)
    dec = ctx.batch_read_array(sb.Column(t, validity is not None, data, metas))
    check_binary({"length": dec.length, "offsets": dec.offsets, "values": dec.values, "validity": dec.validity}, off, dat, validity)
    wo = sb.write_options(opts.default_compression, None if opts.default_compress_ratio < 0 else opts.default_compress_ratio, page,
                          force_codec=opts.force_codec, seed=opts.seed)
    enc = ctx.encode_columns([sb.LeafArray(t, (off, dat), validity=validity)], wo)[0]
    check_binary(oracle_decode_column(t, validity is not None, enc.data, enc.metas), off, dat, validity)
