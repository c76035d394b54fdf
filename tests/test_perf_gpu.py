"""Throughput of the BASELINE.json configs that bench.py does not time (configs[2] strings,
configs[3] nested), at sizes the oracle can still check.  Each case asserts parity (GPU decode ==
oracle decode, bit for bit) and records device time + GB/s in gpurun_out/perf_configs.json;
profiles/ holds the copy of the round.  No throughput assertion: the numbers are measurements."""
import json
import os

import numpy as np
import pytest
import sbo
from helpers import assert_same, assert_same_nested

import strawboat_b200 as sb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RESULTS = {}
PAGE = 8192


from strawboat_b200.workloads import config4_levels, dict_strings  # noqa: E402


def oracle_column(t, nullable, data, metas, nested=None):
    """the oracle's batch read of a WHOLE column body (page loop in C++): every page is compared, at BASELINE size"""
    return sbo.read_column_body(sbo.make_leaf(t, nullable, nested), data, metas)


def timed_decode(ctx, cols, reps=5):
    import torch
    dev = []
    for c in cols:
        td = torch.frombuffer(bytearray(c._keep.tobytes()), dtype=torch.uint8).cuda()
        dev.append(sb.Column(c.type, bool(c.leaf.nullable), td, c.metas,
                             [(c.leaf.nested_kind[i], c.leaf.nested_nullable[i]) for i in range(c.leaf.n_nested)] or None))
    best = None
    for _ in range(reps):
        out = ctx.decode_columns(dev, out="device")
        st = ctx.last_stats()
        out[0]._group.release()
        best = st if best is None or st["device_ms"] < best["device_ms"] else best
    return best


def record(name, st, extra=None):
    r = {"pages": st["pages"], "bytes_in": st["bytes_in"], "bytes_out": st["bytes_out"], "device_us": round(st["device_ms"] * 1e3, 1),
         "decoded_gbs": round(st["bytes_out"] / st["device_ms"] / 1e6, 1),
         "algorithmic_gbs": round((st["bytes_in"] + st["bytes_out"]) / st["device_ms"] / 1e6, 1),
         "codec_pages": st["codec_pages"], "kernel_launches": st["kernel_launches"]}
    r.update(extra or {})
    RESULTS[name] = r
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(RESULTS, open(os.path.join(ROOT, "gpurun_out", "perf_configs.json"), "w"), indent=1)


def test_config3_strings(ctx):
    """nullable Utf8 (i32 offsets) + LargeBinary (i64 offsets), uniq = 1000, 40 % nulls, adaptive on:
    pages written by the GPU encoder (Dict with a nested index block), read by the oracle and by us"""
    rng = np.random.default_rng(42)
    n = int(os.environ.get("SB_PERF_ROWS", 10_000_000))  # BASELINE.json configs[2] size
    cols, total_in = [], 0
    for name, t, large in (("utf8", sb.BINARY, False), ("large_binary", sb.LARGE_BINARY, True)):
        values, validity = dict_strings(rng, n, 1000, 0.4, large)
        enc = ctx.encode_columns([sb.LeafArray(t, values, validity=validity)], sb.write_options(sb.C_LZ4, 2.0, PAGE, seed=42))[0]
        col = sb.Column(t, True, enc.data, enc.metas)
        # parity on EVERY page: the oracle reads what our encoder wrote, we read it too, both equal the input
        ref = oracle_column(t, True, enc.data, enc.metas)
        dec = ctx.batch_read_array(col)
        assert_same(dec, ref, t, True)
        assert np.array_equal(sbo.unpack_bits(dec.validity, n), validity)
        lens = np.diff(dec.offsets.astype(np.int64))
        assert np.array_equal(lens[validity], np.diff(values[0].astype(np.int64))[validity])
        del ref, dec
        cols.append(col)
        st = timed_decode(ctx, [col])
        record(f"config3 {name} {n} rows", st)
    st = timed_decode(ctx, cols)
    record("config3 both columns, one call", st)


def test_config4_nested(ctx):
    """three leaves of List<Struct<a:Int64, b:Float64, c:Utf8>>, each page with its own rep/def streams"""
    nested = [(sbo.N_LIST, True), (sbo.N_STRUCT, True), (sbo.N_PRIMITIVE, True)]
    rng = np.random.default_rng(7)
    rows = int(os.environ.get("SB_PERF_NESTED_ROWS", 4_000_000))  # BASELINE.json configs[3] size
    rep, de, row_start = config4_levels(rng, rows)
    cols = []
    for name, t in (("a_i64", sbo.I64), ("b_f64", sbo.F64), ("c_utf8", sbo.BINARY)):
        pages, metas = [], []
        for r0 in range(0, rows, PAGE):
            e0, e1 = row_start[r0], (row_start[r0 + PAGE] if r0 + PAGE < rows else len(rep))
            prep, pde = rep[e0:e1], de[e0:e1]
            slots = pde >= 2
            valid = pde[slots] == 4
            ns = int(slots.sum())
            if t == sbo.BINARY:
                lens = np.where(valid, rng.integers(1, 6, ns), 0)
                off = np.zeros(ns + 1, np.int32)
                np.cumsum(lens, out=off[1:])
                vals = (off, rng.integers(97, 123, int(off[-1])).astype(np.uint8))
            elif t == sbo.F64:
                vals = rng.integers(0, 1000, ns).astype(np.float64)
            else:
                vals = rng.integers(0, 1 << 40, ns).astype(np.int64)
            block = sbo.compress_values(t, vals, validity=valid, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
            rep_b, def_b = sbo.levels_encode(prep, 1), sbo.levels_encode(pde, 3)
            hdr = np.array([min(PAGE, rows - r0), len(rep_b), len(def_b)], dtype="<u4").tobytes()
            pages.append(hdr + rep_b + def_b + block)
            metas.append((len(pages[-1]), int(e1 - e0)))
        data = b"".join(pages)
        col = sb.Column(t, True, data, metas, nested)
        ref = oracle_column(t, True, data, metas, nested)  # every page of the leaf
        dec = ctx.batch_read_array(col)
        assert_same_nested(dec, ref, t, nested)
        del ref, dec
        cols.append(col)
        record(f"config4 leaf {name} {rows} rows", timed_decode(ctx, [col]))
    record("config4 three leaves, one call", timed_decode(ctx, cols))


def test_config4_nested_written_on_gpu(ctx):
    """the same three leaves, pages written by sb_encode_columns (level streams + leaf value blocks):
    encode device time, then decode of those pages; the oracle reads the first pages to the same arrays"""
    nested = [(sbo.N_LIST, True), (sbo.N_STRUCT, True), (sbo.N_PRIMITIVE, True)]
    rng = np.random.default_rng(7)
    rows = int(os.environ.get("SB_PERF_NESTED_ROWS", 4_000_000))
    rep, de, row_start = config4_levels(rng, rows)
    slots = de >= 2
    valid = de[slots] == 4
    ns = int(slots.sum())
    cols = []
    for name, t in (("a_i64", sbo.I64), ("b_f64", sbo.F64), ("c_utf8", sbo.BINARY)):
        if t == sbo.BINARY:
            lens = np.where(valid, rng.integers(1, 6, ns), 0)
            off = np.zeros(ns + 1, np.int32)
            np.cumsum(lens, out=off[1:])
            vals = (off, rng.integers(97, 123, int(off[-1])).astype(np.uint8))
        elif t == sbo.F64:
            vals = rng.integers(0, 1000, ns).astype(np.float64)
        else:
            vals = rng.integers(0, 1 << 40, ns).astype(np.int64)
        arr = sb.LeafArray(t, vals, validity=valid, nullable=True, nested=nested, rep_levels=rep, def_levels=de, rows=rows)
        best = None
        for _ in range(3):
            enc = ctx.encode_columns([arr], sb.write_options(sb.C_LZ4, 2.0, PAGE, seed=42))[0]
            st = ctx.last_stats()
            best = st if best is None or st["device_ms"] < best["device_ms"] else best
        assert [m[1] for m in enc.metas[:-1]] == [int(row_start[r0 + PAGE] - row_start[r0]) for r0 in range(0, rows - PAGE, PAGE)]
        ref = oracle_column(t, True, enc.data, enc.metas, nested)  # every page: oracle and GPU readers agree ...
        dec = ctx.batch_read_array(sb.Column(t, True, enc.data, enc.metas, nested))
        assert_same_nested(dec, ref, t, nested)
        assert ref["length"] == ns
        if t != sbo.BINARY:  # ... and return the input
            assert np.array_equal(dec.values[valid], vals[valid])
        del ref, dec
        RESULTS[f"config4 encode leaf {name} {rows} rows (GPU writer)"] = {
            "pages": best["pages"], "bytes_in": best["bytes_in"], "bytes_out": best["bytes_out"], "device_us": round(best["device_ms"] * 1e3, 1),
            "encode_gbs": round(best["bytes_in"] / best["device_ms"] / 1e6, 1), "kernel_launches": best["kernel_launches"]}
        col = sb.Column(t, True, enc.data, enc.metas, nested)
        cols.append(col)
        record(f"config4 leaf {name} {rows} rows (GPU-written pages)", timed_decode(ctx, [col]))
    record("config4 three leaves, one call (GPU-written pages)", timed_decode(ctx, cols))
