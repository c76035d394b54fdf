"""Nested columns end to end, with pyarrow as the independent Arrow implementation on both ends:

  pyarrow List<Struct<a:Int64,b:Float64,c:Utf8>> array
    -> its offsets / validity buffers -> sb_nested_levels (device)      == the levels pyarrow's Parquet writer produced
    -> sb_encode_columns (three leaves, pages cut by rows)              (write_nested, src/write/serialize.rs:135-198)
    -> sb_decode_columns -> sb_export_arrow (create_list / create_struct over the first leaf's NestedState)
    -> pyarrow array imported through the C Data Interface              == the input array

plus flat columns, List<Int32>, LargeList<List<Utf8>> and Struct<..> with list children."""
import os
import sys

import numpy as np
import pytest
import sbo

import strawboat_b200 as sb

pytestmark = pytest.mark.gpu
pa = pytest.importorskip("pyarrow")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

PHYS = {pa.int64(): sb.I64, pa.float64(): sb.F64, pa.int32(): sb.I32, pa.string(): sb.BINARY, pa.large_string(): sb.LARGE_BINARY,
        pa.bool_(): sb.BOOL}


def flatten(arr, path=()):
    """pyarrow array -> [(nested path root -> leaf, leaf pyarrow array)], the paths as sb_nested_levels wants them"""
    t = arr.type
    valid = np.asarray(arr.is_valid()) if arr.null_count else None
    if pa.types.is_list(t) or pa.types.is_large_list(t):
        off = np.asarray(arr.offsets).astype(np.int64 if pa.types.is_large_list(t) else np.int32)
        assert off[0] == 0
        if valid is not None:  # compact: null lists own no children
            assert np.all(np.diff(off)[~valid] == 0)
        lvl = {"kind": sb.N_LIST, "nullable": True, "length": len(arr), "offsets": off, "validity": valid}
        return flatten(arr.values[:off[-1]] if len(arr.values) != off[-1] else arr.values, path + (lvl,))
    if pa.types.is_struct(t):
        lvl = {"kind": sb.N_STRUCT, "nullable": True, "length": len(arr), "validity": valid}
        out = []
        for i in range(t.num_fields):
            out += flatten(arr.field(i), path + (lvl,))
        return out
    lvl = {"kind": sb.N_PRIMITIVE, "nullable": True, "length": len(arr), "validity": valid}
    return [(path + (lvl,), arr)]


def field_of(t, name="col"):
    if pa.types.is_list(t) or pa.types.is_large_list(t):
        return sb.make_field(sb.N_LIST, nullable=True, children=[field_of(t.value_type, "item")], name=name, large=pa.types.is_large_list(t))
    if pa.types.is_struct(t):
        return sb.make_field(sb.N_STRUCT, nullable=True, children=[field_of(t.field(i).type, t.field(i).name) for i in range(t.num_fields)], name=name)
    return sb.make_field(sb.N_PRIMITIVE, PHYS[t], nullable=True, name=name, utf8=pa.types.is_string(t) or pa.types.is_large_string(t))


def leaf_array(torch, leaf, nested, rep, de, rows):
    t = PHYS[leaf.type]
    valid = np.asarray(leaf.is_valid())
    vbits = torch.from_numpy(np.packbits(valid, bitorder="little")).cuda()
    if t in (sb.BINARY, sb.LARGE_BINARY):
        odt = np.int64 if t == sb.LARGE_BINARY else np.int32
        strs = [(x.as_py() or "").encode() for x in leaf]
        off = np.concatenate([[0], np.cumsum([len(x) for x in strs])]).astype(odt)
        dat = np.frombuffer(b"".join(strs), np.uint8).copy() if off[-1] else np.zeros(0, np.uint8)
        vals = (torch.from_numpy(off).cuda(), torch.from_numpy(dat).cuda())
    elif t == sb.BOOL:
        bits = np.asarray(leaf.fill_null(False))
        return sb.LeafArray(t, torch.from_numpy(np.packbits(bits, bitorder="little")).cuda(), validity=vbits, nullable=True, length=len(leaf),
                            nested=nested, rep_levels=rep, def_levels=de, rows=rows)
    else:
        vals = torch.from_numpy(np.asarray(leaf.fill_null(0)).astype(sb.NP_OF[t])).cuda()
    return sb.LeafArray(t, vals, validity=vbits, nullable=True, nested=nested, rep_levels=rep, def_levels=de, rows=rows)


def roundtrip(ctx, arr, page_rows=64, default=sb.C_LZ4, ratio=2.0):
    import torch
    leaves = flatten(arr)
    cols, keep = [], []
    for path, leaf in leaves:
        nested = [(d["kind"], d["nullable"]) for d in path]
        if len(path) == 1:  # flat column
            la = leaf_array(torch, leaf, None, None, None, None)
            la.nested = None
            enc = ctx.encode_columns([la], sb.write_options(default, ratio, page_rows, seed=3))[0]
            cols.append(sb.Column(PHYS[leaf.type], True, enc.data, enc.metas))
            continue
        rep, de, n_slots = ctx.nested_levels(list(path))
        assert n_slots == len(leaf)
        keep += [rep, de]
        enc = ctx.encode_columns([leaf_array(torch, leaf, nested, rep, de, len(arr))], sb.write_options(default, ratio, page_rows, seed=3))[0]
        assert sum(m[1] for m in enc.metas) == rep.numel()
        cols.append(sb.Column(PHYS[leaf.type], True, enc.data, enc.metas, nested))
    out = ctx.read_arrow(cols, field_of(arr.type))
    out.validate(full=True)
    assert out.type == arr.type, (out.type, arr.type)
    assert out.equals(arr), (out[:5], arr[:5])
    return out


def test_levels_match_pyarrow_parquet_writer(ctx):
    """sb_nested_levels over the golden array == the (rep, def) pairs in the pyarrow-written Parquet page"""
    import torch
    from make_dremel_golden import build
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dremel_pyarrow.npz"))
    arr = build()
    for (path, leaf), name in zip(flatten(arr), "abc"):
        rep, de, n_slots = ctx.nested_levels(list(path))
        nv = int(G[name + "_num_values"])
        assert rep.numel() == nv and n_slots == len(G["struct_validity"])
        got_rep = torch.as_tensor(sb._DevArray(rep.data_ptr(), 4 * nv, rep), device="cuda").view(torch.int32).cpu().numpy()
        got_def = torch.as_tensor(sb._DevArray(de.data_ptr(), 4 * nv, de), device="cuda").view(torch.int32).cpu().numpy()
        assert np.array_equal(got_rep, sbo.hybrid_rle_decode(G[name + "_rep"].tobytes(), 1, nv))
        assert np.array_equal(got_def, sbo.hybrid_rle_decode(G[name + "_def"].tobytes(), 3, nv))


def test_config4_schema_roundtrip_through_arrow(ctx):
    from make_dremel_golden import build
    arr = build(rows=900, seed=5)
    for page_rows in (64, 1000, None):
        roundtrip(ctx, arr, page_rows)
    roundtrip(ctx, arr, 128, sb.C_NONE, None)
    roundtrip(ctx, arr, 128, sb.C_SNAPPY, 2.0)


def test_other_shapes(ctx):
    rng = np.random.default_rng(3)
    roundtrip(ctx, pa.array([None if rng.random() < 0.2 else int(x) for x in rng.integers(0, 50, 700)], type=pa.int64()))
    roundtrip(ctx, pa.array([None if rng.random() < 0.2 else "s%d" % x for x in rng.integers(0, 50, 700)], type=pa.string()))
    lists = [None if rng.random() < 0.1 else [None if rng.random() < 0.2 else int(x) for x in rng.integers(0, 9, rng.integers(0, 5))] for _ in range(500)]
    roundtrip(ctx, pa.array(lists, type=pa.list_(pa.int32())))
    ll = [None if rng.random() < 0.1 else [None if rng.random() < 0.15 else ["w%d" % x for x in rng.integers(0, 20, rng.integers(0, 4))]
                                             for _ in range(rng.integers(0, 4))] for _ in range(300)]
    roundtrip(ctx, pa.array(ll, type=pa.large_list(pa.list_(pa.string()))))
    st = pa.struct([pa.field("x", pa.int64()), pa.field("tags", pa.list_(pa.string())), pa.field("ok", pa.bool_())])
    rows = [None if rng.random() < 0.1 else {"x": None if rng.random() < 0.2 else int(rng.integers(0, 99)),
                                             "tags": None if rng.random() < 0.2 else ["t%d" % x for x in rng.integers(0, 6, rng.integers(0, 3))],
                                             "ok": None if rng.random() < 0.2 else bool(rng.random() < 0.5)} for _ in range(400)]
    roundtrip(ctx, pa.array(rows, type=st))
    roundtrip(ctx, pa.array([], type=pa.list_(pa.int32())))
