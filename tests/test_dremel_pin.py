"""Pins the Dremel semantics of nested pages (read_validity_nested, src/read/read_basic.rs:65-173;
oracle/FORMAT_ASSUMPTIONS.md #6) against level streams written by PYARROW's Parquet writer for config 4's
schema List<Struct<a:Int64, b:Float64, c:Utf8>> (tests/golden/make_dremel_golden.py made the fixture):
null list / empty list / null struct / null leaf -> (rep, def), entries per row, leaf slots, and the
NestedState the reader rebuilds must equal the Arrow array pyarrow shredded."""
import os

import numpy as np
import pytest
import sbo

NESTED = [(sbo.N_LIST, True), (sbo.N_STRUCT, True), (sbo.N_PRIMITIVE, True)]
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dremel_pyarrow.npz"))
LEAVES = (("a", sbo.I64), ("b", sbo.F64), ("c", sbo.BINARY))


def leaf_values(name, type_):
    if type_ == sbo.BINARY:
        return (G["c_offsets"], G["c_data"])
    return G[name + "_values"]


def make_page(name, type_, opts=None):
    """strawboat nested page around pyarrow's level streams: [rows][rep_len][def_len][rep][def][VALUE_BLOCK]"""
    rep, de = G[name + "_rep"].tobytes(), G[name + "_def"].tobytes()
    block = sbo.compress_values(type_, leaf_values(name, type_), validity=G[name + "_validity"], opts=opts or sbo.make_opts())
    hdr = np.array([int(G["rows"]), len(rep), len(de)], dtype="<u4").tobytes()
    return hdr + rep + de + block, int(G[name + "_num_values"])


def check(res, name, type_, unpack):
    n_slots = len(G["struct_validity"])
    assert res["length"] == n_slots
    lst, stc = res["nested"][0], res["nested"][1]
    assert np.array_equal(np.asarray(lst["offsets"])[:len(G["list_offsets"]) - 1], G["list_offsets"][:-1])
    assert np.array_equal(unpack(lst["validity"], int(G["rows"])), G["list_validity"])
    assert np.array_equal(unpack(stc["validity"], n_slots), G["struct_validity"])
    valid = G[name + "_validity"]
    assert np.array_equal(unpack(res["validity"], n_slots), valid)
    if type_ == sbo.BINARY:
        # valid slots carry their strings; what a null slot holds depends on the codec (Dict repeats a neighbour)
        go, eo = np.asarray(res["offsets"], np.int64), G["c_offsets"].astype(np.int64)
        assert len(go) == n_slots + 1 and go[0] == 0 and np.all(np.diff(go) >= 0) and go[-1] == len(res["values"])
        for i in np.flatnonzero(valid):
            assert bytes(res["values"][go[i]:go[i + 1]]) == bytes(G["c_data"][eo[i]:eo[i + 1]])
    else:
        assert np.array_equal(np.asarray(res["values"])[valid], G[name + "_values"][valid])


def test_fixture_covers_every_shape():
    """level alphabet of the fixture: def 0 null list, 1 empty list, 2 null struct, 3 null leaf, 4 value"""
    de = sbo.hybrid_rle_decode(G["a_def"].tobytes(), 3, int(G["a_num_values"]))
    rep = sbo.hybrid_rle_decode(G["a_rep"].tobytes(), 1, int(G["a_num_values"]))
    assert set(de.tolist()) == {0, 1, 2, 3, 4} and set(rep.tolist()) == {0, 1}
    assert int((rep == 0).sum()) == int(G["rows"])                      # one rep == 0 entry starts every row
    assert int((de >= 2).sum()) == len(G["struct_validity"])            # leaf slots: also under a null struct
    assert np.array_equal(de[de >= 2] >= 3, G["struct_validity"])
    assert np.array_equal(de[de >= 2] == 4, G["a_validity"])
    # null and empty lists own exactly one entry and no slot
    lens = np.diff(G["list_offsets"])
    per_row = np.add.reduceat(np.ones(len(rep), np.int64), np.flatnonzero(rep == 0))
    assert np.array_equal(per_row, np.maximum(lens, 1))
    first = de[np.flatnonzero(rep == 0)]
    assert np.array_equal(first == 0, ~G["list_validity"])
    assert np.array_equal(first == 1, G["list_validity"] & (lens == 0))


@pytest.mark.parametrize("name,type_", LEAVES)
def test_oracle_rebuilds_the_arrow_structure(name, type_):
    page, nv = make_page(name, type_)
    res = sbo.read_column(sbo.make_leaf(type_, True, NESTED), [(page, nv)])
    check(res, name, type_, sbo.unpack_bits)


@pytest.mark.gpu
@pytest.mark.parametrize("name,type_", LEAVES)
@pytest.mark.parametrize("default", [sbo.C_NONE, sbo.C_LZ4])
def test_gpu_rebuilds_the_arrow_structure(ctx, name, type_, default):
    import strawboat_b200 as sb
    page, nv = make_page(name, type_, sbo.make_opts(default, ratio=2.0))
    dec = ctx.batch_read_array(sb.Column(type_, True, page, [(len(page), nv)], NESTED))
    res = {"length": dec.length, "validity": dec.validity, "values": dec.values, "offsets": dec.offsets,
           "nested": [{"offsets": d["offsets"], "validity": d["validity"]} for d in dec.nested]}
    check(res, name, type_, sbo.unpack_bits)
    assert dec.nested[0]["offsets"][-1] == len(G["struct_validity"])  # create_list's final offset
