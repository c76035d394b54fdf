"""CPU: pin the oracle against every known-answer vector available for this path:
upstream's only byte-level KAT (patas pack/unpack, src/compression/double/patas.rs:191-202)
and the hand-derived vectors K1..K8 of SURVEY.md Appendix A.6."""
import numpy as np
import sbo


def hx(s):
    return bytes.fromhex(s.replace("|", " "))


def test_patas_pack_unpack_upstream_kat():
    # patas.rs:193: (692,(1,2,52)), (1026,(2,8,2))
    assert sbo.patas_pack(1, 2, 52) == 692
    assert sbo.patas_pack(2, 8, 2) == 1026
    assert sbo.patas_unpack(692) == (1, 2, 52)
    assert sbo.patas_unpack(1026) == (2, 8, 2)


def test_K1_int64_plain():
    p = sbo.write_page(sbo.I64, np.arange(1, 7, dtype=np.int64))
    assert len(p) == 57
    assert p[:9] == hx("00 30 00 00 00 30 00 00 00")
    assert p[9:] == np.arange(1, 7, dtype="<i8").tobytes()


def test_K2_nullable_int32():
    p = sbo.write_page(sbo.I32, np.array([1, 0, 3], np.int32), validity=np.array([1, 0, 1], bool))
    assert p == hx("02 00 00 00 | 03 05 | 00 | 0C 00 00 00 | 0C 00 00 00 | 01000000 00000000 03000000")


def test_K3_rle():
    v = np.array([20] * 2045 + [10000] * 3, np.uint32)
    p = sbo.write_page(sbo.U32, v, opts=sbo.make_opts(force=sbo.C_RLE))
    assert p == hx("0A | 10 00 00 00 | 00 20 00 00 | FD 07 00 00 14 00 00 00 | 03 00 00 00 10 27 00 00")


def test_K4_onevalue():
    p = sbo.write_page(sbo.U32, np.full(2048, 3, np.uint32), opts=sbo.make_opts(ratio=2.0))
    assert p == hx("0C | 04 00 00 00 | 00 20 00 00 | 03 00 00 00")


def test_K5_dict():
    p = sbo.write_page(sbo.I64, np.array([7, 7, 9, 7], np.int64), opts=sbo.make_opts(force=sbo.C_DICT))
    assert p == hx("0B | 2D 00 00 00 | 20 00 00 00 | 00 | 10 00 00 00 | 10 00 00 00 |"
                   "00000000 00000000 01000000 00000000 | 02 00 00 00 | 07 00 00 00 00 00 00 00 | 09 00 00 00 00 00 00 00")


def test_K7_validity_header():
    p = sbo.write_page(sbo.I64, np.zeros(8192, np.int64), validity=np.ones(8192, bool))
    assert p[:6] == hx("02 04 00 00 81 10") and p[6:6 + 1024] == b"\xff" * 1024


def test_K8_boolean():
    p = sbo.write_page(sbo.BOOL, np.array([1, 1, 1, 0, 0, 0], bool))
    assert p == hx("00 | 01 00 00 00 | 06 00 00 00 | 07")


def test_stat_codec_tree_like_upstream_unit_test():
    """stat.rs:228-269: 20480 x "a" chooses OneValue; forced Dict -> indices OneValue, k = 1;
    forced Freq -> no exceptions."""
    n = 20480
    offsets = np.arange(n + 1, dtype=np.int32)
    data = np.full(n, ord("a"), np.uint8)
    p = sbo.write_page(sbo.BINARY, (offsets, data), opts=sbo.make_opts(sbo.C_LZ4, ratio=1.2))
    assert sbo.stat_page(sbo.BINARY, False, p) == "OneValue"
    p = sbo.write_page(sbo.BINARY, (offsets, data), opts=sbo.make_opts(sbo.C_LZ4, ratio=1.2, force=sbo.C_DICT))
    assert sbo.stat_page(sbo.BINARY, False, p) == "Dict(OneValue)[k=1]"
    p = sbo.write_page(sbo.BINARY, (offsets, data), opts=sbo.make_opts(sbo.C_LZ4, ratio=1.2, force=sbo.C_FREQ))
    assert sbo.stat_page(sbo.BINARY, False, p).startswith("Freq")
    r = sbo.read_column((sbo.BINARY, False), [(p, n)])
    assert np.array_equal(r["offsets"], offsets) and np.array_equal(r["values"], data)
