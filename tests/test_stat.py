"""Page inspector (sb_stat_page, the stat_simple / stat_body walk of src/stat.rs:63-152) against the
oracle's walker on pages the oracle wrote.  Host-only code: runs without a GPU."""
import numpy as np
import pytest
import sbo
from helpers import oracle_encode_column

import strawboat_b200 as sb


def pages_of(data, metas):
    pos = 0
    for ln, _ in metas:
        yield data[pos:pos + ln]
        pos += ln


CASES = [
    ("plain i64", sbo.I64, lambda r, n: r.integers(-2**62, 2**62, n), None, {}),
    ("dict i32", sbo.I32, lambda r, n: r.integers(0, 8, n).astype(np.int32), None, dict(ratio=2.0)),
    ("freq i64", sbo.I64, lambda r, n: np.where(r.random(n) < 0.95, 20, r.integers(10000, 20000, n)), None, dict(ratio=2.0)),
    ("rle i64", sbo.I64, lambda r, n: np.repeat(r.integers(0, 1000, n // 64 + 1), 64)[:n], None, dict(ratio=2.0)),
    ("onevalue nullable", sbo.I64, lambda r, n: np.full(n, 7), 0.3, dict(ratio=2.0)),
    ("sorted i32", sbo.I32, lambda r, n: np.cumsum(r.integers(0, 4, n)).astype(np.int32), None, dict(ratio=1.2)),
    ("f64 lowcard nullable", sbo.F64, lambda r, n: r.integers(0, 8, n).astype(np.float64), 0.2, dict(ratio=2.0)),
    ("f64 patas", sbo.F64, lambda r, n: r.integers(0, 65536, n).astype(np.float64), None, dict(ratio=1.2)),
    ("lz4 i32", sbo.I32, lambda r, n: r.integers(-2**31, 2**31 - 1, n).astype(np.int32), None, dict(default=sbo.C_LZ4, ratio=2.0)),
    ("bool", sbo.BOOL, lambda r, n: r.random(n) < 0.01, 0.1, dict(ratio=2.0)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_same_tree_as_the_oracle(case):
    _, t, gen, p_null, o = case
    rng = np.random.default_rng(3)
    n = 5000
    values = gen(rng, n)
    if t != sbo.BOOL:
        values = values.astype(sbo.NP_OF[t])
    validity = None if p_null is None else rng.random(n) >= p_null
    opts = sbo.make_opts(o.get("default", sbo.C_NONE), ratio=o.get("ratio"))
    data, metas = oracle_encode_column(t, values, validity, validity is not None, 2048, opts, seed=5)
    for page, (ln, nv) in zip(pages_of(data, metas), metas):
        tree, info = sb.stat_page(t, validity is not None, page)
        assert tree == sbo.stat_page(t, validity is not None, page)
        vb = 0 if validity is None else 4 + info["validity_size"]
        assert info["codec"] == page[vb] == info["path"][0]
        assert info["compressed_size"] == int.from_bytes(page[vb + 1:vb + 5], "little")
        assert info["uncompressed_size"] == int.from_bytes(page[vb + 5:vb + 9], "little")
        assert vb + 9 + info["compressed_size"] <= ln
        assert info["depth"] == tree.count("(") + 1


def test_binary_trees():
    rng = np.random.default_rng(4)
    n = 4000
    ids = rng.integers(0, 50, n)
    words = [str(i).encode() for i in range(50)]
    lens = np.array([len(words[i]) for i in ids])
    off = np.zeros(n + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    vals = (off, np.frombuffer(b"".join(words[i] for i in ids), np.uint8))
    for ratio in (None, 2.0):
        data, metas = oracle_encode_column(sbo.BINARY, vals, None, False, 1024, sbo.make_opts(sbo.C_NONE, ratio=ratio), seed=1)
        for page in pages_of(data, metas):
            tree, info = sb.stat_page(sb.BINARY, False, page)
            assert tree == sbo.stat_page(sbo.BINARY, False, page)
            if ratio:
                assert tree.startswith("Dict(") and info["unique_num"] == 50


def test_errors():
    with pytest.raises(sb.StrawboatError) as e:
        sb.stat_page(sb.I64, False, b"\x63" + b"\0" * 8)
    assert e.value.code == sb._capi.SB_OUT_OF_SPEC
    with pytest.raises(sb.StrawboatError) as e:
        sb.stat_page(sb.I64, False, b"\x00\x01")
    assert e.value.code == sb._capi.SB_IO
    with pytest.raises(sb.StrawboatError) as e:
        sb.stat_page(sb.I64, True, b"\xff\xff\xff\x7f" + b"\0" * 20)
    assert e.value.code == sb._capi.SB_IO
