"""CPU: the oracle writer and reader round-trip every case of the reference's integration tests
(tests/it/io.rs) -- the host-side analogue of `assert_eq!(chunk, result)` (io.rs:496,527)."""
import numpy as np
import pytest
import sbo
from helpers import oracle_decode_column, oracle_encode_column

from strawboat_b200.workloads import random_strings

CODECS = [sbo.C_NONE, sbo.C_LZ4, sbo.C_ZSTD]
INT_TYPES = [sbo.I8, sbo.I16, sbo.I32, sbo.I64, sbo.U8, sbo.U16, sbo.U32, sbo.U64]


def rt(type_, values, validity=None, page_size=2048, **kw):
    for default in CODECS:
        for ratio in (None, 2.0):
            opts = sbo.make_opts(default, ratio=ratio, **kw)
            data, metas = oracle_encode_column(type_, values, validity, page_size=page_size, opts=opts)
            ref = oracle_decode_column(type_, validity is not None, data, metas)
            n = ref["length"]
            m = np.ones(n, bool) if validity is None else np.asarray(validity, bool)
            if type_ == sbo.BOOL:
                assert np.array_equal(sbo.unpack_bits(ref["values"], n)[m], np.asarray(values, bool)[m])
            elif type_ in (sbo.BINARY, sbo.LARGE_BINARY):
                off = np.asarray(values[0], np.int64)
                assert np.array_equal(np.diff(ref["offsets"].astype(np.int64))[m], np.diff(off)[m])
            else:
                assert np.array_equal(ref["values"][m], np.asarray(values)[m])
            if validity is not None:
                assert np.array_equal(sbo.unpack_bits(ref["validity"], n), m)


@pytest.mark.parametrize("type_", INT_TYPES + [sbo.F32, sbo.F64])
def test_primitives(type_):
    rng = np.random.default_rng(type_)
    dt = sbo.NP_OF[type_]
    n = 5000
    for v in (rng.integers(0, 100, n).astype(dt), np.full(n, 3, dt), np.repeat(rng.integers(0, 50, n // 50), 50).astype(dt),
              np.where(rng.random(n) < 0.95, 20, rng.integers(0, 100, n)).astype(dt)):
        rt(type_, v)
        rt(type_, v, validity=rng.random(n) > 0.3)
    for force in (sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ):  # the CI matrix (.github/workflows/rust.yml:23-25)
        rt(type_, rng.integers(0, 100, n).astype(dt), validity=rng.random(n) > 0.1, force=force)


def test_bitpacking_and_delta():
    rng = np.random.default_rng(1)
    n = 10240
    for t in (sbo.U32, sbo.I32):
        rt(t, rng.integers(0, 1000, n).astype(sbo.NP_OF[t]))
        rt(t, np.arange(n).astype(sbo.NP_OF[t]))  # test_deletabitpacking (io.rs:144-152)


def test_boolean_and_binary():
    rng = np.random.default_rng(2)
    rt(sbo.BOOL, rng.random(5000) < 0.5)
    rt(sbo.BOOL, rng.random(5000) < 0.5, validity=rng.random(5000) > 0.2, page_size=1001)
    rt(sbo.BOOL, np.ones(5000, bool))
    for large in (False, True):
        o, d, v = random_strings(rng, 6000, 100, 0.3, large=large)
        rt(sbo.LARGE_BINARY if large else sbo.BINARY, (o, d), validity=v)
        for force in (sbo.C_DICT, sbo.C_FREQ):
            rt(sbo.LARGE_BINARY if large else sbo.BINARY, (o, d), validity=v, force=force)
