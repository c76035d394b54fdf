"""Zstandard as the common codec (CommonCompression::Zstd, src/compression/basic.rs:93-97,122-136; the reference's
test matrix runs it: tests/it/io.rs:420-425).

CPU: the DEVICE decoder (strawboat_b200/csrc/sb_zstd.cuh) is compiled for the host with one emulated lane and
checked against frames written by libzstd (the library the `zstd` crate wraps) at several levels, and by pyarrow.
GPU: pages written by the oracle with default_compression = Zstd decode to the oracle's arrays."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("zstd") / "zstd_harness.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(so), os.path.join(ROOT, "tests", "zstd_host_harness.cpp")])
    lib = C.CDLL(str(so))
    lib.zstd_harness_decode.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]

    def decode(comp, n):
        out = np.zeros(max(1, n), np.uint8)
        rc = lib.zstd_harness_decode(comp, len(comp), out.ctypes.data, n)
        return rc, out[:n].tobytes()
    return decode


def corpus(rng):
    yield b"a"
    yield b"abcd" * 1000
    yield bytes(rng.integers(0, 256, 5000, dtype=np.uint8))                              # raw block
    yield b"\0" * 100000                                                                  # RLE literals / long matches
    yield np.repeat(rng.integers(0, 256, 40, dtype=np.uint8), rng.integers(1, 400, 40)).tobytes()
    yield rng.integers(0, 65536, 8192).astype(np.float64).tobytes()                      # 64 KiB page: Huffman literals, 4 streams
    yield np.cumsum(rng.integers(0, 4, 8192)).astype(np.int32).tobytes()
    yield rng.integers(0, 8, 8192).astype(np.int32).tobytes()
    yield rng.integers(0, 1000, 40000).astype(np.int64).tobytes()                         # > 128 KiB: several blocks, repeat modes
    yield ("".join("row %d of the text column, value=%d;" % (i, i * 7919 % 1000) for i in range(3000))).encode()
    yield bytes(rng.integers(97, 101, 300, dtype=np.uint8))                                # small: 1-stream Huffman / direct weights


def test_device_decoder_on_host_vs_libzstd(harness):
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(0)
    for data in corpus(rng):
        frames = [sbo.common_compress(sbo.C_ZSTD, data)]                                   # libzstd level 0 (= 3), as basic.rs:129
        frames += [pa.Codec("zstd", compression_level=lvl).compress(data).to_pybytes() for lvl in (1, 6, 19)]
        for comp in frames:
            rc, out = harness(comp, len(data))
            assert rc == 0 and out == data, (len(data), len(comp))
        comp = frames[0]
        assert harness(comp, len(data) + 1)[0] != 0                                        # frame shorter than the rows
        if len(data) > 1:
            assert harness(comp, len(data) - 1)[0] != 0                                    # dstSize_tooSmall
        assert harness(comp[:-1], len(data))[0] != 0                                       # truncated
    assert harness(b"\x28\xb5\x2f\xfc" + b"\0" * 10, 4)[0] != 0                           # bad magic


def test_device_decoder_survives_corruption(harness):
    """flipped bytes must fail or decode to something -- never crash or run away (the harness is plain C++: a wild
    read would fault here)"""
    rng = np.random.default_rng(1)
    data = rng.integers(0, 65536, 4096).astype(np.float64).tobytes()
    comp = bytearray(sbo.common_compress(sbo.C_ZSTD, data))
    for _ in range(300):
        bad = bytearray(comp)
        for _ in range(int(rng.integers(1, 4))):
            bad[int(rng.integers(4, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        harness(bytes(bad), len(data))


def columns(rng, n):
    from strawboat_b200.workloads import random_strings
    yield sbo.I32, np.cumsum(rng.integers(0, 4, n)).astype(np.int32), None
    yield sbo.I64, rng.integers(0, 50, n), rng.random(n) > 0.2
    yield sbo.F64, rng.integers(0, 65536, n).astype(np.float64), None
    yield sbo.U8, rng.integers(0, 256, n).astype(np.uint8), None
    yield sbo.I64, np.full(n, 7), None
    yield sbo.BOOL, rng.random(n) < 0.3, rng.random(n) > 0.1
    o, d, v = random_strings(rng, n, 40, 0.2)
    yield sbo.BINARY, (o, d), v
    o, d, v = random_strings(rng, n, 5000, 0.0, large=True)
    yield sbo.LARGE_BINARY, (o, d), None


@pytest.mark.gpu
@pytest.mark.parametrize("ratio", [None, 2.0])
def test_gpu_reads_oracle_zstd_pages(ctx, ratio):
    import strawboat_b200 as sb
    rng = np.random.default_rng(1)
    for n in (1, 100, 5000, 20000):
        for t, v, val in columns(rng, n):
            data, metas = oracle_encode_column(t, v, val, page_size=8192, opts=sbo.make_opts(sbo.C_ZSTD, ratio=ratio))
            ref = oracle_decode_column(t, val is not None, data, metas)
            dec = ctx.batch_read_array(sb.Column(t, val is not None, data, metas))
            assert_same(dec, ref, t, val is not None)


@pytest.mark.gpu
def test_gpu_zstd_corrupt_frame(ctx):
    import strawboat_b200 as sb
    rng = np.random.default_rng(2)
    v = rng.integers(0, 1000, 4096).astype(np.int64)
    data, metas = oracle_encode_column(sbo.I64, v, page_size=2048, opts=sbo.make_opts(sbo.C_ZSTD))
    bad = bytearray(data)
    bad[9] ^= 0xff  # magic
    res = ctx.decode_columns([sb.Column(sb.I64, False, bytes(bad), metas)], raise_on_page_error=False)[0]
    assert res.page_status[0] == sb._capi.SB_EXTERNAL and res.page_status[1] == 0
    assert np.array_equal(res.values[2048:], v[2048:])


@pytest.mark.gpu
def test_oracle_reads_gpu_zstd_pages(ctx):
    """default_compression = Zstd on the GPU writer: valid frames (stored blocks) that libzstd reads"""
    import strawboat_b200 as sb
    rng = np.random.default_rng(3)
    for n in (1, 100, 5000, 40000):
        for t, v, val in columns(rng, n):
            for ratio in (None, 2.0):
                enc = ctx.encode_columns([sb.LeafArray(t, v, validity=val)], sb.write_options(sb.C_ZSTD, ratio, 8192 if n < 40000 else None, seed=1))[0]
                ref = oracle_decode_column(t, val is not None, enc.data, enc.metas)
                dec = ctx.batch_read_array(sb.Column(t, val is not None, enc.data, enc.metas))
                assert_same(dec, ref, t, val is not None)
                if t in sbo.NP_OF and val is None:
                    assert np.array_equal(ref["values"].view(np.uint8), np.ascontiguousarray(v, dtype=sbo.NP_OF[t]).view(np.uint8))
