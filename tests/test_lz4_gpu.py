"""GPU parity for LZ4 value blocks (CommonCompression::Lz4, basic.rs:87-91): the streaming
scanner/mover kernel against the oracle (which calls liblz4, the library the reference's lz4
crate wraps) on byte patterns that exercise every path of the decoder: short sequences, long
literal runs, long and self-overlapping matches, matches older than the shared-memory ring,
length-extension chains longer than the rings ("BIG" entries), stored blocks, corrupt
streams."""
import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb

pytestmark = pytest.mark.gpu

LZ4 = sbo.make_opts(sbo.C_LZ4)  # adaptive off: every value block is a raw LZ4 block


def check(ctx, type_, values, page_size, validity=None):
    data, metas = oracle_encode_column(type_, values, validity, validity is not None, page_size, LZ4)
    ref = oracle_decode_column(type_, validity is not None, data, metas)
    dec = ctx.batch_read_array(sb.Column(type_, validity is not None, data, metas))
    assert_same(dec, ref, type_, validity is not None)
    st = ctx.last_stats()
    assert st["codec_pages"].get(sb.C_LZ4, 0) == len(metas)


def as_u8_values(b, type_):
    dt = np.dtype(sbo.NP_OF[type_])
    b = np.frombuffer(bytes(b), dtype=np.uint8)
    b = b[: len(b) // dt.itemsize * dt.itemsize]
    return b.view(dt).copy()


def patterns(rng, nbytes):
    """byte strings with very different LZ4 sequence shapes"""
    out = {}
    out["random"] = rng.integers(0, 256, nbytes, dtype=np.uint8).tobytes()  # stored block
    out["zeros"] = bytes(nbytes)  # one huge match, offset 1
    for period in (2, 3, 5, 7, 13, 31, 32, 33, 100):
        pat = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
        out[f"period{period}"] = (pat * (nbytes // period + 1))[:nbytes]
    # numeric columns: short literals + short matches
    out["small_ints_f64"] = rng.integers(0, 65536, nbytes // 8).astype(np.float64).tobytes()
    out["sorted_i32"] = np.cumsum(rng.integers(0, 4, nbytes // 4)).astype(np.int32).tobytes()
    # a pool of 8-byte values: short matches at distances up to the whole page
    pool = rng.integers(0, 1 << 62, 3000)
    out["pool_far"] = pool[rng.integers(0, len(pool), nbytes // 8)].astype(np.int64).tobytes()
    # long literal runs between repetitive stretches (BIG literals when > 1035 bytes)
    parts = []
    while sum(map(len, parts)) < nbytes:
        parts.append(rng.integers(0, 256, int(rng.integers(1, 6000)), dtype=np.uint8).tobytes())
        parts.append(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 6000)))
    out["lit_match_mix"] = b"".join(parts)[:nbytes]
    # a block repeated far away: long matches with offsets beyond the ring
    blk = rng.integers(0, 256, 10000, dtype=np.uint8).tobytes()
    out["far_repeat"] = (blk * (nbytes // len(blk) + 1))[:nbytes]
    blk = rng.integers(0, 256, 40000, dtype=np.uint8).tobytes()
    out["very_far_repeat"] = (blk * (nbytes // len(blk) + 1))[:nbytes]
    # text-like
    words = [bytes(rng.integers(97, 123, int(rng.integers(2, 12)), dtype=np.uint8)) for _ in range(500)]
    out["text"] = b" ".join(words[int(i)] for i in rng.integers(0, 500, nbytes // 6))[:nbytes]
    return out


@pytest.mark.parametrize("type_", [sbo.U8, sbo.I32, sbo.I64])
def test_lz4_patterns_pages(ctx, type_):
    rng = np.random.default_rng(5)
    W = np.dtype(sbo.NP_OF[type_]).itemsize
    for name, b in patterns(rng, 200_000).items():
        v = as_u8_values(b, type_)
        for rows in (8192, 2048 // W + 3):
            check(ctx, type_, v, rows)


def test_lz4_single_large_page(ctx):
    """max_page_size = None: one block of several MiB; length chains far longer than the rings."""
    rng = np.random.default_rng(6)
    for name, b in patterns(rng, 3_000_000).items():
        check(ctx, sbo.I64, as_u8_values(b, sbo.I64), None)
    # a stored block too large for the stored-block shortcut and a huge literal run inside a
    # compressible block
    check(ctx, sbo.U8, as_u8_values(rng.integers(0, 256, 2_000_000, dtype=np.uint8).tobytes(), sbo.U8), None)
    mix = bytes(100_000) + rng.integers(0, 256, 1_500_000, dtype=np.uint8).tobytes() + bytes(300_000)
    check(ctx, sbo.U8, as_u8_values(mix, sbo.U8), None)


def test_lz4_nullable_and_ragged(ctx):
    rng = np.random.default_rng(7)
    n = 30_001
    v = rng.integers(0, 50, n).astype(np.int64)
    val = rng.random(n) > 0.3
    check(ctx, sbo.I64, v, 4096, validity=val)
    check(ctx, sbo.I64, v[:1500], 1, validity=val[:1500])  # one row per page
    check(ctx, sbo.I16, rng.integers(0, 5, 777).astype(np.int16), 100)


def test_lz4_many_columns_mixed(ctx):
    """many LZ4 pages of different shapes in one launch (job bins, tail of the queue)"""
    rng = np.random.default_rng(8)
    pats = patterns(rng, 400_000)
    cols, refs = [], []
    for name, b in pats.items():
        v = as_u8_values(b, sbo.I64)
        data, metas = oracle_encode_column(sbo.I64, v, None, False, 8192, LZ4)
        cols.append(sb.Column(sbo.I64, False, data, metas))
        refs.append(oracle_decode_column(sbo.I64, False, data, metas))
    decs = ctx.decode_columns(cols)
    for d, r in zip(decs, refs):
        assert_same(d, r, sbo.I64, False)


def test_lz4_corrupt_streams(ctx):
    """every corruption yields a per-page status (SB_EXTERNAL or a clean decode), never a hang,
    a crash or a write outside the page's output"""
    rng = np.random.default_rng(9)
    v = rng.integers(0, 65536, 8192 * 2).astype(np.float64)
    data, metas = oracle_encode_column(sbo.F64, v, None, False, 8192, LZ4)
    good = ctx.batch_read_array(sb.Column(sbo.F64, False, data, metas))
    L0 = metas[0][0]
    for trial in range(40):
        bad = bytearray(data)
        kind = trial % 4
        if kind == 0:  # flip bytes inside the first block
            for _ in range(3):
                bad[9 + int(rng.integers(0, L0 - 9))] = int(rng.integers(0, 256))
        elif kind == 1:  # zero offset
            pos = 9 + int(rng.integers(0, L0 - 20))
            bad[pos:pos + 8] = bytes(8)
        elif kind == 2:  # all 0xff: endless length chains
            pos = 9 + int(rng.integers(0, L0 - 600))
            bad[pos:pos + 500] = b"\xff" * 500
        else:  # wrong uncompressed size in the header
            bad[5:9] = int(rng.integers(1, 1 << 20)).to_bytes(4, "little")
        res = ctx.decode_columns([sb.Column(sbo.F64, False, bytes(bad), metas)], raise_on_page_error=False)[0]
        assert res.page_status[0] in (0, sb._capi.SB_EXTERNAL, sb._capi.SB_IO, sb._capi.SB_PANIC)
        assert res.page_status[1] == 0
        # the second page is untouched by whatever happened to the first
        assert np.array_equal(res.values[8192:].view(np.uint8), good.values[8192:].view(np.uint8))


def test_lz4_own_encoder_chains(ctx):
    """blocks written by this library's LZ4 encoder: its greedy matcher links every match to the
    previous value (offset == value width), the chain case of the mover (pointer jumping)."""
    rng = np.random.default_rng(10)
    wo = sb.write_options(sb.C_LZ4, None, 8192)
    for type_, v in ((sbo.F64, rng.integers(0, 65536, 50_000).astype(np.float64)),
                     (sbo.I64, rng.integers(0, 1 << 20, 50_000).astype(np.int64)),
                     (sbo.I32, np.cumsum(rng.integers(0, 4, 70_001)).astype(np.int32)),
                     (sbo.I16, rng.integers(0, 300, 33_333).astype(np.int16)),
                     (sbo.F32, rng.integers(0, 1000, 20_000).astype(np.float32))):
        enc = ctx.encode_columns([sb.LeafArray(type_, v)], wo)[0]
        ref = oracle_decode_column(type_, False, enc.data, enc.metas)  # liblz4 reads our blocks
        assert np.array_equal(ref["values"].view(np.uint8), v.view(np.uint8))
        dec = ctx.batch_read_array(sb.Column(type_, False, enc.data, enc.metas))
        assert_same(dec, ref, type_, False)


@pytest.mark.parametrize("codec", [sb.C_LZ4, sb.C_SNAPPY])
def test_own_matcher_patterns_and_chunk_junctions(ctx, codec):
    """the CTA-wide matcher of the writer (sb_encode.cuh lz_compress_cta: 4 chunks, parked first matches, junction
    sequences, bodies moved left): every byte pattern above, at page sizes around the chunking rules (one chunk below
    2048 bytes per warp, chunk = ceil(n / 4) rounded to 64, pages beyond 64 KiB), is read back to the input by the
    ORACLE (liblz4 / its Snappy reader) and by the GPU decoder.  A temporary offset below the final one would
    surface as SB_PANIC, an overlap of two chunk bodies as a wrong byte."""
    rng = np.random.default_rng(77)
    sizes = [1, 5, 12, 13, 64, 2047, 2048, 2049, 4096, 8191, 8192, 8193, 8256, 16385, 65535, 65536, 65537, 100003, 262144 + 5]
    pats = patterns(rng, max(sizes))
    # stretches that end or start exactly where a chunk does
    for n in (8192, 65536):
        cl = n // 4
        b = bytearray(rng.integers(0, 256, n, dtype=np.uint8).tobytes())
        b[cl - 40:cl + 40] = bytes(80)                       # a match across the first chunk boundary
        b[2 * cl:2 * cl + 300] = b[2 * cl - 300:2 * cl]      # chunk 2 starts with a copy of chunk 1's tail
        b[3 * cl + 5:4 * cl] = bytes([7]) * (cl - 5)         # last chunk: 5 literals, then one run to the end
        pats["junction%d" % n] = bytes(b)
    for name, buf in pats.items():
        for n in sizes:
            if n > len(buf):
                continue
            v = np.frombuffer(buf[:n], dtype=np.uint8).copy()
            enc = ctx.encode_columns([sb.LeafArray(sbo.U8, v)], sb.write_options(codec, None, None))[0]
            assert len(enc.metas) == 1
            ref = oracle_decode_column(sbo.U8, False, enc.data, enc.metas)
            assert np.array_equal(ref["values"], v), (name, n)
            dec = ctx.batch_read_array(sb.Column(sbo.U8, False, enc.data, enc.metas))
            assert np.array_equal(dec.values, v), (name, n)
