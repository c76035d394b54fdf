"""i128 / i256 leaves (Decimal128 / Decimal256 storage; src/util/mod.rs:77-78, src/write/primitive.rs:71-78,
src/compression/integer/traits.rs:28-39): compress_integer / decompress_integer over 16 / 32-byte integers.
CPU: oracle round trips.  GPU: decode of oracle-written pages (every applicable codec, forced and adaptive),
oracle decode of GPU-written pages, same codec tree as the oracle chooser."""
import numpy as np
import pytest
import sbo
from helpers import assert_same, oracle_decode_column, oracle_encode_column

WIDE = [(sbo.I128, 16), (sbo.I256, 32)]


def wide(values, W):
    """python ints -> little-endian two's complement V{W} array"""
    mask = (1 << (8 * W)) - 1
    return np.frombuffer(b"".join((int(v) & mask).to_bytes(W, "little") for v in values), dtype=f"V{W}").copy()


def shapes(rng, n, W):
    big = 1 << (8 * W - 2)
    yield "random", wide([int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**62)) % big * (1 if rng.random() < 0.5 else -1) for _ in range(n)], W)
    yield "decimal", wide(rng.integers(-10**12, 10**12, n), W)            # sign-extended small magnitudes
    yield "lowcard", wide([int(x) * 10**20 for x in rng.integers(0, 7, n)], W)
    yield "const", wide([-(10**25)] * n, W)
    yield "runs", wide(np.repeat(rng.integers(-10**9, 10**9, n // 23 + 1), 23)[:n], W)
    yield "freq_small", wide(np.where(rng.random(n) < 0.95, 20, rng.integers(0, 10**6, n)), W)      # max < 2^63, >= 256
    yield "freq_negtop", wide(np.where(rng.random(n) < 0.95, -5, rng.integers(-100, -6, n)), W)     # max.as_i64() < 256: Freq refused
    yield "freq_wrap", wide([(1 << 70) + 3 if rng.random() < 0.95 else (1 << 70) + int(rng.integers(4, 200)) for _ in range(n)], W)  # low 64 bits < 256


@pytest.mark.parametrize("type_,W", WIDE)
def test_oracle_roundtrip(type_, W):
    rng = np.random.default_rng(5)
    for n in (0, 1, 300, 1000):
        for name, v in shapes(rng, n, W):
            for validity in (None, rng.random(n) > 0.3):
                for force in (-1, sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ, sbo.C_ONEVALUE):
                    for default in (sbo.C_NONE, sbo.C_LZ4):
                        opts = sbo.make_opts(default, ratio=2.0, force=force)
                        data, metas = oracle_encode_column(type_, v, validity, page_size=256, opts=opts)
                        ref = oracle_decode_column(type_, validity is not None, data, metas)
                        assert ref["length"] == n
                        a, b = ref["values"].view(np.uint8).reshape(-1, W), v.view(np.uint8).reshape(-1, W)
                        keep = slice(None) if validity is None else validity
                        assert np.array_equal(a[keep], b[keep]), (name, force)


@pytest.mark.parametrize("type_,W", WIDE)
def test_oracle_chooser_on_wide_values(type_, W):
    """Bitpacking never applies (size_of::<T>() != 4); Freq needs `max.as_i64() >= 256` on the LOW 64 bits"""
    rng = np.random.default_rng(6)
    trees = {}
    for name, v in shapes(rng, 2048, W):
        page = sbo.write_page(type_, v, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
        trees[name] = sbo.stat_page(type_, False, page)
    assert trees["const"].startswith("OneValue")
    assert trees["lowcard"].startswith("Dict")
    assert trees["freq_small"].startswith("Freq")
    assert not trees["freq_negtop"].startswith("Freq") and not trees["freq_wrap"].startswith("Freq")
    assert trees["runs"].startswith(("Rle", "Dict"))
    assert all("Bitpacking" not in t.split("(")[0] and "Patas" not in t for t in trees.values())


@pytest.mark.gpu
@pytest.mark.parametrize("type_,W", WIDE)
def test_gpu_decodes_oracle_pages(ctx, type_, W):
    import strawboat_b200 as sb
    rng = np.random.default_rng(7)
    for n in (1, 777, 3000):
        for name, v in shapes(rng, n, W):
            for validity in (None, rng.random(n) > 0.3):
                for force in (-1, sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ, sbo.C_ONEVALUE):
                    opts = sbo.make_opts(sbo.C_LZ4, ratio=2.0, force=force)
                    data, metas = oracle_encode_column(type_, v, validity, page_size=1000, opts=opts)
                    ref = oracle_decode_column(type_, validity is not None, data, metas)
                    dec = ctx.batch_read_array(sb.Column(type_, validity is not None, data, metas))
                    assert_same(dec, ref, type_, validity is not None)


@pytest.mark.gpu
@pytest.mark.parametrize("type_,W", WIDE)
def test_oracle_decodes_gpu_pages(ctx, type_, W):
    import strawboat_b200 as sb
    rng = np.random.default_rng(8)
    for n in (1, 640, 2048, 3001):
        for name, v in shapes(rng, n, W):
            for validity in (None, rng.random(n) > 0.3):
                for force in (-1, sbo.C_RLE, sbo.C_DICT, sbo.C_FREQ, sbo.C_ONEVALUE):
                    for default in (sb.C_NONE, sb.C_LZ4):
                        wo = sb.write_options(default, 2.0, 1024, force_codec=force, seed=3)
                        enc = ctx.encode_columns([sb.LeafArray(type_, v, validity=validity)], wo)[0]
                        ref = oracle_decode_column(type_, validity is not None, enc.data, enc.metas)
                        a, b = ref["values"].view(np.uint8).reshape(-1, W), v.view(np.uint8).reshape(-1, W)
                        keep = slice(None) if validity is None else validity
                        assert ref["length"] == n and np.array_equal(a[keep], b[keep]), (name, force, default)
                        if force == -1:  # same codec tree as the oracle chooser, page by page (same sampler seed)
                            pos = 0
                            for pi, ((ln, nv), (o, l)) in enumerate(zip(enc.metas, sb.workloads.split_pages(n, 1024))):
                                val = None if validity is None else validity[o:o + l]
                                op = sbo.write_page(type_, v[o:o + l], val, opts=sbo.make_opts(default, ratio=2.0, seed=3 + pi, float_bitwise=1))
                                got = sbo.stat_page(type_, validity is not None, enc.data[pos:pos + ln])
                                exp = sbo.stat_page(type_, validity is not None, op)
                                assert got.split("[")[0] == exp.split("[")[0], (name, pi, got, exp)
                                if "Lz4" not in exp:
                                    assert enc.data[pos:pos + ln] == op, (name, pi, got)
                                pos += ln
