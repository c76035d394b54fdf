"""CPU: pin the oracle's restatements of third-party byte layouts against independent
implementations available in this image (oracle/FORMAT_ASSUMPTIONS.md)."""
import numpy as np
import pytest
import sbo


def corpus(rng):
    yield b""
    yield b"a"
    yield b"abcd" * 1000
    yield bytes(rng.integers(0, 256, 5000, dtype=np.uint8))
    yield np.repeat(rng.integers(0, 256, 40, dtype=np.uint8), rng.integers(1, 400, 40)).tobytes()
    yield rng.integers(0, 65536, 8192).astype(np.float64).tobytes()
    yield np.cumsum(rng.integers(0, 4, 8192)).astype(np.int32).tobytes()
    yield b"\0" * 100000


def test_lz4_block_own_decoder_vs_liblz4_and_pyarrow():
    """LZ4 raw block (src/compression/basic.rs:88,115: no size prefix).  liblz4 is the C library
    the `lz4` crate wraps; pyarrow's lz4_raw codec is a second independent implementation."""
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("lz4_raw")
    rng = np.random.default_rng(0)
    for data in corpus(rng):
        comp = sbo.lz4_compress(data)                                   # liblz4 LZ4_compress_default
        assert sbo.lz4_decompress(comp, len(data)) == data              # own block decoder
        assert sbo.lz4_decompress(comp, len(data), use_lib=True) == data
        if len(data):
            assert codec.decompress(comp, decompressed_size=len(data)).to_pybytes() == data
            comp2 = codec.compress(data).to_pybytes()                   # pyarrow-produced block
            assert sbo.lz4_decompress(comp2, len(data)) == data


def test_lz4_corrupt_blocks_are_rejected():
    data = b"abcdabcdabcdabcdabcdabcdabcdabcd" * 10
    comp = bytearray(sbo.lz4_compress(data))
    with pytest.raises(sbo.OracleError):
        sbo.lz4_decompress(bytes(comp[:-3]), len(data))
    with pytest.raises(sbo.OracleError):
        sbo.lz4_decompress(bytes(comp), len(data) + 5)


def ref_hybrid_decode(buf, w, n):
    """Independent restatement of the Parquet hybrid RLE / bit-packing spec (Encodings.md)."""
    out, pos = [], 0
    while len(out) < n:
        header, shift = 0, 0
        while True:
            b = buf[pos]
            pos += 1
            header |= (b & 0x7f) << shift
            shift += 7
            if not b & 0x80:
                break
        if header & 1:
            nbytes = min((header >> 1) * w, len(buf) - pos)
            bits = int.from_bytes(buf[pos:pos + nbytes], "little")
            for i in range(nbytes * 8 // w):
                if len(out) < n:
                    out.append((bits >> (i * w)) & ((1 << w) - 1))
            pos += nbytes
        else:
            vb = (w + 7) // 8
            v = int.from_bytes(buf[pos:pos + vb], "little")
            pos += vb
            out += [v] * min(header >> 1, n - len(out))
    return out


def test_hybrid_rle_levels():
    """parquet2 encode_u32 / HybridRleDecoder as used for rep/def levels
    (src/write/serialize.rs:225, src/read/read_basic.rs:83-84)."""
    rng = np.random.default_rng(1)
    for w in (1, 2, 3, 5):
        for n in (0, 1, 7, 8, 9, 1000, 4099):
            lv = rng.integers(0, 1 << w, n).astype(np.uint32)
            enc = sbo.levels_encode(lv, w)
            # one bit-packed run, zero padded to ceil8(n) * w bytes
            groups = (n + 7) // 8
            hdr = (groups << 1) | 1
            ul = 1 if hdr < 0x80 else (2 if hdr < 0x4000 else 3)
            assert len(enc) == ul + groups * w
            assert list(sbo.hybrid_rle_decode(enc, w, n)) == list(lv)
            assert ref_hybrid_decode(enc, w, n) == list(lv)
            # older parquet2 truncated the final partial group: the decoder tolerates a short tail
            need = ul + (n * w + 7) // 8
            assert list(sbo.hybrid_rle_decode(enc[:need], w, n)) == list(lv)


def test_validity_section_is_parquet_bool_levels():
    """write_validity (serialize.rs:200-215): [u32 L][ULEB((ceil8(n)<<1)|1)][bitmap]."""
    rng = np.random.default_rng(2)
    for n in (1, 8, 9, 2048, 8192, 10007):
        v = rng.random(n) > 0.3
        p = sbo.write_page(sbo.I32, np.zeros(n, np.int32), validity=v)
        L = int.from_bytes(p[:4], "little")
        levels = p[4:4 + L]
        assert ref_hybrid_decode(levels, 1, n) == list(v.astype(int))
        assert sbo.value_block_offset((sbo.I32, True), p) == 4 + L


def test_bitpacker4x_layout_properties():
    """BitPacker4x (crate source absent -> layout restated, parity unpinned).
    Checks the documented structure: 16*b bytes, 4 interleaved lanes, LSB-first."""
    rng = np.random.default_rng(3)
    for bits in range(0, 33):
        hi = (1 << bits) - 1
        v = rng.integers(0, hi, 128, endpoint=True).astype(np.uint32) if bits else np.zeros(128, np.uint32)
        if bits:
            v[5] = hi
        nb, packed = sbo.bp4x_compress(v)
        assert nb == bits and len(packed) == 16 * bits
        assert np.array_equal(sbo.bp4x_decompress(packed, nb), v)
        words = np.frombuffer(packed, "<u4")
        for lane in range(4):  # lane l = every 4th output word, values l, l+4, l+8, ...
            stream = int.from_bytes(words[lane::4].tobytes(), "little") if bits else 0
            got = [(stream >> (k * bits)) & hi for k in range(32)]
            assert got == list(v[lane::4])
    v = np.sort(rng.integers(0, 1 << 20, 128)).astype(np.uint32)
    nb, packed = sbo.bp4x_compress(v, initial=7)
    assert np.array_equal(sbo.bp4x_decompress(packed, nb, initial=7), v)
    d = np.diff(np.concatenate([[7], v]).astype(np.int64)).astype(np.uint32)
    assert packed == sbo.bp4x_compress(d, num_bits=nb)[1]  # compress_sorted == compress(deltas)


def test_roaring_portable_format():
    """RoaringFormatSpec without run containers (parity unpinned)."""
    rng = np.random.default_rng(4)
    small = np.sort(rng.choice(60000, 300, replace=False)).astype(np.uint32)
    b = sbo.roaring_serialize(small)
    assert int.from_bytes(b[:4], "little") == 12346 and int.from_bytes(b[4:8], "little") == 1
    assert int.from_bytes(b[8:10], "little") == 0 and int.from_bytes(b[10:12], "little") == 299
    assert int.from_bytes(b[12:16], "little") == 16 and len(b) == 16 + 600
    assert np.array_equal(np.frombuffer(b[16:], "<u2"), small.astype(np.uint16))
    big = np.sort(rng.choice(200000, 90000, replace=False)).astype(np.uint32)  # bitmap + array containers
    b = sbo.roaring_serialize(big)
    assert np.array_equal(sbo.roaring_deserialize(b), big)
    assert len(sbo.roaring_serialize(np.zeros(0, np.uint32))) == 8


def test_snappy_raw_own_codec_vs_pyarrow():
    """Snappy raw format (src/compression/basic.rs:98-105,138-152 -> snap::raw).  The oracle's encoder / decoder
    are restated from the format description; pyarrow's snappy codec wraps Google's C++ library."""
    pa = pytest.importorskip("pyarrow")
    codec = pa.Codec("snappy")
    rng = np.random.default_rng(5)
    for data in corpus(rng):
        comp = sbo.common_compress(sbo.C_SNAPPY, data)
        assert sbo.common_decompress(sbo.C_SNAPPY, comp, len(data)) == data
        if len(data):
            assert codec.decompress(comp, decompressed_size=len(data)).to_pybytes() == data   # Google's decoder reads ours
            comp2 = codec.compress(data).to_pybytes()
            assert sbo.common_decompress(sbo.C_SNAPPY, comp2, len(data)) == data               # we read Google's
    with pytest.raises(sbo.OracleError):
        sbo.common_decompress(sbo.C_SNAPPY, sbo.common_compress(sbo.C_SNAPPY, b"abcdabcdabcd" * 10)[:-2], 120)
    with pytest.raises(sbo.OracleError):
        sbo.common_decompress(sbo.C_SNAPPY, sbo.common_compress(sbo.C_SNAPPY, b"abcdabcdabcd" * 10), 121)
