"""Test helpers: the oracle (oracle/sbo.py) is the checker; strawboat_b200 is the thing checked."""
import numpy as np
import sbo

from strawboat_b200.workloads import split_pages


def slice_values(type_, values, o, l):
    if type_ in (sbo.BINARY, sbo.LARGE_BINARY):
        offsets, data = values[0], values[1]
        return (offsets[o:o + l + 1], data, len(data))  # sliced array keeps the whole values buffer
    return values[o:o + l]


def oracle_encode_column(type_, values, validity=None, nullable=None, page_size=None, opts=None, n=None, seed=0):
    """NativeWriter::encode_chunk page loop for one flat leaf, run on the oracle.
    Returns (column bytes, [(length, num_values)])."""
    if nullable is None:
        nullable = validity is not None
    if n is None:
        n = len(values[0]) - 1 if type_ in (sbo.BINARY, sbo.LARGE_BINARY) else (int(values) if type_ == sbo.NULL else len(values))
    opts = opts or sbo.make_opts()
    out, metas = [], []
    for pi, (o, l) in enumerate(split_pages(n, page_size)):
        opts.seed = seed + pi
        v = l if type_ == sbo.NULL else slice_values(type_, values, o, l)
        val = None if validity is None else validity[o:o + l]
        page = sbo.write_page(type_, v, validity=val, nullable=nullable, opts=opts)
        out.append(page)
        metas.append((len(page), l))
    return b"".join(out), metas


def oracle_decode_column(type_, nullable, data, metas, nested=None):
    pages, pos = [], 0
    for length, nv in metas:
        pages.append((data[pos:pos + length], nv))
        pos += length
    return sbo.read_column(sbo.make_leaf(type_, nullable, nested), pages)


def assert_same(dec, ref, type_, nullable):
    """bit-exact comparison of a strawboat_b200.Decoded (host) with the oracle's column."""
    assert dec.length == ref["length"]
    if type_ == sbo.NULL:
        return
    n = dec.length
    if type_ == sbo.BOOL:
        got = sbo.unpack_bits(dec.values, n)
        exp = sbo.unpack_bits(ref["values"], n)
        assert np.array_equal(got, exp)
    elif type_ in (sbo.BINARY, sbo.LARGE_BINARY):
        assert np.array_equal(dec.offsets, ref["offsets"])
        assert np.array_equal(dec.values, ref["values"])
    else:
        assert dec.values.dtype == ref["values"].dtype
        assert np.array_equal(dec.values.view(np.uint8), ref["values"].view(np.uint8))
    if nullable:
        assert dec.validity is not None
        assert np.array_equal(sbo.unpack_bits(dec.validity, n), sbo.unpack_bits(ref["validity"], n))
    else:
        assert dec.validity is None
