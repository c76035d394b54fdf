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
    if nullable and n == 0 and ref["validity"] is None:
        return  # nothing was ever pushed into the oracle's bitmap
    if nullable:
        assert dec.validity is not None
        assert np.array_equal(sbo.unpack_bits(dec.validity, n), sbo.unpack_bits(ref["validity"], n))
    else:
        assert dec.validity is None


# ---- nested pages (write_nested, src/write/serialize.rs:135-198,217-232) --------------------
def nested_thresholds(nested):
    cum_sum, cum_rep = [0], [0]
    for kind, nullable in nested:
        cum_sum.append(cum_sum[-1] + int(bool(nullable)) + int(kind == sbo.N_LIST))
        cum_rep.append(cum_rep[-1] + int(kind == sbo.N_LIST))
    return cum_sum, cum_rep


def gen_levels(nested, rng, rows, p_null=0.15, max_list=3):
    """Random Dremel (rep, def) entries for `rows` top-level rows of a column whose nesting is
    `nested` = [(kind, nullable)] root -> leaf.  Returns per-row entry lists."""
    cum_sum, cum_rep = nested_thresholds(nested)
    out = []

    def emit(d, rep, dlevel, acc):
        kind, nullable = nested[d]
        if nullable and rng.random() < p_null:
            acc.append((rep, dlevel))
            return
        dd = dlevel + int(bool(nullable))
        if kind == sbo.N_LIST:
            k = int(rng.integers(0, max_list + 1))
            if k == 0:
                acc.append((rep, dd))
                return
            for i in range(k):
                emit(d + 1, rep if i == 0 else cum_rep[d] + 1, dd + 1, acc)
        elif kind == sbo.N_STRUCT:
            emit(d + 1, rep, dd, acc)
        else:
            acc.append((rep, dd))

    for _ in range(rows):
        acc = []
        emit(0, 0, 0, acc)
        out.append(acc)
    return out


def leaf_slots(nested, reps, defs):
    """python restatement of the push rule of read_validity_nested for the LEAF depth only:
    returns (slot count, leaf validity bools)."""
    cum_sum, cum_rep = nested_thresholds(nested)
    D = len(nested)
    valid = []
    for rep, de in zip(reps, defs):
        req = False
        for d in range(D):
            kind, nullable = nested[d]
            right = rep <= cum_rep[d] and de >= cum_sum[d]
            if req or right:
                v = bool(nullable) and de > cum_sum[d]
                req = kind == sbo.N_STRUCT and not v
                if d == D - 1:
                    valid.append(bool(right and ((de != cum_sum[d]) or not nullable)))
    return len(valid), np.array(valid, dtype=bool)


def write_nested_page(type_, nested, row_entries, leaf_values_fn, opts=None):
    """One nested page: [u32 rows][u32 rep_len][u32 def_len][rep][def][VALUE_BLOCK].
    leaf_values_fn(n_slots, validity) -> values for compress_values.  Returns (page, num_values)."""
    cum_sum, cum_rep = nested_thresholds(nested)
    reps = [e[0] for r in row_entries for e in r]
    defs = [e[1] for r in row_entries for e in r]
    w_rep, w_def = int(cum_rep[-1]).bit_length(), int(cum_sum[-1]).bit_length()
    rep_b = sbo.levels_encode(reps, w_rep) if w_rep else b""
    def_b = sbo.levels_encode(defs, w_def) if w_def else b""
    n_slots, lval = leaf_slots(nested, reps, defs)
    values = leaf_values_fn(n_slots, lval)
    leaf_nullable = bool(nested[-1][1])
    block = sbo.compress_values(type_, values, validity=lval if leaf_nullable else None, opts=opts or sbo.make_opts())
    hdr = np.array([len(row_entries), len(rep_b), len(def_b)], dtype="<u4").tobytes()
    return hdr + rep_b + def_b + block, len(reps)


def assert_same_nested(dec, ref, type_, nested):
    """decoded nested leaf == oracle: leaf buffers + NestedState of every depth above the leaf."""
    leaf_nullable = bool(nested[-1][1])
    assert_same(dec, ref, type_, leaf_nullable)
    for d, (kind, nullable) in enumerate(nested[:-1]):
        r, g = ref["nested"][d], dec.nested[d]
        if kind == sbo.N_LIST:
            # the oracle keeps start offsets; create_list appends the end offset (= child length)
            assert g["len"] == len(r["offsets"])
            assert np.array_equal(g["offsets"][:-1], r["offsets"])
            child = None
            if d + 1 == len(nested) - 1:
                child = ref["length"]
            elif nested[d + 1][0] == sbo.N_LIST:
                child = len(ref["nested"][d + 1]["offsets"])
            elif nested[d + 1][1]:
                child = ref["nested"][d + 1]["validity_len"]
            if child is not None:
                assert g["offsets"][-1] == child
        if nullable:
            assert g["len"] == r["validity_len"]
            assert np.array_equal(sbo.unpack_bits(g["validity"], g["len"]), sbo.unpack_bits(r["validity"], g["len"]))
