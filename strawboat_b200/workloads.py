"""Synthetic columns of BASELINE.json's configs, shared by tests and bench.py.

Pure numpy data generation: no codec logic lives here.
"""
import numpy as np

from ._capi import BINARY, BOOL, F64, I32, I64, LARGE_BINARY


def config1(n=1_000_000, seed=42):
    """single non-nullable i64 column, PCG64(42) full range, codec None."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [("c0", I64, rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64, endpoint=True), None)]


def _runs(rng, n, mean_run, dtype):
    lens = rng.geometric(1.0 / mean_run, size=int(n / mean_run * 1.3) + 16)
    while lens.sum() < n:
        lens = np.concatenate([lens, rng.geometric(1.0 / mean_run, size=1024)])
    vals = rng.integers(0, 1 << 40, len(lens)).astype(dtype)
    return np.repeat(vals, lens)[:n]


def config2(n=10_000_000, seed=42):
    """8 primitive columns, one distribution per codec."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cols = []
    cols.append(("c0_i32_sorted", I32, np.cumsum(rng.integers(0, 4, n), dtype=np.int64).astype(np.int32), None))
    cols.append(("c1_i32_lowcard", I32, rng.integers(0, 8, n).astype(np.int32), None))
    cols.append(("c2_i32_random", I32, rng.integers(-2**31, 2**31, n).astype(np.int32), None))
    cols.append(("c3_i64_const", I64, np.full(n, 7_000_000_007, dtype=np.int64), None))
    c4 = np.full(n, 20, dtype=np.int64)
    exc = rng.random(n) < 0.05
    c4[exc] = 10000 + rng.integers(0, 1 << 20, int(exc.sum()))
    cols.append(("c4_i64_freq", I64, c4, None))
    cols.append(("c5_i64_runs", I64, _runs(rng, n, 64, np.int64), None))
    cols.append(("c6_f64_lowcard", F64, rng.integers(0, 8, n).astype(np.float64), None))
    cols.append(("c7_f64_int16", F64, rng.integers(0, 65536, n).astype(np.float64), None))
    return cols


def random_strings(rng, n, uniq, null_density=0.0, large=False, sort_within=0):
    """decimal strings of integers(0, uniq) (tests/it/io.rs:385-397)."""
    ids = rng.integers(0, uniq, n)
    if sort_within:
        ids = ids.reshape(-1, sort_within) if n % sort_within == 0 else ids
        ids = np.sort(ids, axis=-1).reshape(-1)
    table = [str(i).encode() for i in range(uniq)]
    lens = np.array([len(t) for t in table], dtype=np.int64)[ids]
    validity = None
    if null_density > 0:
        validity = rng.random(n) >= null_density
        lens = np.where(validity, lens, 0)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    tab_bytes = np.frombuffer(b"".join(table), dtype=np.uint8)
    tab_off = np.zeros(uniq + 1, dtype=np.int64)
    np.cumsum([len(t) for t in table], out=tab_off[1:])
    # gather bytes
    total = int(offsets[-1])
    row_of = np.repeat(np.arange(n), lens)
    pos_in = np.arange(total) - offsets[:-1][row_of]
    data = tab_bytes[tab_off[ids[row_of]] + pos_in]
    odt = np.int64 if large else np.int32
    return offsets.astype(odt), data.astype(np.uint8), validity


def config3(n=10_000_000, seed=42, uniq=1000, null_density=0.4):
    rng = np.random.Generator(np.random.PCG64(seed))
    o1, d1, v1 = random_strings(rng, n, uniq, null_density, large=False)
    o2, d2, v2 = random_strings(rng, n, uniq, null_density, large=True)
    return [("s0_utf8", BINARY, (o1, d1), v1), ("s1_large_binary", LARGE_BINARY, (o2, d2), v2)]


def dict_strings(rng, n, uniq, null_density, large):
    """config 3 generator at full size (vectorised): decimal strings of integers(0, uniq)
    (tests/it/io.rs:385-397), null rows are empty slots.  Returns ((offsets, data), validity)."""
    width = len(str(uniq - 1))
    table = np.zeros((uniq, width), dtype=np.uint8)
    tlen = np.zeros(uniq, dtype=np.int64)
    for i in range(uniq):
        b = str(i).encode()
        table[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
        tlen[i] = len(b)
    ids = rng.integers(0, uniq, n)
    validity = rng.random(n) >= null_density
    lens = np.where(validity, tlen[ids], 0)
    off = np.zeros(n + 1, dtype=np.int64 if large else np.int32)
    np.cumsum(lens, out=off[1:])
    mask = np.arange(width)[None, :] < lens[:, None]
    return (off, table[ids][mask]), validity


def plain_strings(rng, n):
    """random printable strings of 4..15 bytes (north-star utf8 case: nothing for a codec to find)"""
    lens = rng.integers(4, 16, n)
    off = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    return off, rng.integers(48, 123, int(off[-1]), dtype=np.uint8)


CONFIG4_NESTED = [(1, True), (2, True), (0, True)]  # List<Struct<leaf>>: (N_LIST, N_STRUCT, N_PRIMITIVE), all nullable


def config4_levels(rng, rows):
    """Dremel levels of List<Struct<..>> rows: 10 % null lists, list lengths integers(0,3), 20 % null
    structs / leaves (tests/it/io.rs:280-292,399-415).  nested = CONFIG4_NESTED: max_rep 1, max_def 4.
    Returns (rep, def, first entry of every row)."""
    k = rng.integers(0, 4, rows)
    null_list = rng.random(rows) < 0.1
    cnt = np.where(null_list | (k == 0), 1, k)
    row_of = np.repeat(np.arange(rows), cnt)
    first = np.ones(len(row_of), dtype=bool)
    first[1:] = row_of[1:] != row_of[:-1]
    rep = (~first).astype(np.uint32)
    de = np.full(len(row_of), 4, dtype=np.uint32)
    r = rng.random(len(row_of))
    de[r < 0.2] = 3            # struct valid, leaf null
    de[r < 0.05] = 2           # null struct
    de[(k == 0)[row_of]] = 1   # empty list
    de[null_list[row_of]] = 0  # null list
    return rep, de, np.cumsum(cnt) - cnt


def config4(rows=4_000_000, seed=7):
    """three leaves of List<Struct<a:Int64, b:Float64, c:Utf8>> (config 4): shared levels, per-leaf values over
    the leaf slots (def >= 2), leaf validity def == 4.  Returns (rep, def, row_start, [(name, type, values, validity)])."""
    rng = np.random.default_rng(seed)
    rep, de, row_start = config4_levels(rng, rows)
    slots = de >= 2
    valid = de[slots] == 4
    ns = int(slots.sum())
    lens = np.where(valid, rng.integers(1, 6, ns), 0)
    off = np.zeros(ns + 1, np.int32)
    np.cumsum(lens, out=off[1:])
    leaves = [("a_i64", I64, rng.integers(0, 1 << 40, ns).astype(np.int64), valid),
              ("b_f64", F64, rng.integers(0, 1000, ns).astype(np.float64), valid),
              ("c_utf8", BINARY, (off, rng.integers(97, 123, int(off[-1])).astype(np.uint8)), valid)]
    return rep, de, row_start, leaves


def split_pages(n, page_size):
    """(offset, length) of every page: NativeWriter::encode_chunk, write/common.rs:54-58,79-86."""
    page_size = n if not page_size else min(page_size, n)
    return [(o, min(page_size, n - o)) for o in range(0, n, max(1, page_size))]


__all__ = ["config1", "config2", "config3", "config4", "config4_levels", "CONFIG4_NESTED", "dict_strings", "plain_strings",
           "random_strings", "split_pages", "BOOL"]
