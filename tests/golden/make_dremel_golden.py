"""Generates tests/golden/dremel_pyarrow.npz: Dremel (rep, def) level streams of config 4's schema
`List<Struct<a:Int64, b:Float64, c:Utf8>>` AS WRITTEN BY PYARROW's Parquet writer (an independent
implementation of the Dremel shredding arrow2 performs for the reference), together with the Arrow
structure they came from.

What it pins (oracle/FORMAT_ASSUMPTIONS.md #6): for a null list / empty list / null
struct / null leaf, which (rep, def) pair is written, how many level entries and leaf slots a row owns,
and what `read_validity_nested` (src/read/read_basic.rs:65-173) must rebuild from them: list offsets
and validity, struct validity (children keep a slot under a null struct), leaf validity.

The level bytes are taken verbatim from the file's DataPageV2 (uncompressed, no dictionary): hybrid-RLE
streams that mix RLE and bit-packed runs -- a second writer's streams for the level decoder too.

    python tests/golden/make_dremel_golden.py
"""
import io
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- minimal Thrift compact-protocol reader (just enough for parquet.thrift PageHeader) ----------
class Thrift:
    def __init__(self, buf, pos):
        self.b, self.p = buf, pos

    def varint(self):
        r, s = 0, 0
        while True:
            x = self.b[self.p]
            self.p += 1
            r |= (x & 0x7f) << s
            s += 7
            if not x & 0x80:
                return r

    def zigzag(self):
        v = self.varint()
        return (v >> 1) ^ -(v & 1)

    def skip(self, t):
        if t in (1, 2):
            return
        if t == 3:
            self.p += 1
        elif t in (4, 5, 6):
            self.zigzag()
        elif t == 7:
            self.p += 8
        elif t == 8:
            self.p += self.varint()
        elif t in (9, 10):
            h = self.b[self.p]
            self.p += 1
            n = h >> 4
            if n == 15:
                n = self.varint()
            for _ in range(n):
                self.skip(h & 15)
        elif t == 12:
            self.struct()
        else:
            raise ValueError("thrift type %d" % t)

    def struct(self):
        """{field id: value}; nested structs as dicts, everything else that is not an int / bool skipped"""
        out, fid = {}, 0
        while True:
            h = self.b[self.p]
            self.p += 1
            if h == 0:
                return out
            d, t = h >> 4, h & 15
            fid = fid + d if d else self.zigzag()
            if t in (1, 2):
                out[fid] = t == 1
            elif t in (4, 5, 6):
                out[fid] = self.zigzag()
            elif t == 12:
                out[fid] = self.struct()
            else:
                self.skip(t)


def page_levels(path_or_buf, column):
    """(rep bytes, def bytes, num_values, num_rows) of the single DataPageV2 of `column` in row group 0"""
    buf = path_or_buf
    md = pq.ParquetFile(io.BytesIO(buf)).metadata
    cc = md.row_group(0).column(column)
    assert cc.dictionary_page_offset is None and cc.compression == "UNCOMPRESSED"
    t = Thrift(buf, cc.data_page_offset)
    hdr = t.struct()
    assert hdr[1] == 3, "DataPageV2 expected"
    v2 = hdr[8]
    num_values, num_rows, dl, rl = v2[1], v2[3], v2[5], v2[6]
    body = t.p
    return bytes(buf[body:body + rl]), bytes(buf[body + rl:body + rl + dl]), num_values, num_rows


def build(rows=700, seed=11):
    rng = np.random.default_rng(seed)
    py = []
    for r in range(rows):
        u = rng.random()
        if u < 0.12:
            py.append(None)                       # null list
        elif u < 0.24:
            py.append([])                         # empty list
        else:
            items = []
            for _ in range(int(rng.integers(1, 4))):
                w = rng.random()
                if w < 0.15:
                    items.append(None)            # null struct
                else:
                    items.append({"a": None if rng.random() < 0.2 else int(rng.integers(-2**40, 2**40)),
                                  "b": None if rng.random() < 0.2 else float(rng.integers(0, 1000)),
                                  "c": None if rng.random() < 0.2 else "s%d" % int(rng.integers(0, 50))})
            py.append(items)
    # the four shapes the generator could miss by chance, pinned explicitly at the front and the back
    edge = [None, [], [None], [{"a": None, "b": None, "c": None}], [{"a": 1, "b": 2.0, "c": "x"}, None, {"a": None, "b": 3.0, "c": ""}]]
    py = edge + py + edge[::-1]
    st = pa.struct([pa.field("a", pa.int64()), pa.field("b", pa.float64()), pa.field("c", pa.string())])
    arr = pa.array(py, type=pa.list_(pa.field("item", st)))
    return arr


def main():
    arr = build()
    tbl = pa.table({"col": arr})
    sink = io.BytesIO()
    pq.write_table(tbl, sink, compression="NONE", use_dictionary=False, data_page_version="2.0", write_statistics=False,
                   data_page_size=1 << 30)
    buf = sink.getvalue()
    out = {}
    rows = len(arr)
    # Arrow structure (the truth the reader must rebuild)
    offsets = np.asarray(arr.offsets, dtype=np.int64)
    list_valid = np.asarray(arr.is_valid())
    structs = arr.values                                    # struct slots = list child length
    struct_valid = np.asarray(structs.is_valid())
    out["rows"] = np.int64(rows)
    out["list_offsets"] = offsets                           # rows + 1 entries
    out["list_validity"] = list_valid
    out["struct_validity"] = struct_valid
    for ci, name in enumerate("abc"):
        rep, de, nv, nr = page_levels(buf, ci)
        assert nr == rows
        leaf = structs.field(name)
        out["%s_rep" % name] = np.frombuffer(rep, np.uint8)
        out["%s_def" % name] = np.frombuffer(de, np.uint8)
        out["%s_num_values" % name] = np.int64(nv)
        out["%s_validity" % name] = struct_valid & np.asarray(leaf.is_valid())  # a slot under a null struct is null
        if name == "c":
            s = [x.as_py() or "" for x in leaf]
            s = [v if ok else "" for v, ok in zip(s, out["c_validity"])]
            out["c_offsets"] = np.concatenate([[0], np.cumsum([len(v.encode()) for v in s])]).astype(np.int32)
            out["c_data"] = np.frombuffer("".join(s).encode(), np.uint8)
        else:
            v = np.asarray(leaf.fill_null(0))
            out["%s_values" % name] = np.where(out["%s_validity" % name], v, 0).astype(np.int64 if name == "a" else np.float64)
    np.savez_compressed(os.path.join(HERE, "dremel_pyarrow.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
