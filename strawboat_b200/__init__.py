"""strawboat_b200 -- B200-native page encode/decode backend for the strawboat columnar format.

Thin Python harness over the C ABI (include/strawboat_b200.h).  The compute lives in
csrc/libstrawboat_b200.so (hand-written sm_100a CUDA); this module only marshals buffers.
Function names follow the reference's reader/writer API:

    read.batch_read_array  -> Context.batch_read_array / decode_columns
    read.column_iter_to_arrays(...).next() -> Context.decode_pages
    write.NativeWriter.encode_chunk (page loop) -> Context.encode_columns

There is no CPU fallback: importing works without a GPU (so symbols can be checked), but any
decode/encode call raises StrawboatError(SB_CUDA) when no CUDA device is usable.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (BINARY, BOOL, C_BITPACK, C_DELTABP, C_DICT, C_FREQ, C_LZ4, C_NONE, C_ONEVALUE, C_PATAS,  # noqa: F401
                    C_RLE, C_SNAPPY, C_ZSTD, F32, F64, I8, I16, I32, I64, I128, I256, LARGE_BINARY, MEM_DEVICE, MEM_HOST,
                    N_LIST, N_PRIMITIVE, N_STRUCT, NULL, STATUS_NAMES, U8, U16, U32, U64)

_lib = _capi.load()  # raises ImportError if the CUDA library has not been built

NP_OF = {I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, U8: np.uint8, U16: np.uint16,
         U32: np.uint32, U64: np.uint64, F32: np.float32, F64: np.float64,
         I128: np.dtype("V16"), I256: np.dtype("V32")}  # Decimal128 / Decimal256 storage: opaque 16 / 32-byte little-endian values


class StrawboatError(RuntimeError):
    def __init__(self, code, msg=""):
        name = STATUS_NAMES[code] if 0 <= code < len(STATUS_NAMES) else str(code)
        super().__init__(f"{name}: {msg}")
        self.code = code


def version():
    return _lib.sb_version().decode()


def make_leaf(type_, nullable=False, nested=None):
    lf = _capi.Leaf()
    lf.type = type_
    lf.nullable = int(bool(nullable))
    if nested:
        lf.n_nested = len(nested)
        for i, (k, nu) in enumerate(nested):
            lf.nested_kind[i] = k
            lf.nested_nullable[i] = int(bool(nu))
    return lf


class Column:
    """All pages of one leaf column (input of batch_read_array).

    data  : bytes / numpy uint8 array (host) or a torch CUDA uint8 tensor (device)
    metas : list of (length, num_values) == PageMeta
    """

    def __init__(self, type_, nullable, data, metas, nested=None):
        self.leaf = make_leaf(type_, nullable, nested)
        self.type = type_
        self.metas = list(metas)
        self._keep = None
        self._metas_c = None
        if hasattr(data, "data_ptr"):  # torch tensor
            assert data.is_cuda and data.is_contiguous() and data.element_size() == 1
            self.ptr, self.nbytes, self.mem = data.data_ptr(), data.numel(), MEM_DEVICE
            self._keep = data
        else:
            arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
            self._keep = arr
            self.ptr, self.nbytes, self.mem = (arr.ctypes.data if arr.size else 0), arr.size, MEM_HOST


    def metas_c(self):
        """PageMeta[] for the C ABI (built once per Column)."""
        if self._metas_c is None:
            a = np.array(self.metas, dtype=np.uint64).reshape(-1, 2) if self.metas else np.zeros((1, 2), np.uint64)
            self._metas_np = np.ascontiguousarray(a)
            self._metas_c = C.cast(self._metas_np.ctypes.data, C.POINTER(_capi.PageMeta))
        return self._metas_c


class _DevArray:
    """Zero-copy view of a device buffer for torch.as_tensor(..., device='cuda')."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        self._owner = owner


class Decoded:
    """One decoded Arrow array.  Host mode: numpy arrays.  Device mode: raw pointers that stay
    valid until release()."""

    def __init__(self, ctx, type_, out, idx, group, n_nested=0, copy=True):
        self.type = type_
        self.length = int(out.length)
        self.mem = out.mem
        npg = group.n_pages[idx]
        # one bulk copy (a Python loop over 10^4 pages would dominate a decode call)
        self.page_status = (np.ctypeslib.as_array(out.page_status, shape=(npg,)).copy() if npg else np.zeros(0, np.int32))
        self._group = group
        self.values_ptr, self.values_bytes = out.values, int(out.values_bytes)
        self.offsets_ptr, self.offsets_bytes = out.offsets, int(out.offsets_bytes)
        self.validity_ptr, self.validity_bytes = out.validity, int(out.validity_bytes)
        self.values = self.offsets = self.validity = None
        # nested leaves: NestedState entries of every depth above the leaf (read_validity_nested)
        self.nested = []
        for d in range(max(0, n_nested - 1)):
            self.nested.append({"len": int(out.nested_len[d]), "offsets_ptr": out.nested_offsets[d],
                                "validity_ptr": out.nested_validity[d], "offsets": None, "validity": None})
        if out.mem == MEM_HOST:
            def grab(p, n):
                # zero-copy view of the context's pinned host buffer; valid until release()
                if p and n:
                    a = np.ctypeslib.as_array((C.c_uint8 * int(n)).from_address(int(p)))
                    return a if not copy else a.copy()
                return np.zeros(0, np.uint8) if p or n == 0 else None
            v = grab(out.values, self.values_bytes)
            if type_ in NP_OF and v is not None:
                v = v.view(NP_OF[type_])
            self.values = v
            if type_ in (BINARY, LARGE_BINARY):
                o = grab(out.offsets, self.offsets_bytes)
                self.offsets = o.view(np.int64 if type_ == LARGE_BINARY else np.int32)
            self.validity = grab(out.validity, self.validity_bytes) if out.validity else None
            for nd in self.nested:
                if nd["offsets_ptr"]:
                    nd["offsets"] = grab(nd["offsets_ptr"], (nd["len"] + 1) * 8).view(np.int64)
                if nd["validity_ptr"]:
                    nd["validity"] = grab(nd["validity_ptr"], (nd["len"] + 7) // 8)

    def release(self):
        """give the buffers of this call (all columns decoded together) back to the context."""
        self._group.release()

    def device_view(self, which="values"):
        ptr, n = {"values": (self.values_ptr, self.values_bytes), "offsets": (self.offsets_ptr, self.offsets_bytes),
                  "validity": (self.validity_ptr, self.validity_bytes)}[which]
        return _DevArray(ptr, n, self._group)


class _OutGroup:
    def __init__(self, ctx, outs, n, n_pages):
        self.ctx, self.outs, self.n, self.n_pages = ctx, outs, n, n_pages
        self.released = False

    def release(self):
        if not self.released and self.ctx._h:
            _lib.sb_release_columns(self.ctx._h, self.outs, self.n)
        self.released = True

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Context:
    """sb_ctx: one per host thread; owns a stream and device scratch pools."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        rc = _lib.sb_ctx_create(device, C.byref(h))
        if rc != _capi.SB_OK:
            raise StrawboatError(rc, "sb_ctx_create failed: no usable CUDA device (strawboat_b200 has no CPU fallback)")
        self._h = h
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, stream):
        ptr = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        self._check(_lib.sb_ctx_set_stream(self._h, C.c_void_p(ptr)))

    def close(self):
        if self._h:
            _lib.sb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _capi.SB_OK:
            raise StrawboatError(rc, _lib.sb_last_error(self._h).decode())

    def last_stats(self):
        s = _capi.Stats()
        self._check(_lib.sb_last_stats(self._h, C.byref(s)))
        return {"pages": s.pages, "bytes_in": s.bytes_in, "bytes_out": s.bytes_out,
                "kernel_launches": s.kernel_launches, "device_ms": s.device_ms,
                "main_kernel_ms": s.main_kernel_ms, "lz4_kernel_ms": s.lz4_kernel_ms, "light_kernel_ms": s.light_kernel_ms,
                "lz4_bytes": s.lz4_bytes, "host_ms": s.host_ms,
                "codec_pages": {i: s.codec_pages[i] for i in range(32) if s.codec_pages[i]}}

    # ---- decode ----------------------------------------------------------------------
    def _marshal(self, columns):
        n = len(columns)
        ins = (_capi.ColumnIn * n)()
        keep = []
        for i, col in enumerate(columns):
            ins[i].leaf = col.leaf
            ins[i].bytes = col.ptr
            ins[i].nbytes = col.nbytes
            ins[i].mem = col.mem
            m = col.metas_c()
            keep.append(m)
            ins[i].metas = m
            ins[i].n_pages = len(col.metas)
        return ins, keep

    def decode_columns(self, columns, out="host", raise_on_page_error=True, per_page=False, copy=True):
        """batch_read_array for many leaf columns at once: one concatenated array per column.
        out="host", copy=False returns numpy views of the context's pinned buffers (no extra
        host copy); they stay valid until the returned arrays' group is released."""
        n = len(columns)
        ins, keep = self._marshal(columns)
        outs = (_capi.ColumnOut * n)()
        fn = _lib.sb_decode_pages if per_page else _lib.sb_decode_columns
        rc = fn(self._h, ins, n, MEM_HOST if out == "host" else MEM_DEVICE, outs)
        if rc != _capi.SB_OK and (raise_on_page_error or not any(outs[i]._owner for i in range(n))):
            msg = _lib.sb_last_error(self._h).decode()
            _lib.sb_release_columns(self._h, outs, n)
            raise StrawboatError(rc, msg)
        group = _OutGroup(self, outs, n, [len(c.metas) for c in columns])
        res = [Decoded(self, columns[i].type, outs[i], i, group, columns[i].leaf.n_nested, copy) for i in range(n)]
        if out == "host" and copy:
            group.release()
        return res

    # ---- plan -> allocate -> run, and the asynchronous form --------------------------------------------------
    def plan_columns(self, columns):
        """sb_plan_columns: exact buffer sizes of every column (runs the size pass for binary / nested leaves)"""
        n = len(columns)
        ins, keep = self._marshal(columns)
        sizes = (_capi.ColumnSizes * n)()
        self._check(_lib.sb_plan_columns(self._h, ins, n, sizes))
        return [{"length": int(s.length), "values_bytes": int(s.values_bytes), "offsets_bytes": int(s.offsets_bytes),
                 "validity_bytes": int(s.validity_bytes)} for s in sizes]

    @staticmethod
    def _out_buffers(bufs, n):
        """bufs: per column None or a dict {values, offsets, validity} of writable uint8 buffers (CUDA tensors for
        out="device", pinned / ordinary numpy arrays for out="host")"""
        if bufs is None:
            return None, []
        arr, keep = (_capi.OutBuffers * n)(), []
        for i, b in enumerate(bufs):
            for name in ("values", "offsets", "validity"):
                a = (b or {}).get(name)
                if a is None:
                    continue
                keep.append(a)
                if hasattr(a, "data_ptr"):
                    ptr, cap = a.data_ptr(), a.numel() * a.element_size()
                else:
                    ptr, cap = a.ctypes.data, a.nbytes
                setattr(arr[i], name, ptr)
                setattr(arr[i], name + "_cap", cap)
        return arr, keep

    def decode_columns_async(self, columns, out="host", bufs=None, copy=False):
        """sb_decode_columns_async: returns a handle; handle.wait() -> the decoded columns (as decode_columns).
        Several contexts driven from one thread overlap their copies and kernels."""
        n = len(columns)
        ins, keep = self._marshal(columns)
        outs = (_capi.ColumnOut * n)()
        ob, keep2 = self._out_buffers(bufs, n)
        rc = _lib.sb_decode_columns_async(self._h, ins, n, MEM_HOST if out == "host" else MEM_DEVICE, ob, outs)
        if rc != _capi.SB_OK:
            raise StrawboatError(rc, _lib.sb_last_error(self._h).decode())
        ctx = self

        class Handle:
            def ready(self_):
                return bool(_lib.sb_decode_ready(ctx._h))

            def wait(self_, raise_on_page_error=True):
                rc = _lib.sb_decode_wait(ctx._h)
                self_._keep = (ins, keep, ob, keep2, columns)
                if rc != _capi.SB_OK and (raise_on_page_error or not any(outs[i]._owner for i in range(n))):
                    msg = _lib.sb_last_error(ctx._h).decode()
                    _lib.sb_release_columns(ctx._h, outs, n)
                    raise StrawboatError(rc, msg)
                group = _OutGroup(ctx, outs, n, [len(c.metas) for c in columns])
                res = [Decoded(ctx, columns[i].type, outs[i], i, group, columns[i].leaf.n_nested, copy) for i in range(n)]
                if out == "host" and copy:
                    group.release()
                return res
        return Handle()

    def decode_columns_into(self, columns, bufs, out="device"):
        """sb_decode_columns_into: decode straight into caller-owned buffers (see plan_columns)"""
        return self.decode_columns_async(columns, out=out, bufs=bufs).wait()

    def batch_read_array(self, column, out="host"):
        """read::batch_read::batch_read_array for one leaf column."""
        return self.decode_columns([column], out=out)[0]

    def decode_pages(self, pages, out="host", raise_on_page_error=True):
        """column_iter_to_arrays(...).next(): one array per page; `pages` are one-page Columns."""
        return self.decode_columns(pages, out=out, raise_on_page_error=raise_on_page_error, per_page=True)


# ---- encode ------------------------------------------------------------------------------
def write_options(default_compression=C_NONE, default_compress_ratio=None, max_page_size=None, forbidden=(), force_codec=-1, seed=0):
    """WriteOptions (src/write/common.rs:37-45) + the explicit sampler seed / force-codec knobs."""
    o = _capi.WriteOptions()
    o.default_compression = default_compression
    o.default_compress_ratio = -1.0 if default_compress_ratio is None else float(default_compress_ratio)
    o.max_page_size = 0 if not max_page_size else int(max_page_size)
    m = 0
    for c in forbidden:
        m |= 1 << c
    o.forbidden_mask = m
    o.force_codec = force_codec
    o.seed = seed
    return o


class LeafArray:
    """One flat leaf array (what to_leaves yields, src/write/common.rs:68).

    values   : numpy array (primitives), bool ndarray (BOOL), (offsets, data) for BINARY / LARGE_BINARY;
               torch CUDA tensors are accepted for device-resident input
    validity : bool ndarray, one entry per row (or a packed LSB-first uint8 bitmap when
               `packed_validity`; device tensors are always packed), or None
    """

    def __init__(self, type_, values, validity=None, nullable=None, length=None, packed_validity=False,
                 nested=None, rep_levels=None, def_levels=None, rows=None):
        """nested leaves: `nested` = [(kind, nullable)] root -> leaf, `rep_levels` / `def_levels` = the
        column's Dremel levels (uint32, host arrays or CUDA tensors), `rows` = top-level rows; `values` /
        `validity` then cover the leaf slots."""
        self.type = type_
        self.nullable = (validity is not None) if nullable is None else bool(nullable)
        self._keep = []
        self.nested = nested
        self.rep_ptr = self.def_ptr = 0
        self.n_levels = 0
        self.rows = 0 if rows is None else int(rows)
        if nested:
            for name, lv in (("rep_ptr", rep_levels), ("def_ptr", def_levels)):
                if lv is None:
                    continue
                if hasattr(lv, "data_ptr"):
                    self._keep.append(lv)
                    setattr(self, name, lv.data_ptr())
                    self.n_levels = lv.numel()
                else:
                    a = np.ascontiguousarray(lv, dtype=np.uint32)
                    self._keep.append(a)
                    setattr(self, name, a.ctypes.data if a.size else 0)
                    self.n_levels = len(a)
        self.offsets_ptr = 0
        self.values_bytes = 0
        self.mem = MEM_HOST

        def host(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return (a.ctypes.data if a.size else 0), a.nbytes

        def dev(t):
            self._keep.append(t)
            self.mem = MEM_DEVICE
            return t.data_ptr(), t.numel() * t.element_size()

        is_dev = lambda x: hasattr(x, "data_ptr")  # noqa: E731
        if type_ == NULL:
            self.length = int(values if length is None else length)
            self.values_ptr = 0
        elif type_ == BOOL:
            if is_dev(values):
                assert length is not None, "device BOOL input is a packed bitmap: pass length"
                self.values_ptr, self.values_bytes = dev(values)
                self.length = int(length)
            else:
                self.length = len(values)
                self.values_ptr, self.values_bytes = host(np.packbits(np.asarray(values, dtype=bool), bitorder="little"), np.uint8)
        elif type_ in (BINARY, LARGE_BINARY):
            offsets, data = values[0], values[1]
            if is_dev(offsets):
                self.offsets_ptr, _ = dev(offsets)
                self.values_ptr, self.values_bytes = dev(data)
                self.length = offsets.numel() - 1
            else:
                self.offsets_ptr, _ = host(offsets, np.int64 if type_ == LARGE_BINARY else np.int32)
                self.values_ptr, self.values_bytes = host(data, np.uint8)
                self.length = len(offsets) - 1
        else:
            if is_dev(values):
                self.values_ptr, self.values_bytes = dev(values)
                self.length = values.numel()
            else:
                self.values_ptr, self.values_bytes = host(values, NP_OF[type_])
                self.length = len(values)
        self.validity_ptr = 0
        if validity is not None:
            if is_dev(validity):
                self.validity_ptr, _ = dev(validity)  # packed bitmap
            else:
                v = np.asarray(validity)
                if not packed_validity:
                    assert len(v) == self.length, "validity must have one entry per row"
                    v = np.packbits(v.astype(bool), bitorder="little")
                self.validity_ptr, _ = host(v, np.uint8)


class Encoded:
    """Encoded pages of one leaf column: `data` = the column body (pages back to back),
    `metas` = [(length, num_values)] for the footer (PageMeta, src/lib.rs:71-80)."""

    def __init__(self, out, copy=True):
        self.metas = [(int(out.metas[i].length), int(out.metas[i].num_values)) for i in range(out.n_pages)]
        self.nbytes = int(out.nbytes)
        self.ptr = out.bytes
        self.mem = out.mem
        self.data = None
        if out.mem == MEM_HOST:
            self.data = C.string_at(out.bytes, self.nbytes) if self.nbytes else b""


def _encode_columns(self, arrays, options=None, out="host"):
    """NativeWriter::encode_chunk page loop for flat leaves, on the GPU."""
    options = options or write_options()
    n = len(arrays)
    ins = (_capi.LeafArray * n)()
    for i, a in enumerate(arrays):
        ins[i].leaf = make_leaf(a.type, a.nullable, a.nested)
        ins[i].rep_levels = a.rep_ptr
        ins[i].def_levels = a.def_ptr
        ins[i].n_levels = a.n_levels
        ins[i].rows = a.rows
        ins[i].length = a.length
        ins[i].values = a.values_ptr
        ins[i].values_bytes = a.values_bytes
        ins[i].offsets = a.offsets_ptr
        ins[i].validity = a.validity_ptr
        ins[i].mem = a.mem
    outs = (_capi.EncodedColumn * n)()
    rc = _lib.sb_encode_columns(self._h, ins, n, C.byref(options), MEM_HOST if out == "host" else MEM_DEVICE, outs)
    if rc != _capi.SB_OK:
        msg = _lib.sb_last_error(self._h).decode()
        _lib.sb_release_encoded(self._h, outs, n)
        raise StrawboatError(rc, msg)
    res = [Encoded(outs[i]) for i in range(n)]
    if out == "host":
        _lib.sb_release_encoded(self._h, outs, n)
    else:
        for i, r in enumerate(res):
            r._outs, r._n, r._ctx, r._index = outs, n, self, i
    return res


def _release_encoded(self, encoded):
    if encoded and getattr(encoded[0], "_outs", None) is not None:
        _lib.sb_release_encoded(self._h, encoded[0]._outs, encoded[0]._n)
        for r in encoded:
            r._outs = None


def stat_page(type_, nullable, page, nested=None):
    """stat::stat_simple for one page (src/stat.rs:63-152): (codec tree as text, PageInfo dict).  Host only."""
    info = _capi.PageInfo()
    tree = C.create_string_buffer(256)
    page = bytes(page)
    rc = _lib.sb_stat_page(C.byref(make_leaf(type_, nullable, nested)), page, len(page), C.byref(info), tree, 256)
    if rc != _capi.SB_OK:
        raise StrawboatError(rc, "sb_stat_page")
    d = {f: getattr(info, f) for f, _ in _capi.PageInfo._fields_ if f != "path"}
    d["path"] = [info.path[i] for i in range(min(info.depth, 4))]
    return tree.value.decode(), d


class DeviceLevels:
    """Dremel levels generated on the device (sb_nested_levels): quacks like a CUDA uint32 tensor for LeafArray."""

    def __init__(self, ctx, ptr, n):
        self._ctx, self._ptr, self._n = ctx, ptr, n

    def data_ptr(self):
        return self._ptr

    def numel(self):
        return self._n

    def free(self):
        if self._ptr and self._ctx._h:
            _lib.sb_free_device(self._ctx._h, C.c_void_p(self._ptr))
        self._ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _nested_levels(self, path):
    """arrow2's write_rep_and_def over the `Nested` descriptors of one leaf, on the device.
    path: root -> leaf list of dicts {kind, nullable, length, offsets (int32 / int64 ndarray or CUDA tensor, lists),
    validity (bool ndarray / packed CUDA uint8 tensor / None)}.  Returns (rep, def, n_slots): DeviceLevels usable as
    LeafArray(rep_levels=, def_levels=)."""
    n = len(path)
    lv = (_capi.NestedLevel * n)()
    keep, mem = [], None
    for i, d in enumerate(path):
        lv[i].kind, lv[i].nullable, lv[i].length = d["kind"], int(bool(d.get("nullable"))), int(d["length"])
        for name in ("offsets", "validity"):
            a = d.get(name)
            if a is None:
                continue
            if hasattr(a, "data_ptr"):
                m, ptr = MEM_DEVICE, a.data_ptr()
                if name == "offsets":
                    lv[i].offset_width = a.element_size()
            else:
                m = MEM_HOST
                if name == "validity":
                    a = np.packbits(np.asarray(a, dtype=bool), bitorder="little")
                else:
                    a = np.ascontiguousarray(a)
                    assert a.dtype in (np.int32, np.int64)
                    lv[i].offset_width = a.dtype.itemsize
                ptr = a.ctypes.data if a.size else 0
            keep.append(a)
            assert mem in (None, m), "all nested buffers must live in the same memory space"
            mem = m
            setattr(lv[i], name, ptr)
    rep, de = C.c_void_p(), C.c_void_p()
    nl, ns = C.c_uint64(), C.c_uint64()
    self._check(_lib.sb_nested_levels(self._h, lv, n, MEM_HOST if mem is None else mem, C.byref(rep), C.byref(de), C.byref(nl), C.byref(ns)))
    return DeviceLevels(self, rep.value, nl.value), DeviceLevels(self, de.value, nl.value), int(ns.value)


def make_field(kind, type_=NULL, nullable=False, children=(), name="", utf8=False, large=False):
    """sb_field tree (the Field the reference's readers receive).  Keeps its children alive."""
    f = _capi.Field()
    f.kind, f.type, f.utf8, f.nullable, f.large = kind, type_, int(utf8), int(bool(nullable)), int(large)
    f.name = name.encode()
    if children:
        arr = (_capi.Field * len(children))(*children)
        f.n_children = len(children)
        f.children = C.cast(arr, C.POINTER(_capi.Field))
        f._keep = (arr, children)
    return f


def _read_arrow(self, columns, field):
    """batch_read_array + create_list / create_struct: decode the leaves of one field and assemble them into ONE
    Arrow array handed out through the C Data Interface (imported here with pyarrow; host buffers, zero copy).  The
    array's buffers go back to this context when the last reference to it dies: keep the context alive."""
    import pyarrow as pa
    n = len(columns)
    ins, keep = self._marshal(columns)
    outs = (_capi.ColumnOut * n)()
    rc = _lib.sb_decode_columns(self._h, ins, n, MEM_HOST, outs)
    if rc != _capi.SB_OK:
        msg = _lib.sb_last_error(self._h).decode()
        _lib.sb_release_columns(self._h, outs, n)
        raise StrawboatError(rc, msg)
    arr, sch = _capi.ArrowArray(), _capi.ArrowSchema()
    rc = _lib.sb_export_arrow(self._h, C.byref(field), outs, n, C.byref(arr), C.byref(sch))
    if rc != _capi.SB_OK:
        msg = _lib.sb_last_error(self._h).decode()
        _lib.sb_release_columns(self._h, outs, n)
        raise StrawboatError(rc, msg)
    return pa.Array._import_from_c(C.addressof(arr), C.addressof(sch))


def comm_unique_id():
    """ncclGetUniqueId on the calling rank (rank 0 makes it and hands it to the others out of band)."""
    buf = C.create_string_buffer(128)
    rc = _lib.sb_comm_unique_id(buf)
    if rc != _capi.SB_OK:
        raise StrawboatError(rc, "sb_comm_unique_id: libnccl.so.2 could not be loaded")
    return buf.raw


class Comm:
    """sb_comm: the NCCL communicator of the encode gather (one process per GPU; leaf c lives on rank c mod world)."""

    def __init__(self, ctx, rank, world, unique_id):
        h = C.c_void_p()
        ctx._check(_lib.sb_comm_create(ctx._h, rank, world, unique_id, C.byref(h)))
        self._h, self.ctx, self.rank, self.world = h, ctx, rank, world

    def close(self):
        if self._h:
            _lib.sb_comm_destroy(self._h)
            self._h = None

    def gather_encoded(self, encoded, n_total, writer=0):
        """Collective.  `encoded`: this rank's device-resident Encoded columns in leaf order (from
        Context.encode_columns(..., out="device")).  Returns (list of Encoded in leaf order on the writer / None,
        stats dict).  The writer's columns share one device buffer: the file's body region."""
        n = len(encoded)
        ins = (_capi.EncodedColumn * max(1, n))()
        for i, e in enumerate(encoded):
            ins[i] = e._outs[e._index] if getattr(e, "_outs", None) is not None else e._raw
        outs = (_capi.EncodedColumn * max(1, n_total))() if self.rank == writer else None
        st = _capi.GatherStats()
        self.ctx._check(_lib.sb_gather_encoded(self.ctx._h, self._h, ins, n, n_total, writer, outs, C.byref(st)))
        stats = {"bytes_moved": int(st.bytes_moved), "total_bytes": int(st.total_bytes), "gather_ms": float(st.gather_ms)}
        if self.rank != writer:
            return None, stats
        res = [Encoded(outs[i]) for i in range(n_total)]
        for i, r in enumerate(res):
            r._outs, r._n, r._ctx, r._index = outs, n_total, self.ctx, i
        return res, stats


Context.encode_columns = _encode_columns
Context.release_encoded = _release_encoded
Context.nested_levels = _nested_levels
Context.read_arrow = _read_arrow
