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
                    C_RLE, C_SNAPPY, C_ZSTD, F32, F64, I8, I16, I32, I64, LARGE_BINARY, MEM_DEVICE, MEM_HOST,
                    N_LIST, N_PRIMITIVE, N_STRUCT, NULL, STATUS_NAMES, U8, U16, U32, U64)

_lib = _capi.load()  # raises ImportError if the CUDA library has not been built

NP_OF = {I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, U8: np.uint8, U16: np.uint16,
         U32: np.uint32, U64: np.uint64, F32: np.float32, F64: np.float64}


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

    def __init__(self, ctx, type_, out, idx, group, n_nested=0):
        self.type = type_
        self.length = int(out.length)
        self.mem = out.mem
        self.page_status = [out.page_status[i] for i in range(group.n_pages[idx])]
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
                return np.frombuffer(C.string_at(p, n), dtype=np.uint8).copy() if p and n else (np.zeros(0, np.uint8) if p or n == 0 else None)
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

    def decode_columns(self, columns, out="host", raise_on_page_error=True, per_page=False):
        """batch_read_array for many leaf columns at once: one concatenated array per column."""
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
        res = [Decoded(self, columns[i].type, outs[i], i, group, columns[i].leaf.n_nested) for i in range(n)]
        if out == "host":
            group.release()
        return res

    def batch_read_array(self, column, out="host"):
        """read::batch_read::batch_read_array for one leaf column."""
        return self.decode_columns([column], out=out)[0]

    def decode_pages(self, pages, out="host", raise_on_page_error=True):
        """column_iter_to_arrays(...).next(): one array per page; `pages` are one-page Columns."""
        return self.decode_columns(pages, out=out, raise_on_page_error=raise_on_page_error, per_page=True)
