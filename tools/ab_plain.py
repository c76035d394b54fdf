"""A/B of two builds of the library on one box: plain i64 pages, 16 columns x 1 M rows, device resident.
usage: python tools/ab_plain.py libA.so libB.so   (paths relative to strawboat_b200/csrc)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb
from strawboat_b200 import _capi

ctx0 = sb.Context(0)
rng = np.random.default_rng(1)
v = rng.integers(-2**63, 2**63 - 1, 1_000_000, dtype=np.int64)
val = (rng.random(1_000_000) > 0.1) if os.environ.get("NULLABLE") else None
enc = ctx0.encode_columns([sb.LeafArray(sb.I64, v, validity=val)], sb.write_options(sb.C_NONE, None, 8192))[0]
tens = [torch.frombuffer(bytearray(enc.data), dtype=torch.uint8).cuda() for _ in range(16)]
cols = [sb.Column(sb.I64, val is not None, t, enc.metas) for t in tens]
ins, keep = ctx0._marshal(cols)
for name in sys.argv[1:]:
    L = C.CDLL(os.path.join(ROOT, "strawboat_b200", "csrc", name))
    L.sb_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    L.sb_decode_columns.argtypes = [C.c_void_p, C.POINTER(_capi.ColumnIn), C.c_uint64, C.c_int32, C.POINTER(_capi.ColumnOut)]
    L.sb_release_columns.argtypes = [C.c_void_p, C.POINTER(_capi.ColumnOut), C.c_uint64]
    L.sb_last_stats.argtypes = [C.c_void_p, C.POINTER(_capi.Stats)]
    h = C.c_void_p()
    assert L.sb_ctx_create(0, C.byref(h)) == 0
    best = None
    for _ in range(30):
        outs = (_capi.ColumnOut * 16)()
        assert L.sb_decode_columns(h, ins, 16, 1, outs) == 0
        st = _capi.Stats()
        L.sb_last_stats(h, C.byref(st))
        L.sb_release_columns(h, outs, 16)
        cur = (st.device_ms, st.main_kernel_ms, st.lz4_kernel_ms, st.host_ms)
        best = cur if best is None or cur[0] < best[0] else best
    print(name, "device_us %.1f main_kernel_us %.1f lz4_us %.1f host_us %.1f" % tuple(x * 1e3 for x in best))
