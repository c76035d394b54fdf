"""A/B of library builds on single columns: python tools/ab_cols.py libA.so libB.so ...  (paths relative to strawboat_b200/csrc).
Decodes c4 (Freq with LZ4 exceptions) of configs[1] and the three leaves of configs[3] device to device, prints the best device time."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb
from strawboat_b200 import _capi, workloads as wl
import bench

ctx0 = sb.Context(0)
cols2 = wl.config2(10_000_000, 42)
enc2 = bench.oracle_write_columns(cols2, 42, 8)
dev2, keep2 = bench.to_device_cols(torch, sb, enc2)
rows4 = 4_000_000
rep, de, row_start, leaves = wl.config4(rows4, 7)
nested = wl.CONFIG4_NESTED
enc4 = []
for name, t, v, val in leaves:
    arr = sb.LeafArray(t, v, validity=val, nullable=True, nested=nested, rep_levels=rep, def_levels=de, rows=rows4)
    e = ctx0.encode_columns([arr], sb.write_options(sb.C_LZ4, 2.0, 8192, seed=42))[0]
    enc4.append({"name": name, "type": t, "nullable": True, "data": np.frombuffer(e.data, dtype=np.uint8), "metas": e.metas})
dev4, keep4 = bench.to_device_cols(torch, sb, enc4, nested)
cases = [("c2 all 8", dev2)] + [(e["name"][:7], [d]) for e, d in zip(enc2, dev2) if e["name"][:2] in ("c0", "c1", "c4", "c5", "c6")] + [("c4n " + e["name"], [d]) for e, d in zip(enc4, dev4)] + [("c4n all", dev4)]
for name in sys.argv[1:]:
    L = C.CDLL(os.path.join(ROOT, "strawboat_b200", "csrc", name))
    L.sb_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    L.sb_decode_columns.argtypes = [C.c_void_p, C.POINTER(_capi.ColumnIn), C.c_uint64, C.c_int32, C.POINTER(_capi.ColumnOut)]
    L.sb_release_columns.argtypes = [C.c_void_p, C.POINTER(_capi.ColumnOut), C.c_uint64]
    L.sb_last_stats.argtypes = [C.c_void_p, C.POINTER(_capi.Stats)]
    h = C.c_void_p()
    assert L.sb_ctx_create(0, C.byref(h)) == 0
    res = []
    for tag, cols in cases:
        ins, keep = ctx0._marshal(cols)
        n = len(cols)
        best = None
        for _ in range(6):
            outs = (_capi.ColumnOut * n)()
            rc = L.sb_decode_columns(h, ins, n, 1, outs)
            assert rc == 0, rc
            st = _capi.Stats()
            L.sb_last_stats(h, C.byref(st))
            L.sb_release_columns(h, outs, n)
            best = st.device_ms if best is None or st.device_ms < best else best
        res.append("%s %.0f" % (tag, best * 1e3))
    print(name, "device_us:", " | ".join(res))
