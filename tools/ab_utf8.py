"""A/B of library builds on plain utf8 pages (16 columns x 1 M rows and 1 x 10 M rows, device resident):
python tools/ab_utf8.py libA.so libB.so ...  (paths relative to strawboat_b200/csrc)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb
from strawboat_b200 import _capi

ctx0 = sb.Context(0)
rng = np.random.default_rng(42)
cases = []
for rows, ncols in ((1_000_000, 16), (10_000_000, 1)):
    lens = rng.integers(4, 16, rows)
    off = np.zeros(rows + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    data = rng.integers(48, 123, int(off[-1]), dtype=np.uint8)
    enc = ctx0.encode_columns([sb.LeafArray(sb.BINARY, (off, data))], sb.write_options(sb.C_NONE, None, 8192))[0]
    tens = [torch.frombuffer(bytearray(enc.data), dtype=torch.uint8).cuda() for _ in range(ncols)]
    cases.append(("utf8 %dx%dM" % (ncols, rows // 1_000_000), [sb.Column(sb.BINARY, False, t, enc.metas) for t in tens], tens))
for name in sys.argv[1:]:
    L = C.CDLL(os.path.join(ROOT, "strawboat_b200", "csrc", name))
    L.sb_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
    L.sb_decode_columns.argtypes = [C.c_void_p, C.POINTER(_capi.ColumnIn), C.c_uint64, C.c_int32, C.POINTER(_capi.ColumnOut)]
    L.sb_release_columns.argtypes = [C.c_void_p, C.POINTER(_capi.ColumnOut), C.c_uint64]
    L.sb_last_stats.argtypes = [C.c_void_p, C.POINTER(_capi.Stats)]
    h = C.c_void_p()
    assert L.sb_ctx_create(0, C.byref(h)) == 0
    res = []
    for tag, cols, _ in cases:
        ins, keep = ctx0._marshal(cols)
        n = len(cols)
        best = None
        for _ in range(10):
            outs = (_capi.ColumnOut * n)()
            assert L.sb_decode_columns(h, ins, n, 1, outs) == 0
            st = _capi.Stats()
            L.sb_last_stats(h, C.byref(st))
            L.sb_release_columns(h, outs, n)
            cur = (st.main_kernel_ms, st.device_ms)
            best = cur if best is None or cur[0] < best[0] else best
        res.append("%s kernel %.1f call %.1f us" % (tag, best[0] * 1e3, best[1] * 1e3))
    print(name, " | ".join(res))
