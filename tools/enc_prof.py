"""Where does sb_encode_columns' time go?  configs[1] columns, one column at a time and all together,
with default LZ4 (adaptive), default None (adaptive: statistics + sampling only) and LZ4 with the chooser off.
usage: python tools/enc_prof.py [rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import strawboat_b200 as sb
from strawboat_b200 import workloads

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = sb.Context(0)
cols = workloads.config2(rows, 42)


def run(arrays, wo):
    best = None
    for _ in range(3):
        enc = ctx.encode_columns(arrays, wo)
        st = ctx.last_stats()
        best = st if best is None or st["device_ms"] < best["device_ms"] else best
    return best, sum(len(e.data) for e in enc)


modes = [("lz4+adaptive", sb.write_options(sb.C_LZ4, 2.0, 8192, seed=42)), ("none+adaptive", sb.write_options(sb.C_NONE, 2.0, 8192, seed=42)),
         ("lz4 only", sb.write_options(sb.C_LZ4, None, 8192, seed=42)), ("none only", sb.write_options(sb.C_NONE, None, 8192, seed=42))]
for name, t, v, val in cols:
    arr = [sb.LeafArray(t, v, validity=val)]
    for mn, wo in modes:
        st, nb = run(arr, wo)
        print("%-6s %-14s device_ms %7.3f host_ms %7.3f  in %6.1f MB out %6.1f MB  %6.1f GB/s  hist %s" % (
            name, mn, st["device_ms"], st["host_ms"], st["bytes_in"] / 1e6, nb / 1e6, st["bytes_in"] / st["device_ms"] / 1e6,
            st["codec_pages"]))
arrays = [sb.LeafArray(t, v, validity=val) for (_, t, v, val) in cols]
for mn, wo in modes:
    st, nb = run(arrays, wo)
    print("ALL    %-14s device_ms %7.3f host_ms %7.3f  in %6.1f MB out %6.1f MB  %6.1f GB/s" % (
        mn, st["device_ms"], st["host_ms"], st["bytes_in"] / 1e6, nb / 1e6, st["bytes_in"] / st["device_ms"] / 1e6))
