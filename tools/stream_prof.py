"""sb_decode_pages (one array per page, the streaming reader) against sb_decode_columns on the same pages:
configs[1], device-resident pages, device outputs.  usage: python tools/stream_prof.py [rows]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb
from strawboat_b200 import workloads as wl
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = sb.Context(0)
cols = wl.config2(rows, 42)
enc = bench.oracle_write_columns(cols, 42, 8)
dev, keep = bench.to_device_cols(torch, sb, enc)
pages = []
for c, td in zip(enc, keep):
    pos = 0
    for ln, nv in c["metas"]:
        pages.append(sb.Column(c["type"], c["nullable"], td[pos:pos + ln], [(ln, nv)]))
        pos += ln


def run(fn, n=5):
    best = None
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        t1 = time.perf_counter()
        st = ctx.last_stats()
        for o in out[:1]:
            o._group.release()
        st["wall_ms"] = (t1 - t0) * 1e3
        best = st if best is None or st["wall_ms"] < best["wall_ms"] else best
    return best


a = run(lambda: ctx.decode_columns(dev, out="device"))
b = run(lambda: ctx.decode_pages(pages, out="device"))
for tag, st in (("decode_columns (8 columns)", a), ("decode_pages (%d one-page entries)" % len(pages), b)):
    print("%-42s wall_ms %8.3f host_ms %8.3f device_ms %7.3f -> %.1f GB/s on wall" % (tag, st["wall_ms"], st["host_ms"], st["device_ms"], st["bytes_out"] / st["wall_ms"] / 1e6))
