"""Per-call breakdown of config 3 (strings) and config 4 (nested) decode: device / main kernel / LZ4 kernel / host time
per column.  usage: python tools/c34_prof.py [rows3 rows4]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb
from strawboat_b200 import workloads as wl
import bench

rows3 = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rows4 = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
ctx = sb.Context(0)


def show(tag, st):
    print("%-22s pages %5d device_us %8.1f main_us %8.1f lz4_us %8.1f host_us %8.1f launches %d  out %.1f MB -> %.1f GB/s  codecs %s" % (
        tag, st["pages"], st["device_ms"] * 1e3, st["main_kernel_ms"] * 1e3, st["lz4_kernel_ms"] * 1e3, st["host_ms"] * 1e3, st["kernel_launches"],
        st["bytes_out"] / 1e6, st["bytes_out"] / st["device_ms"] / 1e6, st["codec_pages"]))


cols = wl.config3(rows3, 42)
enc, est = bench.gpu_write_columns(ctx, cols, 42)
print("config3 encode device_ms %.3f" % est["device_ms"])
dev, keep = bench.to_device_cols(torch, sb, enc)
for c, d in zip(enc, dev):
    show("c3 " + c["name"], bench.timed_decode(ctx, [d]))
show("c3 both", bench.timed_decode(ctx, dev))

rep, de, row_start, leaves = wl.config4(rows4, 7)
nested = wl.CONFIG4_NESTED
enc = []
for name, t, v, val in leaves:
    arr = sb.LeafArray(t, v, validity=val, nullable=True, nested=nested, rep_levels=rep, def_levels=de, rows=rows4)
    e = ctx.encode_columns([arr], sb.write_options(sb.C_LZ4, 2.0, 8192, seed=42))[0]
    e = ctx.encode_columns([arr], sb.write_options(sb.C_LZ4, 2.0, 8192, seed=42))[0]
    print("config4 encode %s device_ms %.3f" % (name, ctx.last_stats()["device_ms"]))
    enc.append({"name": name, "type": t, "nullable": True, "data": np.frombuffer(e.data, dtype=np.uint8), "metas": e.metas, "values": v, "validity": val})
dev, keep = bench.to_device_cols(torch, sb, enc, nested)
for c, d in zip(enc, dev):
    show("c4 " + c["name"], bench.timed_decode(ctx, [d]))
show("c4 all", bench.timed_decode(ctx, dev))
