"""configs[1] decode, one call: where the step goes.  device_ms = ev0..ev1 inside the call (classify + kernels),
host_ms = wall time of the call; the bench's step adds the caller's loop.  usage: python tools/c2_prof.py [rows]"""
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
for _ in range(3):
    out = ctx.decode_columns(dev, out="device"); out[0]._group.release()
best = None
for _ in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ctx.decode_columns(dev, out="device")
    t1 = time.perf_counter()
    st = ctx.last_stats()
    out[0]._group.release()
    st["wall_ms"] = (t1 - t0) * 1e3
    best = st if best is None or st["wall_ms"] < best["wall_ms"] else best
print("pages %d wall_ms %.3f host_ms %.3f device_ms %.3f main %.3f lz4 %.3f launches %d out %.1f MB -> %.1f GB/s on wall, %.1f on device" % (
    best["pages"], best["wall_ms"], best["host_ms"], best["device_ms"], best["main_kernel_ms"], best["lz4_kernel_ms"], best["kernel_launches"],
    best["bytes_out"] / 1e6, best["bytes_out"] / best["wall_ms"] / 1e6, best["bytes_out"] / best["device_ms"] / 1e6))
