"""configs[2] utf8 column alone (Dict pages): per-call stats; run with SB_TIMING=1 for the host phases, or under
`ncu --metrics gpu__time_duration.sum` for the size pass / decode pass split.  usage: python tools/c3_prof.py [rows] [lib.so]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb
from strawboat_b200 import workloads as wl
import bench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = sb.Context(0)
rng = np.random.default_rng(42)
v, val = wl.dict_strings(rng, rows, 1000, 0.4, False)
enc, est = bench.gpu_write_columns(ctx, [("s0_utf8", sb.BINARY, v, val)], 42)
dev, keep = bench.to_device_cols(torch, sb, enc)
print("page bytes: min %d max %d" % (min(m[0] for m in enc[0]["metas"]), max(m[0] for m in enc[0]["metas"])) if isinstance(enc[0]["metas"][0], tuple) else "")
for i in range(4):
    out = ctx.decode_columns(dev, out="device")
    st = ctx.last_stats()
    out[0]._group.release()
    print("call %d device_us %.1f main_us %.1f host_us %.1f launches %d" % (i, st["device_ms"] * 1e3, st["main_kernel_ms"] * 1e3, st["host_ms"] * 1e3, st["kernel_launches"]), flush=True)
