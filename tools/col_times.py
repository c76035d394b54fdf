"""per-column device time of the decode kernel on config 2 (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import strawboat_b200 as sb

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cols, _ = bench.build_workload(rows, 42)
ctx = sb.Context(0, stream=torch.cuda.current_stream())
for c in cols:
    if only and not any(c['name'].startswith(o) for o in only):
        continue
    td = torch.from_numpy(c["data"].copy()).cuda()
    col = sb.Column(c["type"], c["nullable"], td, c["metas"])
    best, best_main = 1e9, 1e9
    for _ in range(reps):
        out = ctx.decode_columns([col], out="device")
        st = ctx.last_stats()
        out[0]._group.release()
        best = min(best, st["device_ms"])
        best_main = min(best_main, st["main_kernel_ms"])
    ob = rows * np.dtype(sb.NP_OF[c["type"]]).itemsize
    print(f"{c['name']:18s} in={len(c['data'])/1e6:8.2f}MB out={ob/1e6:7.1f}MB call={best*1e3:9.1f}us main_kernel={best_main*1e3:8.1f}us  alg={(len(c['data'])+ob)/best/1e6:8.1f} GB/s  {c['codecs']}")
