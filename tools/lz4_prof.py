"""diagnostic (build with `make -C strawboat_b200/csrc EXTRA=-DSB_LZ4_PROF`): cycles per phase of the LZ4 mover."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import strawboat_b200 as sb
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "c7"
ctx = sb.Context(0, stream=torch.cuda.current_stream())
cols, _ = bench.build_workload(rows, 42, ctx) if (len(sys.argv) <= 3 or sys.argv[3] == "ours") else bench.build_workload(rows, 42)
lib = sb._lib
out = (C.c_ulonglong * 32)()
names = ["wait", "load+parse", "scan+validate", "literals", "chain", "far", "rounds", "flush", "publish", "periodic_seqs", "batches", "seqs"]
for c in cols:
    if not c["name"].startswith(which):
        continue
    td = torch.from_numpy(c["data"].copy()).cuda()
    col = sb.Column(c["type"], c["nullable"], td, c["metas"])
    ctx.decode_columns([col], out="device")[0]._group.release()
    lib.sb_debug_lz4_prof(out, 1)
    ctx.decode_columns([col], out="device")[0]._group.release()
    st = ctx.last_stats()
    lib.sb_debug_lz4_prof(out, 1)
    v = list(out)
    nb, ns = max(1, v[10]), max(1, v[11])
    print(c["name"], "lz4_kernel_ms", st["lz4_kernel_ms"], "batches", nb, "seqs", ns, "seqs/batch %.1f" % (ns / nb))
    tot = max(1, sum(v[:9]))
    for n, x in zip(names[:9], v[:9]):
        print(f"  {n:14s} {x / nb:9.0f} cyc/batch  {x / ns:7.1f} cyc/seq  {100 * x / tot:5.1f}%")
    print(f"  total          {tot / nb:9.0f} cyc/batch  {tot / ns:7.1f} cyc/seq   sequences on the periodic path: {v[9]} of {ns}")
