"""diagnostic (build with `make -B -C strawboat_b200/csrc EXTRA=-DSB_LZ4_PROF`): cycles per phase of the LZ4 workers (thread 0 of the team).
usage: lz4_prof.py [rows] [column prefix] [ours|oracle]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import strawboat_b200 as sb
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
which = sys.argv[2] if len(sys.argv) > 2 else "c7"
ctx = sb.Context(0, stream=torch.cuda.current_stream())
from strawboat_b200 import workloads as wl
cols = bench.gpu_write_columns(ctx, wl.config2(rows, 42), 42)[0] if (len(sys.argv) > 3 and sys.argv[3] == "ours") else bench.oracle_write_columns(wl.config2(rows, 42), 42, 8)
lib = sb._lib
out = (C.c_ulonglong * 32)()
names = ["wait FULL", "parse", "scan+validate", "far+lit+match", "long runs", "flush", "dependent", "room flush", "-", "dependent_seqs", "batches", "seqs"]
for c in cols:
    if not c["name"].startswith(which):
        continue
    td = torch.from_numpy(np.array(c["data"], copy=True)).cuda()
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
    print(f"  total          {tot / nb:9.0f} cyc/batch  {tot / ns:7.1f} cyc/seq   sequences on the ordered (dependent) path: {v[9]} of {ns}")
