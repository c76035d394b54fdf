"""north-star check: page decode of plain (codec None) i64 / f64 / utf8 columns, device resident.
Pages are written by this library's encoder (default_compression None, adaptive off)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import strawboat_b200 as sb


def fast_strings(rng, n):
    """random printable strings of 4..15 bytes (cheap to generate: the GPU box's CPU time is billed too)"""
    lens = rng.integers(4, 16, n)
    off = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(lens, out=off[1:])
    return off, rng.integers(48, 123, int(off[-1]), dtype=np.uint8)


PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ctx = sb.Context(0, stream=torch.cuda.current_stream())
rng = np.random.default_rng(42)
out = []
for rows, ncols in ((1_000_000, 1), (1_000_000, 16), (10_000_000, 1)):
    cases = [("i64", sb.I64, rng.integers(-2**63, 2**63 - 1, rows, dtype=np.int64), None),
             ("f64", sb.F64, rng.standard_normal(rows), None),
             ("i64 nullable", sb.I64, rng.integers(-2**63, 2**63 - 1, rows, dtype=np.int64), rng.random(rows) > 0.1)]
    cases.append(("utf8", sb.BINARY, fast_strings(rng, rows), None))
    for name, t, v, val in cases:
        wo = sb.write_options(sb.C_NONE, None, 8192)
        enc = ctx.encode_columns([sb.LeafArray(t, v, validity=val)], wo)[0]
        cols = []
        for k in range(ncols):  # distinct device copies: no help from L2 between the columns of a call
            td = torch.frombuffer(bytearray(enc.data), dtype=torch.uint8).cuda()
            cols.append(sb.Column(t, val is not None, td, enc.metas))
        best = None
        for _ in range(6):
            res = ctx.decode_columns(cols, out="device")
            st = ctx.last_stats()
            res[0]._group.release()
            best = st if best is None or st["main_kernel_ms"] < best["main_kernel_ms"] else best
        alg = best["bytes_in"] + best["bytes_out"]
        r = {"case": name, "rows": rows, "columns": ncols, "pages": len(enc.metas) * ncols, "bytes_in": best["bytes_in"], "bytes_out": best["bytes_out"],
             "main_kernel_us": round(best["main_kernel_ms"] * 1e3, 1), "call_device_us": round(best["device_ms"] * 1e3, 1),
             "kernel_gbs": round(alg / best["main_kernel_ms"] / 1e6, 1), "frac_of_peak": round(alg / best["main_kernel_ms"] / 1e6 / PEAK, 3),
             "launches": best["kernel_launches"]}
        out.append(r)
        print(r, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "plain.json"), "w"), indent=1)
