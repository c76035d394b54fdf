#!/usr/bin/env python
"""bench.py -- decoded GB/s of the strawboat page-decode hot path on B200 (BASELINE.json metric).

A "step" = one batched decode of the whole workload: configs[1] of BASELINE.json -- 8 primitive
columns (3 x i32, 3 x i64, 2 x f64) x 10 M rows, 8192 rows/page, default LZ4,
default_compress_ratio 2.0 (adaptive), one distribution per codec.  The pages are
written by the ORACLE writer (the reference's chooser + liblz4): what a reference-written file holds.

  value : Arrow bytes out / device time, page bytes already resident in HBM
          (plan upload + every kernel of sb_decode_columns inside the timed region)
  e2e   : same call with HOST page bytes in pinned memory and HOST Arrow buffers out
          (H2D of the pages + D2H of the decoded buffers inside the timed region)
  roofline : the dominant decode kernel's algorithmic bytes (sum PageMeta.length read + Arrow bytes
          written) / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the oracle (C++ restatement of the Rust reference; the Rust crate
          cannot be built in this image) on the host cores, SAME workload (all rows, all columns).
  extras (N = 1, rank 0; --no-extras skips them): `own_pages` (same workload, pages written by this
          library's GPU encoder), `encode` (+ roofline), `north_star` (plain i64 / f64 / utf8 page decode
          at 1 x 1 M, 16 x 1 M, 1 x 10 M rows: fraction of the HBM peak on call device time and on kernel
          time), `config3` (strings), `config4` (nested).

N > 1: every rank decodes its own 8-column x 10 M-row partition (weak scaling, no data-path
collective: pages are independent); value = total bytes of all ranks / max-over-ranks time.  The
`multi_gpu_encode` object is configs[4]: one mixed-type table sharded by leaf column, encoded on every
rank, gathered to the writer rank with NCCL, framed, re-read and checked.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PAGE_ROWS = 8192


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def oracle():
    """The CPU oracle: used here ONLY to prepare encoded input pages (untimed setup) and for
    the cpu_baseline / --impl reference legs."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sbo
    return sbo


# ------------------------------------------------------------------------------- workloads
def oracle_write_columns(cols, seed, threads):
    """Pages of every column written by the oracle writer (reference chooser, liblz4), page loop in C++."""
    from concurrent.futures import ThreadPoolExecutor
    sbo = oracle()

    def enc(c):
        name, t, v, val = c
        opts = sbo.make_opts(sbo.C_LZ4, ratio=2.0, seed=seed)
        body, metas = sbo.write_column(t, v, val, opts=opts, page_rows=PAGE_ROWS)
        return {"name": name, "type": t, "nullable": val is not None, "data": np.frombuffer(body, dtype=np.uint8),
                "metas": metas, "values": v, "validity": val}

    with ThreadPoolExecutor(max(1, threads)) as ex:
        return list(ex.map(enc, cols))


def gpu_write_columns(ctx, cols, seed, default=None, ratio=2.0):
    """Pages written by this library's encoder (sb_encode_columns).  Returns (columns, encode stats)."""
    import strawboat_b200 as sb
    wo = sb.write_options(sb.C_LZ4 if default is None else default, ratio, PAGE_ROWS, seed=seed)
    arrays = [sb.LeafArray(t, v, validity=val) for (_, t, v, val) in cols]
    best = None
    for _ in range(2):  # second call: pools warm
        enc = ctx.encode_columns(arrays, wo)
        st = ctx.last_stats()
        best = st if best is None or st["device_ms"] < best["device_ms"] else best
    out = []
    for (name, t, v, val), e in zip(cols, enc):
        out.append({"name": name, "type": t, "nullable": val is not None, "data": np.frombuffer(e.data, dtype=np.uint8),
                    "metas": e.metas, "values": v, "validity": val})
    return out, best


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region.  NVML in-process (nvidia_ml_py) when it loads: spawning
    `nvidia-smi` every 50 ms from a process with gigabytes of pinned memory stalled the reader threads of the e2e leg
    for 10-20 ms at a time (steps of 26-30 ms among 10 ms ones); the subprocess form is the fallback."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.reasons = index, [], False, set()
        self.max_mhz = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.max_mhz = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = int(get(self.handle))
        for nm, mask in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
            if bits & mask:
                self.reasons.add(nm)

    def sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(",")
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for nm, v in zip(self.NAMES, out[2:]):
            if "Active" in v and "Not" not in v:
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                self.sample_nvml() if self.nvml else self.sample_smi()
            except Exception:
                pass
            time.sleep(0.005 if self.nvml else 0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": "nvml (in-process)" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------- CPU legs
def cpu_tasks(cols, threads):
    """(leaf, body, metas) decode tasks: one per column, split further by page ranges when the box has more
    cores than columns (pages are independent, so a reference caller can do the same)."""
    sbo = oracle()
    parts = max(1, threads // max(1, len(cols)))
    tasks, out_bytes = [], 0
    for c in cols:
        metas = c["metas"]
        pos = np.concatenate([[0], np.cumsum([m[0] for m in metas])])
        step = (len(metas) + parts - 1) // parts
        for a in range(0, len(metas), max(1, step)):
            b = min(len(metas), a + step)
            tasks.append((sbo.make_leaf(c["type"], c["nullable"]), c["data"][pos[a]:pos[b]], metas[a:b]))
        out_bytes += sum(m[1] for m in metas) * sbo.WIDTH[c["type"]]
    return tasks, out_bytes


def cpu_decode_pass(tasks, threads, pool):
    sbo = oracle()
    t0 = time.perf_counter()
    list(pool.map(lambda j: sbo.read_column_body(j[0], j[1], j[2], fetch=False)["length"], tasks))
    return time.perf_counter() - t0


# ------------------------------------------------------------------------------- device helpers
def timed_decode(ctx, dev_cols, reps=6, key="device_ms"):
    best = None
    for _ in range(reps):
        out = ctx.decode_columns(dev_cols, out="device")
        st = ctx.last_stats()
        out[0]._group.release()
        best = st if best is None or st[key] < best[key] else best
    return best


def to_device_cols(torch, sb, cols, nested=None, copies=1):
    dev, keep = [], []
    for _ in range(copies):
        for c in cols:
            td = torch.from_numpy(np.array(c["data"], copy=True)).cuda()
            keep.append(td)
            dev.append(sb.Column(c["type"], c["nullable"], td, c["metas"], nested))
    return dev, keep


def check_roundtrip(ctx, sb, cols, nested=None):
    """decode(pages) == the source arrays, bit for bit (valid slots only where nulls carry no value)"""
    host = [sb.Column(c["type"], c["nullable"], c["data"], c["metas"], nested) for c in cols]
    for c, r in zip(cols, ctx.decode_columns(host, out="host")):
        v, val = c["values"], c["validity"]
        if isinstance(v, tuple) and val is not None:
            # what a null slot decodes to is codec dependent (Dict / Freq substitute a neighbour): compare the valid rows
            off = np.asarray(v[0], dtype=np.int64)
            lens, glen = np.diff(off)[val], np.diff(r.offsets.astype(np.int64))[val]
            assert np.array_equal(lens, glen), c["name"]
            tot = int(lens.sum())
            row = np.repeat(np.arange(len(lens)), lens)
            pos = np.arange(tot) - np.repeat(np.cumsum(lens) - lens, lens)
            assert np.array_equal(r.values[r.offsets[:-1].astype(np.int64)[val][row] + pos], np.asarray(v[1])[off[:-1][val][row] + pos]), c["name"]
        elif isinstance(v, tuple):
            assert np.array_equal(r.offsets, np.asarray(v[0]) - v[0][0]), c["name"]
            assert np.array_equal(r.values, np.asarray(v[1])[int(v[0][0]):int(v[0][-1])]), c["name"]
        elif val is None:
            assert np.array_equal(r.values.view(np.uint8), np.ascontiguousarray(v).view(np.uint8)), c["name"]
        else:
            assert np.array_equal(r.values[val], np.asarray(v)[val]), c["name"]
        if val is not None and nested is None:
            assert np.array_equal(np.unpackbits(r.validity, bitorder="little")[:len(val)].astype(bool), val), c["name"]


def section_north_star(ctx, torch, sb, peak):
    """plain (codec None) page decode, the north-star cases.  frac_call: algorithmic bytes / device time of the
    whole sb_decode_columns call (table upload, classify, kernels); frac_kernel: / sb_decode_kernel alone."""
    from strawboat_b200 import workloads as wl
    rng = np.random.default_rng(42)
    out = []
    for rows, ncols in ((1_000_000, 1), (1_000_000, 16), (10_000_000, 1)):
        cases = [("i64", sb.I64, rng.integers(-2**63, 2**63 - 1, rows, dtype=np.int64), None),
                 ("f64", sb.F64, rng.standard_normal(rows), None),
                 ("i64 nullable", sb.I64, rng.integers(-2**63, 2**63 - 1, rows, dtype=np.int64), rng.random(rows) > 0.1),
                 ("utf8", sb.BINARY, wl.plain_strings(rng, rows), None)]
        for name, t, v, val in cases:
            enc = ctx.encode_columns([sb.LeafArray(t, v, validity=val)], sb.write_options(sb.C_NONE, None, PAGE_ROWS))[0]
            c = {"name": name, "type": t, "nullable": val is not None, "data": np.frombuffer(enc.data, dtype=np.uint8), "metas": enc.metas,
                 "values": v, "validity": val}
            if ncols == 1:
                check_roundtrip(ctx, sb, [c])
            dev, keep = to_device_cols(torch, sb, [c], copies=ncols)  # distinct device copies: no L2 help between columns
            bc = timed_decode(ctx, dev, key="device_ms", reps=10)
            # the decode kernel that did the work (the light instantiation only runs with SB_SPLIT=1)
            bk = {"main_kernel_ms": max(bc["main_kernel_ms"], bc["light_kernel_ms"])}
            alg = bc["bytes_in"] + bc["bytes_out"]
            out.append({"case": name, "rows": rows, "columns": ncols, "pages": len(enc.metas) * ncols, "algorithmic_bytes": alg,
                        "call_device_us": round(bc["device_ms"] * 1e3, 1), "kernel_us": round(bk["main_kernel_ms"] * 1e3, 1),
                        "call_gbs": round(alg / bc["device_ms"] / 1e6, 1), "kernel_gbs": round(alg / bk["main_kernel_ms"] / 1e6, 1),
                        "frac_call": round(alg / bc["device_ms"] / 1e6 / peak, 3), "frac_kernel": round(alg / bk["main_kernel_ms"] / 1e6 / peak, 3),
                        "launches": bc["kernel_launches"]})
            del dev, keep
    return {"target": "north_star: >= 0.60 of the HBM peak on 1 M-row i64 / f64 / utf8 page decode", "peak_gbs": peak,
            "pages_written_by": "strawboat_b200 GPU encoder, default_compression None, adaptive off", "cases": out}


def section_config3(ctx, torch, sb, peak, rows):
    """configs[2]: nullable Utf8 + LargeBinary, decimal strings of integers(0, 1000), 40 % nulls, adaptive on"""
    from strawboat_b200 import workloads as wl
    rng = np.random.default_rng(42)
    cols = []
    for name, t, large in (("s0_utf8", sb.BINARY, False), ("s1_large_binary", sb.LARGE_BINARY, True)):
        v, val = wl.dict_strings(rng, rows, 1000, 0.4, large)
        cols.append((name, t, v, val))
    enc, est = gpu_write_columns(ctx, cols, 42)
    check_roundtrip(ctx, sb, enc)
    dev, keep = to_device_cols(torch, sb, enc)
    res = {"workload": "configs[2]: nullable Utf8 + LargeBinary x %d rows, uniq 1000, 40 %% nulls, adaptive, pages by the GPU encoder" % rows,
           "columns": []}
    for c, d in zip(enc, dev):
        st = timed_decode(ctx, [d])
        res["columns"].append({"column": c["name"], "codec_pages": st["codec_pages"], "bytes_in": st["bytes_in"], "bytes_out": st["bytes_out"],
                               "device_us": round(st["device_ms"] * 1e3, 1), "decoded_gbs": round(st["bytes_out"] / st["device_ms"] / 1e6, 1)})
    st = timed_decode(ctx, dev)
    alg = st["bytes_in"] + st["bytes_out"]
    res.update({"value": round(st["bytes_out"] / st["device_ms"] / 1e6, 1), "unit": "GB/s decoded, both columns in one call", "device_us": round(st["device_ms"] * 1e3, 1),
                "roofline": {"bound": "hbm", "achieved": round(alg / st["device_ms"] / 1e6, 1), "peak": peak, "unit": "GB/s", "frac": round(alg / st["device_ms"] / 1e6 / peak, 4),
                             "algorithmic_bytes": alg},
                "encode": {"value": round(est["bytes_in"] / est["device_ms"] / 1e6, 2), "unit": "GB/s (Arrow bytes in / device time)", "device_ms": round(est["device_ms"], 3)}})
    return res


def section_config4(ctx, torch, sb, peak, rows):
    """configs[3]: List<Struct<i64, f64, utf8>>, pages with rep / def level streams, written by the GPU encoder"""
    from strawboat_b200 import workloads as wl
    rep, de, row_start, leaves = wl.config4(rows, 7)
    nested = wl.CONFIG4_NESTED
    enc, enc_ms, enc_in = [], 0.0, 0
    for name, t, v, val in leaves:
        arr = sb.LeafArray(t, v, validity=val, nullable=True, nested=nested, rep_levels=rep, def_levels=de, rows=rows)
        best = None
        for _ in range(2):
            e = ctx.encode_columns([arr], sb.write_options(sb.C_LZ4, 2.0, PAGE_ROWS, seed=42))[0]
            st = ctx.last_stats()
            best = st if best is None or st["device_ms"] < best["device_ms"] else best
        enc_ms += best["device_ms"]
        enc_in += best["bytes_in"]
        enc.append({"name": name, "type": t, "nullable": True, "data": np.frombuffer(e.data, dtype=np.uint8), "metas": e.metas, "values": v, "validity": val})
    check_roundtrip(ctx, sb, enc, nested)
    dev, keep = to_device_cols(torch, sb, enc, nested)
    res = {"workload": "configs[3]: List<Struct<a:Int64,b:Float64,c:Utf8>> x %d rows (%d level entries per leaf), 8192 rows/page, pages by the GPU encoder" % (rows, len(rep)),
           "columns": []}
    for c, d in zip(enc, dev):
        st = timed_decode(ctx, [d])
        res["columns"].append({"column": c["name"], "codec_pages": st["codec_pages"], "bytes_in": st["bytes_in"], "bytes_out": st["bytes_out"],
                               "device_us": round(st["device_ms"] * 1e3, 1), "decoded_gbs": round(st["bytes_out"] / st["device_ms"] / 1e6, 1)})
    st = timed_decode(ctx, dev)
    alg = st["bytes_in"] + st["bytes_out"]
    res.update({"value": round(st["bytes_out"] / st["device_ms"] / 1e6, 1), "unit": "GB/s decoded (leaf buffers + NestedState), three leaves in one call",
                "device_us": round(st["device_ms"] * 1e3, 1),
                "roofline": {"bound": "hbm", "achieved": round(alg / st["device_ms"] / 1e6, 1), "peak": peak, "unit": "GB/s", "frac": round(alg / st["device_ms"] / 1e6 / peak, 4),
                             "algorithmic_bytes": alg},
                "encode": {"value": round(enc_in / enc_ms / 1e6, 2), "unit": "GB/s (Arrow + level bytes in / device time)", "device_ms": round(enc_ms, 3)}})
    return res


def section_pipelined(ctx, torch, sb, dev_cols, bytes_out, device, steps, depth=2):
    ctxs = [ctx] + [sb.Context(device) for _ in range(depth - 1)]
    def run(k):
        pend = []
        for i in range(k):
            if len(pend) == depth:
                res = pend.pop(0).wait()
                res[0]._group.release()
            pend.append(ctxs[i % depth].decode_columns_async(dev_cols, out="device"))
        for h in pend:
            res = h.wait()
            res[0]._group.release()
    run(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    for c in ctxs[1:]:
        c.close()
    ms = max(wall, e0.elapsed_time(e1)) / steps
    return {"value": round(bytes_out / ms / 1e6, 1), "unit": "GB/s decoded", "ms_per_step": round(ms, 4), "in_flight": depth, "steps": steps,
            "api": "sb_decode_columns_async + sb_decode_wait on %d contexts, device-resident pages and outputs" % depth}


def kernel_sources_hash():
    h = hashlib.sha256()
    d = os.path.join(ROOT, "strawboat_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from an `ncu --set full` capture of this
    command (tools/ncu_traffic.py writes the file).  Only used when the capture was taken on the kernel sources
    that are running now; otherwise null (a stale figure is worse than none)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        if t.get("kernel_sources_sha") != kernel_sources_hash():
            return None, None
        k = t["kernels"].get(kernel)
        return (int(k["dram_bytes"]), "profiles/r2_ncu_traffic.json (ncu --set full of `bench.py --steps 2 --warmup 3 --no-extras`, same kernel sources)") if k else (None, None)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-columns", action="store_true", help="skip the per-column diagnostic table")
    ap.add_argument("--no-extras", action="store_true", help="skip own_pages / encode / north_star / config3 / config4 / multi_gpu_encode")
    ap.add_argument("--e2e-threads", type=int, default=8, help="contexts (and, in threads mode, host threads) of the e2e leg")
    ap.add_argument("--e2e-mode", default="threads", choices=["threads", "async"])
    ap.add_argument("--pages", default="oracle", choices=["ours", "oracle"],
                    help="who writes the headline's input pages: the oracle writer (reference chooser + liblz4, default) or this library's GPU encoder")
    ap.add_argument("--config3-rows", type=int, default=10_000_000)
    ap.add_argument("--config4-rows", type=int, default=4_000_000)
    ap.add_argument("--config5-rows", type=int, default=int(os.environ.get("SB_CONFIG5_ROWS", 0)),
                    help="rows of the 64-column table of configs[4] (N > 1); 0 = 12.5 M per GPU, i.e. BASELINE's 100 M rows on 8 GPUs "
                         "(weak scaling), when the writer rank's host has the memory to frame the file, else 4 M")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rows = args.rows
    ncpu = os.cpu_count() or 1
    config = {"workload": "configs[1]: 8 primitive columns (3xi32,3xi64,2xf64) x %d rows, %d rows/page, default LZ4, "
                          "default_compress_ratio 2.0 (adaptive), seed 42" % (rows, PAGE_ROWS),
              "rows": rows, "columns": 8, "page_rows": PAGE_ROWS, "l2": "inputs+outputs (>650 MB per step) exceed the 126 MB L2",
              "partitioning": "one 8-column partition per rank, no collective",
              "pages_written_by": "oracle writer (reference chooser, liblz4 blocks)" if args.pages == "oracle" else "strawboat_b200 GPU encoder"}
    base = {"metric": "decoded GB/s (Arrow bytes out)", "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config}
    from strawboat_b200 import workloads as wl

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        from concurrent.futures import ThreadPoolExecutor
        cols = oracle_write_columns(wl.config2(rows, 42), 42, ncpu)
        threads = ncpu
        tasks, ob = cpu_tasks(cols, threads)
        pool = ThreadPoolExecutor(threads)
        t_all = []
        for i in range(args.warmup + args.steps):
            dt = cpu_decode_pass(tasks, threads, pool)
            if i >= args.warmup:
                t_all.append(dt)
        dt = sum(t_all) / len(t_all)
        val = ob / dt / 1e9
        sample = "the whole workload: all %d rows of the 8 columns, %d decode tasks (column x page range) on %d threads" % (rows, len(tasks), threads)
        line = dict(base, impl="reference", value=val, ms_per_step=dt * 1e3,
                    cpu_baseline={"value": val, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample,
                                  "note": "C++ restatement of the Rust reference (oracle/, -O3 -march=native, liblz4); cargo/rustc absent in this image"},
                    e2e={"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import strawboat_b200 as sb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: strawboat_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t0 = time.time()
    ctx = sb.Context(local_rank, stream=torch.cuda.current_stream())
    src = wl.config2(rows, 42 + rank)
    if args.pages == "oracle":
        cols = oracle_write_columns(src, 42 + rank, max(1, ncpu // world))
    else:
        cols, _ = gpu_write_columns(ctx, src, 42 + rank)
    log(f"[rank {rank}] workload built in {time.time() - t0:.1f}s:", {c['name']: len(c['data']) for c in cols})
    dev_cols, keep = to_device_cols(torch, sb, cols)
    host_cols = []
    bytes_in = 0
    for c in cols:
        th = torch.from_numpy(np.array(c["data"], copy=True)).pin_memory()
        keep.append(th)
        host_cols.append(sb.Column(c["type"], c["nullable"], th.numpy(), c["metas"]))
        bytes_in += len(c["data"])
    bytes_out = sum(rows * np.dtype(sb.NP_OF[c["type"]]).itemsize for c in cols)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of what is being timed: decode(pages) == the generated columns, bit for bit
    check_roundtrip(ctx, sb, cols)

    def step_device():
        out = ctx.decode_columns(dev_cols, out="device")
        st = ctx.last_stats()
        out[0]._group.release()
        return st

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches, main_ms, lz4_ms, host_ms, lz4_bytes, light_ms = 0.0, 0, 0.0, 0.0, 0.0, 0, 0.0
    e0.record()
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        st = step_device()
        kernel_ms += st["device_ms"]
        main_ms += st["main_kernel_ms"]
        lz4_ms += st["lz4_kernel_ms"]
        light_ms += st["light_kernel_ms"]
        host_ms += st["host_ms"]
        lz4_bytes = st["lz4_bytes"]
        launches += st["kernel_launches"]
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - tw0) * 1e3
    # the context runs on torch's current stream, so these events bracket all of its work
    ms_total = e0.elapsed_time(e1)
    codec_pages = st["codec_pages"]
    log(f"[rank {rank}] device span {ms_total:.2f} ms, host wall {wall_ms:.2f} ms over {args.steps} steps")
    barrier()

    # e2e: host pages in pinned memory -> host Arrow buffers
    e2e_ms = None
    link = None
    if not args.no_e2e:
        # what the box's host<->device link does on a plain pinned copy (explains the e2e number)
        hp = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dp = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        link = {}
        for name, (a, b) in (("h2d_gbs", (dp, hp)), ("d2h_gbs", (hp, dp))):
            a.copy_(b, non_blocking=True)
            torch.cuda.synchronize()
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            a.copy_(b, non_blocking=True)
            l1.record()
            torch.cuda.synchronize()
            link[name] = round((256 << 20) / l0.elapsed_time(l1) / 1e6, 1)
        del hp, dp

        # The caller-side pattern of the reference (one reader task per column group, databend style): E2E_THREADS
        # host threads, each with its own context (= its own streams), each decoding its share of the columns.  The
        # calls overlap on the device and on the link (H2D of one group with D2H of another); every byte still crosses
        # the link inside the timed region.  Columns with the fewest page bytes go first: their decoded buffers start
        # coming back over the link -- the bottleneck of this leg -- while the large inputs are still going up.
        # (--e2e-mode async: ONE host thread drives the contexts through sb_decode_columns_async / sb_decode_wait.)
        from concurrent.futures import ThreadPoolExecutor
        n_thr = max(1, min(args.e2e_threads, len(host_cols)))
        order = sorted(range(len(host_cols)), key=lambda i: host_cols[i].nbytes)
        groups = [[host_cols[i] for i in order[t::n_thr]] for t in range(n_thr)]
        ctxs = [ctx] + [sb.Context(local_rank) for _ in range(n_thr - 1)]
        pool = ThreadPoolExecutor(n_thr)
        turn = [threading.Event() for _ in range(n_thr + 1)]

        def one(t):
            turn[t].wait()  # submission order = group order (events, no spinning: the cores are shared with the other ranks)
            h = ctxs[t].decode_columns_async(groups[t], out="host")
            turn[t + 1].set()
            res = h.wait()                 # pinned host buffers, zero-copy numpy views
            chk = int(res[0].values[-1])   # touch the result on the host
            res[0].release()
            return chk

        def step_threads():
            for ev in turn:
                ev.clear()
            futs = [pool.submit(one, t) for t in range(n_thr)]
            turn[0].set()
            return sum(f.result() for f in futs)

        def step_async():
            handles = [ctxs[t].decode_columns_async(groups[t], out="host") for t in range(n_thr)]
            chk = 0
            for h in handles:
                res = h.wait()
                chk += int(res[0].values[-1])
                res[0].release()
            return chk

        step_host = step_async if args.e2e_mode == "async" else step_threads

        for _ in range(5):  # pinned pools of every context at their final size, host caches warm
            step_host()
        barrier()
        tw0 = time.perf_counter()
        k2 = max(10, args.steps)
        e2e_steps = []
        for _ in range(k2):
            ts = time.perf_counter()
            step_host()
            e2e_steps.append((time.perf_counter() - ts) * 1e3)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - tw0) * 1e3 / k2  # the mean over ALL timed steps is what is reported
    sampler.stop_flag = True
    sampler.join()

    # per-column device time (diagnostic, rank 0 only, after the timed region): each column decoded alone
    per_column = []
    if rank == 0 and not args.no_columns:
        for c, dc in zip(cols, dev_cols):
            stc = timed_decode(ctx, [dc], reps=3)
            ob = rows * np.dtype(sb.NP_OF[c["type"]]).itemsize
            per_column.append({"column": c["name"], "codec_pages": stc["codec_pages"],
                               "bytes_in": int(len(c["data"])), "bytes_out": int(ob), "device_us": round(stc["device_ms"] * 1e3, 1),
                               "decoded_gbs": round(ob / stc["device_ms"] / 1e6, 1), "algorithmic_gbs": round((len(c["data"]) + ob) / stc["device_ms"] / 1e6, 1)})

    t = torch.tensor([ms_total, e2e_ms or 0.0, kernel_ms, main_ms, lz4_ms, light_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_max, kernel_ms_max, main_ms_max, lz4_ms_max, light_ms_max = t.tolist()
    ms_per_step = ms_total / args.steps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    line = None
    if rank == 0:
        # the two decode kernels run concurrently (sb_lz4_kernel: top-level LZ4 blocks; sb_decode_kernel:
        # everything else); each is measured with its own CUDA events on its own stream.  The roofline
        # object describes the dominant (longer) one; "kernels" lists both and the whole span.
        k_ms = kernel_ms_max / args.steps
        m_ms, l_ms, lt_ms = main_ms_max / args.steps, lz4_ms_max / args.steps, light_ms_max / args.steps
        total_bytes = bytes_in + bytes_out
        main_bytes = total_bytes - lz4_bytes
        kernels = [{"kernel": "sb_lz4_kernel", "ms": l_ms, "algorithmic_bytes": int(lz4_bytes),
                    "gbs": lz4_bytes / (l_ms * 1e-3) / 1e9 if l_ms > 0 else None},
                   {"kernel": "sb_decode_kernel + sb_decode_light_kernel", "ms": max(m_ms, lt_ms), "algorithmic_bytes": int(main_bytes),
                    "gbs": main_bytes / (max(m_ms, lt_ms) * 1e-3) / 1e9 if max(m_ms, lt_ms) > 0 else None,
                    "full_kernel_ms": m_ms, "light_kernel_ms": lt_ms,
                    "note": "the two decode kernels share the non-LZ4 pages (light: flat fixed-width pages with light codec trees) and run "
                            "concurrently with each other and with sb_lz4_kernel: event times include waiting for SMs"},
                   {"kernel": "all kernels of one step (plan upload .. last kernel)", "ms": k_ms, "algorithmic_bytes": int(total_bytes),
                    "gbs": total_bytes / (k_ms * 1e-3) / 1e9}]
        dom = kernels[0] if (l_ms >= max(m_ms, lt_ms) or l_ms >= 0.5 * k_ms) else kernels[1]
        traffic, traffic_src = measured_traffic(dom["kernel"]) if rows == 10_000_000 and args.pages == "oracle" else (None, None)
        line = dict(base, value=world * bytes_out / (ms_per_step * 1e-3) / 1e9, ms_per_step=ms_per_step,
                    gpu_launches=launches, clocks=sampler.result(),
                    roofline={"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                              "frac": dom["gbs"] / peak, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                              "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "kernel_ms": dom["ms"], "traffic": traffic,
                              "traffic_source": traffic_src, "kernels": kernels},
                    host_ms_per_step=host_ms / args.steps, codec_pages=codec_pages, host_cores=ncpu)
        if per_column:
            line["per_column"] = per_column
        if e2e_ms is not None:
            line["e2e"] = {"value": world * bytes_out / (e2e_max * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_max,
                           "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out, "pinned_copy_probe": link,
                           "contexts": n_thr, "host_threads": 1 if args.e2e_mode == "async" else n_thr,
                           "api": "sb_decode_columns_async + sb_decode_wait, one context per column group",
                           "steps": len(e2e_steps), "warmup": 5, "step_ms_rank0": [round(x, 2) for x in e2e_steps],
                           "note": "value = bytes / mean over all timed steps (max over ranks); step_ms_rank0 shows the spread"}

    # ------------------------------------------------------------------ extras
    if not args.no_extras and world == 1:
        t_ex = time.time()
        # same workload, pages written by this library's encoder; the encode call itself with its roofline
        # the same headline workload with TWO calls in flight (two contexts, sb_decode_columns_async / sb_decode_wait):
        # what a reader that keeps several row groups going sees.  The host planning and table upload of call k+1
        # travel under the kernels of call k, and its LZ4 blocks start while call k's main kernel drains.  Reported
        # next to `value`, which stays the one-call-at-a-time number.
        try:
            line["pipelined"] = section_pipelined(ctx, torch, sb, dev_cols, bytes_out, local_rank, args.steps)
        except Exception as e:
            line["pipelined"] = {"error": repr(e)}
        own, est = gpu_write_columns(ctx, src, 42)
        check_roundtrip(ctx, sb, own)
        dev_own, keep_own = to_device_cols(torch, sb, own)
        sto = timed_decode(ctx, dev_own)
        line["own_pages"] = {"value": round(bytes_out / sto["device_ms"] / 1e6, 1), "unit": "GB/s decoded", "device_ms": round(sto["device_ms"], 4),
                             "lz4_kernel_ms": round(sto["lz4_kernel_ms"], 4), "main_kernel_ms": round(sto["main_kernel_ms"], 4),
                             "bytes_in": int(sum(len(c["data"]) for c in own)), "codec_pages": sto["codec_pages"],
                             "note": "configs[1] with pages written by sb_encode_columns (own LZ4 matcher) instead of the oracle writer"}
        del dev_own, keep_own
        enc_alg = est["bytes_in"] + est["bytes_out"]
        line["encode"] = {"value": est["bytes_in"] / (est["device_ms"] * 1e-3) / 1e9, "unit": "GB/s (Arrow bytes in / device time)",
                          "device_ms": est["device_ms"], "bytes_in": est["bytes_in"], "bytes_out": est["bytes_out"], "codec_pages": est["codec_pages"],
                          "launches": est["kernel_launches"],
                          "roofline": {"bound": "hbm", "kernel": "sb_encode_kernel (+ gather)", "achieved": enc_alg / (est["device_ms"] * 1e-3) / 1e9, "peak": peak,
                                       "unit": "GB/s", "frac": enc_alg / (est["device_ms"] * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": int(enc_alg),
                                       "note": "Arrow bytes read once + encoded bytes written once; statistics / sampling / slab traffic not counted"}}
        for name, fn in (("north_star", lambda: section_north_star(ctx, torch, sb, peak)),
                         ("config3", lambda: section_config3(ctx, torch, sb, peak, args.config3_rows)),
                         ("config4", lambda: section_config4(ctx, torch, sb, peak, args.config4_rows))):
            try:
                line[name] = fn()
            except Exception as e:  # an extra must never take the headline down
                line[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        log(f"extras took {time.time() - t_ex:.1f}s")
    if not args.no_extras and world > 1:
        try:
            from strawboat_b200 import parallel
            rows5 = args.config5_rows
            if rows5 <= 0:
                # the writer rank frames the file on the host (body + sink + re-read copy): ~125 bytes per row, 5 copies
                rows5 = min(100_000_000, 12_500_000 * world)
                avail = 0
                try:
                    avail = [int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0]
                except Exception:
                    pass
                flag = torch.tensor([1 if avail >= rows5 * 125 * 8 else 0], device="cuda")
                dist.broadcast(flag, src=0)
                if int(flag.item()) == 0:
                    rows5 = 4_000_000
            res = parallel.bench_config5(ctx, torch, dist, rank, world, rows5, PAGE_ROWS)
            if rank == 0:
                line["multi_gpu_encode"] = res
        except Exception as e:
            if rank == 0:
                line["multi_gpu_encode"] = {"error": repr(e)}

    if rank == 0:
        if world == 1 and not args.no_cpu:
            from concurrent.futures import ThreadPoolExecutor
            sbo = oracle()
            threads = ncpu
            pool = ThreadPoolExecutor(threads)
            tasks1, ob = cpu_tasks(cols, 1)
            dt1 = min(cpu_decode_pass(tasks1, 1, ThreadPoolExecutor(1)) for _ in range(2))
            tasks, ob = cpu_tasks(cols, threads)
            dtn = min(cpu_decode_pass(tasks, threads, pool) for _ in range(4))
            line["cpu_baseline"] = {"value": ob / dtn / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                    "single_thread_value": ob / dt1 / 1e9,
                                    "sample": "the whole workload (all %d rows of the 8 columns), oracle batch decode, %d tasks on %d threads" % (rows, len(tasks), threads)}
            if "encode" in line:
                # the CPU side of the encode number: the oracle writer (stats -> chooser -> codec, liblz4), whole workload
                t0 = time.perf_counter()
                oracle_write_columns(src, 42, threads)
                dt = time.perf_counter() - t0
                line["encode"]["cpu_baseline"] = {"value": line["encode"]["bytes_in"] / dt / 1e9, "unit": "GB/s", "cores": min(threads, len(src)), "kind": "port",
                                                  "sample": "the whole workload, oracle page writer, one thread per column"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
