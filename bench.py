#!/usr/bin/env python
"""bench.py -- decoded GB/s of the strawboat page-decode hot path on B200 (BASELINE.json metric).

A "step" = one batched decode of the whole workload: configs[1] of BASELINE.json -- 8 primitive
columns (3 x i32, 3 x i64, 2 x f64) x 10 M rows, 8192 rows/page, default LZ4,
default_compress_ratio 2.0 (adaptive), one distribution per codec.

  value : Arrow bytes out / device time, page bytes already resident in HBM
          (plan upload + every kernel of sb_decode_columns inside the timed region)
  e2e   : same call with HOST page bytes in pinned memory and HOST Arrow buffers out
          (H2D of the pages + D2H of the decoded buffers inside the timed region)
  roofline : the decode kernel's algorithmic bytes (sum PageMeta.length read + Arrow bytes
          written) / its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the oracle (C++ restatement of the Rust reference; the
          Rust crate cannot be built in this image) timed on the host cores.

N > 1: every rank decodes its own 8-column x 10 M-row partition (weak scaling, no data-path
collective: pages are independent); value = total bytes of all ranks / max-over-ranks time.
"""
import argparse
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


def build_workload(rows, seed, ctx=None):
    """Generate config 2 and encode it page by page (setup, untimed for the decode metric).

    With a Context the pages are written by THIS library's GPU encoder (sb_encode_columns:
    stats -> choose_compressor -> codec, the product path; its device time is reported as the
    `encode` object of the JSON line).  Without one (the --impl reference / cpu_baseline legs,
    which may run where no GPU is visible) the oracle writer is used -- the same chooser, LZ4
    blocks from liblz4."""
    from strawboat_b200 import workloads as wl
    cols = wl.config2(rows, seed)
    if ctx is not None:
        import strawboat_b200 as sb
        wo = sb.write_options(sb.C_LZ4, 2.0, PAGE_ROWS, seed=seed)
        arrays = [sb.LeafArray(t, v, validity=val) for (_, t, v, val) in cols]
        enc = ctx.encode_columns(arrays, wo)
        st = ctx.last_stats()
        out = []
        for (name, t, v, val), e in zip(cols, enc):
            out.append({"name": name, "type": t, "nullable": val is not None, "data": np.frombuffer(e.data, dtype=np.uint8),
                        "metas": e.metas, "values": v, "validity": val, "pages": None, "codecs": {}})
        enc_stats = {"device_ms": st["device_ms"], "bytes_in": int(sum(np.asarray(c[2]).nbytes for c in cols)),
                     "bytes_out": int(sum(len(e.data) for e in enc)), "codec_pages": st["codec_pages"]}
        return out, enc_stats

    from concurrent.futures import ThreadPoolExecutor
    sbo = oracle()

    def enc(c):
        name, t, v, val = c
        opts = sbo.make_opts(sbo.C_LZ4, ratio=2.0)
        pages, metas, trees = [], [], {}
        for pi, o in enumerate(range(0, rows, PAGE_ROWS)):
            opts.seed = seed + pi
            page = sbo.write_page(t, v[o:o + PAGE_ROWS], None if val is None else val[o:o + PAGE_ROWS], opts=opts)
            pages.append(page)
            metas.append((len(page), min(PAGE_ROWS, rows - o)))
            if pi % 97 == 0:
                tr = sbo.stat_page(t, val is not None, page)
                trees[tr] = trees.get(tr, 0) + 1
        return {"name": name, "type": t, "nullable": val is not None, "data": np.frombuffer(b"".join(pages), dtype=np.uint8),
                "metas": metas, "values": v, "validity": val, "pages": pages, "codecs": trees}

    with ThreadPoolExecutor(8) as ex:
        return list(ex.map(enc, cols)), None


def split_pages_of(c):
    """page byte strings of one encoded column (for the CPU legs)"""
    if c["pages"] is None:
        buf, pos, pages = c["data"].tobytes(), 0, []
        for ln, _ in c["metas"]:
            pages.append(buf[pos:pos + ln])
            pos += ln
        c["pages"] = pages
    return c["pages"]


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.reasons = index, [], False, set()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_decode_time(sbo, cols, rows_limit, threads):
    """oracle batch decode (read_integer / read_double page loops) of the first rows_limit rows
    of every column; one thread per column like a per-column reader task."""
    from concurrent.futures import ThreadPoolExecutor
    npages = max(1, rows_limit // PAGE_ROWS)
    jobs = []
    out_bytes = 0
    for c in cols:
        cp = split_pages_of(c)
        pages = [(cp[i], c["metas"][i][1]) for i in range(min(npages, len(cp)))]
        jobs.append((sbo.make_leaf(c["type"], c["nullable"]), pages))
        out_bytes += sum(p[1] for p in pages) * sbo.WIDTH[c["type"]]
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda j: sbo.read_column(j[0], j[1])["length"], jobs))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return out_bytes, best


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
    ap.add_argument("--e2e-threads", type=int, default=8, help="host threads (one context each) of the e2e leg")
    ap.add_argument("--pages", default="ours", choices=["ours", "oracle"],
                    help="who writes the input pages: this library's GPU encoder (default) or the oracle writer (diagnostic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rows = args.rows
    config = {"workload": "configs[1]: 8 primitive columns (3xi32,3xi64,2xf64) x %d rows, %d rows/page, default LZ4, "
                          "default_compress_ratio 2.0 (adaptive), seed 42" % (rows, PAGE_ROWS),
              "rows": rows, "columns": 8, "page_rows": PAGE_ROWS, "l2": "inputs+outputs (>700 MB per step) exceed the 126 MB L2",
              "partitioning": "one 8-column partition per rank, no collective"}
    base = {"metric": "decoded GB/s (Arrow bytes out)", "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        sbo = oracle()
        sample_rows = min(rows, 2_000_000)
        cols, _ = build_workload(sample_rows, 42)
        threads = os.cpu_count() or 1
        t_all = []
        for i in range(args.warmup + args.steps):
            ob, dt = cpu_decode_time(sbo, cols, sample_rows, min(threads, len(cols)))
            if i >= args.warmup:
                t_all.append(dt)
        dt = sum(t_all) / len(t_all)
        val = ob / dt / 1e9
        sample = "first %d rows of each of the 8 columns, one thread per column" % sample_rows
        line = dict(base, impl="reference", value=val, ms_per_step=dt * 1e3,
                    cpu_baseline={"value": val, "unit": "GB/s", "cores": min(threads, len(cols)), "kind": "port", "sample": sample,
                                  "note": "C++ restatement of the Rust reference (oracle/); cargo/rustc absent in this image"},
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
    if args.pages == "oracle":  # diagnostic: liblz4-written LZ4 blocks
        cols, enc_stats = build_workload(rows, 42 + rank)
    else:
        cols, enc_stats = build_workload(rows, 42 + rank, ctx)
    log(f"[rank {rank}] workload built in {time.time() - t0:.1f}s:",
        {c['name']: (len(c['data']), c['codecs']) for c in cols}, enc_stats)
    dev_cols, host_cols, keep = [], [], []
    bytes_in = 0
    for c in cols:
        td = torch.from_numpy(c["data"].copy()).cuda()
        th = torch.from_numpy(c["data"].copy()).pin_memory()
        keep += [td, th]
        dev_cols.append(sb.Column(c["type"], c["nullable"], td, c["metas"]))
        hc = sb.Column(c["type"], c["nullable"], th.numpy(), c["metas"])
        host_cols.append(hc)
        bytes_in += len(c["data"])
    bytes_out = sum(rows * np.dtype(sb.NP_OF[c["type"]]).itemsize for c in cols)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of what is being timed: decode(encode(x)) == x, bit for bit
    res = ctx.decode_columns(host_cols, out="host")
    for c, r in zip(cols, res):
        assert np.array_equal(r.values.view(np.uint8), np.ascontiguousarray(c["values"]).view(np.uint8)), c["name"]
    del res

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
    kernel_ms, launches, main_ms, lz4_ms, host_ms, lz4_bytes = 0.0, 0, 0.0, 0.0, 0.0, 0
    e0.record()
    tw0 = time.perf_counter()
    for _ in range(args.steps):
        st = step_device()
        kernel_ms += st["device_ms"]
        main_ms += st["main_kernel_ms"]
        lz4_ms += st["lz4_kernel_ms"]
        host_ms += st["host_ms"]
        lz4_bytes = st["lz4_bytes"]
        launches += st["kernel_launches"]
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - tw0) * 1e3
    # the context runs on torch's current stream, so these events bracket all of its work
    ms_total = e0.elapsed_time(e1)
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

        # The caller-side pattern of the reference (one reader task per column group, databend style):
        # E2E_THREADS host threads, each with its own context (= its own stream), each decoding its
        # share of the columns.  The calls overlap on the device and on the link (H2D of one group
        # with D2H of another); every byte still crosses the link inside the timed region.
        from concurrent.futures import ThreadPoolExecutor
        n_thr = max(1, min(args.e2e_threads, len(host_cols)))
        # Columns with the fewest page bytes go first (SB_E2E_STAGGER_US apart): their decoded buffers start
        # coming back over the link -- the bottleneck of this leg -- while the large inputs are still going up.
        order = sorted(range(len(host_cols)), key=lambda i: host_cols[i].nbytes)
        groups = [[host_cols[i] for i in order[t::n_thr]] for t in range(n_thr)]
        stagger = float(os.environ.get("SB_E2E_STAGGER_US", "100")) * 1e-6
        ctxs = [ctx] + [sb.Context(local_rank) for _ in range(n_thr - 1)]
        pool = ThreadPoolExecutor(n_thr)

        def one(t):
            if stagger:
                t_go = step_t0[0] + t * stagger
                while time.perf_counter() < t_go:
                    pass
            out = ctxs[t].decode_columns(groups[t], out="host", copy=False)  # pinned host buffers, zero-copy numpy views
            chk = int(out[0].values[-1])  # touch the result on the host
            out[0].release()
            return chk

        step_t0 = [0.0]

        def step_host():
            step_t0[0] = time.perf_counter()
            return sum(pool.map(one, range(n_thr)))

        for _ in range(3):
            step_host()
        barrier()
        tw0 = time.perf_counter()
        k2 = max(5, args.steps // 2)
        for _ in range(k2):
            step_host()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - tw0) * 1e3 / k2
    sampler.stop_flag = True
    sampler.join()

    # per-column device time (diagnostic, rank 0 only, after the timed region): each column decoded alone
    per_column = []
    if rank == 0 and not args.no_columns:
        for c, dc in zip(cols, dev_cols):
            best = None
            for _ in range(3):
                out = ctx.decode_columns([dc], out="device")
                stc = ctx.last_stats()
                out[0]._group.release()
                best = stc["device_ms"] if best is None else min(best, stc["device_ms"])
            ob = rows * np.dtype(sb.NP_OF[c["type"]]).itemsize
            per_column.append({"column": c["name"], "codec_pages": stc["codec_pages"],
                               "bytes_in": int(len(c["data"])), "bytes_out": int(ob), "device_us": round(best * 1e3, 1),
                               "decoded_gbs": round(ob / best / 1e6, 1), "algorithmic_gbs": round((len(c["data"]) + ob) / best / 1e6, 1)})

    t = torch.tensor([ms_total, e2e_ms or 0.0, kernel_ms, main_ms, lz4_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_max, kernel_ms_max, main_ms_max, lz4_ms_max = t.tolist()
    ms_per_step = ms_total / args.steps

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        # the two decode kernels run concurrently (sb_lz4_kernel: top-level LZ4 blocks; sb_decode_kernel:
        # everything else); each is measured with its own CUDA events on its own stream.  The roofline
        # object describes the dominant (longer) one; "kernels" lists both and the whole span.
        k_ms = kernel_ms_max / args.steps
        m_ms, l_ms = main_ms_max / args.steps, lz4_ms_max / args.steps
        total_bytes = bytes_in + bytes_out
        main_bytes = total_bytes - lz4_bytes
        kernels = [{"kernel": "sb_lz4_kernel", "ms": l_ms, "algorithmic_bytes": int(lz4_bytes),
                    "gbs": lz4_bytes / (l_ms * 1e-3) / 1e9 if l_ms > 0 else None},
                   {"kernel": "sb_decode_kernel", "ms": m_ms, "algorithmic_bytes": int(main_bytes),
                    "gbs": main_bytes / (m_ms * 1e-3) / 1e9 if m_ms > 0 else None},
                   {"kernel": "all kernels of one step (plan upload .. last kernel)", "ms": k_ms, "algorithmic_bytes": int(total_bytes),
                    "gbs": total_bytes / (k_ms * 1e-3) / 1e9}]
        # sb_lz4_kernel is placed first (high-priority side stream) and holds the SMs for most of the step; the
        # main kernel's CTAs fill in as LZ4 blocks retire, so its event time includes that wait.  The dominant
        # kernel is the LZ4 one whenever it spans at least half of the step (same order as the serialised ncu list).
        kernels[1]["note"] = "event time includes waiting for SMs behind sb_lz4_kernel when LZ4 blocks are present"
        dom = kernels[0] if (l_ms >= m_ms or l_ms >= 0.5 * k_ms) else kernels[1]
        # DRAM traffic per launch of that kernel: from the committed `ncu --set full` capture of this same command
        # (profiles/r1_ncu_full_config2.csv, written by tools/ncu_summary.py); null when the file is absent
        traffic = None
        try:
            import csv
            rows_ = list(csv.reader(open(os.path.join(ROOT, "profiles", "r1_ncu_full_config2.csv"))))
            col_ = next(i for i, h in enumerate(rows_[0]) if h.startswith(dom["kernel"]))
            byt = {r[0]: float(r[col_]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[r[1]] for r in rows_ if r[0].startswith("dram__bytes_")}
            traffic = int(byt["dram__bytes_read.sum"] + byt["dram__bytes_write.sum"]) if rows == 10_000_000 and enc_stats else None
        except Exception:
            traffic = None
        line = dict(base, value=world * bytes_out / (ms_per_step * 1e-3) / 1e9, ms_per_step=ms_per_step,
                    gpu_launches=launches, clocks=sampler.result(),
                    roofline={"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s",
                              "frac": dom["gbs"] / peak, "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                              "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "kernel_ms": dom["ms"], "traffic": traffic,
                              "traffic_source": "profiles/r1_ncu_full_config2.csv (ncu --set full of this command, dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
                              "kernels": kernels},
                    host_ms_per_step=host_ms / args.steps)
        if per_column:
            line["per_column"] = per_column
        if enc_stats:
            line["encode"] = {"value": enc_stats["bytes_in"] / (enc_stats["device_ms"] * 1e-3) / 1e9, "unit": "GB/s (Arrow bytes in / device time)",
                              "device_ms": enc_stats["device_ms"], "bytes_in": enc_stats["bytes_in"], "bytes_out": enc_stats["bytes_out"],
                              "note": "sb_encode_columns wrote the pages this run decodes (one call, untimed for the decode metric)"}
        line["config"]["pages_written_by"] = "strawboat_b200 GPU encoder" if enc_stats else "oracle writer (liblz4)"
        if e2e_ms is not None:
            line["e2e"] = {"value": world * bytes_out / (e2e_max * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_max,
                           "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out, "pinned_copy_probe": link,
                           "host_threads": min(args.e2e_threads, len(host_cols))}
        if world == 1 and not args.no_cpu:
            sbo = oracle()
            sample_rows = min(rows, 1_000_000)
            ob, dt1 = cpu_decode_time(sbo, cols, sample_rows, 1)
            threads = min(os.cpu_count() or 1, len(cols))
            ob, dtn = cpu_decode_time(sbo, cols, sample_rows, threads)
            line["cpu_baseline"] = {"value": ob / dtn / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                    "single_thread_value": ob / dt1 / 1e9,
                                    "sample": "first %d rows of each of the 8 columns (oracle batch decode, one thread per column)" % sample_rows}
            if "encode" in line:
                # the CPU side of the encode number: the oracle writer (stats -> chooser -> codec, liblz4) on the same sample
                from concurrent.futures import ThreadPoolExecutor

                def enc_col(c):
                    opts = sbo.make_opts(sbo.C_LZ4, ratio=2.0)
                    v, val = c["values"], c["validity"]
                    for pi, o in enumerate(range(0, sample_rows, PAGE_ROWS)):
                        opts.seed = 42 + pi
                        sbo.write_page(c["type"], v[o:min(o + PAGE_ROWS, sample_rows)], None if val is None else val[o:min(o + PAGE_ROWS, sample_rows)], opts=opts)
                    return np.asarray(v[:sample_rows]).nbytes
                best = None
                with ThreadPoolExecutor(threads) as ex:
                    for _ in range(2):
                        t0 = time.perf_counter()
                        nb = sum(ex.map(enc_col, cols))
                        dt = time.perf_counter() - t0
                        best = dt if best is None else min(best, dt)
                line["encode"]["cpu_baseline"] = {"value": nb / best / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
                                                  "sample": "first %d rows of each of the 8 columns (oracle page writer, one thread per column)" % sample_rows}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
