"""Multi-GPU partitioning of the page encode / decode path.

Leaf columns are independent, so ranks take disjoint column sets (`leaf_idx mod world`) and
decode needs no collective.  Encode has one exchange step: the file is column-major with
absolute ColumnMeta offsets and a single footer (src/write/common.rs:76,111-114,
src/write/writer.rs:128-167), so the encoded column bodies are gathered to the writer rank:
  1. all_gather of the per-column encoded sizes,
  2. point-to-point transfer of every rank's bodies to the writer (NCCL over NVLink on GPUs,
     gloo in the CPU tests) -- compressed bytes only,
  3. the writer frames header, bodies in leaf order and footer (fileio.write_file).
"""
import pickle

import torch
import torch.distributed as dist


def shard_columns(n_cols, world, rank):
    """leaf index -> rank placement: column c lives on rank c mod world."""
    return [c for c in range(n_cols) if c % world == rank]


def gather_encoded(local, n_cols, dst=0, device=None, group=None):
    """local: {leaf_idx: (body: bytes | uint8 tensor, metas)} of this rank.
    Returns on `dst` the list [(body bytes, metas)] in leaf order, None elsewhere."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = device if device is not None else torch.device("cpu")
    sizes = torch.zeros(n_cols, dtype=torch.int64, device=dev)
    for c, (body, _) in local.items():
        sizes[c] = body.numel() if torch.is_tensor(body) else len(body)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)            # 8 bytes per column
    metas_blob = pickle.dumps({c: m for c, (_, m) in local.items()})
    metas_all = [None] * world if rank == dst else None
    dist.gather_object(metas_blob, metas_all, dst=dst, group=group)

    def as_tensor(body):
        if torch.is_tensor(body):
            return body.to(dev).reshape(-1)
        return torch.frombuffer(bytearray(body), dtype=torch.uint8).to(dev) if len(body) else torch.zeros(0, dtype=torch.uint8, device=dev)

    mine = sorted(local)
    if rank != dst:
        payload = torch.cat([as_tensor(local[c][0]) for c in mine]) if mine else torch.zeros(0, dtype=torch.uint8, device=dev)
        if payload.numel():
            dist.send(payload, dst, group=group)
        return None
    bodies = {c: as_tensor(local[c][0]) for c in mine}
    for r in range(world):
        if r == dst:
            continue
        sz = all_sizes[r].tolist()
        cols = [c for c in range(n_cols) if c % world == r]
        total = int(sum(sz[c] for c in cols))
        if total == 0:
            for c in cols:
                bodies[c] = torch.zeros(0, dtype=torch.uint8, device=dev)
            continue
        buf = torch.empty(total, dtype=torch.uint8, device=dev)
        dist.recv(buf, r, group=group)
        pos = 0
        for c in cols:
            bodies[c] = buf[pos:pos + int(sz[c])]
            pos += int(sz[c])
    metas = {}
    for blob in metas_all:
        metas.update(pickle.loads(blob))
    return [(bytes(bodies[c].cpu().numpy().tobytes()), metas[c]) for c in range(n_cols)]


# ------------------------------------------------------------------------------------------------------------
# configs[4] of BASELINE.json: one 64-column mixed-type table, generated ON THE DEVICE, leaf c on rank c mod world,
# encoded on every rank, gathered on the writer rank with NCCL from C++ (sb_gather_encoded), framed, re-read, checked.
# ------------------------------------------------------------------------------------------------------------
def config5_schema():
    """64 leaves: 16 i32, 16 i64, 16 f64, 8 Utf8, 4 LargeBinary, 4 Boolean; distributions cycle
    through the ones of configs 2 / 3.  Interleaved so that every rank gets every type."""
    from ._capi import BINARY, BOOL, F64, I32, I64, LARGE_BINARY
    kinds = {I32: ["sorted", "lowcard", "random", "const"], I64: ["const", "freq", "runs", "random"],
             F64: ["lowcard", "int16", "runs", "const"]}
    # eight rows of eight columns, one type per row: `c mod world` (world = 2, 4, 8) then gives every rank the same type mix
    rows = [I32, I64, F64, BINARY, I32, I64, F64, None]
    cols = []
    for c in range(64):
        r, k = divmod(c, 8)
        t = rows[r]
        if t is None:
            cols.append((LARGE_BINARY, "dict", True) if k % 2 == 0 else (BOOL, "p30", False))
        elif t == BINARY:
            cols.append((BINARY, "dict", True))
        else:
            cols.append((t, kinds[t][(k + r // 4) % 4], False))
    return cols


def gen_column_device(torch, spec, n, seed, device):
    """one column of config 5 as device tensors: dict(values | (offsets, data), validity bits (bool tensor) or None)"""
    from ._capi import BINARY, BOOL, F64, I32, I64, LARGE_BINARY
    t, kind, nullable = spec
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    ri = lambda lo, hi, dt=torch.int64: torch.randint(lo, hi, (n,), generator=g, device=device, dtype=dt)  # noqa: E731
    valid = None
    if t in (I32, I64, F64):
        if kind == "sorted":
            v = torch.cumsum(ri(0, 4), 0)
        elif kind == "lowcard":
            v = ri(0, 8)
        elif kind == "random":
            v = ri(-2**31, 2**31) if t == I32 else ri(-2**62, 2**62)
        elif kind == "const":
            v = torch.full((n,), 7_000_007, dtype=torch.int64, device=device)
        elif kind == "freq":
            exc = torch.rand(n, generator=g, device=device) < 0.05
            v = torch.where(exc, 10000 + ri(0, 1 << 20), torch.full((n,), 20, dtype=torch.int64, device=device))
        elif kind == "runs":
            nr = n // 64 + 2
            vals = torch.randint(0, 1 << 30, (nr,), generator=g, device=device)
            v = torch.repeat_interleave(vals, 64)[:n]
        elif kind == "int16":
            v = ri(0, 65536)
        else:
            raise ValueError(kind)
        v = v.to({I32: torch.int32, I64: torch.int64, F64: torch.float64}[t]).contiguous()
        return {"type": t, "values": v, "valid": None, "arrow_bytes": v.numel() * v.element_size()}
    if t == BOOL:
        bits = torch.rand(n, generator=g, device=device) < 0.3
        return {"type": t, "values": bits, "valid": None, "arrow_bytes": (n + 7) // 8}
    # decimal strings of integers(0, 1000), 40 % nulls (empty slots)
    table = torch.zeros((1000, 3), dtype=torch.uint8)
    tlen = torch.zeros(1000, dtype=torch.int64)
    for i in range(1000):
        b = str(i).encode()
        table[i, :len(b)] = torch.tensor(list(b), dtype=torch.uint8)
        tlen[i] = len(b)
    table, tlen = table.to(device), tlen.to(device)
    ids = ri(0, 1000)
    valid = torch.rand(n, generator=g, device=device) >= 0.4
    lens = torch.where(valid, tlen[ids], torch.zeros_like(ids))
    odt = torch.int64 if t == LARGE_BINARY else torch.int32
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(lens, 0)
    total = int(offsets[-1])
    row = torch.repeat_interleave(torch.arange(n, device=device), lens)
    pos = torch.arange(total, device=device) - offsets[:-1][row]
    data = table[ids[row], pos].contiguous()
    off = offsets.to(odt).contiguous()
    return {"type": t, "values": (off, data), "valid": valid,
            "arrow_bytes": off.numel() * off.element_size() + data.numel() + (n + 7) // 8}


def pack_bits_device(torch, bits):
    """bool tensor -> LSB-first packed uint8 tensor (Arrow bitmap)"""
    n = bits.numel()
    pad = (-n) % 8
    b = torch.cat([bits.to(torch.uint8), torch.zeros(pad, dtype=torch.uint8, device=bits.device)]) if pad else bits.to(torch.uint8)
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.uint8, device=bits.device)
    return (b.view(-1, 8) * w).sum(1, dtype=torch.int32).to(torch.uint8).contiguous()


def leaf_array_of(torch, sb, col, n):
    t = col["type"]
    if t == sb.BOOL:
        return sb.LeafArray(t, pack_bits_device(torch, col["values"]), length=n)
    validity = None if col["valid"] is None else pack_bits_device(torch, col["valid"])
    return sb.LeafArray(t, col["values"], validity=validity, nullable=col["valid"] is not None)


def check_column_device(torch, sb, col, dec, n):
    """decoded (device) column == generated column; binary columns on the valid rows (null slots are codec dependent)"""
    dev = col["values"][0].device if isinstance(col["values"], tuple) else col["values"].device
    view = lambda which, dt: torch.as_tensor(dec.device_view(which), device=dev).view(dt)  # noqa: E731
    t = col["type"]
    if t == sb.BOOL:
        return bool(torch.equal(view("values", torch.uint8)[:(n + 7) // 8], pack_bits_device(torch, col["values"])))
    if isinstance(col["values"], tuple):
        off, data = col["values"]
        got_off = view("offsets", off.dtype).to(torch.int64)
        got_val = view("values", torch.uint8)
        valid = col["valid"]
        lens, glens = torch.diff(off.to(torch.int64))[valid], torch.diff(got_off)[valid]
        if not torch.equal(lens, glens):
            return False
        row = torch.repeat_interleave(torch.arange(lens.numel(), device=dev), lens)
        pos = torch.arange(int(lens.sum()), device=dev) - (torch.cumsum(lens, 0) - lens)[row]
        ok = torch.equal(got_val[got_off[:-1][valid][row] + pos], data[off.to(torch.int64)[:-1][valid][row] + pos])
        return bool(ok) and bool(torch.equal(view("validity", torch.uint8)[:(n + 7) // 8], pack_bits_device(torch, valid)))
    v = col["values"]
    return bool(torch.equal(view("values", torch.uint8), v.view(torch.uint8).reshape(-1)))


def bench_config5(ctx, torch, dist, rank, world, rows, page_rows, seed=4242):
    """encode on every rank -> NCCL gather on rank 0 -> file -> re-read -> decode -> check.  Returns the report (rank 0)."""
    import io
    import time

    import strawboat_b200 as sb
    from . import fileio
    dev = torch.device("cuda", torch.cuda.current_device())
    schema = config5_schema()
    n_total = len(schema)
    uid = [sb.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = sb.Comm(ctx, rank, world, uid[0])
    mine = shard_columns(n_total, world, rank)
    cols = [gen_column_device(torch, schema[c], rows, seed + c, dev) for c in mine]
    arrays = [leaf_array_of(torch, sb, col, rows) for col in cols]
    wo = sb.write_options(sb.C_LZ4, 2.0, page_rows, seed=seed)
    arrow_in = sum(c["arrow_bytes"] for c in cols)
    best_enc, enc = None, None
    for _ in range(2):
        if enc is not None:
            ctx.release_encoded(enc)
        torch.cuda.synchronize()
        enc = ctx.encode_columns(arrays, wo, out="device")
        ms = ctx.last_stats()["device_ms"]
        best_enc = ms if best_enc is None else min(best_enc, ms)
    gathered, gst = None, None
    for it in range(2):  # the first gather pays NCCL's connection set-up
        if gathered is not None:
            ctx.release_encoded(gathered)
        torch.cuda.synchronize()
        dist.barrier()
        gathered, gst = comm.gather_encoded(enc, n_total, writer=0)
    t = torch.tensor([best_enc, gst["gather_ms"], float(arrow_in), float(gst["bytes_moved"])], dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    report = None
    if rank == 0:
        t0 = time.time()
        total = gst["total_bytes"]
        body = torch.as_tensor(sb._DevArray(gathered[0].ptr, total, gathered), device=dev).cpu().numpy().tobytes() if total else b""
        sink = io.BytesIO()
        w = fileio.NativeWriter(ctx, sink, b"", wo)
        w.start()
        pos = 0
        for g in gathered:
            w.write_encoded(body[pos:pos + g.nbytes], g.metas)
            pos += g.nbytes
        size = w.finish()
        data = sink.getvalue()
        metas = fileio.read_meta(data)
        ok = len(metas) == n_total and size == len(data)
        # re-read: decode every column from the file bytes and compare with the regenerated source
        for c0 in range(0, n_total, 8):
            group = list(range(c0, min(n_total, c0 + 8)))
            cin = [sb.Column(schema[c][0], schema[c][2], fileio.column_body(data, metas[c]), metas[c][1]) for c in group]
            decs = ctx.decode_columns(cin, out="device")
            for c, d in zip(group, decs):
                src = gen_column_device(torch, schema[c], rows, seed + c, dev)
                ok = ok and check_column_device(torch, sb, src, d, rows)
            decs[0].release()
        enc_ms, gather_ms = float(tmax[0]), float(t[1])
        report = {"workload": "configs[4]: 64 mixed-type columns (16 i32, 16 i64, 16 f64, 8 Utf8, 4 LargeBinary, 4 Boolean) x %d rows, generated on the "
                              "device, leaf c on rank c mod %d, %d rows/page, default LZ4, adaptive" % (rows, world, page_rows),
                  "encode": {"value": float(tsum[2]) / (enc_ms * 1e-3) / 1e9, "unit": "GB/s (Arrow bytes in of all ranks / max-over-ranks device time)",
                             "device_ms": enc_ms, "arrow_bytes_in": int(tsum[2])},
                  "gather": {"collective": "ncclAllGather of (bytes, pages) per column + grouped ncclSend/ncclRecv of the bodies to rank 0, from C++ (sb_gather_encoded)",
                             "bytes_into_writer": gst["bytes_moved"], "ms": gather_ms,
                             "value": gst["bytes_moved"] / (gather_ms * 1e-3) / 1e9 if gather_ms > 0 else None, "unit": "GB/s into the writer rank over NVLink"},
                  "file_bytes": len(data), "columns": n_total, "reread_and_checked": bool(ok), "host_framing_s": round(time.time() - t0, 2)}
    ctx.release_encoded(gathered) if gathered else None
    ctx.release_encoded(enc)
    comm.close()
    return report
