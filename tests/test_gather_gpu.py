"""sb_gather_encoded (multi-GPU encode gather) on ONE GPU: a world of size 1 still runs the C++ path
end to end -- NCCL communicator, ncclAllGather of the layout, staging buffer in leaf order, PageMeta hand-over,
release -- everything but the ncclSend / ncclRecv pairs, which bench.py --gpus N exercises (multi_gpu_encode)."""
import io

import numpy as np
import pytest
import sbo
from helpers import oracle_decode_column

import strawboat_b200 as sb
from strawboat_b200 import fileio

pytestmark = pytest.mark.gpu


def test_gather_world_of_one(ctx):
    rng = np.random.default_rng(0)
    cols = [(sb.I64, rng.integers(0, 100, 5000).astype(np.int64), None),
            (sb.F64, rng.integers(0, 8, 7000).astype(np.float64), rng.random(7000) > 0.2),
            (sb.I32, np.zeros(0, np.int32), None),
            (sb.I32, np.cumsum(rng.integers(0, 4, 3000)).astype(np.int32), None)]
    wo = sb.write_options(sb.C_LZ4, 2.0, 1024, seed=9)
    enc = ctx.encode_columns([sb.LeafArray(t, v, validity=val) for t, v, val in cols], wo, out="device")
    host = ctx.encode_columns([sb.LeafArray(t, v, validity=val) for t, v, val in cols], wo)
    comm = sb.Comm(ctx, 0, 1, sb.comm_unique_id())
    got, st = comm.gather_encoded(enc, len(cols), writer=0)
    assert st["total_bytes"] == sum(h.nbytes for h in host) and st["bytes_moved"] == 0
    import torch
    body = torch.as_tensor(sb._DevArray(got[0].ptr, max(1, st["total_bytes"]), got), device="cuda").cpu().numpy().tobytes()[:st["total_bytes"]]
    pos = 0
    sink = io.BytesIO()
    w = fileio.NativeWriter(ctx, sink, b"", wo)
    w.start()
    for g, h, (t, v, val) in zip(got, host, cols):
        assert g.metas == h.metas and g.nbytes == h.nbytes
        assert body[pos:pos + g.nbytes] == h.data          # same bytes as the single-GPU writer, at the scanned offset
        w.write_encoded(body[pos:pos + g.nbytes], g.metas)
        pos += g.nbytes
    w.finish()
    data = sink.getvalue()
    for cm, (t, v, val) in zip(fileio.read_meta(data), cols):
        ref = oracle_decode_column(t, val is not None, fileio.column_body(data, cm), cm[1])
        keep = slice(None) if val is None else val
        assert ref["length"] == len(v) and np.array_equal(ref["values"][keep], v[keep])
    ctx.release_encoded(got)
    ctx.release_encoded(enc)
    comm.close()
    with pytest.raises(sb.StrawboatError):  # leaf c lives on rank c mod world: wrong local column count
        c2 = sb.Comm(ctx, 0, 1, sb.comm_unique_id())
        try:
            c2.gather_encoded([], 3, writer=0)
        finally:
            c2.close()
