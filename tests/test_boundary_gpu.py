"""The ownership / overlap half of the reader boundary: plan -> allocate -> run into caller-owned
buffers (sb_plan_columns, sb_decode_columns_into) and the asynchronous form (sb_decode_columns_async / sb_decode_wait /
sb_decode_ready), against the synchronous library-allocated path and the oracle."""
import numpy as np
import pytest
import sbo
from helpers import oracle_decode_column, oracle_encode_column

import strawboat_b200 as sb
from strawboat_b200.workloads import random_strings

pytestmark = pytest.mark.gpu


def make_columns(rng, n=20000):
    specs = [(sbo.I64, rng.integers(0, 1000, n), rng.random(n) > 0.2), (sbo.F64, rng.integers(0, 65536, n).astype(np.float64), None),
             (sbo.I32, np.cumsum(rng.integers(0, 4, n)).astype(np.int32), None), (sbo.BOOL, rng.random(n) < 0.4, rng.random(n) > 0.1)]
    o, d, v = random_strings(rng, n, 300, 0.3)
    specs.append((sbo.BINARY, (o, d), v))
    o, d, v = random_strings(rng, n, 100000, 0.0, large=True)
    specs.append((sbo.LARGE_BINARY, (o, d), None))
    cols, refs = [], []
    for t, vals, val in specs:
        data, metas = oracle_encode_column(t, vals, val, page_size=4096, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
        cols.append(sb.Column(t, val is not None, data, metas))
        refs.append(oracle_decode_column(t, val is not None, data, metas))
    return cols, refs


def same(ref, t, values, offsets, validity, n):
    if t == sbo.BOOL:
        assert np.array_equal(sbo.unpack_bits(values, n), sbo.unpack_bits(ref["values"], n))
    elif t in (sbo.BINARY, sbo.LARGE_BINARY):
        odt = np.int64 if t == sbo.LARGE_BINARY else np.int32
        assert np.array_equal(offsets.view(odt)[:n + 1], ref["offsets"]) and np.array_equal(values[:len(ref["values"])], ref["values"])
    else:
        assert np.array_equal(values.view(ref["values"].dtype)[:n].view(np.uint8), ref["values"].view(np.uint8))
    if ref["validity"] is not None:
        assert np.array_equal(sbo.unpack_bits(validity, n), sbo.unpack_bits(ref["validity"], n))


def test_plan_then_decode_into_device_buffers(ctx):
    import torch
    rng = np.random.default_rng(0)
    cols, refs = make_columns(rng)
    sizes = ctx.plan_columns(cols)
    for s, r, c in zip(sizes, refs, cols):
        assert s["length"] == r["length"]
        if c.type in (sbo.BINARY, sbo.LARGE_BINARY):
            assert s["values_bytes"] == len(r["values"]) and s["offsets_bytes"] == r["offsets"].nbytes
        assert s["validity_bytes"] == ((r["length"] + 7) // 8 if r["validity"] is not None else 0)
    al = lambda b: (b + 3) // 4 * 4  # noqa: E731
    bufs = [{"values": torch.full((al(s["values_bytes"]) + 16,), 0xAB, dtype=torch.uint8, device="cuda"),
             "offsets": torch.empty(al(s["offsets_bytes"]) + 16, dtype=torch.uint8, device="cuda") if s["offsets_bytes"] else None,
             "validity": torch.empty(al(s["validity_bytes"]) + 16, dtype=torch.uint8, device="cuda") if s["validity_bytes"] else None} for s in sizes]
    dec = ctx.decode_columns_into(cols, bufs, out="device")
    for d, b, r, c in zip(dec, bufs, refs, cols):
        assert d.values_ptr == b["values"].data_ptr()              # written in place, nothing allocated for it
        n = r["length"]
        same(r, c.type, b["values"].cpu().numpy(), None if b["offsets"] is None else b["offsets"].cpu().numpy(),
             None if b["validity"] is None else b["validity"].cpu().numpy(), n)
    dec[0].release()
    assert int(bufs[0]["values"][0]) == int(refs[0]["values"].view(np.uint8)[0])  # the caller's buffer outlives the release
    small = [dict(b) for b in bufs]
    small[1]["values"] = torch.empty(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(sb.StrawboatError) as e:
        ctx.decode_columns_into(cols, small, out="device")
    assert e.value.code == sb._capi.SB_INVALID_ARG


def test_decode_into_pinned_host_buffers(ctx):
    import torch
    rng = np.random.default_rng(1)
    cols, refs = make_columns(rng, 9000)
    sizes = ctx.plan_columns(cols)
    pin = lambda n: torch.empty(max(16, (n + 3) // 4 * 4), dtype=torch.uint8).pin_memory().numpy()  # noqa: E731
    bufs = [{"values": pin(s["values_bytes"]), "offsets": pin(s["offsets_bytes"]) if s["offsets_bytes"] else None,
             "validity": pin(s["validity_bytes"]) if s["validity_bytes"] else None} for s in sizes]
    dec = ctx.decode_columns_into(cols, bufs, out="host")
    for b, r, c in zip(bufs, refs, cols):
        same(r, c.type, b["values"], b["offsets"], b["validity"], r["length"])
    dec[0].release()


def test_async_calls_on_several_contexts(ctx):
    """one host thread, three contexts: submit, submit, submit, then collect -- results equal the synchronous path"""
    rng = np.random.default_rng(2)
    ctxs = [sb.Context(0) for _ in range(3)]
    work = [make_columns(rng, 6000 + 500 * i) for i in range(3)]
    for _ in range(3):
        handles = [c.decode_columns_async(w[0], out="host") for c, w in zip(ctxs, work)]
        [h.ready() for h in handles]
        for h, (cols, refs) in zip(handles, work):
            dec = h.wait()
            assert h.ready()
            for d, r, c in zip(dec, refs, cols):
                same(r, c.type, d.values if d.values is not None else None, d.offsets, d.validity, r["length"])
            dec[0].release()
    # a second submit on the same context collects the first one; so does an encode
    h1 = ctxs[0].decode_columns_async(work[0][0], out="device")
    h2 = ctxs[0].decode_columns_async(work[1][0], out="device")
    ctxs[0].encode_columns([sb.LeafArray(sb.I64, np.arange(100))])
    h2.wait()[0].release()
    bad = bytearray(work[0][0][1]._keep.tobytes())
    bad[0] = 99  # unknown codec in page 0 of the f64 column
    hb = ctxs[1].decode_columns_async([sb.Column(sbo.F64, False, bytes(bad), work[0][0][1].metas)], out="host")
    res = hb.wait(raise_on_page_error=False)
    assert res[0].page_status[0] == sb._capi.SB_OUT_OF_SPEC and res[0].page_status[1] == 0
    res[0].release()
    for c in ctxs:
        c.close()
