"""Committed golden fixture (tests/golden/pages_v1.npz, made by tests/golden/make_golden.py):
encoded pages of every codec family + their decoded buffers.  CPU: the oracle still produces
exactly these bytes.  GPU: the CUDA decoder reproduces them without any oracle at run time."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pages_v1.npz"))
NAMES = [str(x) for x in G["names"]]


def unpack(bitmap, n):
    return np.unpackbits(np.asarray(bitmap, np.uint8), bitorder="little")[:n]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_decodes_golden_pages(name):
    import sbo
    from helpers import oracle_decode_column
    t, nullable = (int(x) for x in G[name + ".type"])
    metas = [tuple(int(x) for x in m) for m in G[name + ".metas"]]
    ref = oracle_decode_column(t, bool(nullable), G[name + ".data"].tobytes(), metas)
    n = ref["length"]
    if t == sbo.BOOL:
        assert np.array_equal(unpack(ref["values"], n), unpack(G[name + ".values"], n))
    else:
        assert np.array_equal(ref["values"].view(np.uint8), G[name + ".values"])
    if name + ".offsets" in G:
        assert np.array_equal(ref["offsets"].astype(np.int64), G[name + ".offsets"])
    if name + ".validity" in G:
        assert np.array_equal(unpack(ref["validity"], n), unpack(G[name + ".validity"], n))
    assert sbo.stat_page(t, bool(nullable), G[name + ".data"].tobytes()[:metas[0][0]]) == str(G[name + ".tree"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_decodes_golden_pages(ctx, name):
    import strawboat_b200 as sb
    t, nullable = (int(x) for x in G[name + ".type"])
    metas = [tuple(int(x) for x in m) for m in G[name + ".metas"]]
    dec = ctx.batch_read_array(sb.Column(t, bool(nullable), G[name + ".data"].tobytes(), metas))
    n = dec.length
    assert n == sum(m[1] for m in metas)
    if t == sb.BOOL:
        assert np.array_equal(unpack(dec.values, n), unpack(G[name + ".values"], n))
    else:
        assert np.array_equal(dec.values.view(np.uint8), G[name + ".values"])
    if name + ".offsets" in G:
        assert np.array_equal(dec.offsets.astype(np.int64), G[name + ".offsets"])
    if name + ".validity" in G:
        assert np.array_equal(unpack(dec.validity, n), unpack(G[name + ".validity"], n))
