"""Generates tests/golden/pages_v1.npz: encoded pages of every codec family + their decoded
Arrow buffers.  The pages are written and decoded by the CPU oracle (oracle/, the restatement of
the reference path -- the Rust reference itself cannot run in this image), so the fixture freezes
today's oracle behaviour: later oracle or kernel changes that alter any byte are caught on CPU
(test_golden.py::test_oracle_*) and on the GPU (test_golden.py::test_gpu_*), and the GPU test
needs no oracle at run time.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sbo  # noqa: E402
from helpers import oracle_decode_column, oracle_encode_column  # noqa: E402

from strawboat_b200.workloads import random_strings  # noqa: E402


def cases():
    rng = np.random.default_rng(20260101)
    n = 3000
    lz, none = sbo.C_LZ4, sbo.C_NONE
    yield "i64_plain", sbo.I64, rng.integers(-2**62, 2**62, n), None, dict(default=none)
    yield "i64_lz4_null", sbo.I64, rng.integers(0, 50, n), rng.random(n) > 0.3, dict(default=lz)
    yield "i32_dict_bp", sbo.I32, rng.integers(0, 8, 2048 * 2).astype(np.int32), None, dict(default=lz, ratio=2.0)
    yield "i32_delta", sbo.I32, np.cumsum(rng.integers(0, 4, 2048 * 2)).astype(np.int32), None, dict(default=lz, ratio=2.0)
    yield "u32_freq", sbo.U32, np.tile(np.array([20] * 2045 + [10000] * 3, np.uint32), 2), None, dict(default=lz, ratio=2.0)
    yield "i64_rle_null", sbo.I64, np.repeat(rng.integers(0, 1 << 40, n // 50), 50), rng.random(n) > 0.1, dict(default=lz, force=sbo.C_RLE)
    yield "u16_onevalue", sbo.U16, np.full(n, 9, np.uint16), None, dict(default=lz, ratio=2.0)
    yield "f64_dict", sbo.F64, rng.integers(0, 8, n).astype(np.float64), rng.random(n) > 0.2, dict(default=lz, ratio=2.0)
    yield "f64_patas", sbo.F64, np.cumsum(rng.integers(-3, 4, n)) * 0.5, None, dict(default=none, force=sbo.C_PATAS)
    yield "f32_freq", sbo.F32, np.where(rng.random(n) < 0.95, 1.5, rng.standard_normal(n)).astype(np.float32), None, dict(default=none, force=sbo.C_FREQ)
    yield "bool_plain", sbo.BOOL, rng.random(n) < 0.5, rng.random(n) > 0.2, dict(default=none)
    yield "bool_rle", sbo.BOOL, np.repeat(rng.random(30) < 0.5, 100), None, dict(default=lz, ratio=2.0)
    o, d, v = random_strings(rng, n, 50, 0.3)
    yield "utf8_dict", sbo.BINARY, (o, d), v, dict(default=lz, ratio=2.0)
    o, d, v = random_strings(rng, n, 2000, 0.0, large=True)
    yield "large_binary_lz4", sbo.LARGE_BINARY, (o, d), None, dict(default=lz)
    o, d, v = random_strings(rng, n, 30, 0.2)
    yield "utf8_freq", sbo.BINARY, (o, d), v, dict(default=none, force=sbo.C_FREQ)


def main():
    out = {}
    names = []
    for name, t, vals, validity, o in cases():
        opts = sbo.make_opts(o.get("default", 0), ratio=o.get("ratio"), force=o.get("force", -1))
        data, metas = oracle_encode_column(t, vals, validity, page_size=1024, opts=opts, seed=1)
        ref = oracle_decode_column(t, validity is not None, data, metas)
        names.append(name)
        out[name + ".type"] = np.array([t, int(validity is not None)])
        out[name + ".data"] = np.frombuffer(data, np.uint8)
        out[name + ".metas"] = np.array(metas, np.uint64)
        out[name + ".values"] = ref["values"].view(np.uint8)
        if "offsets" in ref:
            out[name + ".offsets"] = ref["offsets"].astype(np.int64)
        if ref["validity"] is not None:
            out[name + ".validity"] = ref["validity"]
        out[name + ".tree"] = np.array([sbo.stat_page(t, validity is not None, data[:metas[0][0]])])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "pages_v1.npz"), **out)
    print("wrote", len(names), "cases,", os.path.getsize(os.path.join(HERE, "pages_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
