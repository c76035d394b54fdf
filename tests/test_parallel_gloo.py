"""CPU, world_size 2 over gloo: the multi-rank host logic of the encode path --
columns sharded `leaf_idx mod world`, encoded bodies gathered to the writer rank, one file with
absolute ColumnMeta offsets.  The page bytes come from the oracle writer here (no GPU in this
container); on a GPU box the same plumbing carries Context.encode_columns output over NCCL."""
import os
import socket

import numpy as np
import sbo
import torch.multiprocessing as mp
from helpers import oracle_decode_column, oracle_encode_column

N_COLS = 5


def make_cols():
    rng = np.random.default_rng(0)
    cols = []
    for c in range(N_COLS):
        t = [sbo.I64, sbo.I32, sbo.F64, sbo.I64, sbo.I32][c]
        cols.append((t, rng.integers(0, 1000, 3000 + 100 * c).astype(sbo.NP_OF[t])))
    return cols


def worker(rank, world, port, out_path):
    import torch.distributed as dist

    from strawboat_b200 import fileio, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cols = make_cols()
    local = {}
    for c in parallel.shard_columns(N_COLS, world, rank):
        t, v = cols[c]
        local[c] = oracle_encode_column(t, v, None, page_size=1024, opts=sbo.make_opts(sbo.C_LZ4, ratio=2.0))
    gathered = parallel.gather_encoded(local, N_COLS, dst=0)
    if rank == 0:
        data, _ = fileio.write_file(gathered)
        with open(out_path, "wb") as f:
            f.write(data)
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_encode_gather_to_writer(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "t.str")
    mp.spawn(worker, args=(2, port, out), nprocs=2, join=True)
    from strawboat_b200 import fileio
    data = open(out, "rb").read()
    metas = fileio.read_meta(data)
    assert len(metas) == N_COLS
    prev_end = 8
    for (t, v), cm in zip(make_cols(), metas):
        assert cm[0] == prev_end  # column-major, absolute offsets, no gaps
        prev_end += sum(p[0] for p in cm[1])
        ref = oracle_decode_column(t, False, fileio.column_body(data, cm), cm[1])
        assert np.array_equal(ref["values"], v)
