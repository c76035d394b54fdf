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
