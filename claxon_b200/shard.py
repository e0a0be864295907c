"""Multi-GPU sharding of a frame batch.

Frames are independent (reference src/frame.rs:603-605: FrameReader keeps no cross-frame state),
so a batch is split into contiguous frame ranges, one per rank, balanced by algorithmic bytes
(frame bytes in + planar i32 out) rather than by count — frame sizes vary ~10x within one file.
No data-path collective is needed when every rank reads its own shard; `scatter_batch` /
`gather_pcm` are the optional single scatter / gather of BASELINE.json's north_star for the case
where rank 0 holds all the bytes.
"""
from __future__ import annotations

import numpy as np


def frame_costs(descs: np.ndarray) -> np.ndarray:
    out = descs["n_channels"].astype(np.uint64) * descs["block_size"].astype(np.uint64) * 4
    return descs["byte_len"].astype(np.uint64) + out


def plan_shards(descs: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous [lo, hi) frame ranges per rank with near-equal cumulative cost."""
    n = int(descs.size)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(world, 1) - 1)
    cum = np.cumsum(frame_costs(descs).astype(np.float64))
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        i = int(np.searchsorted(cum, target, side="left")) + 1
        i = min(max(i, bounds[-1]), n)
        bounds.append(i)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def localize(descs: np.ndarray, lo: int, hi: int):
    """Rebases a shard's descriptors to its own byte / output ranges.
    Returns (local descs, byte_lo, byte_hi, out_lo, out_hi)."""
    d = descs[lo:hi].copy()
    if d.size == 0:
        return d, 0, 0, 0, 0
    b0 = int(d["byte_offset"].min()) & ~15
    b1 = int((d["byte_offset"] + d["byte_len"]).max())
    o0 = int(d["out_offset"].min()) & ~3
    o1 = int((d["out_offset"] + d["n_channels"].astype(np.uint64) * d["block_size"]).max())
    d["byte_offset"] -= np.uint64(b0)
    d["out_offset"] -= np.uint64(o0)
    return d, b0, b1, o0, o1


def scatter_batch(dist, data, descs, src: int = 0, device=None):
    """One scatter of the compressed shards from `src` (torch.distributed; NCCL over NVLink on GPUs,
    gloo on CPU). Returns this rank's (bytes tensor, local descs, out range)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        plan = plan_shards(descs, world)
        parts = [localize(descs, lo, hi) for lo, hi in plan]
        meta = [[(p[0].tobytes(), p[1], p[2], p[3], p[4]) for p in parts]]
    dist.broadcast_object_list(meta, src=src)
    parts_meta = meta[0]
    maxlen = max(p[2] - p[1] for p in parts_meta)
    maxlen = (maxlen + 15) & ~15
    recv = torch.zeros(max(16, maxlen), dtype=torch.uint8, device=device)
    chunks = None
    if rank == src:
        src_t = torch.as_tensor(np.ascontiguousarray(data))
        chunks = []
        for p in parts_meta:
            c = torch.zeros(max(16, maxlen), dtype=torch.uint8, device=device)
            c[: p[2] - p[1]] = src_t[p[1]:p[2]].to(device) if device is not None else src_t[p[1]:p[2]]
            chunks.append(c)
    dist.scatter(recv, chunks, src=src)
    mine = parts_meta[rank]
    from . import DESC_DTYPE
    local = np.frombuffer(mine[0], dtype=DESC_DTYPE).copy()
    return recv[: mine[2] - mine[1]], local, (mine[3], mine[4])
