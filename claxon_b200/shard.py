"""Multi-GPU sharding of a frame batch.

Frames are independent (reference src/frame.rs:603-605: FrameReader keeps no cross-frame state),
so a batch is split into contiguous frame ranges, one per rank, balanced by algorithmic bytes
(frame bytes in + planar i32 out) rather than by count — frame sizes vary ~10x within one file.
No data-path collective is needed when every rank reads its own shard; `scatter_batch` and
`gather_pcm` are the optional single scatter / gather of BASELINE.json's north_star for the case
where one rank holds all the bytes / wants all the PCM: one grouped exchange of exactly-sized
messages each (NCCL over NVLink on GPUs, gloo on CPU), no padding, nothing staged per destination.
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
    """One scatter of the compressed shards from `src`: a single grouped exchange in which `src` sends every
    other rank exactly its shard's bytes (views of one source tensor: nothing is padded or staged per
    destination).  Returns this rank's (bytes tensor, local descs, (out_lo, out_hi))."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = [None]
    if rank == src:
        plan = plan_shards(descs, world)
        parts = [localize(descs, lo, hi) for lo, hi in plan]
        meta = [[(p[0].tobytes(), p[1], p[2], p[3], p[4]) for p in parts]]
    dist.broadcast_object_list(meta, src=src)
    parts_meta = meta[0]
    mine = parts_meta[rank]
    n_mine = mine[2] - mine[1]
    ops = []
    if rank == src:
        # `data`: a numpy array (copied to `device` once) or a torch tensor already where it should be
        src_t = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(data))
        if device is not None and src_t.device != torch.device(device):
            src_t = src_t.to(device)
        recv = src_t[mine[1]:mine[2]]  # the source's own shard: a view, no copy
        for r, p in enumerate(parts_meta):
            if r != src and p[2] > p[1]:
                ops.append(dist.P2POp(dist.isend, src_t[p[1]:p[2]], r))
    else:
        recv = torch.empty(max(1, n_mine), dtype=torch.uint8, device=device)[:n_mine]
        if n_mine:
            ops.append(dist.P2POp(dist.irecv, recv, src))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    from . import DESC_DTYPE
    local = np.frombuffer(mine[0], dtype=DESC_DTYPE).copy()
    return recv, local, (mine[3], mine[4])


def gather_pcm(dist, pcm, out_range, total_elems: int, dst: int = 0, device=None):
    """The matching single gather: every rank's decoded PCM (a tensor of int32 holding its shard's output
    range `out_range` = (lo, hi) in elements of the whole batch's output) to `dst`, again as one grouped
    exchange of exactly-sized messages.  Returns the whole batch's output on `dst` (None elsewhere)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    ranges = [None] * world
    dist.all_gather_object(ranges, (int(out_range[0]), int(out_range[1])))
    ops, whole = [], None
    if rank == dst:
        whole = torch.zeros(max(1, total_elems), dtype=torch.int32, device=device)
        lo, hi = ranges[dst]
        whole[lo:hi] = pcm[: hi - lo]
        for r, (lo, hi) in enumerate(ranges):
            if r != dst and hi > lo:
                ops.append(dist.P2POp(dist.irecv, whole[lo:hi], r))
    else:
        lo, hi = ranges[rank]
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, pcm[: hi - lo].contiguous(), dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return whole
