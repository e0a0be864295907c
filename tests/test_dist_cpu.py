"""world_size-2 gloo test of the multi-GPU host logic (shard plan, the optional scatter of
compressed shards, reassembly of per-rank PCM).  No GPU here: the per-rank decode is done by the
CPU oracle purely as a stand-in checker, the thing under test is the sharding / exchange plumbing."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import claxon_b200 as cb
    from claxon_b200 import synth, shard
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = synth.workload("c4", 66, seed=1234)  # same on every rank; only rank 0's copy is used as source
        descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
        mine, local, (o0, o1) = shard.scatter_batch(dist, b.data if rank == 0 else None, descs if rank == 0 else None)
        mine = mine.numpy()
        # stand-in decode of the local shard
        bad, st, pcm = O.decode_batch(mine, local["byte_offset"], local["byte_len"], local["out_offset"], o1 - o0)
        assert bad == 0
        # every rank's frames are exactly its planned contiguous range
        plan = shard.plan_shards(descs, world)
        lo, hi = plan[rank]
        assert local.size == hi - lo
        whole_bad, _, whole = O.decode_batch(b.data, descs["byte_offset"], descs["byte_len"], descs["out_offset"], out_elems)
        ok = True
        for i in range(local.size):
            g = descs[lo + i]
            n = int(g["n_channels"]) * int(g["block_size"])
            a = pcm[int(local[i]["out_offset"]):int(local[i]["out_offset"]) + n]
            e = whole[int(g["out_offset"]):int(g["out_offset"]) + n]
            ok &= bool(np.array_equal(a, e))
        # the matching gather: rank 0 ends up with the whole batch's PCM
        got = shard.gather_pcm(dist, torch.from_numpy(pcm[: o1 - o0].copy()), (o0, o1), out_elems, dst=0)
        if rank == 0:
            for g in descs:
                n = int(g["n_channels"]) * int(g["block_size"])
                o = int(g["out_offset"])
                ok &= bool(np.array_equal(got.numpy()[o:o + n], whole[o:o + n]))
        else:
            ok &= got is None
        # gather of sample counts: totals must add up
        t = torch.tensor([int(sum(int(d["n_channels"]) * int(d["block_size"]) for d in local))])
        dist.all_reduce(t)
        ok &= int(t.item()) == b.n_samples
        q.put((rank, ok, local.size))
    finally:
        dist.destroy_process_group()


def test_scatter_and_shard_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == 66
