"""One batch, several GPUs, "at most a single NCCL scatter / gather" (BASELINE.json north_star), all on devices:

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_scatter.py [workload] [frames]

Rank 0 generates ONE batch (default: c5, the 8-channel / block 16384 / LPC-32 stress shape) and holds its bytes
in its GPU's memory; `shard.scatter_batch` sends every rank exactly its shard (plan_shards: contiguous frame
ranges balanced on algorithmic bytes) in one grouped NCCL exchange over NVLink; every rank adopts the bytes it
received (device to device, CRC-16 checked on the device), decodes, and `shard.gather_pcm` brings the planar PCM
back to rank 0, which compares it bit for bit with the generator's PCM.  Prints one JSON line with the times
(scatter / decode / gather, each the max over ranks, CUDA events)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import claxon_b200 as cb
from claxon_b200 import synth, shard

workload = sys.argv[1] if len(sys.argv) > 1 else "c5"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", local)
ctx = cb.Context(device=local)
b = descs = None
out_elems = 0
if rank == 0:
    b = synth.workload(workload, frames)
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
meta = [out_elems]
dist.broadcast_object_list(meta, src=0)
out_elems = meta[0]


def timed(fn):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return r, float(t.item())


src_bytes = torch.from_numpy(b.data).to(dev) if rank == 0 else None
for it in range(2):  # first pass warms NCCL up
    (mine, local_descs, (o0, o1)), ms_scatter = timed(lambda: shard.scatter_batch(dist, src_bytes if rank == 0 else None,
                                                                                   descs if rank == 0 else None, src=0, device=dev))
batch = ctx.adopt(mine.data_ptr(), mine.numel(), local_descs, o1 - o0) if local_descs.size else None
t0 = time.perf_counter()
if batch is not None:
    batch.decode(0); batch.sync()
    batch.decode(0); batch.sync()
dist.barrier()
kms = torch.tensor([batch.kernel_ms() if batch is not None else 0.0], device=dev, dtype=torch.float64)
dist.all_reduce(kms, op=dist.ReduceOp.MAX)
ok = True
pcm = torch.zeros(max(1, o1 - o0), dtype=torch.int32, device=dev)
if batch is not None:
    out, res = batch.read()
    ok = bool((res["status"] == 0).all())
    pcm = torch.from_numpy(out[: o1 - o0].copy()).to(dev)
for it in range(2):
    whole, ms_gather = timed(lambda: shard.gather_pcm(dist, pcm, (o0, o1), out_elems, dst=0, device=dev))
okt = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    exact = bool(okt.item()) and out_elems == b.n_samples and bool(torch.equal(whole[:out_elems].cpu(), torch.from_numpy(b.pcm)))
    print(json.dumps({"workload": workload, "frames": frames, "n_gpus": world, "samples": int(b.n_samples), "bit_exact": exact,
                      "scatter_ms": ms_scatter, "scatter_bytes": int(b.data.size), "decode_kernel_ms_max": float(kms.item()),
                      "gather_ms": ms_gather, "gather_bytes": int(4 * out_elems),
                      "decode_msamples_per_s": b.n_samples / (float(kms.item()) / 1e3) / 1e6}))
dist.destroy_process_group()
