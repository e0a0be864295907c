"""Steady-state step time of resident C2 batches with the frame CRC-16 check on the device inside every decode
(batches adopted from device memory) against batches whose CRC verdicts were computed once on the host at creation.

  python tools/exp_crc_cost.py [frames] [batches in flight]"""
import os
import sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "64")
sys.path.insert(0, ".")
import torch
import claxon_b200 as cb
from claxon_b200 import synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = cb.Context(n_streams=nb)
host, dev, keep = [], [], []
samples = 0
for i in range(nb):
    b = synth.workload("c2", frames, seed=100 + i)
    samples = b.n_samples
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    host.append(ctx.upload(b.data, descs, out_elems))
    t = torch.from_numpy(b.data).cuda()
    keep.append(t)
    dev.append(ctx.adopt(t.data_ptr(), t.numel(), descs, out_elems))
for name, bs in (("host CRC at creation", host), ("device CRC in every decode", dev), ("host CRC at creation", host), ("device CRC in every decode", dev)):
    ctx.run_steps(bs, nb * 4, nb)
    ms = sorted(ctx.run_steps(bs, nb * 32, nb) for _ in range(3))[1]
    us = ms / (nb * 32) * 1000
    print(f"{name}: {us:.2f} us/step, {samples / us / 1e3:.1f} Gsamples/s", flush=True)
