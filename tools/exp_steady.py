"""Steady-state step time of device-resident C2 batches (experiment build: CLX_DEBUG_SKIP leaves kernels out)."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ".")
import claxon_b200 as cb
from claxon_b200 import synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = cb.Context(n_streams=64)
batches = []
for i in range(32):
    b = synth.workload("c2", frames, seed=100 + i)
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    batches.append(ctx.upload(b.data, descs, out_elems))
for b in batches:
    b.decode(0); b.sync()
res = {}
for streams in (4, 8, 16, 32, 64):
    ctx.run_steps(batches, 64, streams)
    ms = ctx.run_steps(batches, 640, streams)
    res[streams] = round(ms / 640 * 1000, 2)
print("skip", os.environ.get("CLX_DEBUG_SKIP", "0"), "frames", frames, "us/step by streams", res, flush=True)
