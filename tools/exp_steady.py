"""Steady-state step time of device-resident batches, whole path and per pass (measurement build only).

  CLX_EXPERIMENT=1 python tools/exp_steady.py [frames] [batches in flight] [--wl c2] [--passes 3,1,2]

`--passes`: which passes a batch's graph contains (bit 0 = index pass, bit 1 = decode pass); every batch is
decoded once completely first, so a decode-only graph finds its parameter records in place."""
import ctypes as C
import os
import sys
os.environ.setdefault("CLX_EXPERIMENT", "1")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "64")
sys.path.insert(0, ".")
import claxon_b200 as cb
from claxon_b200 import synth, _lib

args = [a for a in sys.argv[1:] if not a.startswith("--")]
frames = int(args[0]) if len(args) > 0 else 1024
nb = int(args[1]) if len(args) > 1 else 64
wl = sys.argv[sys.argv.index("--wl") + 1] if "--wl" in sys.argv else "c2"
passes = [int(x) for x in (sys.argv[sys.argv.index("--passes") + 1] if "--passes" in sys.argv else "3,1,2").split(",")]
L = _lib.load()
ctx = cb.Context(n_streams=max(2, nb))
batches = []
samples = 0
for i in range(nb):
    b = synth.workload(wl, frames, seed=100 + i)
    samples = b.n_samples
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    batches.append(ctx.upload(b.data, descs, out_elems))
for b in batches:
    b.decode(0); b.sync()
res = {}
dyn = int(sys.argv[sys.argv.index("--dyn") + 1]) if "--dyn" in sys.argv else 0
L.clx_exp_set_dyn_smem(dyn)
for which in passes:
    L.clx_exp_set_which(which)
    for b in batches:
        L.clx_exp_rebuild_graph(ctx._h, b._h)
    ctx.run_steps(batches, nb * 4, nb)
    ms = sorted(ctx.run_steps(batches, nb * 32, nb) for _ in range(3))[1]
    us = ms / (nb * 32) * 1000
    res[which] = (round(us, 2), round(samples / us / 1e3, 1))
print("workload", wl, "frames", frames, "in flight", nb, "extra smem", dyn, "{passes: (us/step, Gsamples/s)}", res, flush=True)
