"""Steady-state step time of device-resident C2 batches as a function of the batches in flight."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ".")
import claxon_b200 as cb
from claxon_b200 import synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
wl = sys.argv[sys.argv.index("--wl") + 1] if "--wl" in sys.argv else "c2"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ctx = cb.Context(n_streams=128, warp_per_frame=('--warp' in sys.argv))
batches = []
for i in range(nb):
    b = synth.workload(wl, frames, seed=100 + i)
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    batches.append(ctx.upload(b.data, descs, out_elems))
for b in batches:
    b.decode(0); b.sync()
if os.environ.get("CLX_SEQ_DEBUG"):
    from claxon_b200 import _lib
    _lib.load().clx_debug_seq_flags(int(os.environ["CLX_SEQ_DEBUG"]))
res = {}
for streams in ((64,) if '--only64' in sys.argv else (32,) if '--only32' in sys.argv else (8, 16, 32, 48, 64, 96, 128)):
    if streams > nb:
        break
    ctx.run_steps(batches[:streams], streams * 4, streams)
    ms = ctx.run_steps(batches[:streams], streams * 24, streams)
    us = ms / (streams * 24) * 1000
    res[streams] = (round(us, 2), round(frames * 8192 / us / 1e3, 1))
print("workload", wl, "frames", frames, "max_connections", os.environ["CUDA_DEVICE_MAX_CONNECTIONS"], "{in flight: (us/step, Gsamples/s)}", res, flush=True)
