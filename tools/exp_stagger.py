"""Does the steady state run the two kernels in lock step?  Same work, but every stream's FIRST step is a
batch of a different (small) size, so the streams drift apart before the timed steps."""
import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "64")
sys.path.insert(0, ".")
import numpy as np
import claxon_b200 as cb
from claxon_b200 import synth
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = cb.Context(n_streams=128)
def up(frames, seed):
    b = synth.workload("c2", frames, seed=seed)
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    return ctx.upload(b.data, descs, out_elems)
normal = [up(1024, 100 + i) for i in range(streams)]
rng = np.random.default_rng(1)
small = [up(int(rng.integers(1, 32)) * 32, 900 + i) for i in range(streams)]
for b in normal + small:
    b.decode(0); b.sync()
rows = 24
plain = normal * rows
ctx.run_steps(plain, len(plain), streams)
ms0 = ctx.run_steps(plain, len(plain), streams)
stag = small + normal * rows
ctx.run_steps(stag, len(stag), streams)
ms1 = ctx.run_steps(stag, len(stag), streams)
print("streams", streams, "lockstep us/step", round(ms0 / len(plain) * 1000, 2), "staggered us/step (incl. the small first row)",
      round(ms1 / (len(plain)) * 1000, 2), flush=True)
