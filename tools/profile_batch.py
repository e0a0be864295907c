"""Decodes one device-resident batch a few times (target for ncu): python tools/profile_batch.py <workload> <frames> <reps> [lane].
`lane`: force the lane-per-frame throughput path (it is what device-resident batches use anyway)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import claxon_b200 as cb
from claxon_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b = synth.workload(name, n)
ctx = cb.Context(lane_per_frame=len(sys.argv) > 4 and sys.argv[4] == 'lane')
descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
dev = ctx.upload(b.data, descs, out_elems)
for i in range(reps):
    dev.decode(0)
    dev.sync()
    print("kernel ms", dev.kernel_ms(), flush=True)
out, res = dev.read()
exact = bool(np.array_equal(out[:b.n_samples], b.pcm)) if out_elems == b.n_samples else "n/a"
print("ok", bool((res["status"] == 0).all()), exact)
