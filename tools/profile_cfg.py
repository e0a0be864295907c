import sys
sys.path.insert(0, ".")
import numpy as np, claxon_b200 as cb
from claxon_b200 import synth
name, n = sys.argv[1], int(sys.argv[2])
b = synth.workload(name, n)
ctx = cb.Context()
descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
dev = ctx.upload(b.data, descs, out_elems)
for i in range(2):
    dev.decode(0); dev.sync()
print("ms", dev.kernel_ms())
