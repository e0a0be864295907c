"""Decodes one device-resident batch a few times (target for ncu)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import claxon_b200 as cb
from claxon_b200 import synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else None
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b = synth.workload(name, n)
ctx = cb.Context()
descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
dev = ctx.upload(b.data, descs, out_elems)
for i in range(reps):
    dev.decode(0)
    dev.sync()
    print("kernel ms", dev.kernel_ms())
out, res = dev.read()
print("ok", bool((res["status"] == 0).all()), bool(np.array_equal(out[:b.n_samples], b.pcm)) if out_elems == b.n_samples else "n/a")

import ctypes
from claxon_b200 import _lib
L = _lib.load()
st = (ctypes.c_ulonglong * 16)()
L.clx_debug_coop_stats(st, 1)
w = max(1, st[0])
if st[0]: print("coop stats: windows", st[0], "rounds/window", st[1]/w, "spec trips/window", st[2]/w, "merge trips/round", st[3]/max(1,st[1]), "codes/window", st[4]/w)

if st[0]: print("cycles/window: spec", st[5]/w, "rounds", st[6]/w, "ranks", st[7]/w, "emit", st[8]/w, "advance", st[9]/w, "| phase1 cycles/warp", st[10]/max(1, 1024*reps), "windows/warp", w/(1024*reps))

if st[13]: print("predict calls", st[13], "narrow", st[12], "cycles/call", st[11]/max(1, st[13]))
