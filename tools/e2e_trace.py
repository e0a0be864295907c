"""Breakdown of the synchronous host-buffer call (CLX_TRACE=1 prints submit / crc / wait / apply per call)."""
import os, sys, time
os.environ["CLX_TRACE"] = "1"
sys.path.insert(0, ".")
import numpy as np
import claxon_b200 as cb
from claxon_b200 import synth
b = synth.workload("c2", 1024)
descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
for kw in ({}, {"lane_per_frame": True}):
    ctx = cb.Context(n_streams=8, **kw)
    p_bytes = ctx.host_alloc(int(b.data.size) + 64); p_bytes[: b.data.size] = b.data
    p_out = ctx.host_alloc(4 * out_elems + 64)
    res = np.zeros(descs.size, dtype=cb.RESULT_DTYPE)
    print("ctx", kw, flush=True)
    for i in range(6):
        t0 = time.perf_counter()
        ctx.decode_frames_raw(p_bytes.ctypes.data, b.data.size, descs.ctypes.data, descs.size, p_out.ctypes.data, out_elems, res.ctypes.data)
        print("  call ms", round((time.perf_counter() - t0) * 1e3, 3), flush=True)
