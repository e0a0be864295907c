"""Breakdown of the synchronous host-buffer call (measurement build: CLX_TRACE=1 prints submit / crc / wait / apply
per call), planar i32 and interleaved i16 output, plus the PCIe copy rates the box can do at all."""
import os, sys, time
os.environ["CLX_TRACE"] = "1"
os.environ.setdefault("CLX_EXPERIMENT", "1")
sys.path.insert(0, ".")
import numpy as np
import torch
import claxon_b200 as cb
from claxon_b200 import synth
b = synth.workload("c2", 1024)
descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
h = torch.empty(32 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    print(name, "32 MiB pinned copy GB/s", round(10 * 32 * 1.048576e-3 / (time.perf_counter() - t0), 1), flush=True)
for kw in ({}, {"lane_per_frame": True}):
    for mode in (cb.OUT_PLANAR_I32, cb.OUT_INTERLEAVED_I16):
        ctx = cb.Context(n_streams=8, **kw)
        p_bytes = ctx.host_alloc(int(b.data.size) + 64); p_bytes[: b.data.size] = b.data
        p_out = ctx.host_alloc(4 * out_elems + 64)
        res = np.zeros(descs.size, dtype=cb.RESULT_DTYPE)
        print("ctx", kw, "mode", mode, flush=True)
        for i in range(5):
            t0 = time.perf_counter()
            ctx.decode_frames_raw(p_bytes.ctypes.data, b.data.size, descs.ctypes.data, descs.size, p_out.ctypes.data, out_elems, res.ctypes.data, mode)
            print("  call ms", round((time.perf_counter() - t0) * 1e3, 3), flush=True)
