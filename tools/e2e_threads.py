"""Do two host threads calling the synchronous host-buffer call overlap?  Per-thread call times and the pair's wall time."""
import os, sys, time, threading
sys.path.insert(0, ".")
import numpy as np
import claxon_b200 as cb
from claxon_b200 import synth
b = synth.workload("c2", 1024)
descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
for verify in (True, False):
    for mode in (cb.OUT_PLANAR_I32, cb.OUT_INTERLEAVED_I16):
        slots = []
        for c in range(2):
            ctx = cb.Context(n_streams=8, verify_crc=verify, host_threads=16)
            pb = ctx.host_alloc(int(b.data.size) + 64); pb[: b.data.size] = b.data
            slots.append((ctx, pb, ctx.host_alloc(4 * out_elems + 64), np.zeros(descs.size, dtype=cb.RESULT_DTYPE), []))
        def worker(slot, n):
            ctx, pb, po, rs, times = slot
            for _ in range(n):
                t0 = time.perf_counter()
                ctx.decode_frames_raw(pb.ctypes.data, b.data.size, descs.ctypes.data, descs.size, po.ctypes.data, out_elems, rs.ctypes.data, mode)
                times.append(time.perf_counter() - t0)
        for s in slots: worker(s, 3)
        for s in slots: s[4].clear()
        t0 = time.perf_counter(); worker(slots[0], 20); one = (time.perf_counter() - t0) / 20
        for s in slots: s[4].clear()
        th = [threading.Thread(target=worker, args=(s, 20)) for s in slots]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        two = (time.perf_counter() - t0) / 40
        print("verify", verify, "mode", mode, "one caller ms/call", round(one * 1e3, 3), "two callers ms/call", round(two * 1e3, 3),
              "per-thread call ms", [round(float(np.median(s[4])) * 1e3, 3) for s in slots], flush=True)
