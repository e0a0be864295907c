#!/usr/bin/env python
"""bench.py — headline benchmark of claxon_b200 (contract: see the task prompt / DESIGN.md §6).

Metric (BASELINE.json): Msamples/s decoded, bit-exact, samples = sum(block_size * channels).
Workload at N=1: BASELINE.json configs[1] ("c2"): batch of 1024 synthetic stereo 16-bit frames,
block size 4096, LPC order 8, Rice parameter 4, mid/side.  One *step* = one pass of the hot path
(`FrameReader::read_next_or_eof` for every frame of the batch) over one such batch.

  value  — kernel-only throughput, inputs resident in HBM.  Steps are issued round-robin over
           `--inflight` distinct device-resident batches (combined footprint > L2, so no step
           finds its inputs or outputs in L2) on `--streams` CUDA streams, i.e. many batches in
           flight: the steady-state regime of a decode service.  A lone 1024-frame batch is
           latency-bound by the serial LPC recurrence and by the sequential
           window chain of the entropy decode; its figure is reported next to it as `single_batch`.
           A step takes ~15 us, so `--steps K` alone would be a sub-millisecond window: the timed
           region is `repeats` x K steps issued back to back (no drain in between; `repeats` is
           chosen so that the region holds >= --min-steps steps), it is measured three times and
           the median region is reported; ms_per_step = region / (repeats * K).  Every batch's CUDA
           graph is instantiated when the batch is created and every batch is decoded once before
           anything is timed, whatever --warmup says.
  e2e    — same metric through the public host-buffer call (`clx_decode_frames`): per step the
           compressed frames go pinned-host -> device and the full planar i32 PCM comes back.
  roofline — HBM: algorithmic bytes (frame bytes read once + planar i32 written once) / device
           time, against the measured copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline — the CPU oracle (a C restatement of claxon; kind "port") on all host cores.

`--impl reference` times that CPU port alone, same config/metric (the reference itself is Rust and
cannot be built in this image: no rustc).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "Msamples/s decoded (bit-exact)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons; `stop(t0, t1)` keeps the samples taken inside the
    timed region [t0, t1] (wall clock), falling back to the nearest ones when the region is shorter than
    the sampling period."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.time() + 3.0
            while not self.rows and time.time() < t_end:  # wait for the first sample: nvidia-smi starts slowly
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                parsed.append((ts, float(f[1]), float(f[2]), [n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [p for p in parsed if t0 is not None and t0 - 0.02 <= p[0] <= t1 + 0.04]
        note = "inside timed region"
        if not inside and parsed:
            mid = ((t0 or 0) + (t1 or 0)) / 2
            inside = sorted(parsed, key=lambda p: abs(p[0] - mid))[:3]
            note = "timed region shorter than the sampling period: nearest samples"
        sm = [p[1] for p in inside]
        reasons = sorted({n for p in inside for n in p[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": inside[0][2] if inside else None,
                "reasons": reasons, "samples": len(sm), "note": note}


def measured_peak_gbs():
    p = os.path.join(HERE, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram bytes per launch of the decode kernel from the committed ncu capture, if any."""
    p = os.path.join(HERE, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def cpu_model():
    """Model name of the host CPU (for the cpu_baseline record); never raises."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_single_thread_rate(batch, min_seconds=1.0):
    """claxon is single-threaded: the same port on ONE host thread; None on any problem."""
    try:
        v, _, _ = cpu_decode_rate(batch, 1, min_seconds)
        return v
    except Exception:
        return None


def cpu_decode_rate(batch, threads, min_seconds):
    from oracle import oracle as O
    offs, lens, poffs = batch.frame_offsets[:-1], batch.frame_lengths, batch.pcm_offsets[:-1]
    out = np.zeros(batch.n_samples, dtype=np.int32)
    O.decode_batch(batch.data, offs, lens, poffs, batch.n_samples, n_threads=threads, out=out)  # warm
    reps, t0 = 0, time.perf_counter()
    while True:
        bad, _, _ = O.decode_batch(batch.data, offs, lens, poffs, batch.n_samples, n_threads=threads, out=out)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    assert bad == 0 and np.array_equal(out, batch.pcm), "CPU oracle output differs from expected PCM"
    return batch.n_samples * reps / dt / 1e6, reps, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--min-steps", type=int, default=4000, help="the timed region holds at least this many steps (repeats x steps)")
    ap.add_argument("--regions", type=int, default=3, help="timed regions; the median is reported")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="claxon_b200", choices=["claxon_b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--frames", type=int, default=None, help="override frames per batch")
    ap.add_argument("--inflight", type=int, default=64, help="distinct device-resident batches cycled")
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=3.0)
    args = ap.parse_args()

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # more hardware work queues than the default 8, so that the batches in flight really overlap
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "64")
    from claxon_b200 import synth

    cfg = synth.workload_config(args.workload, args.frames)
    cfg.seed += 7919 * rank  # weak scaling: every rank decodes its own batch of the same shape
    config = {"workload": f"{args.workload}: {cfg.n_frames} frames x {cfg.n_channels}ch x bs{cfg.block_size}, "
                          f"{cfg.bps}-bit, LPC order {cfg.lpc_min_order}-{cfg.lpc_max_order}, "
                          f"Rice k={cfg.rice_mode}, stereo_mode={cfg.stereo_mode}",
              "frames_per_step": cfg.n_frames, "parallelism": f"frames sharded over {world} GPU(s), no collective"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        batch = synth.generate(cfg)
        cores = os.cpu_count() or 1
        from oracle import oracle as O
        offs, lens, poffs = batch.frame_offsets[:-1], batch.frame_lengths, batch.pcm_offsets[:-1]
        out = np.zeros(batch.n_samples, dtype=np.int32)
        for _ in range(max(1, args.warmup)):
            O.decode_batch(batch.data, offs, lens, poffs, batch.n_samples, n_threads=cores, out=out)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            bad, _, _ = O.decode_batch(batch.data, offs, lens, poffs, batch.n_samples, n_threads=cores, out=out)
        dt = time.perf_counter() - t0
        ok = bad == 0 and np.array_equal(out, batch.pcm)
        v = batch.n_samples * args.steps / dt / 1e6
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/int64",
            "data": "synthetic", "config": config, "bit_exact": bool(ok),
            "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
                             "sample": f"{args.steps} x full {args.workload} batch ({batch.n_samples} samples)",
                             "note": "C restatement of claxon v0.4.3 (oracle/), frames sharded over threads; "
                                     "claxon itself is Rust and cannot be built here (no rustc)"},
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ GPU arm
    import claxon_b200 as cb

    dist = None
    if world > 1:
        # NCCL prints its version banner on stdout at some debug levels; this script's stdout is one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = cb.Context(device=local, n_streams=max(2, args.streams))
    e2e_ctx = cb.Context(device=local, n_streams=8)  # the host-buffer call pipelines up to 8 chunks per batch
    # distinct batches (different content, same shape) so that the working set exceeds L2
    n_distinct = max(1, args.inflight)
    batches, host = [], []
    alg_bytes = None
    for i in range(n_distinct):
        c = synth.workload_config(args.workload, args.frames)
        c.seed = cfg.seed + 1000003 * i
        b = synth.generate(c)
        descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
        batches.append(ctx.upload(b.data, descs, out_elems))
        if i < 2:
            host.append((b, descs, out_elems))
        if alg_bytes is None:
            alg_bytes = int(b.data.size) + 4 * b.n_samples
            n_samples = b.n_samples
            in_bytes = int(b.data.size)
    footprint_mb = n_distinct * (alg_bytes) / 1e6
    config.update({"inflight_batches": n_distinct, "streams": args.streams,
                   "l2": f"steps cycle over {n_distinct} distinct batches, footprint {footprint_mb:.0f} MB > 126 MB L2"})

    # correctness gate: the timed kernels' output must equal the expected PCM bit for bit
    batches[0].decode(0)
    out, res = batches[0].read()
    b0, d0, _ = host[0]
    exact = bool((res["status"] == 0).all())
    for i in range(b0.n_frames):
        o = int(d0[i]["out_offset"]); lo, hi = int(b0.pcm_offsets[i]), int(b0.pcm_offsets[i + 1])
        if not np.array_equal(out[o:o + hi - lo], b0.pcm[lo:hi]):
            exact = False
            break

    # ---- single-batch (latency regime): one batch, serialised steps, flushing nothing (reported only)
    for _ in range(3):
        batches[0].decode(0); batches[0].sync()
    single = []
    for i in range(10):
        bt = batches[(i + 1) % n_distinct]
        bt.decode(0); bt.sync()
        single.append(bt.kernel_ms())
    single_ms = float(np.median(single))

    # ---- steady state: repeats x K steps back to back, several batches in flight
    repeats = max(1, -(-args.min_steps // max(1, args.steps)))
    timed_steps = repeats * args.steps
    sampler = ClockSampler(local)
    sampler.start()
    ctx.run_steps(batches, max(n_distinct, max(args.warmup, 3)), args.streams)  # every batch once, at least
    barrier()
    region_ms = []
    gpu_launches = 0
    t_wall0 = time.time()
    for _ in range(max(1, args.regions)):
        launches1 = ctx.launch_count
        r_ms = ctx.run_steps(batches, timed_steps, args.streams)
        gpu_launches = ctx.launch_count - launches1
        barrier()
        region_ms.append(max_over_ranks(r_ms))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = float(np.median(region_ms))
    value = n_samples * timed_steps * world / (ms / 1e3) / 1e6

    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes * timed_steps / (ms / 1e3) / 1e9
    traffic = load_traffic()

    # ---- end to end through the host-buffer call, pinned memory
    e2e_steps = args.e2e_steps or max(5, min(args.steps, 30))
    hb, hd, hout_elems = host[0]
    ectx = e2e_ctx
    p_bytes = ctx.host_alloc(int(hb.data.size) + 64)
    p_bytes[: hb.data.size] = hb.data
    p_out = ctx.host_alloc(4 * hout_elems + 64)
    out_view = p_out[: 4 * hout_elems].view(np.int32)
    results = np.zeros(hd.size, dtype=cb.RESULT_DTYPE)
    for _ in range(3):
        ectx.decode_frames_raw(p_bytes.ctypes.data, hb.data.size, hd.ctypes.data, hd.size, p_out.ctypes.data,
                               hout_elems, results.ctypes.data)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ectx.decode_frames_raw(p_bytes.ctypes.data, hb.data.size, hd.ctypes.data, hd.size, p_out.ctypes.data,
                               hout_elems, results.ctypes.data)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    e2e_ok = bool((results["status"] == 0).all())
    for i in range(0, hb.n_frames, max(1, hb.n_frames // 64)):
        o = int(hd[i]["out_offset"]); lo, hi = int(hb.pcm_offsets[i]), int(hb.pcm_offsets[i + 1])
        e2e_ok &= bool(np.array_equal(out_view[o:o + hi - lo], hb.pcm[lo:hi]))
    e2e_value = n_samples * e2e_steps * world / e2e_s / 1e6

    cpu = None
    if rank == 0 and args.gpus == 1 and args.cpu_seconds > 0:
        cores = os.cpu_count() or 1
        v, reps, dt = cpu_decode_rate(hb, cores, args.cpu_seconds)
        cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
               "one_thread": cpu_single_thread_rate(hb),
               "sample": f"{reps} x one full {args.workload} batch ({hb.n_samples} samples) in {dt:.1f}s, "
                         f"frames sharded over {cores} threads (one_thread: the same port on a single thread, >= 1 s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / timed_steps, "repeats": repeats, "timed_steps": timed_steps,
            "region_ms": region_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32 samples / int64 accumulate",
            "data": "synthetic", "config": config, "bit_exact": exact and e2e_ok,
            "clocks": clocks, "gpu_launches": int(gpu_launches),
            "single_batch": {"kernel_ms": single_ms, "value": n_samples / (single_ms / 1e3) / 1e6,
                             "unit": "Msamples/s", "note": "one batch, nothing else in flight (latency regime)"},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "steps": e2e_steps,
                    "h2d_bytes_per_step": int(hb.data.size + hd.nbytes),
                    "d2h_bytes_per_step": int(4 * hout_elems + results.nbytes)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg_bytes, "read_only_gbs": in_bytes * timed_steps / (ms / 1e3) / 1e9,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"),
                         "traffic_source": (traffic or {}).get("source")},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
