"""Synthetic FLAC frame batches (forced-parameter mini-encoder, csrc/synth.c).

``generate(SynthConfig(...))`` returns the byte stream, the frame offset table and
the PCM the frames decode to *by construction*.  ``workload(name)`` gives the
BASELINE.json configurations C2..C5.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
from dataclasses import dataclass, field, asdict

import numpy as np

from . import _build


class _Cfg(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_frames", C.c_uint32), ("block_size", C.c_uint32),
        ("tail_block_size", C.c_uint32), ("frames_per_file", C.c_uint32),
        ("n_channels", C.c_uint32), ("bps", C.c_uint32), ("sample_rate_code", C.c_uint32),
        ("stereo_mode", C.c_int32), ("type_mask", C.c_uint32),
        ("lpc_min_order", C.c_uint32), ("lpc_max_order", C.c_uint32),
        ("fixed_min_order", C.c_uint32), ("fixed_max_order", C.c_uint32),
        ("qlp_precision", C.c_uint32), ("rice_mode", C.c_int32),
        ("rice_kmin", C.c_uint32), ("rice_kmax", C.c_uint32),
        ("min_porder", C.c_uint32), ("max_porder", C.c_uint32), ("rice2", C.c_uint32),
        ("wasted_max", C.c_uint32), ("residual_mean", C.c_double),
        ("variable_blocking", C.c_uint32), ("long_unary_per_mille", C.c_uint32),
        ("force_bs16", C.c_uint32),
    ]


TYPE_CONSTANT, TYPE_VERBATIM, TYPE_FIXED, TYPE_LPC = 1, 2, 4, 8
INDEPENDENT, LEFT_SIDE, RIGHT_SIDE, MID_SIDE, RANDOM_STEREO = 0, 8, 9, 10, -1


@dataclass
class SynthConfig:
    seed: int = 0xC1A00002
    n_frames: int = 16
    block_size: int = 4096
    tail_block_size: int = 0
    frames_per_file: int = 0
    n_channels: int = 2
    bps: int = 16
    sample_rate_code: int = 9          # 44.1 kHz
    stereo_mode: int = INDEPENDENT
    type_mask: int = TYPE_LPC
    lpc_min_order: int = 8
    lpc_max_order: int = 8
    fixed_min_order: int = 0
    fixed_max_order: int = 4
    qlp_precision: int = 12
    rice_mode: int = 4                 # >=0 forced k, -1 optimal, -2 k0 in [kmin,kmax] +-1
    rice_kmin: int = 0
    rice_kmax: int = 14
    min_porder: int = 0
    max_porder: int = 0
    rice2: int = 0
    wasted_max: int = 0
    residual_mean: float = 0.0
    variable_blocking: int = 0
    long_unary_per_mille: int = 0
    force_bs16: int = 0


@dataclass
class SynthBatch:
    config: SynthConfig
    data: np.ndarray            # uint8 stream, frames back to back
    frame_offsets: np.ndarray   # uint64 [n_frames + 1]
    pcm: np.ndarray             # int32, planar per frame, back to back (expected decode)
    pcm_offsets: np.ndarray     # uint64 [n_frames + 1] element offsets
    meta: dict = field(default_factory=dict)

    @property
    def n_frames(self) -> int:
        return int(self.frame_offsets.size - 1)

    @property
    def n_samples(self) -> int:
        return int(self.pcm_offsets[-1])

    @property
    def frame_lengths(self) -> np.ndarray:
        return np.diff(self.frame_offsets).astype(np.uint32)


_lib = None


def _synth_lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build_synth())
        L.clxs_generate_mt.restype = C.c_void_p
        L.clxs_generate_mt.argtypes = [C.POINTER(_Cfg), C.c_int]
        for name, res in (("clxs_bytes", C.c_void_p), ("clxs_nbytes", C.c_uint64),
                          ("clxs_frame_offsets", C.c_void_p), ("clxs_pcm", C.c_void_p),
                          ("clxs_pcm_offsets", C.c_void_p), ("clxs_n_samples", C.c_uint64)):
            getattr(L, name).restype = res
            getattr(L, name).argtypes = [C.c_void_p]
        L.clxs_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _copy(ptr: int, count: int, dtype) -> np.ndarray:
    if count == 0:
        return np.zeros(0, dtype=dtype)
    nbytes = count * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def generate(cfg: SynthConfig, n_threads: int | None = None) -> SynthBatch:
    """Deterministic in (cfg): frame i depends only on cfg.seed + i, whatever the thread count."""
    L = _synth_lib()
    c = _Cfg(**asdict(cfg))
    if n_threads is None:
        n_threads = min(32, os.cpu_count() or 1)
    h = L.clxs_generate_mt(C.byref(c), n_threads)
    if not h:
        raise ValueError("invalid synth config")
    try:
        n = cfg.n_frames
        data = _copy(L.clxs_bytes(h), L.clxs_nbytes(h), np.uint8)
        offs = _copy(L.clxs_frame_offsets(h), n + 1, np.uint64)
        pcm = _copy(L.clxs_pcm(h), L.clxs_n_samples(h), np.int32)
        poffs = _copy(L.clxs_pcm_offsets(h), n + 1, np.uint64)
    finally:
        L.clxs_free(h)
    return SynthBatch(cfg, data, offs, pcm, poffs)


# ---------------------------------------------------------------------------
# BASELINE.json workloads
# ---------------------------------------------------------------------------

def workload_config(name: str, n_frames: int | None = None, seed: int | None = None) -> SynthConfig:
    """C2..C5 of BASELINE.json; `n_frames` overrides the batch size (tests use small ones)."""
    name = name.lower()
    if name in ("c2", "c2-ms"):
        cfg = SynthConfig(seed=0xC1A00002, n_frames=1024, block_size=4096, n_channels=2, bps=16,
                          sample_rate_code=9, stereo_mode=MID_SIDE, type_mask=TYPE_LPC,
                          lpc_min_order=8, lpc_max_order=8, qlp_precision=12, rice_mode=4,
                          max_porder=0, residual_mean=11.5)
    elif name == "c2-indep":
        cfg = workload_config("c2")
        cfg.stereo_mode = INDEPENDENT
    elif name == "c3":
        cfg = SynthConfig(seed=0xC1A00003, n_frames=8192, block_size=4096, n_channels=2, bps=24,
                          sample_rate_code=11, stereo_mode=RANDOM_STEREO,
                          type_mask=TYPE_FIXED | TYPE_LPC, lpc_min_order=1, lpc_max_order=12,
                          fixed_min_order=1, fixed_max_order=4, qlp_precision=15, rice_mode=-1,
                          rice_kmin=8, rice_kmax=14, min_porder=0, max_porder=4)
    elif name == "c4":
        cfg = SynthConfig(seed=0xC1A00004, n_frames=110000, block_size=4096, tail_block_size=3140,
                          frames_per_file=11, n_channels=2, bps=16, sample_rate_code=9,
                          stereo_mode=RANDOM_STEREO, type_mask=TYPE_FIXED | TYPE_LPC,
                          lpc_min_order=1, lpc_max_order=12, fixed_min_order=0, fixed_max_order=4,
                          qlp_precision=0, rice_mode=-2, rice_kmin=0, rice_kmax=14,
                          min_porder=0, max_porder=5, force_bs16=0)
    elif name == "c5":
        cfg = SynthConfig(seed=0xC1A00005, n_frames=4096, block_size=16384, n_channels=8, bps=24,
                          sample_rate_code=11, stereo_mode=INDEPENDENT, type_mask=TYPE_LPC,
                          lpc_min_order=32, lpc_max_order=32, qlp_precision=15, rice_mode=-1,
                          rice_kmin=8, rice_kmax=14, min_porder=0, max_porder=9)
    else:
        raise KeyError(name)
    if n_frames is not None:
        cfg.n_frames = n_frames
    if seed is not None:
        cfg.seed = seed
    return cfg


def workload(name: str, n_frames: int | None = None, seed: int | None = None) -> SynthBatch:
    return generate(workload_config(name, n_frames, seed))


# ---------------------------------------------------------------------------
# file mode (C4): 'fLaC' + STREAMINFO + frames, self-verifying through the MD5
# ---------------------------------------------------------------------------

def interleaved_le_bytes(pcm_planar: np.ndarray, n_channels: int, bps: int) -> bytes:
    """Interleaved little-endian PCM as hashed by STREAMINFO's MD5 (reference src/metadata.rs:52-53)."""
    x = pcm_planar.reshape(n_channels, -1).T
    nb = (bps + 7) // 8
    return np.ascontiguousarray(x.astype("<i4")).view(np.uint8).reshape(-1, 4)[:, :nb].tobytes()


def make_file(batch: SynthBatch, first: int, count: int, padding: int = 0) -> bytes:
    """Wraps frames [first, first+count) of `batch` into a FLAC file image."""
    cfg = batch.config
    lens = batch.frame_lengths[first:first + count]
    start, end = int(batch.frame_offsets[first]), int(batch.frame_offsets[first + count])
    blocks = [int(batch.pcm_offsets[i + 1] - batch.pcm_offsets[i]) // cfg.n_channels
              for i in range(first, first + count)]
    md5 = hashlib.md5()
    for i in range(first, first + count):
        seg = batch.pcm[int(batch.pcm_offsets[i]):int(batch.pcm_offsets[i + 1])]
        md5.update(interleaved_le_bytes(seg, cfg.n_channels, cfg.bps))
    rates = [0, 88200, 176400, 192000, 8000, 16000, 22050, 24000, 32000, 44100, 48000, 96000]
    sr = rates[cfg.sample_rate_code] or 44100
    total = sum(blocks)
    si = bytearray()
    si += int(max(16, min(blocks))).to_bytes(2, "big") + int(max(16, max(blocks))).to_bytes(2, "big")
    si += int(lens.min()).to_bytes(3, "big") + int(lens.max()).to_bytes(3, "big")
    packed = (sr << 44) | ((cfg.n_channels - 1) << 41) | ((cfg.bps - 1) << 36) | total
    si += packed.to_bytes(8, "big")
    si += md5.digest()
    out = bytearray(b"fLaC")
    last = 0x80 if padding == 0 else 0
    out += bytes([last | 0]) + len(si).to_bytes(3, "big") + si
    if padding:
        out += bytes([0x80 | 1]) + padding.to_bytes(3, "big") + bytes(padding)
    out += batch.data[start:end].tobytes()
    return bytes(out)
