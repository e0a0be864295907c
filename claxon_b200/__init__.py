"""claxon_b200 — B200-native batched FLAC frame decoder behind claxon's API surface.

Host-side mirror of the reference interface for the per-frame decode path:

    claxon::FlacReader              -> claxon_b200.FlacReader       (reference src/lib.rs:217-470)
    claxon::frame::FrameReader      -> claxon_b200.FrameReader      (src/frame.rs:650-785)
    claxon::frame::Block            -> claxon_b200.Block            (src/frame.rs:402-529)
    claxon::Error                   -> claxon_b200.Error            (src/error.rs:18-32)

plus the batched entry points the GPU wants (`Context.decode_frames`, `demux_frames`).
All sample arithmetic happens in the CUDA library (`libclaxon_b200.so`); this package is
plumbing.  Nothing here imports the test oracle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (OPEN_METADATA_ONLY, OPEN_NO_VORBIS_COMMENT)
from ._lib import (FrameDesc, FrameResult, OPT_NO_VERIFY_CRC, OPT_GENERIC_KERNEL_ONLY, OPT_WARP_PER_FRAME, OPT_LANE_PER_FRAME,
                   FRAME_VARIABLE_BLOCKING, FRAME_CRC16_VERIFIED,
                   OUT_PLANAR_I32, OUT_INTERLEAVED_I32, OUT_INTERLEAVED_I16, OUT_INTERLEAVED_I24)

__all__ = ["Error", "Block", "FrameReader", "FlacReader", "FlacReaderOptions", "StreamInfo", "Context", "DeviceBatch",
           "parse_frame_header", "demux_frames", "open_stream", "ogg_frames", "mp4_frames", "status_str", "DESC_DTYPE", "RESULT_DTYPE"]

# numpy views of the C structs (same layout; asserted below)
DESC_DTYPE = np.dtype([
    ("byte_offset", "<u8"), ("byte_len", "<u4"), ("header_len", "<u2"), ("block_size", "<u2"),
    ("n_channels", "u1"), ("channel_assignment", "u1"), ("bits_per_sample", "u1"), ("flags", "u1"),
    ("sample_rate", "<u4"), ("number", "<u8"), ("out_offset", "<u8")], align=True)
RESULT_DTYPE = np.dtype([("status", "<i4"), ("consumed", "<u4")], align=True)
assert DESC_DTYPE.itemsize == C.sizeof(FrameDesc) and RESULT_DTYPE.itemsize == C.sizeof(FrameResult)

KIND_NONE, KIND_IO, KIND_FORMAT, KIND_UNSUPPORTED, KIND_LIBRARY = range(5)
OK, EOF = 0, 1


def status_str(status: int) -> str:
    return _lib.load().clx_status_str(int(status)).decode()


class Error(Exception):
    """claxon::Error — compares by variant + message like the reference (src/error.rs:34-45)."""

    def __init__(self, status: int, detail: str = ""):
        self.status = int(status)
        self.kind = _lib.load().clx_status_kind(self.status)
        self.message = status_str(self.status)
        super().__init__(self.message + (f" ({detail})" if detail else ""))

    @property
    def variant(self) -> str:
        return {KIND_IO: "IoError", KIND_FORMAT: "FormatError", KIND_UNSUPPORTED: "Unsupported"}.get(
            self.kind, "LibraryError")

    def __eq__(self, other):
        if not isinstance(other, Error):
            return NotImplemented
        if self.kind == KIND_IO or other.kind == KIND_IO:
            return False  # (&IoError(_), _) => false
        return self.kind == other.kind and self.message == other.message

    __hash__ = Exception.__hash__


def _check(status: int, ctx: "Context | None" = None):
    if status != OK:
        detail = ""
        if ctx is not None and status == 91:
            detail = _lib.load().clx_ctx_last_error(ctx._h).decode()
        raise Error(status, detail)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        if data.dtype != np.uint8 or not data.flags.c_contiguous:
            data = np.ascontiguousarray(data, dtype=np.uint8)
        return data
    return np.frombuffer(bytes(data) if not isinstance(data, (bytes, bytearray, memoryview)) else data,
                         dtype=np.uint8)


# ---------------------------------------------------------------------------
# host-side parsing
# ---------------------------------------------------------------------------

@dataclass
class StreamInfo:  # claxon::metadata::StreamInfo (src/metadata.rs:29-54)
    min_block_size: int
    max_block_size: int
    min_frame_size: int | None
    max_frame_size: int | None
    sample_rate: int
    channels: int
    bits_per_sample: int
    samples: int | None
    md5sum: bytes

    @staticmethod
    def _from_c(si) -> "StreamInfo":
        return StreamInfo(si.min_block_size, si.max_block_size, si.min_frame_size or None,
                          si.max_frame_size or None, si.sample_rate, si.channels, si.bits_per_sample,
                          si.samples or None, bytes(si.md5sum))


def parse_frame_header(data, offset: int = 0, flags: int = 0):
    """read_frame_header_or_eof (src/frame.rs:131-316). Returns (status, FrameDesc)."""
    buf = _as_u8(data)
    d = FrameDesc()
    st = _lib.load().clx_parse_frame_header(buf.ctypes.data + offset, buf.size - offset, C.byref(d), flags)
    return st, d


def open_stream(data):
    """FlacReader::new's metadata walk. Returns (StreamInfo, first_frame_offset); raises Error."""
    buf = _as_u8(data)
    si = _lib.StreamInfoC()
    first = C.c_uint64(0)
    _check(_lib.load().clx_open_stream(buf.ctypes.data, buf.size, C.byref(si), C.byref(first)))
    return StreamInfo._from_c(si), first.value


def demux_frames(data, start: int = 0, max_frames: int = 1 << 20, flags: int = 0, threads: int = 1):
    """Finds frame boundaries without decoding. Returns (descs ndarray, next_offset, out_elems, stop_status).
    `threads` != 1: clx_demux_frames_mt on that many host threads (0 = all), same results."""
    buf = _as_u8(data)
    cap = min(max_frames, max(16, (buf.size - start) // 16 + 1))
    while True:
        descs = np.empty(cap, dtype=DESC_DTYPE)  # (filled by the call; zeroing 40 bytes per possible frame costs more than the scan)
        nxt, total, stop = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
        if threads == 1:
            n = _lib.load().clx_demux_frames(buf.ctypes.data, buf.size, start, descs.ctypes.data, cap,
                                             C.byref(nxt), C.byref(total), C.byref(stop), flags)
        else:
            n = _lib.load().clx_demux_frames_mt(buf.ctypes.data, buf.size, start, descs.ctypes.data, cap,
                                                C.byref(nxt), C.byref(total), C.byref(stop), flags, threads)
        if n < cap or cap >= max_frames:
            return descs[:n].copy(), nxt.value, total.value, stop.value
        cap = min(max_frames, cap * 4)


def ogg_frames(data, flags: int = 0):
    """Frames of an in-memory Ogg FLAC file (examples/decode_ogg.rs): (StreamInfo, frame bytes, descs, out_elems);
    the descriptors index the returned byte array (packets may span pages in the file)."""
    buf = _as_u8(data)
    si = _lib.StreamInfoC()
    frames = np.zeros(max(16, buf.size), dtype=np.uint8)
    descs = np.zeros(max(16, buf.size // 8), dtype=DESC_DTYPE)
    n, used, total = C.c_size_t(0), C.c_size_t(0), C.c_uint64(0)
    _check(_lib.load().clx_ogg_frames(buf.ctypes.data, buf.size, C.byref(si), frames.ctypes.data, frames.size,
                                      descs.ctypes.data, descs.size, C.byref(n), C.byref(used), C.byref(total), flags))
    return StreamInfo._from_c(si), frames[: used.value].copy(), descs[: n.value].copy(), int(total.value)


def mp4_frames(data, flags: int = 0):
    """Frames of an in-memory MP4 file with a 'fLaC' track (examples/decode_mp4.rs): (StreamInfo, descs, out_elems);
    the descriptors index `data` itself."""
    buf = _as_u8(data)
    si = _lib.StreamInfoC()
    descs = np.zeros(max(16, buf.size // 8), dtype=DESC_DTYPE)
    n, total = C.c_size_t(0), C.c_uint64(0)
    _check(_lib.load().clx_mp4_frames(buf.ctypes.data, buf.size, C.byref(si), descs.ctypes.data, descs.size, C.byref(n),
                                      C.byref(total), flags))
    return StreamInfo._from_c(si), descs[: n.value].copy(), int(total.value)


def descs_from_offsets(data, offsets, lengths=None, flags: int = 0) -> tuple[np.ndarray, int]:
    """Builds descriptors for frames at known byte offsets (container-provided boundaries,
    cf. reference examples/decode_ogg.rs:107-113). Returns (descs, out_elems)."""
    buf = _as_u8(data)
    offsets = np.asarray(offsets, dtype=np.uint64)
    descs = np.zeros(offsets.size, dtype=DESC_DTYPE)
    L = _lib.load()
    out_at = 0
    d = FrameDesc()
    for i, off in enumerate(offsets):
        off = int(off)
        ln = int(lengths[i]) if lengths is not None else buf.size - off
        st = L.clx_parse_frame_header(buf.ctypes.data + off, ln, C.byref(d), flags)
        if st != OK:
            raise Error(st, f"frame {i}")
        d.byte_offset, d.byte_len, d.out_offset = off, ln, out_at
        descs[i] = np.frombuffer(bytes(d), dtype=DESC_DTYPE)[0]
        out_at += (d.n_channels * d.block_size + 3) & ~3
    return descs, out_at


# ---------------------------------------------------------------------------
# device context
# ---------------------------------------------------------------------------

class Context:
    """clx_ctx: one per host thread / GPU. Raises Error(NO_DEVICE) without a usable GPU."""

    def __init__(self, device: int = 0, verify_crc: bool = True, n_streams: int = 2, generic_only: bool = False,
                 warp_per_frame: bool = False, lane_per_frame: bool = False, host_threads: int = 0):
        """Default: device-resident batches and large calls use the lane-per-frame index pass + lane-per-subframe
        decode pass (csrc/clx_fused.cu); small synchronous host-buffer calls (latency regime) use the
        warp-per-frame path (csrc/clx_coop.cu).  `warp_per_frame` / `lane_per_frame` force one of them everywhere,
        `generic_only` bypasses both (testing, A/B measurements)."""
        self._L = _lib.load()
        flags = ((0 if verify_crc else OPT_NO_VERIFY_CRC) | (OPT_GENERIC_KERNEL_ONLY if generic_only else 0)
                 | (OPT_WARP_PER_FRAME if warp_per_frame else 0) | (OPT_LANE_PER_FRAME if lane_per_frame else 0))
        opts = _lib.Options(device, flags, n_streams, host_threads)
        h = C.c_void_p()
        _check(self._L.clx_ctx_create(C.byref(opts), C.byref(h)))
        self._h = h
        self.verify_crc = verify_crc

    def close(self):
        if getattr(self, "_h", None):
            self._L.clx_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._L.clx_ctx_launch_count(self._h))

    def decode_frames(self, data, descs: np.ndarray, out: np.ndarray | None = None,
                      out_elems: int | None = None, mode: int = OUT_PLANAR_I32):
        """End-to-end host-buffer decode. Returns (out ndarray, results ndarray).  `mode`: OUT_PLANAR_I32 (claxon's
        Block layout, int32) or an interleaved little-endian form: OUT_INTERLEAVED_I32 (int32), _I16 (int16),
        _I24 (uint8, 3 bytes per sample); out_offset / out_elems count samples in every mode."""
        buf = _as_u8(data)
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        if out_elems is None:
            ends = descs["out_offset"] + descs["n_channels"].astype(np.uint64) * descs["block_size"]
            out_elems = int(ends.max()) if descs.size else 0
        if out is None:
            n = max(1, out_elems)
            out = (np.empty(n, dtype=np.int16) if mode == OUT_INTERLEAVED_I16 else
                   np.empty(3 * n, dtype=np.uint8) if mode == OUT_INTERLEAVED_I24 else np.empty(n, dtype=np.int32))
        results = np.zeros(descs.size, dtype=RESULT_DTYPE)
        _check(self._L.clx_decode_frames_to(self._h, buf.ctypes.data, buf.size, descs.ctypes.data, descs.size,
                                            out.ctypes.data, max(1, out_elems), results.ctypes.data, mode), self)
        return out, results

    def decode_frames_raw(self, bytes_ptr: int, nbytes: int, descs_ptr: int, n: int, out_ptr: int,
                          out_elems: int, results_ptr: int, mode: int = OUT_PLANAR_I32):
        """Same call on raw host addresses (pinned buffers owned by the caller)."""
        _check(self._L.clx_decode_frames_to(self._h, bytes_ptr, nbytes, descs_ptr, n, out_ptr, out_elems,
                                            results_ptr, mode), self)

    def run_steps(self, batches: list["DeviceBatch"], steps: int, n_streams: int) -> float:
        """Decodes `steps` batches round-robin over `n_streams` streams; returns device ms (CUDA events)."""
        arr = (C.c_void_p * len(batches))(*[b._h for b in batches])
        ms = C.c_float(0)
        _check(self._L.clx_ctx_run_steps(self._h, arr, len(batches), steps, n_streams, C.byref(ms)), self)
        return float(ms.value)

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer as a uint8 ndarray (freed with host_free)."""
        p = self._L.clx_host_alloc(nbytes)
        if not p:
            raise MemoryError("cudaHostAlloc failed")
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p))

    def host_free(self, arr: np.ndarray):
        self._L.clx_host_free(arr.ctypes.data)

    def upload(self, data, descs: np.ndarray, out_elems: int) -> "DeviceBatch":
        return DeviceBatch(self, data, descs, out_elems)

    def adopt(self, device_ptr: int, nbytes: int, descs: np.ndarray, out_elems: int) -> "DeviceBatch":
        """A batch whose frame bytes already sit in this GPU's memory at `device_ptr` (e.g. a torch tensor's
        data_ptr() after the NCCL scatter of claxon_b200.shard.scatter_batch): copied device to device."""
        return DeviceBatch(self, None, descs, out_elems, device_ptr=device_ptr, nbytes=nbytes)


class DeviceBatch:
    """clx_batch: frames resident in HBM; decode() launches the kernels only."""

    def __init__(self, ctx: Context, data, descs: np.ndarray, out_elems: int, device_ptr: int | None = None,
                 nbytes: int = 0):
        self.ctx = ctx
        self.descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        self.out_elems = int(out_elems)
        h = C.c_void_p()
        if device_ptr is None:
            buf = _as_u8(data)
            self.nbytes = int(buf.size)
            _check(ctx._L.clx_batch_create(ctx._h, buf.ctypes.data, buf.size, self.descs.ctypes.data,
                                           self.descs.size, self.out_elems, C.byref(h)), ctx)
        else:
            self.nbytes = int(nbytes)
            _check(ctx._L.clx_batch_create_ex(ctx._h, device_ptr, self.nbytes, self.descs.ctypes.data, self.descs.size,
                                              self.out_elems, _lib.BATCH_BYTES_ON_DEVICE, C.byref(h)), ctx)
        self._h = h

    def decode(self, stream: int = 0):
        _check(self.ctx._L.clx_batch_decode(self.ctx._h, self._h, stream), self.ctx)

    def sync(self):
        _check(self.ctx._L.clx_batch_sync(self.ctx._h, self._h), self.ctx)

    def kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(self.ctx._L.clx_batch_last_kernel_ms(self.ctx._h, self._h, C.byref(ms)), self.ctx)
        return float(ms.value)

    def read(self):
        out = np.empty(max(1, self.out_elems), dtype=np.int32)
        results = np.zeros(self.descs.size, dtype=RESULT_DTYPE)
        _check(self.ctx._L.clx_batch_read(self.ctx._h, self._h, out.ctypes.data, out.size,
                                          results.ctypes.data), self.ctx)
        return out, results

    @property
    def device_out_ptr(self) -> int:
        return int(self.ctx._L.clx_batch_device_out(self._h) or 0)

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._L.clx_batch_destroy(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ---------------------------------------------------------------------------
# claxon-shaped API
# ---------------------------------------------------------------------------

class Block:
    """claxon::frame::Block (src/frame.rs:402-529): planar samples, channel-major."""

    def __init__(self, time: int, block_size: int, buffer: np.ndarray):
        self._time = int(time)
        self._bs = int(block_size)
        self._buffer = buffer
        self._channels = (buffer.size // block_size) if block_size else 0  # src/frame.rs:418

    @staticmethod
    def empty() -> "Block":
        return Block(0, 0, np.zeros(0, dtype=np.int32))

    def time(self) -> int:
        return self._time

    def len(self) -> int:
        return self._bs * self._channels

    __len__ = len

    def duration(self) -> int:
        return self._bs

    def channels(self) -> int:
        return self._channels

    def channel(self, ch: int) -> np.ndarray:
        if not 0 <= ch < self._channels:
            raise IndexError("channel out of range")  # the reference panics
        return self._buffer[ch * self._bs:(ch + 1) * self._bs]

    def sample(self, ch: int, sample: int) -> int:
        return int(self._buffer[ch * self._bs + sample])

    def into_buffer(self) -> np.ndarray:
        return self._buffer

    def stereo_samples(self):
        if self._channels != 2:
            raise RuntimeError("stereo_samples() must only be called for blocks with two channels.")
        left, right = self.channel(0), self.channel(1)
        return ((int(l), int(r)) for l, r in zip(left, right))


def _ensure_buffer_len(buffer: np.ndarray | None, new_len: int) -> np.ndarray:
    """ensure_buffer_len (src/frame.rs:616-637): exact length, capacity reused when sufficient."""
    if buffer is None:
        return np.zeros(new_len, dtype=np.int32)
    base = buffer.base if isinstance(buffer.base, np.ndarray) and buffer.base.dtype == np.int32 else buffer
    if base.size >= new_len:
        return base[:new_len]
    return np.zeros(new_len, dtype=np.int32)


class FrameReader:
    """claxon::frame::FrameReader over an in-memory byte span positioned at a frame header."""

    def __init__(self, input, ctx: Context | None = None, _flac: bool = False):
        self._ctx = ctx or default_context()
        self._buf = _as_u8(input)
        L = self._ctx._L
        h = C.c_void_p()
        opener = L.clx_reader_open_flac if _flac else L.clx_reader_open_frames
        _check(opener(self._ctx._h, self._buf.ctypes.data, self._buf.size, C.byref(h)), self._ctx)
        self._h = h

    def read_next_or_eof(self, buffer: np.ndarray | None = None) -> Block | None:
        """Decodes the next frame; None at end of stream; raises Error on malformed input."""
        L = self._ctx._L
        st, d = parse_frame_header(self._buf, self.position(), 0 if self._ctx.verify_crc else OPT_NO_VERIFY_CRC)
        if st == EOF:
            return None
        _check(st)
        buffer = _ensure_buffer_len(buffer, d.n_channels * d.block_size)
        bs, ch, t = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        st = L.clx_reader_next(self._h, buffer.ctypes.data, buffer.size, C.byref(bs), C.byref(ch), C.byref(t))
        if st == EOF:
            return None
        _check(st, self._ctx)
        return Block(t.value, bs.value, buffer)

    def read_batch(self, max_frames: int, buffer: np.ndarray | None = None) -> list[Block]:
        """Batched extension: demux + decode up to max_frames frames in one device pass."""
        L = self._ctx._L
        nf, need = C.c_size_t(0), C.c_uint64(0)
        st = L.clx_reader_plan_batch(self._h, max_frames, C.byref(nf), C.byref(need))  # demux ahead: exact buffer size
        if st == EOF:
            return []
        _check(st, self._ctx)
        descs = np.zeros(max_frames, dtype=DESC_DTYPE)
        cap = max(1, int(need.value))
        if buffer is None or buffer.size < cap:
            buffer = np.empty(cap, dtype=np.int32)
        n = C.c_size_t(0)
        st = L.clx_reader_next_batch(self._h, max_frames, buffer.ctypes.data, buffer.size, descs.ctypes.data,
                                     C.byref(n))
        if st == EOF:
            return []
        _check(st, self._ctx)
        blocks = []
        for i in range(n.value):
            d = descs[i]
            o, cnt = int(d["out_offset"]), int(d["n_channels"]) * int(d["block_size"])
            blocks.append(Block(int(d["number"]), int(d["block_size"]), buffer[o:o + cnt]))
        return blocks

    def position(self) -> int:
        return int(self._ctx._L.clx_reader_position(self._h))

    def into_inner(self):
        return self._buf

    def close(self):
        if getattr(self, "_h", None):
            self._ctx._L.clx_reader_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class FlacReaderOptions:
    """claxon::FlacReaderOptions (src/lib.rs:123-151)."""
    metadata_only: bool = False
    read_vorbis_comment: bool = True


def _parse_vorbis_comment(body: np.ndarray):
    """(vendor, [(comment, separator index)]) of a validated VORBIS_COMMENT body (src/metadata.rs:402-513)."""
    raw = body.tobytes()
    at = 4 + int.from_bytes(raw[0:4], "little")
    vendor = raw[4:at].decode("utf-8")
    count = int.from_bytes(raw[at:at + 4], "little")
    at += 4
    comments = []
    while len(raw) - at >= 4 and len(comments) < count:
        n = int.from_bytes(raw[at:at + 4], "little")
        at += 4
        if n == 0:  # zero-length comments occur in the wild and are skipped
            count -= 1
            continue
        c = raw[at:at + n]
        at += n
        comments.append((c.decode("utf-8"), c.index(b"=")))
    return vendor, comments


class FlacReader:
    """claxon::FlacReader (src/lib.rs:207-470) for in-memory streams / files."""

    def __init__(self, data, ctx: Context | None = None, options: FlacReaderOptions | None = None):
        self._options = options or FlacReaderOptions()
        buf = _as_u8(data)
        flags = ((OPEN_METADATA_ONLY if self._options.metadata_only else 0)
                 | (0 if self._options.read_vorbis_comment else OPEN_NO_VORBIS_COMMENT))
        si = _lib.StreamInfoC()
        first, vc_off, vc_len = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        _check(_lib.load().clx_open_stream_ex(buf.ctypes.data, buf.size, flags, C.byref(si), C.byref(first),
                                              C.byref(vc_off), C.byref(vc_len)))
        self._si = StreamInfo._from_c(si)
        self._vendor, self._comments = None, []
        if vc_len.value:
            self._vendor, self._comments = _parse_vorbis_comment(buf[vc_off.value:vc_off.value + vc_len.value])
        # FlacReaderState::Full / MetadataOnly (src/lib.rs:96-104)
        self._frames = None if self._options.metadata_only else FrameReader(buf, ctx, _flac=True)

    @classmethod
    def new(cls, data, ctx: Context | None = None) -> "FlacReader":
        return cls(data, ctx)

    @classmethod
    def new_ext(cls, data, options: FlacReaderOptions, ctx: Context | None = None) -> "FlacReader":
        return cls(data, ctx, options)

    @classmethod
    def open(cls, path, ctx: Context | None = None) -> "FlacReader":
        with open(path, "rb") as f:
            return cls(f.read(), ctx)

    @classmethod
    def open_ext(cls, path, options: FlacReaderOptions, ctx: Context | None = None) -> "FlacReader":
        with open(path, "rb") as f:
            return cls(f.read(), ctx, options)

    def streaminfo(self) -> StreamInfo:
        return self._si

    def vendor(self) -> str | None:
        """Vendor string of the Vorbis comment block, if present (src/lib.rs:318-325)."""
        return self._vendor

    def tags(self):
        """(name, value) pairs of the Vorbis comments, names as stored (src/lib.rs:327-345)."""
        return [(c[:i], c[i + 1:]) for c, i in self._comments]

    def get_tag(self, tag_name: str):
        """Values of every comment whose name equals tag_name ASCII-case-insensitively (src/lib.rs:347-360)."""
        def lower(t):
            return "".join(chr(ord(ch) + 32) if "A" <= ch <= "Z" else ch for ch in t)
        return [c[i + 1:] for c, i in self._comments if lower(c[:i]) == lower(tag_name)]

    def _full(self, what: str) -> FrameReader:
        if self._frames is None:  # the reference panics
            raise RuntimeError(f"FlacReaderOptions::metadata_only must be false to be able to use FlacReader::{what}()")
        return self._frames

    def blocks(self) -> FrameReader:
        return self._full("blocks")

    def samples(self, batch_frames: int = 256):
        """FlacSamples (src/lib.rs:473-519): interleaved samples, channel by channel for each inter-channel
        sample; an error surfaces once, where the bad frame starts, after every sample before it.  Frames are
        decoded `batch_frames` at a time on the device."""
        return self._iter_samples(self._full("samples"), batch_frames)

    def into_samples(self, batch_frames: int = 256):
        """FlacIntoSamples (src/lib.rs:412-435): as samples(), taking the reader with it."""
        frames, self._frames = self._full("into_samples"), None
        return self._iter_samples(frames, batch_frames)

    @staticmethod
    def _iter_samples(frames: FrameReader, batch_frames: int):
        while True:
            blocks = frames.read_batch(batch_frames)  # raises if the very next frame is bad
            if not blocks:
                return
            for block in blocks:
                ch, bs = block.channels(), block.duration()
                yield from block.into_buffer().reshape(ch, bs).T.reshape(-1).tolist()

    def into_inner(self):
        return self._frames.into_inner() if self._frames is not None else None
