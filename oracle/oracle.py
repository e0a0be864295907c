"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

The oracle is a C restatement of claxon v0.4.3's frame-decode semantics
(oracle/claxon_oracle.c).  Only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module;
nothing under ``claxon_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compiles liboracle.so with gcc (a few hundred ms)."""
    src = os.path.join(_HERE, "claxon_oracle.c")
    deps = [src, os.path.join(_HERE, "claxon_oracle.h"),
            os.path.join(_HERE, "..", "include", "clx_status.h")]
    stale = force or not os.path.exists(_SO)
    if not stale and all(os.path.exists(d) for d in deps):
        stale = any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps)
    if stale and os.path.exists(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"] if force else ["make", "-C", _HERE, "-s"])
    return _SO


class FrameHeader(C.Structure):
    _fields_ = [("block_size", C.c_uint32), ("sample_rate", C.c_uint32),
                ("n_channels", C.c_uint32), ("channel_assignment", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("variable_blocking", C.c_uint32),
                ("number", C.c_uint64), ("header_len", C.c_uint32)]


class FrameInfo(C.Structure):
    _fields_ = [("header", FrameHeader), ("time", C.c_uint64), ("consumed", C.c_uint64),
                ("n_samples", C.c_uint32)]


class StreamInfo(C.Structure):
    _fields_ = [("min_block_size", C.c_uint32), ("max_block_size", C.c_uint32),
                ("min_frame_size", C.c_uint32), ("max_frame_size", C.c_uint32),
                ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("samples", C.c_uint64),
                ("md5sum", C.c_uint8 * 16)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        i32p = C.POINTER(C.c_int32)
        L.clxo_extend_sign_u16.restype = C.c_int16
        L.clxo_extend_sign_u16.argtypes = [C.c_uint16, C.c_uint32]
        L.clxo_extend_sign_u32.restype = C.c_int32
        L.clxo_extend_sign_u32.argtypes = [C.c_uint32, C.c_uint32]
        L.clxo_rice_to_signed.restype = C.c_int32
        L.clxo_rice_to_signed.argtypes = [C.c_uint32]
        L.clxo_predict_fixed.argtypes = [C.c_uint32, i32p, C.c_size_t]
        L.clxo_predict_lpc.argtypes = [C.POINTER(C.c_int16), C.c_uint32, C.c_uint32, i32p, C.c_size_t]
        for f in ("clxo_decode_left_side", "clxo_decode_right_side", "clxo_decode_mid_side"):
            getattr(L, f).argtypes = [i32p, C.c_size_t]
        L.clxo_crc8.restype = C.c_uint8
        L.clxo_crc8.argtypes = [C.c_char_p, C.c_size_t]
        L.clxo_crc16.restype = C.c_uint16
        L.clxo_crc16.argtypes = [C.c_char_p, C.c_size_t]
        L.clxo_read_var_length_int.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64),
                                               C.POINTER(C.c_size_t)]
        L.clxo_bit_read.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_int,
                                    C.c_uint32, C.POINTER(C.c_uint32)]
        L.clxo_read_frame_header.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(FrameHeader)]
        L.clxo_decode_subframe.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.c_uint32,
                                           i32p, C.c_uint32]
        L.clxo_decode_frame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.POINTER(FrameInfo), C.c_int]
        L.clxo_open_stream.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(StreamInfo),
                                       C.POINTER(C.c_uint64)]
        L.clxo_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_size_t,
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
        L.clxo_decode_batch_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


@dataclass
class OracleFrame:
    status: int
    info: FrameInfo
    samples: np.ndarray  # planar int32 [n_channels * block_size]; empty on error


def read_frame_header(data, offset: int = 0):
    buf = _u8(data)
    h = FrameHeader()
    st = lib().clxo_read_frame_header(buf.ctypes.data + offset, buf.size - offset, C.byref(h))
    return st, h


def decode_frame(data, offset: int = 0, length: int | None = None, verify_crc: bool = True,
                 fill: int | None = None) -> OracleFrame:
    """FrameReader::read_next_or_eof on the bytes data[offset:offset+length]."""
    buf = _u8(data)
    n = buf.size - offset if length is None else length
    st, h = read_frame_header(buf, offset)
    cap = max(1, h.n_channels * h.block_size) if h.block_size else 1
    out = np.empty(cap, dtype=np.int32)
    if fill is not None:
        out[:] = fill
    info = FrameInfo()
    st = lib().clxo_decode_frame(buf.ctypes.data + offset, n, out.ctypes.data, cap, C.byref(info),
                                 1 if verify_crc else 0)
    if st != 0:
        return OracleFrame(st, info, out[:0] if fill is None else out)
    return OracleFrame(st, info, out[: info.n_samples])


def open_stream(data):
    buf = _u8(data)
    si = StreamInfo()
    first = C.c_uint64(0)
    st = lib().clxo_open_stream(buf.ctypes.data, buf.size, C.byref(si), C.byref(first))
    return st, si, first.value


def decode_stream(data, first_frame: int, max_samples: int, verify_crc: bool = True):
    """Sequential `blocks()` loop; returns (status, n_frames, planar-per-frame samples)."""
    buf = _u8(data)
    out = np.empty(max(1, max_samples), dtype=np.int32)
    nf, ns = C.c_uint64(0), C.c_uint64(0)
    st = lib().clxo_decode_stream(buf.ctypes.data, buf.size, first_frame, out.ctypes.data, out.size,
                                  C.byref(nf), C.byref(ns), 1 if verify_crc else 0)
    return st, nf.value, out[: ns.value]


def decode_batch(data, offsets, lengths, out_offsets, n_out: int, n_threads: int = 1,
                 verify_crc: bool = True, out: np.ndarray | None = None):
    """Batch decode with frames sharded over n_threads independent decoders."""
    buf = _u8(data)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    out_offsets = np.ascontiguousarray(out_offsets, dtype=np.uint64)
    if out is None:
        out = np.empty(max(1, n_out), dtype=np.int32)
    statuses = np.zeros(offsets.size, dtype=np.int32)
    bad = lib().clxo_decode_batch_mt(buf.ctypes.data, offsets.ctypes.data, lengths.ctypes.data,
                                     offsets.size, out.ctypes.data, out_offsets.ctypes.data,
                                     statuses.ctypes.data, n_threads, 1 if verify_crc else 0)
    return bad, statuses, out
