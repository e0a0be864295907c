"""CPU test of the throughput path's per-lane logic (claxon_b200/csrc/clx_lanes.h).

The lanes of the index kernel (one per frame: headers, parameters, skipping the Rice codes of all channels but
the last) and of the decode kernel (one per subframe: Rice decode from the recorded start bit) never talk to
each other, so the code the CUDA kernels run per lane is compiled for the host (tools/seq_host.cpp: plain loads
instead of the shared-memory ring, plus a scalar restatement of the prediction arithmetic) and checked here
against PCM known by construction and against the oracle's per-frame status.  A frame a lane declines (status
-2) is one the generic kernel decodes on the device; frames the oracle accepts must not be declined.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import claxon_b200 as cb
from claxon_b200 import synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "scratch", "seq_host.so")
SRC = os.path.join(ROOT, "tools", "seq_host.cpp")
HDR = os.path.join(ROOT, "claxon_b200", "csrc", "clx_lanes.h")


@pytest.fixture(scope="module")
def harness():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-I", os.path.join(ROOT, "include"), "-o", SO, SRC])
    L = C.CDLL(SO)
    L.seq_host_decode.restype = C.c_int
    L.seq_host_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                  C.c_void_p]
    return L


def run_lane_path(L, b, head_pad):
    descs, out_elems = cb.descs_from_offsets(b.data, b.frame_offsets[:-1], b.frame_lengths)
    data = np.concatenate([b.data, np.zeros(256, np.uint8)])  # slack, as the device buffers have
    out = np.full(out_elems, 0x5A5A5A5A, np.int32)
    res = np.zeros(b.n_frames, dtype=[("status", "<i4"), ("consumed", "<u4")])
    stats = np.zeros(2, np.uint64)
    descs = np.ascontiguousarray(descs)
    L.seq_host_decode(data.ctypes.data, b.data.size, descs.ctypes.data, b.n_frames, int(head_pad), out.ctypes.data,
                      res.ctypes.data, stats.ctypes.data)
    return descs, out, res, stats


def check(L, b, head_pad, max_declined=0.0):
    descs, out, res, stats = run_lane_path(L, b, head_pad)
    declined = 0
    for i in range(b.n_frames):
        if res["status"][i] != 0:
            declined += 1
            continue
        o, n = int(descs[i]["out_offset"]), int(descs[i]["n_channels"]) * int(descs[i]["block_size"])
        lo, hi = int(b.pcm_offsets[i]), int(b.pcm_offsets[i + 1])
        assert np.array_equal(out[o:o + n], b.pcm[lo:hi]), f"frame {i} differs from the generator's PCM"
        assert res["consumed"][i] == b.frame_lengths[i]
    assert declined <= max_declined * b.n_frames, (declined, b.n_frames)
    return stats


CASES = {
    "c2-ms": (synth.workload_config("c2", 40), 0, 0.0),
    "c2-ms-late-start": (synth.workload_config("c2", 40), 3, 0.0),
    "c2-indep": (synth.workload_config("c2-indep", 33), 0, 0.0),
    "c3": (synth.workload_config("c3", 40), 1, 0.0),
    "c4-files": (synth.workload_config("c4", 66), 0, 0.0),
    "c4-files-late-start": (synth.workload_config("c4", 66), 2, 0.0),
    "c5-order32-8ch": (synth.workload_config("c5", 3), 0, 0.0),
    "all-types-wasted-rice2": (synth.SynthConfig(n_frames=128, block_size=1152, n_channels=2, bps=16, stereo_mode=-1,
        type_mask=15, lpc_min_order=1, lpc_max_order=32, qlp_precision=0, rice_mode=-2, rice_kmin=0, rice_kmax=14,
        max_porder=6, rice2=2, wasted_max=5, long_unary_per_mille=100), 0, 0.0),
    "all-types-late-start": (synth.SynthConfig(n_frames=128, block_size=1152, n_channels=2, bps=16, stereo_mode=-1,
        type_mask=15, lpc_min_order=1, lpc_max_order=32, qlp_precision=0, rice_mode=-2, rice_kmin=0, rice_kmax=14,
        max_porder=6, rice2=2, wasted_max=5, long_unary_per_mille=100), 5, 0.0),
    "ragged-3ch-24bit": (synth.SynthConfig(n_frames=77, block_size=1000, tail_block_size=37, n_channels=3, bps=24,
        stereo_mode=0, type_mask=15, lpc_min_order=1, lpc_max_order=12, qlp_precision=0, rice_mode=-1, max_porder=3,
        wasted_max=3), 1, 0.0),
    "tiny-blocks-8bit": (synth.SynthConfig(n_frames=50, block_size=16, tail_block_size=5, n_channels=2, bps=8,
        stereo_mode=-1, type_mask=15, lpc_min_order=1, lpc_max_order=16, qlp_precision=0, rice_mode=-1, max_porder=2,
        force_bs16=1), 0, 0.0),
    "mono-20bit-k0": (synth.SynthConfig(n_frames=20, block_size=4608, n_channels=1, bps=20, type_mask=12,
        lpc_min_order=1, lpc_max_order=12, qlp_precision=14, rice_mode=0, residual_mean=0.4, max_porder=8), 0, 0.0),
    "8ch-12bit-fixed": (synth.SynthConfig(n_frames=33, block_size=576, n_channels=8, bps=12, type_mask=4,
        rice_mode=-1, max_porder=4), 2, 0.0),
    "rice2-big-k": (synth.SynthConfig(n_frames=20, block_size=2048, n_channels=2, bps=24, stereo_mode=9, type_mask=8,
        lpc_min_order=2, lpc_max_order=20, qlp_precision=15, rice_mode=-2, rice_kmin=15, rice_kmax=22, rice2=1,
        max_porder=3), 0, 0.0),
    "max-blocksize": (synth.SynthConfig(n_frames=2, block_size=65535, n_channels=2, bps=16, stereo_mode=10,
        type_mask=8, lpc_min_order=12, lpc_max_order=12, rice_mode=-1, rice_kmax=14), 0, 0.0),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_lane_logic_vs_known_pcm(harness, case):
    cfg, head_pad, max_declined = CASES[case]
    stats = check(harness, synth.generate(cfg), head_pad, max_declined)
    if case.startswith("c2"):
        assert 8 * stats[0] > 50 * stats[1]  # (groups of 8 vs single codes) the shape the fast group is written for stays on the fast path


def test_lane_logic_on_corrupted_frames(harness):
    """Bit flips and truncations: the lane either declines the frame or agrees with the oracle bit for bit
    (status OK, consumed, PCM); it must never crash or run away."""
    base = synth.generate(synth.SynthConfig(n_frames=40, block_size=576, n_channels=2, bps=16, stereo_mode=-1,
                                           type_mask=15, lpc_min_order=1, lpc_max_order=32, qlp_precision=0,
                                           rice_mode=-1, max_porder=4, rice2=2, wasted_max=4))
    rng = np.random.default_rng(7)
    frames = []
    for trial in range(400):
        i = int(rng.integers(0, base.n_frames))
        f = base.data[int(base.frame_offsets[i]):int(base.frame_offsets[i + 1])].copy()
        if trial % 3 == 2:
            f = f[: int(rng.integers(6, f.size))]
        else:
            for _ in range(int(rng.integers(1, 4))):
                f[int(rng.integers(5, min(f.size, 60) if trial % 3 == 0 else f.size))] ^= 1 << int(rng.integers(0, 8))
        st, d = cb.parse_frame_header(f)
        if st == 0:
            frames.append(f)
    data = np.concatenate(frames)
    lens = np.array([f.size for f in frames], np.uint32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    descs, out_elems = cb.descs_from_offsets(data, offs, lens)
    bad, st, ref = O.decode_batch(data, offs, lens, descs["out_offset"], out_elems, n_threads=4, verify_crc=False)
    for head_pad in (0, 2):
        padded = np.concatenate([data, np.zeros(256, np.uint8)])
        out = np.zeros(out_elems, np.int32)
        res = np.zeros(len(frames), dtype=[("status", "<i4"), ("consumed", "<u4")])
        descs = np.ascontiguousarray(descs)
        harness.seq_host_decode(padded.ctypes.data, data.size, descs.ctypes.data, len(frames), head_pad,
                                out.ctypes.data, res.ctypes.data, None)
        agreed = 0
        for i in range(len(frames)):
            if res["status"][i] != 0:
                continue
            assert st[i] == 0, f"frame {i}: lane accepted a frame the oracle rejects with {st[i]}"
            o, n = int(descs[i]["out_offset"]), int(descs[i]["n_channels"]) * int(descs[i]["block_size"])
            assert np.array_equal(out[o:o + n], ref[o:o + n])
            agreed += 1
        assert agreed > 0


def run_frames(L, data, offs, lens, head_pad):
    """Decodes frames (bytes `data`, offsets/lengths) through the lane path; returns descs, out, res."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    descs, out_elems = cb.descs_from_offsets(data, np.asarray(offs, np.uint64), np.asarray(lens, np.uint32))
    descs = np.ascontiguousarray(descs)
    padded = np.concatenate([data, np.zeros(256, np.uint8)])
    out = np.full(max(1, out_elems), 0x5A5A5A5A, np.int32)
    res = np.zeros(len(offs), dtype=[("status", "<i4"), ("consumed", "<u4")])
    L.seq_host_decode(padded.ctypes.data, data.size, descs.ctypes.data, len(offs), int(head_pad), out.ctypes.data,
                      res.ctypes.data, None)
    return descs, out, res


@pytest.mark.parametrize("name", ["pop", "short", "wasted_bits", "non_subset", "empty_vorbis_comment",
                                  "repeated_vorbis_comment"])
def test_lane_logic_on_reference_fixtures(harness, golden, name):
    """The reference's own test streams (PCM pinned by the STREAMINFO MD5 / the order-20 KAT, see test_oracle_golden):
    fixed-4 + Rice (pop), LPC-1 (short), wasted bits + 16-bit block-size code, high-order LPC + Rice2 + mid/side."""
    data = golden[f"{name}__bytes"]
    rows = golden[f"{name}__frames"]
    exp = golden[f"{name}__pcm"]
    frames = [r for r in rows if r[1] == 0 and r[3] > 0]
    offs = [int(r[0]) for r in frames]
    lens = [int(r[3]) for r in frames]
    for head_pad in (0, 1, 4):
        descs, out, res = run_frames(harness, data, offs, lens, head_pad)
        assert (res["status"] == 0).all(), (name, head_pad, res)
        assert np.array_equal(res["consumed"], np.asarray(lens, np.uint32))
        pos = 0
        for i, r in enumerate(frames):
            n = int(r[4] * r[5])
            o = int(descs[i]["out_offset"])
            assert np.array_equal(out[o:o + n], exp[pos:pos + n]), (name, head_pad, i)
            pos += n


@pytest.mark.parametrize("seed", range(8))
def test_lane_logic_random_configs_vs_oracle(harness, seed):
    """Random shapes (every subframe type, orders 1..32, Rice and Rice2, wasted bits, long unary runs, 1..8 channels):
    whatever the lane accepts must equal the oracle bit for bit; what the oracle accepts should rarely be declined."""
    rng = np.random.default_rng(5000 + seed)
    nch = int(rng.integers(1, 9))
    bps = int(rng.choice([8, 12, 16, 20, 24]))
    cfg = synth.SynthConfig(
        seed=int(rng.integers(1, 2**31)), n_frames=int(rng.integers(1, 70)),
        block_size=int(rng.choice([16, 192, 576, 1000, 1152, 2304, 4096, int(rng.integers(1, 5000))])),
        n_channels=nch, bps=bps, stereo_mode=-1 if nch == 2 else 0,
        type_mask=int(rng.integers(1, 16)), lpc_min_order=1, lpc_max_order=int(rng.integers(1, 33)),
        qlp_precision=0, rice_mode=int(rng.choice([-1, -2])), rice_kmin=0, rice_kmax=14,
        max_porder=int(rng.integers(0, 8)), rice2=int(rng.integers(0, 3)), wasted_max=int(rng.integers(0, 6)),
        long_unary_per_mille=int(rng.choice([0, 50])))
    b = synth.generate(cfg)
    offs, lens = b.frame_offsets[:-1], b.frame_lengths
    descs0, out_elems = cb.descs_from_offsets(b.data, offs, lens)
    bad, st, ref = O.decode_batch(b.data, offs, lens, descs0["out_offset"], out_elems, n_threads=4)
    assert bad == 0
    for head_pad in (0, 3):
        descs, out, res = run_frames(harness, b.data, offs, lens, head_pad)
        accepted = 0
        for i in range(b.n_frames):
            if res["status"][i] != 0:
                continue
            accepted += 1
            o, n = int(descs[i]["out_offset"]), int(descs[i]["n_channels"]) * int(descs[i]["block_size"])
            assert np.array_equal(out[o:o + n], ref[o:o + n]), (seed, head_pad, i)
            assert res["consumed"][i] == lens[i]
        assert accepted == b.n_frames, (seed, accepted, b.n_frames)  # no valid stream is declined


def test_lane_logic_wrapping_streams(harness):
    """Streams whose samples wrap around i32 (reference: all arithmetic is wrapping, Appendix A.8): huge Rice2
    residuals exercise the slow code path (codes longer than the 32-bit window) and the u32 wrap of (q << k) | r."""
    cfg = synth.SynthConfig(n_frames=12, block_size=1024, n_channels=2, bps=16, stereo_mode=0, type_mask=8,
                            lpc_min_order=1, lpc_max_order=8, qlp_precision=5, rice_mode=-2, rice_kmin=26,
                            rice_kmax=29, rice2=1, residual_mean=2.0e8, max_porder=1)
    b = synth.generate(cfg)
    assert np.abs(b.pcm.astype(np.int64)).max() > 2**29
    descs, out, res = run_frames(harness, b.data, b.frame_offsets[:-1], b.frame_lengths, 0)
    assert (res["status"] == 0).all()
    for i in range(b.n_frames):
        o = int(descs[i]["out_offset"]); lo, hi = int(b.pcm_offsets[i]), int(b.pcm_offsets[i + 1])
        assert np.array_equal(out[o:o + hi - lo], b.pcm[lo:hi])
