"""Pins the CPU oracle (oracle/claxon_oracle.c) to the reference's own known-answer vectors
(tests/golden/kat.json, transcribed from claxon's in-source unit tests) and to the STREAMINFO MD5
of its fixtures (tests/golden/fixtures.npz, made by tests/golden/make_golden.py)."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
L = O.lib()


def _i32(a):
    return np.array(a, dtype=np.int32)


def test_extend_sign():
    for v, b, exp in KAT["extend_sign_u16"]["cases"]:
        assert L.clxo_extend_sign_u16(v, b) == exp
    for v, b, exp in KAT["extend_sign_u32"]["cases"]:
        assert L.clxo_extend_sign_u32(v, b) == exp


def test_rice_to_signed():
    for v, exp in KAT["rice_to_signed"]["cases"]:
        assert L.clxo_rice_to_signed(v) == exp


def test_predict_fixed():
    for c in KAT["predict_fixed"]["cases"]:
        buf = _i32(c["in"])
        L.clxo_predict_fixed(c["order"], buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size)
        assert buf.tolist() == c["out"]


def test_predict_lpc():
    for c in KAT["predict_lpc"]["cases"]:
        buf = _i32(c["in"])
        coefs = np.array(c["coefs"], dtype=np.int16)
        L.clxo_predict_lpc(coefs.ctypes.data_as(C.POINTER(C.c_int16)), coefs.size, c["shift"],
                           buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size)
        assert buf.tolist() == c["out"]


@pytest.mark.parametrize("name", ["decode_left_side", "decode_right_side", "decode_mid_side"])
def test_decorrelation(name):
    buf = _i32(KAT[name]["in"])
    getattr(L, "clxo_" + name)(buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size)
    assert buf.tolist() == KAT[name]["out"]


def test_mid_side_division_is_shift():
    # the device uses >> 1 where the reference divides by 2: equal because mid*2|side&1 +- side is even
    rng = np.random.default_rng(1)
    a = rng.integers(-2**31, 2**31, 20000, dtype=np.int64).astype(np.int32)
    b = rng.integers(-2**31, 2**31, 20000, dtype=np.int64).astype(np.int32)
    buf = np.concatenate([a, b])
    L.clxo_decode_mid_side(buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size)
    m = ((a.astype(np.uint32) << 1) | (b.astype(np.uint32) & 1)).astype(np.uint32)
    left = ((m + b.astype(np.uint32)).astype(np.uint32).view(np.int32)) >> 1
    right = ((m - b.astype(np.uint32)).astype(np.uint32).view(np.int32)) >> 1
    assert np.array_equal(buf[:20000], left) and np.array_equal(buf[20000:], right)


def test_var_length_int():
    k = KAT["var_length_int"]
    data = bytes(k["bytes"])
    at = 0
    for exp in k["values"]:
        v, used = C.c_uint64(0), C.c_size_t(0)
        assert L.clxo_read_var_length_int(data[at:], len(data) - at, C.byref(v), C.byref(used)) == 0
        assert v.value == exp
        at += used.value
    # two-byte integer with invalid continuation byte, then a lone continuation byte
    v, used = C.c_uint64(0), C.c_size_t(0)
    assert L.clxo_read_var_length_int(data[at:], len(data) - at, C.byref(v), C.byref(used)) == 6
    assert L.clxo_read_var_length_int(data[at + 2:], len(data) - at - 2, C.byref(v), C.byref(used)) == 6


def test_crc_vectors():
    for data, exp in KAT["crc8"]["cases"]:
        assert L.clxo_crc8(bytes(data), len(data)) == exp
    for data, exp in KAT["crc16"]["cases"]:
        assert L.clxo_crc16(bytes(data), len(data)) == exp


def _read(data, pos, kind, bits=0):
    p, v = C.c_uint64(pos), C.c_uint32(0)
    st = L.clxo_bit_read(bytes(data), len(data), C.byref(p), kind, bits, C.byref(v))
    return st, v.value, p.value


def test_bitstream_vectors():
    k = KAT["read_unary"]
    pos = 0
    for exp in k["values"]:
        st, v, pos = _read(k["bytes"], pos, 1)
        assert st == 0 and v == exp
    st, v, pos = _read(k["bytes"], pos, 0, 3)
    assert (st, v) == (0, 2)
    assert _read(k["bytes"], pos, 0, 1)[0] == 2  # UnexpectedEof
    for key in ("read_leq_u8", "read_gt_u8_leq_u16", "read_leq_u16", "read_leq_u32"):
        k = KAT[key]
        pos = 0
        for bits, exp in k["reads"]:
            st, v, pos = _read(k["bytes"], pos, 0, bits)
            assert st == 0 and v == exp, (key, bits)
        if "then_error_bits" in k:
            assert _read(k["bytes"], pos, 0, k["then_error_bits"])[0] == 2
    k = KAT["read_mixed"]
    pos = 0
    for bits, exp in k["first"]:
        st, v, pos = _read(k["bytes"], pos, 0, bits)
        assert v == exp
    for exp in k["samples17"]:
        st, v, pos = _read(k["bytes"], pos, 0, 17)
        assert L.clxo_extend_sign_u32(v, 17) == exp


def _interleaved_md5(pcm, rows, bps):
    pos, parts = 0, []
    for r in rows:
        n = int(r[4] * r[5])
        parts.append(pcm[pos:pos + n].reshape(int(r[5]), int(r[4])).T)
        pos += n
    nb = (bps + 7) // 8
    raw = np.concatenate(parts).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :nb].tobytes()
    return hashlib.md5(raw).hexdigest()


@pytest.mark.parametrize("name", ["pop", "short", "wasted_bits"])
def test_fixture_md5(golden, name):
    data = golden[f"{name}__bytes"]
    st, si, first = O.open_stream(data)
    assert st == 0
    st, nf, pcm = O.decode_stream(data, first, int(si.samples) * si.channels + 16)
    assert st == 0
    rows = golden[f"{name}__frames"]
    rows = rows[rows[:, 1] == 0]
    assert nf == len(rows)
    assert _interleaved_md5(pcm, rows, si.bits_per_sample) == KAT["fixture_md5"][name]
    assert bytes(si.md5sum).hex() == KAT["fixture_md5"][name]
    assert np.array_equal(pcm, golden[f"{name}__pcm"])


def test_fixture_facts(golden):
    # SURVEY.md Appendix C
    pop = golden["pop__pcm"]
    assert pop[:8].tolist() == [0, 2052, 4097, 6126, 8130, 10103, 12036, 13921]
    assert pop[-4:].tolist() == [-8582, -6584, -4560, -2518]
    assert golden["short__pcm"].tolist() == [2, -3, 5, -7]
    wb = golden["wasted_bits__frames"]
    assert wb[1, 7] == 314  # Block::time() quirk: short last frame reports bs * frame number
    ns = golden["non_subset__pcm"]
    assert ns[:3].tolist() == [212872, 209665, 233125] and ns[4096:4099].tolist() == [213604, 211995, 235862]


def test_fuzz_corpus_statuses(golden):
    """Expected first-frame outcome per fuzz file under a normal (CRC-checking) build."""
    exp_open = {"07d9": 30, "5a35": 30, "bb2b": 30, "c377": 30, "c6c1": 30, "d44b": 30}
    exp_frame = {"ca10": 10, "848d": 22, "9208": 13, "b6d3": 13, "1cc7": 14, "74b2": 14, "5b00": 20, "6ecc": 19,
                 "0fd7": 16, "64a1": 16, "7620": 16, "a7f0": 16, "0294": 23, "6710": 23}
    for name in golden["names"]:
        name = str(name)
        if not name.startswith("fuzz__"):
            continue
        short = name[6:10]
        meta, rows = golden[f"{name}__meta"], golden[f"{name}__frames"]
        if short in exp_open:
            assert meta[0] == exp_open[short]
        elif short in exp_frame:
            assert meta[0] == 0 and rows[0, 1] == exp_frame[short], name
        else:
            assert meta[0] != 0  # dies in metadata
        # re-decode now and compare with the stored run
        st, si, first = O.open_stream(golden[f"{name}__bytes"])
        assert st == meta[0]
        if st == 0:
            f = O.decode_frame(golden[f"{name}__bytes"], first)
            assert f.status == rows[0, 1]


def test_all_overwritten_property(golden):
    """fuzz/fuzzers/diff.rs: decoding into buffers prefilled with 13 vs 17 gives identical output."""
    for name in ("pop", "wasted_bits", "non_subset"):
        data = golden[f"{name}__bytes"]
        first = int(golden[f"{name}__meta"][1])
        a = O.decode_frame(data, first, fill=13)
        b = O.decode_frame(data, first, fill=17)
        assert a.status == 0 and np.array_equal(a.samples, b.samples)


# --------------------------------------------------------------------------- a second, independent restatement

_SPEC_SHAPES = [
    # name, SynthConfig keywords — together: every subframe type, fixed 0..4, LPC 1..32, Rice and Rice2, partition
    # orders 0..4, Rice parameters 0..14 (and beyond with Rice2), 8 / 16 / 24 bits, wasted bits, every stereo mode,
    # 1..8 channels, long unary runs
    ("stereo16-mixed", dict(n_frames=10, block_size=192, n_channels=2, bps=16, stereo_mode=-1, type_mask=15, lpc_min_order=1,
                            lpc_max_order=12, qlp_precision=0, rice_mode=-2, rice_kmin=0, rice_kmax=14, max_porder=4, wasted_max=3)),
    ("stereo24-lpc", dict(n_frames=8, block_size=256, n_channels=2, bps=24, stereo_mode=-1, type_mask=8, lpc_min_order=2,
                          lpc_max_order=32, qlp_precision=15, rice_mode=-1, rice_kmin=8, rice_kmax=14, max_porder=3)),
    ("rice2", dict(n_frames=6, block_size=128, n_channels=2, bps=24, stereo_mode=-1, type_mask=12, lpc_min_order=1,
                   lpc_max_order=8, qlp_precision=0, rice_mode=-1, rice2=1, max_porder=2, long_unary_per_mille=300)),
    ("eight-channels-8bit", dict(n_frames=4, block_size=64, n_channels=8, bps=8, stereo_mode=0, type_mask=15, lpc_min_order=1,
                                 lpc_max_order=6, qlp_precision=0, rice_mode=-1, max_porder=2, force_bs16=1)),
    ("mono-tail", dict(n_frames=5, block_size=200, tail_block_size=37, n_channels=1, bps=16, stereo_mode=0, type_mask=15,
                       lpc_min_order=1, lpc_max_order=16, qlp_precision=0, rice_mode=-1, max_porder=3, wasted_max=5)),
]


@pytest.mark.parametrize("name,kw", _SPEC_SHAPES, ids=[s[0] for s in _SPEC_SHAPES])
def test_oracle_against_independent_python_decoder(name, kw):
    """the shapes no MD5-carrying fixture reaches are pinned by a SECOND restatement of the reference —
    tests/spec_decode.py, plain Python written from the reference's control flow — which must give, frame by frame,
    what the C oracle gives (and what the generator put in): samples, bytes consumed, and on damaged frames the
    same claxon error string."""
    from claxon_b200 import synth
    import claxon_b200 as cb
    from tests import spec_decode as D
    b = synth.generate(synth.SynthConfig(seed=0xD0C0 + len(name), **kw))
    rng = np.random.default_rng(len(name))
    outcomes = set()
    for i in range(b.n_frames):
        lo, hi = int(b.frame_offsets[i]), int(b.frame_offsets[i + 1])
        frame = b.data[lo:hi].copy()
        kind, val = D.decode_frame(frame.tobytes())
        of = O.decode_frame(frame)
        assert kind == "ok" and of.status == 0, (name, i, kind, val if kind != "ok" else "", of.status)
        planar = np.array(val[0], dtype=np.int64).reshape(-1)
        assert val[1] == hi - lo == of.info.consumed
        assert np.array_equal(planar, of.samples.astype(np.int64))
        assert np.array_equal(planar, b.pcm[int(b.pcm_offsets[i]): int(b.pcm_offsets[i + 1])].astype(np.int64))
        # the same frame damaged (CRC checks off so that the damage reaches the subframe logic) and truncated
        for trial in range(6):
            bad = frame.copy()
            for _ in range(int(rng.integers(1, 3))):
                bad[int(rng.integers(4, bad.size))] = rng.integers(0, 256)
            if trial % 3 == 2:
                bad = bad[: int(rng.integers(2, bad.size))]
            kind, val = D.decode_frame(bad.tobytes(), verify_crc=False)
            of = O.decode_frame(bad, verify_crc=False)
            if kind == "ok":
                assert of.status == 0 and val[1] == of.info.consumed
                assert np.array_equal(np.array(val[0], dtype=np.int64).reshape(-1), of.samples.astype(np.int64))
            else:
                want = "UnexpectedEof" if val == D.EOF_MSG else val
                assert cb.status_str(of.status) == want, (name, i, trial, cb.status_str(of.status), want)
            outcomes.add("ok" if kind == "ok" else val)
    assert len(outcomes) >= 3, outcomes
