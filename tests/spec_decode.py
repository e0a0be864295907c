"""A third, independent statement of claxon's per-frame decode, in plain Python (slow: small frames only).

Restates `FrameReader::read_next_or_eof` (reference src/frame.rs:667-779), `subframe::decode` and everything below
it (src/subframe.rs:29-91, :184-228, :236-380, :382-474, :492-614, :651-721), the decorrelation
(src/frame.rs:319-389) and the frame CRC-16 (src/crc.rs: polynomial 0x8005, initial value 0, bit by bit here) from the
reference's own control flow, with claxon's error strings.  The oracle (`oracle/claxon_oracle.c`) is pinned by the
reference's known-answer vectors and by three MD5-carrying fixtures; the shapes those do not reach (24-bit, Rice
parameters above 8, partition orders above 1, LPC orders 2..32, every stereo mode, wasted bits, Rice2) are pinned
by comparing the oracle with THIS on synthetic frames (SURVEY.md §8c asks for two independent restatements).

    decode_frame(buf) -> ("eof", None) | ("err", message) | ("ok", (planar samples per channel, bytes consumed))
"""
from __future__ import annotations

from tests import spec_header

EOF_MSG = "unexpected eof"
M32 = 0xFFFFFFFF


class _Eof(Exception):
    pass


class _Fmt(Exception):
    pass


def _i32(v: int) -> int:
    v &= M32
    return v - (1 << 32) if v & 0x80000000 else v


def crc16(data: bytes) -> int:
    crc = 0
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x8005) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


class _Bits:
    """MSB-first bit reader (src/input.rs:415-643: only the bit order and the widths matter)."""

    def __init__(self, buf: bytes, byte_pos: int):
        self.b, self.p = buf, byte_pos * 8

    def read(self, n: int) -> int:
        v = 0
        for _ in range(n):
            if self.p >= len(self.b) * 8:
                raise _Eof()
            v = (v << 1) | ((self.b[self.p >> 3] >> (7 - (self.p & 7))) & 1)
            self.p += 1
        return v

    def unary(self) -> int:
        n = 0
        while self.read(1) == 0:
            n += 1
        return n

    def byte_aligned_end(self) -> int:
        return (self.p + 7) >> 3


def _sext(v: int, bits: int) -> int:
    return v - (1 << bits) if v & (1 << (bits - 1)) else v


def _rice_to_signed(v: int) -> int:
    return _i32((v >> 1) ^ (M32 if v & 1 else 0))


def _residual(r: _Bits, block_size: int, n_warm_up: int):
    method = r.read(2)
    if method > 1:
        raise _Fmt("invalid residual, encountered reserved value")
    order = r.read(4)
    n_part = 1 << order
    per = block_size >> order
    if block_size & ((n_part - 1) & 0xFFFF):
        raise _Fmt("invalid partition order")
    if n_warm_up > per:
        raise _Fmt("invalid residual")
    pbits = 4 if method == 0 else 5
    out = []
    count = per - n_warm_up
    for _ in range(n_part):
        k = r.read(pbits)
        if k == (1 << pbits) - 1:
            raise _Fmt("unencoded binary is not yet implemented")
        for _ in range(count):
            q = r.unary()
            rem = r.read(k)
            out.append(_rice_to_signed(((q << k) & M32) | rem))
        count = per
    return out


_FIXED = {0: [], 1: [1], 2: [-1, 2], 3: [1, -3, 3], 4: [-1, 4, -6, 4]}


def _subframe(r: _Bits, bps: int, bs: int):
    if r.read(1):
        raise _Fmt("invalid subframe header")
    code = r.read(6)
    if code == 0:
        kind, order = "constant", 0
    elif code == 1:
        kind, order = "verbatim", 0
    elif (code & 0b111110) == 0b000010 or (code & 0b111100) == 0b000100 or (code & 0b110000) == 0b010000:
        raise _Fmt("invalid subframe header, encountered reserved value")
    elif (code & 0b111000) == 0b001000:
        order = code & 7
        if order > 4:
            raise _Fmt("invalid subframe header, encountered reserved value")
        kind = "fixed"
    else:
        kind, order = "lpc", (code & 0b011111) + 1
    wasted = 0
    if r.read(1):
        wasted = 1 + r.unary()
    if wasted > 31:
        raise _Fmt("wasted bits per sample must not exceed 31")
    if wasted >= bps:
        raise _Fmt("subframe has no non-wasted bits")
    sf_bps = bps - wasted
    if kind == "constant":
        s = [_sext(r.read(sf_bps), sf_bps)] * bs
    elif kind == "verbatim":
        s = [_sext(r.read(sf_bps), sf_bps) for _ in range(bs)]
    elif kind == "fixed":
        if bs < order:
            raise _Fmt("invalid fixed subframe, order is larger than block size")
        s = [_sext(r.read(sf_bps), sf_bps) for _ in range(order)]
        s += _residual(r, bs, order)
        c = _FIXED[order]
        for i in range(bs - order):  # Wrapping<i32> arithmetic throughout
            pred = 0
            for cj, sj in zip(c, s[i:i + order]):
                pred = _i32(pred + _i32(cj * sj))
            s[i + order] = _i32(pred + s[i + order])
    else:
        if bs < order:
            raise _Fmt("invalid LPC subframe, lpc order is larger than block size")
        s = [_sext(r.read(sf_bps), sf_bps) for _ in range(order)]
        precision = r.read(4) + 1
        if precision - 1 == 15:
            raise _Fmt("invalid subframe, qlp precision value invalid")
        shift = _sext(r.read(5), 5)
        if shift < 0:
            raise _Fmt("a negative quantized linear predictor coefficient shift is not supported, please file a bug.")
        coefs = [0] * order
        for j in reversed(range(order)):  # the first one read multiplies the most recent sample
            coefs[j] = _sext(r.read(precision), precision)
        s += _residual(r, bs, order)
        for i in range(order, bs):  # i64 sum, arithmetic shift, i64 add, truncating cast
            acc = sum(cj * sj for cj, sj in zip(coefs, s[i - order:i]))
            s[i] = _i32((acc >> shift) + s[i])
    if wasted:
        s = [_i32(v << wasted) for v in s]
    return s


def decode_frame(buf, verify_crc: bool = True):
    buf = bytes(buf)
    kind, hdr = spec_header.parse(buf, verify_crc)
    if kind != "ok":
        return kind, hdr
    try:
        if hdr["bits_per_sample"] == 0:
            raise _Fmt("header without bits per sample info")
        bps, bs, ca = hdr["bits_per_sample"], hdr["block_size"], hdr["channel_assignment"]
        r = _Bits(buf, hdr["header_len"])
        if ca < 8:
            ch = [_subframe(r, bps, bs) for _ in range(ca + 1)]
        elif ca == 8:  # left, side
            left = _subframe(r, bps, bs)
            side = _subframe(r, bps + 1, bs)
            ch = [left, [_i32(a - b) for a, b in zip(left, side)]]
        elif ca == 9:  # side, right
            side = _subframe(r, bps + 1, bs)
            right = _subframe(r, bps, bs)
            ch = [[_i32(a + b) for a, b in zip(side, right)], right]
        else:          # mid, side
            mid = _subframe(r, bps, bs)
            side = _subframe(r, bps + 1, bs)
            lefts, rights = [], []
            for m, sd in zip(mid, side):
                m2 = _i32(_i32(m * 2) | (sd & 1))
                a, b = _i32(m2 + sd), _i32(m2 - sd)
                lefts.append(int(a / 2) if a % 2 else a // 2)   # Rust `/`: truncation (the operands are even anyway)
                rights.append(int(b / 2) if b % 2 else b // 2)
            ch = [lefts, rights]
        end = r.byte_aligned_end()           # pad bits are skipped unchecked
        if end + 2 > len(buf):
            raise _Eof()
        if verify_crc and crc16(buf[:end]) != ((buf[end] << 8) | buf[end + 1]):
            raise _Fmt("frame CRC mismatch")
    except _Eof:
        return "err", EOF_MSG
    except _Fmt as e:
        return "err", str(e)
    return "ok", (ch, end + 2)
