"""A third, independent statement of claxon's frame-header parse, in plain Python, for the host tests.

The product's `clx_parse_frame_header` (C++) and the oracle's `clxo_read_frame_header` (C) were written by the same
hand in similar styles; comparing them shows little.  This module restates `read_frame_header_or_eof`
(reference src/frame.rs:131-316), `read_var_length_int` (:64-105) and the CRC-8 (src/crc.rs: polynomial 0x07,
initial value 0, no reflection) directly from the reference's control flow — bit by bit for the CRC, no tables — and
returns claxon's own error strings, so a test can check both the product and the oracle against it through
`clx_status_str`.

    parse(buf) -> ("eof", None) | ("err", message) | ("ok", fields)

`"err", "unexpected eof"` stands for claxon's io::ErrorKind::UnexpectedEof.
"""
from __future__ import annotations

EOF_MSG = "unexpected eof"


class _Eof(Exception):
    pass


class _Fmt(Exception):
    pass


def crc8(data: bytes) -> int:
    crc = 0
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = ((crc << 1) ^ 0x07) & 0xFF if crc & 0x80 else (crc << 1) & 0xFF
    return crc


class _Reader:
    def __init__(self, buf: bytes):
        self.buf, self.pos = bytes(buf), 0

    def u8(self) -> int:
        if self.pos >= len(self.buf):
            raise _Eof()
        self.pos += 1
        return self.buf[self.pos - 1]

    def be_u16(self) -> int:
        hi = self.u8()
        return (hi << 8) | self.u8()


def _var_length_int(r: _Reader) -> int:
    first = r.u8()
    extra, mask_data, mask_mark = 0, 0x7F, 0x80
    while first & mask_mark:
        extra += 1
        mask_data >>= 1
        mask_mark >>= 1
    if extra > 0:
        if extra == 1:
            raise _Fmt("invalid variable-length integer")
        extra -= 1
    result = (first & mask_data) << (6 * extra)
    for i in reversed(range(extra)):
        byte = r.u8()
        if byte & 0xC0 != 0x80:
            raise _Fmt("invalid variable-length integer")
        result |= (byte & 0x3F) << (6 * i)
    return result


_SAMPLE_RATES = {1: 88200, 2: 176400, 3: 192000, 4: 8000, 5: 16000, 6: 22050, 7: 24000, 8: 32000, 9: 44100,
                 10: 48000, 11: 96000}
_BPS = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24}
RESERVED = "invalid frame header, encountered reserved value"


def parse(buf, verify_crc: bool = True):
    r = _Reader(buf)
    # read_be_u16_or_eof (src/input.rs:94-101): if either of the two bytes is missing, that is the end of the stream
    if len(r.buf) < 2:
        return "eof", None
    try:
        sync = r.be_u16()
        if sync & 0xFFFC != 0xFFF8:
            raise _Fmt("frame sync code missing")
        if sync & 0x0002:
            raise _Fmt(RESERVED)
        variable = bool(sync & 1)
        bs_sr = r.u8()
        block_size, bs8, bs16 = 0, False, False
        code = bs_sr >> 4
        if code == 0:
            raise _Fmt(RESERVED)
        elif code == 1:
            block_size = 192
        elif 2 <= code <= 5:
            block_size = 576 << (code - 2)
        elif code == 6:
            bs8 = True
        elif code == 7:
            bs16 = True
        else:
            block_size = 256 << (code - 8)
        sr_code = bs_sr & 15
        sample_rate, sr8, sr16, sr16ten = 0, False, False, False
        if sr_code == 0:
            sample_rate = 0  # "get from streaminfo"
        elif sr_code in _SAMPLE_RATES:
            sample_rate = _SAMPLE_RATES[sr_code]
        elif sr_code == 12:
            sr8 = True
        elif sr_code == 13:
            sr16 = True
        elif sr_code == 14:
            sr16ten = True
        else:
            raise _Fmt("invalid frame header")
        cbr = r.u8()
        ca = cbr >> 4
        if ca < 8:
            n_channels, assignment = ca + 1, ca  # (the descriptor keeps the raw code: 0..7 = n + 1 independent channels)
        elif ca <= 10:
            n_channels, assignment = 2, ca
        else:
            raise _Fmt(RESERVED)
        bps_code = (cbr & 0x0E) >> 1
        if bps_code == 0:
            bps = 0  # "get from streaminfo"
        elif bps_code in _BPS:
            bps = _BPS[bps_code]
        else:
            raise _Fmt(RESERVED)
        if cbr & 1:
            raise _Fmt(RESERVED)
        number = _var_length_int(r)
        if not variable and number > 0x7FFFFFFF:
            raise _Fmt("invalid frame header, frame number too large")
        if bs8:
            block_size = r.u8() + 1
        if bs16:
            v = r.be_u16()
            if v == 0xFFFF:
                raise _Fmt("invalid block size, exceeds 65535")
            block_size = v + 1
        if sr8:
            sample_rate = r.u8()
        if sr16:
            sample_rate = r.be_u16()
        if sr16ten:
            sample_rate = r.be_u16() * 10
        computed = crc8(r.buf[: r.pos])
        presumed = r.u8()
        if verify_crc and computed != presumed:
            raise _Fmt("frame header CRC mismatch")
    except _Eof:
        return "err", EOF_MSG
    except _Fmt as e:
        return "err", str(e)
    return "ok", dict(block_size=block_size, sample_rate=sample_rate, n_channels=n_channels, channel_assignment=assignment,
                      bits_per_sample=bps, variable=int(variable), number=number, header_len=r.pos)
