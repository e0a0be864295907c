"""A third, independent statement of claxon's stream open (metadata walk), in plain Python, for the host tests.

Restates `read_stream_header` + `FlacReader::new_ext` with default options (reference src/lib.rs:186-307) and the
block readers of src/metadata.rs:212-545 from the reference's own control flow, returning claxon's error strings.
See tests/spec_header.py for why a third statement exists.

    open_stream(buf) -> ("err", message) | ("ok", dict(first_frame=..., streaminfo fields ..., vendor=..., comments=[...]))

`"err", "unexpected eof"` stands for io::ErrorKind::UnexpectedEof.
"""
from __future__ import annotations

EOF_MSG = "unexpected eof"


class _Eof(Exception):
    pass


class _Fmt(Exception):
    pass


class _R:
    def __init__(self, buf):
        self.b, self.p = bytes(buf), 0

    def take(self, n: int) -> bytes:
        if self.p + n > len(self.b):
            self.p = len(self.b)
            raise _Eof()
        self.p += n
        return self.b[self.p - n:self.p]

    def be(self, n: int) -> int:
        return int.from_bytes(self.take(n), "big")

    def le32(self) -> int:
        return int.from_bytes(self.take(4), "little")


def _utf8(raw: bytes) -> str:
    try:
        return raw.decode("utf-8")
    except UnicodeDecodeError:
        raise _Fmt("Vorbis comment or vendor string is not valid UTF-8")


def _streaminfo(r: _R) -> dict:
    min_bs, max_bs = r.be(2), r.be(2)
    min_fs, max_fs = r.be(3), r.be(3)
    sr_msb, sr_lsb = r.be(2), r.be(1)
    sample_rate = (sr_msb << 4) | (sr_lsb >> 4)
    channels = ((sr_lsb >> 1) & 7) + 1
    b2 = r.be(1)
    bps = (((sr_lsb & 1) << 4) | (b2 >> 4)) + 1
    samples = ((b2 & 15) << 32) | r.be(4)
    md5 = r.take(16)
    if min_bs > max_bs:
        raise _Fmt("inconsistent bounds, min block size > max block size")
    if min_bs < 16:
        raise _Fmt("invalid block size, must be at least 16")
    if min_fs > max_fs and max_fs != 0:
        raise _Fmt("inconsistent bounds, min frame size > max frame size")
    if sample_rate == 0 or sample_rate > 655350:
        raise _Fmt("invalid sample rate")
    return dict(min_block_size=min_bs, max_block_size=max_bs, min_frame_size=min_fs, max_frame_size=max_fs,
                sample_rate=sample_rate, channels=channels, bits_per_sample=bps, samples=samples, md5sum=md5)


def _vorbis_comment(r: _R, length: int):
    if length < 8:
        raise _Fmt("Vorbis comment block is too short")
    if length > 10 * 1024 * 1024:
        raise _Fmt("Vorbis comment blocks larger than 10 MiB are not supported")
    vendor_len = r.le32()
    if vendor_len > length - 8:
        raise _Fmt("vendor string too long")
    vendor = _utf8(r.take(vendor_len))
    want = r.le32()
    if want >= length // 4:
        raise _Fmt("too many entries for Vorbis comment block")
    left = length - 8 - vendor_len
    comments = []
    while left >= 4 and len(comments) < want:
        n = r.le32()
        left -= 4
        if n > left:
            raise _Fmt("Vorbis comment too long for Vorbis comment block")
        if n == 0:
            want -= 1
            continue
        raw = r.take(n)
        left -= n
        sep = raw.find(b"=")
        if sep < 0:
            raise _Fmt("Vorbis comment does not contain '='")
        if any(x < 0x20 or x > 0x7D for x in raw[:sep]):
            raise _Fmt("Vorbis comment field name contains invalid byte")
        comments.append((_utf8(raw), sep))
    if left != 0:
        raise _Fmt("Vorbis comment block has excess data")
    if len(comments) != want:
        raise _Fmt("Vorbis comment block contains wrong number of entries")
    return vendor, comments


def _block(r: _R, typ: int, length: int):
    if typ == 0:
        if length != 34:
            raise _Fmt("invalid streaminfo metadata block length")
        return "streaminfo", _streaminfo(r)
    if typ == 2:
        if length < 4:
            raise _Fmt("application block length must be at least 4 bytes")
        if length > 10 * 1024 * 1024:
            raise _Fmt("application blocks larger than 10 MiB are not supported")
        r.take(length)
        return "other", None
    if typ == 4:
        return "vorbis", _vorbis_comment(r, length)
    if typ == 127:
        raise _Fmt("invalid metadata block type")
    r.take(length)  # padding, seek table, cue sheet, picture, reserved types: skipped
    return "other", None


def open_stream(buf):
    r = _R(buf)
    try:
        magic = r.be(4)
        if magic != 0x664C6143:
            if magic & 0xFFFFFF00 == 0x49443300:
                raise _Fmt("stream starts with ID3 header rather than FLAC header")
            raise _Fmt("invalid stream header")
        info, vorbis, first = None, None, True
        while True:
            head = r.be(1)
            last, typ = head >> 7, head & 0x7F
            length = r.be(3)
            kind, val = _block(r, typ, length)
            if first:
                if kind != "streaminfo":
                    raise _Fmt("streaminfo block missing")
                info = val
            elif kind == "vorbis":
                if vorbis is not None:
                    raise _Fmt("encountered second Vorbis comment block")
                vorbis = val
            elif kind == "streaminfo":
                raise _Fmt("encountered second streaminfo block")
            first = False
            if last:
                break
    except _Eof:
        return "err", EOF_MSG
    except _Fmt as e:
        return "err", str(e)
    out = dict(info)
    out["first_frame"] = r.p
    out["vendor"], out["comments"] = vorbis if vorbis is not None else (None, [])
    return "ok", out
