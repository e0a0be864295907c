"""Minimal Ogg and MP4 muxers for the container-feed tests (test infrastructure): they wrap frames of a synthetic batch
the way the FLAC-in-Ogg mapping and the 'fLaC' ISO BMFF sample entry store them."""
import struct

import numpy as np

from claxon_b200 import synth


def _ogg_crc(page: bytes) -> int:
    table = []
    for i in range(256):
        r = i << 24
        for _ in range(8):
            r = ((r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if r & 0x80000000 else (r << 1) & 0xFFFFFFFF
        table.append(r)
    c = 0
    for b in page:
        c = ((c << 8) & 0xFFFFFFFF) ^ table[((c >> 24) ^ b) & 0xFF]
    return c


def ogg_mux(packets, serial=0x1234, max_segments=40, foreign_stream=True):
    """Pages of at most `max_segments` lacing values, so that large packets span pages; optionally a second logical
    stream's pages are interleaved (to be ignored by the reader)."""
    laces = []  # (lacing value, payload bytes, first-of-packet)
    for p in packets:
        n = len(p)
        vals = [255] * (n // 255) + [n % 255]
        off = 0
        for i, v in enumerate(vals):
            laces.append((v, p[off:off + v], i == 0))
            off += v
    pages, seq = [], 0
    for i in range(0, len(laces), max_segments):
        chunk = laces[i:i + max_segments]
        continued = not chunk[0][2]
        flags = (1 if continued else 0) | (2 if i == 0 else 0) | (4 if i + max_segments >= len(laces) else 0)
        body = b"".join(c[1] for c in chunk)
        hdr = b"OggS" + bytes([0, flags]) + struct.pack("<QIII", 0, serial, seq, 0) + bytes([len(chunk)]) + bytes(c[0] for c in chunk)
        page = bytearray(hdr + body)
        page[22:26] = struct.pack("<I", _ogg_crc(bytes(page)))
        pages.append(bytes(page))
        seq += 1
        if foreign_stream and i == 0:  # a page of another logical stream right after our first page
            junk = bytearray(b"OggS" + bytes([0, 2]) + struct.pack("<QIII", 0, serial + 1, 0, 0) + bytes([1, 5]) + b"other")
            junk[22:26] = struct.pack("<I", _ogg_crc(bytes(junk)))
            pages.append(bytes(junk))
    return b"".join(pages)


def flac_in_ogg(batch: synth.SynthBatch, with_empty_packet=True) -> bytes:
    n = batch.n_frames
    flac = synth.make_file(batch, 0, n)
    streaminfo = flac[8:42]
    vc = (4).to_bytes(4, "little") + b"test" + (1).to_bytes(4, "little") + (7).to_bytes(4, "little") + b"FOO=bar"
    first = b"\x7fFLAC\x01\x00" + (1).to_bytes(2, "big") + b"fLaC" + bytes([0]) + (34).to_bytes(3, "big") + streaminfo
    header = bytes([0x80 | 4]) + len(vc).to_bytes(3, "big") + vc
    frames = [batch.data[int(batch.frame_offsets[i]):int(batch.frame_offsets[i + 1])].tobytes() for i in range(n)]
    packets = [first, header] + frames + ([b""] if with_empty_packet else [])
    return ogg_mux(packets)


def _box(kind: bytes, body: bytes) -> bytes:
    return struct.pack(">I", 8 + len(body)) + kind + body


def flac_in_mp4(batch: synth.SynthBatch, chunk_plan=(3, 3, 1, 1), co64=False) -> bytes:
    """Frames stored in chunks whose sizes cycle through `chunk_plan` (so that stsc needs several runs)."""
    n = batch.n_frames
    flac = synth.make_file(batch, 0, n)
    streaminfo = flac[8:42]
    frames = [batch.data[int(batch.frame_offsets[i]):int(batch.frame_offsets[i + 1])].tobytes() for i in range(n)]
    # chunk layout
    chunks, i, k = [], 0, 0
    while i < n:
        c = min(chunk_plan[k % len(chunk_plan)], n - i)
        chunks.append(frames[i:i + c])
        i += c
        k += 1
    dfla = _box(b"dfLa", bytes(4) + bytes([0x80]) + (34).to_bytes(3, "big") + streaminfo)
    entry = _box(b"fLaC", bytes(6) + struct.pack(">H", 1) + bytes(8) + struct.pack(">HHHHI", batch.config.n_channels, batch.config.bps, 0, 0, 44100 << 16) + dfla)
    stsd = _box(b"stsd", bytes(4) + struct.pack(">I", 1) + entry)
    stsz = _box(b"stsz", bytes(4) + struct.pack(">II", 0, n) + b"".join(struct.pack(">I", len(f)) for f in frames))
    runs, prev = [], None
    for ci, ch in enumerate(chunks):
        if len(ch) != prev:
            runs.append((ci + 1, len(ch), 1))
            prev = len(ch)
    stsc = _box(b"stsc", bytes(4) + struct.pack(">I", len(runs)) + b"".join(struct.pack(">III", *r) for r in runs))
    ftyp = _box(b"ftyp", b"isom" + bytes(4) + b"isomiso2")

    def moov(offsets):
        if co64:
            stco = _box(b"co64", bytes(4) + struct.pack(">I", len(offsets)) + b"".join(struct.pack(">Q", o) for o in offsets))
        else:
            stco = _box(b"stco", bytes(4) + struct.pack(">I", len(offsets)) + b"".join(struct.pack(">I", o) for o in offsets))
        stbl = _box(b"stbl", stsd + stsz + stsc + stco)
        other = _box(b"trak", _box(b"mdia", _box(b"minf", _box(b"stbl", _box(b"stsd", bytes(4) + struct.pack(">I", 1) + _box(b"mp4a", bytes(28)))))))
        return _box(b"moov", _box(b"mvhd", bytes(100)) + other + _box(b"trak", _box(b"tkhd", bytes(84)) + _box(b"mdia", _box(b"mdhd", bytes(24)) + _box(b"minf", stbl))))

    size0 = len(ftyp) + len(moov([0] * len(chunks))) + 8  # moov's size does not depend on the offsets' values
    offsets, at = [], size0
    for ch in chunks:
        offsets.append(at)
        at += sum(len(f) for f in ch) + 5  # 5 bytes of filler between chunks: frames are contiguous only inside a chunk
    mdat_body = b"".join(b"".join(ch) + b"\x00pad\x00" for ch in chunks)
    return ftyp + moov(offsets) + _box(b"mdat", mdat_body)
