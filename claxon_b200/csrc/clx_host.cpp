// clx_host.cpp — host side of claxon_b200: everything claxon does OUTSIDE the per-frame
// sample arithmetic, restated for flat in-memory byte spans:
//   * frame header parse            reference src/frame.rs:64-105, :131-316
//   * CRC-8 / CRC-16                src/crc.rs:13-112
//   * stream open / metadata walk   src/lib.rs:186-307, src/metadata.rs:214-609
//   * frame demultiplexer (new: claxon learns a frame's length only by decoding it,
//     src/frame.rs:667-779; a batched decoder needs boundaries up front)
// No CUDA in this file.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "claxon_b200.h"

extern "C" {

const char* clx_status_str(int status) { return clx_status_str_inline(status); }
int clx_status_kind(int status) { return clx_status_kind_inline(status); }
uint32_t clx_abi_version(void) { return CLX_ABI_VERSION; }

}  // extern "C"

namespace {

// ---- CRC tables: generated, MSB-first, init 0 (src/crc.rs:13-57) ----
struct CrcTables {
    uint8_t t8[256];
    uint16_t t16[8][256];  // slicing-by-8 for the frame CRC
    CrcTables() {
        for (int i = 0; i < 256; i++) {
            uint8_t c = (uint8_t)i;
            uint16_t d = (uint16_t)(i << 8);
            for (int k = 0; k < 8; k++) {
                c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
                d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1));
            }
            t8[i] = c;
            t16[0][i] = d;
        }
        for (int s = 1; s < 8; s++)
            for (int i = 0; i < 256; i++) {
                uint16_t v = t16[s - 1][i];
                t16[s][i] = (uint16_t)((v << 8) ^ t16[0][v >> 8]);
            }
    }
};
const CrcTables& tabs() {
    static const CrcTables t;
    return t;
}

inline uint16_t crc16_update(uint16_t s, const uint8_t* p, size_t n) {
    const CrcTables& T = tabs();
    while (n >= 8) {
        // state is 16 bits wide: it only mixes into the first two bytes of the block
        uint8_t b0 = (uint8_t)(p[0] ^ (s >> 8)), b1 = (uint8_t)(p[1] ^ (s & 0xff));
        s = (uint16_t)(T.t16[7][b0] ^ T.t16[6][b1] ^ T.t16[5][p[2]] ^ T.t16[4][p[3]] ^ T.t16[3][p[4]] ^
                       T.t16[2][p[5]] ^ T.t16[1][p[6]] ^ T.t16[0][p[7]]);
        p += 8;
        n -= 8;
    }
    while (n--) s = (uint16_t)((s << 8) ^ T.t16[0][(uint8_t)(s >> 8) ^ *p++]);
    return s;
}

// ---- varint (src/frame.rs:64-105) ----
int read_varint(const uint8_t* p, size_t n, uint64_t* value, size_t* used) {
    if (n < 1) return CLX_ERR_IO_UNEXPECTED_EOF;
    const uint8_t first = p[0];
    uint32_t ones = 0;
    while (ones < 8 && (first & (0x80u >> ones))) ones++;
    if (ones == 1) return CLX_ERR_VARINT_INVALID;
    const uint32_t extra = ones >= 2 ? ones - 1 : 0;
    const uint8_t mask = ones == 0 ? 0x7f : (uint8_t)(0x7fu >> ones);
    uint64_t v = ((uint64_t)(first & mask)) << (6 * extra);
    for (uint32_t i = 0; i < extra; i++) {
        if (1 + i >= n) return CLX_ERR_IO_UNEXPECTED_EOF;
        const uint8_t b = p[1 + i];
        if ((b & 0xc0) != 0x80) return CLX_ERR_VARINT_INVALID;
        v |= ((uint64_t)(b & 0x3f)) << (6 * (extra - 1 - i));
    }
    *value = v;
    *used = 1 + extra;
    return CLX_OK;
}

bool utf8_valid(const uint8_t* s, size_t n) {  // Rust String::from_utf8 acceptance set
    size_t i = 0;
    while (i < n) {
        const uint8_t c = s[i];
        if (c < 0x80) { i++; continue; }
        size_t need;
        uint8_t lo = 0x80, hi = 0xbf;
        if (c >= 0xc2 && c <= 0xdf) need = 1;
        else if (c == 0xe0) { need = 2; lo = 0xa0; }
        else if ((c >= 0xe1 && c <= 0xec) || c == 0xee || c == 0xef) need = 2;
        else if (c == 0xed) { need = 2; hi = 0x9f; }
        else if (c == 0xf0) { need = 3; lo = 0x90; }
        else if (c >= 0xf1 && c <= 0xf3) need = 3;
        else if (c == 0xf4) { need = 3; hi = 0x8f; }
        else return false;
        if (i + need >= n) return false;
        if (s[i + 1] < lo || s[i + 1] > hi) return false;
        for (size_t k = 2; k <= need; k++)
            if ((s[i + k] & 0xc0) != 0x80) return false;
        i += need + 1;
    }
    return true;
}

uint32_t le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Validates a VORBIS_COMMENT block exactly as src/metadata.rs:402-513 would.
int check_vorbis(const uint8_t* p, size_t avail, uint32_t length) {
    if (length < 8) return CLX_ERR_VORBIS_TOO_SHORT;
    if (length > 10u * 1024 * 1024) return CLX_ERR_VORBIS_TOO_LARGE;
    size_t at = 0;
    auto need = [&](size_t k) { return at + k <= avail; };
    if (!need(4)) return CLX_ERR_IO_UNEXPECTED_EOF;
    const uint32_t vendor_len = le32(p + at);
    at += 4;
    if (vendor_len > length - 8) return CLX_ERR_VENDOR_TOO_LONG;
    if (!need(vendor_len)) return CLX_ERR_IO_UNEXPECTED_EOF;
    if (!utf8_valid(p + at, vendor_len)) return CLX_ERR_UTF8_INVALID;
    at += vendor_len;
    if (!need(4)) return CLX_ERR_IO_UNEXPECTED_EOF;
    uint32_t comments_len = le32(p + at);
    at += 4;
    if (comments_len >= length / 4) return CLX_ERR_VORBIS_TOO_MANY;
    uint32_t bytes_left = length - 8 - vendor_len, got = 0;
    while (bytes_left >= 4 && got < comments_len) {
        if (!need(4)) return CLX_ERR_IO_UNEXPECTED_EOF;
        const uint32_t clen = le32(p + at);
        at += 4;
        bytes_left -= 4;
        if (clen > bytes_left) return CLX_ERR_VORBIS_COMMENT_TOO_LONG;
        if (clen == 0) { comments_len--; continue; }
        if (!need(clen)) return CLX_ERR_IO_UNEXPECTED_EOF;
        const uint8_t* c = p + at;
        at += clen;
        bytes_left -= clen;
        const uint8_t* eq = (const uint8_t*)memchr(c, '=', clen);
        if (!eq) return CLX_ERR_VORBIS_NO_EQUALS;
        for (const uint8_t* q = c; q < eq; q++)
            if (*q < 0x20 || *q > 0x7d) return CLX_ERR_VORBIS_NAME_INVALID;
        if (!utf8_valid(c, clen)) return CLX_ERR_UTF8_INVALID;
        got++;
    }
    if (bytes_left != 0) return CLX_ERR_VORBIS_EXCESS_DATA;
    if (got != comments_len) return CLX_ERR_VORBIS_WRONG_COUNT;
    return CLX_OK;
}

}  // namespace

extern "C" {

uint8_t clx_crc8(const uint8_t* p, size_t n) {
    const CrcTables& T = tabs();
    uint8_t s = 0;
    for (size_t i = 0; i < n; i++) s = T.t8[s ^ p[i]];
    return s;
}

uint16_t clx_crc16(const uint8_t* p, size_t n) { return crc16_update(0, p, n); }

// read_frame_header_or_eof, src/frame.rs:131-316.  Error precedence follows the order in
// which claxon reads and checks the fields.
int clx_parse_frame_header(const uint8_t* p, size_t n, clx_frame_desc* d, uint32_t flags) {
    memset(d, 0, sizeof *d);
    if (n < 2) return CLX_EOF;  // :140-143 (also the one-byte-left case, src/input.rs:94-101)
    const uint32_t sync = ((uint32_t)p[0] << 8) | p[1];
    if ((sync & 0xfffc) != 0xfff8) return CLX_ERR_SYNC_MISSING;
    if (sync & 2) return CLX_ERR_FRAME_HEADER_RESERVED;
    d->flags = (uint8_t)(sync & 1);  // CLX_FRAME_VARIABLE_BLOCKING
    size_t at = 2;
    if (at + 1 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
    const uint8_t bs_sr = p[at++];
    const uint32_t bs_code = bs_sr >> 4, sr_code = bs_sr & 15;
    uint32_t block_size = 0;
    bool bs8 = false, bs16 = false;
    if (bs_code == 0) return CLX_ERR_FRAME_HEADER_RESERVED;
    else if (bs_code == 1) block_size = 192;
    else if (bs_code <= 5) block_size = 576u << (bs_code - 2);
    else if (bs_code == 6) bs8 = true;
    else if (bs_code == 7) bs16 = true;
    else block_size = 256u << (bs_code - 8);
    static const uint32_t kRates[12] = {0, 88200, 176400, 192000, 8000, 16000,
                                        22050, 24000, 32000, 44100, 48000, 96000};
    bool sr8 = false, sr16 = false, sr16x10 = false;
    if (sr_code < 12) d->sample_rate = kRates[sr_code];
    else if (sr_code == 12) sr8 = true;
    else if (sr_code == 13) sr16 = true;
    else if (sr_code == 14) sr16x10 = true;
    else return CLX_ERR_FRAME_HEADER_INVALID;
    if (at + 1 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
    const uint8_t cbr = p[at++];
    const uint32_t ch = cbr >> 4;
    if (ch < 8) d->n_channels = (uint8_t)(ch + 1);
    else if (ch <= 10) d->n_channels = 2;
    else return CLX_ERR_FRAME_HEADER_RESERVED;
    d->channel_assignment = (uint8_t)ch;
    switch ((cbr >> 1) & 7) {
    case 0: d->bits_per_sample = 0; break;
    case 1: d->bits_per_sample = 8; break;
    case 2: d->bits_per_sample = 12; break;
    case 4: d->bits_per_sample = 16; break;
    case 5: d->bits_per_sample = 20; break;
    case 6: d->bits_per_sample = 24; break;
    default: return CLX_ERR_FRAME_HEADER_RESERVED;
    }
    if (cbr & 1) return CLX_ERR_FRAME_HEADER_RESERVED;
    size_t used = 0;
    int st = read_varint(p + at, n - at, &d->number, &used);
    if (st) return st;
    at += used;
    if (!(d->flags & CLX_FRAME_VARIABLE_BLOCKING) && d->number > 0x7fffffffull) return CLX_ERR_FRAME_NUMBER_TOO_LARGE;
    if (bs8) {
        if (at + 1 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
        block_size = (uint32_t)p[at++] + 1;
    }
    if (bs16) {
        if (at + 2 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
        const uint32_t v = ((uint32_t)p[at] << 8) | p[at + 1];
        at += 2;
        if (v == 0xffff) return CLX_ERR_BLOCK_SIZE_65535;
        block_size = v + 1;
    }
    if (sr8) {
        if (at + 1 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
        d->sample_rate = p[at++];
    }
    if (sr16 || sr16x10) {
        if (at + 2 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
        d->sample_rate = (((uint32_t)p[at] << 8) | p[at + 1]) * (sr16x10 ? 10u : 1u);
        at += 2;
    }
    if (at + 1 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
    const uint8_t computed = clx_crc8(p, at);
    const uint8_t stored = p[at++];
    d->block_size = (uint16_t)block_size;
    d->header_len = (uint16_t)at;
    if (!(flags & CLX_OPT_NO_VERIFY_CRC) && computed != stored) return CLX_ERR_HEADER_CRC_MISMATCH;
    return CLX_OK;
}

// FlacReader::new_ext with default options (src/lib.rs:230-307) on a byte span.
int clx_open_stream(const uint8_t* p, size_t n, clx_streaminfo* si, uint64_t* first_frame) {
    return clx_open_stream_ex(p, n, 0, si, first_frame, nullptr, nullptr);
}

// FlacReader::new_ext (src/lib.rs:230-307) with FlacReaderOptions (src/lib.rs:123-170): the metadata walk stops
// as soon as every desired block has been read (`opts_current.has_desired_blocks()`), and a Vorbis comment
// block that was read but not asked for is dropped again.
int clx_open_stream_ex(const uint8_t* p, size_t n, uint32_t open_flags, clx_streaminfo* si, uint64_t* first_frame,
                       uint64_t* vc_offset, uint32_t* vc_length) {
    const bool metadata_only = (open_flags & CLX_OPEN_METADATA_ONLY) != 0;
    bool want_vc = !(open_flags & CLX_OPEN_NO_VORBIS_COMMENT);  // opts_current.read_vorbis_comment
    const bool keep_vc = want_vc;
    if (vc_offset) *vc_offset = 0;
    if (vc_length) *vc_length = 0;
    if (first_frame) *first_frame = 0;
    memset(si, 0, sizeof *si);
    if (n < 4) return CLX_ERR_IO_UNEXPECTED_EOF;
    const uint32_t magic = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    if (magic != 0x664c6143u)  // "fLaC", src/lib.rs:186-205
        return (magic & 0xffffff00u) == 0x49443300u ? CLX_ERR_STREAM_HEADER_ID3 : CLX_ERR_STREAM_HEADER_INVALID;
    size_t at = 4;
    bool first = true, have_vc = false;
    bool first_block_only_seen = true;  // only the streaminfo block has been handled so far
    for (;;) {
        if (at + 4 > n) return CLX_ERR_IO_UNEXPECTED_EOF;  // block header, src/metadata.rs:214-231
        const bool is_last = (p[at] >> 7) != 0;
        const uint32_t type = p[at] & 0x7f;
        const uint32_t length = ((uint32_t)p[at + 1] << 16) | ((uint32_t)p[at + 2] << 8) | p[at + 3];
        at += 4;
        const uint8_t* body = p + at;
        const size_t avail = n - at;
        bool is_si = false, is_vc = false;
        int st = CLX_OK;
        switch (type) {
        case 0: {  // STREAMINFO, src/metadata.rs:321-400
            if (length != 34) { st = CLX_ERR_STREAMINFO_LENGTH; break; }
            if (avail < 34) { st = CLX_ERR_IO_UNEXPECTED_EOF; break; }
            clx_streaminfo t;
            memset(&t, 0, sizeof t);
            t.min_block_size = ((uint32_t)body[0] << 8) | body[1];
            t.max_block_size = ((uint32_t)body[2] << 8) | body[3];
            t.min_frame_size = ((uint32_t)body[4] << 16) | ((uint32_t)body[5] << 8) | body[6];
            t.max_frame_size = ((uint32_t)body[7] << 16) | ((uint32_t)body[8] << 8) | body[9];
            t.sample_rate = ((uint32_t)body[10] << 12) | ((uint32_t)body[11] << 4) | (body[12] >> 4);
            t.channels = ((body[12] >> 1) & 7u) + 1;
            t.bits_per_sample = ((((uint32_t)body[12] & 1u) << 4) | (body[13] >> 4)) + 1;
            t.samples = ((uint64_t)(body[13] & 15) << 32) | ((uint64_t)body[14] << 24) |
                        ((uint64_t)body[15] << 16) | ((uint64_t)body[16] << 8) | body[17];
            memcpy(t.md5sum, body + 18, 16);
            if (t.min_block_size > t.max_block_size) { st = CLX_ERR_BLOCK_SIZE_BOUNDS; break; }
            if (t.min_block_size < 16) { st = CLX_ERR_BLOCK_SIZE_LT_16; break; }
            if (t.min_frame_size > t.max_frame_size && t.max_frame_size != 0) { st = CLX_ERR_FRAME_SIZE_BOUNDS; break; }
            if (t.sample_rate == 0 || t.sample_rate > 655350) { st = CLX_ERR_SAMPLE_RATE_INVALID; break; }
            if (first) *si = t;
            is_si = true;
            break;
        }
        case 2:  // APPLICATION, src/metadata.rs:524-549
            if (length < 4) st = CLX_ERR_APPLICATION_TOO_SHORT;
            else if (length > 10u * 1024 * 1024) st = CLX_ERR_APPLICATION_TOO_LARGE;
            else if (avail < length) st = CLX_ERR_IO_UNEXPECTED_EOF;
            break;
        case 4:
            st = check_vorbis(body, avail, length);
            is_vc = true;
            if (st == CLX_OK && keep_vc) {
                if (vc_offset) *vc_offset = at;
                if (vc_length) *vc_length = length;
            }
            break;
        case 127:
            st = CLX_ERR_METADATA_BLOCK_TYPE;
            break;
        default:  // padding, seek table, cue sheet, picture, reserved: skipped by length
            if (avail < length) st = CLX_ERR_IO_UNEXPECTED_EOF;
            break;
        }
        if (st) return st;
        if (first) {
            if (!is_si) return CLX_ERR_STREAMINFO_MISSING;  // src/lib.rs:244-248
            first = false;
        } else {
            if (is_vc) {
                if (have_vc) return CLX_ERR_SECOND_VORBIS_COMMENT;
                have_vc = true;
                want_vc = false;  // "We have one, no new one is desired."
            }
            if (is_si) return CLX_ERR_SECOND_STREAMINFO;
        }
        at += length;
        if (is_last) break;
        // early-out once all desired blocks have been collected (src/lib.rs:275-279; the reference makes this
        // test after every block that follows the streaminfo block)
        if (!first_block_only_seen && metadata_only && !want_vc) return CLX_OK;
        first_block_only_seen = false;
    }
    if (first_frame) *first_frame = metadata_only ? 0 : at;
    return CLX_OK;
}

// Frame demultiplexer.  A frame ends where (a) the bytes look like the next frame's sync code
// and (b) the CRC-16 of everything before the two preceding bytes equals those two bytes — or
// at the end of the stream under the same CRC condition.  CRC-16 makes a false boundary inside
// residual data a < 2^-30 event per candidate.  The decode itself re-derives the frame length
// (`clx_frame_result.consumed`), so a wrong guess is detected, never silently accepted.
size_t clx_demux_frames(const uint8_t* bytes, size_t n, uint64_t start, clx_frame_desc* descs, size_t max_frames,
                        uint64_t* next_offset, uint64_t* total_out_elems, int* stop_status, uint32_t flags) {
    size_t count = 0;
    uint64_t pos = start, out_at = total_out_elems ? *total_out_elems : 0;
    int stop = CLX_OK;
    if (start > n) { stop = CLX_EOF; max_frames = 0; }
    while (count < max_frames) {
        clx_frame_desc d;
        const int st = clx_parse_frame_header(bytes + pos, n - pos, &d, flags);
        if (st != CLX_OK) { stop = st; break; }
        d.byte_offset = pos;
        // minimum: header + one byte of subframe data per channel + footer
        const size_t min_len = (size_t)d.header_len + d.n_channels + 2;
        const uint8_t* f = bytes + pos;
        const size_t avail = n - pos;
        size_t found = 0;
        if (avail >= min_len) {
            uint16_t crc = crc16_update(0, f, min_len - 2);  // CRC of f[0 .. i-2) while scanning i
            size_t done = min_len - 2;
            for (size_t i = min_len; i <= avail; i++) {
                const bool at_end = i == avail;
                const bool syncish = !at_end && i + 1 < avail && f[i] == 0xff && (f[i + 1] & 0xfe) == 0xf8;
                if (!at_end && !syncish) {
                    // skip ahead to the next 0xff quickly
                    const uint8_t* nx = (const uint8_t*)memchr(f + i + 1, 0xff, avail - i - 1);
                    i = nx ? (size_t)(nx - f) - 1 : avail - 1;
                    continue;
                }
                crc = crc16_update(crc, f + done, (i - 2) - done);
                done = i - 2;
                const uint16_t stored = (uint16_t)(((uint32_t)f[i - 2] << 8) | f[i - 1]);
                if (crc == stored) {
                    if (!at_end) {  // the next header must at least parse
                        clx_frame_desc nd;
                        if (clx_parse_frame_header(f + i, avail - i, &nd, flags) != CLX_OK) continue;
                    }
                    found = i;
                    break;
                }
            }
        }
        const uint64_t elems = (uint64_t)d.n_channels * d.block_size;
        d.out_offset = out_at;
        out_at += (elems + 3) & ~3ull;
        if (found) {
            d.byte_len = (uint32_t)found;
            d.flags |= CLX_FRAME_CRC16_VERIFIED;
            descs[count++] = d;
            pos += found;
        } else {
            // Boundary unknown (damaged frame or truncated stream): hand the decoder everything
            // that is left; it reports the real outcome.  Always the last descriptor returned.
            d.byte_len = (uint32_t)std::min<size_t>(avail, (size_t)1 << 28);
            descs[count++] = d;
            break;
        }
    }
    if (next_offset) *next_offset = pos;
    if (total_out_elems) *total_out_elems = out_at;
    if (stop_status) *stop_status = stop;
    return count;
}

}  // extern "C"

namespace {

// The largest a frame with this header can be (every subframe verbatim, every sample wasting nothing), with slack:
// how far a worker looks for the end of a frame whose START it is not sure of.
size_t frame_size_bound(const clx_frame_desc& d) {
    const size_t bps = d.bits_per_sample ? d.bits_per_sample : 32;
    return (size_t)d.header_len + (size_t)d.n_channels * (((size_t)d.block_size * (bps + 1) + 7) / 8 + 16) + 64;
}

// One worker of clx_demux_frames_mt: the frames that START in [lo, hi), found without knowing where the stream's
// frames begin.  Candidates are positions whose header parses (sync code, field codes, CRC-8); a candidate counts
// once a CRC-16-confirmed end lies within the size its header allows — from there on the ordinary chain runs
// (every boundary CRC-confirmed) until a frame starts at or after `hi`.
struct DemuxPart {
    uint64_t first = UINT64_MAX;       // start of the first frame of the chain (UINT64_MAX: none found)
    uint64_t end = 0;                  // where the chain stopped: the next frame's start (>= hi), or the stop position
    int stop = CLX_OK;                 // CLX_OK: ran into `hi`; else the status at `end` (CLX_EOF at a clean end)
    bool open_tail = false;            // the last descriptor has an unknown boundary (sequential semantics: the last one)
    std::vector<clx_frame_desc> descs; // out_offset not filled
};

void demux_chain(const uint8_t* bytes, size_t n, uint64_t from, uint64_t hi, uint32_t flags, DemuxPart& part) {
    uint64_t pos = from;
    clx_frame_desc buf[64];
    for (;;) {
        if (pos >= hi) { part.end = pos; part.stop = CLX_OK; return; }
        uint64_t next = pos, total = 0;
        int stop = CLX_OK;
        // frames from pos, a few at a time; stops by itself at the first position that is not a frame start
        const size_t got = clx_demux_frames(bytes, n, pos, buf, 64, &next, &total, &stop, flags);
        for (size_t i = 0; i < got; i++) {
            if (buf[i].byte_offset >= hi) { part.end = buf[i].byte_offset; part.stop = CLX_OK; return; }
            part.descs.push_back(buf[i]);
            if (!(buf[i].flags & CLX_FRAME_CRC16_VERIFIED)) {  // boundary unknown: always the last one
                part.open_tail = true;
                part.end = buf[i].byte_offset;
                part.stop = CLX_OK;
                return;
            }
        }
        if (got < 64) { part.end = next; part.stop = stop == CLX_OK ? CLX_EOF : stop; return; }
        pos = next;
    }
}

void demux_worker(const uint8_t* bytes, size_t n, uint64_t lo, uint64_t hi, bool exact_start, uint32_t flags, DemuxPart& part) {
    if (exact_start) {  // the caller vouches for `lo`: plain chain, errors and all
        part.first = lo;
        demux_chain(bytes, n, lo, hi, flags, part);
        return;
    }
    for (uint64_t c = lo; c < hi && c + 2 <= n; c++) {
        const uint8_t* nx = (const uint8_t*)memchr(bytes + c, 0xff, (size_t)(std::min<uint64_t>(hi, n - 1) - c));
        if (!nx) break;
        c = (uint64_t)(nx - bytes);
        if ((bytes[c + 1] & 0xfe) != 0xf8) continue;
        clx_frame_desc d;
        if (clx_parse_frame_header(bytes + c, n - c, &d, flags) != CLX_OK) continue;
        // confirm: one frame from c, looking no further than a frame with this header can reach
        const size_t window = std::min<size_t>(n - c, frame_size_bound(d) + 16);
        clx_frame_desc one;
        uint64_t next = c, total = 0;
        int stop = CLX_OK;
        if (clx_demux_frames(bytes + c, window, 0, &one, 1, &next, &total, &stop, flags) != 1) continue;
        if (!(one.flags & CLX_FRAME_CRC16_VERIFIED)) continue;
        if (window < n - c && one.byte_len == window) continue;  // "confirmed" only by the window's artificial end
        part.first = c;
        demux_chain(bytes, n, c, hi, flags, part);
        return;
    }
}

}  // namespace

extern "C" {

// clx_demux_frames on `n_threads` host threads: the byte range is cut into equal parts, every worker finds the
// frames that start in its part (see demux_worker) and the parts are stitched in order — a part is accepted only
// if the chain before it ends exactly where it begins; otherwise (a false start that survived CRC-8 and CRC-16:
// a 2^-24 event, or damage) the previous chain simply carries on through it.  Same results as clx_demux_frames,
// descriptor for descriptor.
size_t clx_demux_frames_mt(const uint8_t* bytes, size_t n, uint64_t start, clx_frame_desc* descs, size_t max_frames,
                           uint64_t* next_offset, uint64_t* total_out_elems, int* stop_status, uint32_t flags,
                           uint32_t n_threads) {
    const size_t min_part = 1u << 16;
    if (start > n) n_threads = 1;
    size_t parts = n_threads ? n_threads : std::max(1u, std::thread::hardware_concurrency());
    if (start <= n) parts = std::min<size_t>(parts, std::max<size_t>(1, (n - start) / min_part));
    if (parts <= 1) return clx_demux_frames(bytes, n, start, descs, max_frames, next_offset, total_out_elems, stop_status, flags);
    std::vector<DemuxPart> part(parts);
    std::vector<uint64_t> cut(parts + 1);
    for (size_t t = 0; t <= parts; t++) cut[t] = start + (uint64_t)((n - start) * t / parts);
    {
        // No exception may cross the C ABI: a thread that cannot be started (resource limits) just leaves its part to
        // the stitching pass, which then walks through it sequentially; an allocation failure anywhere falls back to
        // the sequential routine altogether.
        std::vector<std::thread> th;
        bool failed = false;
        try {
            th.reserve(parts);
            for (size_t t = 1; t < parts; t++) {
                try {
                    th.emplace_back([&, t] {
                        try { demux_worker(bytes, n, cut[t], cut[t + 1], false, flags, part[t]); }
                        catch (...) { part[t] = DemuxPart(); }
                    });
                } catch (...) { break; }
            }
            demux_worker(bytes, n, cut[0], cut[1], true, flags, part[0]);
        } catch (...) { failed = true; }
        for (auto& x : th) x.join();
        if (failed) return clx_demux_frames(bytes, n, start, descs, max_frames, next_offset, total_out_elems, stop_status, flags);
    }
    size_t count = 0;
    uint64_t out_at = total_out_elems ? *total_out_elems : 0, pos = start;
    int stop = CLX_OK;
    bool done = false;
    try {
    auto emit = [&](const clx_frame_desc& src) {
        clx_frame_desc d = src;
        const uint64_t elems = (uint64_t)d.n_channels * d.block_size;
        d.out_offset = out_at;
        out_at += (elems + 3) & ~3ull;
        descs[count++] = d;
    };
    for (size_t t = 0; t < parts && !done; t++) {
        DemuxPart* p = &part[t];
        DemuxPart redo;
        if (t > 0 && pos >= cut[t + 1]) continue;  // the chain so far already reaches past this part
        if (t > 0 && p->first != pos) {
            // the part does not begin where the chain ends: let the chain carry on through it
            redo.first = pos;
            demux_chain(bytes, n, pos, cut[t + 1], flags, redo);
            p = &redo;
        }
        for (const clx_frame_desc& d : p->descs) {
            if (count == max_frames) { done = true; break; }
            emit(d);
            pos = d.byte_offset + d.byte_len;
            if (!(d.flags & CLX_FRAME_CRC16_VERIFIED)) { pos = d.byte_offset; done = true; break; }  // unknown boundary: the last one
        }
        if (done) break;
        if (count == max_frames) { done = true; break; }  // (the sequential routine does not look past its last descriptor either)
        pos = p->end;
        if (p->stop != CLX_OK) { stop = p->stop; done = true; }
    }
    if (!done && count < max_frames) {
        // The parts are used up and nothing said stop (a chain that ran into its part's end exactly at the end of
        // the stream, for one): whatever is left, usually nothing, goes the sequential way, which also names the
        // status at the position where it ends.
        uint64_t total = out_at;
        count += clx_demux_frames(bytes, n, pos, descs + count, max_frames - count, &pos, &total, &stop, flags);
        out_at = total;
    }
    } catch (...) {  // (an allocation failed while stitching: *total_out_elems is still untouched)
        return clx_demux_frames(bytes, n, start, descs, max_frames, next_offset, total_out_elems, stop_status, flags);
    }
    if (next_offset) *next_offset = pos;
    if (total_out_elems) *total_out_elems = out_at;
    if (stop_status) *stop_status = stop;
    return count;
}

}  // extern "C"
