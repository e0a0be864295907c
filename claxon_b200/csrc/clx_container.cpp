// clx_container.cpp — container feeds: FLAC frames wrapped in Ogg or in MP4 / ISO BMFF come
// with their boundaries, so the descriptor table the device path wants is read off the container instead of
// being searched for (clx_demux_frames).  What the reference's examples do with the `ogg` and `mp4parse` crates
// (examples/decode_ogg.rs:26-125, examples/decode_mp4.rs:26-167: every Ogg packet after the header packets is
// one frame; every MP4 sample is one frame, located by the chunk offset / sample-to-chunk / sample size tables),
// restated for in-memory files.  Host only.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "claxon_b200.h"

namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t be64(const uint8_t* p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }
uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// STREAMINFO body (34 bytes) -> clx_streaminfo, field packing as src/metadata.rs:321-400
void parse_streaminfo(const uint8_t* b, clx_streaminfo* si) {
    memset(si, 0, sizeof *si);
    si->min_block_size = ((uint32_t)b[0] << 8) | b[1];
    si->max_block_size = ((uint32_t)b[2] << 8) | b[3];
    si->min_frame_size = ((uint32_t)b[4] << 16) | ((uint32_t)b[5] << 8) | b[6];
    si->max_frame_size = ((uint32_t)b[7] << 16) | ((uint32_t)b[8] << 8) | b[9];
    si->sample_rate = ((uint32_t)b[10] << 12) | ((uint32_t)b[11] << 4) | (b[12] >> 4);
    si->channels = ((b[12] >> 1) & 7u) + 1;
    si->bits_per_sample = ((((uint32_t)b[12] & 1u) << 4) | (b[13] >> 4)) + 1;
    si->samples = ((uint64_t)(b[13] & 15) << 32) | ((uint64_t)b[14] << 24) | ((uint64_t)b[15] << 16) | ((uint64_t)b[16] << 8) | b[17];
    memcpy(si->md5sum, b + 18, 16);
}

// Describes the frame at p[0..len) (its exact extent is known), placing it at byte_offset / out_at.
int describe(const uint8_t* p, size_t len, uint64_t byte_offset, uint64_t* out_at, clx_frame_desc* d, uint32_t flags) {
    const int st = clx_parse_frame_header(p, len, d, flags);
    if (st != CLX_OK) return st == CLX_EOF ? (int)CLX_ERR_IO_UNEXPECTED_EOF : st;
    d->byte_offset = byte_offset;
    d->byte_len = (uint32_t)len;
    d->out_offset = *out_at;
    *out_at += ((uint64_t)d->n_channels * d->block_size + 3) & ~3ull;
    return CLX_OK;
}

// Ogg page checksum: CRC-32, polynomial 0x04c11db7, initial value 0, MSB first, over the page with the CRC field zero.
uint32_t ogg_crc(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t r = i << 24;
            for (int k = 0; k < 8; k++) r = (r & 0x80000000u) ? (r << 1) ^ 0x04c11db7u : r << 1;
            table[i] = r;
        }
        ready = true;
    }
    uint32_t c = 0;
    for (size_t i = 0; i < n; i++) c = (c << 8) ^ table[((c >> 24) ^ (i >= 22 && i < 26 ? 0 : p[i])) & 0xff];
    return c;
}

}  // namespace

extern "C" {

int clx_ogg_frames(const uint8_t* ogg, size_t n, clx_streaminfo* si, uint8_t* frames_out, size_t frames_cap,
                   clx_frame_desc* descs, size_t max_frames, size_t* n_frames, size_t* frames_bytes,
                   uint64_t* total_out_elems, uint32_t flags) {
    if (!ogg || !si || !n_frames || !frames_bytes || (!descs && max_frames) || (!frames_out && frames_cap)) return CLX_ERR_INVALID_ARGUMENT;
    *n_frames = 0;
    *frames_bytes = 0;
    if (total_out_elems) *total_out_elems = 0;
    std::vector<uint8_t> packet;   // the packet being assembled (packets may span pages)
    size_t at = 0, packets_seen = 0, headers_left = 0, used = 0;
    uint64_t out_at = 0;
    uint32_t serial = 0;
    bool have_serial = false, open = false;
    while (at < n) {
        if (n - at < 27 || memcmp(ogg + at, "OggS", 4) != 0 || ogg[at + 4] != 0) return CLX_ERR_CONTAINER;
        const uint32_t nseg = ogg[at + 26];
        if (n - at < 27 + nseg) return CLX_ERR_CONTAINER;
        size_t body = 0;
        for (uint32_t i = 0; i < nseg; i++) body += ogg[at + 27 + i];
        const size_t page = 27 + nseg + body;
        if (n - at < page) return CLX_ERR_CONTAINER;
        if (!(flags & CLX_OPT_NO_VERIFY_CRC) && ogg_crc(ogg + at, page) != le32(ogg + at + 22)) return CLX_ERR_CONTAINER;
        const uint32_t this_serial = le32(ogg + at + 14);
        if (!have_serial) { serial = this_serial; have_serial = true; }
        if (this_serial == serial) {  // other logical streams multiplexed into the file are not ours
            const bool continued = (ogg[at + 5] & 1u) != 0;
            if (!continued && open) return CLX_ERR_CONTAINER;  // the previous page left a packet unfinished
            if (continued && !open) packet.clear();              // continuation of a packet we never saw the start of
            const uint8_t* data = ogg + at + 27 + nseg;
            bool skip = continued && !open;
            for (uint32_t i = 0; i < nseg; i++) {
                const uint32_t lace = ogg[at + 27 + i];
                if (!skip) packet.insert(packet.end(), data, data + lace);
                data += lace;
                open = true;
                if (lace < 255) {  // the packet ends here
                    open = false;
                    if (skip) { skip = false; packet.clear(); continue; }
                    if (packets_seen == 0) {
                        // 0x7f "FLAC" major minor, u16 header packets to follow, "fLaC", STREAMINFO block
                        if (packet.size() < 13 + 4 + 34 || packet[0] != 0x7f || memcmp(&packet[1], "FLAC", 4) != 0 ||
                            memcmp(&packet[9], "fLaC", 4) != 0 || (packet[13] & 0x7f) != 0)
                            return CLX_ERR_CONTAINER;
                        headers_left = ((size_t)packet[7] << 8) | packet[8];
                        parse_streaminfo(&packet[17], si);
                    } else if (headers_left > 0) {
                        headers_left--;  // a metadata block (Vorbis comment, ...): not needed to decode
                    } else if (!packet.empty()) {  // empty packets do occur (the reference skips them too)
                        if (*n_frames >= max_frames || used + packet.size() > frames_cap) return CLX_ERR_INVALID_ARGUMENT;
                        memcpy(frames_out + used, packet.data(), packet.size());
                        const int st = describe(frames_out + used, packet.size(), used, &out_at, &descs[*n_frames], flags);
                        if (st != CLX_OK) return st;
                        used += packet.size();
                        ++*n_frames;
                    }
                    packets_seen++;
                    packet.clear();
                }
            }
        }
        at += page;
    }
    if (packets_seen == 0) return CLX_ERR_CONTAINER;
    *frames_bytes = used;
    if (total_out_elems) *total_out_elems = out_at;
    return CLX_OK;
}

namespace {
struct Box { uint32_t type; const uint8_t* body; size_t len; };
// Iterates the boxes in [p, p + n); false on a malformed size.
bool next_box(const uint8_t*& p, size_t& n, Box* b) {
    if (n < 8) return false;
    uint64_t size = be32(p);
    size_t hdr = 8;
    if (size == 1) { if (n < 16) return false; size = be64(p + 8); hdr = 16; }
    else if (size == 0) size = n;  // to the end of the enclosing box
    if (size < hdr || size > n) return false;
    b->type = be32(p + 4);
    b->body = p + hdr;
    b->len = (size_t)size - hdr;
    p += size;
    n -= (size_t)size;
    return true;
}
constexpr uint32_t fourcc(char a, char b, char c, char d) { return ((uint32_t)a << 24) | ((uint32_t)b << 16) | ((uint32_t)c << 8) | (uint32_t)d; }

struct Track {
    bool flac = false, have_si = false;
    clx_streaminfo si{};
    const uint8_t *stsz = nullptr, *stsc = nullptr, *stco = nullptr, *co64 = nullptr;
    size_t stsz_len = 0, stsc_len = 0, stco_len = 0, co64_len = 0;
};

void walk(const uint8_t* p, size_t n, Track* t, int depth) {
    Box b;
    while (n >= 8 && next_box(p, n, &b)) {
        switch (b.type) {
        case fourcc('m', 'd', 'i', 'a'): case fourcc('m', 'i', 'n', 'f'): case fourcc('s', 't', 'b', 'l'):
            if (depth < 8) walk(b.body, b.len, t, depth + 1);
            break;
        case fourcc('s', 't', 's', 'd'): {
            if (b.len < 8) break;
            const uint8_t* q = b.body + 8;  // version/flags, entry count
            size_t m = b.len - 8;
            Box e;
            if (next_box(q, m, &e) && e.type == fourcc('f', 'L', 'a', 'C') && e.len >= 28) {
                t->flac = true;
                const uint8_t* c = e.body + 28;  // past the AudioSampleEntry fields
                size_t cl = e.len - 28;
                Box d;
                while (cl >= 8 && next_box(c, cl, &d))
                    if (d.type == fourcc('d', 'f', 'L', 'a') && d.len >= 4) {
                        const uint8_t* mb = d.body + 4;  // version/flags, then FLAC metadata blocks
                        size_t ml = d.len - 4;
                        while (ml >= 4) {
                            const uint32_t type = mb[0] & 0x7f, len = ((uint32_t)mb[1] << 16) | ((uint32_t)mb[2] << 8) | mb[3];
                            if (len > ml - 4) break;
                            if (type == 0 && len == 34) { parse_streaminfo(mb + 4, &t->si); t->have_si = true; }
                            if (mb[0] & 0x80) break;
                            mb += 4 + len;
                            ml -= 4 + len;
                        }
                    }
            }
            break;
        }
        case fourcc('s', 't', 's', 'z'): t->stsz = b.body; t->stsz_len = b.len; break;
        case fourcc('s', 't', 's', 'c'): t->stsc = b.body; t->stsc_len = b.len; break;
        case fourcc('s', 't', 'c', 'o'): t->stco = b.body; t->stco_len = b.len; break;
        case fourcc('c', 'o', '6', '4'): t->co64 = b.body; t->co64_len = b.len; break;
        default: break;
        }
    }
}
}  // namespace

int clx_mp4_frames(const uint8_t* mp4, size_t n, clx_streaminfo* si, clx_frame_desc* descs, size_t max_frames,
                   size_t* n_frames, uint64_t* total_out_elems, uint32_t flags) {
    if (!mp4 || !si || !n_frames || (!descs && max_frames)) return CLX_ERR_INVALID_ARGUMENT;
    *n_frames = 0;
    if (total_out_elems) *total_out_elems = 0;
    // the first track with a 'fLaC' sample entry (as the reference's example: one track per output)
    Track tr;
    bool found = false;
    const uint8_t* p = mp4;
    size_t left = n;
    Box b;
    while (!found && left >= 8 && next_box(p, left, &b)) {
        if (b.type != fourcc('m', 'o', 'o', 'v')) continue;
        const uint8_t* q = b.body;
        size_t m = b.len;
        Box t;
        while (!found && m >= 8 && next_box(q, m, &t)) {
            if (t.type != fourcc('t', 'r', 'a', 'k')) continue;
            Track cand;
            walk(t.body, t.len, &cand, 0);
            if (cand.flac) { tr = cand; found = true; }
        }
    }
    if (!found || !tr.have_si || !tr.stsz || !tr.stsc || (!tr.stco && !tr.co64) || tr.stsz_len < 12 || tr.stsc_len < 8) return CLX_ERR_CONTAINER;
    *si = tr.si;
    const uint32_t fixed = be32(tr.stsz + 4), count = be32(tr.stsz + 8);
    if (fixed == 0 && tr.stsz_len < 12 + 4ull * count) return CLX_ERR_CONTAINER;
    const uint32_t n_runs = be32(tr.stsc + 4);
    if (tr.stsc_len < 8 + 12ull * n_runs) return CLX_ERR_CONTAINER;
    const bool wide = tr.co64 != nullptr;
    const uint8_t* co = wide ? tr.co64 : tr.stco;
    const size_t co_len = wide ? tr.co64_len : tr.stco_len;
    if (co_len < 8) return CLX_ERR_CONTAINER;
    const uint32_t n_chunks = be32(co + 4);
    if (co_len < 8 + (wide ? 8ull : 4ull) * n_chunks) return CLX_ERR_CONTAINER;
    uint64_t out_at = 0;
    uint32_t sample = 0, run = 0, per_chunk = 0;
    for (uint32_t c = 0; c < n_chunks && sample < count; c++) {
        // "first_chunk" is 1-based; a run holds until the next run's first chunk (examples/decode_mp4.rs:83-93)
        while (run < n_runs && be32(tr.stsc + 8 + 12ull * run) == c + 1) { per_chunk = be32(tr.stsc + 8 + 12ull * run + 4); run++; }
        uint64_t off = wide ? be64(co + 8 + 8ull * c) : be32(co + 8 + 4ull * c);
        for (uint32_t k = 0; k < per_chunk && sample < count; k++, sample++) {
            const uint32_t size = fixed ? fixed : be32(tr.stsz + 12 + 4ull * sample);
            if (off > n || size > n - off) return CLX_ERR_CONTAINER;
            if (*n_frames >= max_frames) return CLX_ERR_INVALID_ARGUMENT;
            const int st = describe(mp4 + off, size, off, &out_at, &descs[*n_frames], flags);
            if (st != CLX_OK) return st;
            ++*n_frames;
            off += size;
        }
    }
    if (total_out_elems) *total_out_elems = out_at;
    return CLX_OK;
}

}  // extern "C"
