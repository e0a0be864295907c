/* claxon_oracle.c — TEST INFRASTRUCTURE ONLY (see claxon_oracle.h).
 *
 * A from-the-spec CPU restatement of claxon v0.4.3's frame decode path.  It is
 * written around a positional bit cursor over a byte span rather than claxon's
 * one-byte cache (`Bitstream{data,bits_left}`): only the observable semantics of
 * src/input.rs:415-643 are kept (MSB-first fields; a read fails with
 * UnexpectedEof exactly when it needs a bit past the end of the input).
 *
 * PARITY PINNED: against every in-source known-answer vector of the reference's unit tests on
 * this path (tests/golden/kat.json, tests/test_oracle_golden.py), against the STREAMINFO MD5 of
 * pop.flac / short.flac / wasted_bits.flac and the two Vorbis-comment variants
 * (tests/golden/fixtures.npz), against the reference's fuzz corpus statuses, and — for the shapes
 * no fixture reaches — frame by frame against a second, independent restatement in plain Python
 * (tests/spec_decode.py, tests/spec_header.py, tests/spec_metadata.py).  The reference itself is
 * Rust; neither the authoring image nor the GPU box has rustc / cargo, so there is no oracle/_ref.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may call this code.
 */
#include "claxon_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* bit cursor                                                                 */
/* ------------------------------------------------------------------------- */

typedef struct {
    const uint8_t* p;
    uint64_t nbits; /* 8 * number of bytes available */
    uint64_t pos;   /* next unread bit, 0 = MSB of p[0] */
    int err;        /* sticky: CLX_ERR_IO_UNEXPECTED_EOF once a read ran off the end */
} bitcur;

/* 64 bits starting at byte `byte_at`, big-endian, zero-filled past the end. */
static inline uint64_t bc_load64(const bitcur* b, uint64_t byte_at) {
    uint64_t nbytes = b->nbits >> 3;
    if (byte_at + 8 <= nbytes) {
        uint64_t w;
        memcpy(&w, b->p + byte_at, 8);
        return __builtin_bswap64(w);
    }
    uint64_t w = 0;
    for (uint64_t i = 0; i < 8; i++) {
        w <<= 8;
        if (byte_at + i < nbytes) w |= b->p[byte_at + i];
    }
    return w;
}

/* Reads `n` (0..32) bits MSB-first.  src/input.rs:515-643 (read_leq_u8/u16/u32,
 * read_gt_u8_leq_u16): all are plain big-endian bit-field extraction. */
static inline uint32_t bc_read(bitcur* b, uint32_t n) {
    if (n == 0) return 0; /* read_leq_u8(0) consumes nothing (src/input.rs:521, :706) */
    if (b->err) return 0;
    if (b->pos + n > b->nbits) { b->err = CLX_ERR_IO_UNEXPECTED_EOF; return 0; }
    uint64_t w = bc_load64(b, b->pos >> 3) << (b->pos & 7); /* (pos&7) + n <= 39 < 64 */
    b->pos += n;
    return (uint32_t)(w >> (64 - n));
}

/* Counts zero bits up to the next one bit and consumes that one bit.
 * src/input.rs:475-511 (read_unary). */
static inline uint32_t bc_unary(bitcur* b) {
    if (b->err) return 0;
    uint32_t n = 0;
    for (;;) {
        if (b->pos >= b->nbits) { b->err = CLX_ERR_IO_UNEXPECTED_EOF; return 0; }
        uint32_t skew = (uint32_t)(b->pos & 7);
        uint64_t w = bc_load64(b, b->pos >> 3) << skew;
        uint64_t valid = 64 - skew;
        if (valid > b->nbits - b->pos) valid = b->nbits - b->pos;
        if (w != 0) {
            uint32_t z = (uint32_t)__builtin_clzll(w);
            if (z < valid) { b->pos += z + 1; return n + z; }
        }
        n += (uint32_t)valid;
        b->pos += valid;
    }
}

int clxo_bit_read(const uint8_t* p, size_t n, uint64_t* bitpos, int kind, uint32_t bits,
                  uint32_t* value) {
    bitcur b = {p, (uint64_t)n * 8, *bitpos, 0};
    uint32_t v = kind == 1 ? bc_unary(&b) : bc_read(&b, bits);
    if (b.err) return b.err;
    *bitpos = b.pos;
    *value = v;
    return CLX_OK;
}

/* ------------------------------------------------------------------------- */
/* scalar helpers                                                             */
/* ------------------------------------------------------------------------- */

/* src/subframe.rs:96-101 */
int16_t clxo_extend_sign_u16(uint16_t val, uint32_t bits) {
    uint16_t up = (uint16_t)(val << (16 - bits));
    return (int16_t)((int16_t)up >> (16 - bits));
}

/* src/subframe.rs:117-122 */
int32_t clxo_extend_sign_u32(uint32_t val, uint32_t bits) {
    uint32_t up = val << (32 - bits);
    return ((int32_t)up) >> (32 - bits);
}

/* src/subframe.rs:157-170: 0,-1,1,-2,2,... */
int32_t clxo_rice_to_signed(uint32_t val) {
    uint32_t half = val >> 1;
    uint32_t mask = 0u - (val & 1u);
    return (int32_t)(half ^ mask);
}

/* src/crc.rs:13-31 + :90-92.  The table claxon embeds is the MSB-first table of
 * polynomial x^8+x^2+x+1 (0x07), init 0; regenerate it rather than embedding. */
static uint8_t g_crc8_tab[256];
static uint16_t g_crc16_tab[256];
static pthread_once_t g_tab_once = PTHREAD_ONCE_INIT;

static void build_tables(void) {
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        for (int k = 0; k < 8; k++) c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
        g_crc8_tab[i] = c;
        /* src/crc.rs:33-57: polynomial 0x8005, MSB-first, init 0. */
        uint16_t d = (uint16_t)(i << 8);
        for (int k = 0; k < 8; k++) d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1));
        g_crc16_tab[i] = d;
    }
}

uint8_t clxo_crc8(const uint8_t* p, size_t n) {
    pthread_once(&g_tab_once, build_tables);
    uint8_t s = 0;
    for (size_t i = 0; i < n; i++) s = g_crc8_tab[s ^ p[i]]; /* src/crc.rs:90-92 */
    return s;
}

uint16_t clxo_crc16(const uint8_t* p, size_t n) {
    pthread_once(&g_tab_once, build_tables);
    uint16_t s = 0;
    for (size_t i = 0; i < n; i++) /* src/crc.rs:110-112 */
        s = (uint16_t)((s << 8) ^ g_crc16_tab[(uint8_t)(s >> 8) ^ p[i]]);
    return s;
}

/* ------------------------------------------------------------------------- */
/* predictors                                                                 */
/* ------------------------------------------------------------------------- */

/* src/subframe.rs:417-474.  All arithmetic is Wrapping<i32>; done here in u32. */
void clxo_predict_fixed(uint32_t order, int32_t* buf, size_t n) {
    static const int32_t rows[5][4] = {
        {0, 0, 0, 0}, {1, 0, 0, 0}, {-1, 2, 0, 0}, {1, -3, 3, 0}, {-1, 4, -6, 4}};
    if (order > 4 || n < order) return;
    for (size_t i = order; i < n; i++) {
        uint32_t pred = 0;
        for (uint32_t j = 0; j < order; j++)
            pred += (uint32_t)rows[order][j] * (uint32_t)buf[i - order + j];
        buf[i] = (int32_t)(pred + (uint32_t)buf[i]);
    }
}

/* src/subframe.rs:524-614.  coefs[j] multiplies buf[i-order+j] (claxon's stored,
 * i.e. reversed-from-stream, order).  i64 products and sum, arithmetic >> shift,
 * + residual in i64, truncate to i32.  The low-order variant's zero padding to 12
 * taps (:543-551, :575-582) is numerically the same `order`-tap recurrence. */
static inline __attribute__((always_inline)) void lpc_run(const int16_t* coefs, uint32_t order,
                                                          uint32_t shift, int32_t* buf, size_t n) {
    for (size_t i = order; i < n; i++) {
        int64_t sum = 0;
        for (uint32_t j = 0; j < order; j++)
            sum += (int64_t)coefs[j] * (int64_t)buf[i - order + j];
        int64_t pred = sum >> shift;
        buf[i] = (int32_t)(uint32_t)(uint64_t)(pred + (int64_t)buf[i]);
    }
}

void clxo_predict_lpc(const int16_t* coefs, uint32_t order, uint32_t shift, int32_t* buf,
                      size_t n) {
    switch (order) { /* constant trip counts let the compiler unroll the tap loop */
#define CASE(o) case o: lpc_run(coefs, o, shift, buf, n); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11)
    CASE(12)
#undef CASE
    default: lpc_run(coefs, order, shift, buf, n); break;
    }
}

/* src/frame.rs:319-334 */
void clxo_decode_left_side(int32_t* buf, size_t n_total) {
    size_t bs = n_total / 2;
    for (size_t i = 0; i < bs; i++)
        buf[bs + i] = (int32_t)((uint32_t)buf[i] - (uint32_t)buf[bs + i]);
}

/* src/frame.rs:345-360 */
void clxo_decode_right_side(int32_t* buf, size_t n_total) {
    size_t bs = n_total / 2;
    for (size_t i = 0; i < bs; i++)
        buf[i] = (int32_t)((uint32_t)buf[i] + (uint32_t)buf[bs + i]);
}

/* src/frame.rs:371-389.  `/ 2` is Rust's truncating signed division. */
void clxo_decode_mid_side(int32_t* buf, size_t n_total) {
    size_t bs = n_total / 2;
    for (size_t i = 0; i < bs; i++) {
        int32_t side = buf[bs + i];
        int32_t mid = (int32_t)(((uint32_t)buf[i] * 2u) | ((uint32_t)side & 1u));
        int32_t sum = (int32_t)((uint32_t)mid + (uint32_t)side);
        int32_t dif = (int32_t)((uint32_t)mid - (uint32_t)side);
        buf[i] = sum / 2;
        buf[bs + i] = dif / 2;
    }
}

/* ------------------------------------------------------------------------- */
/* subframe                                                                   */
/* ------------------------------------------------------------------------- */

/* src/subframe.rs:397-415 (also the warm-up reads at :504, :667). */
static void read_verbatim(bitcur* b, uint32_t bps, int32_t* out, size_t n) {
    for (size_t i = 0; i < n && !b->err; i++) out[i] = clxo_extend_sign_u32(bc_read(b, bps), bps);
}

/* src/subframe.rs:236-380.  `out` has block_size - n_warm_up slots. */
static int read_residual(bitcur* b, uint32_t block_size, uint32_t n_warm_up, int32_t* out) {
    uint32_t method = bc_read(b, 2);
    if (b->err) return b->err;
    if (method > 1) return CLX_ERR_RESIDUAL_RESERVED; /* :245 */
    uint32_t porder = bc_read(b, 4);
    if (b->err) return b->err;
    uint32_t n_part = 1u << porder;
    uint32_t per = block_size >> porder;
    if ((block_size & ((n_part - 1u) & 0xffffu)) != 0) return CLX_ERR_PARTITION_ORDER_INVALID; /* :262 */
    if (n_warm_up > per) return CLX_ERR_RESIDUAL_INVALID;                                    /* :275 */
    uint32_t param_bits = method == 0 ? 4u : 5u; /* :314 vs :362 */
    uint32_t escape = method == 0 ? 15u : 31u;
    size_t at = 0;
    uint32_t len = per - n_warm_up;
    for (uint32_t part = 0; part < n_part; part++) {
        uint32_t k = bc_read(b, param_bits);
        if (b->err) return b->err;
        if (k == escape) return CLX_ERR_UNENCODED_BINARY; /* :317, :365 */
        uint64_t nbytes = b->nbits >> 3;
        for (uint32_t i = 0; i < len; i++) {
            uint64_t byte_at = b->pos >> 3;
            if (byte_at + 8 <= nbytes) { /* whole code inside one 64-bit window: common case */
                uint32_t skew = (uint32_t)(b->pos & 7);
                uint64_t w;
                memcpy(&w, b->p + byte_at, 8);
                w = __builtin_bswap64(w) << skew;
                if (w != 0) {
                    uint32_t z = (uint32_t)__builtin_clzll(w);
                    if (z + 1 + k + skew <= 64) {
                        uint32_t r = k ? (uint32_t)((w << (z + 1)) >> (64 - k)) : 0;
                        out[at + i] = clxo_rice_to_signed((z << k) | r);
                        b->pos += z + 1 + k;
                        continue;
                    }
                }
            }
            uint32_t q = bc_unary(b);
            if (b->err) return b->err;
            uint32_t r = bc_read(b, k);
            if (b->err) return b->err;
            out[at + i] = clxo_rice_to_signed((q << k) | r); /* :340, :346, :376 */
        }
        at += len;
        len = per;
    }
    return CLX_OK;
}

/* src/subframe.rs:184-228 with :29-91, :382-415, :492-516, :651-721 inlined. */
int clxo_decode_subframe(const uint8_t* p, size_t n, uint64_t* bit_pos, uint32_t bps,
                         int32_t* out, uint32_t block_size) {
    bitcur b = {p, (uint64_t)n * 8, *bit_pos, 0};

    /* header: :29-91 */
    uint32_t pad = bc_read(&b, 1);
    if (b.err) return b.err;
    if (pad) return CLX_ERR_SUBFRAME_HEADER_INVALID;
    uint32_t code = bc_read(&b, 6);
    if (b.err) return b.err;
    enum { T_CONST, T_VERB, T_FIXED, T_LPC } type;
    uint32_t order = 0;
    if (code == 0) type = T_CONST;
    else if (code == 1) type = T_VERB;
    else if ((code & 0x3e) == 0x02 || (code & 0x3c) == 0x04 || (code & 0x30) == 0x10)
        return CLX_ERR_SUBFRAME_HEADER_RESERVED;
    else if ((code & 0x38) == 0x08) {
        order = code & 7;
        if (order > 4) return CLX_ERR_SUBFRAME_HEADER_RESERVED;
        type = T_FIXED;
    } else {
        order = (code & 0x1f) + 1;
        type = T_LPC;
    }
    uint32_t wasted = 0;
    uint32_t has_wasted = bc_read(&b, 1);
    if (b.err) return b.err;
    if (has_wasted) {
        wasted = 1 + bc_unary(&b);
        if (b.err) return b.err;
    }
    if (wasted > 31) return CLX_ERR_WASTED_BITS_GT_31;        /* :82 */
    if (wasted >= bps) return CLX_ERR_NO_NON_WASTED_BITS;     /* :198 */
    uint32_t sf_bps = bps - wasted;

    switch (type) {
    case T_CONST: { /* :382-394 */
        int32_t v = clxo_extend_sign_u32(bc_read(&b, sf_bps), sf_bps);
        if (b.err) return b.err;
        for (uint32_t i = 0; i < block_size; i++) out[i] = v;
        break;
    }
    case T_VERB: /* :397-415 */
        read_verbatim(&b, sf_bps, out, block_size);
        if (b.err) return b.err;
        break;
    case T_FIXED: { /* :492-516 */
        if (block_size < order) return CLX_ERR_FIXED_ORDER_GT_BLOCK;
        read_verbatim(&b, sf_bps, out, order);
        if (b.err) return b.err;
        int st = read_residual(&b, block_size, order, out + order);
        if (st) return st;
        clxo_predict_fixed(order, out, block_size);
        break;
    }
    case T_LPC: { /* :651-721 */
        if (block_size < order) return CLX_ERR_LPC_ORDER_GT_BLOCK;
        read_verbatim(&b, sf_bps, out, order);
        if (b.err) return b.err;
        uint32_t prec_m1 = bc_read(&b, 4);
        if (b.err) return b.err;
        if (prec_m1 == 15) return CLX_ERR_QLP_PRECISION_INVALID;
        uint32_t precision = prec_m1 + 1;
        int16_t shift = clxo_extend_sign_u16((uint16_t)bc_read(&b, 5), 5);
        if (b.err) return b.err;
        if (shift < 0) return CLX_ERR_NEGATIVE_QLP_SHIFT;
        int16_t coefs[32];
        /* :696-701: first coefficient in the stream applies to s[i-1]. */
        for (uint32_t t = 0; t < order; t++) {
            uint16_t raw = (uint16_t)bc_read(&b, precision);
            if (b.err) return b.err;
            coefs[order - 1 - t] = clxo_extend_sign_u16(raw, precision);
        }
        int st = read_residual(&b, block_size, order, out + order);
        if (st) return st;
        clxo_predict_lpc(coefs, order, (uint32_t)shift, out, block_size);
        break;
    }
    }

    if (wasted > 0) /* :216-225 */
        for (uint32_t i = 0; i < block_size; i++) out[i] = (int32_t)((uint32_t)out[i] << wasted);

    *bit_pos = b.pos;
    return CLX_OK;
}

/* ------------------------------------------------------------------------- */
/* frame                                                                      */
/* ------------------------------------------------------------------------- */

/* src/frame.rs:64-105 */
int clxo_read_var_length_int(const uint8_t* p, size_t n, uint64_t* value, size_t* consumed) {
    if (n < 1) return CLX_ERR_IO_UNEXPECTED_EOF;
    uint8_t first = p[0];
    uint32_t ones = 0;
    while (ones < 8 && (first & (0x80u >> ones))) ones++;
    uint32_t extra = 0;
    if (ones == 1) return CLX_ERR_VARINT_INVALID; /* :81 */
    if (ones >= 2) extra = ones - 1;
    uint8_t data_mask = (uint8_t)(0x7fu >> ones); /* 0 when ones == 8 */
    if (ones == 0) data_mask = 0x7f;
    uint64_t result = ((uint64_t)(first & data_mask)) << (6 * extra);
    for (uint32_t i = 0; i < extra; i++) {
        if (1 + i >= n) return CLX_ERR_IO_UNEXPECTED_EOF;
        uint8_t byte = p[1 + i];
        if ((byte & 0xc0) != 0x80) return CLX_ERR_VARINT_INVALID; /* :97 */
        result |= ((uint64_t)(byte & 0x3f)) << (6 * (extra - 1 - i));
    }
    *value = result;
    *consumed = 1 + extra;
    return CLX_OK;
}

/* src/frame.rs:131-316 */
int clxo_read_frame_header(const uint8_t* p, size_t n, clxo_frame_header* h) {
    memset(h, 0, sizeof *h);
    if (n < 2) return CLX_EOF; /* :140-143 incl. the one-byte-left case (input.rs:94-101) */
    uint32_t sync = ((uint32_t)p[0] << 8) | p[1];
    if ((sync & 0xfffc) != 0xfff8) return CLX_ERR_SYNC_MISSING;
    if (sync & 2) return CLX_ERR_FRAME_HEADER_RESERVED;
    h->variable_blocking = sync & 1;
    size_t at = 2;
#define NEED(k) do { if (at + (k) > n) return CLX_ERR_IO_UNEXPECTED_EOF; } while (0)
    NEED(1);
    uint8_t bs_sr = p[at++];
    uint32_t bs_code = bs_sr >> 4, sr_code = bs_sr & 15;
    int bs8 = 0, bs16 = 0;
    if (bs_code == 0) return CLX_ERR_FRAME_HEADER_RESERVED;
    else if (bs_code == 1) h->block_size = 192;
    else if (bs_code <= 5) h->block_size = 576u << (bs_code - 2);
    else if (bs_code == 6) bs8 = 1;
    else if (bs_code == 7) bs16 = 1;
    else h->block_size = 256u << (bs_code - 8);
    static const uint32_t rates[12] = {0, 88200, 176400, 192000, 8000, 16000,
                                       22050, 24000, 32000, 44100, 48000, 96000};
    int sr8 = 0, sr16 = 0, sr16x10 = 0;
    if (sr_code < 12) h->sample_rate = rates[sr_code];
    else if (sr_code == 12) sr8 = 1;
    else if (sr_code == 13) sr16 = 1;
    else if (sr_code == 14) sr16x10 = 1;
    else return CLX_ERR_FRAME_HEADER_INVALID; /* :210 */
    NEED(1);
    uint8_t cbr = p[at++];
    uint32_t ch = cbr >> 4;
    if (ch < 8) h->n_channels = ch + 1;
    else if (ch <= 10) h->n_channels = 2;
    else return CLX_ERR_FRAME_HEADER_RESERVED;
    h->channel_assignment = ch;
    switch ((cbr >> 1) & 7) {
    case 0: h->bits_per_sample = 0; break;
    case 1: h->bits_per_sample = 8; break;
    case 2: h->bits_per_sample = 12; break;
    case 4: h->bits_per_sample = 16; break;
    case 5: h->bits_per_sample = 20; break;
    case 6: h->bits_per_sample = 24; break;
    default: return CLX_ERR_FRAME_HEADER_RESERVED;
    }
    if (cbr & 1) return CLX_ERR_FRAME_HEADER_RESERVED;
    size_t used = 0;
    int st = clxo_read_var_length_int(p + at, n - at, &h->number, &used);
    if (st) return st;
    at += used;
    if (!h->variable_blocking && h->number > 0x7fffffffull) return CLX_ERR_FRAME_NUMBER_TOO_LARGE;
    if (bs8) { NEED(1); h->block_size = (uint32_t)p[at++] + 1; }
    if (bs16) {
        NEED(2);
        uint32_t v = ((uint32_t)p[at] << 8) | p[at + 1];
        at += 2;
        if (v == 0xffff) return CLX_ERR_BLOCK_SIZE_65535;
        h->block_size = v + 1;
    }
    if (sr8) { NEED(1); h->sample_rate = p[at++]; }
    if (sr16) { NEED(2); h->sample_rate = ((uint32_t)p[at] << 8) | p[at + 1]; at += 2; }
    if (sr16x10) { NEED(2); h->sample_rate = (((uint32_t)p[at] << 8) | p[at + 1]) * 10; at += 2; }
    NEED(1);
    uint8_t crc = clxo_crc8(p, at);
    uint8_t stored = p[at++];
    h->header_len = (uint32_t)at;
    if (crc != stored) return CLX_ERR_HEADER_CRC_MISMATCH; /* :299 */
#undef NEED
    return CLX_OK;
}

/* src/frame.rs:667-779 */
static int decode_frame_impl(const uint8_t* p, size_t n, int32_t* out, size_t out_cap,
                             clxo_frame_info* info, int verify_crc) {
    memset(info, 0, sizeof *info);
    clxo_frame_header* h = &info->header;
    int st = clxo_read_frame_header(p, n, h);
    if (st == CLX_ERR_HEADER_CRC_MISMATCH && !verify_crc) st = CLX_OK; /* cfg(fuzzing), :295-306 */
    if (st) return st;
    uint32_t bs = h->block_size;
    size_t total = (size_t)h->n_channels * bs;
    info->n_samples = (uint32_t)total;
    if (total > out_cap) return CLX_ERR_INVALID_ARGUMENT;
    if (h->bits_per_sample == 0) return CLX_ERR_NO_BPS_IN_HEADER; /* :687-692 */
    uint32_t bps = h->bits_per_sample;
    uint64_t bit = (uint64_t)h->header_len * 8;
    uint32_t ca = h->channel_assignment;
    if (ca < 8) { /* :706-712 */
        for (uint32_t c = 0; c < h->n_channels; c++) {
            st = clxo_decode_subframe(p, n, &bit, bps, out + (size_t)c * bs, bs);
            if (st) return st;
        }
    } else {
        /* side channel carries one extra bit: :717, :725, :736 */
        uint32_t bps0 = ca == 9 ? bps + 1 : bps;
        uint32_t bps1 = ca == 9 ? bps : bps + 1;
        st = clxo_decode_subframe(p, n, &bit, bps0, out, bs);
        if (st) return st;
        st = clxo_decode_subframe(p, n, &bit, bps1, out + bs, bs);
        if (st) return st;
        if (ca == 8) clxo_decode_left_side(out, 2 * (size_t)bs);
        else if (ca == 9) clxo_decode_right_side(out, 2 * (size_t)bs);
        else clxo_decode_mid_side(out, 2 * (size_t)bs);
    }
    /* :744-750 pad bits to the byte boundary are skipped unchecked */
    uint64_t end = (bit + 7) / 8;
    if (end + 2 > n) return CLX_ERR_IO_UNEXPECTED_EOF; /* :754 read_be_u16 */
    uint16_t stored = (uint16_t)(((uint32_t)p[end] << 8) | p[end + 1]);
    if (verify_crc && clxo_crc16(p, (size_t)end) != stored) return CLX_ERR_FRAME_CRC_MISMATCH;
    info->consumed = end + 2;
    /* :771-774 — note: THIS frame's block size multiplies the frame number */
    info->time = h->variable_blocking ? h->number : (uint64_t)bs * h->number;
    return CLX_OK;
}

int clxo_decode_frame(const uint8_t* p, size_t n, int32_t* out, size_t out_cap,
                      clxo_frame_info* info, int verify_crc) {
    return decode_frame_impl(p, n, out, out_cap, info, verify_crc);
}

/* ------------------------------------------------------------------------- */
/* stream                                                                     */
/* ------------------------------------------------------------------------- */

/* Rust's String::from_utf8 acceptance set (well-formed UTF-8, no surrogates,
 * no overlongs, <= U+10FFFF). */
static int utf8_ok(const uint8_t* s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) { i++; continue; }
        size_t need;
        uint32_t lo = 0x80, hi = 0xbf;
        if (c >= 0xc2 && c <= 0xdf) need = 1;
        else if (c == 0xe0) { need = 2; lo = 0xa0; }
        else if (c >= 0xe1 && c <= 0xec) need = 2;
        else if (c == 0xed) { need = 2; hi = 0x9f; }
        else if (c >= 0xee && c <= 0xef) need = 2;
        else if (c == 0xf0) { need = 3; lo = 0x90; }
        else if (c >= 0xf1 && c <= 0xf3) need = 3;
        else if (c == 0xf4) { need = 3; hi = 0x8f; }
        else return 0;
        if (i + need >= n) return 0; /* truncated sequence */
        if (s[i + 1] < lo || s[i + 1] > hi) return 0;
        for (size_t k = 2; k <= need; k++)
            if ((s[i + k] & 0xc0) != 0x80) return 0;
        i += need + 1;
    }
    return 1;
}

static uint32_t rd_le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* src/metadata.rs:402-513: validation only. */
static int check_vorbis_block(const uint8_t* p, size_t avail, uint32_t length) {
    if (length < 8) return CLX_ERR_VORBIS_TOO_SHORT;
    if (length > 10u * 1024 * 1024) return CLX_ERR_VORBIS_TOO_LARGE;
    size_t at = 0;
#define NEEDV(k) do { if (at + (size_t)(k) > avail) return CLX_ERR_IO_UNEXPECTED_EOF; } while (0)
    NEEDV(4);
    uint32_t vendor_len = rd_le32(p + at); at += 4;
    if (vendor_len > length - 8) return CLX_ERR_VENDOR_TOO_LONG;
    NEEDV(vendor_len);
    if (!utf8_ok(p + at, vendor_len)) return CLX_ERR_UTF8_INVALID;
    at += vendor_len;
    NEEDV(4);
    uint32_t comments_len = rd_le32(p + at); at += 4;
    if (comments_len >= length / 4) return CLX_ERR_VORBIS_TOO_MANY;
    uint32_t bytes_left = length - 8 - vendor_len;
    uint32_t got = 0;
    while (bytes_left >= 4 && got < comments_len) {
        NEEDV(4);
        uint32_t clen = rd_le32(p + at); at += 4;
        bytes_left -= 4;
        if (clen > bytes_left) return CLX_ERR_VORBIS_COMMENT_TOO_LONG;
        if (clen == 0) { comments_len -= 1; continue; }
        NEEDV(clen);
        const uint8_t* c = p + at;
        at += clen;
        bytes_left -= clen;
        size_t sep = clen;
        for (size_t i = 0; i < clen; i++) if (c[i] == '=') { sep = i; break; }
        if (sep == clen) return CLX_ERR_VORBIS_NO_EQUALS;
        for (size_t i = 0; i < sep; i++)
            if (c[i] < 0x20 || c[i] > 0x7d) return CLX_ERR_VORBIS_NAME_INVALID;
        if (!utf8_ok(c, clen)) return CLX_ERR_UTF8_INVALID;
        got++;
    }
#undef NEEDV
    if (bytes_left != 0) return CLX_ERR_VORBIS_EXCESS_DATA;
    if (got != comments_len) return CLX_ERR_VORBIS_WRONG_COUNT;
    return CLX_OK;
}

/* src/lib.rs:186-307 (FlacReader::new with default options) over
 * src/metadata.rs:214-400, :515-549, :557-609. */
int clxo_open_stream(const uint8_t* p, size_t n, clxo_streaminfo* si, uint64_t* first_frame) {
    memset(si, 0, sizeof *si);
    if (n < 4) return CLX_ERR_IO_UNEXPECTED_EOF;
    uint32_t magic = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    if (magic != 0x664c6143u)
        return (magic & 0xffffff00u) == 0x49443300u ? CLX_ERR_STREAM_HEADER_ID3
                                                    : CLX_ERR_STREAM_HEADER_INVALID;
    size_t at = 4;
    int have_si = 0, have_vc = 0, first = 1;
    for (;;) {
        if (at + 4 > n) return CLX_ERR_IO_UNEXPECTED_EOF;
        int is_last = p[at] >> 7;
        uint32_t type = p[at] & 0x7f;
        uint32_t length = ((uint32_t)p[at + 1] << 16) | ((uint32_t)p[at + 2] << 8) | p[at + 3];
        at += 4;
        const uint8_t* body = p + at;
        size_t avail = n - at;
        int st = CLX_OK;
        int is_si = 0, is_vc = 0;
        switch (type) {
        case 0:
            if (length != 34) { st = CLX_ERR_STREAMINFO_LENGTH; break; }
            if (avail < 34) { st = CLX_ERR_IO_UNEXPECTED_EOF; break; }
            {
                clxo_streaminfo t;
                memset(&t, 0, sizeof t);
                t.min_block_size = ((uint32_t)body[0] << 8) | body[1];
                t.max_block_size = ((uint32_t)body[2] << 8) | body[3];
                t.min_frame_size = ((uint32_t)body[4] << 16) | ((uint32_t)body[5] << 8) | body[6];
                t.max_frame_size = ((uint32_t)body[7] << 16) | ((uint32_t)body[8] << 8) | body[9];
                t.sample_rate = ((uint32_t)body[10] << 12) | ((uint32_t)body[11] << 4) | (body[12] >> 4);
                t.channels = ((body[12] >> 1) & 7) + 1;
                t.bits_per_sample = (((uint32_t)(body[12] & 1) << 4) | (body[13] >> 4)) + 1;
                t.samples = ((uint64_t)(body[13] & 15) << 32) | ((uint64_t)body[14] << 24) |
                            ((uint64_t)body[15] << 16) | ((uint64_t)body[16] << 8) | body[17];
                memcpy(t.md5sum, body + 18, 16);
                if (t.min_block_size > t.max_block_size) { st = CLX_ERR_BLOCK_SIZE_BOUNDS; break; }
                if (t.min_block_size < 16) { st = CLX_ERR_BLOCK_SIZE_LT_16; break; }
                if (t.min_frame_size > t.max_frame_size && t.max_frame_size != 0) {
                    st = CLX_ERR_FRAME_SIZE_BOUNDS; break;
                }
                if (t.sample_rate == 0 || t.sample_rate > 655350) { st = CLX_ERR_SAMPLE_RATE_INVALID; break; }
                if (!have_si) *si = t;
                is_si = 1;
            }
            break;
        case 2:
            if (length < 4) { st = CLX_ERR_APPLICATION_TOO_SHORT; break; }
            if (length > 10u * 1024 * 1024) { st = CLX_ERR_APPLICATION_TOO_LARGE; break; }
            if (avail < length) st = CLX_ERR_IO_UNEXPECTED_EOF;
            break;
        case 4:
            st = check_vorbis_block(body, avail, length);
            is_vc = 1;
            break;
        case 127:
            st = CLX_ERR_METADATA_BLOCK_TYPE;
            break;
        default: /* padding, seektable, cuesheet, picture, reserved: skipped by length */
            if (avail < length) st = CLX_ERR_IO_UNEXPECTED_EOF;
            break;
        }
        if (st) return st;
        if (first) { /* lib.rs:244-248 */
            if (!is_si) return CLX_ERR_STREAMINFO_MISSING;
            have_si = 1;
            first = 0;
        } else {
            if (is_vc) { if (have_vc) return CLX_ERR_SECOND_VORBIS_COMMENT; have_vc = 1; }
            if (is_si) return CLX_ERR_SECOND_STREAMINFO;
        }
        at += length;
        if (is_last) break;
    }
    *first_frame = at;
    return CLX_OK;
}

int clxo_decode_stream(const uint8_t* p, size_t n, uint64_t first_frame, int32_t* out,
                       size_t out_cap, uint64_t* n_frames, uint64_t* n_samples_total,
                       int verify_crc) {
    uint64_t at = first_frame, frames = 0, samples = 0;
    int st = CLX_OK;
    while (1) {
        clxo_frame_info info;
        st = decode_frame_impl(p + at, n - (size_t)at, out + samples, out_cap - (size_t)samples,
                               &info, verify_crc);
        if (st == CLX_EOF) { st = CLX_OK; break; }
        if (st) break;
        at += info.consumed;
        samples += info.n_samples;
        frames++;
    }
    *n_frames = frames;
    *n_samples_total = samples;
    return st;
}

/* ------------------------------------------------------------------------- */
/* multi-threaded batch (CPU baseline)                                        */
/* ------------------------------------------------------------------------- */

typedef struct {
    const uint8_t* bytes;
    const uint64_t* offsets;
    const uint32_t* lengths;
    size_t lo, hi;
    int32_t* out;
    const uint64_t* out_offsets;
    int32_t* statuses;
    int verify_crc;
    int bad;
} mt_job;

static void* mt_worker(void* arg) {
    mt_job* j = (mt_job*)arg;
    for (size_t i = j->lo; i < j->hi; i++) {
        clxo_frame_info info;
        int st = decode_frame_impl(j->bytes + j->offsets[i], j->lengths[i],
                                   j->out + j->out_offsets[i], (size_t)1 << 40, &info,
                                   j->verify_crc);
        if (j->statuses) j->statuses[i] = st;
        if (st) j->bad++;
    }
    return NULL;
}

int clxo_decode_batch_mt(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* lengths,
                         size_t n_frames, int32_t* out, const uint64_t* out_offsets,
                         int32_t* statuses, int n_threads, int verify_crc) {
    pthread_once(&g_tab_once, build_tables);
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_frames && n_frames > 0) n_threads = (int)n_frames;
    mt_job* jobs = (mt_job*)calloc((size_t)n_threads, sizeof(mt_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    for (int t = 0; t < n_threads; t++) {
        jobs[t] = (mt_job){bytes, offsets, lengths, n_frames * (size_t)t / (size_t)n_threads,
                           n_frames * (size_t)(t + 1) / (size_t)n_threads, out, out_offsets,
                           statuses, verify_crc, 0};
        if (t > 0) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    mt_worker(&jobs[0]);
    int bad = jobs[0].bad;
    for (int t = 1; t < n_threads; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
    free(jobs);
    free(th);
    return bad;
}
