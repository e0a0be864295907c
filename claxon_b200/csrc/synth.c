/* synth.c — forced-parameter FLAC frame generator ("mini-encoder") for claxon_b200.
 *
 * Produces valid FLAC frames of an exactly prescribed shape (block size, channel
 * assignment, subframe type/order, Rice parameter / partition order, wasted bits,
 * Rice2, ...) for parity tests and the benchmark workloads of BASELINE.json.
 * Frames are built by *synthesis*: residuals are drawn from a two-sided geometric
 * distribution, the decoder recurrence is run forward to obtain the PCM, and the
 * residuals are then Rice coded with the prescribed parameters — so the PCM a
 * correct decoder must return is known by construction and returned alongside.
 *
 * Bitstream layout follows the FLAC format as consumed by the reference decoder
 * (frame header: reference src/frame.rs:131-316; subframe header / warm-up / LPC
 * parameters / residual: src/subframe.rs:29-91, :236-380, :651-721; CRC-8/16:
 * src/crc.rs).  Never emits the Rice escape code (the reference rejects it,
 * src/subframe.rs:317-319).  Plain C, no dependencies; not part of the decode path.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct clxs_config {
    uint64_t seed;
    uint32_t n_frames;
    uint32_t block_size;
    uint32_t tail_block_size;   /* if != 0: every frames_per_file-th frame uses this size */
    uint32_t frames_per_file;   /* 0 = one long stream; else frame numbers restart */
    uint32_t n_channels;        /* 1..8 (2 for stereo modes) */
    uint32_t bps;               /* 8, 12, 16, 20 or 24 */
    uint32_t sample_rate_code;  /* 4-bit code 0..11 */
    int32_t stereo_mode;        /* 0 independent, 8 L/S, 9 R/S, 10 M/S, -1 random of the four */
    uint32_t type_mask;         /* bit0 constant, bit1 verbatim, bit2 fixed, bit3 lpc */
    uint32_t lpc_min_order, lpc_max_order;     /* 1..32 */
    uint32_t fixed_min_order, fixed_max_order; /* 0..4 */
    uint32_t qlp_precision;     /* 1..15, 0 = random in 8..15 */
    int32_t rice_mode;          /* >= 0: forced k; -1: optimal per partition; -2: k0 per subframe in
                                   [kmin,kmax], partitions k0-1..k0+1 */
    uint32_t rice_kmin, rice_kmax;
    uint32_t min_porder, max_porder;
    uint32_t rice2;             /* 0 never, 1 always, 2 random per subframe */
    uint32_t wasted_max;        /* 0 = never; else ~1/4 of subframes get 1..wasted_max wasted bits */
    double residual_mean;       /* mean |e|; 0 = derive from k (2^k * 0.7) */
    uint32_t variable_blocking; /* 1 = sample numbers in the header */
    uint32_t long_unary_per_mille; /* chance (per 1000 subframes) to inject one huge residual */
    uint32_t force_bs16;        /* always use the 16-bit explicit block size code */
} clxs_config;

typedef struct clxs_stream {
    uint8_t* bytes;
    size_t nbytes, cap;
    uint64_t* frame_offsets; /* n_frames + 1 */
    int32_t* pcm;            /* planar per frame, back to back */
    uint64_t* pcm_offsets;   /* n_frames + 1 (elements) */
    uint32_t n_frames;
    uint64_t n_samples;
} clxs_stream;

/* ---- rng ---- */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t* x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed);
}
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static double rng_unit(rng_t* r) { return (double)((rng_next(r) >> 11) + 1) * (1.0 / 9007199254740993.0); }
static uint32_t rng_range(rng_t* r, uint32_t lo, uint32_t hi) { /* inclusive */
    if (hi <= lo) return lo;
    return lo + (uint32_t)(rng_next(r) % (uint64_t)(hi - lo + 1));
}

/* ---- crc ---- */
static uint8_t crc8_tab[256];
static uint16_t crc16_tab[256];
static int tabs_ready = 0;
static void init_tabs(void) {
    if (tabs_ready) return;
    for (int i = 0; i < 256; i++) {
        uint8_t c = (uint8_t)i;
        uint16_t d = (uint16_t)(i << 8);
        for (int k = 0; k < 8; k++) {
            c = (uint8_t)((c & 0x80) ? ((c << 1) ^ 0x07) : (c << 1));
            d = (uint16_t)((d & 0x8000) ? ((d << 1) ^ 0x8005) : (d << 1));
        }
        crc8_tab[i] = c;
        crc16_tab[i] = d;
    }
    tabs_ready = 1;
}

/* ---- bit writer over the growing stream buffer ---- */
typedef struct { clxs_stream* st; uint64_t acc; uint32_t nacc; } bitw;
static void ensure(clxs_stream* s, size_t extra) {
    if (s->nbytes + extra <= s->cap) return;
    size_t ncap = s->cap * 2 + extra + 4096;
    s->bytes = (uint8_t*)realloc(s->bytes, ncap);
    s->cap = ncap;
}
static void put_byte(clxs_stream* s, uint8_t b) { ensure(s, 1); s->bytes[s->nbytes++] = b; }
static void bw_put(bitw* w, uint32_t value, uint32_t nbits) { /* nbits <= 32 */
    if (nbits == 0) return;
    uint64_t v = nbits == 32 ? value : (value & ((1u << nbits) - 1u));
    w->acc = (w->acc << nbits) | v;
    w->nacc += nbits;
    while (w->nacc >= 8) {
        put_byte(w->st, (uint8_t)(w->acc >> (w->nacc - 8)));
        w->nacc -= 8;
    }
}
static void bw_zeros(bitw* w, uint64_t n) {
    while (n >= 24) { bw_put(w, 0, 24); n -= 24; }
    bw_put(w, 0, (uint32_t)n);
}
static void bw_align(bitw* w) { if (w->nacc) bw_put(w, 0, 8 - w->nacc); }

/* ---- predictor design ---- */
/* Random stable all-pole model through reflection coefficients (step-up), quantised
 * to `precision` bits with the largest shift that fits. a[j] predicts from x[n-1-j]. */
static void design_lpc(rng_t* r, uint32_t order, uint32_t precision, double damp, int32_t* qc,
                       uint32_t* shift_out) {
    double a[33], tmp[33];
    memset(a, 0, sizeof a);
    for (uint32_t m = 1; m <= order; m++) {
        double limit = (m == 1 ? 0.95 : 0.6) * damp;
        double k = (2.0 * rng_unit(r) - 1.0) * limit;
        if (m == 1) k = fabs(k) * 0.5 + 0.45 * damp; /* audio-like: strongly positive first lag */
        for (uint32_t j = 1; j < m; j++) tmp[j] = a[j] - k * a[m - j];
        for (uint32_t j = 1; j < m; j++) a[j] = tmp[j];
        a[m] = k;
    }
    double amax = 1e-9;
    for (uint32_t j = 1; j <= order; j++) if (fabs(a[j]) > amax) amax = fabs(a[j]);
    int32_t lim = (1 << (precision - 1)) - 1;
    int sh = (int)floor(log2((double)lim / amax));
    if (sh > 15) sh = 15;
    if (sh < 0) sh = 0;
    for (uint32_t j = 1; j <= order; j++) {
        double v = a[j] * (double)(1 << sh);
        int32_t q = (int32_t)lrint(v);
        if (q > lim) q = lim;
        if (q < -lim - 1) q = -lim - 1;
        qc[j - 1] = q;
    }
    *shift_out = (uint32_t)sh;
}

static int32_t draw_residual(rng_t* r, double mean) {
    double u = rng_unit(r);
    double m = floor(-log(u) * mean);
    if (m > 1.0e9) m = 1.0e9;
    int32_t v = (int32_t)m;
    return (rng_next(r) & 1) ? -v : v;
}

static inline uint32_t zigzag(int32_t e) { return ((uint32_t)e << 1) ^ (uint32_t)(e >> 31); }

static uint64_t rice_bits(const int32_t* e, uint32_t n, uint32_t k) {
    uint64_t bits = 0;
    for (uint32_t i = 0; i < n; i++) bits += (uint64_t)(zigzag(e[i]) >> k) + 1 + k;
    return bits;
}

typedef struct {
    int type; /* 0 const, 1 verbatim, 2 fixed, 3 lpc */
    uint32_t order, wasted, precision, shift, rice2, porder;
    int32_t coefs[32];
} sf_plan;

/* Generates one coded channel: fills x[0..bs) (pre-wasted-shift values) and writes the
 * subframe bits.  `sf_bps` is the width warm-up/verbatim samples are stored with. */
static void gen_subframe(const clxs_config* cfg, rng_t* r, bitw* w, uint32_t bs, uint32_t bps,
                         int32_t* x, int32_t* e) {
    sf_plan p;
    memset(&p, 0, sizeof p);
    /* type */
    uint32_t mask = cfg->type_mask ? cfg->type_mask : 8u;
    int choices[4], nch = 0;
    for (int t = 0; t < 4; t++) if (mask & (1u << t)) choices[nch++] = t;
    p.type = choices[rng_next(r) % (uint64_t)nch];
    if (p.type == 2) p.order = rng_range(r, cfg->fixed_min_order, cfg->fixed_max_order > 4 ? 4 : cfg->fixed_max_order);
    if (p.type == 3) p.order = rng_range(r, cfg->lpc_min_order ? cfg->lpc_min_order : 1,
                                         cfg->lpc_max_order ? cfg->lpc_max_order : 1);
    if (p.order > bs) p.order = bs;
    if (p.type == 3 && p.order == 0) { p.type = 2; }
    if (cfg->wasted_max && (rng_next(r) & 3) == 0) p.wasted = rng_range(r, 1, cfg->wasted_max);
    if (p.wasted >= bps) p.wasted = bps - 1;
    uint32_t sf_bps = bps - p.wasted;
    int64_t lim = ((int64_t)1 << (sf_bps - 1)) - 1;
    p.rice2 = cfg->rice2 == 2 ? (uint32_t)(rng_next(r) & 1) : cfg->rice2;
    uint32_t kcap = p.rice2 ? 30u : 14u;

    /* partition order: 2^po | bs and order <= bs >> po */
    uint32_t po = rng_range(r, cfg->min_porder, cfg->max_porder);
    while (po > 0 && ((bs & ((1u << po) - 1u)) != 0 || p.order > (bs >> po))) po--;
    p.porder = po;

    /* residual scale */
    int32_t k0 = cfg->rice_mode;
    if (cfg->rice_mode < 0) k0 = (int32_t)rng_range(r, cfg->rice_kmin, cfg->rice_kmax);
    if ((uint32_t)k0 > kcap) k0 = (int32_t)kcap;
    double mean = cfg->residual_mean > 0 ? cfg->residual_mean : ldexp(0.72, k0);
    if (mean < 0.3) mean = 0.3;

    /* ---- header ---- */
    uint32_t code = p.type == 0 ? 0u : p.type == 1 ? 1u : p.type == 2 ? (8u | p.order) : (32u | (p.order - 1));
    bw_put(w, 0, 1);
    bw_put(w, code, 6);
    if (p.wasted) { bw_put(w, 1, 1); bw_zeros(w, p.wasted - 1); bw_put(w, 1, 1); }
    else bw_put(w, 0, 1);

    if (p.type == 0) {
        int32_t v = (int32_t)((int64_t)(rng_next(r) % (uint64_t)(2 * lim + 1)) - lim);
        bw_put(w, (uint32_t)v, sf_bps);
        for (uint32_t i = 0; i < bs; i++) x[i] = (int32_t)((uint32_t)v << p.wasted);
        return;
    }
    if (p.type == 1) {
        for (uint32_t i = 0; i < bs; i++) {
            x[i] = (int32_t)((int64_t)(rng_next(r) % (uint64_t)(2 * lim + 1)) - lim);
            bw_put(w, (uint32_t)x[i], sf_bps);
            x[i] = (int32_t)((uint32_t)x[i] << p.wasted);
        }
        return;
    }

    static const int32_t fixed_rows[5][4] = {{0,0,0,0},{1,0,0,0},{2,-1,0,0},{3,-3,1,0},{4,-6,4,-1}};
    uint32_t nres = bs - p.order;
    if (p.type == 2) {
        /* Fixed predictors are unstable as synthesis filters (poles at z=1), so build the
         * signal first — a slow sinusoid plus white noise — and ANALYSE it: e = x - pred. */
        static const double binom[5] = {1.0, 2.0, 6.0, 20.0, 70.0}; /* C(2o,o): noise gain^2 */
        p.shift = 0;
        for (uint32_t j = 0; j < p.order; j++) p.coefs[j] = fixed_rows[p.order][j];
        double wf = 6.283185307179586 / (double)rng_range(r, 400, 4000);
        double amp = (double)lim * 0.25, cap = 0.5 * mean / pow(wf, (double)p.order);
        if (amp > cap) amp = cap;
        double noise = mean / (0.8 * sqrt(binom[p.order]));
        if (noise * 8 > (double)lim * 0.5) noise = (double)lim / 16.0;
        double ph = rng_unit(r) * 6.283185307179586;
        for (uint32_t i = 0; i < bs; i++) {
            int32_t nz = draw_residual(r, noise);
            double v = amp * sin(wf * i + ph) + (double)nz;
            if (v > (double)lim) v = (double)lim;
            if (v < (double)(-lim - 1)) v = (double)(-lim - 1);
            x[i] = (int32_t)lrint(v);
        }
        for (uint32_t i = p.order; i < bs; i++) {
            uint32_t pred = 0;
            for (uint32_t j = 0; j < p.order; j++) pred += (uint32_t)p.coefs[j] * (uint32_t)x[i - 1 - j];
            e[i - p.order] = (int32_t)((uint32_t)x[i] - pred);
        }
    } else {
        /* LPC: random stable all-pole model, SYNTHESISED from drawn residuals; retried with
         * more damping until every sample is representable. */
        double damp = 1.0;
        for (int attempt = 0;; attempt++) {
            p.precision = cfg->qlp_precision ? cfg->qlp_precision : rng_range(r, 8, 15);
            design_lpc(r, p.order, p.precision, damp, p.coefs, &p.shift);
            double amp = (double)lim * 0.02 * damp;
            double base = (2.0 * rng_unit(r) - 1.0) * amp, slope = (2.0 * rng_unit(r) - 1.0) * mean;
            for (uint32_t i = 0; i < p.order; i++) {
                double v = base + slope * i;
                if (v > (double)lim) v = (double)lim;
                if (v < (double)(-lim - 1)) v = (double)(-lim - 1);
                x[i] = (int32_t)lrint(v);
            }
            for (uint32_t i = 0; i < nres; i++) e[i] = draw_residual(r, mean);
            int ok = 1;
            for (uint32_t i = p.order; i < bs; i++) {
                int64_t acc = 0;
                for (uint32_t j = 0; j < p.order; j++) acc += (int64_t)p.coefs[j] * (int64_t)x[i - 1 - j];
                int64_t v = (acc >> p.shift) + (int64_t)e[i - p.order];
                if ((v > lim || v < -lim - 1) && attempt < 12) { ok = 0; break; }
                x[i] = (int32_t)(uint32_t)(uint64_t)v;
            }
            if (ok) break;
            damp *= 0.7;
        }
    }
    if (cfg->long_unary_per_mille && nres > 0 && rng_range(r, 0, 999) < cfg->long_unary_per_mille) {
        /* one outlier residual: exercises unary runs far longer than a machine word */
        uint32_t at = rng_range(r, 0, nres - 1);
        e[at] += (int32_t)((rng_next(r) & 1) ? -1 : 1) * (int32_t)rng_range(r, 200, 4000) * (int32_t)(mean + 1);
        for (uint32_t i = p.order; i < bs; i++) { /* re-synthesise: x is whatever the decoder computes */
            int64_t acc = 0;
            for (uint32_t j = 0; j < p.order; j++) acc += (int64_t)p.coefs[j] * (int64_t)x[i - 1 - j];
            x[i] = (int32_t)(uint32_t)(uint64_t)((acc >> p.shift) + (int64_t)e[i - p.order]);
        }
    }

    /* ---- warm-up, LPC parameters ---- */
    for (uint32_t i = 0; i < p.order; i++) bw_put(w, (uint32_t)x[i], sf_bps);
    if (p.type == 3) {
        bw_put(w, p.precision - 1, 4);
        bw_put(w, p.shift, 5);
        for (uint32_t j = 0; j < p.order; j++) bw_put(w, (uint32_t)p.coefs[j], p.precision);
    }
    /* ---- residual ---- */
    bw_put(w, p.rice2 ? 1u : 0u, 2);
    bw_put(w, p.porder, 4);
    uint32_t per = bs >> p.porder, at = 0, len = per - p.order;
    for (uint32_t part = 0; part < (1u << p.porder); part++) {
        uint32_t k;
        if (cfg->rice_mode >= 0) k = (uint32_t)cfg->rice_mode;
        else if (cfg->rice_mode == -1) {
            uint64_t best = ~0ull;
            k = 0;
            for (uint32_t t = 0; t <= kcap; t++) {
                uint64_t b = rice_bits(e + at, len, t);
                if (b < best) { best = b; k = t; }
            }
        } else {
            int32_t kk = k0 + (int32_t)rng_range(r, 0, 2) - 1;
            if (kk < (int32_t)cfg->rice_kmin) kk = (int32_t)cfg->rice_kmin;
            if (kk > (int32_t)cfg->rice_kmax) kk = (int32_t)cfg->rice_kmax;
            k = (uint32_t)kk;
        }
        if (k > kcap) k = kcap;
        bw_put(w, k, p.rice2 ? 5 : 4);
        for (uint32_t i = 0; i < len; i++) {
            uint32_t v = zigzag(e[at + i]);
            bw_zeros(w, v >> k);
            bw_put(w, 1, 1);
            bw_put(w, v & ((k == 0) ? 0u : ((1u << k) - 1u)), k);
        }
        at += len;
        len = per;
    }
    if (p.wasted) /* src/subframe.rs:216-225: decoded samples are shifted back up */
        for (uint32_t i = 0; i < bs; i++) x[i] = (int32_t)((uint32_t)x[i] << p.wasted);
}

static void put_varint(clxs_stream* s, uint64_t v) {
    if (v < 0x80) { put_byte(s, (uint8_t)v); return; }
    int extra = 1;
    while (extra < 6 && v >= ((uint64_t)1 << (6 * extra + (6 - extra)))) extra++;
    /* first byte: (extra+1) leading ones, a zero, then 6-extra data bits */
    uint8_t lead = (uint8_t)(0xff << (7 - extra));
    put_byte(s, (uint8_t)(lead | (uint8_t)(v >> (6 * extra))));
    for (int i = extra - 1; i >= 0; i--) put_byte(s, (uint8_t)(0x80 | ((v >> (6 * i)) & 0x3f)));
}

static void gen_frame(const clxs_config* cfg, clxs_stream* s, uint32_t index, int32_t* pcm,
                      int32_t* scratch) {
    rng_t r;
    rng_seed(&r, cfg->seed + (uint64_t)index * 0x632be59bd9b4e019ull + 1);
    uint32_t bs = cfg->block_size;
    uint64_t number = index;
    if (cfg->frames_per_file) {
        number = index % cfg->frames_per_file;
        if (cfg->tail_block_size && number == cfg->frames_per_file - 1) bs = cfg->tail_block_size;
    } else if (cfg->tail_block_size && index == cfg->n_frames - 1) bs = cfg->tail_block_size;
    uint32_t nch = cfg->n_channels;
    int32_t mode = cfg->stereo_mode;
    if (mode == -1) { static const int32_t m[4] = {0, 8, 9, 10}; mode = m[rng_next(&r) & 3]; }
    if (nch != 2) mode = 0;
    uint32_t ca = mode ? (uint32_t)mode : nch - 1;

    size_t start = s->nbytes;
    /* ---- frame header ---- */
    put_byte(s, 0xff);
    put_byte(s, (uint8_t)(0xf8 | (cfg->variable_blocking ? 1 : 0)));
    uint32_t bs_code = 0;
    if (!cfg->force_bs16) {
        if (bs == 192) bs_code = 1;
        for (uint32_t c = 2; c <= 5; c++) if (bs == (576u << (c - 2))) bs_code = c;
        for (uint32_t c = 8; c <= 15; c++) if (bs == (256u << (c - 8))) bs_code = c;
        if (!bs_code) bs_code = bs <= 256 ? 6 : 7;
    } else bs_code = 7;
    put_byte(s, (uint8_t)((bs_code << 4) | (cfg->sample_rate_code & 15)));
    uint32_t bps_code = cfg->bps == 8 ? 1 : cfg->bps == 12 ? 2 : cfg->bps == 16 ? 4 : cfg->bps == 20 ? 5 : 6;
    put_byte(s, (uint8_t)((ca << 4) | (bps_code << 1)));
    if (cfg->variable_blocking) {
        uint64_t first = cfg->frames_per_file ? (uint64_t)(index % cfg->frames_per_file) * cfg->block_size
                                              : (uint64_t)index * cfg->block_size;
        put_varint(s, first);
    } else put_varint(s, number);
    if (bs_code == 6) put_byte(s, (uint8_t)(bs - 1));
    if (bs_code == 7) { put_byte(s, (uint8_t)((bs - 1) >> 8)); put_byte(s, (uint8_t)(bs - 1)); }
    uint8_t c8 = 0;
    for (size_t i = start; i < s->nbytes; i++) c8 = crc8_tab[c8 ^ s->bytes[i]];
    put_byte(s, c8);

    /* ---- subframes ---- */
    bitw w = {s, 0, 0};
    for (uint32_t c = 0; c < nch; c++) {
        uint32_t bps = cfg->bps;
        if ((mode == 8 || mode == 10) && c == 1) bps++;
        if (mode == 9 && c == 0) bps++;
        gen_subframe(cfg, &r, &w, bs, bps, pcm + (size_t)c * bs, scratch);
    }
    bw_align(&w);
    /* expected decoded output: undo the inter-channel decorrelation (src/frame.rs:319-389) */
    if (mode == 8)
        for (uint32_t i = 0; i < bs; i++) pcm[bs + i] = (int32_t)((uint32_t)pcm[i] - (uint32_t)pcm[bs + i]);
    else if (mode == 9)
        for (uint32_t i = 0; i < bs; i++) pcm[i] = (int32_t)((uint32_t)pcm[i] + (uint32_t)pcm[bs + i]);
    else if (mode == 10)
        for (uint32_t i = 0; i < bs; i++) {
            uint32_t side = (uint32_t)pcm[bs + i];
            uint32_t mid = ((uint32_t)pcm[i] << 1) | (side & 1u);
            pcm[i] = ((int32_t)(mid + side)) >> 1;
            pcm[bs + i] = ((int32_t)(mid - side)) >> 1;
        }
    uint16_t c16 = 0;
    for (size_t i = start; i < s->nbytes; i++)
        c16 = (uint16_t)((c16 << 8) ^ crc16_tab[(uint8_t)(c16 >> 8) ^ s->bytes[i]]);
    put_byte(s, (uint8_t)(c16 >> 8));
    put_byte(s, (uint8_t)c16);
}

typedef struct {
    const clxs_config* cfg;
    uint32_t lo, hi;
    clxs_stream part;      /* bytes of frames [lo, hi) */
    uint64_t* lens;        /* per-frame byte length */
    int32_t* pcm;          /* global pcm buffer */
    const uint64_t* pcm_offsets;
} gen_job;

static void* gen_worker(void* arg) {
    gen_job* j = (gen_job*)arg;
    int32_t* scratch = (int32_t*)malloc(sizeof(int32_t) * 65536);
    for (uint32_t i = j->lo; i < j->hi; i++) {
        size_t before = j->part.nbytes;
        gen_frame(j->cfg, &j->part, i, j->pcm + j->pcm_offsets[i], scratch);
        j->lens[i] = j->part.nbytes - before;
    }
    free(scratch);
    return NULL;
}

clxs_stream* clxs_generate_mt(const clxs_config* cfg, int n_threads) {
    init_tabs();
    if (!cfg || cfg->n_channels < 1 || cfg->n_channels > 8 || cfg->block_size < 1 ||
        cfg->block_size > 65535)
        return NULL;
    clxs_stream* s = (clxs_stream*)calloc(1, sizeof *s);
    s->n_frames = cfg->n_frames;
    s->frame_offsets = (uint64_t*)calloc((size_t)cfg->n_frames + 1, sizeof(uint64_t));
    s->pcm_offsets = (uint64_t*)calloc((size_t)cfg->n_frames + 1, sizeof(uint64_t));
    uint64_t total = 0;
    for (uint32_t i = 0; i < cfg->n_frames; i++) {
        uint32_t bs = cfg->block_size;
        if (cfg->tail_block_size) {
            if (cfg->frames_per_file ? (i % cfg->frames_per_file == cfg->frames_per_file - 1)
                                     : (i == cfg->n_frames - 1))
                bs = cfg->tail_block_size;
        }
        s->pcm_offsets[i] = total;
        total += (uint64_t)bs * cfg->n_channels;
    }
    s->pcm_offsets[cfg->n_frames] = total;
    s->n_samples = total;
    s->pcm = (int32_t*)malloc(sizeof(int32_t) * (size_t)(total ? total : 1));
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > cfg->n_frames) n_threads = cfg->n_frames ? (int)cfg->n_frames : 1;
    gen_job* jobs = (gen_job*)calloc((size_t)n_threads, sizeof(gen_job));
    pthread_t* th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
    uint64_t* lens = (uint64_t*)calloc((size_t)cfg->n_frames + 1, sizeof(uint64_t));
    for (int t = 0; t < n_threads; t++) {
        jobs[t].cfg = cfg;
        jobs[t].lo = (uint32_t)((uint64_t)cfg->n_frames * (uint64_t)t / (uint64_t)n_threads);
        jobs[t].hi = (uint32_t)((uint64_t)cfg->n_frames * (uint64_t)(t + 1) / (uint64_t)n_threads);
        jobs[t].lens = lens;
        jobs[t].pcm = s->pcm;
        jobs[t].pcm_offsets = s->pcm_offsets;
        jobs[t].part.cap = (size_t)(total / (uint64_t)n_threads * cfg->bps / 8 / 2) + 65536;
        jobs[t].part.bytes = (uint8_t*)malloc(jobs[t].part.cap);
        if (t > 0) pthread_create(&th[t], NULL, gen_worker, &jobs[t]);
    }
    gen_worker(&jobs[0]);
    size_t nbytes = jobs[0].part.nbytes;
    for (int t = 1; t < n_threads; t++) { pthread_join(th[t], NULL); nbytes += jobs[t].part.nbytes; }
    s->bytes = (uint8_t*)malloc(nbytes + 64);
    s->cap = nbytes + 64;
    for (int t = 0; t < n_threads; t++) {
        memcpy(s->bytes + s->nbytes, jobs[t].part.bytes, jobs[t].part.nbytes);
        s->nbytes += jobs[t].part.nbytes;
        free(jobs[t].part.bytes);
    }
    memset(s->bytes + s->nbytes, 0, 64);
    uint64_t at = 0;
    for (uint32_t i = 0; i < cfg->n_frames; i++) { s->frame_offsets[i] = at; at += lens[i]; }
    s->frame_offsets[cfg->n_frames] = at;
    free(lens); free(jobs); free(th);
    return s;
}

clxs_stream* clxs_generate(const clxs_config* cfg) { return clxs_generate_mt(cfg, 1); }

const uint8_t* clxs_bytes(const clxs_stream* s) { return s->bytes; }
uint64_t clxs_nbytes(const clxs_stream* s) { return s->nbytes; }
const uint64_t* clxs_frame_offsets(const clxs_stream* s) { return s->frame_offsets; }
const int32_t* clxs_pcm(const clxs_stream* s) { return s->pcm; }
const uint64_t* clxs_pcm_offsets(const clxs_stream* s) { return s->pcm_offsets; }
uint64_t clxs_n_samples(const clxs_stream* s) { return s->n_samples; }
void clxs_free(clxs_stream* s) {
    if (!s) return;
    free(s->bytes); free(s->frame_offsets); free(s->pcm); free(s->pcm_offsets); free(s);
}
