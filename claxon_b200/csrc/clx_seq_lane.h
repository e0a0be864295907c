// clx_seq_lane.h — the per-lane half of the sequential entropy decoder (clx_seq.cu).
//
// One LANE decodes one FRAME's bitstream sequentially: subframe header, warm-up samples, LPC
// parameters, residual header and the Rice partitions (reference src/subframe.rs:29-91, :236-380,
// :382-415, :651-701), channel after channel, exactly in the order claxon reads them — so there is
// nothing to speculate about and nothing to scan: a lane always knows its bit position.  32 frames
// advance per warp instruction.  The lane does not predict: it leaves the residuals in a scratch
// laid out for the prediction kernel (one lane per subframe) and the predictor parameters in a
// per-subframe record.
//
// The logic lives in this header, free of CUDA built-ins, so that the very same code runs on the
// host inside the test harness (tools/seq_host.cpp, driven by tests/test_seq_host.py): the lanes of a
// warp never talk to each other, which makes the kernel a plain loop over lanes on a CPU.  All
// memory traffic goes through the `IO` policy: on the device a shared-memory ring fed by cp.async
// and vector stores to the scratch, on the host plain loads and stores.
//
// Anything irregular (malformed input of any kind, the Rice escape code, a residual that does not
// fit the scratch's sample width) is not guessed at: the frame is flagged and the generic kernel
// (clx_decode.cu), which owns claxon's error precedence, decodes it afterwards.
#ifndef CLX_SEQ_LANE_H
#define CLX_SEQ_LANE_H
#include <stdint.h>

#include "claxon_b200.h"

#ifdef __CUDACC__
#define CLX_HD __host__ __device__ __forceinline__
#else
#define CLX_HD inline
#endif

namespace clx {

// One per subframe; written by the entropy lane, read by the prediction lane.
struct SeqParams {
    int32_t order;      // predictor order; 0 = the residual is the sample (verbatim / fixed-0)
    int32_t shift;      // qlp shift (0 for fixed predictors)
    int32_t wasted;     // wasted bits per sample (src/subframe.rs:216-225)
    uint32_t absum;     // sum |coef|
    int16_t coefs[32];  // coefs[j] multiplies s[t-1-j]
    int32_t warm[32];   // warm-up samples s[0..order)
};

// Residual scratch.  The 32 frames of an entropy warp and one channel share a block of rows; a row
// is 32 lanes x 16 bytes, lane l's 16 bytes holding 8 consecutive residuals of its subframe as i16
// (narrow: streams of at most 16 bits per sample) or 4 as i32 (wide).  A warp-wide 16-byte store
// or load of one row is therefore one fully coalesced 512-byte transaction, for the entropy lanes
// (lane = frame) and for the prediction lanes (lane = subframe) alike.
constexpr uint32_t SEQ_ROW_BYTES = 512;
template <bool NARROW>
CLX_HD uint32_t seq_rows_for(uint32_t max_bs) { return NARROW ? (max_bs + 7) / 8 : (max_bs + 3) / 4; }
template <bool NARROW>
CLX_HD uint64_t seq_elem_offset(uint32_t t) {  // byte offset of residual t inside the lane's column
    return NARROW ? (uint64_t)(t >> 3) * SEQ_ROW_BYTES + (t & 7) * 2 : (uint64_t)(t >> 2) * SEQ_ROW_BYTES + (t & 3) * 4;
}

CLX_HD uint32_t hd_clz(uint32_t v) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__clz((int)v);
#else
    return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}
CLX_HD uint32_t hd_msb(uint32_t v) {  // index of the most significant set bit (v != 0)
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
#else
    return 31u - (uint32_t)__builtin_clz(v | 1u);
#endif
}
CLX_HD uint32_t hd_neg_lsb(uint32_t u) {  // 0 - (u & 1)
#ifdef __CUDA_ARCH__
    int32_t r;
    asm("bfe.s32 %0, %1, 0, 1;" : "=r"(r) : "r"(u));
    return (uint32_t)r;
#else
    return 0u - (u & 1u);
#endif
}
// upper 32 bits of (hi:lo) << (n & 31)
CLX_HD uint32_t hd_fsl(uint32_t hi, uint32_t lo, uint32_t n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, n);
#else
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
#endif
}
CLX_HD uint32_t hd_bswap(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0, 0x0123);
#else
    return __builtin_bswap32(v);
#endif
}
CLX_HD int32_t hd_sext(uint32_t v, uint32_t bits) {  // bits in [1, 32]
    return ((int32_t)(v << (32 - bits))) >> (32 - bits);
}

enum : uint32_t { SEQ_SUBFRAME = 0, SEQ_PART = 1, SEQ_RUN = 2, SEQ_DONE = 3 };

// IO policy (see DeviceIO in clx_seq.cu and HostIO in tools/seq_host.cpp):
//   uint32_t word(uint32_t wi)            big-endian word `wi` of the frame (relative to its 16-byte aligned base)
//   void ensure(uint32_t bitpos)          the next 2048 bits from bitpos are readable through word()
//   bool prefetch_group(uint32_t bitpos)  steady-state refill, once per fast group; false: take the slow path
//   void seek_next(uint32_t wi), uint32_t next_raw()   sequential word reads, bytes as stored (the register window's refill)
//   void select_channel(uint32_t ch)      subsequent stores go to channel ch's rows
//   void store8(uint32_t t, const int32_t (&e)[8])   residuals t..t+7, t % 8 == 0
//   void store1(uint32_t t, int32_t e)
template <class IO, bool NARROW>
struct SeqLane {
    IO io;
    SeqParams* params;  // the frame's CH records
    uint32_t bs, nch, ca, fbps, bit0, limit, byte_len;
    uint32_t o;           // bit cursor, relative to the frame's 16-byte aligned base
    // Register window (maintained while mode == SEQ_RUN): big-endian words o>>5 and (o>>5)+1, and word
    // (o>>5)+2 as loaded (little-endian): it is byte-swapped only when it moves up, one word later, so the
    // swap never waits for the shared-memory load that produced it.
    uint32_t W0, W1, W2;
    uint32_t mode, ch, t, n_left, parts_left, per, order, pbits;
    uint32_t k, K, Kneg, K30, c32k, thr;
    uint32_t consumed;
    bool ok, slow_next, first_part;

    CLX_HD void init(const clx_frame_desc& d, SeqParams* p, uint32_t max_channels) {
        params = p;
        bs = d.block_size; nch = d.n_channels; ca = d.channel_assignment; fbps = d.bits_per_sample;
        byte_len = d.byte_len;
        bit0 = (uint32_t)(d.byte_offset & 15) * 8;
        limit = bit0 + d.byte_len * 8;
        o = bit0 + (uint32_t)d.header_len * 8;
        W0 = W1 = W2 = 0;
        mode = SEQ_SUBFRAME; ch = 0; t = 0; n_left = 0; parts_left = 0; per = 0; order = 0; pbits = 4;
        k = 0; K = 1; Kneg = 0xffffffffu; K30 = 30; c32k = 32; thr = 1;
        consumed = 0;
        ok = true; slow_next = false; first_part = false;
        if (nch > max_channels || fbps == 0 || (NARROW && fbps > 16)) fail();
    }
    CLX_HD void fail() { ok = false; mode = SEQ_DONE; }
    CLX_HD bool done() const { return mode == SEQ_DONE; }
    CLX_HD bool fast_ready() const { return mode == SEQ_RUN && !slow_next && n_left >= 8 && (t & 7u) == 0; }

    CLX_HD uint32_t peek32(uint32_t pos) { return hd_fsl(io.word(pos >> 5), io.word((pos >> 5) + 1), pos); }
    CLX_HD uint32_t bits(uint32_t pos, uint32_t n) { return n ? peek32(pos) >> (32 - n) : 0u; }  // n <= 32
    CLX_HD void window_seek() {
        const uint32_t wi = o >> 5;
        W0 = io.word(wi); W1 = io.word(wi + 1);
        io.seek_next(wi + 2);
        W2 = io.next_raw();
    }
    CLX_HD void emit1(int32_t e) {  // one residual through the slow path
        if (NARROW && (e < -32768 || e > 32767)) { fail(); return; }
        io.store1(t, e);
        t++;
    }
    // after the last residual of a partition
    CLX_HD void advance() {
        if (parts_left) { mode = SEQ_PART; return; }
        if (o > limit) { fail(); return; }
        ch++;
        mode = SEQ_SUBFRAME;
    }

    // ---- eight Rice codes (src/subframe.rs:336-348) ----
    // Precondition fast_ready().  A code is taken here when it fits the 32-bit window `hi`
    // (unary + terminator + k bits <= 32) and, for the narrow scratch, its value fits 16 bits: both
    // are the single comparison hi >= thr.  The eight codes are decoded straight through; if any of
    // them failed the comparison, what came after it is meaningless (but harmless: nothing is
    // stored, and every shared-memory address is masked into the lane's ring), the cursor is put
    // back and the slow path takes the codes one by one until the next group boundary.
    CLX_HD void fast_group() {
        if (!io.prefetch_group(o)) { slow_next = true; return; }  // ring not far enough ahead: the slow path refills it
        const uint32_t o0 = o;
        int32_t e[8];
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t hi = hd_fsl(W0, W1, o);
            bad = bad || hi < thr;
            const uint32_t m = hd_msb(hi);        // terminator at bit m: unary quotient q = 31 - m
            const uint32_t v = hi >> ((m - k) & 31u);  // K + r
            const uint32_t u = m * Kneg + (v + K30);  // (q << k) | r = (30 - m) * K + v
            // rice_to_signed (src/subframe.rs:157-170): (u >> 1) ^ -(u & 1).  With u = 2h + b that is h for b = 0 and
            // ~h = h - u for b = 1, i.e. h + (-b) * u in wrapping arithmetic: the multiply-add runs on the FMA pipe,
            // which this loop leaves half idle, instead of a third ALU operation (the ALU pipe is what bounds it).
            e[i] = (int32_t)((u >> 1) + hd_neg_lsb(u) * u);
            const uint32_t on = o + c32k - m;     // o + q + 1 + k
            if ((on ^ o) >> 5) { W0 = W1; W1 = hd_bswap(W2); W2 = io.next_raw(); }
            o = on;
        }
        if (bad) {
            o = o0;  // the slow path re-reads the window (and re-seeds the sequential reads) after its code
            slow_next = true;
            return;
        }
        io.store8(t, e);
        t += 8;
        n_left -= 8;
        if (n_left == 0) advance();
    }

    // ---- one Rice code of any shape ----
    CLX_HD void slow_code() {
        slow_next = false;
        uint32_t q = 0;
        uint32_t v;
        for (;;) {
            io.ensure(o);
            v = peek32(o);
            if (v != 0) break;
            q += 32; o += 32;
            if (o > limit) { fail(); return; }
        }
        const uint32_t z = hd_clz(v);
        q += z;
        o += z + 1;
        const uint32_t r = bits(o, k);
        o += k;
        const uint32_t u = (q << k) | r;  // wrapping, as the reference's u32 arithmetic
        emit1((int32_t)((u >> 1) ^ (0u - (u & 1u))));
        if (!ok) return;
        n_left--;
        if (n_left == 0) advance();
        else window_seek();
    }

    // ---- partition header (src/subframe.rs:310-319, :358-367) ----
    CLX_HD void do_part() {
        io.ensure(o);
        k = bits(o, pbits);
        o += pbits;
        if (k == (1u << pbits) - 1u) { fail(); return; }  // escape code: Unsupported in the reference
        n_left = first_part ? per - order : per;
        first_part = false;
        parts_left--;
        K = 1u << k;
        K30 = 30u * K;
        Kneg = 0u - K;
        c32k = 32u + k;
        // fast-path bound on the unary quotient q: q + 1 + k <= 32, and (narrow) (q + 1) << k <= 65536
        uint32_t qmax = 31u - k;
        if (NARROW) {
            if (k > 16) { fail(); return; }
            const uint32_t fit = (65536u >> k) - 1u;
            qmax = qmax < fit ? qmax : fit;
        }
        thr = 1u << (31u - qmax);
        mode = SEQ_RUN;
        if (n_left == 0) advance();
        else window_seek();
    }

    // ---- subframe header, warm-up, predictor parameters, residual header ----
    CLX_HD void do_subframe() {
        if (ch >= nch) {  // frame footer: pad to the byte boundary, the CRC-16 must be readable (src/frame.rs:744-754)
            const uint32_t end_byte = (o - bit0 + 7) >> 3;
            consumed = end_byte + 2;
            if (o > limit || consumed > byte_len) ok = false;
            mode = SEQ_DONE;
            return;
        }
        io.ensure(o);
        io.select_channel(ch);
        uint32_t bps = fbps;
        if (ca == 9) bps += (ch == 0);                   // side/right: side first (src/frame.rs:725)
        else if (ca == 8 || ca == 10) bps += (ch == 1);  // src/frame.rs:717, :736
        SeqParams* sp = params + ch;
        // subframe header (src/subframe.rs:29-91)
        const uint32_t head = bits(o, 8);
        o += 8;
        if (head & 0x80u) { fail(); return; }
        const uint32_t code = (head >> 1) & 0x3fu;
        order = 0;
        int type;
        if (code == 0) type = 0;
        else if (code == 1) type = 1;
        else if ((code & 0x3eu) == 0x02u || (code & 0x3cu) == 0x04u || (code & 0x30u) == 0x10u) { fail(); return; }
        else if ((code & 0x38u) == 0x08u) { order = code & 7u; if (order > 4) { fail(); return; } type = 2; }
        else { order = (code & 0x1fu) + 1; type = 3; }
        uint32_t wasted = 0;
        if (head & 1u) {
            const uint32_t v = peek32(o);
            if (v == 0) { fail(); return; }  // > 31 wasted bits: an error for the generic kernel to name
            const uint32_t q = hd_clz(v);
            wasted = q + 1;
            o += q + 1;
        }
        if (wasted >= bps) { fail(); return; }
        const uint32_t sfbps = bps - wasted;
        if (sfbps > 32) { fail(); return; }
        if ((type == 2 || type == 3) && order > bs) { fail(); return; }
        sp->wasted = (int32_t)wasted;
        if (type == 0) {  // constant (src/subframe.rs:382-394): an order-1 predictor with coefficient 1 over zero residuals
            const int32_t v = hd_sext(bits(o, sfbps), sfbps);
            o += sfbps;
            sp->order = 1; sp->shift = 0; sp->absum = 1; sp->coefs[0] = 1; sp->warm[0] = v;
            const int32_t zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            t = 1;
            while (t < bs) {
                if ((t & 7u) == 0 && t + 8 <= bs) { io.store8(t, zeros); t += 8; }
                else { io.store1(t, 0); t++; }
            }
            if (o > limit) { fail(); return; }
            ch++;
            return;  // mode stays SEQ_SUBFRAME
        }
        if (type == 1) {  // verbatim (src/subframe.rs:397-415): residuals of an order-0 predictor
            sp->order = 0; sp->shift = 0; sp->absum = 0;
            t = 0;
            for (uint32_t i = 0; i < bs && ok; i++) {
                if ((i & 15u) == 0) io.ensure(o);
                emit1(hd_sext(bits(o, sfbps), sfbps));
                o += sfbps;
                if (o > limit) fail();
            }
            if (!ok) return;
            ch++;
            return;
        }
        for (uint32_t i = 0; i < order; i++) {  // warm-up
            if ((i & 15u) == 0) io.ensure(o);
            sp->warm[i] = hd_sext(bits(o, sfbps), sfbps);
            o += sfbps;
        }
        if (o > limit) { fail(); return; }
        io.ensure(o);
        uint32_t shift = 0, absum = 0;
        if (type == 3) {  // src/subframe.rs:669-701
            const uint32_t pq = bits(o, 9);  // 4-bit precision-1, 5-bit signed shift
            o += 9;
            const uint32_t prec_m1 = pq >> 5;
            if (prec_m1 == 15) { fail(); return; }
            const uint32_t precision = prec_m1 + 1;
            const int32_t sh = hd_sext(pq & 31u, 5);
            if (sh < 0) { fail(); return; }
            shift = (uint32_t)sh;
            for (uint32_t j = 0; j < order; j++) {
                const int32_t c = hd_sext(bits(o, precision), precision);
                o += precision;
                sp->coefs[j] = (int16_t)c;
                absum += (uint32_t)(c < 0 ? -c : c);
            }
        } else {  // rows of src/subframe.rs:427-431; coefs[0] multiplies s[t-1]
            const uint32_t packed = order == 1 ? 0x00000001u : order == 2 ? 0x0000ff02u
                                  : order == 3 ? 0x0001fd03u : order == 4 ? 0xff04fa04u : 0u;
            for (uint32_t j = 0; j < order; j++) {
                const int32_t c = (int32_t)(int8_t)(packed >> (8 * j));
                sp->coefs[j] = (int16_t)c;
                absum += (uint32_t)(c < 0 ? -c : c);
            }
        }
        sp->order = (int32_t)order; sp->shift = (int32_t)shift; sp->absum = absum;
        if (o > limit) { fail(); return; }
        io.ensure(o);
        // residual header (src/subframe.rs:236-304)
        const uint32_t rh = bits(o, 6);  // 2-bit coding method, 4-bit partition order
        o += 6;
        const uint32_t method = rh >> 4, po = rh & 15u;
        if (method > 1) { fail(); return; }
        const uint32_t n_part = 1u << po;
        if ((bs & ((n_part - 1u) & 0xffffu)) != 0) { fail(); return; }
        per = bs >> po;
        if (order > per) { fail(); return; }
        pbits = method == 0 ? 4u : 5u;
        parts_left = n_part;
        first_part = true;
        t = order;
        mode = SEQ_PART;
    }

    // everything that is not a fast group
    CLX_HD void slow_step() {
        if (mode == SEQ_SUBFRAME) do_subframe();
        if (mode == SEQ_PART) do_part();
        else if (mode == SEQ_RUN && !fast_ready()) slow_code();
    }
};

}  // namespace clx
#endif
