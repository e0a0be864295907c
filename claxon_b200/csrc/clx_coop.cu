// clx_coop.cu — warp-cooperative frame decode: the fast path of claxon_b200.
//
// One CTA decodes a group of G frames held entirely in shared memory, in three phases:
//
//   1. ENTROPY (one warp per frame).  The warp walks its frame's bitstream — subframe header,
//      warm-up samples, LPC parameters, residual header, Rice partitions (reference
//      src/subframe.rs:29-91, :236-380, :382-415, :651-701) — and decodes the Rice residuals of a
//      partition 1024 bits at a time with all 32 lanes: each lane owns one 32-bit word of the
//      window, finds the unary terminators of its word for the speculated code phase, the
//      phases are made exact with a shuffle fix-point (a lane's entry phase is the previous
//      lane's exit phase; chains that meet a terminator of the speculated chain merge with it),
//      ballot-free popcount + shuffle prefix scans give every code its output rank and the end
//      of the previous code (hence its quotient), and each lane then emits its codes
//      (rice_to_signed, src/subframe.rs:157-170) straight into the frame's sample buffer.
//   2. PREDICTION (one lane per subframe).  predict_fixed / predict_lpc_* (src/subframe.rs:
//      417-474, :524-614) are strictly serial recurrences — the floor in `>> qlp_shift` makes
//      them non-associative — so every subframe of the group gets one lane, coefficients and
//      history register-resident, i64 accumulate, arithmetic shift, truncating cast, in place.
//   3. OUTPUT (all threads).  Wasted-bits shift (src/subframe.rs:216-225), inter-channel
//      decorrelation (src/frame.rs:319-389) and coalesced 16-byte stores of the planar Block
//      layout (src/frame.rs:477-481).
//
// Anything this path does not handle exactly — malformed input of any kind, the Rice escape
// code, unary runs longer than a window, frames larger than the shared-memory budget — is not
// guessed at: the frame is flagged and the generic lane-per-frame kernel (clx_decode.cu), which
// reproduces claxon's error precedence, decodes it afterwards.
#include <cuda_runtime.h>
#include <stdint.h>

#include "claxon_b200.h"
#include <algorithm>
#include <cstdlib>

#include "clx_internal.h"

namespace clx {

struct SubParams {   // one per subframe, shared memory
    int32_t order;   // predictor order; 0 = nothing to predict (constant / verbatim / fixed-0)
    int32_t shift;   // qlp shift (0 for fixed predictors)
    int32_t wasted;  // wasted bits per sample
    uint32_t narrow; // 0: predicted with the i64 accumulator; else sum|coef| of a subframe predicted with
                     // the i32 accumulator (phase 3 verifies that this was exact)
    int16_t coefs[32];  // coefs[j] multiplies s[t-1-j]
};

struct GroupHeader {  // per frame of the group, shared memory
    int32_t ok;        // 1 = decoded by this kernel, 0 = flagged for the generic kernel / absent
    uint32_t consumed; // bytes incl. CRC-16
};

// ---------------------------------------------------------------------------------
// Warp-wide bit window: 64 consecutive big-endian words of the frame in two registers per lane
// (X = words b0..b0+31, Y = words b0+32..b0+63).  Y is loaded one window ahead of its use, so
// HBM/L2 latency is off the critical path; all field extraction is shuffle + funnel shift.
// ---------------------------------------------------------------------------------
struct Win {
    const uint32_t* base;  // 4-byte aligned global address at or before the frame's first byte
    uint32_t wlim;         // first word index that lies outside the byte buffer (reads give 0)
    uint32_t b0;           // word index of X lane 0
    uint32_t X, Y;
};

// X holds big-endian (byte-swapped) words ready for bit arithmetic; Y holds them RAW as loaded, so
// that nothing touches a freshly loaded register until it moves into X a window later.
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint32_t win_ldg(const Win& w, uint32_t idx) {
    uint32_t v = 0;
    if (idx < w.wlim) v = __ldg(w.base + idx);
    return v;
}
__device__ __forceinline__ void win_prime(Win& w, uint32_t bitpos, uint32_t lane) {
    w.b0 = bitpos >> 5;
    w.X = bswap32(win_ldg(w, w.b0 + lane));
    w.Y = win_ldg(w, w.b0 + 32 + lane);
}
// Slides the window so that the word containing `bitpos` is X lane 0.
__device__ __forceinline__ void win_advance(Win& w, uint32_t bitpos, uint32_t lane) {
    const uint32_t d = (bitpos >> 5) - w.b0;
    if (d == 0) return;
    if (d > 32) { win_prime(w, bitpos, lane); return; }
    const uint32_t src = (lane + d) & 31;
    const uint32_t xs = __shfl_sync(0xffffffffu, w.X, src);
    const uint32_t ys = __shfl_sync(0xffffffffu, w.Y, src);
    const bool low = lane + d < 32;
    w.X = low ? xs : bswap32(ys);
    w.b0 += d;
    uint32_t fresh = 0;
    if (!low) fresh = win_ldg(w, w.b0 + 32 + lane);  // consumed one window later
    w.Y = low ? ys : fresh;
}
// 32 bits starting at `bitpos` (uniform across the warp); requires bitpos>>5 in [b0, b0+30].
__device__ __forceinline__ uint32_t win_peek32(const Win& w, uint32_t bitpos) {
    const uint32_t i = (bitpos >> 5) - w.b0;
    const uint32_t w0 = __shfl_sync(0xffffffffu, w.X, i & 31);
    const uint32_t w1 = __shfl_sync(0xffffffffu, w.X, (i + 1) & 31);  // callers keep i <= 30
    return __funnelshift_l(w1, w0, bitpos & 31);
}
__device__ __forceinline__ uint32_t top_bits(uint32_t v, uint32_t n) {  // n in [0,32]
    return n >= 32 ? v : __funnelshift_l(v, 0, n);
}
// Per-lane variant: lane-specific bit positions inside the window (word index <= b0+30).
__device__ __forceinline__ uint32_t win_peek32_lane(const Win& w, uint32_t bitpos) {
    const uint32_t i = (bitpos >> 5) - w.b0;
    const uint32_t w0 = __shfl_sync(0xffffffffu, w.X, i & 31);
    const uint32_t w1 = __shfl_sync(0xffffffffu, w.X, (i + 1) & 31);  // callers keep i <= 30
    return __funnelshift_l(w1, w0, bitpos & 31);
}
__device__ __forceinline__ int32_t sext(uint32_t v, uint32_t bits) {
    return ((int32_t)(v << (32 - bits))) >> (32 - bits);
}

// Terminators of one word for a search that starts at bit `o`: walks code by code (terminator,
// then k remainder bits), merging with the speculated chain (tm0, x0) as soon as both hit the
// same terminator.  Returns the terminator mask and the exit phase (search offset in the next word).
__device__ __forceinline__ void walk_word(uint32_t W, uint32_t o, uint32_t k, bool merge, uint32_t tm0, uint32_t x0,
                                          uint32_t& tm, uint32_t& x) {
    tm = 0;
    x = 0;
    for (;;) {
        const uint32_t m = W & (0xffffffffu >> o);
        if (m == 0) { x = 0; break; }
        const uint32_t t = __clz(m);
        const uint32_t bit = 0x80000000u >> t;
        if (merge && (tm0 & bit)) {
            tm |= tm0 & (bit | (bit - 1));
            x = x0;
            break;
        }
        tm |= bit;
        o = t + 1 + k;
        if (o >= 32) { x = o - 32; break; }
    }
}

// Decodes up to `n_rem` Rice codes with parameter k starting at bit `P` from the current window
// into out[0..); returns the number decoded (0 = cannot make progress here) and advances P to
// the end of the last decoded code.
__device__ __forceinline__ uint32_t rice_window(const Win& w, uint32_t& P, uint32_t k, uint32_t n_rem, int32_t* out,
                                                uint32_t lane) {
    // Lanes 0..30 own one word each (992 bits per window); lane 31 only lends its word as the
    // right-hand neighbour, so the window never reads Y, whose newest words may still be in flight.
    const uint32_t W = lane < 31 ? w.X : 0u;
    const uint32_t WN = __shfl_down_sync(0xffffffffu, w.X, 1);
    const uint32_t s = P - (w.b0 << 5);  // < 32: search offset of lane 0

    // speculated chain: every lane assumes its word starts a fresh search
    uint32_t tm0, x0;
    walk_word(W, 0, k, false, 0, 0, tm0, x0);
    // fix-point on the entry phases
    uint32_t e = __shfl_up_sync(0xffffffffu, x0, 1);
    if (lane == 0) e = s;
    uint32_t tm = tm0, x = x0;
    bool need = e != 0;
    for (;;) {
        if (need) walk_word(W, e, k, true, tm0, x0, tm, x);
        uint32_t xe = __shfl_up_sync(0xffffffffu, x, 1);
        need = lane > 0 && xe != e;
        if (need) e = xe;
        if (!__any_sync(0xffffffffu, need)) break;
    }
    // ranks
    const uint32_t cnt = __popc(tm);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    uint32_t excl = incl - cnt;
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total > n_rem) {  // the partition ends inside this window: keep the first n_rem codes
        if (excl >= n_rem) tm = 0;
        else if (incl > n_rem) {
            uint32_t keep = 0, tmp = tm;
            for (uint32_t j = 0; j < n_rem - excl; j++) {
                const uint32_t bit = 0x80000000u >> __clz(tmp);
                keep |= bit;
                tmp &= ~bit;
            }
            tm = keep;
        }
        total = n_rem;
    }
    if (total == 0) return 0;
    // end of the last code at or before each lane (window-relative bit offset)
    uint32_t lane_end = 0;
    if (tm) lane_end = (lane << 5) + (32 - __ffs(tm)) + 1 + k;
    uint32_t endi = lane_end;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, endi, d);
        if (lane >= (uint32_t)d) endi = max(endi, v);
    }
    uint32_t start = __shfl_up_sync(0xffffffffu, endi, 1);
    if (lane == 0) start = s;
    start = max(start, s);
    const uint32_t new_end = __shfl_sync(0xffffffffu, endi, 31);
    // emit
    uint32_t idx = excl, rest = tm;
    while (rest) {
        const uint32_t t = __clz(rest);
        rest &= ~(0x80000000u >> t);
        const uint32_t pos = (lane << 5) + t;
        const uint32_t q = pos - start;
        const uint32_t hi = __funnelshift_lc(WN, W, t + 1);
        const uint32_t r = __funnelshift_l(hi, 0, k);
        const uint32_t u = (q << k) | r;
        out[idx++] = (int32_t)((u >> 1) ^ (0u - (u & 1u)));
        start = pos + 1 + k;
    }
    P = (w.b0 << 5) + new_end;
    return total;
}

// ---------------------------------------------------------------------------------
// Phase 2 helper: the recurrence for one subframe, TAPS taps, in place.
// ---------------------------------------------------------------------------------
// One step of the recurrence for U consecutive samples.  v[0..TAPS) = history (oldest first),
// v[TAPS+i] = sample i of this trip.  Terms that only involve history are summed first (they do
// not depend on this trip's samples), the terms with fresh samples last, most recent last — the
// serial chain per sample is then one IMAD.WIDE, the shift and the residual add.
template <int TAPS, int U, typename ACC>
__device__ __forceinline__ void predict_trip(int32_t (&v)[TAPS + U], const int32_t (&c)[TAPS], const int32_t (&r)[U],
                                             uint32_t shift) {
    ACC part[U];
#pragma unroll
    for (int i = 0; i < U; i++) {
        ACC acc = 0;
#pragma unroll
        for (int j = 0; j < TAPS; j++)  // c[j] multiplies v[i + TAPS - 1 - j]; history only here
            if (i + TAPS - 1 - j < TAPS) acc += (ACC)c[j] * (ACC)v[i + TAPS - 1 - j];
        part[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < U; i++) {
        ACC acc = part[i];
#pragma unroll
        for (int j = TAPS - 1; j >= 0; j--)  // fresh samples, oldest first
            if (i + TAPS - 1 - j >= TAPS) acc += (ACC)c[j] * (ACC)v[i + TAPS - 1 - j];
        v[TAPS + i] = (int32_t)(acc >> shift) + r[i];
    }
}

// The recurrence for one subframe per lane, in place.  Lanes run in lockstep on t; the bulk of the
// block is decoded by a predicate-free loop, the ragged head (warm-up, differing orders) and tail
// (differing block sizes) by a guarded one.
//
// ACC = long long is the reference's arithmetic verbatim (i64 products and sum).  ACC = int is the
// same recurrence with 32-bit wrapping multiply-adds — 2.3x cheaper on this chip — and yields
// bit-identical samples whenever no sum of products leaves the i32 range, i.e. whenever
// sum|coef| * max|sample| < 2^31: the caller picks it only where that is expected, and phase 3
// re-checks it against the samples actually produced (if it ever fails the frame is re-decoded by
// the generic kernel, so the output never depends on the shortcut).
template <int TAPS, int U, typename ACC>
__device__ __forceinline__ void predict_inplace(int32_t* buf, uint32_t bs, uint32_t order, uint32_t shift,
                                                const int16_t* coefs, bool active) {
    int32_t c[TAPS], h[TAPS];  // c[j] multiplies s[t-1-j]; h[j] = s[t-1-j]
#pragma unroll
    for (int j = 0; j < TAPS; j++) {
        c[j] = (active && (uint32_t)j < order) ? (int32_t)coefs[j] : 0;
        // Opaque to the optimiser: otherwise the i16 -> i64 promotion is folded into a full 64-bit
        // multiply (3 instructions) instead of one signed 32x32+64 IMAD.WIDE per tap.
        asm volatile("" : "+r"(c[j]));
        h[j] = 0;
    }
    const uint32_t max_bs = __reduce_max_sync(0xffffffffu, active ? bs : 0u);
    const uint32_t min_bs = __reduce_min_sync(0xffffffffu, active ? bs : 0xffffffffu);
    const uint32_t max_order = __reduce_max_sync(0xffffffffu, active ? order : 0u);
    const uint32_t head_end = min(max_bs, (max_order + (uint32_t)U - 1) / (uint32_t)U * (uint32_t)U);
    const uint32_t bulk_end = head_end + (min_bs > head_end ? (min_bs - head_end) / (uint32_t)U * (uint32_t)U : 0u);

    auto guarded = [&](uint32_t t0, uint32_t t1) {  // one sample at a time, every condition checked
        for (uint32_t t = t0; t < t1; t++) {
            const bool inside = active && t < bs;
            int32_t val = inside ? buf[t] : 0;
            if (t >= order) {
                long long acc = 0;
#pragma unroll
                for (int j = 0; j < TAPS; j++) acc += (long long)c[j] * (long long)h[j];
                val += sizeof(ACC) == 8 ? (int32_t)(acc >> shift) : (int32_t)((int32_t)acc >> shift);
                if (inside) buf[t] = val;
            }
#pragma unroll
            for (int j = TAPS - 1; j > 0; j--) h[j] = h[j - 1];
            h[0] = val;
        }
    };
    guarded(0, head_end);
    if (bulk_end > head_end) {
        int32_t v[TAPS + U];
#pragma unroll
        for (int j = 0; j < TAPS; j++) v[j] = h[TAPS - 1 - j];
        int32_t rn[U];  // residuals are fetched one trip ahead: shared-memory latency stays off the chain
#pragma unroll
        for (int i = 0; i < U; i++) rn[i] = buf[head_end + i];
        for (uint32_t t = head_end; t < bulk_end; t += U) {
            int32_t r[U];
#pragma unroll
            for (int i = 0; i < U; i++) r[i] = rn[i];
            if (t + U < bulk_end) {
#pragma unroll
                for (int i = 0; i < U; i++) rn[i] = buf[t + U + i];
            }
            predict_trip<TAPS, U, ACC>(v, c, r, shift);
            if (active) {
#pragma unroll
                for (int i = 0; i < U; i++) buf[t + i] = v[TAPS + i];
            }
#pragma unroll
            for (int j = 0; j < TAPS; j++) v[j] = v[j + U];
        }
#pragma unroll
        for (int j = 0; j < TAPS; j++) h[j] = v[TAPS - 1 - j];
    }
    guarded(bulk_end, max_bs);
}

__device__ __forceinline__ void decor(uint32_t ca, int32_t a, int32_t b, int32_t& o0, int32_t& o1) {
    if (ca == 8) { o0 = a; o1 = (int32_t)((uint32_t)a - (uint32_t)b); }
    else if (ca == 9) { o0 = (int32_t)((uint32_t)a + (uint32_t)b); o1 = b; }
    else {
        const uint32_t m = ((uint32_t)a << 1) | ((uint32_t)b & 1u);
        o0 = ((int32_t)(m + (uint32_t)b)) >> 1;
        o1 = ((int32_t)(m - (uint32_t)b)) >> 1;
    }
}

// ---------------------------------------------------------------------------------
// The kernel.  blockDim = 32 * G; dynamic shared memory = G * frame_stride * 4 + tables.
// ---------------------------------------------------------------------------------
constexpr int COOP_MAX_G = 8;
constexpr int COOP_MAX_CH = 8;

__global__ void __launch_bounds__(COOP_MAX_G * 32)
decode_frames_coop_kernel(const uint8_t* __restrict__ bytes, uint64_t buf_bytes, const clx_frame_desc* __restrict__ descs,
                          uint32_t n_frames, int32_t* __restrict__ out, clx_frame_result* __restrict__ results,
                          int* __restrict__ need_generic, uint32_t G, uint32_t frame_stride /* i32 elements */,
                          uint32_t CH /* channel slots per frame = max channels in the batch */,
                          uint32_t dbg /* timing experiments only: bit1 skip phase 2, bit2 skip phase 3 */) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    int32_t* s_buf = reinterpret_cast<int32_t*>(smem_raw);
    SubParams* s_par = reinterpret_cast<SubParams*>(smem_raw + (size_t)G * frame_stride * 4);
    GroupHeader* s_hdr = reinterpret_cast<GroupHeader*>(s_par + G * CH);

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t fidx = blockIdx.x * G + warp;

    // =========================================================================== phase 1
    {
        bool ok = fidx < n_frames;
        clx_frame_desc d;
        if (ok) d = descs[fidx];
        else { d.block_size = 0; d.n_channels = 0; d.bits_per_sample = 0; d.channel_assignment = 0; d.byte_len = 0;
               d.byte_offset = 0; d.header_len = 0; d.out_offset = 0; }
        const uint32_t bs = d.block_size, nch = d.n_channels, ca = d.channel_assignment;
        if (ok && ((uint64_t)bs * nch > frame_stride || nch > CH || d.bits_per_sample == 0)) ok = false;
        int32_t* fbuf = s_buf + (size_t)warp * frame_stride;
        Win w;
        const uint64_t aligned = d.byte_offset & ~3ull;
        w.base = reinterpret_cast<const uint32_t*>(bytes + aligned);
        w.wlim = (uint32_t)min((buf_bytes - aligned) >> 2, (uint64_t)0x7fffffffu);
        const uint32_t bit0 = (uint32_t)(d.byte_offset & 3) * 8;
        const uint32_t limit = bit0 + d.byte_len * 8;
        uint32_t P = bit0 + (uint32_t)d.header_len * 8;
        if (ok) win_prime(w, P, lane);
        else { w.b0 = 0; w.X = 0; w.Y = 0; }

        for (uint32_t ch = 0; ok && ch < nch; ch++) {
            uint32_t bps = d.bits_per_sample;
            if (ca == 9) bps += (ch == 0);
            else if (ca == 8 || ca == 10) bps += (ch == 1);
            int32_t* sbuf = fbuf + (size_t)ch * bs;
            SubParams* sp = &s_par[warp * CH + ch];
            win_advance(w, P, lane);
            // ---- subframe header (src/subframe.rs:29-91) ----
            uint32_t head = win_peek32(w, P) >> 24;
            P += 8;
            if (head & 0x80u) { ok = false; break; }
            const uint32_t code = (head >> 1) & 0x3fu;
            uint32_t order = 0;
            int type;
            if (code == 0) type = 0;
            else if (code == 1) type = 1;
            else if ((code & 0x3eu) == 0x02u || (code & 0x3cu) == 0x04u || (code & 0x30u) == 0x10u) { ok = false; break; }
            else if ((code & 0x38u) == 0x08u) { order = code & 7u; if (order > 4) { ok = false; break; } type = 2; }
            else { order = (code & 0x1fu) + 1; type = 3; }
            uint32_t wasted = 0;
            if (head & 1u) {
                const uint32_t v = win_peek32(w, P);
                if (v == 0) { ok = false; break; }  // > 31 wasted bits: an error for the generic kernel to name
                const uint32_t q = __clz(v);
                wasted = q + 1;
                P += q + 1;
            }
            if (wasted >= bps) { ok = false; break; }
            const uint32_t sfbps = bps - wasted;
            if (sfbps > 30) { ok = false; break; }
            if ((type == 2 || type == 3) && order > bs) { ok = false; break; }
            if (lane == 0) { sp->order = 0; sp->shift = 0; sp->wasted = (int32_t)wasted; sp->narrow = 0; }
            if (type == 0) {  // constant (src/subframe.rs:382-394)
                const int32_t v = sext(top_bits(win_peek32(w, P), sfbps), sfbps);
                P += sfbps;
                for (uint32_t i = lane; i < bs; i += 32) sbuf[i] = v;
                if (P > limit) { ok = false; break; }
                continue;
            }
            // ---- verbatim samples: the whole subframe, or the warm-up (src/subframe.rs:397-415) ----
            const uint32_t n_raw = type == 1 ? bs : order;
            for (uint32_t i0 = 0; i0 < n_raw; i0 += 32) {
                win_advance(w, P, lane);
                const uint32_t i = i0 + lane;
                const uint32_t pos = P + lane * sfbps;
                const uint32_t v = win_peek32_lane(w, pos);
                if (i < n_raw) sbuf[i] = sext(top_bits(v, sfbps), sfbps);
                P += min(32u, n_raw - i0) * sfbps;
            }
            if (P > limit) { ok = false; break; }
            if (type == 1) continue;
            // ---- predictor parameters (src/subframe.rs:427-431, :669-701) ----
            win_advance(w, P, lane);
            uint32_t shift = 0;
            if (type == 3) {
                const uint32_t pq = win_peek32(w, P) >> 23;
                P += 9;
                const uint32_t prec_m1 = pq >> 5;
                if (prec_m1 == 15) { ok = false; break; }
                const uint32_t precision = prec_m1 + 1;
                const int32_t sh = sext(pq & 31u, 5);
                if (sh < 0) { ok = false; break; }
                shift = (uint32_t)sh;
                const uint32_t v = win_peek32_lane(w, P + lane * precision);
                if (lane < order) sp->coefs[lane] = (int16_t)sext(top_bits(v, precision), precision);
                P += order * precision;
            } else if (lane < 4) {
                // Pascal rows with alternating sign; coefs[0] multiplies s[t-1]
                // row `order` of {1}, {2,-1}, {3,-3,1}, {4,-6,4,-1}, packed one nibble-pair per entry
                const uint32_t packed = order == 1 ? 0x00000001u : order == 2 ? 0x0000ff02u
                                      : order == 3 ? 0x0001fd03u : order == 4 ? 0xff04fa04u : 0u;
                sp->coefs[lane] = (int16_t)(int8_t)(packed >> (8 * lane));
            }
            if (lane == 0) { sp->order = (int32_t)order; sp->shift = (int32_t)shift; }
            // ---- residual (src/subframe.rs:236-380) ----
            win_advance(w, P, lane);
            const uint32_t rh = win_peek32(w, P) >> 26;
            P += 6;
            const uint32_t method = rh >> 4, po = rh & 15u;
            if (method > 1) { ok = false; break; }
            const uint32_t n_part = 1u << po;
            if ((bs & ((n_part - 1u) & 0xffffu)) != 0) { ok = false; break; }
            const uint32_t per = bs >> po;
            if (order > per) { ok = false; break; }
            const uint32_t pbits = method == 0 ? 4u : 5u;
            uint32_t at = order;
            for (uint32_t part = 0; ok && part < n_part; part++) {
                win_advance(w, P, lane);
                const uint32_t k = win_peek32(w, P) >> (32 - pbits);
                P += pbits;
                if (k == (1u << pbits) - 1u) { ok = false; break; }  // escape code: Unsupported in the reference
                uint32_t n_rem = part == 0 ? per - order : per;
                while (n_rem > 0) {
                    win_advance(w, P, lane);
                    const uint32_t got = rice_window(w, P, k, n_rem, sbuf + at, lane);
                    if (got == 0 || P > limit) { ok = false; break; }
                    at += got;
                    n_rem -= got;
                }
            }
            if (P > limit) ok = false;
        }
        // frame footer: pad to the byte boundary, the CRC-16 must be readable (src/frame.rs:744-754)
        uint32_t consumed = 0;
        if (ok) {
            const uint32_t end_byte = (P - bit0 + 7) >> 3;
            consumed = end_byte + 2;
            if (P > limit || consumed > d.byte_len) ok = false;
        }
        if (lane == 0) {
            s_hdr[warp].ok = ok ? 1 : 0;
            s_hdr[warp].consumed = consumed;
            if (fidx < n_frames) {
                clx_frame_result res;
                res.status = ok ? (int32_t)CLX_OK : (int32_t)CLX_INTERNAL_NEED_GENERIC;
                res.consumed = consumed;
                results[fidx] = res;
                if (!ok) *need_generic = 1;
            }
        }
    }
    __syncthreads();

    // =========================================================================== phase 2
    if (!(dbg & 2))
    // subframe slot q = frame * CH + channel; lanes of warp w take slots [32w, 32w+32)
    {
        const uint32_t slots = G * CH;
        for (uint32_t q0 = warp * 32; q0 < slots; q0 += blockDim.x) {
            const uint32_t q = q0 + lane;
            const uint32_t f = q / CH, c = q % CH;
            bool active = false, narrow_ok = false;
            uint32_t bs = 0, order = 0, shift = 0, absum = 0;
            const int16_t* coefs = s_par[0].coefs;
            SubParams* lane_sp = nullptr;
            int32_t* sbuf = s_buf;
            if (q < slots && s_hdr[f].ok) {
                const uint32_t gf = blockIdx.x * G + f;
                const uint32_t nch = descs[gf].n_channels;
                if (c < nch) {
                    bs = descs[gf].block_size;
                    SubParams* sp = &s_par[f * CH + c];
                    lane_sp = sp;
                    order = (uint32_t)sp->order;
                    shift = (uint32_t)sp->shift;
                    coefs = sp->coefs;
                    sbuf = s_buf + (size_t)f * frame_stride + (size_t)c * bs;
                    active = order > 0;
                    if (active) {
                        for (uint32_t j = 0; j < order; j++) absum += (uint32_t)abs((int)coefs[j]);
                        // nominal sample width of this subframe (one extra bit for a side channel)
                        const uint32_t ca = descs[gf].channel_assignment;
                        uint32_t bits = descs[gf].bits_per_sample;
                        if (ca == 9) bits += (c == 0); else if (ca == 8 || ca == 10) bits += (c == 1);
                        narrow_ok = ((unsigned long long)absum << bits) < (1ull << 31);
                    }
                }
            }
            const uint32_t max_order = __reduce_max_sync(0xffffffffu, active ? order : 0u);
            if (max_order == 0) continue;
            // idle lanes read (never write) some valid buffer so that the bulk loop needs no guards
            const uint32_t some = __ffs(__ballot_sync(0xffffffffu, active)) - 1;
            const unsigned long long alias = __shfl_sync(0xffffffffu, (unsigned long long)sbuf, some);
            if (!active) sbuf = reinterpret_cast<int32_t*>(alias);
            // i32 accumulator where sum|coef| * 2^sample_bits leaves headroom in i32 for every lane
            const bool all_narrow = __all_sync(0xffffffffu, !active || narrow_ok);
            if (active && lane_sp != nullptr) lane_sp->narrow = all_narrow ? absum : 0u;
            if (all_narrow) {
                if (max_order <= 4) predict_inplace<4, 4, int>(sbuf, bs, order, shift, coefs, active);
                else if (max_order <= 8) predict_inplace<8, 8, int>(sbuf, bs, order, shift, coefs, active);
                else if (max_order <= 12) predict_inplace<12, 4, int>(sbuf, bs, order, shift, coefs, active);
                else predict_inplace<32, 4, int>(sbuf, bs, order, shift, coefs, active);
            } else {
                if (max_order <= 4) predict_inplace<4, 4, long long>(sbuf, bs, order, shift, coefs, active);
                else if (max_order <= 8) predict_inplace<8, 8, long long>(sbuf, bs, order, shift, coefs, active);
                else if (max_order <= 12) predict_inplace<12, 4, long long>(sbuf, bs, order, shift, coefs, active);
                else predict_inplace<32, 4, long long>(sbuf, bs, order, shift, coefs, active);
            }
        }
    }
    __syncthreads();

    // =========================================================================== phase 3
    for (uint32_t f = 0; f < G; f++) {
        if (!s_hdr[f].ok || (dbg & 4)) continue;
        const uint32_t gf = blockIdx.x * G + f;
        const clx_frame_desc d = descs[gf];
        const uint32_t bs = d.block_size, nch = d.n_channels, ca = d.channel_assignment;
        const int32_t* fbuf = s_buf + (size_t)f * frame_stride;
        int32_t* o = out + d.out_offset;
        const SubParams* sp = &s_par[f * CH];
        const bool vec = ((bs & 3) == 0) && ((d.out_offset & 3) == 0);
        // Subframes predicted with the i32 accumulator: exact iff sum|coef| * max|sample| < 2^31.
        for (uint32_t c = 0; c < nch; c++) {
            const uint32_t absum = sp[c].narrow;
            if (absum == 0) continue;  // warp-uniform (shared memory broadcast)
            uint32_t m = 0;
            const int32_t* cb = fbuf + (size_t)c * bs;
            for (uint32_t t = threadIdx.x; t < bs; t += blockDim.x) {
                const int32_t v = cb[t];
                m = max(m, (uint32_t)(v < 0 ? 0u - (uint32_t)v : (uint32_t)v));
            }
            m = __reduce_max_sync(0xffffffffu, m);
            if ((unsigned long long)absum * m >= (1ull << 31) && lane == 0) {
                results[gf].status = CLX_INTERNAL_NEED_GENERIC;  // benign race: every writer stores the same value
                *need_generic = 1;
            }
        }
        if (ca >= 8) {
            const uint32_t w0 = (uint32_t)sp[0].wasted, w1 = (uint32_t)sp[1].wasted;
            if (vec) {
                for (uint32_t t = threadIdx.x * 4; t < bs; t += blockDim.x * 4) {
                    int4 a = *reinterpret_cast<const int4*>(fbuf + t);
                    int4 b = *reinterpret_cast<const int4*>(fbuf + bs + t);
                    int4 x, y;
                    decor(ca, (int32_t)((uint32_t)a.x << w0), (int32_t)((uint32_t)b.x << w1), x.x, y.x);
                    decor(ca, (int32_t)((uint32_t)a.y << w0), (int32_t)((uint32_t)b.y << w1), x.y, y.y);
                    decor(ca, (int32_t)((uint32_t)a.z << w0), (int32_t)((uint32_t)b.z << w1), x.z, y.z);
                    decor(ca, (int32_t)((uint32_t)a.w << w0), (int32_t)((uint32_t)b.w << w1), x.w, y.w);
                    *reinterpret_cast<int4*>(o + t) = x;
                    *reinterpret_cast<int4*>(o + bs + t) = y;
                }
            } else {
                for (uint32_t t = threadIdx.x; t < bs; t += blockDim.x) {
                    int32_t x, y;
                    decor(ca, (int32_t)((uint32_t)fbuf[t] << w0), (int32_t)((uint32_t)fbuf[bs + t] << w1), x, y);
                    o[t] = x;
                    o[bs + t] = y;
                }
            }
        } else {
            for (uint32_t c = 0; c < nch; c++) {
                const uint32_t wst = (uint32_t)sp[c].wasted;
                const int32_t* cb = fbuf + (size_t)c * bs;
                int32_t* oc = o + (size_t)c * bs;
                if (vec) {
                    for (uint32_t t = threadIdx.x * 4; t < bs; t += blockDim.x * 4) {
                        int4 a = *reinterpret_cast<const int4*>(cb + t);
                        a.x = (int32_t)((uint32_t)a.x << wst); a.y = (int32_t)((uint32_t)a.y << wst);
                        a.z = (int32_t)((uint32_t)a.z << wst); a.w = (int32_t)((uint32_t)a.w << wst);
                        *reinterpret_cast<int4*>(oc + t) = a;
                    }
                } else {
                    for (uint32_t t = threadIdx.x; t < bs; t += blockDim.x) oc[t] = (int32_t)((uint32_t)cb[t] << wst);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// launch helper: returns false when the batch cannot use this kernel (G would be 0)
// ---------------------------------------------------------------------------------
bool coop_plan(uint32_t max_frame_elems, uint32_t max_channels, uint32_t n_frames, int sm_count, size_t smem_budget,
               CoopPlan* plan) {
    plan->G = 0;
    if (max_frame_elems == 0 || n_frames == 0 || max_channels == 0 || max_channels > COOP_MAX_CH) return false;
    const uint32_t stride = ((max_frame_elems + 3) & ~3u) + 4;  // +4 words: frames land on different banks
    const size_t per_frame = (size_t)stride * 4 + max_channels * sizeof(SubParams) + sizeof(GroupHeader);
    uint32_t g = (uint32_t)std::min<size_t>(COOP_MAX_G, smem_budget / per_frame);
    if (g == 0) return false;
    // fill the machine in as few waves as possible, then prefer small groups (more CTAs in flight)
    const uint32_t want = (n_frames + (uint32_t)sm_count - 1) / (uint32_t)sm_count;
    if (want < g) g = std::max<uint32_t>(1, want);
    plan->G = g;
    plan->frame_stride = stride;
    plan->channels = max_channels;
    plan->smem_bytes = (size_t)g * per_frame;
    return true;
}

cudaError_t launch_coop(const uint8_t* d_bytes, uint64_t buf_bytes, const clx_frame_desc* d_descs, uint32_t n_frames,
                        int32_t* d_out, clx_frame_result* d_results, int* d_need_generic, const CoopPlan& plan,
                        cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(decode_frames_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    static const uint32_t dbg = getenv("CLX_COOP_DEBUG") ? (uint32_t)atoi(getenv("CLX_COOP_DEBUG")) : 0u;
    dim3 grid((n_frames + plan.G - 1) / plan.G), block(32 * plan.G);
    decode_frames_coop_kernel<<<grid, block, plan.smem_bytes, stream>>>(d_bytes, buf_bytes, d_descs, n_frames, d_out,
                                                                        d_results, d_need_generic, plan.G,
                                                                        plan.frame_stride, plan.channels, dbg);
    return cudaGetLastError();
}

}  // namespace clx
