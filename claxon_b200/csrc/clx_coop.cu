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
// Warp-wide bit window: 256 consecutive big-endian words of the frame, 16-byte aligned.
// Lane l holds words 4l..4l+3 of the current 128-word window in X (byte-swapped, ready for bit
// arithmetic) and the same slice of the NEXT window in Y, raw as loaded: Y is fetched one window
// ahead with a single coalesced 16-byte load per lane and nothing touches it until it slides into
// X, so HBM/L2 latency stays off the critical path.  All field extraction is shuffle + funnel shift.
// ---------------------------------------------------------------------------------
__device__ unsigned long long g_coop_stats[16];  // debug counters
constexpr uint32_t WPL = 4;            // words per lane
constexpr uint32_t WIN_WORDS = 32 * WPL;

struct Win {
    const uint4* base;  // 16-byte aligned global address at or before the frame's first byte
    uint32_t qlim;      // first 16-byte group index that lies outside the byte buffer (reads give 0)
    uint32_t b0;        // word index (multiple of 4) of X[0] of lane 0
    uint32_t X[WPL], Y[WPL];
    uint32_t F[WPL];    // words loaded by the previous slide, not yet merged into Y (pending != 0)
    uint32_t pending;
};

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
__device__ __forceinline__ uint4 win_ldg(const Win& w, uint32_t quad) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (quad < w.qlim) v = __ldg(w.base + quad);
    return v;
}
__device__ __forceinline__ void win_prime(Win& w, uint32_t bitpos, uint32_t lane) {
    w.b0 = (bitpos >> 5) & ~3u;
    const uint4 x = win_ldg(w, (w.b0 >> 2) + lane), y = win_ldg(w, (w.b0 >> 2) + 32 + lane);
    w.X[0] = bswap32(x.x); w.X[1] = bswap32(x.y); w.X[2] = bswap32(x.z); w.X[3] = bswap32(x.w);
    w.Y[0] = y.x; w.Y[1] = y.y; w.Y[2] = y.z; w.Y[3] = y.w;
    w.pending = 0;
}
// Slides the window so that the 16-byte group containing `bitpos` is lane 0's.
__device__ __forceinline__ void win_advance(Win& w, uint32_t bitpos, uint32_t lane) {
    const uint32_t nb0 = (bitpos >> 5) & ~3u;
    const uint32_t d = (nb0 - w.b0) >> 2;  // lanes to shift by
    if (d == 0) return;
    if (d > 32) { win_prime(w, bitpos, lane); return; }
    // The words fetched by the previous slide are only now folded into Y: a register written by a
    // load is not touched until a whole window later, so the load's latency is never waited for.
    if (w.pending) {
#pragma unroll
        for (uint32_t j = 0; j < WPL; j++) w.Y[j] = w.F[j];
    }
    const uint32_t src = (lane + d) & 31;
    const bool low = lane + d < 32;
    w.b0 = nb0;
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) {
        const uint32_t xs = __shfl_sync(0xffffffffu, w.X[j], src);
        const uint32_t ys = __shfl_sync(0xffffffffu, w.Y[j], src);
        w.X[j] = low ? xs : bswap32(ys);
        w.Y[j] = ys;  // lanes with !low get their real Y from F at the next slide
    }
    w.pending = low ? 0u : 1u;
    if (!low) {
        const uint4 fresh = win_ldg(w, (nb0 >> 2) + 32 + lane);
        w.F[0] = fresh.x; w.F[1] = fresh.y; w.F[2] = fresh.z; w.F[3] = fresh.w;
    }
}
// Word `i` (0 .. WIN_WORDS-1, uniform across the warp) of the window.
__device__ __forceinline__ uint32_t win_word(const Win& w, uint32_t i) {
    const uint32_t j = i & 3;
    const uint32_t mine = j == 0 ? w.X[0] : j == 1 ? w.X[1] : j == 2 ? w.X[2] : w.X[3];
    return __shfl_sync(0xffffffffu, mine, (i >> 2) & 31);
}
// 32 bits starting at `bitpos` (uniform); requires the word of bitpos to be at most WIN_WORDS-2.
__device__ __forceinline__ uint32_t win_peek32(const Win& w, uint32_t bitpos) {
    const uint32_t i = (bitpos >> 5) - w.b0;
    return __funnelshift_l(win_word(w, i + 1), win_word(w, i), bitpos & 31);
}
__device__ __forceinline__ uint32_t top_bits(uint32_t v, uint32_t n) {  // n in [0,32]
    return n >= 32 ? v : __funnelshift_l(v, 0, n);
}
// Per-lane variant: every lane asks for its own bit position inside the window.
__device__ __forceinline__ uint32_t win_peek32_lane(const Win& w, uint32_t bitpos) {
    const uint32_t i = (bitpos >> 5) - w.b0, i1 = i + 1;
    uint32_t a[WPL], b[WPL];
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) {
        a[j] = __shfl_sync(0xffffffffu, w.X[j], (i >> 2) & 31);
        b[j] = __shfl_sync(0xffffffffu, w.X[j], (i1 >> 2) & 31);
    }
    const uint32_t w0 = (i & 3) == 0 ? a[0] : (i & 3) == 1 ? a[1] : (i & 3) == 2 ? a[2] : a[3];
    const uint32_t w1 = (i1 & 3) == 0 ? b[0] : (i1 & 3) == 1 ? b[1] : (i1 & 3) == 2 ? b[2] : b[3];
    return __funnelshift_l(w1, w0, bitpos & 31);
}
__device__ __forceinline__ int32_t sext(uint32_t v, uint32_t bits) {
    return ((int32_t)(v << (32 - bits))) >> (32 - bits);
}

// ---------------------------------------------------------------------------------
// Rice decode of one window.  Vocabulary: a *search* looks for the next unary terminator (a 1 bit)
// from some bit offset; after a terminator at t the next search starts at t+1+k.  The *phase* of a
// 32-bit word is the offset at which the first search inside it starts.  A word walked from phase e
// yields its terminator mask and its exit phase (the next word's entry phase).
//
// Every lane owns WPL consecutive words and walks them simultaneously (independent dependency
// chains, interleaved by the unrolled loop).  Phases are first speculated — every word assumes its
// left neighbour exits as it would from phase 0 — and then corrected by a fix-point: a word whose
// entry phase changed is re-walked, but only until it meets a terminator of its own phase-0 chain,
// from where both chains coincide.
// ---------------------------------------------------------------------------------
#ifdef CLX_COOP_STATS
#define COOP_STAT(i, v) do { const unsigned long long sv_ = (unsigned long long)(v); if (lane == 0) atomicAdd(&g_coop_stats[i], sv_); } while (0)
#define COOP_CLOCK() clock64()
#else
#define COOP_STAT(i, v) do { (void)(v); } while (0)
#define COOP_CLOCK() 0ll
#endif

struct Walk {
    uint32_t tm[WPL];  // terminator masks (bit 31 = first bit of the word)
    uint32_t x[WPL];   // exit phases
};

// select without control flow: the optimiser otherwise turns chains of ?: back into branches, which
// serialises the WPL interleaved walks of a lane.
__device__ __forceinline__ uint32_t sel32(bool c, uint32_t a, uint32_t b) {
    uint32_t r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.u32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"((uint32_t)c));
    return r;
}

// One code of one word, branch-free (everything is a select, so the WPL words of a lane interleave
// in the instruction stream and their dependency chains overlap).  `o` = current search offset,
// 64 once the word is finished.  `stm` = speculated terminator mask to merge with (0: never merge).
__device__ __forceinline__ void walk_step(uint32_t W, uint32_t k, uint32_t stm, uint32_t sx, uint32_t& o, uint32_t& tm,
                                          uint32_t& x) {
    const bool active = o < 32;
    const uint32_t m = W & __funnelshift_rc(0xffffffffu, 0u, o);  // o >= 32 -> 0
    const uint32_t t = __clz(m);                                   // 32 when nothing is left
    const uint32_t bit = __funnelshift_rc(0x80000000u, 0u, t);     // 0 when nothing is left
    const bool hit = bit != 0;
    const bool mrg = (stm & bit) != 0;
    const uint32_t no = t + 1 + k;
    const bool spill = no >= 32;
    tm |= sel32(mrg, stm & (bit | (bit - 1)), bit);
    x = sel32(mrg, sx, sel32(hit, sel32(spill, no - 32, x), sel32(active, 0u, x)));
    o = sel32(active && hit && !mrg && !spill, no, 64u);
}

// Walks the words flagged in `todo` from phases e[]; when `merge`, a walk stops at the first
// terminator shared with the speculated chain.  Results for unflagged words are left untouched.
__device__ __forceinline__ uint32_t walk_words(const uint32_t (&W)[WPL], const uint32_t (&e)[WPL], uint32_t k, uint32_t todo,
                                           bool merge, const Walk& spec, Walk& out) {
    uint32_t trips = 0;
    uint32_t o[WPL];
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) {
        const bool go = (todo >> j) & 1u;
        o[j] = sel32(go, e[j], 64u);
        out.tm[j] = sel32(go, 0u, out.tm[j]);
        out.x[j] = sel32(go, 0u, out.x[j]);
    }
    for (;;) {
#pragma unroll
        for (uint32_t j = 0; j < WPL; j++)
            walk_step(W[j], k, merge ? spec.tm[j] : 0u, spec.x[j], o[j], out.tm[j], out.x[j]);
        bool more = false;
#pragma unroll
        for (uint32_t j = 0; j < WPL; j++) more |= o[j] < 32;
        trips++;
        if (!__any_sync(0xffffffffu, more)) break;
    }
    return trips;
}

// Decodes up to `n_rem` Rice codes with parameter k starting at bit `P` from the current window
// into out[0..); returns the number decoded (0 = cannot make progress here) and advances P to
// the end of the last decoded code.  The very last word of the window is never owned (it only
// lends its bits as the right-hand neighbour), so Y is not read here.
__device__ __forceinline__ uint32_t rice_window(const Win& w, uint32_t& P, uint32_t k, uint32_t n_rem, int32_t* out,
                                                uint32_t lane) {
    const uint32_t s = P - (w.b0 << 5);      // < 128: where the first search starts
    const uint32_t first = s >> 5;           // virtual word that contains it (lane 0)
    uint32_t W[WPL], WN[WPL];
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) W[j] = w.X[j];
    const uint32_t right = __shfl_down_sync(0xffffffffu, w.X[0], 1);
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) WN[j] = j + 1 < WPL ? w.X[j + 1] : right;
    // words that take part: from `first` up to the last-but-one word of the window
    uint32_t live = (1u << WPL) - 1u;
    if (lane == 0) live &= ~((1u << first) - 1u);
    if (lane == 31) live &= ~(1u << (WPL - 1));
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++)
        if (!(live & (1u << j))) W[j] = 0;  // a dead word has no terminators and exits with phase 0

    long long tc0 = COOP_CLOCK();
    // 1. speculated chains: phase 0 everywhere
    Walk spec, cur;
    const uint32_t zero[WPL] = {0, 0, 0, 0};
    COOP_STAT(0, 1);
    COOP_STAT(2, walk_words(W, zero, k, live, false, spec, spec));
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++)
        if (!(live & (1u << j))) { spec.tm[j] = 0; spec.x[j] = 0; }
    cur = spec;
    long long tc1 = COOP_CLOCK(); COOP_STAT(5, tc1 - tc0);
    // 2. fix-point on the entry phases
    uint32_t e[WPL] = {0, 0, 0, 0};
    for (;;) {
        const uint32_t left = __shfl_up_sync(0xffffffffu, cur.x[WPL - 1], 1);
        uint32_t todo = 0;
#pragma unroll
        for (uint32_t j = 0; j < WPL; j++) {
            uint32_t want = j == 0 ? left : cur.x[j - 1];
            if (lane == 0 && j <= first) want = j == first ? (s & 31) : 0;
            if (!(live & (1u << j))) want = 0;
            if (want != e[j]) { e[j] = want; todo |= 1u << j; }
        }
        todo &= live;
        if (!__any_sync(0xffffffffu, todo != 0)) break;
        COOP_STAT(1, 1);
        COOP_STAT(3, walk_words(W, e, k, todo, true, spec, cur));
    }
    long long tc2 = COOP_CLOCK(); COOP_STAT(6, tc2 - tc1);
    // 3. ranks
    uint32_t cnt[WPL], lane_cnt = 0;
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) { cnt[j] = __popc(cur.tm[j]); lane_cnt += cnt[j]; }
    uint32_t incl = lane_cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t rank[WPL];
    rank[0] = incl - lane_cnt;
#pragma unroll
    for (uint32_t j = 1; j < WPL; j++) rank[j] = rank[j - 1] + cnt[j - 1];
    if (total > n_rem) {  // the partition ends inside this window: keep the first n_rem codes
#pragma unroll
        for (uint32_t j = 0; j < WPL; j++) {
            if (rank[j] >= n_rem) cur.tm[j] = 0;
            else if (rank[j] + cnt[j] > n_rem) {
                uint32_t keep = 0, tmp = cur.tm[j];
                for (uint32_t i = 0; i < n_rem - rank[j]; i++) {
                    const uint32_t bit = 0x80000000u >> __clz(tmp);
                    keep |= bit;
                    tmp &= ~bit;
                }
                cur.tm[j] = keep;
            }
        }
        total = n_rem;
    }
    if (total == 0) return 0;
    // 4. end of the last code at or before each word (window-relative bit offsets)
    uint32_t wend[WPL], lane_end = 0;
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) {
        wend[j] = cur.tm[j] ? ((lane * WPL + j) << 5) + (32 - __ffs(cur.tm[j])) + 1 + k : 0;
        lane_end = max(lane_end, wend[j]);
    }
    uint32_t endi = lane_end;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, endi, d);
        if (lane >= (uint32_t)d) endi = max(endi, v);
    }
    uint32_t before = __shfl_up_sync(0xffffffffu, endi, 1);
    if (lane == 0) before = 0;
    const uint32_t new_end = __shfl_sync(0xffffffffu, endi, 31);
    uint32_t start[WPL];
    start[0] = max(before, s);
#pragma unroll
    for (uint32_t j = 1; j < WPL; j++) start[j] = max(start[j - 1], wend[j - 1]);
    long long tc3 = COOP_CLOCK(); COOP_STAT(7, tc3 - tc2);
    // 5. emit: the WPL words of a lane advance together, one code each per trip (predicated)
    uint32_t rest[WPL];
#pragma unroll
    for (uint32_t j = 0; j < WPL; j++) rest[j] = cur.tm[j];
    for (;;) {
        bool more = false;
#pragma unroll
        for (uint32_t j = 0; j < WPL; j++) {
            const bool have = rest[j] != 0;
            const uint32_t t = __clz(rest[j]) & 31;
            rest[j] &= ~__funnelshift_rc(0x80000000u, 0u, sel32(have, t, 32u));
            const uint32_t pos = ((lane * WPL + j) << 5) + t;
            const uint32_t q = pos - start[j];
            const uint32_t hi = __funnelshift_lc(WN[j], W[j], t + 1);
            const uint32_t r = __funnelshift_l(hi, 0, k);
            const uint32_t u = (q << k) | r;
            const uint32_t val = (u >> 1) ^ (0u - (u & 1u));
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.u32 [%0], %1;\n\t}" ::"l"(out + rank[j]), "r"(val),
                         "r"((uint32_t)have)
                         : "memory");
            rank[j] += (uint32_t)have;
            start[j] = sel32(have, pos + 1 + k, start[j]);
            more |= rest[j] != 0;
        }
        if (!__any_sync(0xffffffffu, more)) break;
    }
    COOP_STAT(4, total);
    COOP_STAT(8, COOP_CLOCK() - tc3);
    P = (w.b0 << 5) + new_end;
    return total;
}

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
// Shared-space accessors (LDS/STS with a 32-bit address).  `volatile` keeps their order among
// themselves and relative to barriers; no memory clobber, so arithmetic schedules freely around them.
__device__ __forceinline__ int32_t lds32(uint32_t addr) {
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, int32_t v) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v));
}

// `sbuf` is the subframe's buffer as a 32-bit shared-space address.
template <int TAPS, int U, typename ACC>
__device__ __forceinline__ void predict_inplace(uint32_t sbuf, uint32_t bs, uint32_t order, uint32_t shift,
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
            int32_t val = inside ? lds32(sbuf + 4 * t) : 0;
            if (t >= order) {
                long long acc = 0;
#pragma unroll
                for (int j = 0; j < TAPS; j++) acc += (long long)c[j] * (long long)h[j];
                val += sizeof(ACC) == 8 ? (int32_t)(acc >> shift) : (int32_t)((int32_t)acc >> shift);
                if (inside) sts32(sbuf + 4 * t, val);
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
        for (int i = 0; i < U; i++) rn[i] = lds32(sbuf + 4 * (head_end + i));
        for (uint32_t t = head_end; t < bulk_end; t += U) {
            int32_t r[U];
#pragma unroll
            for (int i = 0; i < U; i++) r[i] = rn[i];
            // prefetch of the next trip (reads up to U samples past the bulk on the last trip: still
            // inside the subframe buffer or the next one, never used)
#pragma unroll
            for (int i = 0; i < U; i++) rn[i] = lds32(sbuf + 4 * (t + U + i));
            predict_trip<TAPS, U, ACC>(v, c, r, shift);
            if (active) {
#pragma unroll
                for (int i = 0; i < U; i++) sts32(sbuf + 4 * (t + i), v[TAPS + i]);
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
    long long tp0 = COOP_CLOCK();
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
        const uint64_t aligned = d.byte_offset & ~15ull;
        w.base = reinterpret_cast<const uint4*>(bytes + aligned);
        w.qlim = (uint32_t)min((buf_bytes - aligned) >> 4, (uint64_t)0x7ffffffu);
        const uint32_t bit0 = (uint32_t)(d.byte_offset & 15) * 8;
        const uint32_t limit = bit0 + d.byte_len * 8;
        uint32_t P = bit0 + (uint32_t)d.header_len * 8;
        if (ok) win_prime(w, P, lane);
        else { w.b0 = 0; w.pending = 0; for (uint32_t j = 0; j < WPL; j++) { w.X[j] = 0; w.Y[j] = 0; w.F[j] = 0; } }

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
                    long long ta = COOP_CLOCK();
                    win_advance(w, P, lane);
                    COOP_STAT(9, COOP_CLOCK() - ta);
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
    COOP_STAT(10, COOP_CLOCK() - tp0);
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
                        // valid streams keep |sample| <= 2^(bits-1); anything beyond is caught by the phase-3 check
                        narrow_ok = ((unsigned long long)absum << (bits - 1)) < (1ull << 31);
                    }
                }
            }
            const uint32_t max_order = __reduce_max_sync(0xffffffffu, active ? order : 0u);
            if (max_order == 0) continue;
            // idle lanes read (never write) some valid buffer so that the bulk loop needs no guards
            const uint32_t some = __ffs(__ballot_sync(0xffffffffu, active)) - 1;
            uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sbuf);
            const uint32_t alias = __shfl_sync(0xffffffffu, saddr, some);
            if (!active) saddr = alias;
            long long tl0 = COOP_CLOCK();
            // i32 accumulator where sum|coef| * 2^sample_bits leaves headroom in i32 for every lane
            const bool all_narrow = __all_sync(0xffffffffu, !active || narrow_ok);
            if (active && lane_sp != nullptr) lane_sp->narrow = all_narrow ? absum : 0u;
            if (all_narrow) {
                if (max_order <= 4) predict_inplace<4, 4, int>(saddr, bs, order, shift, coefs, active);
                else if (max_order <= 8) predict_inplace<8, 8, int>(saddr, bs, order, shift, coefs, active);
                else if (max_order <= 12) predict_inplace<12, 4, int>(saddr, bs, order, shift, coefs, active);
                else predict_inplace<32, 4, int>(saddr, bs, order, shift, coefs, active);
            } else {
                if (max_order <= 4) predict_inplace<4, 4, long long>(saddr, bs, order, shift, coefs, active);
                else if (max_order <= 8) predict_inplace<8, 8, long long>(saddr, bs, order, shift, coefs, active);
                else if (max_order <= 12) predict_inplace<12, 4, long long>(saddr, bs, order, shift, coefs, active);
                else predict_inplace<32, 4, long long>(saddr, bs, order, shift, coefs, active);
            }
            COOP_STAT(11, COOP_CLOCK() - tl0);
            COOP_STAT(12, all_narrow ? 1 : 0);
            COOP_STAT(13, 1);
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

extern "C" void clx_debug_coop_stats(unsigned long long* out8, int reset) {
    cudaMemcpyFromSymbol(out8, clx::g_coop_stats, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(clx::g_coop_stats, z, sizeof z); }
}
