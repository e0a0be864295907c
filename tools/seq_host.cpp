// seq_host.cpp — host harness for the per-lane entropy logic of the sequential path
// (claxon_b200/csrc/clx_seq_lane.h).  Test infrastructure only (tests/test_seq_host.py): it runs the
// very code the CUDA kernel runs per lane, with plain loads/stores instead of the shared-memory ring
// and the scratch stores, then applies a scalar restatement of the prediction kernel's arithmetic
// so the result can be compared with known PCM.  Build: g++ -O2 -shared -fPIC -Iinclude.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../claxon_b200/csrc/clx_seq_lane.h"

namespace {

struct HostIO {
    const uint8_t* base = nullptr;  // frame's 16-byte aligned base
    uint64_t avail = 0;             // bytes readable from base
    uint8_t* column = nullptr;      // lane's column of the current channel's rows
    uint8_t* frame_rows = nullptr;  // (w, 0) rows + lane * 16
    uint64_t channel_stride = 0;
    bool narrow = false;
    uint32_t word(uint32_t wi) const {
        uint8_t b[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            const uint64_t p = (uint64_t)wi * 4 + i;
            if (p < avail) b[i] = base[p];
        }
        return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3];
    }
    uint32_t wnext = 0;
    void seek_next(uint32_t wi) { wnext = wi; }
    uint32_t next_raw() { return __builtin_bswap32(word(wnext++)); }
    void ensure(uint32_t) {}
    bool prefetch_group(uint32_t) { return true; }
    void select_channel(uint32_t ch) { column = frame_rows + ch * channel_stride; }
    void store1(uint32_t t, int32_t e) {
        if (narrow) { int16_t v = (int16_t)e; memcpy(column + clx::seq_elem_offset<true>(t), &v, 2); }
        else memcpy(column + clx::seq_elem_offset<false>(t), &e, 4);
    }
    void store8(uint32_t t, const int32_t (&e)[8]) {
        for (uint32_t i = 0; i < 8; i++) store1(t + i, e[i]);
    }
};

template <bool NARROW>
void run(const uint8_t* bytes, uint64_t nbytes, const clx_frame_desc* descs, uint32_t n, uint32_t CH, uint32_t max_bs,
         int32_t* out, clx_frame_result* results, uint64_t* stats) {
    const uint32_t n_warps = (n + 31) / 32;
    const uint64_t rows = clx::seq_rows_for<NARROW>(max_bs);
    std::vector<uint8_t> scratch((size_t)n_warps * CH * rows * clx::SEQ_ROW_BYTES, 0xCD);
    std::vector<clx::SeqParams> params((size_t)n_warps * 32 * CH);
    memset(params.data(), 0xCD, params.size() * sizeof(clx::SeqParams));
    for (uint32_t f = 0; f < n; f++) {
        const clx_frame_desc& d = descs[f];
        clx::SeqLane<HostIO, NARROW> L;
        const uint64_t aligned = d.byte_offset & ~15ull;
        L.io.base = bytes + aligned;
        L.io.avail = nbytes - aligned;
        L.io.narrow = NARROW;
        L.io.channel_stride = rows * clx::SEQ_ROW_BYTES;
        L.io.frame_rows = scratch.data() + (uint64_t)(f / 32) * CH * rows * clx::SEQ_ROW_BYTES + (f % 32) * 16;
        L.io.column = L.io.frame_rows;
        L.init(d, params.data() + (size_t)f * CH, CH);
        uint64_t fast = 0, slow = 0;
        while (!L.done()) {
            if (L.fast_ready()) { L.fast_group(); fast++; }
            else { L.slow_step(); slow++; }
        }
        if (stats) { stats[0] += fast; stats[1] += slow; }
        results[f].status = L.ok ? 0 : -2;
        results[f].consumed = L.consumed;
        if (!L.ok) continue;
        // scalar restatement of the prediction kernel (i64 accumulate, src/subframe.rs:524-614; wasted
        // shift :216-225; decorrelation src/frame.rs:319-389), reading the scratch the way its lanes do
        const uint32_t bs = d.block_size;
        int32_t* fo = out + d.out_offset;
        for (uint32_t c = 0; c < d.n_channels; c++) {
            const clx::SeqParams& sp = params[(size_t)f * CH + c];
            const uint8_t* col = L.io.frame_rows + c * L.io.channel_stride;
            int32_t* s = fo + (size_t)c * bs;
            for (uint32_t t = 0; t < bs; t++) {
                if (t < (uint32_t)sp.order) { s[t] = sp.warm[t]; continue; }
                int32_t r;
                if (NARROW) { int16_t v; memcpy(&v, col + clx::seq_elem_offset<true>(t), 2); r = v; }
                else memcpy(&r, col + clx::seq_elem_offset<false>(t), 4);
                long long acc = 0;
                for (int j = 0; j < sp.order; j++) acc += (long long)sp.coefs[j] * (long long)s[t - 1 - j];
                s[t] = (int32_t)((uint32_t)(int32_t)(acc >> sp.shift) + (uint32_t)r);
            }
            for (uint32_t t = 0; t < bs; t++) s[t] = (int32_t)((uint32_t)s[t] << sp.wasted);
        }
        if (d.channel_assignment >= 8) {
            int32_t *a = fo, *b = fo + bs;
            for (uint32_t t = 0; t < bs; t++) {
                const uint32_t x = (uint32_t)a[t], y = (uint32_t)b[t];
                if (d.channel_assignment == 8) b[t] = (int32_t)(x - y);
                else if (d.channel_assignment == 9) a[t] = (int32_t)(x + y);
                else {
                    const uint32_t m = (x << 1) | (y & 1u);
                    a[t] = ((int32_t)(m + y)) >> 1;
                    b[t] = ((int32_t)(m - y)) >> 1;
                }
            }
        }
    }
}

}  // namespace

extern "C" int seq_host_decode(const uint8_t* bytes, uint64_t nbytes, const clx_frame_desc* descs, uint32_t n, int narrow,
                               int32_t* out, clx_frame_result* results, uint64_t* stats) {
    uint32_t max_ch = 1, max_bs = 1;
    for (uint32_t i = 0; i < n; i++) {
        if (descs[i].n_channels > max_ch) max_ch = descs[i].n_channels;
        if (descs[i].block_size > max_bs) max_bs = descs[i].block_size;
    }
    uint32_t CH = 1;
    while (CH < max_ch) CH <<= 1;
    if (narrow) run<true>(bytes, nbytes, descs, n, CH, max_bs, out, results, stats);
    else run<false>(bytes, nbytes, descs, n, CH, max_bs, out, results, stats);
    return 0;
}
