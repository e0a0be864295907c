/* claxon_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the per-frame decode semantics of the reference
 * decoder ruuda/claxon v0.4.3.  It is the checker the CUDA path is compared
 * against; nothing in the product library (claxon_b200/) links, imports or
 * calls it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may use it.
 *
 * Parity pin: this restatement reproduces every in-source known-answer vector
 * of the reference's own unit tests on this path and the STREAMINFO MD5 of every
 * bundled fixture that carries one (tests/test_oracle_golden.py).  The reference
 * itself is Rust and cannot be compiled in this image (no rustc), so there is no
 * oracle/_ref build.
 *
 * Every function cites the reference file:line whose behaviour it restates
 * (paths relative to the reference checkout).
 */
#ifndef CLAXON_ORACLE_H
#define CLAXON_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/clx_status.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Frame header as parsed by src/frame.rs:131-316. */
typedef struct clxo_frame_header {
    uint32_t block_size;         /* 1..65535 */
    uint32_t sample_rate;        /* 0 = "get from streaminfo" */
    uint32_t n_channels;         /* 1..8 */
    uint32_t channel_assignment; /* raw 4-bit code: 0..7 independent, 8 L/S, 9 R/S, 10 M/S */
    uint32_t bits_per_sample;    /* 0 = absent from header */
    uint32_t variable_blocking;  /* 0 fixed (number = frame #), 1 variable (number = sample #) */
    uint64_t number;             /* frame or sample number */
    uint32_t header_len;         /* bytes, incl. the CRC-8 */
} clxo_frame_header;

/* Result of decoding one frame (src/frame.rs:667-779). */
typedef struct clxo_frame_info {
    clxo_frame_header header;
    uint64_t time;          /* Block::time(): src/frame.rs:771-774 */
    uint64_t consumed;      /* bytes from sync through CRC-16 when status==CLX_OK */
    uint32_t n_samples;     /* n_channels * block_size (Block::len()) */
} clxo_frame_info;

/* STREAMINFO (src/metadata.rs:321-400). */
typedef struct clxo_streaminfo {
    uint32_t min_block_size, max_block_size;
    uint32_t min_frame_size, max_frame_size; /* 0 = unknown */
    uint32_t sample_rate, channels, bits_per_sample;
    uint64_t samples;                        /* 0 = unknown */
    uint8_t md5sum[16];
} clxo_streaminfo;

/* --- primitives with known-answer vectors in the reference's unit tests --- */
int16_t clxo_extend_sign_u16(uint16_t val, uint32_t bits);          /* src/subframe.rs:96-101 */
int32_t clxo_extend_sign_u32(uint32_t val, uint32_t bits);          /* src/subframe.rs:117-122 */
int32_t clxo_rice_to_signed(uint32_t val);                          /* src/subframe.rs:157-170 */
void clxo_predict_fixed(uint32_t order, int32_t* buf, size_t n);    /* src/subframe.rs:417-474 */
void clxo_predict_lpc(const int16_t* coefs, uint32_t order, uint32_t shift,
                      int32_t* buf, size_t n);                      /* src/subframe.rs:524-614 */
void clxo_decode_left_side(int32_t* buf, size_t n_total);           /* src/frame.rs:319-334 */
void clxo_decode_right_side(int32_t* buf, size_t n_total);          /* src/frame.rs:345-360 */
void clxo_decode_mid_side(int32_t* buf, size_t n_total);            /* src/frame.rs:371-389 */
uint8_t clxo_crc8(const uint8_t* p, size_t n);                      /* src/crc.rs:13-31, :90-92 */
uint16_t clxo_crc16(const uint8_t* p, size_t n);                    /* src/crc.rs:33-57, :110-112 */
/* Returns status; *consumed = bytes read. src/frame.rs:64-105 */
int clxo_read_var_length_int(const uint8_t* p, size_t n, uint64_t* value, size_t* consumed);
/* MSB-first bit-field reads at a bit offset (src/input.rs:415-643 semantics).
 * kind: 0 = read `bits` bits (<=32), 1 = read_unary.  Returns status. */
int clxo_bit_read(const uint8_t* p, size_t n, uint64_t* bitpos, int kind, uint32_t bits,
                  uint32_t* value);

/* --- frame level --- */
/* Parses a frame header at p[0..n).  CLX_EOF when fewer than 2 bytes remain. */
int clxo_read_frame_header(const uint8_t* p, size_t n, clxo_frame_header* h);

/* Decodes subframes [0, n_channels) of a frame whose header has been parsed; `p`
 * points at the frame's first (sync) byte.  Exposed so tests can address single
 * subframes: bit_pos in/out is relative to p. */
int clxo_decode_subframe(const uint8_t* p, size_t n, uint64_t* bit_pos, uint32_t bps,
                         int32_t* out, uint32_t block_size);

/* FrameReader::read_next_or_eof on an in-memory byte span: decodes the frame that
 * starts at p[0].  `out` must hold n_channels*block_size i32 (out_cap elements);
 * if it is too small the function returns CLX_ERR_INVALID_ARGUMENT after filling
 * info->header so that the caller can retry.  verify_crc=0 mimics cfg(fuzzing). */
int clxo_decode_frame(const uint8_t* p, size_t n, int32_t* out, size_t out_cap,
                      clxo_frame_info* info, int verify_crc);

/* --- stream level (host side of FlacReader::new: src/lib.rs:186-307) --- */
/* Walks 'fLaC' + metadata blocks; returns the first-frame byte offset.
 * read_vorbis_comment mirrors FlacReaderOptions (vorbis blocks are validated the
 * way src/metadata.rs:402-513 does; contents are not returned). */
int clxo_open_stream(const uint8_t* p, size_t n, clxo_streaminfo* si, uint64_t* first_frame);

/* Decodes a whole stream's frames sequentially (the claxon `blocks()` loop).
 * Writes interleaved?=0 planar-per-frame samples back to back into out.
 * Returns the status of the first failing frame or CLX_OK at clean EOF.
 * n_frames / n_samples_total are filled either way. */
int clxo_decode_stream(const uint8_t* p, size_t n, uint64_t first_frame, int32_t* out,
                       size_t out_cap, uint64_t* n_frames, uint64_t* n_samples_total,
                       int verify_crc);

/* Multi-threaded batch decode used by the CPU baseline: frames are given as
 * (offset, length) pairs, statically sharded over n_threads threads, each an
 * independent single-threaded decoder (claxon itself has no threading).
 * out_offsets[i] = element offset of frame i in `out`.  Returns number of frames
 * whose status != CLX_OK (statuses written per frame when `statuses` != NULL). */
int clxo_decode_batch_mt(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* lengths,
                         size_t n_frames, int32_t* out, const uint64_t* out_offsets,
                         int32_t* statuses, int n_threads, int verify_crc);

#ifdef __cplusplus
}
#endif
#endif
