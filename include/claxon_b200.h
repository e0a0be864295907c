/* claxon_b200.h — C ABI of the B200-native batched FLAC frame decoder.
 *
 * This is the drop-in boundary for the per-frame decode path of ruuda/claxon
 * v0.4.3.  claxon has no FFI of its own; its boundary is the Rust API
 *   FrameReader::read_next_or_eof(&mut self, Vec<i32>) -> Result<Option<Block>>
 *                                                     (reference src/frame.rs:667)
 *   FlacReader::{new, streaminfo, blocks, samples}    (src/lib.rs:217-435)
 *   Block::{time, len, duration, channels, channel, sample, into_buffer}
 *                                                     (src/frame.rs:402-529)
 * A Rust `extern "C"` shim binds exactly the entry points below:
 * the host parses frame headers (clx_parse_frame_header == src/frame.rs:131-316),
 * ships raw frame bitstreams to the device, and everything below the header parse
 * and above the CRC-16 footer check — subframe::decode (src/subframe.rs:184-228),
 * decode_residual (:236-380), predict_fixed (:417-474), predict_lpc_* (:524-614),
 * decode_{left,right,mid}_side (src/frame.rs:319-389) — runs in sm_100a kernels.
 *
 * Plain pointers and sizes only; no torch / C++ types cross this boundary.
 * There is NO CPU fallback: if no CUDA device is usable the create call fails
 * with CLX_ERR_NO_DEVICE.
 */
#ifndef CLAXON_B200_H
#define CLAXON_B200_H

#include <stddef.h>
#include <stdint.h>
#include "clx_status.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CLX_ABI_VERSION 1

/* ------------------------------------------------------------------------- */
/* Frame descriptor: the parsed frame header + where the frame's bytes are.   */
/* Mirrors claxon's private FrameHeader (src/frame.rs:41-48).                 */
/* ------------------------------------------------------------------------- */
typedef struct clx_frame_desc {
    uint64_t byte_offset;        /* offset of the frame's first sync byte in the byte buffer */
    uint32_t byte_len;           /* bytes AVAILABLE to this frame from byte_offset: the exact frame
                                    length (sync..CRC-16) when known, else an upper bound (e.g. to
                                    the end of the stream).  Reads past it are UnexpectedEof. */
    uint16_t header_len;         /* frame header bytes incl. the CRC-8 */
    uint16_t block_size;         /* 1..65535 inter-channel samples */
    uint8_t n_channels;          /* 1..8 */
    uint8_t channel_assignment;  /* raw 4-bit code: 0..7 independent(n+1), 8 L/S, 9 R/S, 10 M/S */
    uint8_t bits_per_sample;     /* 8/12/16/20/24; 0 = not in the header (-> Unsupported) */
    uint8_t flags;               /* CLX_FRAME_* */
    uint32_t sample_rate;        /* Hz, 0 = "from streaminfo"; not used by the decode */
    uint64_t number;             /* coded frame / sample number */
    uint64_t out_offset;         /* element (i32) offset of this frame's samples in the output;
                                    multiple of 4 recommended (vectorised stores).  Elements of
                                    `out` that lie between the first and the last frame of a call but
                                    belong to no frame (alignment gaps) are unspecified afterwards. */
} clx_frame_desc;

#define CLX_FRAME_VARIABLE_BLOCKING 1u /* `number` is a sample number, else a frame number */
#define CLX_FRAME_CRC16_VERIFIED 2u    /* set by clx_demux_frames: CRC-16 of the first byte_len-2 bytes
                                          already matched the footer (saves the host a second pass) */

/* Per-frame outcome. */
typedef struct clx_frame_result {
    int32_t status;      /* clx_status; CLX_OK when the frame decoded and its CRC-16 matched */
    uint32_t consumed;   /* bytes from the sync code through the CRC-16 (valid when status is
                            CLX_OK or CLX_ERR_FRAME_CRC_MISMATCH) */
} clx_frame_result;

/* STREAMINFO (src/metadata.rs:29-54). 0 = unknown for frame sizes / samples. */
typedef struct clx_streaminfo {
    uint32_t min_block_size, max_block_size;
    uint32_t min_frame_size, max_frame_size;
    uint32_t sample_rate, channels, bits_per_sample;
    uint64_t samples;
    uint8_t md5sum[16];
} clx_streaminfo;

typedef struct clx_options {
    int32_t device;        /* CUDA device ordinal */
    uint32_t flags;        /* CLX_OPT_* */
    uint32_t n_streams;    /* internal CUDA streams for host<->device pipelining (0 = default 2) */
    uint32_t host_threads; /* host threads for the CRC-16 pass (0 = default: min(32, hardware threads));
                              give each rank its share when several ranks run on one host */
} clx_options;
#define CLX_OPT_NO_VERIFY_CRC 1u /* mimic claxon's cfg(fuzzing): skip CRC-8/CRC-16 checks */
#define CLX_OPT_GENERIC_KERNEL_ONLY 2u /* testing: bypass the fast path */
#define CLX_OPT_WARP_PER_FRAME 4u      /* the warp-per-frame fast path everywhere (default: only for small synchronous calls) */
#define CLX_OPT_LANE_PER_FRAME 8u      /* the lane-per-frame fast path everywhere, also for small synchronous calls */

typedef struct clx_ctx clx_ctx;     /* one per host thread / GPU; owns device scratch */
typedef struct clx_batch clx_batch; /* a device-resident batch (bytes + descriptors + output) */

/* ------------------------------------------------------------------------- */
/* status helpers                                                             */
/* ------------------------------------------------------------------------- */
const char* clx_status_str(int status);  /* claxon's verbatim error string */
int clx_status_kind(int status);         /* clx_error_kind */
uint32_t clx_abi_version(void);

/* ------------------------------------------------------------------------- */
/* host-side parsing (no GPU needed)                                          */
/* ------------------------------------------------------------------------- */

/* read_frame_header_or_eof (src/frame.rs:131-316) on p[0..n).  Fills every field of
 * `d` except byte_offset/byte_len/out_offset.  CLX_EOF when fewer than 2 bytes remain. */
int clx_parse_frame_header(const uint8_t* p, size_t n, clx_frame_desc* d, uint32_t flags);

/* FlacReader::new (src/lib.rs:217-307, default options): checks 'fLaC', walks the
 * metadata blocks, returns STREAMINFO and the offset of the first frame. */
int clx_open_stream(const uint8_t* p, size_t n, clx_streaminfo* si, uint64_t* first_frame);
/* FlacReader::new_ext with FlacReaderOptions (src/lib.rs:123-170, :230-307).  CLX_OPEN_METADATA_ONLY: stop
 * as soon as every desired block has been read; *first_frame is then 0 and the stream cannot be decoded
 * (claxon panics in blocks()/samples()).  CLX_OPEN_NO_VORBIS_COMMENT == read_vorbis_comment: false.
 * *vc_offset / *vc_length locate the body of the (validated) VORBIS_COMMENT block, 0 / 0 when there is
 * none or it was not asked for: little-endian u32 vendor length, vendor string, u32 comment count, then
 * per comment a u32 length and "NAME=value" in UTF-8 (what FlacReader::vendor/tags/get_tag expose,
 * src/lib.rs:318-360, src/metadata.rs:134-211); zero-length comments are to be skipped. */
#define CLX_OPEN_METADATA_ONLY 1u
#define CLX_OPEN_NO_VORBIS_COMMENT 2u
int clx_open_stream_ex(const uint8_t* p, size_t n, uint32_t open_flags, clx_streaminfo* si, uint64_t* first_frame,
                       uint64_t* vc_offset, uint32_t* vc_length);

/* Frame demultiplexer: finds frame boundaries in bytes[start..n) without decoding:
 * sync code + header parse + CRC-8, then the first later sync position at which the
 * CRC-16 over the candidate span matches (or the end of the stream).  Writes up to
 * `max_frames` descriptors with exact byte_len and out_offset laid out back to back
 * (each frame aligned to 4 elements).  Stops at the first position that is not a
 * valid frame start and reports that header's status in *stop_status (CLX_EOF at a
 * clean end).  Returns the number of descriptors written. */
size_t clx_demux_frames(const uint8_t* bytes, size_t n, uint64_t start, clx_frame_desc* descs,
                        size_t max_frames, uint64_t* next_offset, uint64_t* total_out_elems,
                        int* stop_status, uint32_t flags);

/* The same on `n_threads` host threads (0 = one per hardware thread): the byte range is cut into equal parts, every
 * worker finds the frames that START in its part — a start counts once its header parses (sync code, field codes,
 * CRC-8) and a CRC-16-confirmed end lies within the size that header allows — and the parts are stitched in order,
 * a part being accepted only if the chain before it ends exactly where it begins.  Descriptor for descriptor the
 * result of clx_demux_frames; the scan (dominated by the CRC-16 of every byte) runs at n_threads times its rate. */
size_t clx_demux_frames_mt(const uint8_t* bytes, size_t n, uint64_t start, clx_frame_desc* descs,
                           size_t max_frames, uint64_t* next_offset, uint64_t* total_out_elems,
                           int* stop_status, uint32_t flags, uint32_t n_threads);

/* Container feeds: a wrapper that stores one FLAC frame per packet / sample hands over
 * the frame boundaries, so no search (clx_demux_frames) is needed.  What the reference leaves to the `ogg`
 * and `mp4parse` crates in examples/decode_ogg.rs:26-125 and examples/decode_mp4.rs:26-167.
 *
 * clx_ogg_frames: an in-memory Ogg file with the FLAC mapping (first packet: 0x7F "FLAC" version, number of
 * header packets, "fLaC", STREAMINFO; then the header packets; then one frame per packet; empty packets
 * skipped; page CRCs verified unless CLX_OPT_NO_VERIFY_CRC).  Packets may span pages, so the frames are
 * copied back to back into `frames_out` (capacity frames_cap; the file's size always suffices) and the
 * descriptors index THAT buffer; byte_len is exact.
 * clx_mp4_frames: an in-memory MP4 / ISO BMFF file; the first track with a 'fLaC' sample entry.  STREAMINFO
 * comes from its dfLa box, frame extents from stsz + stsc + stco/co64; samples are contiguous in the file, so
 * the descriptors index the file's own bytes.
 * Both fill out_offset back to back (4-element aligned) and return CLX_ERR_CONTAINER for a malformed wrapper,
 * a frame-header status if a packet / sample is not a frame, CLX_ERR_INVALID_ARGUMENT if max_frames or
 * frames_cap is too small. */
int clx_ogg_frames(const uint8_t* ogg, size_t n, clx_streaminfo* si, uint8_t* frames_out, size_t frames_cap,
                   clx_frame_desc* descs, size_t max_frames, size_t* n_frames, size_t* frames_bytes,
                   uint64_t* total_out_elems, uint32_t flags);
int clx_mp4_frames(const uint8_t* mp4, size_t n, clx_streaminfo* si, clx_frame_desc* descs, size_t max_frames,
                   size_t* n_frames, uint64_t* total_out_elems, uint32_t flags);

uint8_t clx_crc8(const uint8_t* p, size_t n);   /* src/crc.rs: poly 0x07, init 0 */
uint16_t clx_crc16(const uint8_t* p, size_t n); /* src/crc.rs: poly 0x8005, init 0 */

/* ------------------------------------------------------------------------- */
/* device path                                                                */
/* ------------------------------------------------------------------------- */
int clx_ctx_create(const clx_options* opts, clx_ctx** out);
void clx_ctx_destroy(clx_ctx* ctx);
const char* clx_ctx_last_error(const clx_ctx* ctx); /* CUDA error text for CLX_ERR_CUDA */

/* The per-frame decode path, end to end with HOST buffers (the call a binding makes):
 * copies the frame bytes to the device, runs the decode kernels, copies the planar
 * i32 PCM (Block layout: buffer[ch*block_size + i], src/frame.rs:477-481) back into
 * `out` at descs[i].out_offset, and fills results[i].  The frame CRC-16 is verified
 * (unless CLX_OPT_NO_VERIFY_CRC) after the subframes, so subframe errors pre-empt
 * "frame CRC mismatch" exactly as in src/frame.rs:752-763.  A failed frame never
 * aborts the batch; its output region is fully overwritten (never stale). */
int clx_decode_frames(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs,
                      size_t n_frames, int32_t* out, size_t out_elems, clx_frame_result* results);

/* Output stage: the same call with the samples delivered INTERLEAVED — for each
 * inter-channel sample every channel in turn, the order FlacSamples yields them in (src/lib.rs:473-519)
 * — as little-endian integers of 2, 3 or 4 bytes: what a WAV writer stores (examples/decode.rs:48-62)
 * and what the STREAMINFO MD5 is defined over (src/metadata.rs:52-53).  The conversion runs on the
 * device, so 16-bit audio crosses PCIe as 2 bytes per sample.  `out_elems` and descs[i].out_offset
 * count SAMPLES (elements of 2 / 3 / 4 bytes); a frame occupies n_channels * block_size of them as in
 * the planar layout.  CLX_OUT_INTERLEAVED_I16 / _I24 require every frame's bits_per_sample to be at
 * most 16 / 24 (else CLX_ERR_INVALID_ARGUMENT); a sample that nevertheless does not fit (only an invalid
 * stream has them) is truncated to the element size like `sample as i16` in examples/decode.rs:52.
 * CLX_OUT_PLANAR_I32 is clx_decode_frames. */
#define CLX_OUT_PLANAR_I32 0u
#define CLX_OUT_INTERLEAVED_I32 1u
#define CLX_OUT_INTERLEAVED_I16 2u
#define CLX_OUT_INTERLEAVED_I24 3u
int clx_decode_frames_to(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs,
                         size_t n_frames, void* out, size_t out_elems, clx_frame_result* results, uint32_t mode);

/* Device-resident variant: upload once, decode many times (kernel-only timing), read back. */
int clx_batch_create(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs,
                     size_t n_frames, size_t out_elems, clx_batch** out);
/* The same with `bytes` in DEVICE memory of the context's GPU (CLX_BATCH_BYTES_ON_DEVICE): e.g. a shard that
 * arrived over NVLink by the one NCCL scatter (claxon_b200/shard.py).  Nothing of it ever passes through the host:
 * the frame CRC-16 is then checked on the device as part of every decode. */
#define CLX_BATCH_BYTES_ON_DEVICE 1u
int clx_batch_create_ex(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs,
                        size_t n_frames, size_t out_elems, uint32_t batch_flags, clx_batch** out);
int clx_batch_decode(clx_ctx* ctx, clx_batch* b, uint32_t stream_index); /* async on an internal stream */
int clx_batch_sync(clx_ctx* ctx, clx_batch* b);
int clx_batch_read(clx_ctx* ctx, clx_batch* b, int32_t* out, size_t out_elems, clx_frame_result* results);
void clx_batch_destroy(clx_ctx* ctx, clx_batch* b);
/* Steady-state throughput: decodes `steps` batches back to back (step i = batches[i % n_batches]
 * on internal stream i % n_streams, so several batches are in flight) and returns the device time
 * from first launch to last completion, measured with CUDA events. */
int clx_ctx_run_steps(clx_ctx* ctx, clx_batch** batches, size_t n_batches, uint32_t steps, uint32_t n_streams,
                      float* total_ms);
/* Raw device pointers of a batch (for zero-copy consumers, e.g. torch / NCCL). */
void* clx_batch_device_out(clx_batch* b);
void* clx_batch_device_bytes(clx_batch* b);
/* Events-based timing of the kernels of the last `clx_batch_decode` on this batch (ms). */
int clx_batch_last_kernel_ms(clx_ctx* ctx, clx_batch* b, float* ms);
/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
uint64_t clx_ctx_launch_count(const clx_ctx* ctx);
void* clx_ctx_stream(clx_ctx* ctx, uint32_t stream_index); /* cudaStream_t */

/* Pinned (page-locked) host memory so that the copies inside clx_decode_frames are truly
 * asynchronous; optional — any host pointer works. */
void* clx_host_alloc(size_t bytes);
void clx_host_free(void* p);

/* ------------------------------------------------------------------------- */
/* claxon-shaped reader facade (FlacReader / FrameReader over the calls above) */
/* ------------------------------------------------------------------------- */
typedef struct clx_reader clx_reader;

/* FrameReader::new over an in-memory span positioned at a frame header
 * (src/frame.rs:652); `clx_reader_open_flac` is FlacReader::new + blocks(). */
int clx_reader_open_frames(clx_ctx* ctx, const uint8_t* bytes, size_t n, clx_reader** out);
int clx_reader_open_flac(clx_ctx* ctx, const uint8_t* bytes, size_t n, clx_reader** out);
int clx_reader_streaminfo(const clx_reader* r, clx_streaminfo* si);
/* read_next_or_eof: decodes ONE frame through the device path.  `buffer`/`capacity`
 * is the recycled Vec<i32>; on CLX_OK the block_size / channels / time outputs describe
 * the Block and `buffer[0 .. channels*block_size)` holds it.  If capacity is too small the
 * call returns CLX_ERR_INVALID_ARGUMENT with block_size and channels set, consuming nothing.
 * CLX_EOF == Ok(None). */
int clx_reader_next(clx_reader* r, int32_t* buffer, size_t capacity, uint32_t* block_size,
                    uint32_t* channels, uint64_t* time);
/* Batched extension, step 1 (optional): demuxes up to max_frames frames ahead of the reader's
 * position WITHOUT decoding and reports how many were found and how many output elements
 * clx_reader_next_batch(max_frames) needs for all of them (each frame aligned to 4 elements).
 * The demux is cached: the following next_batch with the same max_frames does not repeat it.
 * CLX_EOF at a clean end of stream; a header-level error if the very first frame has one. */
int clx_reader_plan_batch(clx_reader* r, size_t max_frames, size_t* n_frames, uint64_t* out_elems);
/* Batched extension: demuxes up to max_frames frames ahead and decodes them in one
 * device launch.  Frames that do not fit `capacity` are left for the next call; if not even the
 * first one fits the call returns CLX_ERR_INVALID_ARGUMENT (size the buffer with plan_batch).  Stops before the first frame that fails (that frame's status is
 * returned by the next call).  Returns the number of frames decoded in *n_decoded. */
int clx_reader_next_batch(clx_reader* r, size_t max_frames, int32_t* buffer, size_t capacity,
                          clx_frame_desc* descs, size_t* n_decoded);
uint64_t clx_reader_position(const clx_reader* r); /* byte offset of the next frame */
void clx_reader_close(clx_reader* r);

#ifdef __cplusplus
}
#endif
#endif /* CLAXON_B200_H */
