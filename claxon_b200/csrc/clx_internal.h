// clx_internal.h — shared between the kernels (clx_decode.cu) and the C ABI (clx_api.cu).
#ifndef CLX_INTERNAL_H
#define CLX_INTERNAL_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "claxon_b200.h"

// Device-only marker: the frame has an LPC order above what the first kernel instance keeps in
// registers and must be decoded by the 32-tap instance.  Never returned through the C ABI.
#define CLX_INTERNAL_NEED_HIGH_ORDER (-1)

namespace clx {
// Decodes `n_frames` frames described by d_descs from d_bytes (256-byte aligned; buf_bytes = allocated
// size, a multiple of 64 with at least 128 bytes of slack after the last frame) into
// d_out / d_results on `stream`.  d_need_hi is a 4-byte device scratch word.
cudaError_t launch_decode(const uint8_t* d_bytes, uint64_t buf_bytes, const clx_frame_desc* d_descs,
                          uint32_t n_frames, int32_t* d_out, clx_frame_result* d_results, int* d_need_hi,
                          cudaStream_t stream, uint64_t* launches);
}  // namespace clx
#endif
