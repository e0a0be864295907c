// clx_output.cu — the output stage after the decode path: planar i32 (claxon's Block layout,
// reference src/frame.rs:477-481) -> interleaved little-endian samples of 2, 3 or 4 bytes, the form
// FlacSamples yields them in (src/lib.rs:473-519: for each inter-channel sample, every channel in turn),
// WAV writers store them in (examples/decode.rs:48-62) and the STREAMINFO MD5 is defined over
// (src/metadata.rs:52-53).  Runs on the device so that 16-bit audio crosses PCIe as 2 bytes per sample.
#include <cuda_runtime.h>
#include <stdint.h>

#include "claxon_b200.h"
#include "clx_internal.h"

namespace clx {

constexpr uint32_t IL_TILE = 4096;  // interleaved elements per CTA

// One CTA per (tile, frame).  Element i of a frame's interleaved block is sample t = i / n_channels of
// channel c = i % n_channels, i.e. planar element c * block_size + t.  Writes are coalesced; the reads
// are n_channels coalesced streams.
template <int ESIZE>
__global__ void __launch_bounds__(256)
interleave_kernel(const clx_frame_desc* __restrict__ descs, uint32_t n_frames, const int32_t* __restrict__ planar,
                  uint8_t* __restrict__ dst) {
    const uint32_t f = blockIdx.y;
    if (f >= n_frames) return;
    const clx_frame_desc d = descs[f];
    const uint32_t nch = d.n_channels, bs = d.block_size, total = nch * bs;
    const uint32_t base = blockIdx.x * IL_TILE;
    if (base >= total) return;
    const int32_t* src = planar + d.out_offset;
    uint8_t* out = dst + d.out_offset * (uint64_t)ESIZE;
    if (ESIZE == 2 && nch == 2 && (d.out_offset & 1) == 0) {  // stereo 16-bit: one 4-byte store per sample pair
        for (uint32_t i = base / 2 + threadIdx.x; i < min(base + IL_TILE, total) / 2; i += 256) {
            const uint32_t l = (uint32_t)src[i] & 0xffffu, r = (uint32_t)src[bs + i];
            reinterpret_cast<uint32_t*>(out)[i] = l | (r << 16);
        }
        return;
    }
    for (uint32_t i = base + threadIdx.x; i < min(base + IL_TILE, total); i += 256) {
        const uint32_t t = i / nch, c = i - t * nch;
        const int32_t v = src[c * bs + t];
        if (ESIZE == 4) reinterpret_cast<int32_t*>(out)[i] = v;
        else if (ESIZE == 2) reinterpret_cast<int16_t*>(out)[i] = (int16_t)v;
        else {
            out[3 * (uint64_t)i] = (uint8_t)v;
            out[3 * (uint64_t)i + 1] = (uint8_t)(v >> 8);
            out[3 * (uint64_t)i + 2] = (uint8_t)(v >> 16);
        }
    }
}

uint32_t output_elem_size(uint32_t mode) {
    return mode == CLX_OUT_INTERLEAVED_I16 ? 2u : mode == CLX_OUT_INTERLEAVED_I24 ? 3u : 4u;
}

cudaError_t launch_interleave(const clx_frame_desc* d_descs, uint32_t n_frames, uint32_t max_frame_elems, const int32_t* d_planar,
                              void* d_dst, uint32_t mode, cudaStream_t stream) {
    if (n_frames == 0 || mode == CLX_OUT_PLANAR_I32) return cudaSuccess;
    const uint32_t tiles = (max_frame_elems + IL_TILE - 1) / IL_TILE;
    for (uint32_t f0 = 0; f0 < n_frames; f0 += 65535) {  // gridDim.y limit
        const uint32_t nf = min(65535u, n_frames - f0);
        dim3 grid(tiles, nf);
        if (mode == CLX_OUT_INTERLEAVED_I16)
            interleave_kernel<2><<<grid, 256, 0, stream>>>(d_descs + f0, nf, d_planar, (uint8_t*)d_dst);
        else if (mode == CLX_OUT_INTERLEAVED_I24)
            interleave_kernel<3><<<grid, 256, 0, stream>>>(d_descs + f0, nf, d_planar, (uint8_t*)d_dst);
        else
            interleave_kernel<4><<<grid, 256, 0, stream>>>(d_descs + f0, nf, d_planar, (uint8_t*)d_dst);
    }
    return cudaGetLastError();
}

}  // namespace clx
