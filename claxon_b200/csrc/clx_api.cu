// clx_api.cu — the device half of the C ABI (include/claxon_b200.h): context, device-resident
// batches, the end-to-end host-buffer decode call and the claxon-shaped reader facade.
//
// There is deliberately no CPU decode path in this library: without a usable CUDA device
// every entry point below fails with CLX_ERR_NO_DEVICE / CLX_ERR_CUDA.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "claxon_b200.h"
#include "clx_internal.h"

// A few long-lived host threads for the per-call CRC-16 pass: spawning std::threads per call costs more
// than the checksums themselves (6 MB per C2 batch).
class HostPool {
public:
    explicit HostPool(unsigned n) {
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    unsigned size() const { return (unsigned)workers_.size(); }
    // Runs fn(part, parts) for part = 0 .. parts-1: part 0 on the caller, the rest on the workers.
    void run(unsigned parts, const std::function<void(unsigned, unsigned)>& fn) {
        parts = std::max(1u, std::min(parts, size() + 1));
        if (parts > 1) {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            parts_ = parts;
            pending_ = parts - 1;
            gen_++;
        }
        if (parts > 1) cv_.notify_all();
        fn(0, parts);
        if (parts > 1) {
            std::unique_lock<std::mutex> l(m_);
            done_.wait(l, [this] { return pending_ == 0; });
            fn_ = nullptr;
        }
    }

private:
    void loop(unsigned idx) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(unsigned, unsigned)>* fn;
            unsigned parts;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
                parts = parts_;
            }
            if (fn && idx + 1 < parts) {
                (*fn)(idx + 1, parts);
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned, unsigned)>* fn_ = nullptr;
    unsigned parts_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct clx_ctx {
    int device = 0;
    uint32_t flags = 0;
    std::vector<cudaStream_t> streams;
    std::string last_error;
    uint64_t launches = 0;
    int sm_count = 148;
    size_t smem_budget = 227 * 1024;
    bool use_coop = true;
    bool warp_per_frame = false;  // CLX_OPT_WARP_PER_FRAME: the warp-per-frame fast path (clx_coop.cu) everywhere
    bool lane_per_frame_always = false;  // CLX_OPT_LANE_PER_FRAME: clx_fused.cu even for small synchronous calls
    // grow-only device scratch for clx_decode_frames, one set per stream (chunk pipelining)
    struct Scratch {
        uint8_t* d_bytes = nullptr; size_t bytes_cap = 0;
        clx_frame_desc* d_descs = nullptr; size_t descs_cap = 0;
        int32_t* d_out = nullptr; size_t out_cap = 0;
        clx_frame_result* d_results = nullptr; size_t results_cap = 0;
        int* d_need_hi = nullptr;
        uint8_t* d_params = nullptr; size_t params_cap = 0;  // fast path: per-subframe predictor parameters
        uint8_t* d_conv = nullptr; size_t conv_cap = 0;      // interleaved output modes: converted samples
    };
    std::vector<Scratch> scratch;
    // pinned staging for the small per-frame tables: pageable memory would make the "async" copies
    // synchronous and serialise the chunk pipeline
    clx_frame_desc* h_descs = nullptr; size_t h_descs_cap = 0;   // rebased descriptors (H2D)
    clx_frame_result* h_results = nullptr; size_t h_results_cap = 0;  // results (D2H)
    std::vector<uint8_t> crc_verdict;     // per frame: CRC-16 of the claimed span matched
    unsigned host_threads = 1;
    HostPool* pool = nullptr;   // created on first use
};

struct clx_batch {
    uint8_t* d_bytes = nullptr; size_t nbytes = 0, buf_bytes = 0;
    clx_frame_desc* d_descs = nullptr;
    int32_t* d_out = nullptr; size_t out_elems = 0;
    clx_frame_result* d_results = nullptr;
    int* d_need_hi = nullptr;
    void* d_params = nullptr;
    uint32_t n_frames = 0;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaStream_t last_stream = nullptr;
    clx::CoopPlan plan;
    // The batch's launch sequence (flag reset + kernels) captured once as a CUDA graph: a decode is then
    // one graph launch instead of five stream operations, which matters when a step is ~25 us.
    cudaGraphExec_t graph = nullptr;
    uint64_t graph_launches = 0;   // kernel launches inside the graph
    cudaEvent_t ev_idle = nullptr; // without a graph: the previous decode of this batch has finished
    // Frame CRC-16 (src/frame.rs:752-763): the batch's bytes never change, so the checksum of every claimed
    // span is taken once, on the host, when the batch is created; clx_batch_read applies it.
    std::vector<uint8_t> crc_ok;
    std::vector<uint64_t> h_offset;
    std::vector<uint32_t> h_len;
    std::vector<uint32_t> order;   // device position -> caller's frame index (empty: identity), see shape_order()
    bool device_crc = false;       // bytes came from device memory: the CRC-16 check runs on the device, inside the graph
};

namespace {

int cuda_fail(clx_ctx* ctx, cudaError_t e, const char* what) {
    if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return CLX_ERR_CUDA;
}
#define CU(ctx, call)                                              \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
    } while (0)

template <typename T>
int grow(clx_ctx* ctx, T*& ptr, size_t& cap, size_t need, size_t slack) {
    if (need <= cap && ptr) return CLX_OK;
    if (ptr) CU(ctx, cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    size_t want = need + need / 4 + slack;
    CU(ctx, cudaMalloc((void**)&ptr, want * sizeof(T)));
    cap = want;
    return CLX_OK;
}

// Host-side frame CRC-16 (src/frame.rs:752-763) for device-resident batches: their bytes never change, so the
// checksum of every claimed span is taken once, when the batch is created (the host-buffer call checks on the
// device instead, clx_crc.cu).  It takes effect only when the subframes decoded, so subframe errors keep their
// precedence over "frame CRC mismatch"; a frame that ended somewhere else than claimed gets a second look.
void precompute_crc_range(const uint8_t* bytes, const clx_frame_desc* descs, uint8_t* verdict, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
        const clx_frame_desc& d = descs[i];
        if (d.flags & CLX_FRAME_CRC16_VERIFIED) { verdict[i] = 1; continue; }  // demuxer already matched it
        if (d.byte_len < 2) { verdict[i] = 0; continue; }
        const uint8_t* f = bytes + d.byte_offset;
        const uint16_t stored = (uint16_t)(((uint32_t)f[d.byte_len - 2] << 8) | f[d.byte_len - 1]);
        verdict[i] = clx_crc16(f, d.byte_len - 2) == stored ? 1 : 0;
    }
}

void precompute_crc(clx_ctx* ctx, const uint8_t* bytes, const clx_frame_desc* descs, size_t n) {
    ctx->crc_verdict.assign(n, 0);
    if (ctx->flags & CLX_OPT_NO_VERIFY_CRC) return;
    uint8_t* verdict = ctx->crc_verdict.data();
    unsigned nt = std::min<unsigned>(ctx->host_threads, (unsigned)std::max<size_t>(1, n / 32));
    if (nt <= 1) return precompute_crc_range(bytes, descs, verdict, 0, n);
    if (!ctx->pool) ctx->pool = new HostPool(ctx->host_threads - 1);
    ctx->pool->run(nt, [&](unsigned part, unsigned parts) {
        precompute_crc_range(bytes, descs, verdict, n * part / parts, n * (part + 1) / parts);
    });
}

void build_graph(clx_ctx* ctx, clx_batch* b);

// Everything the kernels assume about a caller-supplied descriptor (the C ABI does not trust it).
bool valid_desc(const clx_frame_desc& d, size_t nbytes, size_t out_elems) {
    const uint64_t elems = (uint64_t)d.n_channels * d.block_size;
    const uint32_t bps = d.bits_per_sample;
    return d.byte_offset <= nbytes && d.byte_len <= nbytes - d.byte_offset && d.header_len <= d.byte_len &&
           d.n_channels >= 1 && d.n_channels <= 8 && d.block_size != 0 && d.byte_len <= (1u << 28) &&
           !(d.channel_assignment >= 8 && d.n_channels != 2) && d.channel_assignment <= 10 &&
           (bps == 0 || (bps >= 4 && bps <= 32)) &&  // 0: "not in the header" -> Unsupported, as the reference
           d.out_offset <= out_elems && elems <= out_elems - d.out_offset;
}

// The kernels map consecutive descriptors onto the lanes of a warp, and a warp advances at the pace of its longest
// block: frames of one shape belong next to each other.  Fills `order` (position -> frame index) with the frames
// [lo, hi) grouped by (channels, block size), stream order kept inside a group; returns false (order untouched) if
// they are all of one shape already — the usual case: a file's frames differ only in its last block.
bool shape_order(const clx_frame_desc* descs, size_t lo, size_t hi, std::vector<uint32_t>& order) {
    bool mixed = false;
    for (size_t i = lo + 1; i < hi && !mixed; i++)
        mixed = descs[i].block_size != descs[lo].block_size || descs[i].n_channels != descs[lo].n_channels;
    if (!mixed) return false;
    order.resize(hi - lo);
    for (size_t i = lo; i < hi; i++) order[i - lo] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const uint32_t ka = ((uint32_t)descs[a].n_channels << 16) | descs[a].block_size;
        const uint32_t kb = ((uint32_t)descs[b].n_channels << 16) | descs[b].block_size;
        return ka > kb;
    });
    return true;
}

// Chooses how a set of frames maps onto the cooperative kernel (frames per CTA, shared memory).
// Two fast paths, two regimes.  The lane-per-frame path (clx_fused.cu) has the fewest instructions per
// sample and is what a stream of batches should use; but a lane walks its whole frame alone, so one call
// takes ~0.4 ms of device time however few frames it holds.  A synchronous host-buffer call with a few
// thousand frames and nothing else in flight is latency-bound: there the warp-per-frame path
// (clx_coop.cu: 32 lanes share a frame, ~0.25 ms per 1024 frames) finishes sooner and lets the PCM
// copy-out start earlier.  `latency_call` = the plan is for such a call.
constexpr size_t kLatencyRegimeFrames = 4096;
clx::CoopPlan make_plan(const clx_ctx* ctx, const clx_frame_desc* descs, size_t n, bool latency_call = false) {
    clx::CoopPlan plan;
    if (!ctx->use_coop) return plan;
    uint32_t max_elems = 0, max_ch = 0, max_bs = 0, max_bps = 0;
    for (size_t i = 0; i < n; i++) {
        max_elems = std::max<uint32_t>(max_elems, (uint32_t)descs[i].n_channels * descs[i].block_size);
        max_ch = std::max<uint32_t>(max_ch, descs[i].n_channels);
        max_bs = std::max<uint32_t>(max_bs, descs[i].block_size);
        max_bps = std::max<uint32_t>(max_bps, descs[i].bits_per_sample);
    }
    // (Frames with many channels and long blocks — BASELINE.json's stress shape, 8 x 16384 — are latency-bound on
    // either fast path: the index lane walks (channels - 1) * block_size Rice codes alone.  Measured on 512 such
    // frames: 13.0 ms through the lane-per-frame path, 16.8 ms through the warp-per-frame path; no special case.)
    if (clx::coop_plan(max_elems, max_ch, (uint32_t)n, ctx->sm_count, ctx->smem_budget, &plan) && !ctx->warp_per_frame &&
        !(latency_call && !ctx->lane_per_frame_always)) {
        plan.G = 2;
        plan.max_bs = max_bs;
    }
    return plan;
}

}  // namespace

extern "C" {

int clx_ctx_create(const clx_options* opts, clx_ctx** out) {
    if (!out) return CLX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return CLX_ERR_NO_DEVICE;
    clx_ctx* ctx = new clx_ctx();
    ctx->device = opts ? opts->device : 0;
    ctx->flags = opts ? opts->flags : 0;
    if (ctx->device < 0 || ctx->device >= count) { delete ctx; return CLX_ERR_NO_DEVICE; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) { delete ctx; return CLX_ERR_NO_DEVICE; }
    if (clx::crc16_init() != cudaSuccess) { delete ctx; return CLX_ERR_CUDA; }
    uint32_t ns = opts && opts->n_streams ? opts->n_streams : 2;
    ns = std::min<uint32_t>(ns, 128);
    ctx->streams.resize(ns);
    for (uint32_t i = 0; i < ns; i++)
        if (cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return CLX_ERR_CUDA;
        }
    ctx->scratch.resize(ns);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess) {
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_budget = prop.sharedMemPerBlockOptin;
    }
    if (opts && (opts->flags & CLX_OPT_GENERIC_KERNEL_ONLY)) ctx->use_coop = false;
    if (opts && (opts->flags & CLX_OPT_WARP_PER_FRAME)) ctx->warp_per_frame = true;
    if (opts && (opts->flags & CLX_OPT_LANE_PER_FRAME)) ctx->lane_per_frame_always = true;
    ctx->host_threads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    if (opts && opts->host_threads) ctx->host_threads = std::max(1u, std::min(64u, opts->host_threads));
    *out = ctx;
    return CLX_OK;
}

void clx_ctx_destroy(clx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto& s : ctx->scratch) {
        cudaFree(s.d_bytes); cudaFree(s.d_descs); cudaFree(s.d_out); cudaFree(s.d_results); cudaFree(s.d_need_hi);
        cudaFree(s.d_params);
        cudaFree(s.d_conv);
    }
    for (auto s : ctx->streams) cudaStreamDestroy(s);
    if (ctx->h_descs) cudaFreeHost(ctx->h_descs);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    delete ctx->pool;
    delete ctx;
}

const char* clx_ctx_last_error(const clx_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t clx_ctx_launch_count(const clx_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* clx_ctx_stream(clx_ctx* ctx, uint32_t i) { return ctx ? (void*)ctx->streams[i % ctx->streams.size()] : nullptr; }

// ---------------------------------------------------------------------------------
// end-to-end decode with host buffers
// ---------------------------------------------------------------------------------
int clx_decode_frames(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs,
                      size_t n_frames, int32_t* out, size_t out_elems, clx_frame_result* results) {
    return clx_decode_frames_to(ctx, bytes, nbytes, descs, n_frames, out, out_elems, results, CLX_OUT_PLANAR_I32);
}

int clx_decode_frames_to(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs,
                         size_t n_frames, void* out_v, size_t out_elems, clx_frame_result* results, uint32_t mode) {
    if (!ctx || (!bytes && nbytes) || (!descs && n_frames) || (!results && n_frames) || mode > CLX_OUT_INTERLEAVED_I24)
        return CLX_ERR_INVALID_ARGUMENT;
    if (n_frames == 0) return CLX_OK;
    uint8_t* const out = static_cast<uint8_t*>(out_v);
    const size_t esize = clx::output_elem_size(mode);
    const uint32_t max_bps = mode == CLX_OUT_INTERLEAVED_I16 ? 16u : mode == CLX_OUT_INTERLEAVED_I24 ? 24u : 32u;
    for (size_t i = 0; i < n_frames; i++)
        if (descs[i].bits_per_sample > max_bps) return CLX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!out) return CLX_ERR_INVALID_ARGUMENT;
    for (size_t i = 0; i < n_frames; i++)
        if (!valid_desc(descs[i], nbytes, out_elems)) return CLX_ERR_INVALID_ARGUMENT;
    const double t0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    // Chunks of frames are pipelined over the context's streams: H2D of chunk i+1 and D2H of
    // chunk i-1 overlap the kernels of chunk i.  A chunk covers a contiguous byte range and a
    // contiguous output range (descriptors in stream order, as the demuxer emits them).
    const size_t ns = ctx->streams.size();
    size_t n_chunks = std::min<size_t>(ns, std::max<size_t>(1, n_frames / 128));
    for (size_t i = 1; i < n_frames && n_chunks > 1; i++)  // chunking needs stream order on both sides
        if (descs[i].byte_offset < descs[i - 1].byte_offset || descs[i].out_offset < descs[i - 1].out_offset)
            n_chunks = 1;
    if (ctx->h_descs_cap < n_frames) {
        if (ctx->h_descs) cudaFreeHost(ctx->h_descs);
        if (ctx->h_results) cudaFreeHost(ctx->h_results);
        ctx->h_descs = nullptr; ctx->h_results = nullptr; ctx->h_descs_cap = ctx->h_results_cap = 0;
        const size_t cap = n_frames + n_frames / 2 + 1024;
        CU(ctx, cudaHostAlloc((void**)&ctx->h_descs, cap * sizeof(clx_frame_desc), cudaHostAllocDefault));
        CU(ctx, cudaHostAlloc((void**)&ctx->h_results, cap * sizeof(clx_frame_result), cudaHostAllocDefault));
        ctx->h_descs_cap = ctx->h_results_cap = cap;
    }
    memcpy(ctx->h_descs, descs, n_frames * sizeof(clx_frame_desc));
    struct Span { size_t f0, f1; uint64_t b0, b1, o0, o1; };
    std::vector<Span> spans;
    std::vector<uint32_t> device_order;  // empty: the device sees the frames in the caller's order
    for (size_t c = 0; c < n_chunks; c++) {
        Span s{n_frames * c / n_chunks, n_frames * (c + 1) / n_chunks, ~0ull, 0, ~0ull, 0};
        for (size_t i = s.f0; i < s.f1; i++) {
            const clx_frame_desc& d = descs[i];
            s.b0 = std::min<uint64_t>(s.b0, d.byte_offset & ~15ull);
            s.b1 = std::max<uint64_t>(s.b1, d.byte_offset + d.byte_len);
            s.o0 = std::min<uint64_t>(s.o0, d.out_offset);
            s.o1 = std::max<uint64_t>(s.o1, d.out_offset + (uint64_t)d.n_channels * d.block_size);
        }
        // The device copy of the chunk's output keeps the host layout's alignment (offset mod 4 elements, so
        // 16-byte stores stay possible) but the copy back covers exactly [o0, o1): it never touches an element
        // before the chunk's first frame, so neighbouring chunks cannot overlap on the host side.
        // device order of the chunk's frames: grouped by shape (position p of the chunk holds frame order[p])
        std::vector<uint32_t> order;
        if (shape_order(descs, s.f0, s.f1, order)) {
            if (device_order.empty()) {
                device_order.resize(n_frames);
                for (size_t i = 0; i < n_frames; i++) device_order[i] = (uint32_t)i;
            }
            for (size_t p = 0; p < order.size(); p++) {
                device_order[s.f0 + p] = order[p];
                ctx->h_descs[s.f0 + p] = descs[order[p]];
            }
        }
        for (size_t i = s.f0; i < s.f1; i++) {
            ctx->h_descs[i].byte_offset -= s.b0;
            ctx->h_descs[i].out_offset -= s.o0 & ~3ull;
        }
        spans.push_back(s);
    }
    size_t enqueued = 0;  // chunks with work in flight
    // On any failure after the first enqueue: nothing may still be writing into the caller's buffers on return.
    auto drain = [&](int rc) {
        for (size_t c = 0; c < enqueued && c < n_chunks; c++) cudaStreamSynchronize(ctx->streams[c]);
        return rc;
    };
#define CUD(call)                                                         \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return drain(cuda_fail(ctx, e_, #call));   \
    } while (0)
    for (size_t c = 0; c < n_chunks; c++) {
        const Span& s = spans[c];
        clx_ctx::Scratch& sc = ctx->scratch[c];
        cudaStream_t st = ctx->streams[c];
        const size_t nb = (size_t)(s.b1 - s.b0), nf = s.f1 - s.f0, lead = (size_t)(s.o0 & 3), no = (size_t)(s.o1 - s.o0);
        int rc;
        const size_t nb_pad = ((nb + 63) & ~(size_t)63) + 128;  // whole 64-byte TMA chunks + look-ahead
        if ((rc = grow(ctx, sc.d_bytes, sc.bytes_cap, nb_pad, 4096))) return drain(rc);
        if ((rc = grow(ctx, sc.d_descs, sc.descs_cap, nf, 64))) return drain(rc);
        if ((rc = grow(ctx, sc.d_out, sc.out_cap, lead + no + 4, 4096))) return drain(rc);
        if ((rc = grow(ctx, sc.d_results, sc.results_cap, nf, 64))) return drain(rc);
        if (!sc.d_need_hi) CUD(cudaMalloc((void**)&sc.d_need_hi, 4 * sizeof(int)));
        const clx::CoopPlan plan = make_plan(ctx, descs + s.f0, nf, n_frames <= kLatencyRegimeFrames);
        if ((rc = grow(ctx, sc.d_params, sc.params_cap, clx::coop_params_bytes(plan, (uint32_t)nf) + 16, 4096))) return drain(rc);
        if (mode != CLX_OUT_PLANAR_I32 && (rc = grow(ctx, sc.d_conv, sc.conv_cap, (lead + no + 4) * esize, 4096))) return drain(rc);
        enqueued = c + 1;
        CUD(cudaMemcpyAsync(sc.d_bytes, bytes + s.b0, nb, cudaMemcpyHostToDevice, st));
        CUD(cudaMemcpyAsync(sc.d_descs, ctx->h_descs + s.f0, nf * sizeof(clx_frame_desc), cudaMemcpyHostToDevice, st));
        CUD(clx::launch_decode(sc.d_bytes, nb_pad, sc.d_descs, (uint32_t)nf, sc.d_out, sc.d_results, sc.d_need_hi,
                               sc.d_params, plan, st, &ctx->launches));
        if (!(ctx->flags & CLX_OPT_NO_VERIFY_CRC)) {  // src/frame.rs:752-763, after the subframes, on the device
            CUD(clx::launch_crc16(sc.d_bytes, sc.d_descs, (uint32_t)nf, sc.d_results, st));
            ctx->launches++;
        }
        if (mode == CLX_OUT_PLANAR_I32) {
            CUD(cudaMemcpyAsync(out + s.o0 * esize, sc.d_out + lead, no * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        } else {
            uint32_t max_elems = 0;
            for (size_t i = s.f0; i < s.f1; i++) max_elems = std::max<uint32_t>(max_elems, (uint32_t)descs[i].n_channels * descs[i].block_size);
            CUD(clx::launch_interleave(sc.d_descs, (uint32_t)nf, max_elems, sc.d_out, sc.d_conv, mode, st));
            ctx->launches++;
            CUD(cudaMemcpyAsync(out + s.o0 * esize, sc.d_conv + lead * esize, no * esize, cudaMemcpyDeviceToHost, st));
        }
        CUD(cudaMemcpyAsync(ctx->h_results + s.f0, sc.d_results, nf * sizeof(clx_frame_result), cudaMemcpyDeviceToHost, st));
    }
#ifdef CLX_EXPERIMENT
    static const bool trace = getenv("CLX_TRACE") != nullptr;
#else
    constexpr bool trace = false;
#endif
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t1 = trace ? now() : 0;
    for (size_t c = 0; c < n_chunks; c++) CUD(cudaStreamSynchronize(ctx->streams[c]));
#undef CUD
    const double t2 = trace ? now() : 0;
    if (device_order.empty()) memcpy(results, ctx->h_results, n_frames * sizeof(clx_frame_result));
    else
        for (size_t p = 0; p < n_frames; p++) results[device_order[p]] = ctx->h_results[p];
    if (trace)
        fprintf(stderr, "[clx] frames=%zu chunks=%zu submit=%.3f ms wait=%.3f ms results=%.3f ms\n", n_frames, n_chunks, t1 - t0,
                t2 - t1, now() - t2);
    return CLX_OK;
}

// ---------------------------------------------------------------------------------
// device-resident batches
// ---------------------------------------------------------------------------------
int clx_batch_create(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs, size_t n_frames,
                     size_t out_elems, clx_batch** out) {
    return clx_batch_create_ex(ctx, bytes, nbytes, descs, n_frames, out_elems, 0, out);
}

int clx_batch_create_ex(clx_ctx* ctx, const uint8_t* bytes, size_t nbytes, const clx_frame_desc* descs, size_t n_frames,
                        size_t out_elems, uint32_t batch_flags, clx_batch** out) {
    if (!ctx || !out || (!bytes && nbytes) || (!descs && n_frames)) return CLX_ERR_INVALID_ARGUMENT;
    const bool on_device = (batch_flags & CLX_BATCH_BYTES_ON_DEVICE) != 0;
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n_frames; i++)
        if (!valid_desc(descs[i], nbytes, out_elems)) return CLX_ERR_INVALID_ARGUMENT;
    clx_batch* b = new clx_batch();
    b->nbytes = nbytes;
    b->buf_bytes = ((nbytes + 63) & ~(size_t)63) + 128;  // whole 64-byte TMA chunks + look-ahead
    b->out_elems = out_elems;
    b->n_frames = (uint32_t)n_frames;
    b->plan = make_plan(ctx, descs, n_frames);
    cudaError_t e = cudaMalloc((void**)&b->d_bytes, b->buf_bytes);
    if (e == cudaSuccess) e = cudaMemset(b->d_bytes, 0, b->buf_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_descs, std::max<size_t>(1, n_frames) * sizeof(clx_frame_desc));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_out, (out_elems + 4) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_results, std::max<size_t>(1, n_frames) * sizeof(clx_frame_result));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->d_need_hi, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_params, clx::coop_params_bytes(b->plan, b->n_frames) + 16);
    if (e == cudaSuccess) e = cudaMemcpy(b->d_bytes, bytes, nbytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        if (shape_order(descs, 0, n_frames, b->order)) {
            std::vector<clx_frame_desc> sorted(n_frames);
            for (size_t p = 0; p < n_frames; p++) sorted[p] = descs[b->order[p]];
            e = cudaMemcpy(b->d_descs, sorted.data(), n_frames * sizeof(clx_frame_desc), cudaMemcpyHostToDevice);
        } else e = cudaMemcpy(b->d_descs, descs, n_frames * sizeof(clx_frame_desc), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_start);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev_stop);
    if (e != cudaSuccess) {
        clx_batch_destroy(ctx, b);
        return cuda_fail(ctx, e, "clx_batch_create");
    }
    if (!(ctx->flags & CLX_OPT_NO_VERIFY_CRC)) {
        if (on_device) b->device_crc = true;  // no host copy to checksum: clx_crc.cu, as part of every decode
        else {
            precompute_crc(ctx, bytes, descs, n_frames);
            b->crc_ok = ctx->crc_verdict;
        }
    }
    b->h_offset.resize(n_frames);
    b->h_len.resize(n_frames);
    for (size_t i = 0; i < n_frames; i++) { b->h_offset[i] = descs[i].byte_offset; b->h_len[i] = descs[i].byte_len; }
    build_graph(ctx, b);
    *out = b;
    return CLX_OK;
}

}  // extern "C"

namespace {
// Captures the batch's launch sequence once; called from clx_batch_create so that no decode ever pays for
// (or is timed with) a graph instantiation.
void build_graph(clx_ctx* ctx, clx_batch* b) {
#ifdef CLX_EXPERIMENT
    if (getenv("CLX_NO_GRAPH")) return;
#endif
    if (b->n_frames == 0) return;
    cudaStream_t st = ctx->streams[0];
    cudaGraph_t g = nullptr;
    uint64_t n = 0;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        cudaError_t e1 = clx::launch_decode(b->d_bytes, b->buf_bytes, b->d_descs, b->n_frames, b->d_out, b->d_results,
                                            b->d_need_hi, b->d_params, b->plan, st, &n);
        if (e1 == cudaSuccess && b->device_crc) { e1 = clx::launch_crc16(b->d_bytes, b->d_descs, b->n_frames, b->d_results, st); n++; }
        e = cudaStreamEndCapture(st, &g);
        if (e1 != cudaSuccess) e = e1;
    }
    if (e == cudaSuccess && g) e = cudaGraphInstantiate(&b->graph, g, 0);
    if (g) cudaGraphDestroy(g);
    if (e != cudaSuccess || !b->graph) {
        b->graph = nullptr;
        cudaGetLastError();
    } else b->graph_launches = n;
}

// Enqueues one decode of a device-resident batch on `st`, through the batch's graph when there is one.
int enqueue_batch(clx_ctx* ctx, clx_batch* b, cudaStream_t st) {
    if (b->graph) {
        CU(ctx, cudaGraphLaunch(b->graph, st));
        ctx->launches += b->graph_launches;
        return CLX_OK;
    }
    // No graph: two decodes of one batch share its flag words and parameter records, so they must not overlap.
    if (b->ev_idle) CU(ctx, cudaStreamWaitEvent(st, b->ev_idle, 0));
    CU(ctx, clx::launch_decode(b->d_bytes, b->buf_bytes, b->d_descs, b->n_frames, b->d_out, b->d_results, b->d_need_hi,
                               b->d_params, b->plan, st, &ctx->launches));
    if (b->device_crc) { CU(ctx, clx::launch_crc16(b->d_bytes, b->d_descs, b->n_frames, b->d_results, st)); ctx->launches++; }
    if (!b->ev_idle) CU(ctx, cudaEventCreateWithFlags(&b->ev_idle, cudaEventDisableTiming));
    CU(ctx, cudaEventRecord(b->ev_idle, st));
    return CLX_OK;
}
}  // namespace

extern "C" {

int clx_batch_decode(clx_ctx* ctx, clx_batch* b, uint32_t stream_index) {
    if (!ctx || !b) return CLX_ERR_INVALID_ARGUMENT;
    cudaStream_t st = ctx->streams[stream_index % ctx->streams.size()];
    b->last_stream = st;
    CU(ctx, cudaEventRecord(b->ev_start, st));
    int rc = enqueue_batch(ctx, b, st);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(b->ev_stop, st));
    return CLX_OK;
}

int clx_batch_sync(clx_ctx* ctx, clx_batch* b) {
    if (!ctx || !b) return CLX_ERR_INVALID_ARGUMENT;
    if (b->last_stream) CU(ctx, cudaStreamSynchronize(b->last_stream));
    return CLX_OK;
}

int clx_batch_last_kernel_ms(clx_ctx* ctx, clx_batch* b, float* ms) {
    if (!ctx || !b || !ms) return CLX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaEventSynchronize(b->ev_stop));
    CU(ctx, cudaEventElapsedTime(ms, b->ev_start, b->ev_stop));
    return CLX_OK;
}

int clx_batch_read(clx_ctx* ctx, clx_batch* b, int32_t* out, size_t out_elems, clx_frame_result* results) {
    if (!ctx || !b) return CLX_ERR_INVALID_ARGUMENT;
    int rc = clx_batch_sync(ctx, b);
    if (rc) return rc;
    if (out) CU(ctx, cudaMemcpy(out, b->d_out, std::min(out_elems, b->out_elems) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (results) {
        if (b->order.empty()) {
            CU(ctx, cudaMemcpy(results, b->d_results, b->n_frames * sizeof(clx_frame_result), cudaMemcpyDeviceToHost));
        } else {
            std::vector<clx_frame_result> dev(b->n_frames);
            CU(ctx, cudaMemcpy(dev.data(), b->d_results, b->n_frames * sizeof(clx_frame_result), cudaMemcpyDeviceToHost));
            for (size_t p = 0; p < b->n_frames; p++) results[b->order[p]] = dev[p];
        }
        if (!(ctx->flags & CLX_OPT_NO_VERIFY_CRC) && !b->device_crc) {
            std::vector<uint8_t> tmp;
            for (size_t i = 0; i < b->n_frames; i++) {
                if (results[i].status != CLX_OK) continue;
                bool ok;
                const uint32_t consumed = results[i].consumed;
                if (consumed == b->h_len[i]) ok = b->crc_ok[i] != 0;
                else {  // the frame ended before the end of the span it was given: checksum what it did consume
                    if (consumed < 2 || consumed > b->h_len[i]) ok = false;
                    else {
                        tmp.resize(consumed);
                        CU(ctx, cudaMemcpy(tmp.data(), b->d_bytes + b->h_offset[i], consumed, cudaMemcpyDeviceToHost));
                        const uint16_t stored = (uint16_t)(((uint32_t)tmp[consumed - 2] << 8) | tmp[consumed - 1]);
                        ok = clx_crc16(tmp.data(), consumed - 2) == stored;
                    }
                }
                if (!ok) results[i].status = CLX_ERR_FRAME_CRC_MISMATCH;
            }
        }
    }
    return CLX_OK;
}

void clx_batch_destroy(clx_ctx* ctx, clx_batch* b) {
    (void)ctx;
    if (!b) return;
    cudaFree(b->d_bytes); cudaFree(b->d_descs); cudaFree(b->d_out); cudaFree(b->d_results); cudaFree(b->d_need_hi);
    cudaFree(b->d_params);
    if (b->graph) cudaGraphExecDestroy(b->graph);
    if (b->ev_idle) cudaEventDestroy(b->ev_idle);
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    if (b->ev_stop) cudaEventDestroy(b->ev_stop);
    delete b;
}

// Decodes `steps` batches back to back, step i taking batches[i % n_batches] on internal stream
// i % n_streams, and returns the device time (CUDA events) from the first launch to the last
// completion.  This is the steady-state "many batches in flight" regime of a decode service.
int clx_ctx_run_steps(clx_ctx* ctx, clx_batch** batches, size_t n_batches, uint32_t steps, uint32_t n_streams,
                      float* total_ms) {
    if (!ctx || !batches || n_batches == 0 || !total_ms) return CLX_ERR_INVALID_ARGUMENT;
    CU(ctx, cudaSetDevice(ctx->device));
    n_streams = std::max<uint32_t>(1, std::min<uint32_t>(n_streams, (uint32_t)ctx->streams.size()));
    cudaEvent_t start = nullptr, stop = nullptr;
    std::vector<cudaEvent_t> done(n_streams, nullptr);
    auto body = [&]() -> int {
        CU(ctx, cudaEventCreate(&start));
        CU(ctx, cudaEventCreate(&stop));
        for (auto& e : done) CU(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaStream_t s0 = ctx->streams[0];
        CU(ctx, cudaEventRecord(start, s0));
        for (uint32_t s = 1; s < n_streams; s++) CU(ctx, cudaStreamWaitEvent(ctx->streams[s], start, 0));
        for (uint32_t i = 0; i < steps; i++) {
            clx_batch* b = batches[i % n_batches];
            cudaStream_t st = ctx->streams[i % n_streams];
            b->last_stream = st;
            int rc = enqueue_batch(ctx, b, st);
            if (rc) return rc;
        }
        for (uint32_t s = 1; s < n_streams; s++) {
            CU(ctx, cudaEventRecord(done[s], ctx->streams[s]));
            CU(ctx, cudaStreamWaitEvent(s0, done[s], 0));
        }
        CU(ctx, cudaEventRecord(stop, s0));
        CU(ctx, cudaEventSynchronize(stop));
        CU(ctx, cudaEventElapsedTime(total_ms, start, stop));
        return CLX_OK;
    };
    const int rc = body();
    if (rc != CLX_OK) cudaDeviceSynchronize();  // nothing of this call may still be running when it returns
    if (start) cudaEventDestroy(start);
    if (stop) cudaEventDestroy(stop);
    for (auto& e : done)
        if (e) cudaEventDestroy(e);
    return rc;
}

void* clx_batch_device_out(clx_batch* b) { return b ? b->d_out : nullptr; }
void* clx_batch_device_bytes(clx_batch* b) { return b ? b->d_bytes : nullptr; }

// Pinned host memory for callers that want true asynchronous copies.
void* clx_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void clx_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"

// ---------------------------------------------------------------------------------
// reader facade: claxon::FrameReader / FlacReader::blocks() over the batched device path
// ---------------------------------------------------------------------------------
struct clx_reader {
    clx_ctx* ctx = nullptr;
    const uint8_t* bytes = nullptr;
    size_t n = 0;
    uint64_t pos = 0;
    bool have_si = false;
    clx_streaminfo si{};
    std::vector<clx_frame_desc> descs;
    std::vector<clx_frame_result> results;
    // frames demuxed ahead by clx_reader_plan_batch / clx_reader_next_batch, valid while pos == plan_pos
    size_t plan_n = 0, plan_max = 0;
    uint64_t plan_pos = ~0ull, plan_elems = 0;
    int plan_stop = CLX_OK;
};

namespace {
// Upper bound on the bytes a sane encoder spends on a frame: verbatim coding plus slack.
size_t sane_frame_bound(const clx_frame_desc& d) {
    const size_t per_ch = ((size_t)d.block_size * (d.bits_per_sample + 2u)) / 8 + 256;
    return (size_t)d.header_len + (size_t)d.n_channels * per_ch + 2;
}
uint64_t block_time(const clx_frame_desc& d) {  // src/frame.rs:771-774
    return (d.flags & CLX_FRAME_VARIABLE_BLOCKING) ? d.number : (uint64_t)d.block_size * d.number;
}
}  // namespace

extern "C" {

int clx_reader_open_frames(clx_ctx* ctx, const uint8_t* bytes, size_t n, clx_reader** out) {
    if (!ctx || !out || (!bytes && n)) return CLX_ERR_INVALID_ARGUMENT;
    clx_reader* r = new clx_reader();
    r->ctx = ctx;
    r->bytes = bytes;
    r->n = n;
    *out = r;
    return CLX_OK;
}

int clx_reader_open_flac(clx_ctx* ctx, const uint8_t* bytes, size_t n, clx_reader** out) {
    if (!ctx || !out || (!bytes && n)) return CLX_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    clx_streaminfo si;
    uint64_t first = 0;
    int st = clx_open_stream(bytes, n, &si, &first);
    if (st) return st;
    clx_reader* r = new clx_reader();
    r->ctx = ctx;
    r->bytes = bytes;
    r->n = n;
    r->pos = first;
    r->have_si = true;
    r->si = si;
    *out = r;
    return CLX_OK;
}

int clx_reader_streaminfo(const clx_reader* r, clx_streaminfo* si) {
    if (!r || !si || !r->have_si) return CLX_ERR_INVALID_ARGUMENT;
    *si = r->si;
    return CLX_OK;
}

uint64_t clx_reader_position(const clx_reader* r) { return r ? r->pos : 0; }
void clx_reader_close(clx_reader* r) { delete r; }

int clx_reader_next(clx_reader* r, int32_t* buffer, size_t capacity, uint32_t* block_size, uint32_t* channels,
                    uint64_t* time) {
    if (!r) return CLX_ERR_INVALID_ARGUMENT;
    if (r->pos > r->n) return CLX_EOF;
    clx_frame_desc d;
    const size_t avail = r->n - r->pos;
    int st = clx_parse_frame_header(r->bytes + r->pos, avail, &d, r->ctx->flags);
    if (st) return st;  // CLX_EOF == Ok(None)
    if (block_size) *block_size = d.block_size;
    if (channels) *channels = d.n_channels;
    const size_t elems = (size_t)d.n_channels * d.block_size;
    if (capacity < elems || !buffer) return CLX_ERR_INVALID_ARGUMENT;  // ensure_buffer_len is the caller's job
    d.byte_offset = r->pos;
    d.out_offset = 0;
    clx_frame_result res{};
    // A frame does not carry its length; give the device a generous window and widen it to the
    // rest of the stream in the (pathological) case the frame turns out to be longer.
    size_t window = std::min(avail, sane_frame_bound(d));
    for (;;) {
        d.byte_len = (uint32_t)std::min<size_t>(window, (size_t)1 << 28);
        st = clx_decode_frames(r->ctx, r->bytes, r->n, &d, 1, buffer, capacity, &res);
        if (st) return st;
        if (res.status == CLX_ERR_IO_UNEXPECTED_EOF && window < avail) { window = avail; continue; }
        break;
    }
    if (res.status != CLX_OK) return res.status;
    r->pos += res.consumed;
    if (time) *time = block_time(d);
    return CLX_OK;
}

}  // extern "C"

namespace {
// Demuxes up to max_frames frames ahead of the reader's position (once per position).
void plan(clx_reader* r, size_t max_frames) {
    if (r->plan_pos == r->pos && r->plan_max == max_frames) return;
    r->descs.resize(max_frames);
    uint64_t next = r->pos, total = 0;
    int stop = CLX_OK;
    size_t n = 0;
    if (r->pos > r->n) stop = CLX_EOF;
    else n = clx_demux_frames(r->bytes, r->n, r->pos, r->descs.data(), max_frames, &next, &total, &stop, r->ctx->flags);
    r->plan_n = n; r->plan_max = max_frames; r->plan_pos = r->pos; r->plan_elems = total; r->plan_stop = stop;
}
}  // namespace

extern "C" {

int clx_reader_plan_batch(clx_reader* r, size_t max_frames, size_t* n_frames, uint64_t* out_elems) {
    if (!r || !n_frames || !out_elems) return CLX_ERR_INVALID_ARGUMENT;
    *n_frames = 0;
    *out_elems = 0;
    if (max_frames == 0) return CLX_OK;
    plan(r, max_frames);
    if (r->plan_n == 0) return r->plan_stop;  // header-level error or CLX_EOF
    *n_frames = r->plan_n;
    *out_elems = r->plan_elems;
    return CLX_OK;
}

int clx_reader_next_batch(clx_reader* r, size_t max_frames, int32_t* buffer, size_t capacity, clx_frame_desc* descs,
                          size_t* n_decoded) {
    if (!r || !n_decoded || !descs) return CLX_ERR_INVALID_ARGUMENT;
    *n_decoded = 0;
    if (max_frames == 0) return CLX_OK;
    plan(r, max_frames);
    size_t n = r->plan_n;
    if (n == 0) return r->plan_stop;  // header-level error or CLX_EOF
    // fit the caller's buffer
    while (n > 0) {
        const clx_frame_desc& last = r->descs[n - 1];
        if (last.out_offset + (uint64_t)last.n_channels * last.block_size <= capacity) break;
        n--;
    }
    if (n == 0 || !buffer) return CLX_ERR_INVALID_ARGUMENT;  // not even the first frame fits: see clx_reader_plan_batch
    r->results.resize(n);
    int st = clx_decode_frames(r->ctx, r->bytes, r->n, r->descs.data(), n, buffer, capacity, r->results.data());
    if (st) return st;
    size_t good = 0;
    while (good < n && r->results[good].status == CLX_OK) good++;
    if (good == 0) return r->results[0].status;  // the very first frame failed
    for (size_t i = 0; i < good; i++) {
        descs[i] = r->descs[i];
        descs[i].byte_len = r->results[i].consumed;
        descs[i].number = block_time(r->descs[i]);  // Block::time()
    }
    const clx_frame_desc& lastd = r->descs[good - 1];
    r->pos = lastd.byte_offset + r->results[good - 1].consumed;
    *n_decoded = good;
    return CLX_OK;
}

}  // extern "C"

#ifdef CLX_EXPERIMENT
// Measurement builds only (libclaxon_b200_exp.so, tools/exp_*.py): choose which passes of the throughput path a
// batch's graph contains, and rebuild a batch's graph after changing the choice.
extern "C" void clx_exp_set_which(int which) { clx::g_exp_which = which; }
extern "C" void clx_exp_set_dyn_smem(int bytes) { clx::g_exp_dyn_smem = bytes; }
extern "C" int clx_exp_rebuild_graph(clx_ctx* ctx, clx_batch* b) {
    if (!ctx || !b) return CLX_ERR_INVALID_ARGUMENT;
    cudaDeviceSynchronize();
    if (b->graph) { cudaGraphExecDestroy(b->graph); b->graph = nullptr; }
    build_graph(ctx, b);
    return b->graph ? CLX_OK : CLX_ERR_CUDA;
}
#endif
