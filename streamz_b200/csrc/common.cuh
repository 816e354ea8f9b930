// Shared internals of libstreamz_b200: context, error plumbing, growable device scratch.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/streamz_b200.h"

namespace szb {

void set_error(const char* fmt, ...);

#define SZB_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::szb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return SZB_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define SZB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            ::szb::set_error(__VA_ARGS__); \
            return SZB_ERR_INVALID;       \
        }                                 \
    } while (0)

#define SZB_TRY(expr)                     \
    do {                                  \
        szb_status _s = (expr);           \
        if (_s != SZB_OK) return _s;      \
    } while (0)

// A device buffer that only grows; reused across calls so steady-state calls do no cudaMalloc.
struct DevBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    szb_status reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

struct PinnedBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    szb_status reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

// Pinned staging for small host -> device uploads (segment tables, clip offsets) issued with cudaMemcpyAsync from *_dev entry
// points that return without synchronising: a slot is only rewritten after the copy that last read it has completed (event),
// so a second call cannot overwrite the table of a launch that is still queued.  kSlots calls may be in flight before a call
// has to wait for the oldest upload.
struct StagingRing {
    static constexpr int kSlots = 8;
    PinnedBuf buf[kSlots];
    cudaEvent_t ev[kSlots] = {};
    bool busy[kSlots] = {};
    int next = 0;
    int cur = -1;
    szb_status acquire(size_t bytes, void** out);      // waits for the slot's previous upload, grows it, returns its pointer
    szb_status uploaded(cudaStream_t stream);          // call right after enqueueing the copy that reads the acquired slot
    void release();
};

}  // namespace szb

struct szb_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;  // chunk pipeline of the host entry points
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;   // szb_timer_*
    uint64_t launches = 0;
    uint64_t graph_launches = 0;   // cudaGraphLaunch calls (captured two-step training graphs)
    bool small_steps = true;   // epochs whose batches have <= 32 rows run as one persistent cooperative kernel (SZB_NO_SMALL_KERNEL=1: off)
    int graph_max_rows = 1 << 30;   // largest batch whose epoch replays the captured two-step graph (SZB_GRAPH_MAX_ROWS); round 2 stopped at
                                    // 256 rows -- large batches gain little on an idle host (77.4 vs 79.0 us per batch-4096 step) but no longer
                                    // depend on the host keeping up with nine launches per 80 us (bench.py runs next to its clock sampler)
    bool graph_peers = false;   // ... also in multi-GPU steps with the peer-memory exchange (SZB_GRAPH_PEERS=1; the exchange step number then
                                // lives in device memory).  Tested at world 2 (identical results, same 102 us per step): no gain, so off
    bool graphs = true;   // small-batch training epochs replay a captured two-step CUDA graph (SZB_NO_GRAPHS=1 turns it off)
    bool pdl = true;   // programmatic dependent launch between the kernels of a training step (SZB_NO_PDL=1 turns it off)
    // Launch structure of a tensor-core training step (SZB_STEP_FUSE=<bits>, default 4; 0 = eleven launches per step: batch
    // kernel, 3 forward GEMMs, softmax, dW3, dX2, dW2, dX1, dW1, update).  Measured on B200, batch 4096, 60-512-256-100, TMA GEMMs:
    //   4: input gradients first, then the three weight-gradient GEMMs as ONE grouped launch            87.9 -> 78.4 us  (default)
    //   1: softmax / cross-entropy in the epilogue of the layer-3 GEMM (n_out <= 128, 32 CTAs of 128 rows instead of 64 + a
    //      separate 128-CTA kernel): 84.7 us with four threads per row (108 us with one) -- SLOWER, off
    //   2: the batch kernel (gather by permutation, dropout, transposed copy) runs for 32 steps at a time (batches > 256 rows):
    //      78.6 us -- no gain (with programmatic dependent launch the per-step batch kernel was already hidden), off
    int step_fuse = 4;
    bool gemm_tma = true;   // dense-layer GEMMs fetch their operands by TMA and split the producer work over two warp groups
                            // (gemm_tma.cuh); SZB_GEMM_TMA=0 selects gemm_tc_ta_kernel (cp.async)
    bool gemm_ta = true;    // dense-layer GEMMs take the A operand from tensor memory (gemm_tc_ta_kernel); SZB_GEMM_TA=0 selects the
                            // shared-memory-operand kernel (round 2: full GPU suite green with it, 97.1 vs 101.0 us per batch-4096 step)
    // per-launch timing of the extraction kernel (roofline figure)
    bool ktime_on = false;
    double ktime_ms = 0.0;
    uint64_t ktime_launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ktime_pending;
    std::vector<cudaEvent_t> pipe_events;                // chunk pipeline of szb_extract_batch
    // device-resident resample -> extract, OPTIONAL: the 44.1 kHz intermediate in a two-slot ring sized to stay in L2
    // (capi.cu).  Off by default -- measured on B200 (profiles/r02_l2_ring_sweep.txt): every setting is slower than the
    // single-chunk pipeline (21.7 ms per configs[1] step): 29.6 ms at 96 MB chunks, 37.8 ms at 32 MB; relaunching the
    // persistent extraction kernel per chunk costs more than the HBM round trip it saves (neither kernel is HBM-bound).
    // szb_ctx_set_fused_resample / SZB_FUSED_RESAMPLE=1: FIR inside the extraction kernel's staging (no 44.1 kHz intermediate
    // in HBM).  Bit-identical, but measured SLOWER on B200 (31.5 ms against 21.7 ms per configs[1] step: the FIR's 51 tap
    // registers do not fit beside the extraction kernel's state at 96 registers per thread, and its phases serialise behind
    // block barriers instead of overlapping across four independent CTAs), so it is off by default.
    bool fuse_resample = false;
    int l2_chunk_mb = 0;                                 // SZB_L2_CHUNK_MB (0 = one chunk, intermediate through HBM)
    int l2_streams = 2;                                  // SZB_L2_STREAMS: resampler of chunk k + 1 under the tail of extract k
    cudaEvent_t ring_ev[5] = {};                         // [0,1] resample done, [2,3] extract done (per slot), [4] tables uploaded
    // scratch
    szb::DevBuf segs, counter, pcm, feats, taps, labels, misc, probs, x;
    szb::StagingRing h_stage;
    // raw-audio training loops (loops.cu): the files' PCM + one augmented clip, and a constant label column
    szb::DevBuf loop_pcm, loop_labels;
    uint64_t loop_labels_n = 0;
    uint32_t loop_labels_class = 0;
    uint32_t taps_rate = 0;  // rate the taps buffer currently holds
    // NCCL (loaded lazily with dlopen; see comm.cu)
    void* nccl_comm = nullptr;
    cudaStream_t comm_stream = nullptr;   // gradient all-reduces overlapped with the rest of the backward pass
    cudaEvent_t ev_comm = nullptr;
    int rank = 0, world = 1;
    // Gradient exchange over NVLink peer memory (comm.cu): every rank's exchange region [flags | inbox | red] is mapped
    // into every other rank with CUDA IPC, and the update kernels (mlp.cu: p2p_exchange) run a two-shot all-reduce made of
    // peer stores in the same launch as the SGD update -- no NCCL call inside a training step.  Off (NCCL all-reduce) when
    // IPC / peer access is missing.
    static constexpr int kMaxPeers = 16;
    bool p2p_on = false;
    size_t p2p_cap = 0;                       // floats per buffer (inbox, red)
    uint32_t p2p_step = 0;                    // steps exchanged so far (flag value of the next step = p2p_step + 1)
    float* p2p_inbox[kMaxPeers] = {};         // [world][slice] partial slices received from every rank ([rank] is local memory)
    float* p2p_red[kMaxPeers] = {};           // the reduced gradient vector as every rank receives it
    uint32_t* p2p_flags[kMaxPeers] = {};      // flag block of every rank: [s] / [16 + s] = last step rank s finished phase 0 / 1 for
    szb::DevBuf p2p_counters;                 // private last-CTA tickets of the two phases
    int p2p_max_blocks = 0;                   // co-resident CTA capacity of the update kernels (0 = not queried yet)
    int p2p_mode = 0;                         // 0 = by size (one-shot below 6 MB of outgoing copies per step), 1 = one-shot, 2 = two-shot,
                                              // 3 = one-shot with the flag inside every 8-byte packet (no flag round; tensor-core path)
    bool p2p_ll_auto = false;                 // mode 0 prefers the packet protocol when the vector fits (SZB_P2P_LL=1).  Measured at N = 2: 101.9 us
                                              // per step against 102.6 us for the one-shot exchange with its flag round -- what an exchange costs
                                              // is the time until the peer's data is visible, with or without flags; off (twice the bytes at N = 8)
    bool p2p_early_push = false;              // one-shot exchange: finished gradient tiles are stored into the peers by the grouped weight-gradient
                                              // launch instead of the update kernel (SZB_P2P_EARLY_PUSH=1).  Measured at N = 2: 102.5 us per step
                                              // against 100.5 us without -- the flag round, not the data, is what the exchange costs; off
    bool p2p_trace_on = false;                // szb_comm_peer_trace: CTA 0 of the update kernel records its phase times
    void* p2p_region = nullptr;               // local allocation backing p2p_flags / p2p_inbox / p2p_red of this rank
};

namespace szb {
// Launch on the context's stream with the programmatic-stream-serialization attribute: the kernel may be scheduled while
// its predecessor in the stream is still draining.  ONLY for kernels that execute griddepcontrol.wait before their first
// global-memory access (and never write anything earlier); everything else keeps the <<<>>> launch.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(szb_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // Single-GPU steps only: with the overlapped NCCL all-reduces of a multi-GPU step, the early-launched CTAs of the next
    // GEMM take the SMs the collective's kernel would have started on, and the two ranks stall each other (measured at
    // N = 2: 132 ms per epoch with PDL against 48 ms without).
    // (The peer-memory exchange launches no collective kernel, so it keeps PDL.)
    cfg.numAttrs = (ctx->pdl && (ctx->world == 1 || ctx->p2p_on)) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}  // namespace szb
