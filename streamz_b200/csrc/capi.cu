// C-ABI layer: context management, tables, and the front-end entry points of include/streamz_b200.h.
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "common.cuh"
#include "frontend.cuh"
#include "tables.hpp"

namespace szb {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

szb_status DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return SZB_OK;
    release();
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cap = 0;
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return SZB_ERR_ALLOC;
    }
    cap = want;
    return SZB_OK;
}
void DevBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}
szb_status PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return SZB_OK;
    release();
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMallocHost(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cap = 0;
        set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
        return SZB_ERR_ALLOC;
    }
    cap = want;
    return SZB_OK;
}
void PinnedBuf::release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
}

szb_status StagingRing::acquire(size_t bytes, void** out) {
    const int s = next;
    next = (next + 1) % kSlots;
    if (busy[s]) {
        SZB_CUDA(cudaEventSynchronize(ev[s]));
        busy[s] = false;
    }
    SZB_TRY(buf[s].reserve(bytes ? bytes : 1));
    cur = s;
    *out = buf[s].ptr;
    return SZB_OK;
}
szb_status StagingRing::uploaded(cudaStream_t stream) {
    if (cur < 0) return SZB_OK;
    if (!ev[cur]) SZB_CUDA(cudaEventCreateWithFlags(&ev[cur], cudaEventDisableTiming));
    SZB_CUDA(cudaEventRecord(ev[cur], stream));
    busy[cur] = true;
    cur = -1;
    return SZB_OK;
}
void StagingRing::release() {
    for (int s = 0; s < kSlots; ++s) {
        if (busy[s]) cudaEventSynchronize(ev[s]);
        if (ev[s]) cudaEventDestroy(ev[s]);
        ev[s] = nullptr;
        busy[s] = false;
        buf[s].release();
    }
}

static szb_status drain_ktime(szb_ctx* ctx) {
    for (auto& pr : ctx->ktime_pending) {
        SZB_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        SZB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        ctx->ktime_ms += ms;
        ctx->ktime_launches += 1;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    ctx->ktime_pending.clear();
    return SZB_OK;
}

// Window offsets of a batch; clips are at `rate` Hz (resampled length floor(n * 44100 / rate), lib.rs:196).
static void batch_layout(const uint64_t* clip_off, uint32_t n_clips, uint32_t rate, std::vector<uint64_t>& off44,
                         std::vector<uint64_t>& win_off) {
    off44.resize(size_t(n_clips) + 1);
    win_off.resize(size_t(n_clips) + 1);
    off44[0] = 0;
    win_off[0] = 0;
    for (uint32_t c = 0; c < n_clips; ++c) {
        const uint64_t n_in = clip_off[c + 1] - clip_off[c];
        const uint64_t n44 = rate == SZB_SAMPLE_RATE ? n_in : szb_resample_out_len(n_in, rate);
        // resampled clips are laid out back to back, each start rounded up to 8 samples (16-byte aligned rows)
        off44[c + 1] = rate == SZB_SAMPLE_RATE ? clip_off[c + 1] - clip_off[0] : ((off44[c] + n44 + 7) & ~uint64_t(7));
        win_off[c + 1] = win_off[c] + szb_num_windows(n44);
    }
}

}  // namespace szb

using namespace szb;

extern "C" {

const char* szb_version(void) { return "streamz_b200 0.1.0 (sm_100a)"; }
const char* szb_last_error(void) { return g_last_error.c_str(); }

szb_status szb_ctx_create(int32_t device, void* stream, szb_ctx** out) {
    SZB_REQUIRE(out != nullptr, "szb_ctx_create: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        set_error("no CUDA device available (%s); streamz_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SZB_ERR_NO_DEVICE;
    }
    SZB_REQUIRE(device >= 0 && device < n_dev, "szb_ctx_create: device %d out of range (%d devices)", device, n_dev);
    SZB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SZB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only and has no fallback", device, prop.major,
                  prop.minor);
        return SZB_ERR_NO_DEVICE;
    }
    szb_ctx* ctx = new szb_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("SZB_NO_PDL")) ctx->pdl = !(e[0] == '1');
    if (const char* e = getenv("SZB_NO_GRAPHS")) ctx->graphs = !(e[0] == '1');
    if (const char* e = getenv("SZB_GRAPH_MAX_ROWS")) ctx->graph_max_rows = std::max(0, atoi(e));
    if (const char* e = getenv("SZB_GRAPH_PEERS")) ctx->graph_peers = (e[0] == '1');
    if (const char* e = getenv("SZB_NO_SMALL_KERNEL")) ctx->small_steps = !(e[0] == '1');
    if (const char* e = getenv("SZB_GEMM_TA")) ctx->gemm_ta = (e[0] == '1');
    if (const char* e = getenv("SZB_STEP_FUSE")) ctx->step_fuse = atoi(e) & 7;
    if (const char* e = getenv("SZB_GEMM_TMA")) ctx->gemm_tma = (e[0] == '1');
    if (const char* e = getenv("SZB_P2P_EARLY_PUSH")) ctx->p2p_early_push = (e[0] == '1');
    if (const char* e = getenv("SZB_P2P_LL")) ctx->p2p_ll_auto = (e[0] == '1');
    if (const char* e = getenv("SZB_L2_CHUNK_MB")) ctx->l2_chunk_mb = std::max(0, atoi(e));
    if (const char* e = getenv("SZB_FUSED_RESAMPLE")) ctx->fuse_resample = (e[0] == '1');
    if (const char* e = getenv("SZB_L2_STREAMS")) ctx->l2_streams = atoi(e) > 1 ? 2 : 1;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
        ctx->own_stream = false;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            set_error("cudaStreamCreate failed");
            return SZB_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking);
    cudaEventCreate(&ctx->ev_start);
    cudaEventCreate(&ctx->ev_stop);
    for (cudaEvent_t& e : ctx->ring_ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    szb_status s = upload_frontend_tables();
    if (s != SZB_OK) {
        szb_ctx_destroy(ctx);
        return s;
    }
    *out = ctx;
    return SZB_OK;
}

void szb_ctx_destroy(szb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    szb_comm_destroy(ctx);
    for (auto& pr : ctx->ktime_pending) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    for (cudaEvent_t e : ctx->pipe_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ring_ev) if (e) cudaEventDestroy(e);
    for (DevBuf* b : { &ctx->segs, &ctx->counter, &ctx->pcm, &ctx->feats, &ctx->taps, &ctx->labels, &ctx->misc, &ctx->probs,
                       &ctx->x, &ctx->loop_pcm, &ctx->loop_labels, &ctx->p2p_counters })
        b->release();
    ctx->h_stage.release();
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

szb_status szb_ctx_sync(szb_ctx* ctx) {
    SZB_REQUIRE(ctx, "szb_ctx_sync: ctx is NULL");
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}
int32_t szb_ctx_sm_count(const szb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
szb_status szb_ctx_set_fused_resample(szb_ctx* ctx, int32_t enable) {
    SZB_REQUIRE(ctx, "szb_ctx_set_fused_resample: ctx is NULL");
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->fuse_resample = enable != 0;
    return SZB_OK;
}
szb_status szb_ctx_set_l2_ring(szb_ctx* ctx, int32_t chunk_mb, int32_t streams) {
    SZB_REQUIRE(ctx && chunk_mb >= 0 && (streams == 1 || streams == 2), "szb_ctx_set_l2_ring: chunk_mb >= 0, streams 1 or 2");
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->l2_chunk_mb = chunk_mb;
    ctx->l2_streams = streams;
    return SZB_OK;
}
uint64_t szb_ctx_launch_count(const szb_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t szb_ctx_graph_launch_count(const szb_ctx* ctx) { return ctx ? ctx->graph_launches : 0; }

szb_status szb_timer_start(szb_ctx* ctx) {
    SZB_REQUIRE(ctx, "szb_timer_start: ctx is NULL");
    SZB_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    return SZB_OK;
}
szb_status szb_timer_stop(szb_ctx* ctx, float* elapsed_ms) {
    SZB_REQUIRE(ctx && elapsed_ms, "szb_timer_stop: NULL argument");
    SZB_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    SZB_CUDA(cudaEventSynchronize(ctx->ev_stop));
    SZB_CUDA(cudaEventElapsedTime(elapsed_ms, ctx->ev_start, ctx->ev_stop));
    return SZB_OK;
}
szb_status szb_kernel_timing(szb_ctx* ctx, int32_t enable) {
    SZB_REQUIRE(ctx, "szb_kernel_timing: ctx is NULL");
    ctx->ktime_on = enable != 0;
    return SZB_OK;
}
szb_status szb_kernel_timing_read(szb_ctx* ctx, double* total_ms, uint64_t* launches, int32_t reset) {
    SZB_REQUIRE(ctx, "szb_kernel_timing_read: ctx is NULL");
    SZB_TRY(drain_ktime(ctx));
    if (total_ms) *total_ms = ctx->ktime_ms;
    if (launches) *launches = ctx->ktime_launches;
    if (reset) {
        ctx->ktime_ms = 0.0;
        ctx->ktime_launches = 0;
    }
    return SZB_OK;
}

szb_status szb_dev_alloc(szb_ctx* ctx, size_t bytes, void** dptr) {
    SZB_REQUIRE(ctx && dptr, "szb_dev_alloc: NULL argument");
    SZB_CUDA(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return SZB_ERR_ALLOC;
    }
    return SZB_OK;
}
szb_status szb_dev_free(szb_ctx* ctx, void* dptr) {
    SZB_REQUIRE(ctx, "szb_dev_free: ctx is NULL");
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    SZB_CUDA(cudaFree(dptr));
    return SZB_OK;
}
szb_status szb_memcpy_h2d(szb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    SZB_REQUIRE(ctx, "szb_memcpy_h2d: ctx is NULL");
    SZB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}
szb_status szb_memcpy_d2h(szb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    SZB_REQUIRE(ctx, "szb_memcpy_d2h: ctx is NULL");
    SZB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}
szb_status szb_host_alloc_pinned(size_t bytes, void** hptr) {
    SZB_REQUIRE(hptr, "szb_host_alloc_pinned: NULL argument");
    cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        set_error("cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return SZB_ERR_ALLOC;
    }
    return SZB_OK;
}
szb_status szb_host_free_pinned(void* hptr) {
    SZB_CUDA(cudaFreeHost(hptr));
    return SZB_OK;
}

// ---- tables ---------------------------------------------------------------------------------------------------------
szb_status szb_table_mel(float* out) {
    SZB_REQUIRE(out, "szb_table_mel: out is NULL");
    const auto d = mel_filterbank_dense();
    std::memcpy(out, d.data(), d.size() * sizeof(float));
    return SZB_OK;
}
szb_status szb_table_dct(float* out) {
    SZB_REQUIRE(out, "szb_table_dct: out is NULL");
    const auto d = dct2_rows();
    std::memcpy(out, d.data(), d.size() * sizeof(float));
    return SZB_OK;
}
szb_status szb_table_resample_taps(uint32_t rate, float* out, uint32_t* L, uint32_t* M) {
    SZB_REQUIRE(rate > 0, "szb_table_resample_taps: rate is 0");
    uint32_t l, m;
    resample_ratio(rate, l, m);
    if (L) *L = l;
    if (M) *M = m;
    if (out) {
        const auto c = resample_taps(rate);
        std::memcpy(out, c.data(), c.size() * sizeof(float));
    }
    return SZB_OK;
}

// ---- front end ------------------------------------------------------------------------------------------------------
uint64_t szb_num_windows(uint64_t n) { return n < SZB_WINDOW_SIZE ? 0 : (n - SZB_WINDOW_SIZE) / SZB_HOP_SIZE + 1; }
uint64_t szb_resample_out_len(uint64_t n_in, uint32_t rate) {
    if (rate == 0) return 0;
    return uint64_t((unsigned __int128)n_in * SZB_SAMPLE_RATE / rate);
}

uint64_t szb_extract_batch_windows(const uint64_t* clip_off, uint32_t n_clips, uint32_t rate) {
    if (!clip_off || rate == 0) return 0;
    uint64_t total = 0;
    for (uint32_t c = 0; c < n_clips; ++c) {
        const uint64_t n_in = clip_off[c + 1] - clip_off[c];
        total += szb_num_windows(rate == SZB_SAMPLE_RATE ? n_in : szb_resample_out_len(n_in, rate));
    }
    return total;
}

szb_status szb_downmix_to_mono(szb_ctx* ctx, const int16_t* in, uint64_t n, uint32_t channels, int16_t* mono,
                               uint64_t mono_cap, uint64_t* n_mono) {
    SZB_REQUIRE(ctx && n_mono && (in || n == 0), "szb_downmix_to_mono: NULL argument");
    const uint64_t ch = channels <= 1 ? 1 : channels;
    const uint64_t n_out = (n + ch - 1) / ch;
    *n_mono = n_out;
    SZB_REQUIRE(mono_cap >= n_out && (mono || n_out == 0), "szb_downmix_to_mono: capacity %llu < %llu",
                (unsigned long long)mono_cap, (unsigned long long)n_out);
    if (n_out == 0) return SZB_OK;
    if (ch == 1) {  // lib.rs:173-175
        std::memcpy(mono, in, n * sizeof(int16_t));
        return SZB_OK;
    }
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->pcm.reserve(n * 2));
    SZB_TRY(ctx->misc.reserve(n_out * 2));
    SZB_CUDA(cudaMemcpyAsync(ctx->pcm.ptr, in, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(launch_downmix(ctx, ctx->pcm.as<int16_t>(), n, uint32_t(ch), ctx->misc.as<int16_t>(), n_out));
    SZB_CUDA(cudaMemcpyAsync(mono, ctx->misc.ptr, n_out * 2, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

static unsigned long long host_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// augment (lib.rs:103-116).  Clip-level draws from the seed: noise_level U(0, 0.005), gain U(0.95, 1.05),
// shift in [0, min(len, 800)) (lib.rs:105-107); the reference uses an unseeded thread_rng.
szb_status szb_augment_params(uint64_t seed, uint64_t n_samples, float* noise_level, float* gain, uint64_t* shift) {
    SZB_REQUIRE(noise_level && gain && shift, "szb_augment_params: NULL argument");
    const unsigned long long k = host_splitmix64(seed ^ 0xA06DE27ull);
    const float u0 = float(uint32_t(host_splitmix64(k ^ 1) >> 40)) * 5.9604644775390625e-08f;
    const float u1 = float(uint32_t(host_splitmix64(k ^ 2) >> 40)) * 5.9604644775390625e-08f;
    *noise_level = 0.005f * u0;
    *gain = 0.95f + 0.1f * u1;
    const uint64_t range = std::min<uint64_t>(n_samples, SZB_WINDOW_SIZE);
    *shift = range ? host_splitmix64(k ^ 3) % range : 0;
    return SZB_OK;
}

szb_status szb_augment_dev(szb_ctx* ctx, const int16_t* d_in, uint64_t n, uint64_t seed, int16_t* d_out) {
    SZB_REQUIRE(ctx && (n == 0 || (d_in && d_out)), "szb_augment_dev: NULL argument");
    SZB_REQUIRE(d_in != d_out || n == 0, "szb_augment_dev: in-place augmentation is not supported (circular shift)");
    SZB_CUDA(cudaSetDevice(ctx->device));
    float nl, gain;
    uint64_t shift;
    SZB_TRY(szb_augment_params(seed, n, &nl, &gain, &shift));
    return launch_augment(ctx, d_in, n, shift, gain, nl, host_splitmix64(seed ^ 0x5EEDull), d_out);
}

szb_status szb_augment(szb_ctx* ctx, const int16_t* in, uint64_t n, uint64_t seed, int16_t* out) {
    SZB_REQUIRE(ctx && (n == 0 || (in && out)), "szb_augment: NULL argument");
    if (n == 0) return SZB_OK;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->pcm.reserve(n * 2));
    SZB_TRY(ctx->misc.reserve(n * 2));
    SZB_CUDA(cudaMemcpyAsync(ctx->pcm.ptr, in, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(szb_augment_dev(ctx, ctx->pcm.as<int16_t>(), n, seed, ctx->misc.as<int16_t>()));
    SZB_CUDA(cudaMemcpyAsync(out, ctx->misc.ptr, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

szb_status szb_resample_to_44100(szb_ctx* ctx, const int16_t* in, uint64_t n_in, uint32_t rate, int16_t* out,
                                 uint64_t out_cap, uint64_t* n_out) {
    SZB_REQUIRE(ctx && n_out && (in || n_in == 0), "szb_resample_to_44100: NULL argument");
    SZB_REQUIRE(rate > 0, "szb_resample_to_44100: rate is 0");
    const uint64_t n = rate == SZB_SAMPLE_RATE ? n_in : szb_resample_out_len(n_in, rate);
    *n_out = n;
    SZB_REQUIRE(out_cap >= n && (out || n == 0), "szb_resample_to_44100: capacity %llu < %llu",
                (unsigned long long)out_cap, (unsigned long long)n);
    if (n == 0) return SZB_OK;
    if (rate == SZB_SAMPLE_RATE) {  // lib.rs:187-189
        std::memcpy(out, in, n * sizeof(int16_t));
        return SZB_OK;
    }
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->pcm.reserve(n_in * 2));
    SZB_TRY(ctx->misc.reserve(n * 2 + 64));
    SZB_TRY(ctx->labels.reserve(4 * sizeof(uint64_t)));
    const uint64_t offs[4] = { 0, n_in, 0, n };
    SZB_CUDA(cudaMemcpyAsync(ctx->labels.ptr, offs, sizeof offs, cudaMemcpyHostToDevice, ctx->stream));
    SZB_CUDA(cudaMemcpyAsync(ctx->pcm.ptr, in, n_in * 2, cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(launch_resample(ctx, ctx->pcm.as<int16_t>(), ctx->labels.as<uint64_t>(), ctx->labels.as<uint64_t>() + 2, 1, n,
                            rate, ctx->misc.as<int16_t>()));
    SZB_CUDA(cudaMemcpyAsync(out, ctx->misc.ptr, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

// Shared driver of the two batch entry points.  With `h_pcm` / `h_feats` set (host entry point) the batch is cut into
// chunks of clips that flow through a three-stage pipeline -- H2D copy (copy_in stream), resample + extract (the
// context's stream), D2H copy (copy_out stream) -- so the PCIe transfers of neighbouring chunks overlap the kernels.
static szb_status extract_batch_impl(szb_ctx* ctx, const int16_t* d_pcm, const int16_t* h_pcm, const uint64_t* clip_off,
                                     uint32_t n_clips, uint32_t rate, float* d_feats, float* h_feats, uint64_t cap_windows,
                                     uint64_t* win_off) {
    for (uint32_t c = 0; c < n_clips; ++c)
        SZB_REQUIRE(clip_off[c + 1] >= clip_off[c], "extract_batch: clip_off not monotone at %u", c);
    SZB_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint64_t> off44, woff;
    batch_layout(clip_off, n_clips, rate, off44, woff);
    std::memcpy(win_off, woff.data(), woff.size() * sizeof(uint64_t));
    const uint64_t total = woff[n_clips];
    SZB_REQUIRE(cap_windows >= total, "extract_batch: capacity %llu windows < %llu", (unsigned long long)cap_windows,
                (unsigned long long)total);
    if (total == 0) return SZB_OK;
    SZB_REQUIRE(d_pcm && d_feats, "extract_batch: NULL buffer");
    const bool resample = rate != SZB_SAMPLE_RATE;
    const bool piped = h_pcm != nullptr;
    const uint64_t first = clip_off[0];

    // chunking: ~48 MB of traffic per chunk for the host call (PCIe pipeline).  Device-resident call with a resampler in
    // front: chunks whose 44.1 kHz intermediate is a few tens of MB, written by the resampler into a two-slot ring that never
    // leaves the 126 MB L2 -- the extraction kernel of the chunk reads it back from L2 and the slot is overwritten two chunks
    // later, still dirty in cache, so the 2 x 8.8 GB round trip through HBM of the unchunked pipeline disappears (ncu DRAM
    // bytes: profiles/).  Device-resident call at 44.1 kHz: one chunk.
    // fused mode: no resampler launch at all -- the extraction kernel filters every 33-hop tile from the original-rate clip in
    // shared memory (frontend.cu).  Needs 16-byte aligned clip starts (cp.async) and a supported rate.
    bool fused = resample && ctx->fuse_resample && fused_resample_supported(rate) && (reinterpret_cast<uintptr_t>(d_pcm) & 15) == 0;
    for (uint32_t c = 0; c < n_clips && fused; ++c)
        if (woff[c + 1] > woff[c] && (((clip_off[c] - first) & 7) != 0 || woff[c + 1] - woff[c] > 10000000ull)) fused = false;
    const uint64_t l2_chunk_bytes = uint64_t(ctx->l2_chunk_mb) << 20;
    const bool ring = resample && !fused && !piped && l2_chunk_bytes > 0;
    std::vector<uint32_t> chunk_begin{ 0 };
    if (ring) {
        uint64_t acc = 0;
        for (uint32_t c = 0; c < n_clips; ++c) {
            const uint64_t b = (off44[c + 1] - off44[c]) * 2;
            if (acc > 0 && acc + b > l2_chunk_bytes) {
                chunk_begin.push_back(c);
                acc = 0;
            }
            acc += b;
        }
    } else if (piped) {
        const uint64_t target = 48ull << 20;
        uint64_t acc = 0;
        for (uint32_t c = 0; c < n_clips; ++c) {
            acc += (clip_off[c + 1] - clip_off[c]) * 2 + (woff[c + 1] - woff[c]) * SZB_FEATURE_SIZE * 4;
            if (acc >= target && c + 1 < n_clips) {
                chunk_begin.push_back(c + 1);
                acc = 0;
            }
        }
    }
    chunk_begin.push_back(n_clips);
    const size_t n_chunks = chunk_begin.size() - 1;

    const int16_t* d_pcm44 = d_pcm;   // device layout: clip c starts at d_pcm[clip_off[c] - first] (rate 44.1k) ...
    uint64_t* d_in_off = nullptr;
    uint64_t* d_out_off = nullptr;
    std::vector<uint64_t> off_ring;   // ring mode: clip c's 44.1 kHz samples start at misc[off_ring[c]] (slot = chunk & 1)
    if (resample && !fused) {         // ... or at misc[off44[c]] after the resampler
        if (ring) {
            uint64_t slot = 0;
            for (size_t k = 0; k + 1 < chunk_begin.size(); ++k) slot = std::max(slot, off44[chunk_begin[k + 1]] - off44[chunk_begin[k]]);
            slot = (slot + 127) & ~uint64_t(127);
            off_ring.resize(size_t(n_clips) + 1, 0);
            for (size_t k = 0; k + 1 < chunk_begin.size(); ++k)
                for (uint32_t c = chunk_begin[k]; c < chunk_begin[k + 1]; ++c)
                    off_ring[c] = (k & 1) * slot + (off44[c] - off44[chunk_begin[k]]);
            SZB_TRY(ctx->misc.reserve(2 * slot * 2 + 64));
        } else {
            SZB_TRY(ctx->misc.reserve(off44[n_clips] * 2 + 64));
        }
        SZB_TRY(ctx->labels.reserve((size_t(n_clips) + 1) * 2 * sizeof(uint64_t)));
        void* hp = nullptr;
        SZB_TRY(ctx->h_stage.acquire((size_t(n_clips) + 1) * 2 * sizeof(uint64_t), &hp));
        uint64_t* h = static_cast<uint64_t*>(hp);
        for (uint32_t c = 0; c <= n_clips; ++c) {
            h[c] = clip_off[c] - first;
            h[n_clips + 1 + c] = ring ? off_ring[c] : off44[c];
        }
        SZB_CUDA(cudaMemcpyAsync(ctx->labels.ptr, h, (size_t(n_clips) + 1) * 2 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
        SZB_TRY(ctx->h_stage.uploaded(ctx->stream));
        d_in_off = ctx->labels.as<uint64_t>();
        d_out_off = d_in_off + n_clips + 1;
        d_pcm44 = ctx->misc.as<int16_t>();
    }
    // segment table of every chunk, uploaded once
    std::vector<Segment> segs;
    std::vector<size_t> seg_begin(n_chunks + 1, 0);
    std::vector<uint64_t> seg_off(size_t(n_clips) + 1);
    for (uint32_t c = 0; c <= n_clips; ++c) seg_off[c] = (resample && !fused) ? (ring ? off_ring[c] : off44[c]) : clip_off[c] - first;
    std::vector<uint64_t> n_in_of;
    if (fused) {
        n_in_of.resize(n_clips);
        for (uint32_t c = 0; c < n_clips; ++c) n_in_of[c] = clip_off[c + 1] - clip_off[c];
    }
    for (size_t k = 0; k < n_chunks; ++k) {
        seg_begin[k] = segs.size();
        build_segments(seg_off.data(), woff.data(), chunk_begin[k], chunk_begin[k + 1], ctx->sm_count, segs, fused ? n_in_of.data() : nullptr);
    }
    seg_begin[n_chunks] = segs.size();
    SZB_TRY(upload_segments(ctx, segs, uint32_t(n_chunks)));
    // 128-bit PCM loads need every clip to start on a 16-byte boundary (always true after the resampler: its output
    // rows are padded to 8 samples; true for caller-packed 44.1 kHz batches whose offsets are multiples of 8)
    bool aligned16 = (reinterpret_cast<uintptr_t>(d_pcm44) & 15) == 0;
    for (uint32_t c = 0; c < n_clips && aligned16; ++c)
        if (woff[c + 1] > woff[c] && (seg_off[c] & 7) != 0) aligned16 = false;

    if (piped) {
        while (ctx->pipe_events.size() < 2 * n_chunks + 1) {
            cudaEvent_t e;
            SZB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->pipe_events.push_back(e);
        }
        // the copy streams must not start before the uploads above (and any earlier work on the stream) are done
        SZB_CUDA(cudaEventRecord(ctx->pipe_events[2 * n_chunks], ctx->stream));
        SZB_CUDA(cudaStreamWaitEvent(ctx->copy_in, ctx->pipe_events[2 * n_chunks], 0));
        SZB_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_events[2 * n_chunks], 0));
    }
    for (size_t k = 0; k < n_chunks; ++k) {
        const uint32_t c0 = chunk_begin[k], c1 = chunk_begin[k + 1];
        if (piped) {
            const uint64_t s0 = clip_off[c0] - first, s1 = clip_off[c1] - first;
            if (s1 > s0)
                SZB_CUDA(cudaMemcpyAsync(const_cast<int16_t*>(d_pcm) + s0, h_pcm + clip_off[c0], (s1 - s0) * 2, cudaMemcpyHostToDevice,
                                         ctx->copy_in));
            SZB_CUDA(cudaEventRecord(ctx->pipe_events[2 * k], ctx->copy_in));
            SZB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->pipe_events[2 * k], 0));
        }
        if (resample && !fused) {
            uint64_t max_out = 0;
            for (uint32_t c = c0; c < c1; ++c) max_out = std::max(max_out, szb_resample_out_len(clip_off[c + 1] - clip_off[c], rate));
            if (ring && ctx->l2_streams > 1) {
                // resampler of chunk k on the side stream: it may run under the tail of extract(k - 1); it must wait for
                // extract(k - 2), the last reader of its ring slot, and for the tables uploaded on the main stream
                cudaStream_t side = ctx->copy_in;
                if (k == 0) {
                    SZB_CUDA(cudaEventRecord(ctx->ring_ev[4], ctx->stream));
                    SZB_CUDA(cudaStreamWaitEvent(side, ctx->ring_ev[4], 0));
                }
                if (k >= 2) SZB_CUDA(cudaStreamWaitEvent(side, ctx->ring_ev[2 + (k & 1)], 0));     // extract(k - 2) done
                cudaStream_t main_stream = ctx->stream;
                ctx->stream = side;
                const szb_status st = launch_resample(ctx, d_pcm, d_in_off + c0, d_out_off + c0, c1 - c0, max_out, rate, ctx->misc.as<int16_t>());
                ctx->stream = main_stream;
                SZB_TRY(st);
                SZB_CUDA(cudaEventRecord(ctx->ring_ev[k & 1], side));
                SZB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ring_ev[k & 1], 0));
            } else {
                SZB_TRY(launch_resample(ctx, d_pcm, d_in_off + c0, d_out_off + c0, c1 - c0, max_out, rate, ctx->misc.as<int16_t>()));
            }
        }
        SZB_TRY(launch_extract(ctx, d_pcm44, seg_begin[k], seg_begin[k + 1] - seg_begin[k], uint32_t(k), d_feats, aligned16, fused ? rate : 0u));
        if (ring && ctx->l2_streams > 1) SZB_CUDA(cudaEventRecord(ctx->ring_ev[2 + (k & 1)], ctx->stream));
        if (piped) {
            SZB_CUDA(cudaEventRecord(ctx->pipe_events[2 * k + 1], ctx->stream));
            SZB_CUDA(cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_events[2 * k + 1], 0));
            const uint64_t w0 = woff[c0], w1 = woff[c1];
            if (w1 > w0)
                SZB_CUDA(cudaMemcpyAsync(h_feats + w0 * SZB_FEATURE_SIZE, d_feats + w0 * SZB_FEATURE_SIZE,
                                         (w1 - w0) * SZB_FEATURE_SIZE * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
        }
    }
    if (piped) {
        SZB_CUDA(cudaStreamSynchronize(ctx->copy_out));
        SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SZB_OK;
}

szb_status szb_extract_batch_dev(szb_ctx* ctx, const int16_t* d_pcm, const uint64_t* clip_off, uint32_t n_clips,
                                 uint32_t rate, float* d_feats, uint64_t cap_windows, uint64_t* win_off) {
    SZB_REQUIRE(ctx && clip_off && win_off, "szb_extract_batch_dev: NULL argument");
    SZB_REQUIRE(rate > 0, "szb_extract_batch_dev: rate is 0");
    return extract_batch_impl(ctx, d_pcm ? d_pcm + clip_off[0] : nullptr, nullptr, clip_off, n_clips, rate, d_feats, nullptr,
                              cap_windows, win_off);
}

szb_status szb_extract_batch(szb_ctx* ctx, const int16_t* pcm, const uint64_t* clip_off, uint32_t n_clips, uint32_t rate,
                             float* feats, uint64_t cap_windows, uint64_t* win_off) {
    SZB_REQUIRE(ctx && clip_off && win_off, "szb_extract_batch: NULL argument");
    SZB_REQUIRE(rate > 0, "szb_extract_batch: rate is 0");
    const uint64_t total = szb_extract_batch_windows(clip_off, n_clips, rate);
    SZB_REQUIRE(cap_windows >= total, "szb_extract_batch: capacity %llu windows < %llu", (unsigned long long)cap_windows,
                (unsigned long long)total);
    const uint64_t first = n_clips ? clip_off[0] : 0, last = n_clips ? clip_off[n_clips] : 0;
    const uint64_t n_samples = last - first;
    if (total == 0 || n_samples == 0) {
        win_off[0] = 0;
        for (uint32_t c = 0; c < n_clips; ++c) win_off[c + 1] = 0;
        return SZB_OK;
    }
    SZB_REQUIRE(pcm && feats, "szb_extract_batch: NULL buffer");
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->pcm.reserve(n_samples * 2 + 64));
    SZB_TRY(ctx->feats.reserve(total * SZB_FEATURE_SIZE * sizeof(float)));
    return extract_batch_impl(ctx, ctx->pcm.as<int16_t>(), pcm, clip_off, n_clips, rate, ctx->feats.as<float>(), feats, cap_windows,
                              win_off);
}

szb_status szb_extract(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, float* feats, uint64_t cap_windows,
                       uint64_t* n_windows) {
    SZB_REQUIRE(ctx && n_windows, "szb_extract: NULL argument");
    const uint64_t off[2] = { 0, n_samples };
    uint64_t woff[2] = { 0, 0 };
    *n_windows = szb_num_windows(n_samples);
    SZB_TRY(szb_extract_batch(ctx, pcm, off, 1, SZB_SAMPLE_RATE, feats, cap_windows, woff));
    return SZB_OK;
}

// Shared body of szb_extract_range(_dev): stages the samples of frames [f_lo, f_hi) and runs one segment whose PCM origin
// is the (virtual) start of the clip, so the kernel indexes hops exactly as it does for a whole clip.
static szb_status extract_range_impl(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t w_begin, uint64_t w_end,
                                     float* d_out) {
    const uint64_t n_total = szb_num_windows(n_samples);
    const uint64_t f_lo = w_begin >= 2 ? w_begin - 2 : 0, f_hi = std::min<uint64_t>(w_end + 2, n_total);
    const uint64_t s_lo = f_lo * 400, s_hi = (f_hi - 1) * 400 + 800;         // samples the frames f_lo .. f_hi - 1 cover
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->pcm.reserve((s_hi - s_lo) * 2 + 64));
    SZB_CUDA(cudaMemcpyAsync(ctx->pcm.ptr, pcm + s_lo, (s_hi - s_lo) * 2, cudaMemcpyHostToDevice, ctx->stream));
    Segment sg{};
    sg.pcm_off = 0ull - s_lo;            // (modular) offset of the clip's sample 0 relative to the staged buffer
    sg.out_row = 0ull - w_begin;         // row w lands at d_out[(w - w_begin) * 60]
    sg.n_total = uint32_t(n_total);
    sg.w_begin = uint32_t(w_begin);
    sg.w_end = uint32_t(w_end);
    std::vector<Segment> segs{ sg };
    SZB_TRY(upload_segments(ctx, segs, 1));
    // s_lo is a multiple of 400 samples = 800 bytes = 50 x 16: the staged buffer keeps the clip's 16-byte phase
    SZB_TRY(launch_extract(ctx, ctx->pcm.as<int16_t>(), 0, 1, 0, d_out, true));
    return SZB_OK;
}

szb_status szb_extract_range_dev(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t w_begin, uint64_t w_end,
                                 float* d_feats, uint64_t cap_windows) {
    SZB_REQUIRE(ctx, "szb_extract_range_dev: ctx is NULL");
    const uint64_t n_total = szb_num_windows(n_samples);
    SZB_REQUIRE(w_begin <= w_end && w_end <= n_total, "szb_extract_range_dev: windows [%llu, %llu) outside the clip's %llu",
                (unsigned long long)w_begin, (unsigned long long)w_end, (unsigned long long)n_total);
    SZB_REQUIRE(n_total < (1ull << 32), "szb_extract_range_dev: clip too long");
    SZB_REQUIRE(cap_windows >= w_end - w_begin, "szb_extract_range_dev: capacity %llu windows < %llu",
                (unsigned long long)cap_windows, (unsigned long long)(w_end - w_begin));
    if (w_end == w_begin) return SZB_OK;
    SZB_REQUIRE(pcm && d_feats, "szb_extract_range_dev: NULL buffer");
    return extract_range_impl(ctx, pcm, n_samples, w_begin, w_end, d_feats);
}

szb_status szb_extract_range(szb_ctx* ctx, const int16_t* pcm, uint64_t n_samples, uint64_t w_begin, uint64_t w_end, float* feats,
                             uint64_t cap_windows) {
    SZB_REQUIRE(ctx, "szb_extract_range: ctx is NULL");
    const uint64_t n = w_end >= w_begin ? w_end - w_begin : 0;
    if (n) SZB_TRY(ctx->feats.reserve(n * SZB_FEATURE_SIZE * sizeof(float)));
    SZB_TRY(szb_extract_range_dev(ctx, pcm, n_samples, w_begin, w_end, ctx->feats.as<float>(), cap_windows));
    if (n == 0) return SZB_OK;
    SZB_REQUIRE(feats, "szb_extract_range: NULL buffer");
    SZB_CUDA(cudaMemcpyAsync(feats, ctx->feats.ptr, n * SZB_FEATURE_SIZE * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

}  // extern "C"
