// Multi-GPU plumbing: one process per GPU; batch-parallel training all-reduces the flattened gradient (+ the
// [n_used, loss] tail) once per step with NCCL over NVLink/NVSwitch -- the only collective of the path
// (BASELINE.json north_star).  Extraction shards clips across ranks and needs no collective.
//
// NCCL is bound at run time with dlopen so the library has no link-time NCCL dependency: a host that never calls
// szb_comm_* (single GPU, or the Rust CLI) does not need libnccl at all, and inside a PyTorch process the already
// loaded libnccl.so.2 is reused instead of a second copy.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace szb {

struct Id128 { char bytes[128]; };  // ncclUniqueId, passed by value

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static szb_status load_nccl() {
    if (g_nccl.handle) return SZB_OK;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return SZB_ERR_NCCL;
    }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        set_error("libnccl is missing a required symbol");
        dlclose(h);
        return SZB_ERR_NCCL;
    }
    g_nccl.handle = h;
    return SZB_OK;
}

#define SZB_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != 0) {                                                                                   \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");  \
            return SZB_ERR_NCCL;                                                                         \
        }                                                                                                \
    } while (0)

szb_status comm_allreduce_f32(szb_ctx* ctx, float* buf, size_t n) {
    if (ctx->world <= 1) return SZB_OK;
    SZB_REQUIRE(ctx->nccl_comm, "all-reduce requested but szb_comm_init was not called");
    // ncclFloat32 = 7, ncclSum = 0
    SZB_NCCL(g_nccl.AllReduce(buf, buf, n, 7, 0, ctx->nccl_comm, ctx->stream));
    ctx->launches += 1;
    return SZB_OK;
}

// All-reduce of one gradient slice on the communication stream, ordered after everything enqueued so far on the
// context's stream: the reduction of layer-3 gradients then runs while the backward GEMMs of layers 2 and 1 execute.
szb_status comm_allreduce_overlapped(szb_ctx* ctx, float* buf, size_t n) {
    if (ctx->world <= 1 || n == 0) return SZB_OK;
    SZB_REQUIRE(ctx->nccl_comm, "all-reduce requested but szb_comm_init was not called");
    if (!ctx->comm_stream) SZB_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    if (!ctx->ev_comm) SZB_CUDA(cudaEventCreateWithFlags(&ctx->ev_comm, cudaEventDisableTiming));
    SZB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->stream));
    SZB_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_comm, 0));
    SZB_NCCL(g_nccl.AllReduce(buf, buf, n, 7, 0, ctx->nccl_comm, ctx->comm_stream));
    ctx->launches += 1;
    return SZB_OK;
}

// The context's stream waits for every overlapped all-reduce issued so far.
szb_status comm_join(szb_ctx* ctx) {
    if (ctx->world <= 1 || !ctx->comm_stream) return SZB_OK;
    SZB_CUDA(cudaEventRecord(ctx->ev_comm, ctx->comm_stream));
    SZB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_comm, 0));
    return SZB_OK;
}

// ---- gradient exchange over peer memory ------------------------------------------------------------------------------
// Region layout per rank: [flag block, 256 B: u32 words 0-15 flag1 per source rank, 16-31 flag2, 48 spare (teardown
// rendezvous)][inbox: cap floats][red: cap floats]  (protocol: mlp.cu, p2p_exchange).
constexpr size_t kP2pCapFloats = size_t(1) << 20;   // 4 MB per buffer: nets up to ~1 M parameters (C = 1000: 420 k)
constexpr size_t kP2pFlagBytes = 256;
constexpr size_t kP2pSpareWord = 48;

static void p2p_teardown(szb_ctx* ctx) {
    for (int r = 0; r < ctx->world && r < szb_ctx::kMaxPeers; ++r) {
        if (r != ctx->rank && ctx->p2p_flags[r]) cudaIpcCloseMemHandle(ctx->p2p_flags[r]);
        ctx->p2p_flags[r] = nullptr;
        ctx->p2p_inbox[r] = nullptr;
        ctx->p2p_red[r] = nullptr;
    }
    if (ctx->p2p_region) cudaFree(ctx->p2p_region);
    ctx->p2p_region = nullptr;
    ctx->p2p_on = false;
    ctx->p2p_cap = 0;
}

// Maps every rank's region into this process.  Any failure on any rank leaves ALL ranks on the NCCL path (the outcome is
// agreed with a min all-reduce), so the ranks never disagree about which exchange a step uses.
static szb_status p2p_setup(szb_ctx* ctx) {
    const int W = ctx->world, me = ctx->rank;
    int ok = (W <= szb_ctx::kMaxPeers && g_nccl.AllGather) ? 1 : 0;
    // two-shot: [inbox: cap][red: cap]; one-shot: [2 step parities][W source ranks][cap]
    const size_t bytes = kP2pFlagBytes + 2 * size_t(W) * kP2pCapFloats * sizeof(float);
    cudaIpcMemHandle_t mine{};
    if (ok && cudaMalloc(&ctx->p2p_region, bytes) != cudaSuccess) { ctx->p2p_region = nullptr; ok = 0; }
    if (ok && cudaMemset(ctx->p2p_region, 0, bytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, ctx->p2p_region) != cudaSuccess) ok = 0;
    cudaGetLastError();
    // exchange the handles (and the ok flags) through NCCL itself: no extra plumbing between the processes
    struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
    static_assert(sizeof(Msg) == 128, "Msg size");
    Msg msg{}; msg.h = mine; msg.ok = ok;
    Msg* d_all = nullptr;
    SZB_CUDA(cudaMalloc(&d_all, sizeof(Msg) * size_t(W + 1)));
    SZB_CUDA(cudaMemcpy(d_all + W, &msg, sizeof msg, cudaMemcpyHostToDevice));
    if (g_nccl.AllGather) {
        SZB_NCCL(g_nccl.AllGather(d_all + W, d_all, sizeof(Msg), /*ncclUint8*/ 1, ctx->nccl_comm, ctx->stream));
    }
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<Msg> all(static_cast<size_t>(W));
    SZB_CUDA(cudaMemcpy(all.data(), d_all, sizeof(Msg) * size_t(W), cudaMemcpyDeviceToHost));
    for (int r = 0; r < W; ++r) ok = ok && all[size_t(r)].ok;
    if (ok) {
        for (int r = 0; r < W && ok; ++r) {
            void* base = ctx->p2p_region;
            if (r != me && cudaIpcOpenMemHandle(&base, all[size_t(r)].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            ctx->p2p_flags[r] = static_cast<uint32_t*>(base);
            ctx->p2p_inbox[r] = reinterpret_cast<float*>(static_cast<char*>(base) + kP2pFlagBytes);
            ctx->p2p_red[r] = ctx->p2p_inbox[r] + kP2pCapFloats;
        }
        cudaGetLastError();
    }
    // agree on the outcome
    float* d_ok = reinterpret_cast<float*>(d_all);
    const float f_ok = ok ? 1.f : 0.f;
    SZB_CUDA(cudaMemcpy(d_ok, &f_ok, sizeof f_ok, cudaMemcpyHostToDevice));
    SZB_NCCL(g_nccl.AllReduce(d_ok, d_ok, 1, /*ncclFloat32*/ 7, /*ncclMin*/ 3, ctx->nccl_comm, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    float agreed = 0.f;
    SZB_CUDA(cudaMemcpy(&agreed, d_ok, sizeof agreed, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    if (agreed < 0.5f) {
        p2p_teardown(ctx);
        return SZB_OK;
    }
    SZB_TRY(ctx->p2p_counters.reserve(1280));           // [0, 8): tickets; [64, 128): phase trace (8 x u64); [256, 1280): per-tile tickets of
    SZB_CUDA(cudaMemset(ctx->p2p_counters.ptr, 0, 1280));  // the grouped weight-gradient launch (early push, gemm_tma.cuh)
    ctx->p2p_on = true;
    ctx->p2p_cap = kP2pCapFloats;
    ctx->p2p_step = 0;
    return SZB_OK;
}

}  // namespace szb

using namespace szb;

extern "C" {

szb_status szb_comm_unique_id(uint8_t id[128]) {
    SZB_REQUIRE(id, "szb_comm_unique_id: id is NULL");
    SZB_TRY(load_nccl());
    SZB_NCCL(g_nccl.GetUniqueId(id));
    return SZB_OK;
}

szb_status szb_comm_init(szb_ctx* ctx, const uint8_t id[128], int32_t rank, int32_t world) {
    SZB_REQUIRE(ctx && id, "szb_comm_init: NULL argument");
    SZB_REQUIRE(world >= 1 && rank >= 0 && rank < world, "szb_comm_init: bad rank %d / world %d", rank, world);
    SZB_REQUIRE(!ctx->nccl_comm, "szb_comm_init: communicator already initialised");
    ctx->rank = rank;
    ctx->world = world;
    if (world == 1) return SZB_OK;
    SZB_TRY(load_nccl());
    SZB_CUDA(cudaSetDevice(ctx->device));
    Id128 uid;
    std::memcpy(uid.bytes, id, 128);
    SZB_NCCL(g_nccl.CommInitRank(&ctx->nccl_comm, world, uid, rank));
    return SZB_OK;
}

szb_status szb_comm_peer_exchange(szb_ctx* ctx, int32_t enable, int32_t* active) {
    SZB_REQUIRE(ctx, "szb_comm_peer_exchange: ctx is NULL");
    if (enable >= 1 && enable <= 4) ctx->p2p_mode = enable == 1 ? 0 : enable - 1;      // 1: choose by size, 2: one-shot, 3: two-shot, 4: packets
    if (ctx->world > 1 && ctx->nccl_comm) {
        SZB_CUDA(cudaSetDevice(ctx->device));
        if (enable && !ctx->p2p_on) SZB_TRY(p2p_setup(ctx));
        if (!enable && ctx->p2p_on) {
            cudaStreamSynchronize(ctx->stream);
            float* d = reinterpret_cast<float*>(ctx->p2p_flags[ctx->rank]) + kP2pSpareWord;
            SZB_NCCL(g_nccl.AllReduce(d, d, 1, 7, 0, ctx->nccl_comm, ctx->stream));   // peers may still be reading
            SZB_CUDA(cudaStreamSynchronize(ctx->stream));
            p2p_teardown(ctx);
        }
    }
    if (active) *active = ctx->p2p_on ? 1 : 0;
    return SZB_OK;
}

// Diagnostics of the fused exchange: enable != 0 starts recording (and clears) the phase times of CTA 0 of the update kernel;
// ns[0..5] = mean nanoseconds per step spent in: scatter stores, first publish (fence + ticket + flag), wait for flag1, reduce +
// broadcast, second publish, wait for flag2; ns[6] = 0; ns[7] = steps recorded.
szb_status szb_comm_peer_trace(szb_ctx* ctx, int32_t enable, double* ns) {
    SZB_REQUIRE(ctx, "szb_comm_peer_trace: ctx is NULL");
    if (!ctx->p2p_on) { if (ns) for (int i = 0; i < 8; ++i) ns[i] = 0.0; return SZB_OK; }
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[8] = {};
    unsigned char* base = ctx->p2p_counters.as<unsigned char>() + 64;
    SZB_CUDA(cudaMemcpy(h, base, sizeof h, cudaMemcpyDeviceToHost));
    if (ns) for (int i = 0; i < 8; ++i) ns[i] = i == 7 ? double(h[7]) : (h[7] ? double(h[i]) / double(h[7]) : 0.0);
    SZB_CUDA(cudaMemset(base, 0, sizeof h));
    ctx->p2p_trace_on = enable != 0;
    return SZB_OK;
}

szb_status szb_comm_destroy(szb_ctx* ctx) {
    if (!ctx) return SZB_OK;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaSetDevice(ctx->device);
        if (ctx->p2p_on) {
            // peers may still be reading this rank's gradients: rendezvous before the region goes away
            cudaStreamSynchronize(ctx->stream);
            float* d = reinterpret_cast<float*>(ctx->p2p_flags[ctx->rank]) + kP2pSpareWord;        // spare word of the flag block
            g_nccl.AllReduce(d, d, 1, 7, 0, ctx->nccl_comm, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
        }
        p2p_teardown(ctx);
        g_nccl.CommDestroy(ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
    if (ctx->ev_comm) { cudaEventDestroy(ctx->ev_comm); ctx->ev_comm = nullptr; }
    ctx->world = 1;
    ctx->rank = 0;
    return SZB_OK;
}

int32_t szb_comm_world(const szb_ctx* ctx) { return ctx ? ctx->world : 1; }

}  // extern "C"
