// Small-batch training epochs of SimpleNeuralNet (streamz-rs/src/lib.rs:599-622 around train_batch, lib.rs:1002-1060) as ONE
// persistent cooperative kernel.
//
// Why: the reference trains with batch 8 (main.rs:36), and at that size a step made of eleven tensor-core kernels costs 78 us of
// fixed per-kernel latency for a few microseconds of arithmetic (CUDA-graph replay: 74 us).  Here the whole epoch is
// one launch: the grid walks the steps itself and meets at a grid barrier between the layers -- six barriers per step, FP32
// CUDA-core arithmetic (B <= 32 rows cannot fill a 128-row MMA tile anyway), weights updated in place with no gradient buffer.
//
// Mapping: LANE = BATCH ROW (B <= 32).  Activations and deltas live transposed in a small L2-resident scratch ([feature][32]), so
// a warp reads one feature of all rows with one coalesced load and a weight with one uniform load.  Everything another CTA wrote
// during the kernel (weights, scratch) is read with L1-bypassing loads.
//   A  per column i of W1 (warp task): finish the PREVIOUS step's W1 / b1 update for that column, then h1[:, i] = relu(x W1 + b1)
//   B  per column j of W2 (4 warps split K = 512): h2[:, j] = tanh(h1 W2 + b2)
//   C  per class c (8 warps split K = 256): z[:, c] = h2 W3 + b3          C2  per row: softmax, d3 = p - t, loss (lib.rs:1023-1028)
//   D  per row j of W3: d2[:, j] = (d3 W3^T) (1 - h2^2), then W3[j, :] -= s h2^T d3        (lib.rs:1029-1034, 1051-1052)
//   E  per row i of W2 (8 warps split K = 256): d1[:, i] = (d2 W2^T) [h1 > 0], then W2[i, :] -= s h1^T d2   (lib.rs:1035-1040)
// with s = lr / (number of windows that survived dropout) (lib.rs:1047); a window whose inputs all dropped to zero contributes
// nothing and is not counted (lib.rs:607-609).
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "mlp.cuh"

namespace szb {

constexpr int kSmallB = 32;        // rows per step (lane = row)
constexpr int kSmallThreads = 256; // 8 warps per CTA
constexpr int kSmallMaxC = 256;    // classes: d3^T (C x 32 floats) is staged in shared memory for the W3 update

__device__ __forceinline__ unsigned long long small_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct SmallArgs {
    float* P;                                   // [w1 | b1 | w2 | b2 | w3 | b3]
    size_t off_b1, off_w2, off_b2, off_w3, off_b3;
    int n_in, h1, h2, C;
    const float* feats; const uint32_t* labels; const uint32_t* perm; const uint8_t* keep;
    const uint32_t* step_off;                   // [n_steps + 1] rows of perm per step
    uint32_t n_steps;
    float prob, lr;
    unsigned long long key;
    float* scratch;                             // h1T | h2T | zT (logits, then d3) | d2T | d1T, each [features][32]
    unsigned int* barrier;                      // [0] arrivals (monotonic), [1] generation
    double* stats;                              // [0] += loss, [1] += windows used
};

// Grid barrier for a cooperative launch (all CTAs resident).  Monotonic counters: no reset, no ABA.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int target = (gen + 1) * gridDim.x;
        if (atomicAdd(&bar[0], 1u) + 1 == target) {
            atomicExch(&bar[1], gen + 1);
        } else {
            unsigned long long t0 = 0;
            for (unsigned int spin = 0;; ++spin) {
                unsigned int g;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory");
                if (int(g - (gen + 1)) >= 0) break;
                if ((spin & 0xFFFFu) == 0xFFFFu) {      // a missing CTA would be a bug, not a wait: never hang the GPU
                    unsigned long long t;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                    if (t0 == 0) t0 = t;
                    else if (t - t0 > 20ull * 1000000000ull) __trap();
                }
            }
        }
        __threadfence();
    }
    gen += 1;
    __syncthreads();
}

__global__ void __launch_bounds__(kSmallThreads) train_small_kernel(const SmallArgs a) {
    extern __shared__ float sm[];
    // per-CTA shared memory: the batch (current and previous step) as [k][32] so that lane = row reads are conflict-free,
    // validity / labels, partial sums of the K-split tasks, and the delta tile of the update phases
    float* xs = sm;                                         // [2][n_in][32]
    float* part = xs + 2 * a.n_in * kSmallB;                // [8 warps][32]
    float* dts = part + 8 * kSmallB;                        // [max(h2, C)][33]: d2^T or d3^T of the step
    __shared__ uint32_t s_lab[2][kSmallB];
    __shared__ uint8_t s_valid[2][kSmallB];
    __shared__ float s_nused[2];
    __shared__ float s_hrow[kSmallB];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = blockIdx.x * (kSmallThreads / 32) + warp, nwarps = gridDim.x * (kSmallThreads / 32);
    const int I = a.n_in, H1 = a.h1, H2 = a.h2, C = a.C;
    float* W1 = a.P; float* b1 = a.P + a.off_b1; float* W2 = a.P + a.off_w2; float* b2 = a.P + a.off_b2;
    float* W3 = a.P + a.off_w3; float* b3 = a.P + a.off_b3;
    float* h1T = a.scratch; float* h2T = h1T + size_t(H1) * kSmallB; float* zT = h2T + size_t(H2) * kSmallB;
    float* d2T = zT + size_t(C) * kSmallB; float* d1T = d2T + size_t(H2) * kSmallB;
    unsigned int gen = 0;
    {   // the barrier words carry over from earlier launches of this net: start from the current generation
        unsigned int g;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(a.barrier + 1) : "memory");
        gen = g;
    }
    float scale_prev = 0.f;                                 // lr / n_used of the previous step (its W1 update is still owed)
    int B_prev = 0;

    for (uint32_t step = 0; step <= a.n_steps; ++step) {
        const int cur = step & 1, prv = cur ^ 1;
        const bool live = step < a.n_steps;                 // the extra pass only settles the last step's W1 update
        const uint32_t r0 = live ? a.step_off[step] : 0;
        const int B = live ? int(a.step_off[step + 1] - r0) : 0;
        // ---- gather the step's rows (every CTA keeps its own copy): perm, dropout, all-zero test (lib.rs:604-609) ----------
        if (live) {
            for (int r = warp; r < kSmallB; r += kSmallThreads / 32) {
                bool any = false;
                uint32_t w = 0;
                if (r < B) {
                    w = a.perm[r0 + r];
                    for (int i = lane; i < I; i += 32) {
                        float v = __ldg(a.feats + size_t(w) * I + i);
                        if (a.keep) {
                            if (!a.keep[size_t(w) * I + i]) v = 0.f;
                        } else if (a.prob > 0.f) {
                            const unsigned long long u = small_splitmix64(a.key ^ (static_cast<unsigned long long>(w) * 64ull + uint32_t(i)));
                            if (float(uint32_t(u >> 40)) * 5.9604644775390625e-08f < a.prob) v = 0.f;
                        }
                        xs[(cur * I + i) * kSmallB + r] = v;
                        any |= v != 0.f;
                    }
                } else {
                    for (int i = lane; i < I; i += 32) xs[(cur * I + i) * kSmallB + r] = 0.f;
                }
                any = __any_sync(0xffffffffu, any);
                if (lane == 0) {
                    s_valid[cur][r] = any ? 1 : 0;
                    s_lab[cur][r] = r < B ? a.labels[w] : 0xffffffffu;
                }
            }
        }
        __syncthreads();
        if (live && tid == 0) {
            int n = 0;
            for (int r = 0; r < B; ++r) n += s_valid[cur][r];
            s_nused[cur] = float(n);
        }
        __syncthreads();
        const float n_used = live ? s_nused[cur] : 0.f;
        const float scale = n_used > 0.f ? a.lr / n_used : 0.f;       // empty batch: nothing moves (lib.rs:1003-1005)

        // ---- A: per column i of W1: previous step's update of the column, then the forward pass through it ------------------
        for (int i = gwarp; i < H1; i += nwarps) {
            if (scale_prev != 0.f) {
                const float dprev = lane < B_prev ? __ldcg(d1T + size_t(i) * kSmallB + lane) : 0.f;     // d1 of the previous step
                part[warp * kSmallB + lane] = dprev;
                __syncwarp();
                for (int k0 = 0; k0 < I; k0 += 32) {                                                    // (uniform trip count per warp)
                    const int k = k0 + lane;
                    if (k < I) {
                        float g = 0.f;
                        for (int b = 0; b < B_prev; ++b) g = fmaf(xs[(prv * I + k) * kSmallB + b], part[warp * kSmallB + b], g);
                        W1[size_t(k) * H1 + i] = __ldcg(W1 + size_t(k) * H1 + i) - scale_prev * g;      // lib.rs:1055
                    }
                }
                float gb = dprev;
#pragma unroll
                for (int o = 16; o; o >>= 1) gb += __shfl_xor_sync(0xffffffffu, gb, o);
                if (lane == 0) b1[i] = __ldcg(b1 + i) - scale_prev * gb;                                // lib.rs:1056
                __syncwarp();
            }
            if (live) {
                float acc = __ldcg(b1 + i);
#pragma unroll 4
                for (int k = 0; k < I; ++k) acc = fmaf(xs[(cur * I + k) * kSmallB + lane], __ldcg(W1 + size_t(k) * H1 + i), acc);
                h1T[size_t(i) * kSmallB + lane] = lane < B ? fmaxf(acc, 0.f) : 0.f;                     // lib.rs:1016-1017
            }
        }
        if (!live) break;
        grid_barrier(a.barrier, gen);

        // ---- B: h2 = tanh(h1 W2 + b2): CTA task = two columns, four warps split K = H1 per column --------------------------
        for (int t = blockIdx.x; t < (H2 + 1) / 2; t += gridDim.x) {
            const int j = 2 * t + (warp >> 2), q = warp & 3;
            float acc = 0.f;
            if (j < H2) {
                const int k0 = (H1 * q) / 4, k1 = (H1 * (q + 1)) / 4;
#pragma unroll 8
                for (int k = k0; k < k1; ++k) acc = fmaf(__ldcg(h1T + size_t(k) * kSmallB + lane), __ldcg(W2 + size_t(k) * H2 + j), acc);
            }
            part[warp * kSmallB + lane] = acc;
            __syncthreads();
            if (q == 0 && j < H2) {
                const float* p = part + (warp >> 2) * 4 * kSmallB + lane;
                const float v = tanhf(((p[0] + p[kSmallB]) + (p[2 * kSmallB] + p[3 * kSmallB])) + __ldcg(b2 + j));   // lib.rs:1018-1019
                h2T[size_t(j) * kSmallB + lane] = lane < B ? v : 0.f;
            }
            __syncthreads();
        }
        grid_barrier(a.barrier, gen);

        // ---- C: logits, CTA task = one class, eight warps split K = H2 -------------------------------------------------------
        for (int c = blockIdx.x; c < C; c += gridDim.x) {
            const int k0 = (H2 * warp) / 8, k1 = (H2 * (warp + 1)) / 8;
            float acc = 0.f;
#pragma unroll 8
            for (int k = k0; k < k1; ++k) acc = fmaf(__ldcg(h2T + size_t(k) * kSmallB + lane), __ldcg(W3 + size_t(k) * C + c), acc);
            part[warp * kSmallB + lane] = acc;
            __syncthreads();
            if (warp == 0) {
                float v = __ldcg(b3 + c);
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) v += part[w8 * kSmallB + lane];
                zT[size_t(c) * kSmallB + lane] = v;                                                     // lib.rs:1020-1021
            }
            __syncthreads();
        }
        grid_barrier(a.barrier, gen);

        // ---- C2: softmax, d3 = p - t, loss with the pre-update weights: one warp per row, lanes over classes -----------------
        for (int b = B + gwarp; b < kSmallB; b += nwarps)                                               // rows beyond the batch: no delta
            for (int c = lane; c < C; c += 32) zT[size_t(c) * kSmallB + b] = 0.f;
        for (int b = gwarp; b < B; b += nwarps) {
            const bool ok = s_valid[cur][b] != 0;
            const uint32_t label = s_lab[cur][b];
            float mx = -INFINITY;
            for (int c = lane; c < C; c += 32) mx = fmaxf(mx, __ldcg(zT + size_t(c) * kSmallB + b));
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int c = lane; c < C; c += 32) sum += expf(__ldcg(zT + size_t(c) * kSmallB + b) - mx);
#pragma unroll
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            float loss = 0.f;
            for (int c = lane; c < C; c += 32) {
                const float p = expf(__ldcg(zT + size_t(c) * kSmallB + b) - mx) / sum;                   // lib.rs:1023-1026
                const float t = uint32_t(c) == label ? 1.f : 0.f;                                       // label >= C: all-zero target
                zT[size_t(c) * kSmallB + b] = ok ? p - t : 0.f;                                         // lib.rs:1028
                if (ok && uint32_t(c) == label) loss = -logf(fmaxf(p, 1e-12f));                         // lib.rs:611-615
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
            if (lane == 0 && ok) {
                atomicAdd(a.stats, double(loss));
                atomicAdd(a.stats + 1, 1.0);
            }
        }
        grid_barrier(a.barrier, gen);

        // ---- D: per row j of W3 (warp task): d2[:, j], then the row's update ------------------------------------------------
        for (int i = tid; i < C * kSmallB; i += kSmallThreads) dts[(i >> 5) * 33 + (i & 31)] = __ldcg(zT + i);   // d3^T of the step
        __syncthreads();
        for (int j = gwarp; j < H2; j += nwarps) {
            const float h = __ldcg(h2T + size_t(j) * kSmallB + lane);
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc = fmaf(dts[c * 33 + lane], __ldcg(W3 + size_t(j) * C + c), acc);
            d2T[size_t(j) * kSmallB + lane] = acc * (1.f - h * h);                                      // lib.rs:1034
            part[warp * kSmallB + lane] = h;
            __syncwarp();
            if (scale != 0.f)
                for (int c = lane; c < C; c += 32) {                                                    // (no warp-collective inside)
                    float g = 0.f;
                    for (int b = 0; b < B; ++b) g = fmaf(part[warp * kSmallB + b], dts[c * 33 + b], g);
                    W3[size_t(j) * C + c] = __ldcg(W3 + size_t(j) * C + c) - scale * g;                 // lib.rs:1051
                }
            __syncwarp();
        }
        if (scale != 0.f)
            for (int c = gwarp * 32 + lane; c < C; c += nwarps * 32) {                                   // b3 (lib.rs:1052)
                float g = 0.f;
                for (int b = 0; b < B; ++b) g += dts[c * 33 + b];
                b3[c] = __ldcg(b3 + c) - scale * g;
            }
        grid_barrier(a.barrier, gen);

        // ---- E: per row i of W2 (CTA task, eight warps split K = H2): d1[:, i], then the row's update ------------------------
        for (int i = tid; i < H2 * kSmallB; i += kSmallThreads) dts[(i >> 5) * 33 + (i & 31)] = __ldcg(d2T + i);   // d2^T of the step
        __syncthreads();
        for (int i = blockIdx.x; i < H1; i += gridDim.x) {
            const int k0 = (H2 * warp) / 8, k1 = (H2 * (warp + 1)) / 8;
            float acc = 0.f;
#pragma unroll 8
            for (int k = k0; k < k1; ++k) acc = fmaf(dts[k * 33 + lane], __ldcg(W2 + size_t(i) * H2 + k), acc);
            part[warp * kSmallB + lane] = acc;
            __syncthreads();
            const float h = __ldcg(h1T + size_t(i) * kSmallB + lane);
            if (warp == 0) {
                float v = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) v += part[w8 * kSmallB + lane];
                d1T[size_t(i) * kSmallB + lane] = h > 0.f ? v : 0.f;                                    // lib.rs:1039-1040
                s_hrow[lane] = h;
            }
            __syncthreads();
            if (scale != 0.f)
                for (int j = tid; j < H2; j += kSmallThreads) {
                    float g = 0.f;
                    for (int b = 0; b < B; ++b) g = fmaf(s_hrow[b], dts[j * 33 + b], g);
                    W2[size_t(i) * H2 + j] = __ldcg(W2 + size_t(i) * H2 + j) - scale * g;               // lib.rs:1053
                }
            __syncthreads();
        }
        if (scale != 0.f)
            for (int j = gwarp * 32 + lane; j < H2; j += nwarps * 32) {                                  // b2 (lib.rs:1054)
                float g = 0.f;
                for (int b = 0; b < B; ++b) g += dts[j * 33 + b];
                b2[j] = __ldcg(b2 + j) - scale * g;
            }
        grid_barrier(a.barrier, gen);
        scale_prev = scale;
        B_prev = B;
    }
}

// Runs the steps of an epoch in the persistent kernel when they qualify.  *done = false leaves the caller on the ordinary path.
szb_status train_epoch_small(szb_net* net, const float* d_feats, const uint32_t* d_labels, const uint32_t* d_perm, const uint32_t* step_sizes,
                             uint32_t n_steps, float lr, float dropout, unsigned long long key, const uint8_t* d_keep, bool* done) {
    *done = false;
    szb_ctx* ctx = net->ctx;
    if (!ctx->small_steps || ctx->world != 1 || net->small_failed || n_steps == 0 || net->n_out > uint32_t(kSmallMaxC)) return SZB_OK;
    uint32_t max_b = 0;
    for (uint32_t i = 0; i < n_steps; ++i) max_b = std::max(max_b, step_sizes[i]);
    if (max_b == 0 || max_b > uint32_t(kSmallB)) return SZB_OK;
    const size_t dts_rows = std::max<size_t>(net->h2, net->n_out);
    const size_t smem = (size_t(2) * net->n_in * kSmallB + 8 * kSmallB + dts_rows * 33) * sizeof(float);
    if (smem > 160 * 1024) return SZB_OK;
    if (net->small_grid == 0) {
        int per_sm = 0, coop = 0;
        SZB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        if (!coop) { net->small_failed = true; return SZB_OK; }
        SZB_CUDA(cudaFuncSetAttribute(train_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(160 * 1024)));
        SZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, train_small_kernel, kSmallThreads, smem));
        if (per_sm < 1) { net->small_failed = true; return SZB_OK; }
        net->small_grid = ctx->sm_count;                   // one CTA per SM: the phases are latency-bound, more CTAs only lengthen the barriers
    }
    const size_t scratch_floats = (size_t(net->h1) * 2 + size_t(net->h2) * 2 + net->n_out) * kSmallB;
    SZB_TRY(net->small_scratch.reserve(scratch_floats * sizeof(float) + 64));
    if (!net->small_barrier.ptr) {
        SZB_TRY(net->small_barrier.reserve(2 * sizeof(unsigned int)));
        SZB_CUDA(cudaMemsetAsync(net->small_barrier.ptr, 0, 2 * sizeof(unsigned int), ctx->stream));
    }
    SZB_TRY(net->small_steps.reserve((size_t(n_steps) + 1) * sizeof(uint32_t)));
    void* hp = nullptr;
    SZB_TRY(ctx->h_stage.acquire((size_t(n_steps) + 1) * sizeof(uint32_t), &hp));
    uint32_t* off = static_cast<uint32_t*>(hp);
    off[0] = 0;
    for (uint32_t i = 0; i < n_steps; ++i) off[i + 1] = off[i] + step_sizes[i];
    SZB_CUDA(cudaMemcpyAsync(net->small_steps.ptr, hp, (size_t(n_steps) + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(ctx->h_stage.uploaded(ctx->stream));

    SmallArgs a{};
    a.P = net->params.as<float>();
    a.off_b1 = net->off_b1(); a.off_w2 = net->off_w2(); a.off_b2 = net->off_b2(); a.off_w3 = net->off_w3(); a.off_b3 = net->off_b3();
    a.n_in = int(net->n_in); a.h1 = int(net->h1); a.h2 = int(net->h2); a.C = int(net->n_out);
    a.feats = d_feats; a.labels = d_labels; a.perm = d_perm; a.keep = d_keep;
    a.step_off = net->small_steps.as<uint32_t>();
    a.n_steps = n_steps;
    a.prob = dropout; a.lr = lr; a.key = key;
    a.scratch = net->small_scratch.as<float>();
    a.barrier = net->small_barrier.as<unsigned int>();
    a.stats = net->stats.as<double>();
    void* args[] = { &a };
    const cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(train_small_kernel), dim3(net->small_grid), dim3(kSmallThreads),
                                                      args, smem, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        net->small_failed = true;                          // e.g. the device is shared and the grid cannot be co-resident: ordinary path
        return SZB_OK;
    }
    ctx->launches += 1;
    net->wt_dirty = true;                                  // the tensor-core path's transposed copies no longer match the weights
    *done = true;
    return SZB_OK;
}

}  // namespace szb
