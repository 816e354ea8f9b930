// SimpleNeuralNet (streamz-rs/src/lib.rs:745-1060) on the GPU: batched forward, cross-entropy backward, SGD, and the
// per-window aggregation of identify_speaker_list (lib.rs:1383-1411).
//
// This file holds the FP32 path for arbitrary layer sizes (the reference's constructor is generic, lib.rs:767, and its
// own test uses a 4-3-2-2 net, lib.rs:1834): a register-tiled SIMT GEMM with the layer epilogues fused
// (bias+ReLU, bias+tanh, bias, *tanh', *ReLU').  Results follow the reference's arithmetic in FP32 exactly up to
// summation order.  See DESIGN.md "MLP kernels".
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "gemm_tma.cuh"
#include "mlp.cuh"

namespace szb {

// ---- counter RNG shared with the oracle (oracle/streamz_oracle.py: dropout_keep_mask) -------------------------------
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ unsigned long long dropout_key(unsigned long long seed, unsigned long long stream) {
    return splitmix64(seed + stream * 0xD1B54A32D192ED03ull);
}
// true = keep.  u = 24 random bits * 2^-24 in [0,1); dropped when u < prob (rng.gen::<f32>() < prob, lib.rs:125)
__host__ __device__ __forceinline__ bool dropout_keep(unsigned long long key, unsigned long long row, uint32_t i, float prob) {
    const unsigned long long u = splitmix64(key ^ (row * 64ull + i));
    const float r = float(uint32_t(u >> 40)) * 5.9604644775390625e-08f;
    return !(r < prob);
}

// ---- SIMT GEMM with fused epilogues ----------------------------------------------------------------------------------
enum Epi : int { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_TANH = 2, EPI_MUL_DTANH = 3, EPI_MUL_DRELU = 4, EPI_ATOMIC = 5 };

constexpr int GB = 64;   // block tile M and N
constexpr int GK = 16;   // block tile K
constexpr int GT = 256;  // threads, each computes 4 x 4

// C[M,N] (+)= op(A)[M,K] * op(B)[K,N].  TA: A is stored [K][M] (lda = row stride); TB: B is stored [N][K].
// blockIdx.z splits K (only with EPI_ATOMIC).
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(GT) gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                 const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                                 const float* __restrict__ bias, const float* __restrict__ aux, int ldaux,
                                                 int k_chunk) {
    __shared__ float sA[GK][GB + 4];
    __shared__ float sB[GK][GB + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
    const int kb = blockIdx.z * k_chunk, ke = min(K, kb + k_chunk);
    float acc[4][4] = {};
    for (int k0 = kb; k0 < ke; k0 += GK) {
        // load A tile -> sA[k][m]
        for (int i = tid; i < GB * GK; i += GT) {
            int m, k;
            if (TA) { m = i % GB; k = i / GB; } else { k = i % GK; m = i / GK; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < ke) v = TA ? A[size_t(gk) * lda + gm] : A[size_t(gm) * lda + gk];
            sA[k][m] = v;
        }
        for (int i = tid; i < GB * GK; i += GT) {
            int n, k;
            if (TB) { k = i % GK; n = i / GK; } else { n = i % GB; k = i / GB; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < N && gk < ke) v = TB ? B[size_t(gn) * ldb + gk] : B[size_t(gk) * ldb + gn];
            sB[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (EPI == EPI_BIAS) v += bias[gn];
            if (EPI == EPI_BIAS_RELU) { v += bias[gn]; v = v > 0.f ? v : 0.f; }           // lib.rs:882
            if (EPI == EPI_BIAS_TANH) v = tanhf(v + bias[gn]);                             // lib.rs:883
            if (EPI == EPI_MUL_DTANH) { const float h = aux[size_t(gm) * ldaux + gn]; v *= (1.f - h * h); }  // lib.rs:1034
            if (EPI == EPI_MUL_DRELU) v = aux[size_t(gm) * ldaux + gn] > 0.f ? v : 0.f;   // lib.rs:1040
            if (EPI == EPI_ATOMIC) atomicAdd(&C[size_t(gm) * ldc + gn], v);
            else C[size_t(gm) * ldc + gn] = v;
        }
    }
}

template <bool TA, bool TB, int EPI>
static szb_status gemm(szb_ctx* ctx, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                       const float* bias, const float* aux, int ldaux) {
    if (M <= 0 || N <= 0) return SZB_OK;
    dim3 grid((N + GB - 1) / GB, (M + GB - 1) / GB, 1);
    int k_chunk = K;
    if (EPI == EPI_ATOMIC) {
        const int tiles = grid.x * grid.y;
        int split = std::max(1, std::min((ctx->sm_count * 2 + tiles - 1) / tiles, (K + 4 * GK - 1) / (4 * GK)));
        k_chunk = ((K + split - 1) / split + GK - 1) / GK * GK;
        grid.z = (K + k_chunk - 1) / k_chunk;
    }
    gemm_kernel<TA, TB, EPI><<<grid, GT, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux, k_chunk);
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    return SZB_OK;
}

// ---- row-wise softmax with the consumers fused ----------------------------------------------------------------------
// One warp per row.  mode bits: 1 = write probabilities, 2 = training (delta3 = p - t in place of z, loss, count),
// 4 = identify histogram (argmax_last / threshold), 8 = accumulate per-class probability sums.
__global__ void softmax_kernel(float* __restrict__ z, int rows, int C, int ldz, int mode, float* __restrict__ probs,
                               const uint32_t* __restrict__ labels, const float* __restrict__ target_vec,
                               const uint8_t* __restrict__ valid, float* __restrict__ tail, float threshold,
                               unsigned long long* __restrict__ hist, float* __restrict__ sums, float* __restrict__ zT, int ldzT,
                               const unsigned long long* __restrict__ clip_off = nullptr, uint32_t n_clips = 0,
                               unsigned long long row_base = 0, uint32_t* __restrict__ clip_hist = nullptr) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* zr = z + size_t(row) * ldz;
    const bool ok = valid ? valid[row] != 0 : true;
    const uint32_t label = labels ? labels[row] : 0xffffffffu;
    float best = -1.f;
    int best_c = -1;
    // per-element consumer shared by both paths
    auto consume = [&](int c, float p) {
        if (mode & 1) probs[size_t(row) * C + c] = p;
        if (mode & 2) {
            const float t = target_vec ? target_vec[c] : (uint32_t(c) == label ? 1.f : 0.f);
            const float d3 = ok ? p - t : 0.f;                                    // lib.rs:1028
            zr[c] = d3;
            if (zT) zT[size_t(c) * ldzT + row] = d3;
            if (ok && !target_vec && uint32_t(c) == label) atomicAdd(&tail[1], -logf(fmaxf(p, 1e-12f)));  // lib.rs:611-615
        }
        if (mode & 4) { if (p >= best) { best = p; best_c = c; } }                // last maximal element wins
        if ((mode & 8) && ok) atomicAdd(&sums[c], p);
    };
    if (C <= 128) {
        // the whole row lives in registers: one read of the logits, one expf per element
        float v[4], e[4];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            v[q] = c < C ? zr[c] : -INFINITY;
            mx = fmaxf(mx, v[q]);                                                 // lib.rs:887
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            e[q] = lane + 32 * q < C ? expf(v[q] - mx) : 0.f;                     // lib.rs:888
            sum += e[q];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c < C) consume(c, e[q] / sum);                                    // lib.rs:890
        }
    } else {
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, zr[c]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int c = lane; c < C; c += 32) sum += expf(zr[c] - mx);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        for (int c = lane; c < C; c += 32) consume(c, expf(zr[c] - mx) / sum);
    }
    if (mode & 4) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, o);
            if (ob > best || (ob == best && oc > best_c)) { best = ob; best_c = oc; }
        }
        if (lane == 0 && best_c >= 0 && best >= threshold) {                                 // lib.rs:1398-1400
            if (clip_hist) {     // batched form: the window's clip = last c with clip_off[c] <= global row (empty clips skipped)
                const unsigned long long g = row_base + uint32_t(row);
                uint32_t lo = 0, hi = n_clips;
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (clip_off[mid] <= g) lo = mid; else hi = mid;
                }
                atomicAdd(&clip_hist[size_t(lo) * C + best_c], 1u);
            } else {
                atomicAdd(&hist[best_c], 1ull);
            }
        }
    }
    if ((mode & 2) && lane == 0 && ok) atomicAdd(&tail[0], 1.f);
}


// Training softmax for C <= 128 (one CTA = 32 rows, one warp per row): softmax, delta3 = p - t (lib.rs:1028) written
// in place AND as a [C][32]-row tile staged in shared memory so the transposed copy leaves with 128-byte coalesced
// stores; loss (lib.rs:611-615) and surviving-window count are reduced inside the CTA: two atomics per 32 rows.
__global__ void __launch_bounds__(1024) softmax_train_kernel(float* __restrict__ z, int rows, int C, const uint32_t* __restrict__ labels,
                                                            const float* __restrict__ target_vec, const uint8_t* __restrict__ valid,
                                                            float* __restrict__ tail, float* __restrict__ zT, int ldzT) {
    __shared__ float tile[128][33];
    __shared__ float s_loss[32], s_cnt[32];
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
    const int row0 = blockIdx.x * 32, row = row0 + wrow;
    float loss = 0.f, cnt = 0.f;
    if (row < rows) {
        float* zr = z + size_t(row) * C;
        const bool ok = valid ? valid[row] != 0 : true;
        const uint32_t label = labels ? labels[row] : 0xffffffffu;
        float v[4], e[4];
        float mx = -INFINITY;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            v[q] = c < C ? zr[c] : -INFINITY;
            mx = fmaxf(mx, v[q]);                                                 // lib.rs:1023
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            e[q] = lane + 32 * q < C ? expf(v[q] - mx) : 0.f;                     // lib.rs:1024
            sum += e[q];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            if (c < C) {
                const float p = e[q] / sum;                                       // lib.rs:1026
                const float t = target_vec ? target_vec[c] : (uint32_t(c) == label ? 1.f : 0.f);
                const float d3 = ok ? p - t : 0.f;
                zr[c] = d3;
                tile[c][wrow] = d3;
                if (ok && !target_vec && uint32_t(c) == label) loss = -logf(fmaxf(p, 1e-12f));
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
        cnt = ok ? 1.f : 0.f;
    }
    if (lane == 0) { s_loss[wrow] = loss; s_cnt[wrow] = cnt; }
    __syncthreads();
    const int nrows = min(32, rows - row0);
    for (int c = wrow; c < C; c += 32)
        if (lane < nrows) zT[size_t(c) * ldzT + row0 + lane] = tile[c][lane];
    if (wrow == 0) {
        float l = s_loss[lane], n = s_cnt[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            l += __shfl_xor_sync(0xffffffffu, l, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (lane == 0) {
            if (n > 0.f) atomicAdd(&tail[0], n);
            if (l != 0.f) atomicAdd(&tail[1], l);
        }
    }
}

// Column sums of D[rows][cols] accumulated into g[cols] (bias gradients, lib.rs:1033,1038,1044).
__global__ void colsum_kernel(const float* __restrict__ D, int rows, int cols, int ld, float* __restrict__ g, int rows_per_block) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float acc = 0.f;
    for (int r = r0; r < r1; ++r) acc += D[size_t(r) * ld + c];
    atomicAdd(&g[c], acc);
}

// Gather rows perm[s..s+B) of feats, apply input dropout (lib.rs:119-129, no rescale), flag windows left all-zero
// (lib.rs:607-609).  One block = 32 batch rows, a warp per row (or several rows per warp); the transposed copy the
// weight-gradient GEMM reads is staged in shared memory ([n_in][33] floats, when it fits) and leaves as 128-byte row segments.
// The body is shared by prep_batch_kernel and by the extra CTAs of sgd_fused_kernel that prepare the NEXT batch.
struct PrepArgs {
    const float* feats; const uint32_t* labels_all; const uint32_t* perm;
    int B, n_in;
    const uint8_t* keep; int keep_by_row;   // keep indexed by window id (1) or by batch row (0)
    float prob; unsigned long long key;
    float* xb; float* xbT; uint32_t* lab; uint8_t* valid;
    float* h1T; int h1; float* h2T; int h2;
    int use_tile;
    int blocks_per_step;    // > 0: the launch prepares SEVERAL consecutive steps of B rows each (block b serves step b / blocks_per_step);
};                          //      step j's rows follow step j - 1's in xb / lab / valid, its transposed block in xbT

__device__ __forceinline__ void prep_batch_block(const PrepArgs& a, const int block, const uint32_t* __restrict__ perm,
                                                 const unsigned long long key, float* prep_tile /* [n_in][33] when use_tile */) {
    const int B = a.B, n_in = a.n_in;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int row0 = block * 32;
    for (int r = warp; r < 32; r += nw) {
        const int row = row0 + r;
        if (row >= B) break;
        const uint32_t w = perm ? perm[row] : uint32_t(row);
        bool any = false;
        for (int i = lane; i < n_in; i += 32) {
            float v = a.feats[size_t(w) * n_in + i];
            if (a.keep) {
                if (!a.keep[size_t(a.keep_by_row ? w : uint32_t(row)) * n_in + i]) v = 0.f;
            } else if (a.prob > 0.f) {
                if (!dropout_keep(key, w, uint32_t(i), a.prob)) v = 0.f;
            }
            a.xb[size_t(row) * n_in + i] = v;
            if (a.xbT) {
                if (a.use_tile) prep_tile[i * 33 + r] = v;
                else a.xbT[size_t(i) * B + row] = v;
            }
            any |= (v != 0.f);
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) {
            a.valid[row] = any ? 1 : 0;
            if (a.lab) a.lab[row] = a.labels_all ? a.labels_all[w] : 0xffffffffu;
        }
    }
    if (!a.xbT) return;
    const int nrows = min(32, B - row0);
    if (a.use_tile) {
        __syncthreads();
        for (int i = warp; i < n_in; i += nw)
            if (lane < nrows) a.xbT[size_t(i) * B + row0 + lane] = prep_tile[i * 33 + lane];
    }
    // row of ones under each transposed activation: the weight-gradient GEMM then yields the bias gradient
    if (warp == 0 && lane < nrows) {
        a.xbT[size_t(n_in) * B + row0 + lane] = 1.f;
        a.h1T[size_t(a.h1) * B + row0 + lane] = 1.f;
        a.h2T[size_t(a.h2) * B + row0 + lane] = 1.f;
    }
}

__global__ void prep_batch_kernel(const __grid_constant__ PrepArgs a, StepParams* __restrict__ sp) {
    extern __shared__ float prep_tile[];   // [n_in][33] when use_tile
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const uint32_t* perm = a.perm;
    unsigned long long key = a.key;
    if (sp) {                              // captured step: position in the shuffled order and dropout key come from device memory
        perm += sp->cursor;
        key = sp->key;
        // multi-GPU: the flag value of this step's gradient exchange (read by the update kernel, the last kernel of the step)
        if (blockIdx.x == 0 && threadIdx.x == 0) sp->p2p_step += 1u;
    }
    if (a.blocks_per_step > 0) {
        // several steps of the epoch in one launch (szb_net_train_epoch_steps_dev: large batches): the batch kernel leaves the
        // per-step critical path.  Nothing here depends on the weights, only on the permutation and the dropout stream.
        const int j = int(blockIdx.x) / a.blocks_per_step;
        PrepArgs b = a;
        b.xb += size_t(j) * a.B * a.n_in;
        if (b.xbT) b.xbT += size_t(j) * (a.n_in + 1) * a.B;
        if (b.lab) b.lab += size_t(j) * a.B;
        b.valid += size_t(j) * a.B;
        prep_batch_block(b, int(blockIdx.x) % a.blocks_per_step, perm + size_t(j) * a.B, key, prep_tile);
        return;
    }
    prep_batch_block(a, blockIdx.x, perm, key, prep_tile);
}

// theta -= (lr / n_used) * g   (lib.rs:1047-1059); n_used comes from the (all-reduced) gradient tail.
__global__ void sgd_kernel(float* __restrict__ params, const float* __restrict__ grads, size_t n, float lr,
                           double* __restrict__ stats) {
    const float n_used = grads[n];
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats) {
        stats[0] += double(grads[n + 1]);  // loss
        stats[1] += double(n_used);
    }
    if (n_used <= 0.f) return;             // empty batch: no-op (lib.rs:1003-1005)
    const float scale = lr / n_used;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        params[i] -= grads[i] * scale;
}

// ---- multi-GPU gradient exchange over NVLink peer memory, fused into the update kernels ---------------------------------
// Every rank's exchange region (comm.cu: [flags | inbox | red]) is mapped into all ranks with CUDA IPC.  A step's gradient
// vector V = G[0, nv) (nv = parameters + the two [n_used, loss, ..] tail blocks) is cut into `world` contiguous slices of
// slice_g 16-byte groups; rank r OWNS slice r.  The exchange is a two-shot all-reduce made of posted peer STORES only --
// no rank ever waits on a remote load -- in the same launch as the SGD update:
//   phase 0  scatter : every rank stores slice d of its private gradient into inbox[me] of rank d   (1/world of V per peer)
//            ...  the last CTA to finish publishes flag1[me] = step on every peer (release, system scope)
//   phase 1  reduce  : wait flag1 from all ranks (acquire); the owner sums its `world` inbox rows IN RANK ORDER (its own
//            row straight from G) and stores the reduced slice into `red` of EVERY rank;  last CTA publishes flag2[me]
//   phase 2  update  : wait flag2 from all ranks; red (local memory) now holds the whole reduced vector: the ordinary
//            update runs on it.  Each element is summed by exactly one rank in a fixed order and broadcast, so the replicas
//            stay bit-identical (tested) and equal the old one-shot pull bit for bit.
// Round 1's one-shot pull read all `world` full vectors per rank with dependent 4-byte peer loads (0.39 per-GPU efficiency at
// 8 GPUs, 259 us per step) and needed a 753 KB device-to-device copy per step; here a rank moves 2 (world-1)/world of V in
// posted stores and the copy is gone (phase 0 reads G directly).
// Buffer reuse needs no double buffering: a rank enters step t+1 only after it has seen flag2(t) from every rank, and a rank
// publishes flag2(t) only after it has finished reading its inbox(t); `red` of rank d is rewritten in phase 1 of step t+1,
// which the writers enter only after flag1(t+1) from d, i.e. after d's step-t kernel (the reader of red) has completed.
struct P2pArgs {
    float* inbox[szb_ctx::kMaxPeers];        // two-shot: inbox of every rank: [world][slice_g * 4] floats, row = source rank
    float* red[szb_ctx::kMaxPeers];          // two-shot: reduced-vector buffer of every rank
    float* area[szb_ctx::kMaxPeers];         // one-shot: [2 (step parity)][world (source rank)][cap] full gradient vectors
    size_t cap;                              // floats per vector slot of `area`
    int one_shot;                            // 1: every rank stores its WHOLE vector into every peer (one flag round); 2: the same with
                                             // the flag inside every 8-byte packet (no flag round at all: ll_send / ll_recv below)
    uint32_t* flags[szb_ctx::kMaxPeers];     // flag block of every rank: [0, 16) flag1 per source rank, [16, 32) flag2
    unsigned int* counters;                  // private: [0] CTAs done with phase 0, [1] with phase 1 (last-CTA detection)
    unsigned long long* trace;               // optional (szb_comm_peer_trace): CTA 0 accumulates %globaltimer deltas per phase
    int rank, world;
    uint32_t step;
    uint32_t n4, slice_g;                    // 16-byte groups in V, groups per slice
    int nblk;                                // CTAs taking part in the exchange: blocks [0, nblk) of the grid (0 = the whole grid)
    uint32_t sent4;                          // one-shot: 16-byte groups [0, sent4) were already stored into the peers by the grouped
};                                           // weight-gradient launch (gemm_tma.cuh: PeerPush); the update kernel sends the rest

__device__ __forceinline__ void p2p_wait_flags(const uint32_t* f, int world, uint32_t step) {
    if (int(threadIdx.x) < world) {
        const uint32_t* p = f + threadIdx.x;
        uint32_t seen = 0;
        unsigned long long t0 = 0;
        for (uint32_t spin = 0;; ++spin) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p) : "memory");
            if (int32_t(seen - step) >= 0) break;
            __nanosleep(32);
            if ((spin & 0xFFFFu) == 0xFFFFu) {             // a peer that is minutes late has died: never hang the GPU
                unsigned long long t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 180ull * 1000000000ull) __trap();
            }
        }
    }
    __syncthreads();
}

// Called by every thread of every CTA when this CTA's stores of a phase are issued: the LAST CTA of the grid to arrive
// publishes `step` into flag word (base + rank) of every peer.  (stores -> bar -> fence.sys -> ticket) is the release
// pattern, (ticket -> fence.sys -> flag store.release) chains it to the peers' acquire loads.
__device__ __forceinline__ void p2p_publish(const P2pArgs& a, unsigned int* counter, int flag_base, const uint32_t step) {
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int ticket = atomicAdd(counter, 1u);
        s_last = ticket == (a.nblk ? unsigned(a.nblk) : gridDim.x) - 1;
        if (s_last) *counter = 0u;                         // every CTA has arrived; the next step finds it zero
    }
    __syncthreads();
    if (s_last && int(threadIdx.x) < a.world)      // (the release store orders everything the barrier above made visible to this thread)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flags[threadIdx.x] + flag_base + a.rank), "r"(step) : "memory");
}

// Phases 0 and 1.  On return (all threads of the CTA) the local `red` buffer holds the rank-ordered sum of every rank's
// V; read it with L1-bypassing loads (peers wrote it).
__device__ __forceinline__ unsigned long long p2p_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// trace slot i accumulates the nanoseconds CTA 0 spent up to the end of phase part i (0 scatter, 1 publish, 2 wait flag1,
// 3 reduce + broadcast, 4 publish, 5 wait flag2); slot 7 counts the steps
#define P2P_TRACE(i) do { if (tr) { const unsigned long long tn = p2p_now(); atomicAdd(a.trace + (i), tn - t_prev); t_prev = tn; } } while (0)

__device__ __forceinline__ const float* p2p_exchange(const P2pArgs& a, const float* __restrict__ G, const uint32_t step) {
    const int me = a.rank, W = a.world;
    const bool tr = a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long t_prev = tr ? p2p_now() : 0ull;
    const size_t nthreads = size_t(a.nblk ? unsigned(a.nblk) : gridDim.x) * blockDim.x, t0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const float4* G4 = reinterpret_cast<const float4*>(G);
    if (a.one_shot) {
        // ONE-SHOT: the whole vector goes to every peer (W - 1 posted stores per 16-byte group), one flag round, and every rank
        // adds the W vectors itself while it updates (p2p_grad_at).  Moves (W - 1) x the vector per rank instead of 2 (W - 1) / W,
        // but saves a whole publish / wait round (~12 us): the faster exchange while the vector is small against the link
        // (753 KB x 7 peers = 7 us of NVLink time at W = 8).  Slots alternate by step parity: a rank overwrites slot p two steps
        // later, after every peer has published the step in between, i.e. has finished the update that read slot p.
        const size_t slot4 = (size_t(step & 1u) * W + me) * (a.cap / 4);
        for (size_t i = a.sent4 + t0; i < a.n4; i += nthreads) {
            const float4 v = G4[i];
#pragma unroll
            for (int r = 0; r < szb_ctx::kMaxPeers; ++r)
                if (r < W && r != me) reinterpret_cast<float4*>(a.area[r])[slot4 + i] = v;
        }
        P2P_TRACE(0);
        p2p_publish(a, a.counters + 0, 0, step);
        P2P_TRACE(1);
        p2p_wait_flags(a.flags[me], W, step);
        P2P_TRACE(2);
        if (tr) atomicAdd(a.trace + 7, 1ull);
        return nullptr;
    }
    // ---- phase 0: scatter my slices to their owners (my own slice stays in G)
    for (size_t i = t0; i < a.n4; i += nthreads) {
        const uint32_t d = uint32_t(i / a.slice_g);
        if (int(d) == me) continue;
        const float4 v = G4[i];
        reinterpret_cast<float4*>(a.inbox[d])[size_t(me) * a.slice_g + (i - size_t(d) * a.slice_g)] = v;
    }
    P2P_TRACE(0);
    p2p_publish(a, a.counters + 0, 0, step);
    P2P_TRACE(1);
    // ---- phase 1: reduce my slice in rank order, broadcast it
    p2p_wait_flags(a.flags[me], W, step);
    P2P_TRACE(2);
    const size_t g0 = size_t(me) * a.slice_g;
    const size_t mine = g0 < a.n4 ? min(size_t(a.slice_g), size_t(a.n4) - g0) : 0;
    const float4* in4 = reinterpret_cast<const float4*>(a.inbox[me]);
    for (size_t o = t0; o < mine; o += nthreads) {
        float4 v[szb_ctx::kMaxPeers];                       // all rows requested before the first add
#pragma unroll
        for (int r = 0; r < szb_ctx::kMaxPeers; ++r)
            if (r < W) v[r] = r == me ? G4[g0 + o] : __ldcg(in4 + size_t(r) * a.slice_g + o);
        float4 s = v[0];
#pragma unroll
        for (int r = 1; r < szb_ctx::kMaxPeers; ++r)
            if (r < W) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
#pragma unroll
        for (int r = 0; r < szb_ctx::kMaxPeers; ++r)
            if (r < W) reinterpret_cast<float4*>(a.red[r])[g0 + o] = s;
    }
    P2P_TRACE(3);
    p2p_publish(a, a.counters + 1, 16, step);
    P2P_TRACE(4);
    // ---- phase 2 entry: everybody's reduced slice has landed here
    p2p_wait_flags(a.flags[me] + 16, W, step);
    P2P_TRACE(5);
    if (tr) atomicAdd(a.trace + 7, 1ull);
    return a.red[me];
}

// Element idx of the step's reduced gradient after p2p_exchange: the two-shot exchange left it in R; after the one-shot
// exchange it is the rank-ordered sum of this rank's private value and the peers' copies in the local area.
__device__ __forceinline__ float p2p_grad_at(const P2pArgs& a, const float* __restrict__ R, const float* __restrict__ G, size_t idx, const uint32_t step) {
    if (!a.one_shot) return __ldcg(R + idx);
    const float* base = a.area[a.rank] + size_t(step & 1u) * a.world * a.cap + idx;
    float v[szb_ctx::kMaxPeers];
#pragma unroll
    for (int r = 0; r < szb_ctx::kMaxPeers; ++r)
        if (r < a.world) v[r] = r == a.rank ? G[idx] : __ldcg(base + size_t(r) * a.cap);
    float s = v[0];
#pragma unroll
    for (int r = 1; r < szb_ctx::kMaxPeers; ++r)
        if (r < a.world) s += v[r];
    return s;
}

// ---- one-shot exchange without a flag round ("LL": the protocol NCCL uses for small messages) -----------------------------
// A flag round costs ~12 us (block barrier, system-scope fence = the acknowledgement round trip of the posted NVLink stores,
// ticket, release store, propagation, the peers' acquire polls) against ~1.5 us of data movement (szb_comm_peer_trace), and
// sending the data earlier does not shorten it (PeerPush, measured).  So the data carries its own flag: element idx of the
// step's vector travels as ONE 8-byte packet {value, step}; an aligned 8-byte store is performed as a whole, so a receiver
// that reads the packet with one 8-byte load and finds the step number has the value that belongs to it.  No fence, no
// ticket, no flag, no CTA waits for another CTA of its own grid: every thread sends the elements it will update, then polls
// the peers' packets of exactly those elements and adds them in rank order (replicas stay bit-identical, and equal to the
// other two protocols bit for bit).  Costs twice the bytes; slots alternate by step parity exactly as in the one-shot exchange
// (a slot is rewritten two steps later, after every peer has sent the step in between, i.e. finished the update that read it).
// A slot last written by one of the other protocols holds plain floats where this one expects flags: a false match would need a
// gradient value whose bits equal the step counter (a denormal below 1e-38); the region starts zeroed and steps count from 1.
__device__ __forceinline__ unsigned long long* ll_slot(const P2pArgs& a, int dst, int src, const uint32_t step) {
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(a.area[dst]) +
                                                 (size_t(step & 1u) * size_t(a.world) + size_t(src)) * a.cap * sizeof(float));
}
__device__ __forceinline__ void ll_send(const P2pArgs& a, size_t idx, float v, const uint32_t step) {
    const uint32_t bits = __float_as_uint(v);
#pragma unroll
    for (int r = 0; r < szb_ctx::kMaxPeers; ++r)
        if (r < a.world && r != a.rank)
            asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(ll_slot(a, r, a.rank, step) + idx), "r"(bits), "r"(step) : "memory");
}
__device__ __forceinline__ float ll_recv(const P2pArgs& a, int src, size_t idx, const uint32_t step) {
    const unsigned long long* p = ll_slot(a, a.rank, src, step) + idx;
    uint32_t bits = 0, flag = 0;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(flag) : "l"(p) : "memory");
        if (flag == step) break;
        if ((spin & 0xFFFFu) == 0xFFFFu) {                 // a peer that is minutes late has died: never hang the GPU
            const unsigned long long t = p2p_now();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 180ull * 1000000000ull) __trap();
        }
    }
    return __uint_as_float(bits);
}
// rank-ordered sum of element idx over all ranks: this rank's value from its private gradient, the others from their packets
__device__ __forceinline__ float ll_grad_at(const P2pArgs& a, const float* __restrict__ G, size_t idx, const uint32_t step) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < szb_ctx::kMaxPeers; ++r)
        if (r < a.world) {
            const float v = r == a.rank ? G[idx] : ll_recv(a, r, idx, step);
            s = r == 0 ? v : s + v;
        }
    return s;
}

// Tensor-core path: the SGD update, the refresh of the transposed weight copies the GEMMs read, and the zeroing of the
// gradient vector for the next step in ONE pass (three launches -- sgd, transpose, 753 KB memset -- become one).
// One block = one 32 x 32 tile of a weight matrix W[k][n]: rows of P and G are read and written as 128-byte segments
// (lanes along n), the updated tile crosses a padded shared-memory tile and leaves for WT[n][k] as 128-byte segments along
// k.  The blocks past the last tile update the biases.  (The first version walked WT linearly with two integer divisions
// per element and lane-strided reads of P and G: 8 us per step against ~3 us now.)
// Multi-GPU (a.world > 1): the kernel starts with the two-shot gradient exchange above; every CTA of the grid must be
// resident at once (the launch keeps the grid below the device's capacity for this kernel).
__global__ void __launch_bounds__(256) sgd_fused_kernel(float* __restrict__ P, float* __restrict__ G, float* __restrict__ WT, int n_in,
                                                        int h1, int h2, int n_out, size_t off_b1, size_t off_w2, size_t off_b2,
                                                        size_t off_w3, size_t off_b3, size_t off_wt2, size_t off_wt3, size_t np, int parity,
                                                        float lr, double* __restrict__ stats, const __grid_constant__ P2pArgs a,
                                                        StepParams* __restrict__ sp, int advance) {
    __shared__ float tile[32][33];
    tc::pdl_launch_dependents();
    tc::pdl_wait();
    const int n_upd = int(gridDim.x);
    if (sp) {                              // captured step (CUDA graph): learning rate from device memory; move on to the next batch
        lr = sp->lr;                       // (the next step's batch kernel reads the position only after this grid has completed)
        if (blockIdx.x == 0 && threadIdx.x == 0) sp->cursor += uint32_t(advance);
    }
    const bool peers = a.world > 1;
    // the step's flag value: from the launch, or -- in a captured step, whose parameters are frozen -- from device memory, where the
    // step's batch kernel has just counted it up
    const uint32_t step = sp ? sp->p2p_step : a.step;
    const int t1 = ((n_in + 31) / 32) * ((h1 + 31) / 32), t2 = ((h1 + 31) / 32) * ((h2 + 31) / 32),
              t3 = ((h2 + 31) / 32) * ((n_out + 31) / 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t nb = size_t(h1) + h2 + n_out;
    const int nblk_tiles = t1 + t2 + t3;
    // biases: spread over the blocks starting behind the last tile's block
    const size_t bt = size_t((int(blockIdx.x) + n_upd - (nblk_tiles % n_upd)) % n_upd) * blockDim.x + threadIdx.x;
    auto tile_of = [&](int t, int& K, int& N, size_t& off_w, size_t& off_wt, int& k0, int& n0) {
        int tt;
        if (t < t1) { K = n_in; N = h1; tt = t; off_w = 0; off_wt = 0; }
        else if (t < t1 + t2) { K = h1; N = h2; tt = t - t1; off_w = off_w2; off_wt = off_wt2; }
        else { K = h2; N = n_out; tt = t - t1 - t2; off_w = off_w3; off_wt = off_wt3; }
        const int ntn = (N + 31) / 32;
        k0 = (tt / ntn) * 32; n0 = (tt % ntn) * 32;
    };
    auto bias_idx = [&](size_t i) -> size_t {
        return i < size_t(h1) ? off_b1 + i : (i < size_t(h1) + h2 ? off_b2 + (i - h1) : off_b3 + (i - h1 - h2));
    };
    if (peers && a.one_shot == 2) {
        // ---- exchange inside the packets (ll_send / ll_recv): pass 1 sends every element this thread will update, pass 2
        // polls the peers' packets of the same elements.  The thread that sends an element is the one that zeroes it.
        __shared__ float s_tail[2];
        if (blockIdx.x == 0 && threadIdx.x < 8) ll_send(a, np + threadIdx.x, G[np + threadIdx.x], step);    // [n_used, loss] blocks first
        for (int t = blockIdx.x; t < nblk_tiles; t += n_upd) {
            int K, N, k0, n0; size_t off_w, off_wt;
            tile_of(t, K, N, off_w, off_wt, k0, n0);
#pragma unroll
            for (int r = warp; r < 32; r += 8) {
                const int k = k0 + r, n = n0 + lane;
                if (k < K && n < N) { const size_t idx = off_w + size_t(k) * N + n; ll_send(a, idx, G[idx], step); }
            }
        }
        for (size_t i = bt; i < nb; i += size_t(n_upd) * blockDim.x) { const size_t idx = bias_idx(i); ll_send(a, idx, G[idx], step); }
        if (threadIdx.x < 2) s_tail[threadIdx.x] = ll_grad_at(a, G, np + 4 * parity + threadIdx.x, step);
        __syncthreads();
        const float n_used_ll = s_tail[0];
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (stats) { stats[0] += double(s_tail[1]); stats[1] += double(n_used_ll); }
            G[np + 4 * (1 - parity)] = 0.f;            // the next step's tail block (nobody reads it in this launch)
            G[np + 4 * (1 - parity) + 1] = 0.f;
        }
        const float scale_ll = n_used_ll > 0.f ? lr / n_used_ll : 0.f;
        for (int t = blockIdx.x; t < nblk_tiles; t += n_upd) {
            int K, N, k0, n0; size_t off_w, off_wt;
            tile_of(t, K, N, off_w, off_wt, k0, n0);
            __syncthreads();
#pragma unroll
            for (int r = warp; r < 32; r += 8) {
                const int k = k0 + r, n = n0 + lane;
                if (k < K && n < N) {
                    const size_t idx = off_w + size_t(k) * N + n;
                    const float p = P[idx] - ll_grad_at(a, G, idx, step) * scale_ll;
                    P[idx] = p;
                    G[idx] = 0.f;
                    tile[r][lane] = p;
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = warp; r < 32; r += 8) {
                const int n = n0 + r, k = k0 + lane;
                if (n < N && k < K) WT[off_wt + size_t(n) * K + k] = tile[lane][r];
            }
        }
        for (size_t i = bt; i < nb; i += size_t(n_upd) * blockDim.x) {
            const size_t idx = bias_idx(i);
            P[idx] -= ll_grad_at(a, G, idx, step) * scale_ll;
            G[idx] = 0.f;
        }
        return;
    }
    const float* R = peers ? p2p_exchange(a, G, step) : G;     // the step's (reduced) gradient vector
    auto grad_at = [&](size_t idx) -> float { return peers ? p2p_grad_at(a, R, G, idx, step) : R[idx]; };
    const float n_used = grad_at(np + 4 * parity);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (stats) {
            stats[0] += double(grad_at(np + 4 * parity + 1));  // loss
            stats[1] += double(n_used);
        }
        // the next step's tail block: nobody reads it during this launch (in a multi-GPU step every CTA of this rank has
        // finished reading G in phase 0 before any CTA gets here)
        G[np + 4 * (1 - parity)] = 0.f;
        G[np + 4 * (1 - parity) + 1] = 0.f;
    }
    const float scale = n_used > 0.f ? lr / n_used : 0.f;   // empty batch: gradients are zero, nothing moves (lib.rs:1003-1005)
    for (int t = blockIdx.x; t < nblk_tiles; t += n_upd) {
        int K, N, k0, n0;
        size_t off_w, off_wt;
        tile_of(t, K, N, off_w, off_wt, k0, n0);
        __syncthreads();                                  // the previous tile of this block has been read out
#pragma unroll
        for (int r = warp; r < 32; r += 8) {
            const int k = k0 + r, n = n0 + lane;
            if (k < K && n < N) {
                const size_t idx = off_w + size_t(k) * N + n;
                const float p = P[idx] - grad_at(idx) * scale;
                P[idx] = p;
                G[idx] = 0.f;
                tile[r][lane] = p;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = warp; r < 32; r += 8) {
            const int n = n0 + r, k = k0 + lane;
            if (n < N && k < K) WT[off_wt + size_t(n) * K + k] = tile[lane][r];
        }
    }
    for (size_t i = bt; i < nb; i += size_t(n_upd) * blockDim.x) {
        const size_t idx = bias_idx(i);
        P[idx] -= grad_at(idx) * scale;
        G[idx] = 0.f;
    }
}

// FP32 (CUDA-core) path of a multi-GPU step: the same exchange, then the plain update theta -= (lr / sum n_used) * sum g.
__global__ void __launch_bounds__(256) sgd_p2p_kernel(float* __restrict__ params, const float* __restrict__ G,
                                                      const __grid_constant__ P2pArgs a, size_t n, float lr, double* __restrict__ stats) {
    const float* R = p2p_exchange(a, G, a.step);
    const float n_used = p2p_grad_at(a, R, G, n, a.step), loss = p2p_grad_at(a, R, G, n + 1, a.step);
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats) {
        stats[0] += double(loss);
        stats[1] += double(n_used);
    }
    if (n_used <= 0.f) return;             // empty global batch: no-op (lib.rs:1003-1005)
    const float scale = lr / n_used;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        params[i] -= p2p_grad_at(a, R, G, i, a.step) * scale;
}

__global__ void init_uniform_kernel(float* __restrict__ w, size_t n, unsigned long long key, unsigned long long base) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const unsigned long long u = splitmix64(key ^ (base + i));
        w[i] = float(uint32_t(u >> 40)) * 5.9604644775390625e-08f - 0.5f;  // U[-0.5, 0.5) (lib.rs:770)
    }
}


// ---- embeddings (embed lib.rs:895-900, forward_embedding lib.rs:1073-1079) --------------------------
// Exact per-column order statistic of E[n][ld] by 4-pass radix select on order-preserving keys (one CTA per column):
// out[c] = value of rank r0 (and r1 when r1 != r0, averaged) -- the median of lib.rs:1438-1447 / 1487-1496.
__device__ __forceinline__ uint32_t f32_key(float f) {
    const uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_f32(uint32_t k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
__global__ void column_median_kernel(const float* __restrict__ E, uint32_t n, int ld, float* __restrict__ out) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_rank;
    const int c = blockIdx.x;
    const uint32_t ranks[2] = { (n % 2 == 0) ? n / 2 - 1 : n / 2, n / 2 };
    float vals[2];
    for (int which = 0; which < 2; ++which) {
        if (which == 1 && ranks[1] == ranks[0]) { vals[1] = vals[0]; break; }
        if (threadIdx.x == 0) { s_prefix = 0; s_rank = ranks[which]; }
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t k = f32_key(E[size_t(i) * ld + c]);
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t r = s_rank, b = 0;
                for (; b < 256; ++b) {
                    if (r < hist[b]) break;
                    r -= hist[b];
                }
                s_rank = r;
                s_prefix = prefix | (b << shift);
            }
            __syncthreads();
        }
        vals[which] = key_f32(s_prefix);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = (ranks[0] == ranks[1]) ? vals[0] : (vals[0] + vals[1]) / 2.0f;   // lib.rs:1441-1445
}

// out[c] = sum_r E[r][c] (mean embedding, lib.rs:1455-1462); deterministic: one CTA per 32 columns, fixed order
__global__ void column_sum_kernel(const float* __restrict__ E, uint32_t n, int cols, int ld, float* __restrict__ out) {
    __shared__ float part[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    float acc = 0.f;
    if (c < cols)
        for (uint32_t r = ry; r < n; r += 8) acc += E[size_t(r) * ld + c];
    part[ry][threadIdx.x & 31] = acc;
    __syncthreads();
    if (ry == 0 && c < cols) {
        float t = 0.f;
        for (int q = 0; q < 8; ++q) t += part[q][threadIdx.x & 31];
        out[c] = t;
    }
}

szb_status net_reserve_rows(szb_net* net, uint64_t rows) {
    if (rows <= net->cap_rows) return SZB_OK;
    SZB_TRY(net->xb.reserve(rows * net->n_in * 4));
    SZB_TRY(net->lab.reserve(rows * 4));
    SZB_TRY(net->valid.reserve(rows));
    SZB_TRY(net->a_h1.reserve(rows * net->h1 * 4));
    SZB_TRY(net->a_h2.reserve(rows * net->h2 * 4));
    SZB_TRY(net->a_z.reserve(rows * net->n_out * 4));
    SZB_TRY(net->d_2.reserve(rows * net->h2 * 4));
    SZB_TRY(net->d_1.reserve(rows * net->h1 * 4));
    SZB_TRY(net->xbT.reserve(rows * (net->n_in + 1) * 4));   // + the row of ones (bias gradients)
    SZB_TRY(net->h1T.reserve(rows * (net->h1 + 1) * 4));
    SZB_TRY(net->h2T.reserve(rows * (net->h2 + 1) * 4));
    SZB_TRY(net->zT.reserve(rows * net->n_out * 4));
    SZB_TRY(net->d2T.reserve(rows * net->h2 * 4));
    SZB_TRY(net->d1T.reserve(rows * net->h1 * 4));
    net->cap_rows = rows;
    return SZB_OK;
}

// wt1 = w1^T, wt2 = w2^T, wt3 = w3^T in one launch (the weights total < 2 MB and live in L2)
__global__ void transpose_weights_kernel(const float* __restrict__ P, float* __restrict__ WT, int n_in, int h1, int h2, int n_out,
                                         size_t off_w2, size_t off_w3, size_t off_wt2, size_t off_wt3) {
    const size_t n1 = size_t(n_in) * h1, n2 = size_t(h1) * h2, n3 = size_t(h2) * n_out;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += size_t(gridDim.x) * blockDim.x) {
        if (i < n1) {                       // i indexes wt1[n][k], n < h1, k < n_in
            const int n = int(i / n_in), k = int(i % n_in);
            WT[i] = P[size_t(k) * h1 + n];
        } else if (i < n1 + n2) {
            const size_t j = i - n1;
            const int n = int(j / h1), k = int(j % h1);
            WT[off_wt2 + j] = P[off_w2 + size_t(k) * h2 + n];
        } else {
            const size_t j = i - n1 - n2;
            const int n = int(j / h2), k = int(j % h2);
            WT[off_wt3 + j] = P[off_w3 + size_t(k) * n_out + n];
        }
    }
}

static szb_status refresh_wt(szb_net* net) {
    if (!net->wt_dirty) return SZB_OK;
    szb_ctx* ctx = net->ctx;
    SZB_TRY(net->wt.reserve(net->n_wt() * 4));
    const size_t n = net->n_wt();
    transpose_weights_kernel<<<int(std::min<size_t>((n + 255) / 256, size_t(ctx->sm_count) * 4)), 256, 0, ctx->stream>>>(
        net->params.as<float>(), net->wt.as<float>(), int(net->n_in), int(net->h1), int(net->h2), int(net->n_out), net->off_w2(),
        net->off_w3(), net->off_wt2(), net->off_wt3());
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    net->wt_dirty = false;
    return SZB_OK;
}

template <int EPI>
static szb_status gemm_tc(szb_net* net, tc::GemmArgs g, int split_k = 1) {
    // 128 x 128 tiles unless that grid would leave most of the 148 SMs idle (the per-step GEMMs are small): then
    // 128 x 64 tiles double the CTA count
    const int tiles128 = ((g.M + 127) / 128) * ((g.N + 127) / 128) * std::max(1, split_k);
    const bool narrow = tiles128 * 10 < net->ctx->sm_count * 7 && g.N > 64;
    bool done = false;
    if (net->precision == 2) {
        if (narrow) SZB_TRY((tc::launch_gemm_tma<64, 1, EPI>(net->ctx, g, split_k, &done)));
        else SZB_TRY((tc::launch_gemm_tma<128, 1, EPI>(net->ctx, g, split_k, &done)));
    } else {
        if (narrow) SZB_TRY((tc::launch_gemm_tma<64, 3, EPI>(net->ctx, g, split_k, &done)));
        else SZB_TRY((tc::launch_gemm_tma<128, 3, EPI>(net->ctx, g, split_k, &done)));
    }
    if (done) return SZB_OK;
    if (net->precision == 2)
        return narrow ? tc::launch_gemm_tc<64, 1, EPI>(net->ctx, g, split_k) : tc::launch_gemm_tc<128, 1, EPI>(net->ctx, g, split_k);
    return narrow ? tc::launch_gemm_tc<64, 3, EPI>(net->ctx, g, split_k) : tc::launch_gemm_tc<128, 3, EPI>(net->ctx, g, split_k);
}

// What the layer-3 epilogue of a training step needs to finish the forward pass itself (tma_epilogue_softmax)
struct SoftmaxCe {
    const uint32_t* labels; const float* target_vec; const uint8_t* valid;
    float* tail;     // [n_used, loss] block of the gradient vector this step accumulates into
    float* zT;       // transposed delta3 [n_out][B]
};

// forward for rows already in d_x (device); leaves logits in a_z; activations in a_h1 / a_h2.
// With `ce` (training, tensor-core path) layer 3 may run softmax / cross-entropy in its epilogue: *ce_done then says that a_z
// already holds delta3 (and ce->zT its transpose) and the loss / count are in ce->tail.
static szb_status forward_rows(szb_net* net, const float* d_x, int B, bool training = false, const SoftmaxCe* ce = nullptr,
                               bool* ce_done = nullptr) {
    if (ce_done) *ce_done = false;
    szb_ctx* ctx = net->ctx;
    float* P = net->params.as<float>();
    if (net->precision != 0) {
        SZB_TRY(refresh_wt(net));
        const float* WT = net->wt.as<float>();
        const int I = int(net->n_in), H1 = int(net->h1), H2 = int(net->h2), C = int(net->n_out);
        tc::GemmArgs g{};
        g.A = d_x; g.lda = I; g.B = WT + net->off_wt1(); g.ldb = I; g.C = net->a_h1.as<float>(); g.ldc = H1;
        g.CT = training ? net->h1T.as<float>() : nullptr; g.ldct = B; g.bias = P + net->off_b1(); g.M = B; g.N = H1; g.K = I;
        SZB_TRY(gemm_tc<tc::TC_BIAS_RELU>(net, g));
        g = tc::GemmArgs{};
        g.A = net->a_h1.as<float>(); g.lda = H1; g.B = WT + net->off_wt2(); g.ldb = H1; g.C = net->a_h2.as<float>(); g.ldc = H2;
        g.CT = training ? net->h2T.as<float>() : nullptr; g.ldct = B; g.bias = P + net->off_b2(); g.M = B; g.N = H2; g.K = H1;
        SZB_TRY(gemm_tc<tc::TC_BIAS_TANH>(net, g));
        g = tc::GemmArgs{};
        g.A = net->a_h2.as<float>(); g.lda = H2; g.B = WT + net->off_wt3(); g.ldb = H2; g.C = net->a_z.as<float>(); g.ldc = C;
        g.bias = P + net->off_b3(); g.M = B; g.N = C; g.K = H2;
        if (ce && ce_done && C <= 128) {
            tc::GemmArgs gs = g;
            gs.CT = ce->zT; gs.ldct = B; gs.labels = ce->labels; gs.target_vec = ce->target_vec; gs.valid = ce->valid; gs.tail = ce->tail;
            if (net->precision == 2) SZB_TRY(tc::launch_gemm_tma_softmax<1>(net->ctx, gs, ce_done));
            else SZB_TRY(tc::launch_gemm_tma_softmax<3>(net->ctx, gs, ce_done));
            if (*ce_done) return SZB_OK;
        }
        SZB_TRY(gemm_tc<tc::TC_BIAS>(net, g));
        return SZB_OK;
    }
    SZB_TRY((gemm<false, false, EPI_BIAS_RELU>(ctx, B, net->h1, net->n_in, d_x, net->n_in, P + net->off_w1(), net->h1,
                                                net->a_h1.as<float>(), net->h1, P + net->off_b1(), nullptr, 0)));
    SZB_TRY((gemm<false, false, EPI_BIAS_TANH>(ctx, B, net->h2, net->h1, net->a_h1.as<float>(), net->h1, P + net->off_w2(),
                                                net->h2, net->a_h2.as<float>(), net->h2, P + net->off_b2(), nullptr, 0)));
    SZB_TRY((gemm<false, false, EPI_BIAS>(ctx, B, net->n_out, net->h2, net->a_h2.as<float>(), net->h2, P + net->off_w3(),
                                           net->n_out, net->a_z.as<float>(), net->n_out, P + net->off_b3(), nullptr, 0)));
    return SZB_OK;
}


// Second hidden layer for rows in d_x: relu2 == 0 -> tanh (embed), 1 -> ReLU (forward_embedding).  out: [B][h2].
static szb_status embed_rows(szb_net* net, const float* d_x, int B, int relu2, float* d_out) {
    szb_ctx* ctx = net->ctx;
    float* P = net->params.as<float>();
    const int I = int(net->n_in), H1 = int(net->h1), H2 = int(net->h2);
    if (net->precision != 0) {
        SZB_TRY(refresh_wt(net));
        const float* WT = net->wt.as<float>();
        tc::GemmArgs g{};
        g.A = d_x; g.lda = I; g.B = WT + net->off_wt1(); g.ldb = I; g.C = net->a_h1.as<float>(); g.ldc = H1;
        g.bias = P + net->off_b1(); g.M = B; g.N = H1; g.K = I;
        SZB_TRY(gemm_tc<tc::TC_BIAS_RELU>(net, g));
        g = tc::GemmArgs{};
        g.A = net->a_h1.as<float>(); g.lda = H1; g.B = WT + net->off_wt2(); g.ldb = H1; g.C = d_out; g.ldc = H2;
        g.bias = P + net->off_b2(); g.M = B; g.N = H2; g.K = H1;
        return relu2 ? gemm_tc<tc::TC_BIAS_RELU>(net, g) : gemm_tc<tc::TC_BIAS_TANH>(net, g);
    }
    SZB_TRY((gemm<false, false, EPI_BIAS_RELU>(ctx, B, H1, I, d_x, I, P + net->off_w1(), H1, net->a_h1.as<float>(), H1, P + net->off_b1(),
                                                nullptr, 0)));
    if (relu2)
        return gemm<false, false, EPI_BIAS_RELU>(ctx, B, H2, H1, net->a_h1.as<float>(), H1, P + net->off_w2(), H2, d_out, H2,
                                                 P + net->off_b2(), nullptr, 0);
    return gemm<false, false, EPI_BIAS_TANH>(ctx, B, H2, H1, net->a_h1.as<float>(), H1, P + net->off_w2(), H2, d_out, H2, P + net->off_b2(),
                                             nullptr, 0);
}

static szb_status launch_softmax(szb_net* net, int B, int mode, float* probs, const uint32_t* labels, const float* target_vec,
                                 const uint8_t* valid, float threshold, unsigned long long* hist, float* sums, float* zT = nullptr) {
    if (B <= 0) return SZB_OK;
    const int wpb = 8;
    float* tail = net->grads.as<float>() + net->n_params() + 4 * net->tail_parity;
    if (mode == 2 && zT && net->n_out <= 128) {   // tensor-core training path
        SZB_CUDA(launch_pdl(net->ctx, softmax_train_kernel, dim3((B + 31) / 32), dim3(1024), 0, net->a_z.as<float>(), B, int(net->n_out),
                            labels, target_vec, valid, tail, zT, B));
        net->ctx->launches += 1;
        return SZB_OK;
    }
    softmax_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, net->ctx->stream>>>(net->a_z.as<float>(), B, int(net->n_out),
                                                                          int(net->n_out), mode, probs, labels, target_vec,
                                                                          valid, tail, threshold, hist, sums, zT, B);
    SZB_CUDA(cudaGetLastError());
    net->ctx->launches += 1;
    return SZB_OK;
}

// One optimiser step on the B rows staged in net->xb (+ lab / valid): forward, CE backward, [all-reduce], SGD.
szb_status comm_allreduce_f32(szb_ctx* ctx, float* buf, size_t n);  // comm.cu
szb_status comm_allreduce_overlapped(szb_ctx* ctx, float* buf, size_t n);
szb_status comm_join(szb_ctx* ctx);

// Where the step's batch sits: the net's own batch buffers (filled by launch_prep right before the step), or one step's slice
// of the chunk buffers that a multi-step launch of the batch kernel filled (szb_net_train_epoch_steps_dev).
struct BatchView {
    const float* xb; const float* xbT; const uint32_t* lab; const uint8_t* valid;
};

static szb_status train_step_staged(szb_net* net, int B, const float* target_vec, float lr, StepParams* sp = nullptr,
                                    const BatchView* view = nullptr) {
    szb_ctx* ctx = net->ctx;
    float* P = net->params.as<float>();
    const size_t np = net->n_params();
    // multi-GPU with peer-mapped exchange buffers (comm.cu).  The gradient is accumulated in private memory (the split-K
    // atomics of the weight-gradient GEMMs run 2.2x slower on IPC-exported memory: 105 vs 48 ms per epoch at N = 2) and the
    // update kernel scatters it to the slice owners itself -- no staging copy.
    const bool p2p = ctx->world > 1 && ctx->p2p_on && np + kGradTail + 4 * size_t(ctx->world) + 4 <= ctx->p2p_cap;
    float* G = net->grads.as<float>();
    // single-context tensor-core steps end in sgd_fused_kernel, which leaves the gradient vector zeroed for the next step
    const bool fused = net->precision != 0;   // (with the peer exchange the same kernel also sums the ranks' gradients)
    if (!net->grads_zero || (!fused && net->tail_parity != 0)) {
        SZB_CUDA(cudaMemsetAsync(G, 0, (np + kGradTail) * sizeof(float), ctx->stream));
        net->tail_parity = 0;
    }
    net->grads_zero = false;
    P2pArgs a{};
    a.world = 1;
    int p2p_blocks = 0;
    if (p2p) {
        const size_t nv = np + kGradTail;
        a.n4 = uint32_t((nv + 3) / 4);
        a.slice_g = (a.n4 + uint32_t(ctx->world) - 1) / uint32_t(ctx->world);
        for (int r = 0; r < ctx->world; ++r) {
            a.inbox[r] = ctx->p2p_inbox[r];
            a.red[r] = ctx->p2p_red[r];
            a.area[r] = ctx->p2p_inbox[r];
            a.flags[r] = ctx->p2p_flags[r];
        }
        a.cap = ctx->p2p_cap;
        // one round of flags instead of two while (world - 1) copies of the vector are a few microseconds of NVLink time
        a.one_shot = ctx->p2p_mode == 1 || (ctx->p2p_mode != 2 && size_t(ctx->world - 1) * nv * sizeof(float) <= (size_t(6) << 20)) ? 1 : 0;
        // flag inside the packets (no flag round): tensor-core update kernel only; a packet slot holds cap / 2 elements
        if ((ctx->p2p_mode == 3 || (ctx->p2p_mode == 0 && ctx->p2p_ll_auto)) && net->precision != 0 && 2 * nv <= ctx->p2p_cap &&
            size_t(ctx->world - 1) * nv * 8 <= (size_t(32) << 20))
            a.one_shot = 2;
        a.counters = ctx->p2p_counters.as<unsigned int>();
        a.trace = ctx->p2p_trace_on ? reinterpret_cast<unsigned long long*>(ctx->p2p_counters.as<unsigned char>() + 64) : nullptr;
        a.rank = ctx->rank; a.world = ctx->world; a.step = ++ctx->p2p_step;
        // the exchange makes the CTAs of a launch wait for one another: the whole grid has to be resident
        if (ctx->p2p_max_blocks == 0) {
            int per_sm_fused = 0, per_sm_plain = 0;
            SZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_fused, sgd_fused_kernel, 256, 0));
            SZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_plain, sgd_p2p_kernel, 256, 0));
            ctx->p2p_max_blocks = std::max(1, std::min(per_sm_fused, per_sm_plain)) * ctx->sm_count;
        }
        p2p_blocks = ctx->p2p_max_blocks;
    }
    bool reduced = p2p;     // gradient slices already all-reduced (overlapped) inside the backward pass, or exchanged below
    if (B > 0) {
        const BatchView own{net->xb.as<float>(), net->xbT.as<float>(), net->lab.as<uint32_t>(), net->valid.as<uint8_t>()};
        const BatchView& bv = view ? *view : own;
        const float* xb = bv.xb;
        const bool use_tc = net->precision != 0;
        // layer 3 ends in softmax / cross-entropy: inside the GEMM's epilogue when a tile holds a whole row of logits, else
        // as a kernel of its own
        SoftmaxCe ce{bv.lab, target_vec, bv.valid, net->grads.as<float>() + np + 4 * net->tail_parity, net->zT.as<float>()};
        bool ce_done = false;
        SZB_TRY(forward_rows(net, xb, B, true, (use_tc && (ctx->step_fuse & 1)) ? &ce : nullptr, &ce_done));
        if (!ce_done)
            SZB_TRY(launch_softmax(net, B, 2, nullptr, bv.lab, target_vec, bv.valid, 0.f, nullptr, nullptr, use_tc ? net->zT.as<float>() : nullptr));
        float* d3 = net->a_z.as<float>();
        const int C = int(net->n_out), H1 = int(net->h1), H2 = int(net->h2), I = int(net->n_in);
        if (use_tc) {
            auto split_for = [&](int M, int N) {
                const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
                return std::max(1, ctx->sm_count / tiles);
            };
            tc::GemmArgs gw[3]{}, gx2{}, gx1{};
            // gW3[h2][C] = H2^T dZ                                                    (lib.rs:1029-1032)
            gw[0].A = net->h2T.as<float>(); gw[0].lda = B; gw[0].B = net->zT.as<float>(); gw[0].ldb = B; gw[0].C = G + net->off_w3(); gw[0].ldc = C;
            gw[0].M = H2 + 1; gw[0].N = C; gw[0].K = B;   // row H2 of the product is sum_b dZ = the b3 gradient, which sits right
                                                          // behind w3 in the flattened gradient (lib.rs:1033)
            // d2 = (dZ W3^T) * (1 - H2^2)                                              (lib.rs:1034)
            gx2.A = d3; gx2.lda = C; gx2.B = P + net->off_w3(); gx2.ldb = C; gx2.C = net->d_2.as<float>(); gx2.ldc = H2; gx2.CT = net->d2T.as<float>();
            gx2.ldct = B; gx2.aux = net->h2T.as<float>(); gx2.ldaux = B; gx2.M = B; gx2.N = H2; gx2.K = C;   // factor read from H2^T: coalesced
            // gW2[h1][h2] = H1^T d2                                                    (lib.rs:1035-1037)
            gw[1].A = net->h1T.as<float>(); gw[1].lda = B; gw[1].B = net->d2T.as<float>(); gw[1].ldb = B; gw[1].C = G + net->off_w2(); gw[1].ldc = H2;
            gw[1].M = H1 + 1; gw[1].N = H2; gw[1].K = B;  // + b2 gradient (lib.rs:1038)
            // d1 = (d2 W2^T) * [H1 > 0]                                                (lib.rs:1039-1040)
            // only d1^T is consumed (by the layer-1 weight gradient): the row-major copy is not written
            gx1.A = net->d_2.as<float>(); gx1.lda = H2; gx1.B = P + net->off_w2(); gx1.ldb = H2; gx1.C = nullptr; gx1.ldc = H1;
            gx1.CT = net->d1T.as<float>(); gx1.ldct = B; gx1.aux = net->h1T.as<float>(); gx1.ldaux = B; gx1.M = B; gx1.N = H1; gx1.K = H2;
            // gW1[n_in][h1] = X^T d1                                                   (lib.rs:1041-1043)
            gw[2].A = bv.xbT; gw[2].lda = B; gw[2].B = net->d1T.as<float>(); gw[2].ldb = B; gw[2].C = G + net->off_w1(); gw[2].ldc = H1;
            gw[2].M = I + 1; gw[2].N = H1; gw[2].K = B;   // + b1 gradient (lib.rs:1044)
            bool grouped = false;
            if ((ctx->step_fuse & 4) && (p2p || ctx->world == 1)) {
                // The weight gradients depend only on activations and deltas: the two input-gradient GEMMs (the dependent chain)
                // go first, then all three weight gradients share ONE launch.  (With NCCL the per-layer order below is kept: each
                // layer's all-reduce overlaps the GEMMs of the layers before it.)
                SZB_TRY(gemm_tc<tc::TC_MUL_DTANH>(net, gx2));
                SZB_TRY(gemm_tc<tc::TC_MUL_DRELU>(net, gx1));
                // one-shot peer exchange: the launch also stores every finished gradient tile into the peers (PeerPush)
                tc::PeerPush push{};
                bool pushed = false;
                if (p2p && a.one_shot && ctx->p2p_early_push) {
                    for (int r = 0; r < ctx->world; ++r) push.area[r] = a.area[r];
                    push.slot_off = (size_t(a.step & 1u) * size_t(a.world) + size_t(a.rank)) * (a.cap / 4) * 4;
                    push.G = G;
                    push.counters = ctx->p2p_counters.as<unsigned int>() + 64;
                    push.world = a.world; push.rank = a.rank;
                }
                if (net->precision == 2) SZB_TRY(tc::launch_gemm_tma_group<1>(ctx, gw, 3, &grouped, &push, &pushed));
                else SZB_TRY(tc::launch_gemm_tma_group<3>(ctx, gw, 3, &grouped, &push, &pushed));
                if (pushed) a.sent4 = uint32_t(np / 4);       // [0, np) = the three products' outputs; the tail (and the group that
                                                              // straddles np) is still sent by the update kernel
                if (!grouped) {
                    if (net->precision == 2) SZB_TRY(tc::launch_gemm_tc_group<1>(ctx, gw, 3, &grouped));
                    else SZB_TRY(tc::launch_gemm_tc_group<3>(ctx, gw, 3, &grouped));
                }
                if (!grouped) {
                    SZB_TRY(gemm_tc<tc::TC_ATOMIC>(net, gw[0], split_for(H2 + 1, C)));
                    SZB_TRY(gemm_tc<tc::TC_ATOMIC>(net, gw[1], split_for(H1 + 1, H2)));
                    SZB_TRY(gemm_tc<tc::TC_ATOMIC>(net, gw[2], split_for(I + 1, H1)));
                }
            } else {
                SZB_TRY(gemm_tc<tc::TC_ATOMIC>(net, gw[0], split_for(H2 + 1, C)));
                // multi-GPU: [w3 | b3 | n_used, loss] is final -> reduce it across ranks while layers 2 and 1 run
                if (!p2p) SZB_TRY(comm_allreduce_overlapped(ctx, G + net->off_w3(), np + kGradTail - net->off_w3()));
                SZB_TRY(gemm_tc<tc::TC_MUL_DTANH>(net, gx2));
                SZB_TRY(gemm_tc<tc::TC_ATOMIC>(net, gw[1], split_for(H1 + 1, H2)));
                if (!p2p) SZB_TRY(comm_allreduce_overlapped(ctx, G + net->off_w2(), net->off_w3() - net->off_w2()));
                SZB_TRY(gemm_tc<tc::TC_MUL_DRELU>(net, gx1));
                SZB_TRY(gemm_tc<tc::TC_ATOMIC>(net, gw[2], split_for(I + 1, H1)));
                if (!p2p) {
                    SZB_TRY(comm_allreduce_overlapped(ctx, G, net->off_w2()));
                    SZB_TRY(comm_join(ctx));
                }
            }
            reduced = true;
        } else {
        // layer 3
        SZB_TRY((gemm<true, false, EPI_ATOMIC>(ctx, H2, C, B, net->a_h2.as<float>(), H2, d3, C, G + net->off_w3(), C, nullptr,
                                                nullptr, 0)));
        SZB_TRY((gemm<false, true, EPI_MUL_DTANH>(ctx, B, H2, C, d3, C, P + net->off_w3(), C, net->d_2.as<float>(), H2, nullptr,
                                                   net->a_h2.as<float>(), H2)));
        // layer 2
        SZB_TRY((gemm<true, false, EPI_ATOMIC>(ctx, H1, H2, B, net->a_h1.as<float>(), H1, net->d_2.as<float>(), H2,
                                                G + net->off_w2(), H2, nullptr, nullptr, 0)));
        SZB_TRY((gemm<false, true, EPI_MUL_DRELU>(ctx, B, H1, H2, net->d_2.as<float>(), H2, P + net->off_w2(), H2,
                                                   net->d_1.as<float>(), H1, nullptr, net->a_h1.as<float>(), H1)));
        // layer 1
        SZB_TRY((gemm<true, false, EPI_ATOMIC>(ctx, I, H1, B, xb, I, net->d_1.as<float>(), H1, G + net->off_w1(), H1, nullptr,
                                                nullptr, 0)));
        // bias gradients
        auto colsum = [&](const float* D, int cols, float* g) -> szb_status {
            const int rpb = 128;
            dim3 grid((cols + 127) / 128, (B + rpb - 1) / rpb);
            colsum_kernel<<<grid, 128, 0, ctx->stream>>>(D, B, cols, cols, g, rpb);
            SZB_CUDA(cudaGetLastError());
            ctx->launches += 1;
            return SZB_OK;
        };
        SZB_TRY(colsum(d3, C, G + net->off_b3()));
        SZB_TRY(colsum(net->d_2.as<float>(), H2, G + net->off_b2()));
        SZB_TRY(colsum(net->d_1.as<float>(), H1, G + net->off_b1()));
        }
    }
    if (ctx->world > 1 && !reduced) {
        if (net->precision != 0) {
            // a rank whose slice of this batch is empty must still issue the SAME collectives as its peers
            SZB_TRY(comm_allreduce_overlapped(ctx, G + net->off_w3(), np + kGradTail - net->off_w3()));
            SZB_TRY(comm_allreduce_overlapped(ctx, G + net->off_w2(), net->off_w3() - net->off_w2()));
            SZB_TRY(comm_allreduce_overlapped(ctx, G, net->off_w2()));
            SZB_TRY(comm_join(ctx));
        } else {
            SZB_TRY(comm_allreduce_f32(ctx, G, np + kGradTail));
        }
    }
    if (fused) {
        SZB_TRY(net->wt.reserve(net->n_wt() * 4));
        const int tiles = int(((net->n_in + 31) / 32) * ((net->h1 + 31) / 32) + ((net->h1 + 31) / 32) * ((net->h2 + 31) / 32) +
                              ((net->h2 + 31) / 32) * ((net->n_out + 31) / 32));
        int blocks = std::max(1, std::min(tiles + 1, ctx->sm_count * 4));
        if (p2p) blocks = std::min(blocks, p2p_blocks);
        SZB_CUDA(launch_pdl(ctx, sgd_fused_kernel, dim3(blocks), dim3(256), 0, P, G, net->wt.as<float>(), int(net->n_in), int(net->h1),
                            int(net->h2), int(net->n_out), net->off_b1(), net->off_w2(), net->off_b2(), net->off_w3(), net->off_b3(),
                            net->off_wt2(), net->off_wt3(), np, net->tail_parity, lr, net->stats.as<double>(), a, sp, B));
    } else if (p2p) {
        const int blocks = int(std::min<size_t>((np / 4 + 255) / 256, size_t(std::min(ctx->sm_count * 2, p2p_blocks))));
        sgd_p2p_kernel<<<std::max(1, blocks), 256, 0, ctx->stream>>>(P, G, a, np, lr, net->stats.as<double>());
    } else {
        const int blocks = int(std::min<size_t>((np + 255) / 256, size_t(ctx->sm_count) * 4));
        sgd_kernel<<<blocks, 256, 0, ctx->stream>>>(P, G, np, lr, net->stats.as<double>());
    }
    SZB_CUDA(cudaGetLastError());
    ctx->launches += 1;
    if (fused) {
        net->wt_dirty = false;
        net->grads_zero = true;
        net->tail_parity ^= 1;
    } else {
        net->wt_dirty = true;
    }
    return SZB_OK;
}

static PrepArgs make_prep_args(szb_net* net, const float* d_feats, const uint32_t* d_labels, const uint32_t* d_perm, int B,
                               const uint8_t* d_keep, int keep_by_row, float prob, unsigned long long key, size_t* tile_bytes) {
    PrepArgs a{};
    a.feats = d_feats; a.labels_all = d_labels; a.perm = d_perm; a.B = B; a.n_in = int(net->n_in);
    a.keep = d_keep; a.keep_by_row = keep_by_row; a.prob = prob; a.key = key;
    a.xb = net->xb.as<float>(); a.xbT = net->precision != 0 ? net->xbT.as<float>() : nullptr;
    a.lab = net->lab.as<uint32_t>(); a.valid = net->valid.as<uint8_t>();
    a.h1T = net->h1T.as<float>(); a.h1 = int(net->h1); a.h2T = net->h2T.as<float>(); a.h2 = int(net->h2);
    const size_t bytes = size_t(net->n_in) * 33 * sizeof(float);
    a.use_tile = bytes <= 48 * 1024 ? 1 : 0;
    *tile_bytes = a.use_tile ? bytes : size_t(0);
    return a;
}

static szb_status launch_prep(szb_net* net, const float* d_feats, const uint32_t* d_labels, const uint32_t* d_perm, int B,
                              const uint8_t* d_keep, int keep_by_row, float prob, unsigned long long key, StepParams* sp = nullptr) {
    if (B <= 0) return SZB_OK;
    const int wpb = B >= 32 ? 32 : 8;     // one warp per row when the block's 32 rows exist: every gather load in flight at once
    size_t tile_bytes = 0;
    const PrepArgs a = make_prep_args(net, d_feats, d_labels, d_perm, B, d_keep, keep_by_row, prob, key, &tile_bytes);
    SZB_CUDA(launch_pdl(net->ctx, prep_batch_kernel, dim3((B + 31) / 32), dim3(wpb * 32), tile_bytes, a, sp));
    net->ctx->launches += 1;
    return SZB_OK;
}

// Small batches are launch-latency-bound (11 kernels per step, a few microseconds of work each; BASELINE configs[0]: batch 8):
// two consecutive steps (one of each tail parity) are captured ONCE as a CUDA graph -- programmatic dependent launch edges
// included -- and replayed for the rest of the epoch and for later epochs; what changes from step to step or epoch to epoch
// (position in the permutation, dropout key, learning rate) is read from StepParams in device memory.
static szb_status capture_step_graph(szb_net* net, const StepGraphKey& k, float prob) {
    szb_ctx* ctx = net->ctx;
    if (net->step_graph) { cudaGraphExecDestroy(net->step_graph); net->step_graph = nullptr; }
    const uint64_t launches0 = ctx->launches;
    const uint32_t p2p_step0 = ctx->p2p_step;           // a captured step takes its exchange step number from device memory
    StepParams* sp = net->step_params.as<StepParams>();
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); net->step_graph_failed = true; return SZB_OK; }
    szb_status st = SZB_OK;
    for (int rep = 0; rep < 2 && st == SZB_OK; ++rep) {
        st = launch_prep(net, static_cast<const float*>(k.feats), static_cast<const uint32_t*>(k.labels), static_cast<const uint32_t*>(k.perm), k.B,
                         static_cast<const uint8_t*>(k.keep), 1, prob, 0ull, sp);
        if (st == SZB_OK) st = train_step_staged(net, k.B, nullptr, 0.f, sp);
    }
    const cudaError_t e_end = cudaStreamEndCapture(ctx->stream, &graph);
    net->step_graph_kernels = uint32_t(ctx->launches - launches0);   // kernels in one replay (two steps)
    ctx->launches = launches0;                          // nothing ran
    ctx->p2p_step = p2p_step0;
    if (st != SZB_OK || e_end != cudaSuccess || !graph || cudaGraphInstantiate(&net->step_graph, graph, 0) != cudaSuccess) {
        cudaGetLastError();
        net->step_graph = nullptr;
        net->step_graph_failed = true;                  // e.g. a driver without programmatic edges in captured graphs: stay on plain launches
    } else {
        net->step_graph_key = k;
    }
    if (graph) cudaGraphDestroy(graph);
    return SZB_OK;
}

static szb_status net_alloc(szb_ctx* ctx, uint32_t n_in, uint32_t h1, uint32_t h2, uint32_t n_out, szb_net** out) {
    SZB_REQUIRE(ctx && out, "net: NULL argument");
    SZB_REQUIRE(n_in > 0 && h1 > 0 && h2 > 0 && n_out > 0, "net: layer sizes must be positive (%u,%u,%u,%u)", n_in, h1, h2, n_out);
    SZB_CUDA(cudaSetDevice(ctx->device));
    szb_net* net = new szb_net();
    net->ctx = ctx;
    net->n_in = n_in; net->h1 = h1; net->h2 = h2; net->n_out = n_out;
    szb_status s = net->params.reserve(net->n_params() * 4);
    if (s == SZB_OK) s = net->grads.reserve((net->n_params() + kGradTail) * 4);
    if (s == SZB_OK) s = net->stats.reserve(2 * sizeof(double));
    if (s != SZB_OK) { szb_net_destroy(net); return s; }
    cudaMemsetAsync(net->params.ptr, 0, net->n_params() * 4, ctx->stream);
    cudaMemsetAsync(net->stats.ptr, 0, 2 * sizeof(double), ctx->stream);
    *out = net;
    return SZB_OK;
}

}  // namespace szb

using namespace szb;

extern "C" {

szb_status szb_net_create(szb_ctx* ctx, uint32_t n_in, uint32_t h1, uint32_t h2, uint32_t n_out, uint64_t seed, szb_net** out) {
    SZB_TRY(net_alloc(ctx, n_in, h1, h2, n_out, out));
    szb_net* net = *out;
    const unsigned long long key = splitmix64(seed ^ 0x57A3A2B200ull);
    float* P = net->params.as<float>();
    struct { size_t off, n; } blocks[3] = { { net->off_w1(), size_t(n_in) * h1 }, { net->off_w2(), size_t(h1) * h2 },
                                            { net->off_w3(), size_t(h2) * n_out } };
    for (auto& b : blocks) {
        init_uniform_kernel<<<int(std::min<size_t>((b.n + 255) / 256, 1024)), 256, 0, ctx->stream>>>(P + b.off, b.n, key, b.off);
        SZB_CUDA(cudaGetLastError());
        ctx->launches += 1;
    }
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

szb_status szb_net_from_weights(szb_ctx* ctx, uint32_t n_in, uint32_t h1, uint32_t h2, uint32_t n_out, const float* w1,
                                const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                                szb_net** out) {
    SZB_REQUIRE(w1 && b1 && w2 && b2 && w3 && b3, "szb_net_from_weights: NULL weights");
    SZB_TRY(net_alloc(ctx, n_in, h1, h2, n_out, out));
    szb_net* net = *out;
    std::vector<float> flat(net->n_params());
    std::memcpy(&flat[net->off_w1()], w1, size_t(n_in) * h1 * 4);
    std::memcpy(&flat[net->off_b1()], b1, size_t(h1) * 4);
    std::memcpy(&flat[net->off_w2()], w2, size_t(h1) * h2 * 4);
    std::memcpy(&flat[net->off_b2()], b2, size_t(h2) * 4);
    std::memcpy(&flat[net->off_w3()], w3, size_t(h2) * n_out * 4);
    std::memcpy(&flat[net->off_b3()], b3, size_t(n_out) * 4);
    SZB_CUDA(cudaMemcpyAsync(net->params.ptr, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

szb_status szb_net_get_weights(szb_net* net, float* w1, float* b1, float* w2, float* b2, float* w3, float* b3) {
    SZB_REQUIRE(net, "szb_net_get_weights: net is NULL");
    std::vector<float> flat(net->n_params());
    SZB_CUDA(cudaSetDevice(net->ctx->device));
    SZB_CUDA(cudaMemcpyAsync(flat.data(), net->params.ptr, flat.size() * 4, cudaMemcpyDeviceToHost, net->ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(net->ctx->stream));
    if (w1) std::memcpy(w1, &flat[net->off_w1()], size_t(net->n_in) * net->h1 * 4);
    if (b1) std::memcpy(b1, &flat[net->off_b1()], size_t(net->h1) * 4);
    if (w2) std::memcpy(w2, &flat[net->off_w2()], size_t(net->h1) * net->h2 * 4);
    if (b2) std::memcpy(b2, &flat[net->off_b2()], size_t(net->h2) * 4);
    if (w3) std::memcpy(w3, &flat[net->off_w3()], size_t(net->h2) * net->n_out * 4);
    if (b3) std::memcpy(b3, &flat[net->off_b3()], size_t(net->n_out) * 4);
    return SZB_OK;
}

szb_status szb_net_dims(const szb_net* net, uint32_t dims[4]) {
    SZB_REQUIRE(net && dims, "szb_net_dims: NULL argument");
    dims[0] = net->n_in; dims[1] = net->h1; dims[2] = net->h2; dims[3] = net->n_out;
    return SZB_OK;
}
uint32_t szb_net_output_size(const szb_net* net) { return net ? net->n_out : 0; }

szb_status szb_net_set_precision(szb_net* net, int32_t mode) {
    SZB_REQUIRE(net && mode >= 0 && mode <= 2, "szb_net_set_precision: mode must be 0 (FP32), 1 (3xTF32) or 2 (TF32)");
    net->precision = mode;
    net->wt_dirty = true;
    return SZB_OK;
}
int32_t szb_net_get_precision(const szb_net* net) { return net ? net->precision : -1; }

szb_status szb_net_add_output_class(szb_net* net, const float* new_col, uint64_t seed) {
    SZB_REQUIRE(net, "szb_net_add_output_class: net is NULL");
    const uint32_t C = net->n_out, H2 = net->h2;
    std::vector<float> w1(size_t(net->n_in) * net->h1), b1(net->h1), w2(size_t(net->h1) * H2), b2(H2), w3(size_t(H2) * C), b3(C);
    SZB_TRY(szb_net_get_weights(net, w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data()));
    std::vector<float> nw3(size_t(H2) * (C + 1)), nb3(C + 1, 0.f);
    const unsigned long long key = splitmix64(seed ^ 0xADD0C1A55ull);
    for (uint32_t r = 0; r < H2; ++r) {
        for (uint32_t c = 0; c < C; ++c) nw3[size_t(r) * (C + 1) + c] = w3[size_t(r) * C + c];   // lib.rs:803-806
        nw3[size_t(r) * (C + 1) + C] = new_col ? new_col[r]                                         // lib.rs:807-809
                                               : float(uint32_t(splitmix64(key ^ r) >> 40)) * 5.9604644775390625e-08f - 0.5f;
    }
    for (uint32_t c = 0; c < C; ++c) nb3[c] = b3[c];                                                // lib.rs:812-815
    szb_net* fresh = nullptr;
    SZB_TRY(szb_net_from_weights(net->ctx, net->n_in, net->h1, H2, C + 1, w1.data(), b1.data(), w2.data(), b2.data(), nw3.data(),
                                 nb3.data(), &fresh));
    // adopt the new buffers in place so the caller's handle stays valid
    net->params.release(); net->grads.release();
    net->params = fresh->params; net->grads = fresh->grads;
    fresh->params = DevBuf(); fresh->grads = DevBuf();
    net->n_out = C + 1;
    net->wt_dirty = true;
    net->grads_zero = false;      // fresh, uninitialised gradient buffer
    net->tail_parity = 0;
    net->a_z.release(); net->zT.release(); net->cap_rows = 0;
    szb_net_destroy(fresh);
    return SZB_OK;
}

void szb_net_destroy(szb_net* net) {
    if (!net) return;
    if (net->ctx) { cudaSetDevice(net->ctx->device); cudaStreamSynchronize(net->ctx->stream); }
    if (net->step_graph) cudaGraphExecDestroy(net->step_graph);
    net->step_params.release();
    net->small_scratch.release(); net->small_barrier.release(); net->small_steps.release();
    for (DevBuf* b : { &net->params, &net->grads, &net->xb, &net->lab, &net->valid, &net->a_h1, &net->a_h2, &net->a_z, &net->d_2,
                       &net->d_1, &net->stats, &net->perm, &net->hist, &net->wt, &net->xbT, &net->h1T, &net->h2T, &net->zT, &net->d2T, &net->d1T,
                       &net->chunk_xb, &net->chunk_xbT, &net->chunk_lab, &net->chunk_valid })
        b->release();
    delete net;
}

// Rows are processed in chunks so the activation scratch stays bounded.
static constexpr uint64_t kChunkRows = 1u << 16;

szb_status szb_net_forward_dev(szb_net* net, const float* d_x, uint64_t B, float* d_probs) {
    SZB_REQUIRE(net, "szb_net_forward_dev: net is NULL");
    if (B == 0) return SZB_OK;
    SZB_REQUIRE(d_x && d_probs, "szb_net_forward_dev: NULL buffer");
    SZB_CUDA(cudaSetDevice(net->ctx->device));
    SZB_TRY(net_reserve_rows(net, std::min(B, kChunkRows)));
    for (uint64_t r0 = 0; r0 < B; r0 += kChunkRows) {
        const int nb = int(std::min(kChunkRows, B - r0));
        SZB_TRY(forward_rows(net, d_x + r0 * net->n_in, nb));
        SZB_TRY(launch_softmax(net, nb, 1, d_probs + r0 * net->n_out, nullptr, nullptr, nullptr, 0.f, nullptr, nullptr));
    }
    return SZB_OK;
}

szb_status szb_net_forward(szb_net* net, const float* x, uint64_t B, float* probs) {
    SZB_REQUIRE(net, "szb_net_forward: net is NULL");
    if (B == 0) return SZB_OK;
    SZB_REQUIRE(x && probs, "szb_net_forward: NULL buffer");
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->x.reserve(B * net->n_in * 4));
    SZB_TRY(ctx->probs.reserve(B * net->n_out * 4));
    SZB_CUDA(cudaMemcpyAsync(ctx->x.ptr, x, B * net->n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(szb_net_forward_dev(net, ctx->x.as<float>(), B, ctx->probs.as<float>()));
    SZB_CUDA(cudaMemcpyAsync(probs, ctx->probs.ptr, B * net->n_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

static szb_status read_stats(szb_net* net, double* loss_sum, uint64_t* n_used) {
    double h[2] = { 0, 0 };
    SZB_CUDA(cudaMemcpyAsync(h, net->stats.ptr, sizeof h, cudaMemcpyDeviceToHost, net->ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(net->ctx->stream));
    if (loss_sum) *loss_sum = h[0];
    if (n_used) *n_used = uint64_t(h[1] + 0.5);
    return SZB_OK;
}

szb_status szb_net_train_batch(szb_net* net, const float* x, uint64_t B, const float* target, float lr) {
    SZB_REQUIRE(net, "szb_net_train_batch: net is NULL");
    if (B == 0) return SZB_OK;  // lib.rs:1003-1005
    SZB_REQUIRE(x && target, "szb_net_train_batch: NULL buffer");
    SZB_REQUIRE(B <= (1u << 24), "szb_net_train_batch: batch of %llu rows is too large", (unsigned long long)B);
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(net_reserve_rows(net, B));
    SZB_TRY(ctx->probs.reserve(size_t(net->n_out) * 4));
    SZB_TRY(ctx->x.reserve(B * net->n_in * 4));
    SZB_CUDA(cudaMemcpyAsync(ctx->x.ptr, x, B * net->n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    SZB_CUDA(cudaMemcpyAsync(ctx->probs.ptr, target, size_t(net->n_out) * 4, cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(launch_prep(net, ctx->x.as<float>(), nullptr, nullptr, int(B), nullptr, 0, 0.f, 0));
    SZB_CUDA(cudaMemsetAsync(net->valid.ptr, 1, B, ctx->stream));  // the reference's train_batch uses every row it is given
    SZB_TRY(train_step_staged(net, int(B), ctx->probs.as<float>(), lr));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

szb_status szb_net_train_batch_labels(szb_net* net, const float* x, const uint32_t* labels, uint64_t B, float lr,
                                      const uint8_t* keep, double* loss_sum, uint64_t* n_used) {
    SZB_REQUIRE(net, "szb_net_train_batch_labels: net is NULL");
    if (loss_sum) *loss_sum = 0.0;
    if (n_used) *n_used = 0;
    if (B == 0) return SZB_OK;
    SZB_REQUIRE(x && labels, "szb_net_train_batch_labels: NULL buffer");
    SZB_REQUIRE(B <= (1u << 24), "szb_net_train_batch_labels: batch of %llu rows is too large", (unsigned long long)B);
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(net_reserve_rows(net, B));
    SZB_TRY(ctx->x.reserve(B * net->n_in * 4));
    SZB_TRY(ctx->labels.reserve(B * 4));
    SZB_CUDA(cudaMemcpyAsync(ctx->x.ptr, x, B * net->n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    SZB_CUDA(cudaMemcpyAsync(ctx->labels.ptr, labels, B * 4, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* d_keep = nullptr;
    if (keep) {
        SZB_TRY(ctx->misc.reserve(B * net->n_in));
        SZB_CUDA(cudaMemcpyAsync(ctx->misc.ptr, keep, B * net->n_in, cudaMemcpyHostToDevice, ctx->stream));
        d_keep = ctx->misc.as<uint8_t>();
    }
    SZB_CUDA(cudaMemsetAsync(net->stats.ptr, 0, 2 * sizeof(double), ctx->stream));
    SZB_TRY(launch_prep(net, ctx->x.as<float>(), ctx->labels.as<uint32_t>(), nullptr, int(B), d_keep, 0, 0.f, 0));
    SZB_TRY(train_step_staged(net, int(B), nullptr, lr));
    return read_stats(net, loss_sum, n_used);
}

szb_status szb_net_train_epoch_steps_dev(szb_net* net, const float* d_feats, const uint32_t* d_labels, uint64_t n,
                                         const uint32_t* perm, uint64_t n_perm, const uint32_t* step_sizes, uint32_t n_steps, float lr,
                                         float dropout, uint64_t seed, uint64_t stream, const uint8_t* d_keep, double* loss_sum,
                                         uint64_t* n_used) {
    SZB_REQUIRE(net, "szb_net_train_epoch_steps_dev: net is NULL");
    if (loss_sum) *loss_sum = 0.0;
    if (n_used) *n_used = 0;
    if (n_steps == 0) return SZB_OK;
    SZB_REQUIRE(step_sizes && (perm || n_perm == 0), "szb_net_train_epoch_steps_dev: NULL buffer");
    SZB_REQUIRE(n <= 0xffffffffull, "szb_net_train_epoch_steps_dev: more than 2^32 windows");
    uint64_t total = 0, max_b = 0;
    for (uint32_t i = 0; i < n_steps; ++i) {
        total += step_sizes[i];
        max_b = std::max<uint64_t>(max_b, step_sizes[i]);
    }
    SZB_REQUIRE(total == n_perm, "szb_net_train_epoch_steps_dev: step sizes sum to %llu, perm has %llu rows", (unsigned long long)total,
                (unsigned long long)n_perm);
    SZB_REQUIRE(n_perm == 0 || (d_feats && d_labels), "szb_net_train_epoch_steps_dev: NULL device buffer");
    // the counter of the library's dropout stream packs the feature index into 6 bits (dropout_keep): wider inputs would
    // reuse the draws of the next window.  Explicit decisions (d_keep) have no such limit.
    SZB_REQUIRE(dropout <= 0.f || d_keep || net->n_in <= 64,
                "szb_net_train_epoch_steps_dev: the counter dropout stream supports n_in <= 64 (net has %u); pass d_keep", net->n_in);
    for (uint64_t i = 0; i < n_perm; ++i)
        SZB_REQUIRE(perm[i] < n, "szb_net_train_epoch_steps_dev: perm[%llu] = %u out of range", (unsigned long long)i, perm[i]);
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(net_reserve_rows(net, std::max<uint64_t>(1, max_b)));
    SZB_TRY(net->perm.reserve(std::max<uint64_t>(1, n_perm) * 4));
    if (n_perm) SZB_CUDA(cudaMemcpyAsync(net->perm.ptr, perm, n_perm * 4, cudaMemcpyHostToDevice, ctx->stream));
    SZB_CUDA(cudaMemsetAsync(net->stats.ptr, 0, 2 * sizeof(double), ctx->stream));
    const unsigned long long key = dropout_key(seed, stream);
    {   // batches of <= 32 rows (the reference's default is 8): the whole epoch in one persistent kernel (train_small.cu)
        bool done = false;
        SZB_TRY(train_epoch_small(net, d_feats, d_labels, net->perm.as<uint32_t>(), step_sizes, n_steps, lr, dropout, key, d_keep, &done));
        if (done) return read_stats(net, loss_sum, n_used);
    }
    uint64_t s = 0;
    uint32_t i = 0;
    auto plain_step = [&]() -> szb_status {
        const int B = int(step_sizes[i]);
        SZB_TRY(launch_prep(net, d_feats, d_labels, net->perm.as<uint32_t>() + s, B, d_keep, 1, dropout, key));
        SZB_TRY(train_step_staged(net, B, nullptr, lr));   // B == 0 still joins the all-reduce of a multi-GPU step
        s += step_sizes[i];
        ++i;
        return SZB_OK;
    };
    // ---- large batches: the batch kernel runs for kPrepChunk steps at a time (step_fuse bit 2) ------------------------------------
    // Gathering, dropping out and transposing a batch depends only on the permutation and the dropout stream, never on the
    // weights, so it does not have to sit between two steps: one launch fills the batch buffers of the next kPrepChunk equal-sized
    // steps (64 MB for 32 steps of 4096 windows), and each step reads its slice.  Per step that is 1/32 of a (larger, better
    // filled) launch instead of a ~4 us kernel on the critical path.  Batches of <= 256 rows replay a captured graph instead.
    if ((ctx->step_fuse & 2) && net->precision != 0 && step_sizes[0] > 256) {
        constexpr uint32_t kPrepChunk = 32;
        const uint32_t B0 = step_sizes[0];
        const size_t ni = net->n_in;
        while (i < n_steps && step_sizes[i] == B0) {
            uint32_t cnt = 0;
            while (i + cnt < n_steps && cnt < kPrepChunk && step_sizes[i + cnt] == B0) ++cnt;
            SZB_TRY(net->chunk_xb.reserve(size_t(cnt) * B0 * ni * 4));
            SZB_TRY(net->chunk_xbT.reserve(size_t(cnt) * B0 * (ni + 1) * 4));
            SZB_TRY(net->chunk_lab.reserve(size_t(cnt) * B0 * 4));
            SZB_TRY(net->chunk_valid.reserve(size_t(cnt) * B0));
            size_t tile_bytes = 0;
            PrepArgs a = make_prep_args(net, d_feats, d_labels, net->perm.as<uint32_t>() + s, int(B0), d_keep, 1, dropout, key, &tile_bytes);
            a.xb = net->chunk_xb.as<float>(); a.xbT = net->chunk_xbT.as<float>(); a.lab = net->chunk_lab.as<uint32_t>();
            a.valid = net->chunk_valid.as<uint8_t>();
            a.blocks_per_step = int((B0 + 31) / 32);
            SZB_CUDA(launch_pdl(ctx, prep_batch_kernel, dim3(cnt * a.blocks_per_step), dim3(1024), tile_bytes, a, static_cast<StepParams*>(nullptr)));
            ctx->launches += 1;
            for (uint32_t j = 0; j < cnt; ++j) {
                const BatchView bv{a.xb + size_t(j) * B0 * ni, a.xbT + size_t(j) * B0 * (ni + 1), a.lab + size_t(j) * B0, a.valid + size_t(j) * B0};
                SZB_TRY(train_step_staged(net, int(B0), nullptr, lr, nullptr, &bv));
            }
            s += uint64_t(cnt) * B0;
            i += cnt;
        }
    }
    // ---- small batches: replay a captured two-step graph (see capture_step_graph) -------------------------------------------
    const int B0 = int(step_sizes[0]);
    uint32_t n_full = 0;
    while (n_full < n_steps && int(step_sizes[n_full]) == B0) ++n_full;
    // multi-GPU: only with the peer-memory exchange (no collective call inside a step), and when the vector fits its buffers
    const bool peers_ok = ctx->world == 1 || (ctx->p2p_on && ctx->graph_peers && net->n_params() + kGradTail + 4 * size_t(ctx->world) + 4 <= ctx->p2p_cap);
    if (ctx->graphs && peers_ok && net->precision != 0 && !net->step_graph_failed && B0 > 0 && B0 <= ctx->graph_max_rows && n_full >= 8) {
        SZB_TRY(net->step_params.reserve(sizeof(StepParams)));
        StepGraphKey k;
        k.B = B0; k.precision = net->precision; k.feats = d_feats; k.labels = d_labels; k.keep = d_keep; k.perm = net->perm.ptr;
        k.params = net->params.ptr; k.prob = dropout; k.cap_rows = net->cap_rows; k.n_out = net->n_out;
        k.exchange = ctx->world > 1 ? 1 + ctx->p2p_mode : 0;
        if (!net->grads_zero || net->wt_dirty) {       // the first steps of a net run plainly: they set up the state a captured step assumes
            SZB_TRY(plain_step());
            SZB_TRY(plain_step());
        }
        k.parity = net->step_graph ? net->step_graph_key.parity : net->tail_parity;
        if (net->step_graph && net->step_graph_key == k && net->tail_parity != k.parity) SZB_TRY(plain_step());   // an odd step count left the
        k.parity = net->tail_parity;                                                                            // other tail block current
        if (!net->step_graph || !(net->step_graph_key == k)) SZB_TRY(capture_step_graph(net, k, dropout));
        if (net->step_graph && net->step_graph_key == k) {
            const uint32_t n_pairs = (n_full - i) / 2;
            void* hp = nullptr;
            SZB_TRY(ctx->h_stage.acquire(sizeof(StepParams), &hp));
            *static_cast<StepParams*>(hp) = StepParams{ uint32_t(s), lr, key, ctx->p2p_step, 0u };
            SZB_CUDA(cudaMemcpyAsync(net->step_params.ptr, hp, sizeof(StepParams), cudaMemcpyHostToDevice, ctx->stream));
            SZB_TRY(ctx->h_stage.uploaded(ctx->stream));
            for (uint32_t p = 0; p < n_pairs; ++p) SZB_CUDA(cudaGraphLaunch(net->step_graph, ctx->stream));
            ctx->launches += uint64_t(n_pairs) * net->step_graph_kernels;   // 2 x (batch kernel, the step's GEMMs, [softmax,] update)
            ctx->graph_launches += n_pairs;
            if (ctx->world > 1) ctx->p2p_step += 2 * n_pairs;      // the replayed steps counted it up in device memory
            i += 2 * n_pairs;
            s += uint64_t(2 * n_pairs) * uint64_t(B0);
        }
    }
    while (i < n_steps) SZB_TRY(plain_step());
    return read_stats(net, loss_sum, n_used);   // loss and count are global (all-reduced) in a multi-GPU run
}

szb_status szb_net_train_epoch_dev(szb_net* net, const float* d_feats, const uint32_t* d_labels, uint64_t n,
                                   const uint32_t* perm, uint64_t n_perm, uint32_t batch, float lr, float dropout, uint64_t seed,
                                   uint64_t stream, const uint8_t* d_keep, double* loss_sum, uint64_t* n_used) {
    SZB_REQUIRE(net, "szb_net_train_epoch_dev: net is NULL");
    if (batch == 0) batch = 1;  // lib.rs:602 batch_size.max(1)
    std::vector<uint32_t> sizes;
    for (uint64_t s = 0; s < n_perm; s += batch) sizes.push_back(uint32_t(std::min<uint64_t>(batch, n_perm - s)));
    return szb_net_train_epoch_steps_dev(net, d_feats, d_labels, n, perm, n_perm, sizes.data(), uint32_t(sizes.size()), lr, dropout, seed,
                                         stream, d_keep, loss_sum, n_used);
}

szb_status szb_dropout_keep_mask(uint64_t seed, uint64_t stream, const uint64_t* rows, uint64_t n_rows, uint32_t n_in, float prob,
                                 uint8_t* keep) {
    SZB_REQUIRE(keep && (rows || n_rows == 0), "szb_dropout_keep_mask: NULL argument");
    SZB_REQUIRE(n_in <= 64, "szb_dropout_keep_mask: n_in %u > 64 (the counter packs the feature index in 6 bits)", n_in);
    const unsigned long long key = dropout_key(seed, stream);
    for (uint64_t r = 0; r < n_rows; ++r)
        for (uint32_t i = 0; i < n_in; ++i)
            keep[r * n_in + i] = (prob <= 0.f || dropout_keep(key, rows[r], i, prob)) ? 1 : 0;
    return SZB_OK;
}


// ---- embeddings (N1) ---------------------------------------------------------------------------------------------------
szb_status szb_net_embedding_size(const szb_net* net, uint32_t* size) {
    SZB_REQUIRE(net && size, "szb_net_embedding_size: NULL argument");
    *size = net->h2;
    return SZB_OK;
}

// stages the rows (host) and leaves the [n][h2] embeddings in net->d_2-sized scratch; returns the device pointer
static szb_status embed_all(szb_net* net, const float* feats, uint64_t n, int relu2, float** d_emb) {
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->x.reserve(std::max<uint64_t>(1, n) * net->n_in * 4));
    SZB_TRY(ctx->probs.reserve(std::max<uint64_t>(1, n) * net->h2 * 4));
    if (n) SZB_CUDA(cudaMemcpyAsync(ctx->x.ptr, feats, n * net->n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (n) SZB_TRY(net_reserve_rows(net, std::min(n, kChunkRows)));
    for (uint64_t r0 = 0; r0 < n; r0 += kChunkRows) {
        const int nb = int(std::min(kChunkRows, n - r0));
        SZB_TRY(embed_rows(net, ctx->x.as<float>() + r0 * net->n_in, nb, relu2, ctx->probs.as<float>() + r0 * net->h2));
    }
    *d_emb = ctx->probs.as<float>();
    return SZB_OK;
}

szb_status szb_net_embed(szb_net* net, const float* x, uint64_t B, int32_t relu2, float* out) {
    SZB_REQUIRE(net && (B == 0 || (x && out)), "szb_net_embed: NULL argument");
    if (B == 0) return SZB_OK;
    float* d = nullptr;
    SZB_TRY(embed_all(net, x, B, relu2, &d));
    SZB_CUDA(cudaMemcpyAsync(out, d, B * net->h2 * 4, cudaMemcpyDeviceToHost, net->ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(net->ctx->stream));
    return SZB_OK;
}

static void l2_normalize(std::vector<float>& v) {   // normalize(), lib.rs:132-139
    float norm = 0.f;
    for (float x : v) norm += x * x;
    norm = std::sqrt(norm);
    if (norm > 1e-6f)
        for (float& x : v) x /= norm;
}

szb_status szb_net_embedding_mean(szb_net* net, const float* feats, uint64_t n, float* out) {
    SZB_REQUIRE(net && out && (feats || n == 0), "szb_net_embedding_mean: NULL argument");
    std::vector<float> acc(net->h2, 0.f);
    if (n > 0) {   // extract_embedding_from_features, lib.rs:1453-1475: forward_embedding (ReLU, ReLU), mean, normalise
        float* d = nullptr;
        SZB_TRY(embed_all(net, feats, n, 1, &d));
        szb_ctx* ctx = net->ctx;
        SZB_TRY(net->hist.reserve(size_t(net->h2) * 4));
        column_sum_kernel<<<(net->h2 + 31) / 32, 256, 0, ctx->stream>>>(d, uint32_t(n), int(net->h2), int(net->h2), net->hist.as<float>());
        SZB_CUDA(cudaGetLastError());
        ctx->launches += 1;
        SZB_CUDA(cudaMemcpyAsync(acc.data(), net->hist.ptr, size_t(net->h2) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SZB_CUDA(cudaStreamSynchronize(ctx->stream));
        for (float& v : acc) v /= float(n);
    }
    l2_normalize(acc);
    std::memcpy(out, acc.data(), acc.size() * 4);
    return SZB_OK;
}

szb_status szb_net_embedding_median(szb_net* net, const float* feats, uint64_t n, int32_t relu2, float* out) {
    SZB_REQUIRE(net && out && (feats || n == 0), "szb_net_embedding_median: NULL argument");
    SZB_REQUIRE(n <= 0xffffffffull, "szb_net_embedding_median: too many windows");
    std::vector<float> emb(net->h2, 0.f);
    if (n > 0) {   // median_embedding_from_features (relu2 = 1, lib.rs:1478-1500) / extract_embedding (relu2 = 0, lib.rs:1418-1450)
        float* d = nullptr;
        SZB_TRY(embed_all(net, feats, n, relu2, &d));
        szb_ctx* ctx = net->ctx;
        SZB_TRY(net->hist.reserve(size_t(net->h2) * 4));
        column_median_kernel<<<net->h2, 256, 0, ctx->stream>>>(d, uint32_t(n), int(net->h2), net->hist.as<float>());
        SZB_CUDA(cudaGetLastError());
        ctx->launches += 1;
        SZB_CUDA(cudaMemcpyAsync(emb.data(), net->hist.ptr, size_t(net->h2) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        SZB_CUDA(cudaStreamSynchronize(ctx->stream));
        l2_normalize(emb);
    }
    std::memcpy(out, emb.data(), emb.size() * 4);   // no windows: zero vector, not normalised (lib.rs:1429-1431)
    return SZB_OK;
}

float szb_cosine_similarity(const float* a, const float* b, uint32_t n) {   // lib.rs:1531-1540
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (uint32_t i = 0; i < n; ++i) { dot += a[i] * b[i]; na += a[i] * a[i]; nb += b[i] * b[i]; }
    na = std::sqrt(na); nb = std::sqrt(nb);
    return (na == 0.f || nb == 0.f) ? 0.f : dot / (na * nb);
}

// identify_speaker_from_embedding (lib.rs:1503-1529): the centroid with the largest cosine similarity wins if it clears the
// threshold, which is relaxed to 0.7 x threshold while fewer than 20 speakers are known.  The reference walks a HashMap
// (unspecified order, strict '>'): ties go to whichever entry comes first; here to the first in the caller's order.
szb_status szb_match_embedding(const float* emb, const float* centroids, const uint64_t* ids, uint32_t n, uint32_t dim, float threshold,
                               uint64_t* best_id, float* best_sim) {
    SZB_REQUIRE(best_id && (n == 0 || (emb && centroids)), "szb_match_embedding: NULL argument");
    float bs = -3.402823466e+38f;                           // f32::MIN
    uint64_t bi = UINT64_MAX;                               // usize::MAX
    for (uint32_t i = 0; i < n; ++i) {
        const float sim = szb_cosine_similarity(emb, centroids + size_t(i) * dim, dim);
        if (sim > bs) { bs = sim; bi = ids ? ids[i] : i; }
    }
    const float dynamic_threshold = n < 20 ? threshold * 0.7f : threshold;
    *best_id = bs > dynamic_threshold ? bi : UINT64_MAX;
    if (best_sim) *best_sim = bs;
    return SZB_OK;
}

// ---- aggregation ----------------------------------------------------------------------------------------------------
static szb_status identify_dev(szb_net* net, const float* d_feats, uint64_t n, float threshold, uint64_t* counts, float* sums) {
    szb_ctx* ctx = net->ctx;
    const uint32_t C = net->n_out;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(net->hist.reserve(size_t(C) * (8 + 4)));
    unsigned long long* d_hist = net->hist.as<unsigned long long>();
    float* d_sums = reinterpret_cast<float*>(d_hist + C);
    SZB_CUDA(cudaMemsetAsync(net->hist.ptr, 0, size_t(C) * 12, ctx->stream));
    if (n > 0) {
        SZB_TRY(net_reserve_rows(net, std::min(n, kChunkRows)));
        for (uint64_t r0 = 0; r0 < n; r0 += kChunkRows) {
            const int nb = int(std::min(kChunkRows, n - r0));
            SZB_TRY(forward_rows(net, d_feats + r0 * net->n_in, nb));
            SZB_TRY(launch_softmax(net, nb, counts ? 4 : 8, nullptr, nullptr, nullptr, nullptr, threshold, d_hist, d_sums));
        }
    }
    if (counts) SZB_CUDA(cudaMemcpyAsync(counts, d_hist, size_t(C) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (sums) SZB_CUDA(cudaMemcpyAsync(sums, d_sums, size_t(C) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

szb_status szb_identify_counts_dev(szb_net* net, const float* d_feats, uint64_t n, float threshold, uint64_t* counts) {
    SZB_REQUIRE(net && counts && (d_feats || n == 0), "szb_identify_counts_dev: NULL argument");
    return identify_dev(net, d_feats, n, threshold, counts, nullptr);
}

szb_status szb_identify_counts(szb_net* net, const float* feats, uint64_t n, float threshold, uint64_t* counts) {
    SZB_REQUIRE(net && counts && (feats || n == 0), "szb_identify_counts: NULL argument");
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    if (n) {
        SZB_TRY(ctx->x.reserve(n * net->n_in * 4));
        SZB_CUDA(cudaMemcpyAsync(ctx->x.ptr, feats, n * net->n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    return identify_dev(net, ctx->x.as<float>(), n, threshold, counts, nullptr);
}

szb_status szb_identify_counts_batch_dev(szb_net* net, const float* d_feats, const uint64_t* win_off, uint32_t n_clips, float threshold,
                                         uint32_t* counts) {
    SZB_REQUIRE(net && win_off && counts, "szb_identify_counts_batch_dev: NULL argument");
    if (n_clips == 0) return SZB_OK;
    for (uint32_t c = 0; c < n_clips; ++c)
        SZB_REQUIRE(win_off[c + 1] >= win_off[c], "szb_identify_counts_batch_dev: win_off not monotone at %u", c);
    const uint64_t n = win_off[n_clips] - win_off[0];
    szb_ctx* ctx = net->ctx;
    const uint32_t C = net->n_out;
    SZB_CUDA(cudaSetDevice(ctx->device));
    const size_t hist_bytes = size_t(n_clips) * C * sizeof(uint32_t), off_bytes = (size_t(n_clips) + 1) * sizeof(uint64_t);
    SZB_TRY(net->hist.reserve(hist_bytes + off_bytes + 16));
    uint32_t* d_hist = net->hist.as<uint32_t>();
    unsigned long long* d_off = reinterpret_cast<unsigned long long*>(net->hist.as<unsigned char>() + ((hist_bytes + 15) & ~size_t(15)));
    SZB_CUDA(cudaMemsetAsync(d_hist, 0, hist_bytes, ctx->stream));
    void* hp = nullptr;
    SZB_TRY(ctx->h_stage.acquire(off_bytes, &hp));
    std::memcpy(hp, win_off, off_bytes);
    SZB_CUDA(cudaMemcpyAsync(d_off, hp, off_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SZB_TRY(ctx->h_stage.uploaded(ctx->stream));
    if (n > 0) {
        SZB_REQUIRE(d_feats, "szb_identify_counts_batch_dev: d_feats is NULL");
        constexpr uint64_t kRows = 1u << 17;             // larger chunks than the per-clip path: few launches over a whole shard
        SZB_TRY(net_reserve_rows(net, std::min(n, kRows)));
        for (uint64_t r0 = 0; r0 < n; r0 += kRows) {
            const int nb = int(std::min(kRows, n - r0));
            SZB_TRY(forward_rows(net, d_feats + (win_off[0] + r0) * net->n_in, nb));
            const int wpb = 8;
            softmax_kernel<<<(nb + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(net->a_z.as<float>(), nb, int(C), int(C), 4, nullptr, nullptr, nullptr,
                                                                              nullptr, nullptr, threshold, nullptr, nullptr, nullptr, nb, d_off, n_clips,
                                                                              win_off[0] + r0, d_hist);
            SZB_CUDA(cudaGetLastError());
            ctx->launches += 1;
        }
    }
    SZB_CUDA(cudaMemcpyAsync(counts, d_hist, hist_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SZB_CUDA(cudaStreamSynchronize(ctx->stream));
    return SZB_OK;
}

szb_status szb_identify_sums(szb_net* net, const float* feats, uint64_t n, float* sums) {
    SZB_REQUIRE(net && sums && (feats || n == 0), "szb_identify_sums: NULL argument");
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    if (n) {
        SZB_TRY(ctx->x.reserve(n * net->n_in * 4));
        SZB_CUDA(cudaMemcpyAsync(ctx->x.ptr, feats, n * net->n_in * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    return identify_dev(net, ctx->x.as<float>(), n, 0.f, nullptr, sums);
}

szb_status szb_identify_speaker_list(szb_net* net, const int16_t* pcm, uint64_t n_samples, float threshold, uint32_t* speakers,
                                     uint32_t cap, uint32_t* n_speakers) {
    SZB_REQUIRE(net && n_speakers, "szb_identify_speaker_list: NULL argument");
    SZB_REQUIRE(net->n_in == SZB_FEATURE_SIZE, "szb_identify_speaker_list: net input size %u != 60", net->n_in);
    szb_ctx* ctx = net->ctx;
    *n_speakers = 0;
    const uint64_t n = szb_num_windows(n_samples);
    std::vector<uint64_t> counts(net->n_out, 0);
    if (n > 0) {
        SZB_REQUIRE(pcm, "szb_identify_speaker_list: pcm is NULL");
        SZB_CUDA(cudaSetDevice(ctx->device));
        // extraction output stays on the device and feeds the forward pass directly (lib.rs:1390-1392)
        SZB_TRY(ctx->pcm.reserve(n_samples * 2 + 64));
        SZB_TRY(ctx->feats.reserve(n * SZB_FEATURE_SIZE * 4));
        SZB_CUDA(cudaMemcpyAsync(ctx->pcm.ptr, pcm, n_samples * 2, cudaMemcpyHostToDevice, ctx->stream));
        const uint64_t off[2] = { 0, n_samples };
        uint64_t woff[2];
        SZB_TRY(szb_extract_batch_dev(ctx, ctx->pcm.as<int16_t>(), off, 1, SZB_SAMPLE_RATE, ctx->feats.as<float>(), n, woff));
        SZB_TRY(identify_dev(net, ctx->feats.as<float>(), n, threshold, counts.data(), nullptr));
    }
    // lib.rs:1403-1410: keep count > 0, stable sort by count descending
    std::vector<std::pair<uint32_t, uint64_t>> pairs;
    for (uint32_t c = 0; c < net->n_out; ++c)
        if (counts[c] > 0) pairs.emplace_back(c, counts[c]);
    std::stable_sort(pairs.begin(), pairs.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
    *n_speakers = uint32_t(pairs.size());
    SZB_REQUIRE(cap >= pairs.size() && (speakers || pairs.empty()), "szb_identify_speaker_list: capacity %u < %zu", cap, pairs.size());
    for (size_t i = 0; i < pairs.size(); ++i) speakers[i] = pairs[i].first;
    return SZB_OK;
}

}  // extern "C"
