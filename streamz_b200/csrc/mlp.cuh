// Internal interface of the SimpleNeuralNet kernels (mlp.cu).
#pragma once
#include "common.cuh"

namespace szb {
// Per-epoch parameters of a CAPTURED training step (CUDA graph): they live in device memory so that one instantiated graph
// serves every step of every epoch -- the batch kernel reads its position in the shuffled order and the dropout key from here,
// the update kernel the learning rate, and it advances the position for the next step.
struct StepParams {
    uint32_t cursor;              // first row of the step in the epoch's permutation
    float lr;
    unsigned long long key;       // dropout key of the epoch (seed, stream)
    uint32_t p2p_step;            // multi-GPU: gradient exchanges done so far; the batch kernel of a step counts it up, the update
    uint32_t pad;                 // kernel uses it as the step's flag value (szb_ctx::p2p_step follows on the host)
};
struct StepGraphKey {             // everything a captured step bakes in
    int B = 0, precision = -1, parity = -1;
    const void *feats = nullptr, *labels = nullptr, *keep = nullptr, *perm = nullptr, *params = nullptr;
    float prob = -1.f;
    uint64_t cap_rows = 0;
    uint32_t n_out = 0;
    int exchange = 0;             // 0 = single GPU, 1 + szb_ctx::p2p_mode otherwise
    bool operator==(const StepGraphKey& o) const {
        return B == o.B && precision == o.precision && parity == o.parity && feats == o.feats && labels == o.labels && keep == o.keep &&
               perm == o.perm && params == o.params && prob == o.prob && cap_rows == o.cap_rows && n_out == o.n_out && exchange == o.exchange;
    }
};
}  // namespace szb

struct szb_net {
    szb_ctx* ctx = nullptr;
    uint32_t n_in = 0, h1 = 0, h2 = 0, n_out = 0;
    // flattened parameters  [w1 | b1 | w2 | b2 | w3 | b3]  and gradients in the same order followed by
    // [n_used, loss_sum_lo, ...] so that ONE all-reduce covers a whole step.
    szb::DevBuf params, grads;
    // activations / scratch for up to cap_rows rows
    szb::DevBuf xb, lab, valid, a_h1, a_h2, a_z, d_2, d_1, stats, perm, hist;
    // tensor-core path: transposed weights [wt1 (h1 x n_in) | wt2 (h2 x h1) | wt3 (n_out x h2)] and the transposed
    // activations / deltas ([features][rows]) that make every GEMM of a step "TN" (gemm_tc.cuh)
    szb::DevBuf wt, xbT, h1T, h2T, zT, d2T, d1T;
    szb::DevBuf chunk_xb, chunk_xbT, chunk_lab, chunk_valid;   // batch buffers of several steps filled by one launch (large batches)
    bool wt_dirty = true;
    bool grads_zero = false;      // the gradient vector (except the other tail block) is known to be all zeros
    int tail_parity = 0;          // which tail block the current step accumulates [n_used, loss] into
    int precision = 1;            // 0 = FP32 SIMT, 1 = 3xTF32 tensor cores (default), 2 = TF32 tensor cores
    uint64_t cap_rows = 0, cap_rows_t = 0;
    // small-batch epochs (launch-latency-bound: 11 launches per step): two consecutive steps captured once as a CUDA graph
    // and replayed (mlp.cu: szb_net_train_epoch_steps_dev)
    // batches of <= 32 rows: the whole epoch as one persistent cooperative kernel (train_small.cu)
    szb::DevBuf small_scratch, small_barrier, small_steps;
    int small_grid = 0;
    bool small_failed = false;
    szb::DevBuf step_params;
    cudaGraphExec_t step_graph = nullptr;
    uint32_t step_graph_kernels = 0;   // kernel launches inside one replay of step_graph (two steps)
    szb::StepGraphKey step_graph_key;
    bool step_graph_failed = false;
    std::vector<std::vector<std::string>> file_lists;  // lib.rs:757, host-side only
    // Saved speaker embeddings with their quality metrics (lib.rs:760-761, set_embeddings / embeddings lib.rs:870-877) and
    // the optional hidden encoding layer (lib.rs:752-754).  Host-side state that model.npz carries (lib.rs:1099-1127,
    // 1168-1264); the hot path never reads the encoding layer, it is kept only so that load -> save loses nothing.
    std::vector<float> emb, emb_mean, emb_std;         // [emb_n][emb_dim], [emb_n], [emb_n]
    uint32_t emb_n = 0, emb_dim = 0;
    std::vector<float> w4, b4;                         // [w4_rows][b4.size()] row-major, [n]
    uint32_t w4_rows = 0;

    size_t off_w1() const { return 0; }
    size_t off_b1() const { return size_t(n_in) * h1; }
    size_t off_w2() const { return off_b1() + h1; }
    size_t off_b2() const { return off_w2() + size_t(h1) * h2; }
    size_t off_w3() const { return off_b2() + h2; }
    size_t off_b3() const { return off_w3() + size_t(h2) * n_out; }
    size_t n_params() const { return off_b3() + n_out; }
    size_t off_wt1() const { return 0; }
    size_t off_wt2() const { return size_t(h1) * n_in; }
    size_t off_wt3() const { return off_wt2() + size_t(h2) * h1; }
    size_t n_wt() const { return off_wt3() + size_t(n_out) * h2; }
};

namespace szb {
// floats appended to the gradient vector: two [n_used, loss, spare, spare] blocks.  Steps alternate between them (tail
// parity): the fused update kernel of a step reads its own block and zeroes the other one for the next step, so no block is
// ever zeroed while another CTA may still be reading it.
constexpr size_t kGradTail = 8;
szb_status net_reserve_rows(szb_net* net, uint64_t rows);
szb_status train_epoch_small(szb_net* net, const float* d_feats, const uint32_t* d_labels, const uint32_t* d_perm, const uint32_t* step_sizes,
                             uint32_t n_steps, float lr, float dropout, unsigned long long key, const uint8_t* d_keep, bool* done);
}  // namespace szb
