// Internal interface of the SimpleNeuralNet kernels (mlp.cu).
#pragma once
#include "common.cuh"

struct szb_net {
    szb_ctx* ctx = nullptr;
    uint32_t n_in = 0, h1 = 0, h2 = 0, n_out = 0;
    // flattened parameters  [w1 | b1 | w2 | b2 | w3 | b3]  and gradients in the same order followed by
    // [n_used, loss_sum_lo, ...] so that ONE all-reduce covers a whole step.
    szb::DevBuf params, grads;
    // activations / scratch for up to cap_rows rows
    szb::DevBuf xb, lab, valid, a_h1, a_h2, a_z, d_2, d_1, stats, perm, hist;
    uint64_t cap_rows = 0;
    std::vector<std::vector<std::string>> file_lists;  // lib.rs:757, host-side only

    size_t off_w1() const { return 0; }
    size_t off_b1() const { return size_t(n_in) * h1; }
    size_t off_w2() const { return off_b1() + h1; }
    size_t off_b2() const { return off_w2() + size_t(h1) * h2; }
    size_t off_w3() const { return off_b2() + h2; }
    size_t off_b3() const { return off_w3() + size_t(h2) * n_out; }
    size_t n_params() const { return off_b3() + n_out; }
};

namespace szb {
constexpr size_t kGradTail = 4;  // floats appended to the gradient vector: [n_used, loss, spare, spare]
szb_status net_reserve_rows(szb_net* net, uint64_t rows);
}  // namespace szb
