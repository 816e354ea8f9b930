// Raw-audio training loops of the reference on top of the device-resident entry points (host code, no kernels of its own):
//   pretrain_network  (streamz-rs/src/lib.rs:348-397): per epoch  augment -> extract -> shuffle -> dropout / train_batch chunks
//   train_from_files  (lib.rs:668-732):               per file, per epoch  lr * 0.99^step  and one pretrain_network epoch
// The augmented clip, its windows and the labels never leave the GPU; only the shuffled order (n_windows u32) is uploaded
// per epoch.  Every random draw the reference takes from an unseeded thread_rng derives from `seed` here:
//   augmentation of epoch e of file f :  szb_augment_params(szb_loop_seed(seed, f, e)) + the per-sample counter stream
//   shuffle                           :  szb_shuffle_perm(szb_loop_seed(seed, f, e), 0, n)          (lib.rs:370)
//   input dropout                     :  counter RNG keyed (szb_loop_seed(seed, f, e), stream 0)    (lib.rs:375)
// so the oracle (oracle/streamz_oracle.py: pretrain_network, train_from_files) restates the same run draw for draw.
#include <cstring>
#include <vector>

#include "common.cuh"
#include "mlp.cuh"

namespace szb {

static unsigned long long loop_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// f32::powi as compiler-rt's __powisf2 evaluates it (binary exponentiation in float32): lr * 0.99f32.powi(step), lib.rs:709.
static float powi_f32(float a, int b) {
    const bool recip = b < 0;
    float r = 1.f;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.f / r : r;
}

// One epoch of pretrain_network on a clip that is already on the device.  d_pcm: the clip; d_aug: scratch of the same length.
static szb_status pretrain_epoch_dev(szb_net* net, const int16_t* d_pcm, int16_t* d_aug, uint64_t n_samples, uint32_t target_class,
                                     float lr, float dropout, uint32_t batch, uint64_t epoch_seed, double* loss_sum, uint64_t* n_used) {
    szb_ctx* ctx = net->ctx;
    const uint64_t n = szb_num_windows(n_samples);
    *loss_sum = 0.0;
    *n_used = 0;
    if (n == 0) return SZB_OK;                                                  // lib.rs:369-371: no windows, no chunks
    SZB_TRY(szb_augment_dev(ctx, d_pcm, n_samples, epoch_seed, d_aug));         // lib.rs:368
    SZB_TRY(ctx->feats.reserve(n * SZB_FEATURE_SIZE * sizeof(float)));
    const uint64_t off[2] = { 0, n_samples };
    uint64_t woff[2];
    SZB_TRY(szb_extract_batch_dev(ctx, d_aug, off, 1, SZB_SAMPLE_RATE, ctx->feats.as<float>(), n, woff));   // lib.rs:369
    std::vector<uint32_t> perm(n);
    SZB_TRY(szb_shuffle_perm(epoch_seed, 0, n, perm.data()));                   // lib.rs:370
    if (ctx->loop_labels_n < n || ctx->loop_labels_class != target_class) {     // every window carries the file's label
        std::vector<uint32_t> lab(n, target_class);
        SZB_TRY(ctx->loop_labels.reserve(n * sizeof(uint32_t)));
        SZB_CUDA(cudaMemcpyAsync(ctx->loop_labels.ptr, lab.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        SZB_CUDA(cudaStreamSynchronize(ctx->stream));                           // lab is a temporary
        ctx->loop_labels_n = n;
        ctx->loop_labels_class = target_class;
    }
    return szb_net_train_epoch_dev(net, ctx->feats.as<float>(), ctx->loop_labels.as<uint32_t>(), n, perm.data(), n, batch, lr, dropout,
                                   epoch_seed, 0, nullptr, loss_sum, n_used);   // lib.rs:371-390
}

}  // namespace szb

using namespace szb;

extern "C" {

uint64_t szb_loop_seed(uint64_t seed, uint32_t file, uint32_t epoch) {
    return loop_splitmix64(loop_splitmix64(seed ^ 0x7EA1F11E5ull) + (uint64_t(file) << 32 | epoch));
}

// Fisher-Yates from the back (the order rand's SliceRandom::shuffle walks, lib.rs:370/601) with this repo's counter stream:
// for i = n-1 .. 1: j = splitmix64(key ^ i) % (i + 1); swap(p[i], p[j]);  key = splitmix64(seed + stream * 0xD1B54A32D192ED03 + 0x5F)
szb_status szb_shuffle_perm(uint64_t seed, uint64_t stream, uint64_t n, uint32_t* perm) {
    SZB_REQUIRE(perm || n == 0, "szb_shuffle_perm: perm is NULL");
    SZB_REQUIRE(n <= 0xffffffffull, "szb_shuffle_perm: more than 2^32 rows");
    for (uint64_t i = 0; i < n; ++i) perm[i] = uint32_t(i);
    const unsigned long long key = loop_splitmix64(seed + stream * 0xD1B54A32D192ED03ull + 0x5Full);
    for (uint64_t i = n; i-- > 1;) {
        const uint64_t j = loop_splitmix64(key ^ i) % (i + 1);
        const uint32_t t = perm[i];
        perm[i] = perm[j];
        perm[j] = t;
    }
    return SZB_OK;
}

float szb_lr_decay(float lr, int32_t step) { return lr * powi_f32(0.99f, step); }   // lib.rs:709

szb_status szb_net_pretrain_network(szb_net* net, const int16_t* pcm, uint64_t n_samples, uint32_t target_class, uint32_t epochs,
                                    float lr, float dropout, uint32_t batch, uint64_t seed, double* loss_sum, uint64_t* n_used) {
    SZB_REQUIRE(net, "szb_net_pretrain_network: net is NULL");
    SZB_REQUIRE(net->n_in == SZB_FEATURE_SIZE, "szb_net_pretrain_network: net input size %u != 60", net->n_in);
    if (loss_sum) *loss_sum = 0.0;
    if (n_used) *n_used = 0;
    if (epochs == 0 || szb_num_windows(n_samples) == 0) return SZB_OK;           // lib.rs:392-396: count == 0 -> 0.0
    SZB_REQUIRE(pcm, "szb_net_pretrain_network: pcm is NULL");
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    SZB_TRY(ctx->loop_pcm.reserve(2 * (n_samples * 2 + 64)));
    int16_t* d_pcm = ctx->loop_pcm.as<int16_t>();
    int16_t* d_aug = d_pcm + ((n_samples + 31) & ~uint64_t(31));                  // keeps the 16-byte alignment of the clip start
    SZB_CUDA(cudaMemcpyAsync(d_pcm, pcm, n_samples * 2, cudaMemcpyHostToDevice, ctx->stream));
    double total = 0.0;
    uint64_t count = 0;
    for (uint32_t e = 0; e < epochs; ++e) {
        double l = 0.0;
        uint64_t c = 0;
        SZB_TRY(pretrain_epoch_dev(net, d_pcm, d_aug, n_samples, target_class, lr, dropout, batch, szb_loop_seed(seed, 0, e), &l, &c));
        total += l;
        count += c;
    }
    if (loss_sum) *loss_sum = total;
    if (n_used) *n_used = count;
    return SZB_OK;
}

szb_status szb_net_train_from_files(szb_net* net, const int16_t* pcm, const uint64_t* clip_off, const uint32_t* classes, uint32_t n_files,
                                    uint32_t epochs, float lr, float dropout, uint32_t batch, uint64_t seed, double* loss_sum,
                                    uint64_t* n_used) {
    SZB_REQUIRE(net, "szb_net_train_from_files: net is NULL");
    SZB_REQUIRE(net->n_in == SZB_FEATURE_SIZE, "szb_net_train_from_files: net input size %u != 60", net->n_in);
    if (loss_sum) *loss_sum = 0.0;
    if (n_used) *n_used = 0;
    if (n_files == 0 || epochs == 0) return SZB_OK;
    SZB_REQUIRE(pcm && clip_off && classes, "szb_net_train_from_files: NULL argument");
    szb_ctx* ctx = net->ctx;
    SZB_CUDA(cudaSetDevice(ctx->device));
    uint64_t max_len = 0;
    for (uint32_t f = 0; f < n_files; ++f) {
        SZB_REQUIRE(clip_off[f + 1] >= clip_off[f], "szb_net_train_from_files: clip_off not monotone at %u", f);
        max_len = std::max(max_len, clip_off[f + 1] - clip_off[f]);
    }
    // every clip starts on a 16-byte boundary of the device buffer so the extraction kernel keeps its 128-bit loads
    std::vector<uint64_t> d_off(size_t(n_files) + 1, 0);
    for (uint32_t f = 0; f < n_files; ++f) d_off[f + 1] = (d_off[f] + (clip_off[f + 1] - clip_off[f]) + 31) & ~uint64_t(31);
    const uint64_t aug_off = d_off[n_files];
    SZB_TRY(ctx->loop_pcm.reserve((aug_off + max_len + 64) * 2));
    int16_t* base = ctx->loop_pcm.as<int16_t>();
    for (uint32_t f = 0; f < n_files; ++f) {
        const uint64_t len = clip_off[f + 1] - clip_off[f];
        if (len) SZB_CUDA(cudaMemcpyAsync(base + d_off[f], pcm + clip_off[f], len * 2, cudaMemcpyHostToDevice, ctx->stream));
    }
    double total = 0.0;
    uint64_t count = 0;
    int32_t step = 0;
    // The reference walks the files with rayon and serialises the epochs of all files under one write lock (lib.rs:691,
    // 710), so their interleaving is scheduler-dependent; file-major order is the single-thread serialisation of it.
    for (uint32_t f = 0; f < n_files; ++f) {
        const uint64_t len = clip_off[f + 1] - clip_off[f];
        for (uint32_t e = 0; e < epochs; ++e) {
            const float lr_scaled = szb_lr_decay(lr, step++);                    // lib.rs:709 (the step counts even for clips without windows)
            double l = 0.0;
            uint64_t c = 0;
            SZB_TRY(pretrain_epoch_dev(net, base + d_off[f], base + aug_off, len, classes[f], lr_scaled, dropout, batch,
                                       szb_loop_seed(seed, f, e), &l, &c));
            total += l;
            count += c;
        }
    }
    if (loss_sum) *loss_sum = total;
    if (n_used) *n_used = count;
    return SZB_OK;
}

}  // extern "C"
