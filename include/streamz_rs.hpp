// streamz_rs.hpp -- C++17 host-side mirror of the reference's public Rust API for the hot path, over the C ABI of
// streamz_b200.h.  The reference (crate `streamz_rs`, streamz-rs/src/lib.rs) is compiled Rust; there is no Rust toolchain
// in this build environment, so the host layer a Rust maintainer would write (rust/src/lib.rs in this repo, uncompiled)
// is provided here in C++ with the same names, argument meaning and error behaviour:
//
//   Rust (lib.rs)                                                   C++ (namespace streamz_rs)
//   FeatureExtractor::new / extract                 239, 261        FeatureExtractor{} / extract()
//   with_thread_extractor                           271             with_thread_extractor(f)
//   downmix_to_mono / resample_to_44100             172, 186        same names; resample throws on error (Result<_, Box<dyn Error>>)
//   SimpleNeuralNet::{new, forward, train, train_batch, output_size, add_output_class, record_training_file,
//                     file_lists, save, load}       767-1282        same names; save/load throw (Result)
//   pretrain_from_features / train_from_feature_map 582, 632        same names, same argument order
//   identify_speaker / _with_threshold / _with_threshold_feats / identify_speaker_list   1285-1411   same names
//   load_cached_features                            558             same name (caller supplies the decoded samples)
//
// Infallible Rust functions stay infallible in spirit: an empty input gives an empty output; a CUDA failure (which has
// no analogue in the reference) throws streamz_rs::Error.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <optional>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <tuple>
#include <vector>

#include "streamz_b200.h"

namespace streamz_rs {

constexpr uint32_t DEFAULT_SAMPLE_RATE = SZB_SAMPLE_RATE;  // lib.rs:25
constexpr size_t WINDOW_SIZE = SZB_WINDOW_SIZE;            // lib.rs:26
constexpr size_t MFCC_SIZE = SZB_MFCC_SIZE;                // lib.rs:28
constexpr size_t FEATURE_SIZE = SZB_FEATURE_SIZE;          // lib.rs:30-34
constexpr float DEFAULT_DROPOUT = SZB_DEFAULT_DROPOUT;     // lib.rs:36

struct Error : std::runtime_error {
    szb_status status;
    Error(szb_status s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(szb_status s) {
    if (s != SZB_OK) throw Error(s, szb_last_error());
}

using Windows = std::vector<std::vector<float>>;  // Vec<Vec<f32>>

namespace detail {
inline std::vector<float> flatten(const Windows& w, size_t width) {
    std::vector<float> flat;
    flat.reserve(w.size() * width);
    for (const auto& row : w) {
        if (row.size() != width) throw Error(SZB_ERR_INVALID, "window has the wrong number of features");  // ndarray would panic
        flat.insert(flat.end(), row.begin(), row.end());
    }
    return flat;
}
inline Windows unflatten(const std::vector<float>& flat, size_t rows, size_t width) {
    Windows w(rows);
    for (size_t r = 0; r < rows; ++r) w[r].assign(flat.begin() + r * width, flat.begin() + (r + 1) * width);
    return w;
}
}  // namespace detail

// One CUDA context (stream + scratch) per thread, like the reference's thread-local extractor (lib.rs:266-268).
class Context {
   public:
    explicit Context(int device = 0) { check(szb_ctx_create(device, nullptr, &ctx_)); }
    ~Context() { szb_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    szb_ctx* get() const { return ctx_; }
    static Context& thread_default() {
        thread_local Context c(0);
        return c;
    }

   private:
    szb_ctx* ctx_ = nullptr;
};

inline std::vector<int16_t> downmix_to_mono(const std::vector<int16_t>& samples, size_t channels) {   // lib.rs:172
    std::vector<int16_t> out((samples.size() + std::max<size_t>(channels, 1) - 1) / std::max<size_t>(channels, 1));
    uint64_t n = 0;
    check(szb_downmix_to_mono(Context::thread_default().get(), samples.data(), samples.size(), uint32_t(channels), out.data(),
                              out.size(), &n));
    out.resize(n);
    return out;
}

inline std::vector<int16_t> resample_to_44100(const std::vector<int16_t>& samples, uint32_t from_rate) {   // lib.rs:186
    std::vector<int16_t> out(from_rate == DEFAULT_SAMPLE_RATE ? samples.size() : szb_resample_out_len(samples.size(), from_rate));
    uint64_t n = 0;
    check(szb_resample_to_44100(Context::thread_default().get(), samples.data(), samples.size(), from_rate, out.data(), out.size(), &n));
    out.resize(n);
    return out;
}

class FeatureExtractor {   // lib.rs:231-264
   public:
    FeatureExtractor() : ctx_(&Context::thread_default()) {}   // tables are uploaded when the context is created
    explicit FeatureExtractor(Context& ctx) : ctx_(&ctx) {}
    Windows extract(const std::vector<int16_t>& samples) const {   // lib.rs:261
        const uint64_t n = szb_num_windows(samples.size());
        std::vector<float> flat(n * FEATURE_SIZE);
        uint64_t got = 0;
        check(szb_extract(ctx_->get(), samples.data(), samples.size(), flat.data(), n, &got));
        return detail::unflatten(flat, got, FEATURE_SIZE);
    }
    // The rayon loop of main.rs:500-508 (clips at `rate` Hz are resampled first, like batch_resample, lib.rs:541).
    std::vector<Windows> extract_batch(const std::vector<std::vector<int16_t>>& clips, uint32_t rate = DEFAULT_SAMPLE_RATE) const {
        std::vector<uint64_t> off(clips.size() + 1, 0), woff(clips.size() + 1, 0);
        std::vector<int16_t> pcm;
        for (size_t i = 0; i < clips.size(); ++i) {
            pcm.insert(pcm.end(), clips[i].begin(), clips[i].end());
            off[i + 1] = pcm.size();
        }
        const uint64_t total = szb_extract_batch_windows(off.data(), uint32_t(clips.size()), rate);
        std::vector<float> flat(total * FEATURE_SIZE);
        check(szb_extract_batch(ctx_->get(), pcm.data(), off.data(), uint32_t(clips.size()), rate, flat.data(), total, woff.data()));
        std::vector<Windows> out(clips.size());
        for (size_t i = 0; i < clips.size(); ++i) {
            const size_t n = size_t(woff[i + 1] - woff[i]);
            std::vector<float> part(flat.begin() + woff[i] * FEATURE_SIZE, flat.begin() + woff[i + 1] * FEATURE_SIZE);
            out[i] = detail::unflatten(part, n, FEATURE_SIZE);
        }
        return out;
    }
    Context& context() const { return *ctx_; }

   private:
    Context* ctx_;
};

template <class F>
auto with_thread_extractor(F&& f) {   // lib.rs:271
    thread_local FeatureExtractor ex;
    return f(ex);
}

class SimpleNeuralNet {   // lib.rs:745-1282
   public:
    SimpleNeuralNet(size_t input, size_t hidden1, size_t hidden2, size_t output, uint64_t seed = std::random_device{}())   // ::new, lib.rs:767
        : ctx_(&Context::thread_default()) {
        check(szb_net_create(ctx_->get(), uint32_t(input), uint32_t(hidden1), uint32_t(hidden2), uint32_t(output), seed, &net_));
    }
    ~SimpleNeuralNet() { szb_net_destroy(net_); }
    SimpleNeuralNet(SimpleNeuralNet&& o) noexcept : ctx_(o.ctx_), net_(o.net_), sample_rate_(o.sample_rate_), bits_(o.bits_) { o.net_ = nullptr; }
    SimpleNeuralNet(const SimpleNeuralNet&) = delete;

    size_t output_size() const { return szb_net_output_size(net_); }   // lib.rs:792
    size_t input_size() const {
        uint32_t d[4];
        check(szb_net_dims(net_, d));
        return d[0];
    }
    void add_output_class() { check(szb_net_add_output_class(net_, nullptr, std::random_device{}())); }   // lib.rs:797
    void set_dataset_specs(uint32_t sample_rate, uint16_t bits) { sample_rate_ = sample_rate; bits_ = bits; }   // lib.rs:823
    void record_training_file(size_t cls, const std::string& path) { check(szb_net_record_training_file(net_, uint32_t(cls), path.c_str())); }

    std::vector<float> forward(const std::vector<float>& bits) const {   // lib.rs:880
        std::vector<float> p(output_size());
        if (bits.size() != input_size()) throw Error(SZB_ERR_INVALID, "forward: input size mismatch");
        check(szb_net_forward(net_, bits.data(), 1, p.data()));
        return p;
    }
    Windows forward_batch(const Windows& rows) const {   // the batched call the GPU wants; same arithmetic per row
        const std::vector<float> flat = detail::flatten(rows, input_size());
        std::vector<float> p(rows.size() * output_size());
        check(szb_net_forward(net_, flat.data(), rows.size(), p.data()));
        return detail::unflatten(p, rows.size(), output_size());
    }
    void train(const std::vector<float>& bits, const std::vector<float>& target, float lr) { train_batch({ bits }, target, lr); }   // lib.rs:954
    void train_batch(const Windows& batch, const std::vector<float>& target, float lr) {   // lib.rs:1002
        if (batch.empty()) return;   // lib.rs:1003-1005
        if (target.size() != output_size()) throw Error(SZB_ERR_INVALID, "train_batch: target size mismatch");
        const std::vector<float> flat = detail::flatten(batch, input_size());
        check(szb_net_train_batch(net_, flat.data(), batch.size(), target.data(), lr));
    }
    // set_embeddings / embeddings (lib.rs:869-877): (embedding, mean similarity, std similarity) per speaker, saved with the model
    using SpeakerEmbedding = std::tuple<std::vector<float>, float, float>;
    void set_embeddings(const std::vector<SpeakerEmbedding>& embeds) {
        const size_t dim = embeds.empty() ? 0 : std::get<0>(embeds[0]).size();
        std::vector<float> flat, mean, sd;
        for (const auto& e : embeds) {
            flat.insert(flat.end(), std::get<0>(e).begin(), std::get<0>(e).end());
            mean.push_back(std::get<1>(e));
            sd.push_back(std::get<2>(e));
        }
        check(szb_net_set_embeddings(net_, flat.data(), mean.data(), sd.data(), uint32_t(embeds.size()), uint32_t(dim)));
    }
    std::vector<SpeakerEmbedding> embeddings() const {
        uint32_t n = 0, dim = 0;
        check(szb_net_get_embeddings(net_, nullptr, nullptr, nullptr, 0, &n, &dim));
        std::vector<float> flat(size_t(n) * dim), mean(n), sd(n);
        if (n) check(szb_net_get_embeddings(net_, flat.data(), mean.data(), sd.data(), n, &n, &dim));
        std::vector<SpeakerEmbedding> out;
        for (uint32_t i = 0; i < n; ++i)
            out.emplace_back(std::vector<float>(flat.begin() + size_t(i) * dim, flat.begin() + size_t(i + 1) * dim), mean[i], sd[i]);
        return out;
    }
    void save(const std::string& path) const { check(szb_net_save(net_, path.c_str(), sample_rate_, bits_)); }   // lib.rs:1081
    static SimpleNeuralNet load(const std::string& path) {   // lib.rs:1132
        SimpleNeuralNet n;
        uint32_t sr = 0, bits = 0;
        check(szb_net_load(n.ctx_->get(), path.c_str(), &n.net_, &sr, &bits));
        n.sample_rate_ = sr;
        n.bits_ = uint16_t(bits);
        return n;
    }
    szb_net* handle() const { return net_; }
    Context& context() const { return *ctx_; }

   private:
    SimpleNeuralNet() : ctx_(&Context::thread_default()) {}
    Context* ctx_;
    szb_net* net_ = nullptr;
    uint32_t sample_rate_ = DEFAULT_SAMPLE_RATE;
    uint16_t bits_ = 16;
};

// lib.rs:582-628.  `rng` supplies the shuffle (the reference uses an unseeded thread_rng); dropout decisions come from the
// library's counter RNG keyed by (seed, epoch).
inline float pretrain_from_features(SimpleNeuralNet& net, const Windows& windows, size_t target_class, size_t num_classes, size_t epochs,
                                    float lr, float dropout, size_t batch_size, std::mt19937_64* rng = nullptr, uint64_t seed = 0) {
    if (num_classes != net.output_size()) throw Error(SZB_ERR_INVALID, "pretrain_from_features: num_classes != output_size");
    if (windows.empty() || epochs == 0) return 0.0f;
    std::mt19937_64 local(seed);
    if (!rng) rng = &local;
    const size_t n = windows.size();
    const std::vector<float> flat = detail::flatten(windows, net.input_size());
    std::vector<uint32_t> labels(n, uint32_t(target_class)), perm(n);
    szb_ctx* ctx = net.context().get();
    void *d_feats = nullptr, *d_labels = nullptr;
    check(szb_dev_alloc(ctx, flat.size() * 4, &d_feats));
    check(szb_dev_alloc(ctx, n * 4, &d_labels));
    double total = 0.0;
    uint64_t count = 0;
    try {
        check(szb_memcpy_h2d(ctx, d_feats, flat.data(), flat.size() * 4));
        check(szb_memcpy_h2d(ctx, d_labels, labels.data(), n * 4));
        for (size_t e = 0; e < epochs; ++e) {
            for (size_t i = 0; i < n; ++i) perm[i] = uint32_t(i);
            std::shuffle(perm.begin(), perm.end(), *rng);   // lib.rs:600-601
            double loss = 0.0;
            uint64_t used = 0;
            check(szb_net_train_epoch_dev(net.handle(), static_cast<const float*>(d_feats), static_cast<const uint32_t*>(d_labels), n,
                                          perm.data(), n, uint32_t(std::max<size_t>(batch_size, 1)), lr, dropout, seed, e, nullptr, &loss,
                                          &used));
            total += loss;
            count += used;
        }
    } catch (...) {
        szb_dev_free(ctx, d_feats);
        szb_dev_free(ctx, d_labels);
        throw;
    }
    szb_dev_free(ctx, d_feats);
    szb_dev_free(ctx, d_labels);
    return count ? float(total / double(count)) : 0.0f;   // lib.rs:623-627
}

// lib.rs:348-397: `epochs` times augment -> extract -> shuffle -> dropout / train_batch chunks; one device-resident C call.
// Every draw the reference takes from thread_rng (augmentation, shuffle, dropout) derives from `seed` (szb_loop_seed).
inline float pretrain_network(SimpleNeuralNet& net, const std::vector<int16_t>& samples, size_t target_class, size_t num_classes,
                              size_t epochs, float lr, float dropout, size_t batch_size, uint64_t seed = 0) {
    if (num_classes != net.output_size()) throw Error(SZB_ERR_INVALID, "pretrain_network: num_classes != output_size");
    double loss = 0.0;
    uint64_t used = 0;
    check(szb_net_pretrain_network(net.handle(), samples.data(), samples.size(), uint32_t(target_class), uint32_t(epochs), lr, dropout,
                                   uint32_t(batch_size ? batch_size : 1), seed, &loss, &used));
    return used ? float(loss / double(used)) : 0.0f;   // lib.rs:392-396
}

// lib.rs:668-732 on decoded 44.1 kHz clips (decoding stays with the caller, lib.rs:696): every (file, epoch) is one
// pretrain_network epoch at lr * 0.99^step (lib.rs:708-709), file-major; records the training files (lib.rs:723).
struct TrainingClip { std::string path; std::vector<int16_t> samples; size_t cls; };
inline float train_from_files(SimpleNeuralNet& net, const std::vector<TrainingClip>& files, size_t num_speakers, size_t epochs, float lr,
                              float dropout, size_t batch_size, uint64_t seed = 0) {
    if (num_speakers != net.output_size()) throw Error(SZB_ERR_INVALID, "train_from_files: num_speakers != output_size");
    net.set_dataset_specs(DEFAULT_SAMPLE_RATE, 16);   // lib.rs:703-706
    std::vector<int16_t> pcm;
    std::vector<uint64_t> off{ 0 };
    std::vector<uint32_t> classes;
    for (const auto& f : files) {
        pcm.insert(pcm.end(), f.samples.begin(), f.samples.end());
        off.push_back(pcm.size());
        classes.push_back(uint32_t(f.cls));
    }
    double loss = 0.0;
    uint64_t used = 0;
    check(szb_net_train_from_files(net.handle(), pcm.data(), off.data(), classes.data(), uint32_t(files.size()), uint32_t(epochs), lr, dropout,
                                   uint32_t(batch_size ? batch_size : 1), seed, &loss, &used));
    for (const auto& f : files) net.record_training_file(f.cls, f.path);
    return used ? float(loss / double(used)) : 0.0f;
}

// lib.rs:632-665: files one after the other, each for all its epochs; mean of the per-file losses.
inline float train_from_feature_map(SimpleNeuralNet& net, const std::map<std::string, Windows>& feature_map,
                                    const std::vector<std::pair<std::string, size_t>>& files, size_t epochs, float lr, float dropout,
                                    size_t batch_size, uint64_t seed = 0) {
    float total = 0.f;
    size_t count = 0;
    for (const auto& [path, cls] : files) {
        auto it = feature_map.find(path);
        if (it == feature_map.end()) continue;
        total += pretrain_from_features(net, it->second, cls, net.output_size(), epochs, lr, dropout, batch_size, nullptr, seed + count);
        net.record_training_file(cls, path);
        ++count;
    }
    return count ? total / float(count) : 0.f;
}

namespace detail {
inline size_t argmax_last(const std::vector<float>& v) {   // max_by keeps the last maximal element (lib.rs:1298-1301)
    size_t best = 0;
    for (size_t i = 1; i < v.size(); ++i)
        if (v[i] >= v[best]) best = i;
    return best;
}
inline std::vector<float> prob_sums(const SimpleNeuralNet& net, const Windows& windows) {
    std::vector<float> sums(net.output_size(), 0.f);
    const std::vector<float> flat = flatten(windows, net.input_size());
    check(szb_identify_sums(net.handle(), flat.data(), windows.size(), sums.data()));
    return sums;
}
}  // namespace detail

inline size_t identify_speaker(const SimpleNeuralNet& net, const std::vector<int16_t>& sample, const FeatureExtractor& extractor) {   // lib.rs:1285
    const std::vector<float> sums = detail::prob_sums(net, extractor.extract(sample));
    return sums.empty() ? 0 : detail::argmax_last(sums);
}

inline std::optional<size_t> identify_speaker_with_threshold_feats(const SimpleNeuralNet& net, const Windows& windows, float threshold) {   // lib.rs:1346
    if (net.output_size() <= 1 || windows.empty()) return std::nullopt;
    const std::vector<float> sums = detail::prob_sums(net, windows);
    const size_t best = detail::argmax_last(sums);
    if (sums[best] / float(windows.size()) >= threshold) return best;
    return std::nullopt;
}

inline std::optional<size_t> identify_speaker_with_threshold(const SimpleNeuralNet& net, const std::vector<int16_t>& sample, float threshold,
                                                             const FeatureExtractor& extractor) {   // lib.rs:1307
    if (net.output_size() <= 1) return std::nullopt;
    return identify_speaker_with_threshold_feats(net, extractor.extract(sample), threshold);
}

inline std::vector<size_t> identify_speaker_list(const SimpleNeuralNet& net, const std::vector<int16_t>& sample, float threshold,
                                                 const FeatureExtractor& /*extractor*/) {   // lib.rs:1383
    std::vector<uint32_t> out(std::max<size_t>(net.output_size(), 1));
    uint32_t n = 0;
    check(szb_identify_speaker_list(net.handle(), sample.data(), sample.size(), threshold, out.data(), uint32_t(out.size()), &n));
    return std::vector<size_t>(out.begin(), out.begin() + n);
}

// ---- embeddings and cosine matching (lib.rs:893-905, 1073-1079, 1413-1540) -------------------------------------------
inline std::vector<float> embed(const SimpleNeuralNet& net, const std::vector<float>& bits) {   // lib.rs:895
    uint32_t h2 = 0;
    check(szb_net_embedding_size(net.handle(), &h2));
    std::vector<float> out(h2);
    check(szb_net_embed(net.handle(), bits.data(), 1, 0, out.data()));
    return out;
}
inline std::vector<float> forward_embedding(const SimpleNeuralNet& net, const std::vector<float>& input) {   // lib.rs:1073
    uint32_t h2 = 0;
    check(szb_net_embedding_size(net.handle(), &h2));
    std::vector<float> out(h2);
    check(szb_net_embed(net.handle(), input.data(), 1, 1, out.data()));
    return out;
}
inline std::vector<float> extract_embedding_from_features(const SimpleNeuralNet& net, const Windows& feats) {   // lib.rs:1453
    uint32_t h2 = 0;
    check(szb_net_embedding_size(net.handle(), &h2));
    std::vector<float> out(h2);
    const std::vector<float> flat = detail::flatten(feats, net.input_size());
    check(szb_net_embedding_mean(net.handle(), flat.data(), feats.size(), out.data()));
    return out;
}
inline std::vector<float> median_embedding_from_features(const SimpleNeuralNet& net, const Windows& feats) {   // lib.rs:1478
    uint32_t h2 = 0;
    check(szb_net_embedding_size(net.handle(), &h2));
    std::vector<float> out(h2);
    const std::vector<float> flat = detail::flatten(feats, net.input_size());
    check(szb_net_embedding_median(net.handle(), flat.data(), feats.size(), 1, out.data()));
    return out;
}
inline std::vector<float> extract_embedding(const SimpleNeuralNet& net, const std::vector<int16_t>& sample, const FeatureExtractor& extractor) {   // lib.rs:1418
    uint32_t h2 = 0;
    check(szb_net_embedding_size(net.handle(), &h2));
    std::vector<float> out(h2);
    const Windows w = extractor.extract(sample);
    const std::vector<float> flat = detail::flatten(w, net.input_size());
    check(szb_net_embedding_median(net.handle(), flat.data(), w.size(), 0, out.data()));
    return out;
}
inline float cosine_similarity(const std::vector<float>& a, const std::vector<float>& b) {   // lib.rs:1531
    return szb_cosine_similarity(a.data(), b.data(), uint32_t(std::min(a.size(), b.size())));
}

// lib.rs:558-579; decoding stays on the host and is out of scope, so the caller passes the decoded mono 44.1 kHz samples.
inline Windows load_cached_features(const std::string& path, const std::function<std::vector<int16_t>(const std::string&)>& load_audio_samples,
                                    const FeatureExtractor& extractor) {
    char cache[4096];
    check(szb_feature_cache_path(path.c_str(), cache, sizeof cache));
    uint64_t rows = 0, cols = 0;
    if (szb_npy_read_f32(cache, nullptr, 0, &rows, &cols) == SZB_OK) {
        std::vector<float> flat(rows * cols);
        check(szb_npy_read_f32(cache, flat.data(), flat.size(), &rows, &cols));
        return detail::unflatten(flat, rows, cols);
    }
    Windows feats = extractor.extract(load_audio_samples(path));
    if (!feats.empty()) {
        const std::vector<float> flat = detail::flatten(feats, feats[0].size());
        (void)szb_npy_write_f32(cache, flat.data(), feats.size(), feats[0].size());   // `let _ = write_npy(..)`, lib.rs:576
    }
    return feats;
}

}  // namespace streamz_rs
