// C++ host-layer test: the reference's own unit test (lib.rs:1832-1851) and the C1 flow (main.rs:500-508, 658-666 ->
// extract, train_from_feature_map, save/load, identify_speaker_list) written against include/streamz_rs.hpp exactly as a
// Rust caller writes them against streamz_rs.  Built and run by tests/test_gpu_cpp_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "streamz_rs.hpp"

using namespace streamz_rs;

#define REQUIRE(c)                                                        \
    do {                                                                  \
        if (!(c)) {                                                       \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                     \
        }                                                                 \
    } while (0)

static std::vector<int16_t> tone_clip(int speaker, double seconds) {
    const size_t n = size_t(seconds * 44100);
    std::vector<int16_t> s(n);
    uint32_t lcg = 12345u + speaker;
    for (size_t i = 0; i < n; ++i) {
        const double t = double(i) / 44100.0;
        double x = 0.0;
        for (int h = 1; h <= 12; ++h) x += std::sin(2 * M_PI * (110.0 + 70.0 * speaker) * h * t) / h;
        lcg = lcg * 1664525u + 1013904223u;
        x = 0.3 * x + 0.02 * (double(lcg >> 8) / 8388608.0 - 1.0);
        s[i] = int16_t(std::lround(std::fmax(-1.0, std::fmin(1.0, x * 0.5)) * 32767.0));
    }
    return s;
}

int main() {
    // --- weights_change_after_training (lib.rs:1832-1851) ---
    {
        SimpleNeuralNet net(4, 3, 2, 2);
        const std::vector<float> x = { 0.1f, -0.2f, 0.3f, 0.4f };
        const std::vector<float> before = net.forward(x);
        net.train_batch({ x }, { 1.0f, 0.0f }, 0.1f);
        const std::vector<float> after = net.forward(x);
        REQUIRE(before.size() == 2 && after.size() == 2);
        REQUIRE(after[0] > before[0]);   // one SGD step towards class 0 must raise its probability
        REQUIRE(std::fabs(after[0] + after[1] - 1.0f) < 1e-5f);
    }
    // --- C1-shaped flow ---
    FeatureExtractor ex;
    REQUIRE(ex.extract(std::vector<int16_t>(799)).empty());                      // lib.rs:289
    std::map<std::string, Windows> fmap;
    std::vector<std::pair<std::string, size_t>> files;
    for (int spk = 0; spk < 2; ++spk)
        for (int k = 0; k < 2; ++k) {
            const std::string path = "spk" + std::to_string(spk) + "_" + std::to_string(k) + ".wav";
            fmap[path] = with_thread_extractor([&](const FeatureExtractor& e) { return e.extract(tone_clip(spk, 1.0 + 0.1 * k)); });
            REQUIRE(fmap[path].size() == szb_num_windows(size_t((1.0 + 0.1 * k) * 44100)) && fmap[path][0].size() == FEATURE_SIZE);
            files.emplace_back(path, size_t(spk));
        }
    SimpleNeuralNet net(FEATURE_SIZE, 512, 256, 2, 1);
    // train_from_feature_map visits the files one after the other (lib.rs:643-658); several one-epoch rounds interleave
    // the two speakers so that the last file does not wipe out the first
    const float l0 = train_from_feature_map(net, fmap, files, 1, 0.01f, DEFAULT_DROPOUT, 8, 1);
    float l1 = l0;
    for (int round = 0; round < 12; ++round) l1 = train_from_feature_map(net, fmap, files, 1, 0.01f, DEFAULT_DROPOUT, 8, 2 + round);
    std::printf("loss %.4f -> %.4f\n", l0, l1);
    REQUIRE(std::isfinite(l0) && l1 < l0);
    const auto who0 = identify_speaker_list(net, tone_clip(0, 1.0), 0.5f, ex);
    const auto who1 = identify_speaker_list(net, tone_clip(1, 1.0), 0.5f, ex);
    REQUIRE(!who0.empty() && who0[0] == 0);
    REQUIRE(!who1.empty() && who1[0] == 1);
    REQUIRE(identify_speaker(net, tone_clip(1, 0.7), ex) == 1);
    REQUIRE(identify_speaker_with_threshold(net, tone_clip(0, 0.7), 0.5f, ex).value_or(99) == 0);
    REQUIRE(!identify_speaker_with_threshold(net, tone_clip(0, 0.7), 1.5f, ex).has_value());
    // --- raw-audio training loops (lib.rs:348-397, 668-732): augment -> extract -> train per epoch, lr * 0.99^step ---
    {
        SimpleNeuralNet raw(FEATURE_SIZE, 512, 256, 2, 4);
        const float first = pretrain_network(raw, tone_clip(0, 0.6), 0, 2, 1, 0.01f, DEFAULT_DROPOUT, 8, 21);
        REQUIRE(std::isfinite(first) && first > 0.f);
        REQUIRE(pretrain_network(raw, std::vector<int16_t>(500), 0, 2, 3, 0.01f, DEFAULT_DROPOUT, 8, 21) == 0.0f);   // lib.rs:392-396
        std::vector<TrainingClip> clips = { { "a.wav", tone_clip(0, 0.5), 0 }, { "b.wav", tone_clip(1, 0.5), 1 } };
        const float l_start = train_from_files(raw, clips, 2, 1, 0.01f, DEFAULT_DROPOUT, 8, 30);
        float lf = l_start;
        for (int round = 1; round < 8; ++round) lf = train_from_files(raw, clips, 2, 1, 0.01f, DEFAULT_DROPOUT, 8, 30 + round);
        std::printf("train_from_files loss %.4f -> %.4f\n", l_start, lf);
        REQUIRE(std::isfinite(lf) && lf < l_start);       // both speakers interleaved: the two-class loss falls over the rounds
        REQUIRE(identify_speaker(raw, tone_clip(1, 0.5), ex) == 1 && identify_speaker(raw, tone_clip(0, 0.5), ex) == 0);
    }
    // --- save / load round trip (lib.rs:1081-1282), with the speaker embeddings the CLI stores before saving (main.rs:845-856) ---
    const char* path = "/tmp/streamz_b200_cpp_model.npz";
    net.set_embeddings({ { std::vector<float>(256, 0.5f), 0.9f, 0.01f }, { std::vector<float>(256, -0.25f), 0.8f, 0.02f } });
    net.save(path);
    SimpleNeuralNet back = SimpleNeuralNet::load(path);
    REQUIRE(back.output_size() == 2);
    REQUIRE(back.embeddings().size() == 2 && std::get<0>(back.embeddings()[1])[7] == -0.25f && std::get<1>(back.embeddings()[0]) == 0.9f);
    const auto p0 = net.forward(fmap.begin()->second[3]), p1 = back.forward(fmap.begin()->second[3]);
    REQUIRE(p0[0] == p1[0] && p0[1] == p1[1]);
    net.add_output_class();
    REQUIRE(net.output_size() == 3 && net.forward(fmap.begin()->second[0]).size() == 3);
    // resampler contract: identity at 44.1 kHz, length floor(n * 44100 / rate)
    const std::vector<int16_t> c = tone_clip(0, 0.25);
    REQUIRE(resample_to_44100(c, 44100) == c);
    REQUIRE(resample_to_44100(std::vector<int16_t>(16000, 100), 16000).size() == 44100);
    std::puts("host_api_test ok");
    return 0;
}
