// Standalone check of the tensor-core integer resampler (streamz_b200/csrc/resample_tc.cuh) against a scalar CPU evaluation
// of the same integer specification, plus a timing of the C2-sized batch.
// cd tools/micro/resample_tc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o resample_tc_test resample_tc_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "resample_tc_host.hpp"
namespace szb { void set_error(const char*, ...) {} }
using namespace szb;

static void cpu_ref(const std::vector<int16_t>& x, uint32_t rate, std::vector<int16_t>& y) {
    uint32_t L, M; resample_ratio(rate, L, M);
    const auto cq = resample_taps_q(rate);
    const uint64_t n_out = uint64_t(x.size()) * 44100ull / rate;
    y.resize(n_out);
    for (uint64_t j = 0; j < n_out; ++j) {
        const uint64_t pos = j * M; const int64_t i0 = int64_t(pos / L); const uint32_t p = uint32_t(pos % L);
        long long acc = 0;
        for (int t = 0; t < 16; ++t) { const int64_t i = i0 - 7 + t; if (i >= 0 && i < int64_t(x.size())) acc += (long long)cq[size_t(p) * 16 + t] * x[size_t(i)]; }
        long long v = (acc + 16384) >> 15; if (v < -32768) v = -32768; if (v > 32767) v = 32767;
        y[j] = int16_t(v);
    }
}

int main(int argc, char** argv) {
    const uint32_t rate = argc > 1 ? uint32_t(atoi(argv[1])) : 16000;
    const int n_clips = argc > 2 ? atoi(argv[2]) : 3;
    rtc::Plan plan;
    if (!rtc::make_plan(rate, plan)) { printf("rate %u not supported by the tensor-core kernel\n", rate); return 2; }
    printf("rate %u: L = %d, chunks = %d, K-steps per chunk:", rate, plan.geom.L, plan.geom.n_chunks);
    for (int n = 0; n < plan.geom.n_chunks; ++n) printf(" %d+%d", plan.geom.ks_first[n], plan.geom.ks_count[n]);
    printf("\n");
    // ragged clips incl. one shorter than a tile and one with an odd length
    std::vector<std::vector<int16_t>> clips;
    srand(7);
    for (int c = 0; c < n_clips; ++c) {
        const size_t n = c == 0 ? rate * 3 + 37 : (c == 1 ? 5000 : rate * 10);
        std::vector<int16_t> x(n);
        for (size_t i = 0; i < n; ++i) x[i] = int16_t((rand() % 65536) - 32768);
        if (c == 2) for (size_t i = 0; i < n; ++i) x[i] = (i / 40) % 2 ? 32767 : -32768;     // full-scale square wave: clamps
        clips.push_back(x);
    }
    std::vector<unsigned long long> in_off(n_clips + 1, 0), out_off(n_clips + 1, 0);
    uint64_t max_out = 0;
    for (int c = 0; c < n_clips; ++c) {
        in_off[c + 1] = in_off[c] + ((clips[c].size() + 7) & ~size_t(7));                      // clips start on 16 bytes
        const uint64_t n_out = uint64_t(clips[c].size()) * 44100ull / rate;
        out_off[c + 1] = (out_off[c] + n_out + 7) & ~7ull;
        max_out = std::max(max_out, n_out);
    }
    // note: in_off[c+1] - in_off[c] is used as the clip length by the kernel, so keep exact lengths in a second table
    std::vector<unsigned long long> in_len_off(n_clips + 1, 0);
    std::vector<int16_t> pcm(in_off[n_clips] + 8, 0);
    for (int c = 0; c < n_clips; ++c) std::copy(clips[c].begin(), clips[c].end(), pcm.begin() + in_off[c]);
    int16_t *d_in, *d_out; unsigned long long *d_ioff, *d_ooff; uint8_t* d_bt;
    cudaMalloc(&d_in, pcm.size() * 2); cudaMalloc(&d_out, (out_off[n_clips] + 8) * 2);
    cudaMalloc(&d_ioff, (n_clips + 1) * 8); cudaMalloc(&d_ooff, (n_clips + 1) * 8); cudaMalloc(&d_bt, plan.btiles.size());
    cudaMemcpy(d_in, pcm.data(), pcm.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ooff, out_off.data(), (n_clips + 1) * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_bt, plan.btiles.data(), plan.btiles.size(), cudaMemcpyHostToDevice);
    const size_t smem = rtc::smem_bytes(plan.geom.n_chunks);
    cudaFuncSetAttribute(rtc::resample_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    long bad_total = 0;
    for (int c = 0; c < n_clips; ++c) {          // one clip per launch so that exact (unpadded) lengths can be passed
        const unsigned long long io[2] = { in_off[c], in_off[c] + clips[c].size() }, oo[2] = { out_off[c], 0 };
        cudaMemcpy(d_ioff, io, 16, cudaMemcpyHostToDevice); cudaMemcpy(d_ooff, oo, 16, cudaMemcpyHostToDevice);
        const uint64_t n_out = uint64_t(clips[c].size()) * 44100ull / rate;
        rtc::Args a{ d_in, d_ioff, d_ooff, d_bt, d_out, 1u, uint32_t(((n_out + plan.geom.L - 1) / plan.geom.L + rtc::kRows - 1) / rtc::kRows), rate };
        cudaMemset(d_out + out_off[c], 0x55, n_out * 2);
        rtc::resample_tc_kernel<<<148, rtc::kThreads, smem>>>(plan.geom, a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<int16_t> got(n_out), want;
        cudaMemcpy(got.data(), d_out + out_off[c], n_out * 2, cudaMemcpyDeviceToHost);
        cpu_ref(clips[c], rate, want);
        long bad = 0; long first = -1;
        for (size_t i = 0; i < n_out; ++i) if (got[i] != want[i]) { if (first < 0) first = long(i); ++bad; }
        printf("clip %d: %zu in -> %llu out, mismatches %ld%s\n", c, clips[c].size(), (unsigned long long)n_out, bad, bad ? "" : "  (bit-exact)");
        if (bad) printf("   first at %ld: got %d want %d\n", first, got[first], want[first]);
        bad_total += bad;
    }
    // timing: 10 000 clips of 10 s (the C2 batch) unless the device is small
    {
        const int N = 10000; const size_t n = size_t(rate) * 10;
        int16_t *b_in, *b_out; unsigned long long *b_io, *b_oo;
        const uint64_t n_out = uint64_t(n) * 44100ull / rate, stride_out = (n_out + 7) & ~7ull;
        if (cudaMalloc(&b_in, N * n * 2) == cudaSuccess && cudaMalloc(&b_out, N * stride_out * 2) == cudaSuccess) {
            cudaMemset(b_in, 1, N * n * 2);
            std::vector<unsigned long long> io(N + 1), oo(N + 1);
            for (int c = 0; c <= N; ++c) { io[c] = c * n; oo[c] = c * stride_out; }
            cudaMalloc(&b_io, (N + 1) * 8); cudaMalloc(&b_oo, (N + 1) * 8);
            cudaMemcpy(b_io, io.data(), (N + 1) * 8, cudaMemcpyHostToDevice); cudaMemcpy(b_oo, oo.data(), (N + 1) * 8, cudaMemcpyHostToDevice);
            rtc::Args a{ b_in, b_io, b_oo, d_bt, b_out, uint32_t(N), uint32_t(((n_out + plan.geom.L - 1) / plan.geom.L + rtc::kRows - 1) / rtc::kRows), rate };
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            rtc::resample_tc_kernel<<<148, rtc::kThreads, smem>>>(plan.geom, a);
            cudaEventRecord(e0);
            for (int i = 0; i < 3; ++i) rtc::resample_tc_kernel<<<148, rtc::kThreads, smem>>>(plan.geom, a);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("C2 batch (%d clips x 10 s @ %u Hz): %.3f ms per launch (%s)\n", N, rate, ms / 3, cudaGetErrorString(cudaGetLastError()));
        }
    }
    printf("%s\n", bad_total ? "FAIL" : "OK");
    return bad_total ? 1 : 0;
}
