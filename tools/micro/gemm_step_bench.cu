// Microbenchmark: the eight tcgen05 GEMMs of one SimpleNeuralNet training step (batch 4096, 60-512-256-100) launched
// through the library's own gemm_tc.cuh, in step order, timed (a) per GEMM with CUDA events inside the chain and (b) per
// phase inside every CTA with %globaltimer stamps (SZB_GEMM_TRACE): entry, set-up done, first k-block staged, all k-blocks
// staged, last MMA complete, epilogue done.  Not part of the library; build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DSZB_GEMM_TRACE -I streamz_b200/csrc \
//        tools/micro/gemm_step_bench.cu -o tools/micro/gemm_step_bench
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gemm_tc.cuh"
#include "gemm_tma.cuh"

namespace szb {
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
}  // namespace szb

using namespace szb;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

static float* dalloc(size_t n, float scale, unsigned seed) {
    std::vector<float> h(n);
    unsigned s = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        h[i] = scale * (float((s >> 8) & 0xFFFF) / 32768.f - 1.f);
    }
    float* d;
    CK(cudaMalloc(&d, n * sizeof(float)));
    CK(cudaMemcpy(d, h.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

struct Job {
    const char* name;
    int epi;
    tc::GemmArgs g;
    int split;
    bool narrow;
    int passes;
};

template <int EPI>
static szb_status run_epi(szb_ctx* ctx, const Job& j) {
    if (ctx->gemm_tma) {                 // third generation: TMA operand fetch, two producer groups (gemm_tma.cuh)
        bool done = false;
        szb_status s;
        if (j.passes == 1) s = j.narrow ? tc::launch_gemm_tma<64, 1, EPI>(ctx, j.g, j.split, &done) : tc::launch_gemm_tma<128, 1, EPI>(ctx, j.g, j.split, &done);
        else s = j.narrow ? tc::launch_gemm_tma<64, 3, EPI>(ctx, j.g, j.split, &done) : tc::launch_gemm_tma<128, 3, EPI>(ctx, j.g, j.split, &done);
        if (s != SZB_OK || done) return s;
        fprintf(stderr, "TMA launch fell back\n");
    }
    if (j.passes == 1)
        return j.narrow ? tc::launch_gemm_tc<64, 1, EPI>(ctx, j.g, j.split) : tc::launch_gemm_tc<128, 1, EPI>(ctx, j.g, j.split);
    return j.narrow ? tc::launch_gemm_tc<64, 3, EPI>(ctx, j.g, j.split) : tc::launch_gemm_tc<128, 3, EPI>(ctx, j.g, j.split);
}
static void run(szb_ctx* ctx, const Job& j) {
    szb_status s = SZB_OK;
    switch (j.epi) {
        case tc::TC_BIAS: s = run_epi<tc::TC_BIAS>(ctx, j); break;
        case tc::TC_BIAS_RELU: s = run_epi<tc::TC_BIAS_RELU>(ctx, j); break;
        case tc::TC_BIAS_TANH: s = run_epi<tc::TC_BIAS_TANH>(ctx, j); break;
        case tc::TC_MUL_DTANH: s = run_epi<tc::TC_MUL_DTANH>(ctx, j); break;
        case tc::TC_MUL_DRELU: s = run_epi<tc::TC_MUL_DRELU>(ctx, j); break;
        default: s = run_epi<tc::TC_ATOMIC>(ctx, j); break;
    }
    if (s != SZB_OK) exit(2);
}

int main(int argc, char** argv) {
    const int passes = argc > 1 ? atoi(argv[1]) : 3;
    const int iters = argc > 2 ? atoi(argv[2]) : 50;
    const int B = 4096, I = 60, H1 = 512, H2 = 256, C = 100;
    szb_ctx ctx;
    ctx.pdl = argc > 3 ? atoi(argv[3]) != 0 : true;
    const int kernel_gen = argc > 4 ? atoi(argv[4]) : 0;     // 0: operands from shared memory, 1: A in tensor memory (gemm_tc_ta_kernel), 2: + TMA (gemm_tma_kernel)
    const bool use_ta = kernel_gen != 0;
    ctx.gemm_tma = false;
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    ctx.sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreate(&ctx.stream));

    float* xb = dalloc(size_t(B) * I, 1.f, 1);
    float* xbT = dalloc(size_t(I + 1) * B, 1.f, 2);
    float* wt1 = dalloc(size_t(H1) * I, .5f, 3);
    float* wt2 = dalloc(size_t(H2) * H1, .5f, 4);
    float* wt3 = dalloc(size_t(C) * H2, .5f, 5);
    float* w2 = dalloc(size_t(H1) * H2, .5f, 6);
    float* w3 = dalloc(size_t(H2) * C, .5f, 7);
    float* bias = dalloc(1024, .1f, 8);
    float* h1 = dalloc(size_t(B) * H1, 1.f, 9);
    float* h1T = dalloc(size_t(H1 + 1) * B, 1.f, 10);
    float* h2 = dalloc(size_t(B) * H2, 1.f, 11);
    float* h2T = dalloc(size_t(H2 + 1) * B, 1.f, 12);
    float* z = dalloc(size_t(B) * C, 1.f, 13);
    float* zT = dalloc(size_t(C) * B, 1.f, 14);
    float* d2 = dalloc(size_t(B) * H2, 1.f, 15);
    float* d2T = dalloc(size_t(H2) * B, 1.f, 16);
    float* d1 = dalloc(size_t(B) * H1, 1.f, 17);
    float* d1T = dalloc(size_t(H1) * B, 1.f, 18);
    float* G = dalloc(size_t(300000), 0.f, 19);

    auto narrow_for = [&](int M, int N, int split) {
        const int tiles128 = ((M + 127) / 128) * ((N + 127) / 128) * std::max(1, split);
        return tiles128 * 10 < ctx.sm_count * 7 && N > 64;
    };
    auto split_for = [&](int M, int N) {
        const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
        return std::max(1, ctx.sm_count / tiles);
    };
    std::vector<Job> jobs;
    auto add = [&](const char* name, int epi, const float* A, int lda, const float* Bm, int ldb, float* Cm, int ldc, float* CT, int ldct,
                   const float* aux, int ldaux, int M, int N, int K, int split) {
        Job j{};
        j.name = name; j.epi = epi; j.split = split; j.passes = passes;
        j.g.A = A; j.g.lda = lda; j.g.B = Bm; j.g.ldb = ldb; j.g.C = Cm; j.g.ldc = ldc; j.g.CT = CT; j.g.ldct = ldct;
        j.g.bias = bias; j.g.aux = aux; j.g.ldaux = ldaux; j.g.M = M; j.g.N = N; j.g.K = K;
        j.narrow = narrow_for(M, N, epi == tc::TC_ATOMIC ? split : 1);
        jobs.push_back(j);
    };
    add("fwd1  X*W1   relu ", tc::TC_BIAS_RELU, xb, I, wt1, I, h1, H1, h1T, B, nullptr, 0, B, H1, I, 1);
    add("fwd2  H1*W2  tanh ", tc::TC_BIAS_TANH, h1, H1, wt2, H1, h2, H2, h2T, B, nullptr, 0, B, H2, H1, 1);
    add("fwd3  H2*W3  bias ", tc::TC_BIAS, h2, H2, wt3, H2, z, C, nullptr, 0, nullptr, 0, B, C, H2, 1);
    add("dW3   H2T*dZ  atom", tc::TC_ATOMIC, h2T, B, zT, B, G, C, nullptr, 0, nullptr, 0, H2 + 1, C, B, split_for(H2 + 1, C));
    add("dX2   dZ*W3T dtanh", tc::TC_MUL_DTANH, z, C, w3, C, d2, H2, d2T, B, h2T, B, B, H2, C, 1);
    add("dW2   H1T*d2  atom", tc::TC_ATOMIC, h1T, B, d2T, B, G + 30000, H2, nullptr, 0, nullptr, 0, H1 + 1, H2, B, split_for(H1 + 1, H2));
    add("dX1   d2*W2T drelu", tc::TC_MUL_DRELU, d2, H2, w2, H2, nullptr, H1, d1T, B, h1T, B, B, H1, H2, 1);
    add("dW1   XT*d1   atom", tc::TC_ATOMIC, xbT, B, d1T, B, G + 170000, H1, nullptr, 0, nullptr, 0, I + 1, H1, B, split_for(I + 1, H1));

    const int n_chain = int(jobs.size());
    // epilogue experiments (traced only, not part of the timed chain): same K loop as dX1, different outputs
    float* big = dalloc(size_t(H1) * (B + 32), 1.f, 20);
    add("x: aux loads only  ", tc::TC_MUL_DRELU, d2, H2, w2, H2, nullptr, H1, nullptr, B, h1T, B, B, H1, H2, 1);
    add("x: CT stores only  ", tc::TC_BIAS_RELU, d2, H2, w2, H2, nullptr, H1, d1T, B, nullptr, 0, B, H1, H2, 1);
    add("x: C stores only   ", tc::TC_BIAS_RELU, d2, H2, w2, H2, d1, H1, nullptr, B, nullptr, 0, B, H1, H2, 1);
    add("x: CT ld=B+32      ", tc::TC_BIAS_RELU, d2, H2, w2, H2, nullptr, H1, big, B + 32, nullptr, 0, B, H1, H2, 1);
    add("x: no output       ", tc::TC_BIAS_RELU, d2, H2, w2, H2, nullptr, H1, nullptr, B, nullptr, 0, B, H1, H2, 1);
    const int nj = int(jobs.size());
    if (use_ta) {
        // correctness first: every GEMM of the step with A from shared memory and with A from tensor memory, same inputs
        bool all_ok = true;
        for (int k = 0; k < n_chain; ++k) {
            const Job& j = jobs[k];
            const size_t nc = j.g.C ? size_t(j.g.M) * j.g.ldc : 0, nt = j.g.CT ? size_t(j.g.N) * j.g.ldct : 0;
            std::vector<float> c0(nc), c1(nc), t0(nt), t1(nt);
            for (int pass = 0; pass < 2; ++pass) {
                ctx.gemm_ta = pass == 1;
                ctx.gemm_tma = pass == 1 && kernel_gen == 2;
                if (nc) CK(cudaMemsetAsync(j.g.C, 0, nc * 4, ctx.stream));
                if (nt) CK(cudaMemsetAsync(j.g.CT, 0, nt * 4, ctx.stream));
                run(&ctx, j);
                CK(cudaStreamSynchronize(ctx.stream));
                if (nc) CK(cudaMemcpy((pass ? c1 : c0).data(), j.g.C, nc * 4, cudaMemcpyDeviceToHost));
                if (nt) CK(cudaMemcpy((pass ? t1 : t0).data(), j.g.CT, nt * 4, cudaMemcpyDeviceToHost));
            }
            double dc = 0, mc = 0, dt = 0, mt = 0;
            for (size_t i = 0; i < nc; ++i) { dc = std::max(dc, double(std::fabs(c0[i] - c1[i]))); mc = std::max(mc, double(std::fabs(c0[i]))); }
            for (size_t i = 0; i < nt; ++i) { dt = std::max(dt, double(std::fabs(t0[i] - t1[i]))); mt = std::max(mt, double(std::fabs(t0[i]))); }
            const bool ok = (nc == 0 || dc <= 2e-5 * mc) && (nt == 0 || dt <= 2e-5 * mt) && (nc == 0 || mc > 0);
            all_ok = all_ok && ok;
            printf("check %-20s C: max|diff| %.3e of max %.3e   CT: max|diff| %.3e of max %.3e   %s\n", j.name, dc, mc, dt, mt, ok ? "ok" : "MISMATCH");
        }
        printf("generation-%d kernel vs shared-memory kernel: %s\n", kernel_gen, all_ok ? "all GEMMs agree" : "MISMATCH");
        ctx.gemm_ta = true;
        ctx.gemm_tma = kernel_gen == 2;
    }
    std::vector<cudaEvent_t> ev(nj + 1);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    for (int w = 0; w < 3; ++w)
        for (auto& j : jobs) run(&ctx, j);
    CK(cudaStreamSynchronize(ctx.stream));

    // (a) whole chain, no events in between
    cudaEvent_t c0, c1;
    CK(cudaEventCreate(&c0));
    CK(cudaEventCreate(&c1));
    CK(cudaEventRecord(c0, ctx.stream));
    for (int it = 0; it < iters; ++it)
        for (int k = 0; k < n_chain; ++k) run(&ctx, jobs[k]);
    CK(cudaEventRecord(c1, ctx.stream));
    CK(cudaStreamSynchronize(ctx.stream));
    float chain_ms = 0;
    CK(cudaEventElapsedTime(&chain_ms, c0, c1));
    printf("passes %d: chain of %d GEMMs: %.1f us per step (%d iterations)\n", passes, n_chain, chain_ms * 1000.f / iters, iters);

    // (b) per GEMM, events inside the chain
    std::vector<double> acc(nj, 0.0);
    for (int it = 0; it < iters; ++it) {
        CK(cudaEventRecord(ev[0], ctx.stream));
        for (int k = 0; k < nj; ++k) {
            run(&ctx, jobs[k]);
            CK(cudaEventRecord(ev[k + 1], ctx.stream));
        }
        CK(cudaStreamSynchronize(ctx.stream));
        for (int k = 0; k < nj; ++k) {
            float ms;
            CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            acc[k] += ms;
        }
    }
#ifdef SZB_GEMM_TRACE
    // (c) per-CTA phase stamps
    const int max_ctas = 1024;
    unsigned long long* d_trace;
    CK(cudaMalloc(&d_trace, size_t(nj) * max_ctas * 8 * sizeof(unsigned long long)));
    CK(cudaMemset(d_trace, 0, size_t(nj) * max_ctas * 8 * sizeof(unsigned long long)));
    for (int k = 0; k < nj; ++k) {
        unsigned long long* p = d_trace + size_t(k) * max_ctas * 8;
        CK(cudaMemcpyToSymbolAsync(tc::g_gemm_trace, &p, sizeof(p), 0, cudaMemcpyHostToDevice, ctx.stream));
        run(&ctx, jobs[k]);
    }
    CK(cudaStreamSynchronize(ctx.stream));
    std::vector<unsigned long long> tr(size_t(nj) * max_ctas * 8);
    CK(cudaMemcpy(tr.data(), d_trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    printf("%-20s %5s %5s | %7s | mean over CTAs, us after CTA entry: %6s %6s %6s %6s %6s %6s | %6s %6s\n", "gemm", "BN", "ctas", "evt us",
           "setup", "kb0", "staged", "mma", "epi0", "epiAll", "span", "spread");
#endif
    double total = 0;
    for (int k = 0; k < nj; ++k) {
        total += acc[k] / iters;
        const Job& j = jobs[k];
#ifdef SZB_GEMM_TRACE
        const unsigned long long* p = tr.data() + size_t(k) * max_ctas * 8;
        int n = 0;
        double s[7] = {0}, ghz = 0;
        unsigned long long tmin = ~0ull, tmax = 0, emax = 0;
        for (int c = 0; c < max_ctas; ++c) {
            if (p[c * 8] == 0) continue;
            ++n;
            for (int q = 1; q < 7; ++q)
                if (p[c * 8 + q]) s[q] += double(p[c * 8 + q] - p[c * 8]);
            tmin = std::min(tmin, p[c * 8]);
            emax = std::max(emax, p[c * 8]);
            tmax = std::max(tmax, p[c * 8 + 6]);
            if (p[c * 8 + 4] > p[c * 8 + 1]) ghz += double(p[c * 8 + 7]) / double(p[c * 8 + 4] - p[c * 8 + 1]);
        }
        printf("%-20s %5d %5d | %7.2f | %39s %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f | %6.2f %6.2f\n", j.name, j.narrow ? 64 : 128, n,
               acc[k] * 1000.0 / iters, "", s[1] / n / 1e3, s[2] / n / 1e3, s[3] / n / 1e3, s[4] / n / 1e3, s[5] / n / 1e3, s[6] / n / 1e3,
               double(tmax - tmin) / 1e3, double(emax - tmin) / 1e3);
        printf("%-20s SM clock between set-up and last MMA: %.3f GHz\n", "", ghz / n);
#else
        printf("%-20s BN %3d | %7.2f us\n", j.name, j.narrow ? 64 : 128, acc[k] * 1000.0 / iters);
#endif
    }
    printf("sum of per-GEMM event times (incl. the x: experiments): %.1f us\n", total * 1000.0);
    return 0;
}
