// Microbenchmark: FP32 issue rate per SM by operand form (register / constant-bank / immediate), 8 independent chains per
// thread, at 5 and 2 warps per scheduler.  nvcc -arch=sm_100a -O3 fp_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
struct P { float v[16]; };
template <int MODE>
__global__ void k(float* out, int iters, const __grid_constant__ P p, const float* g) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    float b[8], c[8];
    for (int i = 0; i < 8; ++i) { b[i] = g[i + threadIdx.x % 3]; c[i] = g[8 + i + threadIdx.x % 5]; }   // true registers
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b[i], c[i]);                  // FFMA R, R, R (3 distinct registers)
            if (MODE == 1) a[i] = a[i] + b[i];                             // FADD R, R
            if (MODE == 2) a[i] = fmaf(a[i], p.v[i], c[i]);                // FFMA R, c[][], R
            if (MODE == 3) a[i] = fmaf(a[i], 1.0001f, c[i]);               // FFMA R, imm, R
            if (MODE == 4) a[i] = fmaf(b[i], p.v[i], a[i]);                // FFMA acc chain: R(b), c[][], R(acc)
            if (MODE == 5) a[i] = fmaf(b[i], c[(i + 1) & 7], a[i]);        // FFMA acc chain, 3 registers
            if (MODE == 6) a[i] = a[i] * b[i];                             // FMUL R, R
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* d, const float* g, int threads) {
    const int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    P p; for (int i = 0; i < 16; ++i) p.v[i] = 1.0f + 1e-4f * i;
    k<MODE><<<148, threads>>>(d, 100, p, g); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, iters, p, g); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = double(iters) * 8 * threads * 148;
    printf("%-34s threads %4d  %8.3f ms  %6.1f lane-instr/clk/SM @1965 MHz\n", name, threads, ms, inst / (ms * 1e-3) / 148 / 1.965e9);
}
int main() { float* d; cudaMalloc(&d, 148 * 1024 * 4); float* g; cudaMalloc(&g, 64 * 4); float h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0f + 1e-5f * i; cudaMemcpy(g, h, sizeof h, cudaMemcpyHostToDevice);
    for (int threads : {640, 256, 160}) {
        run<0>("FFMA a=a*R+R", d, g, threads); run<1>("FADD a=a+R", d, g, threads); run<2>("FFMA a=a*c[]+R", d, g, threads);
        run<3>("FFMA a=a*imm+R", d, g, threads); run<4>("FFMA a=R*c[]+a", d, g, threads); run<5>("FFMA a=R*R+a", d, g, threads); run<6>("FMUL a=a*R", d, g, threads);
    }
    return 0; }
