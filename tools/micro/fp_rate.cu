// Microbenchmark: sustained FP32 issue rate per SM for the instruction forms the front-end kernel uses, at the
// kernel's occupancy (640 threads = 5 warps per scheduler, 1 CTA per SM).  nvcc -arch=sm_100a -O3 fp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(640, 1) k(float* out, int iters, float s) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    float b = s, c = s * 0.5f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], b, c);                 // FFMA 3 registers
            if (MODE == 1) a[i] = a[i] + b;                         // FADD
            if (MODE == 2) a[i] = a[i] * b;                         // FMUL
            if (MODE == 3) a[i] = (i & 1) ? fmaf(a[i], b, c) : a[i] + b;   // mixed
            if (MODE == 4) a[i] = fmaf(a[i], 1.0001f, c);           // FFMA with immediate
        }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, float* d) {
    const int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, 640>>>(d, 100, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148, 640>>>(d, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double inst = double(iters) * 8 * 640 * 148;          // thread-instructions
    double per_clk_sm = inst / (ms * 1e-3) / 148 / (clk_khz * 1e3);
    printf("%-28s %8.3f ms  %6.1f lane-instr/clk/SM (at nominal %d MHz)\n", name, ms, per_clk_sm, clk_khz / 1000);
}
int main() { float* d; cudaMalloc(&d, 148 * 640 * 4);
    run<0>("FFMA reg,reg,reg", d); run<1>("FADD reg,reg", d); run<2>("FMUL reg,reg", d); run<3>("FADD/FFMA mixed", d); run<4>("FFMA reg,imm,reg", d);
    return 0; }
