// Microbenchmark: cost of __syncthreads() for a 640-thread CTA (1 CTA/SM), alone and with a little work between barriers.
#include <cstdio>
#include <cuda_runtime.h>
template <int WORK>
__global__ void __launch_bounds__(640, 1) k(float* out, int iters) {
    __shared__ float s[640];
    float a = threadIdx.x;
    s[threadIdx.x] = a;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int b = 0; b < 7; ++b) {
            if (WORK) { a = fmaf(a, 1.0001f, s[(threadIdx.x + b * 32 + it) % 640]); s[threadIdx.x] = a; }
            __syncthreads();
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + s[0];
}
template <int WORK> void run(const char* name, float* d) {
    const int iters = 2000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<WORK><<<148, 640>>>(d, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<WORK><<<148, 640>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-24s %8.3f ms  %.1f ns per barrier (%.0f cycles @1.965 GHz)\n", name, ms, ms * 1e6 / (iters * 7.0), ms * 1e6 / (iters * 7.0) * 1.965);
}
int main() { float* d; cudaMalloc(&d, 148 * 640 * 4); run<0>("barrier only", d); run<1>("barrier + LDS/FFMA/STS", d); return 0; }
