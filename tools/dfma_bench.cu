// tools/dfma_bench.cu — measures the non-tensor fp64 FMA rate of the device (SURVEY §8d risk note).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    double* d; cudaMalloc(&d, sizeof(double) * blocks * threads);
    k<<<blocks, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 8 * (double)iters * blocks * threads;
    printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"ms\": %.3f}\n", p.name, p.multiProcessorCount, fl / ms / 1e9, ms);
    return 0;
}
