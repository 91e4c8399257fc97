// Scratch micro-benchmark: dependent-issue latency of fp64 FMA / ADD, fp32 FMA, LDS on one warp, and fp64 FMA
// throughput of an SM with 8 and 16 warps (B200).  nvcc -arch=sm_100a -O3 -o dfma_lat dfma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, long long *cyc, int n, double a, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = x * a + b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dadd(double *out, long long *cyc, int n, double b) {
    double x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = x + b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma(float *out, long long *cyc, int n, float a, float b) {
    float x = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = x * a + b;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dfma_ilp(double *out, long long *cyc, int n, double a, double b) {   // 4 independent chains per thread
    double x0 = out[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b; }
    }
    __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_bar(long long *cyc, int n) {
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double *d; float *f; long long *c, h[4];
    cudaMalloc(&d, 1024 * 8); cudaMemset(d, 0, 1024 * 8); cudaMalloc(&f, 1024 * 4); cudaMemset(f, 0, 1024 * 4); cudaMalloc(&c, 64);
    const int n = 1000;
    for (int rep = 0; rep < 2; ++rep) {
        k_dfma<<<1, 32>>>(d, c, n, 1.0000001, 1e-9); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA dependent latency (1 warp): %.2f cycles/op\n", (double)h[0] / (16. * n));
        k_dadd<<<1, 32>>>(d, c, n, 1e-9); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DADD dependent latency (1 warp): %.2f cycles/op\n", (double)h[0] / (16. * n));
        k_ffma<<<1, 32>>>(f, c, n, 1.0000001f, 1e-9f); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("FFMA dependent latency (1 warp): %.2f cycles/op\n", (double)h[0] / (16. * n));
        for (int nt : {32, 128, 256, 512, 1024}) {
            k_dfma_ilp<<<1, nt>>>(d, c, n, 1.0000001, 1e-9); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("DFMA 4 chains/thread, %4d threads: %.2f cycles per warp-instruction slot, %.1f FMA/clk/SM\n", nt,
                            (double)h[0] / (16. * n), 16. * n * nt / (double)h[0]);
        }
        for (int nt : {32, 128, 256, 512}) {
            k_bar<<<1, nt>>>(c, 1000); cudaMemcpy(h, c, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("__syncthreads, %4d threads: %.1f cycles\n", nt, (double)h[0] / 1000.);
        }
    }
    return 0;
}
