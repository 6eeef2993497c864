// Probe: FP64 FMA latency and throughput per SM (decides how the Cholesky of G is written, whiten.cu).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, int n) {
    double a = threadIdx.x * 1e-9 + 1.0, b = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fma(a, b, c);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = a; }
}
__global__ void thr(double* out, int n) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-9 + k;
    const double b = 1.0000001, c = 1e-7;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
    long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n / 8; out[1] = s; }
}
__global__ void lat32(double* out, int n) {
    float a = threadIdx.x * 1e-9f + 1.0f, b = 1.0000001f, c = 1e-7f;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = fmaf(a, b, c);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = a; }
}
int main() {
    double *d, h[2]; cudaMalloc(&d, 16);
    for (int threads : {32, 128, 256, 1024}) {
        lat<<<1, threads>>>(d, 4096); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("threads %4d: dependent DFMA %.1f cycles each", threads, h[0]);
        thr<<<1, threads>>>(d, 2048); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf(" | 8 independent chains: %.2f cycles per DFMA per thread (=> %.1f DFMA lanes/cycle/SM)", h[0], threads / h[0]);
        lat32<<<1, threads>>>(d, 4096); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf(" | dependent FFMA %.1f\n", h[0]);
    }
    return 0;
}
