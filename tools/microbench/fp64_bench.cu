// FP64 latency / throughput microbenchmark for B200 (roofline denominators for the camera solve).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double *out, long long *cyc, double a, double b, int n)
{
    double x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rsqrt(double *out, long long *cyc, double a, int n)
{
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = rsqrt(x) + a;
    long long t1 = clock64();
    out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void lat_bar(long long *cyc, int n)
{
    __shared__ double s[256];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { s[threadIdx.x] = i; __syncthreads(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0 + (long long)s[1] * 0;
}
template <int ILP>
__global__ void thr(double *out, double a, double b, int n)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = a + k;
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], b, a);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; long long *cyc, h[3];
    cudaMalloc(&out, 1 << 26); cudaMalloc(&cyc, 64);
    lat<<<1, 32>>>(out, cyc, 1.0, 0.999, 4096);
    lat_rsqrt<<<1, 32>>>(out, cyc, 1.5, 1024);
    lat_bar<<<1, 256>>>(cyc, 1024);
    cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %.1f cycles\n", h[0] / 4096.0);
    printf("rsqrt(double)+DADD dependent latency: %.1f cycles\n", h[1] / 1024.0);
    printf("STS + __syncthreads (256 threads): %.1f cycles\n", h[2] / 1024.0);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int warps = 4; warps <= 32; warps *= 2) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int n = 20000, blocks = p.multiProcessorCount;
        thr<8><<<blocks, warps * 32>>>(out, 1.0, 0.999, 100);
        cudaEventRecord(e0);
        thr<8><<<blocks, warps * 32>>>(out, 1.0, 0.999, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * n * (double)blocks * warps * 32;
        printf("FP64 FMA throughput, %2d warps/SM x ILP 8: %.2f TFLOP/s (%.1f FMA/clk/SM at %d MHz)\n", warps, flops / ms / 1e9,
               flops / 2 / (ms * 1e-3) / blocks / (p.clockRate * 1e3), p.clockRate / 1000);
    }
    return 0;
}
