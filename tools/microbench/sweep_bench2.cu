// Ablation of the ROW-OWNER panel sweep (kernels_chol.cu): cycles per column step.
#include <cstdio>
#include <cuda_runtime.h>
#define TS 48
template <int V>
__global__ void __launch_bounds__(128) sweep(double *out, long long *cyc, const double *in)
{
    __shared__ __align__(16) double colbuf[2][2 * TS + 4];
    const int tid = threadIdx.x;
    const bool active = tid <= 2 * TS;
    double row[TS];
#pragma unroll
    for (int c = 0; c < TS; ++c) row[c] = ((tid % TS) == c ? 100.0 : 0.0) + in[(tid * 48 + c) % 97] * 0.01;
    bool bad = false;
    if (active) colbuf[0][tid] = row[0];
    long long t0 = clock64();
#pragma unroll
    for (int j = 0; j < TS; ++j) {
        __syncthreads();
        if (active) {
            const double *cb = colbuf[j & 1];
            double piv = 100.0;
            if (V != 3) { piv = cb[j]; bad |= !(piv > 0.0 && piv < 1e300); }
            double f;
            if (V == 2 || V == 3) f = row[j] * 0.01; else if (V == 4) f = row[j] / piv; else f = row[j] * __drcp_rn(piv);
            if (j + 1 < TS) { row[j + 1] -= f * cb[j + 1]; colbuf[(j + 1) & 1][tid] = row[j + 1]; }
            if (V == 0 || V == 4 || V == 5) {
#pragma unroll
                for (int c = j + 2; c < TS; ++c) row[c] -= f * cb[c];
            }
            if (V != 5) row[j] *= rsqrt(piv);
        }
    }
    long long t1 = clock64();
    double s = bad;
#pragma unroll
    for (int c = 0; c < TS; ++c) s += row[c];
    out[tid] = s;
    if (tid == 0) cyc[V] = t1 - t0;
}
int main()
{
    double *out, *in; long long *cyc, h[8];
    cudaMalloc(&out, 4096); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 64);
    double hin[97]; for (int i = 0; i < 97; ++i) hin[i] = (i * 37 % 101) / 101.0;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        sweep<0><<<1, 128>>>(out, cyc, in); sweep<1><<<1, 128>>>(out, cyc, in); sweep<2><<<1, 128>>>(out, cyc, in);
        sweep<3><<<1, 128>>>(out, cyc, in); sweep<4><<<1, 128>>>(out, cyc, in); sweep<5><<<1, 128>>>(out, cyc, in);
    }
    cudaMemcpy(h, cyc, 48, cudaMemcpyDeviceToHost);
    const char *names[] = {"full", "no bulk update (chain only)", "chain only, constant reciprocal", "chain only, constant reciprocal, no pivot load",
                           "full with IEEE division instead of __drcp_rn", "full without the final rsqrt scaling"};
    for (int l = 0; l < 6; ++l) printf("variant %d (%s): %.0f cycles per column step\n", l, names[l], h[l] / 48.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
