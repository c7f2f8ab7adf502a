// Ablation microbenchmark of the panel column sweep (kernels_chol.cu:k_panel): cycles per column step
// for cumulative subsets of the work, one CTA of 128 threads (6x3 register blocks of D and A).
#include <cstdio>
#include <cuda_runtime.h>
#define TS 48
#define CBS (2 * TS + 2)
template <int LEVEL>
__global__ void __launch_bounds__(128) sweep(double *out, long long *cyc, const double *in)
{
    __shared__ __align__(16) double colbuf[2][CBS];
    __shared__ double B2[TS * (TS + 2)];
    const int tid = threadIdx.x, tr = tid % 8, tc = tid / 8;
    double d[6][3], a[6][3], bq[3] = {0.1, 0.2, 0.3};
    for (int p = 0; p < 6; ++p) for (int q = 0; q < 3; ++q) {
        const int r = tr * 6 + p, c = tc * 3 + q;
        d[p][q] = (r == c ? 100.0 : 0.0) + in[(r * 48 + c) % 97] * 0.01;
        a[p][q] = in[(r * 7 + c) % 97] * 0.01;
    }
    if (tc == 0) { for (int p = 0; p < 6; ++p) { colbuf[0][tr * 6 + p] = d[p][0]; colbuf[0][TS + tr * 6 + p] = a[p][0]; } colbuf[0][2 * TS] = bq[0]; }
    const double *pld = colbuf[0] + tr * 6, *pli = colbuf[0] + TS + tr * 6, *plc = colbuf[0] + tc * 3;
    double lrow_p[6] = {0}, ld_p[6] = {0}, lc_p[3] = {0}, lb_p = 0, yk = 0;
    bool bad = false;
    long long t0 = clock64();
#pragma unroll 1
    for (int jb = 0; jb < 16; ++jb) {
#pragma unroll
        for (int JQ = 0; JQ < 3; ++JQ) {
            const int QN = (JQ + 1) % 3, QP = JQ, Q3 = 3 - QP - QN;
            const int j = jb * 3 + JQ, par = j & 1;
            __syncthreads();
            double rs = 1.0, piv = 1.0;
            if (LEVEL >= 1 && LEVEL != 6) { piv = colbuf[par][j]; bad |= !(piv > 0.0 && piv < 1e300); rs = (LEVEL == 8) ? 1.0 / piv : rsqrt(piv); } if (LEVEL == 6) rs = 0.1;
            double lrow[6], ld[6], lc[3], lb = 0;
            if (LEVEL >= 2) {
                const double2 *vd = reinterpret_cast<const double2 *>(pld + par * CBS), *vr = reinterpret_cast<const double2 *>(pli + par * CBS);
#pragma unroll
                for (int p = 0; p < 3; ++p) { double2 x = vd[p], y = vr[p]; ld[2 * p] = x.x * rs; ld[2 * p + 1] = x.y * rs; lrow[2 * p] = y.x * rs; lrow[2 * p + 1] = y.y * rs; }
#pragma unroll
                for (int q = 0; q < 3; ++q) lc[q] = plc[par * CBS + q] * rs;
                lb = colbuf[par][2 * TS] * rs;
            } else { for (int p = 0; p < 6; ++p) { ld[p] = rs; lrow[p] = rs; } for (int q = 0; q < 3; ++q) lc[q] = rs; }
            if (LEVEL >= 3 && LEVEL != 7) {
#pragma unroll
                for (int p = 0; p < 6; ++p) { d[p][QN] -= ld_p[p] * lc_p[QN]; a[p][QN] -= lrow_p[p] * lc_p[QN]; }
                bq[QN] -= lb_p * lc_p[QN];
#pragma unroll
                for (int p = 0; p < 6; ++p) { d[p][QN] -= ld[p] * lc[QN]; a[p][QN] -= lrow[p] * lc[QN]; }
                bq[QN] -= lb * lc[QN];
            }
            if (tc == (JQ == 2 ? jb + 1 : jb)) {
                double *cbn = colbuf[par ^ 1];
                double2 *cd = reinterpret_cast<double2 *>(cbn + tr * 6), *ca = reinterpret_cast<double2 *>(cbn + TS + tr * 6);
#pragma unroll
                for (int p = 0; p < 3; ++p) { cd[p] = make_double2(d[2 * p][QN], d[2 * p + 1][QN]); ca[p] = make_double2(a[2 * p][QN], a[2 * p + 1][QN]); }
                if (tr == 0) cbn[2 * TS] = bq[QN];
            }
            if (LEVEL == 4 || LEVEL == 5) {
#pragma unroll
                for (int p = 0; p < 6; ++p) { d[p][Q3] -= ld_p[p] * lc_p[Q3]; a[p][Q3] -= lrow_p[p] * lc_p[Q3]; }
                bq[Q3] -= lb_p * lc_p[Q3];
            }
            if (LEVEL == 5 && tc == jb) {
                double2 *po = reinterpret_cast<double2 *>(B2 + j * (TS + 2) + tr * 6);
#pragma unroll
                for (int p = 0; p < 3; ++p) po[p] = make_double2(lrow[2 * p], lrow[2 * p + 1]);
                if (tr == 0) yk += lb;
            }
#pragma unroll
            for (int p = 0; p < 6; ++p) { ld_p[p] = ld[p]; lrow_p[p] = lrow[p]; }
#pragma unroll
            for (int q = 0; q < 3; ++q) lc_p[q] = lc[q];
            lb_p = lb;
        }
    }
    long long t1 = clock64();
    double s = yk + (bad ? 1 : 0);
    for (int p = 0; p < 6; ++p) for (int q = 0; q < 3; ++q) s += d[p][q] + a[p][q];
    out[tid] = s + B2[tid] + bq[0];
    if (tid == 0) cyc[LEVEL] = t1 - t0;
}
int main()
{
    double *out, *in; long long *cyc, h[16];
    cudaMalloc(&out, 4096); cudaMalloc(&in, 4096); cudaMalloc(&cyc, 128);
    double hin[97]; for (int i = 0; i < 97; ++i) hin[i] = (i * 37 % 101) / 101.0;
    cudaMemcpy(in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        sweep<0><<<1, 128>>>(out, cyc, in); sweep<1><<<1, 128>>>(out, cyc, in); sweep<2><<<1, 128>>>(out, cyc, in);
        sweep<3><<<1, 128>>>(out, cyc, in); sweep<4><<<1, 128>>>(out, cyc, in); sweep<5><<<1, 128>>>(out, cyc, in); sweep<6><<<1, 128>>>(out, cyc, in); sweep<7><<<1, 128>>>(out, cyc, in); sweep<8><<<1, 128>>>(out, cyc, in);
    }
    cudaMemcpy(h, cyc, 72, cudaMemcpyDeviceToHost);
    const char *names[] = {"barrier + publish only", "+ pivot rsqrt", "+ column loads and scaling", "+ critical next-column update", "+ deferred bulk update", "+ factor column store", "level 3 with constant rs (no pivot read, no rsqrt)", "level 2 + publish (no update)", "level 3 with 1/piv instead of rsqrt"};
    for (int l = 0; l < 9; ++l) printf("level %d (%s): %.0f cycles per column step\n", l, names[l], h[l] / 48.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
