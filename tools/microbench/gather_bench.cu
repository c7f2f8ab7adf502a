// gather_bench.cu -- how fast can a B200 gather 144-byte blocks (one W block of the pair pass) from an array much larger than L2?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
// Variants: lane-per-block with 16-byte loads (9 LDG.128), lane-per-block with 32-byte loads (4 LDG.256 + 1 LDG.128), nine lanes
// per block (coalesced 16-byte pieces), and a sequential stream for reference; occupancy varied through the block count.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

template <int MODE, int ILP>
__global__ void k_gather(const double *__restrict__ W, const int *__restrict__ idx, long long n, double *__restrict__ out)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    double acc = 0.0;
    if (MODE == 2) {                                   // nine lanes per block: lane -> (block slot, 16-byte piece)
        const int lane = threadIdx.x & 31, slot = lane / 9, piece = lane % 9;
        const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = stride >> 5;
        for (long long t = wid * 3; t < n; t += nw * 3) {
            if (slot < 3 && t + slot < n) {
                const double2 v = __ldg(reinterpret_cast<const double2 *>(W + (size_t)idx[t + slot] * 18) + piece);
                acc += v.x + v.y;
            }
        }
    } else {
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride * ILP) {
#pragma unroll
            for (int u = 0; u < ILP; ++u) {
                const long long tt = t + u * stride;
                if (tt >= n) break;
                const double *p = W + (size_t)(MODE == 3 ? tt : idx[tt]) * 18;
                if (MODE == 0 || MODE == 3) {
#pragma unroll
                    for (int q = 0; q < 9; ++q) { const double2 v = __ldg(reinterpret_cast<const double2 *>(p) + q); acc += v.x + v.y; }
                } else {
                    const bool odd = (reinterpret_cast<unsigned long long>(p) & 16ull) != 0;
                    const double *q = odd ? p + 2 : p;
                    double a, b, c, d;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { ldg256(q + 4 * j, a, b, c, d); acc += a + b + c + d; }
                    const double2 e = __ldg(reinterpret_cast<const double2 *>(odd ? p : p + 16));
                    acc += e.x + e.y;
                }
            }
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

int main()
{
    const long long nblk = 5000000, ngather = 10000000;          // 720 MB of blocks, 10 M gathers (the off-diagonal triples of the headline workload)
    double *W, *out; int *idx;
    CK(cudaMalloc(&W, nblk * 144)); CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&idx, ngather * 4));
    CK(cudaMemset(W, 0, nblk * 144));
    int *h = (int *)malloc(ngather * 4);
    unsigned long long s = 88172645463325252ull;
    for (long long i = 0; i < ngather; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % nblk); }
    CK(cudaMemcpy(idx, h, ngather * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char *names[4] = {"lane per block, 9 x LDG.128", "lane per block, 4 x LDG.256 + LDG.128", "nine lanes per block (coalesced pieces)", "sequential stream, 9 x LDG.128"};
    printf("%-42s %8s %6s %10s %10s\n", "variant", "warps/SM", "ILP", "ms", "GB/s");
    for (int mode = 0; mode < 4; ++mode)
        for (int wps : {8, 16, 32, 64})
            for (int ilp : {1, 2, 4}) {
                if (mode == 2 && ilp > 1) continue;
                const int grid = 148 * wps / 8;                  // 256-thread blocks
                float best = 1e9;
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaEventRecord(e0));
#define RUN(M, I) k_gather<M, I><<<grid, 256>>>(W, idx, ngather, out)
                    if (mode == 0) { if (ilp == 1) RUN(0, 1); else if (ilp == 2) RUN(0, 2); else RUN(0, 4); }
                    if (mode == 1) { if (ilp == 1) RUN(1, 1); else if (ilp == 2) RUN(1, 2); else RUN(1, 4); }
                    if (mode == 2) RUN(2, 1);
                    if (mode == 3) { if (ilp == 1) RUN(3, 1); else if (ilp == 2) RUN(3, 2); else RUN(3, 4); }
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                printf("%-42s %8d %6d %10.3f %10.1f\n", names[mode], wps, ilp, best, ngather * 144.0 / (best * 1e6));
            }
    return 0;
}
