// per-node latency of a CUDA graph made of dependent kernel nodes (the panel chain of the camera solve)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void node(int *p, int work)
{
    extern __shared__ double sm[];
    if (threadIdx.x == 0 && blockIdx.x == 0) { int v = p[0]; for (int i = 0; i < work; ++i) v = v * 3 + 1; p[0] = v; }
}
int main()
{
    int *p; cudaMalloc(&p, 4); cudaMemset(p, 0, 4);
    cudaStream_t s; cudaStreamCreate(&s);
    for (int smem : {0, 37632, 75264}) for (int grid : {1, 137}) {
        cudaFuncSetAttribute(node, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int k = 0; k < 250; ++k) node<<<grid, 128, smem, s>>>(p, 0);
        cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
        cudaEventRecord(e0, s);
        for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, s);
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("graph of 250 dependent nodes, grid %3d x 128 thr, %5d B dyn smem: %.2f us per node\n", grid, smem, ms * 1e3 / 2500);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
