// Microbenchmark: read bandwidth as a function of working-set size (L2 vs HBM) on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2bw tools/l2bw.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const float4 *p, size_t n, float *sink, int reps)
{
    float acc = 0.f;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = __ldcg(p + i);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *sink = acc;
}
__global__ void cp(const float4 *p, float4 *o, size_t n, int reps)
{
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            o[i] = __ldcg(p + i);
}
int main()
{
    float *buf, *buf2, *sink;
    size_t maxb = 1ull << 30;
    cudaMalloc(&buf, maxb); cudaMalloc(&buf2, maxb); cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t sizes[] = {4, 8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 512, 1024};
    for (size_t mb : sizes) {
        size_t bytes = mb << 20, n = bytes / 16;
        int reps = (int)((8ull << 30) / bytes); if (reps < 2) reps = 2;
        for (int blocks : {148 * 4, 148 * 8}) {
            rd<<<blocks, 512>>>((float4 *)buf, n, sink, 2);
            cudaEventRecord(e0);
            rd<<<blocks, 512>>>((float4 *)buf, n, sink, reps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("read  %5zu MB blocks %4d: %8.1f GB/s\n", mb, blocks, (double)bytes * reps / ms / 1e6);
        }
        int creps = reps / 2 + 1;
        cp<<<148 * 8, 512>>>((float4 *)buf, (float4 *)buf2, n, 2);
        cudaEventRecord(e0);
        cp<<<148 * 8, 512>>>((float4 *)buf, (float4 *)buf2, n, creps);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("copy  %5zu MB (r+w bytes)   : %8.1f GB/s\n", mb, 2.0 * bytes * creps / ms / 1e6);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
