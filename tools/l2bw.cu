/* tools/l2bw.cu -- measurement aid: achievable L2 / HBM bandwidth of simple float4 read, write and copy kernels
 * as a function of the working-set size (the K1 -> K2 hand-over lives in L2; this gives the second roofline). */
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const float4* __restrict__ p, size_t n, float* sink)
{
    float4 a = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = p[i]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (a.x + a.y + a.z + a.w == 123.456f) *sink = a.x;
}
__global__ void k_copy(const float4* __restrict__ p, float4* __restrict__ q, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) q[i] = p[i];
}
__global__ void k_write(float4* __restrict__ q, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) q[i] = make_float4(1, 2, 3, 4);
}
int main()
{
    float4 *a, *b; float* sink;
    const size_t maxb = 1ull << 30;
    cudaMalloc(&a, maxb); cudaMalloc(&b, maxb); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, maxb); cudaMemset(b, 0, maxb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, thr = 256;
    for (size_t mb : {4, 8, 16, 32, 48, 64, 96, 128, 256, 1024}) {
        const size_t bytes = mb << 20, n = bytes / 16;
        const int reps = (int)((8ull << 30) / bytes) + 1;
        float ms[3];
        for (int mode = 0; mode < 3; mode++) {
            for (int w = 0; w < 3; w++) { if (mode == 0) k_read<<<grid, thr>>>(a, n, sink); else if (mode == 1) k_copy<<<grid, thr>>>(a, b, n / 2); else k_write<<<grid, thr>>>(b, n); }
            cudaEventRecord(e0);
            for (int r = 0; r < reps; r++) { if (mode == 0) k_read<<<grid, thr>>>(a, n, sink); else if (mode == 1) k_copy<<<grid, thr>>>(a, b, n / 2); else k_write<<<grid, thr>>>(b, n); }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms[mode], e0, e1); ms[mode] /= reps;
        }
        printf("working set %5zu MiB: read %7.0f GB/s   copy(r+w, half each) %7.0f GB/s   write %7.0f GB/s\n", mb, bytes / ms[0] / 1e6,
               bytes / ms[1] / 1e6, bytes / ms[2] / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
