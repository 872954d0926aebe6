/* tools/peerstore.cu -- measurement aid: bandwidth of SM-issued stores into a PEER GPU's memory over NVLink as a function of
 * the store width and of how many bytes a warp writes contiguously (the extract kernel's sink stores are 8 B per lane,
 * 128 contiguous bytes per half warp).  One process, two GPUs, peer access enabled.  nvcc -O3 -arch=sm_100a -o peerstore.bin */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

/* every warp writes runs of RUN bytes (RUN = 32 lanes x VEC bytes or 16 lanes x VEC bytes), runs of one warp 4 KiB apart */
template <int VEC, int LANES> __global__ void k_store(char* dst, size_t bytes, int iters)
{
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int run = LANES * VEC;                               /* contiguous bytes per group of LANES lanes */
    const size_t nruns = bytes / run;
    for (int it = 0; it < iters; it++) {
        for (size_t r = (size_t)warp * (32 / LANES) + lane / LANES; r < nruns; r += (size_t)nwarps * (32 / LANES)) {
            /* scatter the runs so that consecutive runs of a warp are not adjacent (rows of different channels) */
            const size_t rr = (r * 2654435761ull) % nruns;
            char* p = dst + rr * run + (lane % LANES) * VEC;
            if (VEC == 8) *reinterpret_cast<float2*>(p) = make_float2((float)it, (float)r);
            else *reinterpret_cast<float4*>(p) = make_float4((float)it, (float)r, 1.f, 2.f);
        }
    }
}
template <int VEC, int LANES> static void run(const char* name, char* dst, size_t bytes, int ctas)
{
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_store<VEC, LANES><<<ctas, 256>>>(dst, bytes, 1); CK(cudaDeviceSynchronize());
    const int iters = 8;
    CK(cudaEventRecord(a)); k_store<VEC, LANES><<<ctas, 256>>>(dst, bytes, iters); CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-44s %4d CTAs: %.0f GB/s\n", name, ctas, (double)bytes * iters / (ms * 1e-3) / 1e9);
}
int main()
{
    int n = 0; CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs two GPUs\n"); return 0; }
    const size_t bytes = 1ull << 29;
    char *local, *peer;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&local, bytes)); CK(cudaDeviceEnablePeerAccess(1, 0));
    for (int ctas : {148, 592, 1184}) {
        run<8, 16>("local  8 B/lane, 128 B runs", local, bytes, ctas);
        run<8, 16>("peer   8 B/lane, 128 B runs", peer, bytes, ctas);
        run<8, 32>("peer   8 B/lane, 256 B runs", peer, bytes, ctas);
        run<16, 16>("peer  16 B/lane, 256 B runs", peer, bytes, ctas);
        run<16, 32>("peer  16 B/lane, 512 B runs", peer, bytes, ctas);
    }
    return 0;
}
