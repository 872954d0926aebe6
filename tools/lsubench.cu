/* tools/lsubench.cu -- measurement aid: is the load/store path of an SM bound by BYTES or by INSTRUCTIONS?  Global loads that hit
 * L1 / L2 and global stores, 64-bit against 128-bit per lane, same bytes per thread.  nvcc -O3 -arch=sm_100a -o lsubench.bin */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

/* every CTA streams over its own window of `win` bytes (16 KiB: L1 resident, 4 MiB: L2 resident) */
template <int VEC, bool STORE> __global__ void __launch_bounds__(256) k_ls(char* buf, size_t win, int iters, float* sink)
{
    char* base = buf + (size_t)blockIdx.x * win;
    const size_t per_iter = 256 * 64;                       /* bytes a CTA touches per inner step: 64 B per thread */
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        const size_t off = ((size_t)it * per_iter) % win;
        if (VEC == 8) {
            float2* p = reinterpret_cast<float2*>(base + off);
#pragma unroll
            for (int k = 0; k < 8; k++) {                   /* 8 x 8 B per thread, a warp covers 256 contiguous bytes per k */
                if (STORE) p[k * 256 + threadIdx.x] = make_float2(acc, (float)k);
                else { const float2 v = __ldg(p + k * 256 + threadIdx.x); acc += v.x + v.y; }
            }
        } else {
            float4* p = reinterpret_cast<float4*>(base + off);
#pragma unroll
            for (int k = 0; k < 4; k++) {                   /* 4 x 16 B per thread, a warp covers 512 contiguous bytes per k */
                if (STORE) p[k * 256 + threadIdx.x] = make_float4(acc, (float)k, 1.f, 2.f);
                else { const float4 v = __ldg(p + k * 256 + threadIdx.x); acc += v.x + v.y + v.z + v.w; }
            }
        }
    }
    if (acc == 1234.5f) *sink = acc;
}
template <int VEC, bool STORE> static void run(const char* name, char* buf, size_t win, int ctas_per_sm, int nsm, float* sink)
{
    const int iters = 2000;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_ls<VEC, STORE><<<nsm * ctas_per_sm, 256>>>(buf, win, 50, sink); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a)); k_ls<VEC, STORE><<<nsm * ctas_per_sm, 256>>>(buf, win, iters, sink); CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
    const double bytes = (double)iters * 256 * 64 * nsm * ctas_per_sm;
    printf("%-40s window %7zu B, %d CTAs/SM: %7.0f GB/s = %5.1f B/clk/SM\n", name, win, ctas_per_sm, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / nsm / 1.92e9);
}
int main()
{
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nsm = pr.multiProcessorCount;
    char* buf; float* sink; CK(cudaMalloc(&buf, (size_t)nsm * 8 * (4u << 20) / 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 0, (size_t)nsm * 8 * (4u << 20) / 4));
    for (int c : {2, 4, 8}) {
        run<8, false>("LDG.64  (L1 resident)", buf, 16u << 10, c, nsm, sink);
        run<16, false>("LDG.128 (L1 resident)", buf, 16u << 10, c, nsm, sink);
        run<8, false>("LDG.64  (L2 resident)", buf, 1u << 20, c, nsm, sink);
        run<16, false>("LDG.128 (L2 resident)", buf, 1u << 20, c, nsm, sink);
        run<8, true>("STG.64  (L2 resident)", buf, 1u << 20, c, nsm, sink);
        run<16, true>("STG.128 (L2 resident)", buf, 1u << 20, c, nsm, sink);
    }
    return 0;
}
