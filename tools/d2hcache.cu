// measurement aid: what makes a D2H copy into a pinned buffer slow?  (found while chasing a 0.75 ms copy of 10 MB in the activity path)
//  - NOT the buffer's age as a DMA target (48 buffers in turn), not small against huge pages, not a source just written by a kernel;
//  - the destination's lines sitting in the caches of several CPU cores: 12 threads that have read the buffer make the next
//    D2H 6-10x slower; clflushopt of the lines after reading them removes the penalty at no measurable cost.
// build: nvcc -O2 -arch=sm_100a -diag-suppress 1650 -Xcompiler -mclflushopt tools/d2hcache.cu -o tools/d2hcache.bin
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <thread>

static float d2h(void* h, const void* d, size_t n, cudaStream_t s, cudaEvent_t e0, cudaEvent_t e1)
{
    cudaEventRecord(e0, s); cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
static void* thp_pinned(size_t bytes)
{
    const size_t two = (size_t)2 << 20; bytes = (bytes + two - 1) / two * two;
    void* p = 0; if (posix_memalign(&p, two, bytes)) return 0;
    madvise(p, bytes, MADV_HUGEPAGE);
    memset(p, 0, bytes);
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) { free(p); return 0; }
    return p;
}
__global__ void fill(float* p, size_t n, float v) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v; }
#include <immintrin.h>
#include <chrono>
static double read_flush(const void* h, size_t n, int nthreads, int mode)   /* mode 0: read, 1: read then clflushopt per line, 2: read + clflushopt interleaved */
{
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=]() {
            const char* base = (const char*)h; size_t a = n / 64 * t / nthreads * 64, b = n / 64 * (t + 1) / nthreads * 64;
            unsigned long long acc = 0;
            if (mode == 2) {
                for (size_t i = a; i < b; i += 64) { const unsigned long long* q = (const unsigned long long*)(base + i); for (int k = 0; k < 8; k++) acc += q[k]; _mm_clflushopt((void*)(base + i)); }
            } else {
                for (size_t i = a; i < b; i += 8) acc += *(const unsigned long long*)(base + i);
                if (mode == 1) for (size_t i = a; i < b; i += 64) _mm_clflushopt((void*)(base + i));
            }
            _mm_sfence();
            static volatile unsigned long long sink; sink = acc;
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
static void read_all(const void* h, size_t n, int nthreads, bool whole)
{
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([=]() {
            const unsigned long long* q = (const unsigned long long*)h; size_t a = 0, b = n / 8;
            if (!whole) { a = n / 8 * t / nthreads; b = n / 8 * (t + 1) / nthreads; }
            unsigned long long acc = 0; for (size_t i = a; i < b; i++) acc += q[i];
            static volatile unsigned long long sink; sink = acc;
        });
    for (auto& x : th) x.join();
}
int main()
{
    const size_t n = 10246656;             /* one cfg3 call's burst samples */
    void* d; cudaMalloc(&d, 64 << 20);
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r"); char line[128] = "?"; if (f) { fgets(line, sizeof(line), f); fclose(f); }
    printf("transparent_hugepage: %s", line);
    {
        void* h; cudaHostAlloc(&h, n, cudaHostAllocDefault); memset(h, 0, n);
        for (int i = 0; i < 3; i++) d2h(h, d, n, s, e0, e1);
        float t0 = d2h(h, d, n, s, e0, e1);
        fill<<<296, 256, 0, s>>>((float*)d, n / 4, 1.0f); float t1 = d2h(h, d, n, s, e0, e1);
        read_all(h, n, 12, false); float t2 = d2h(h, d, n, s, e0, e1);
        read_all(h, n, 12, true); float t3 = d2h(h, d, n, s, e0, e1);
        read_all(h, n, 12, false); fill<<<296, 256, 0, s>>>((float*)d, n / 4, 2.0f); float t4 = d2h(h, d, n, s, e0, e1);
        printf("D2H %.1f MB: warm %.3f ms | source just written by a kernel %.3f | destination read by 12 threads (slices) %.3f | (all of it each) %.3f | both %.3f\n",
               n / 1e6, t0, t1, t2, t3, t4);
        for (int mode = 0; mode < 3; mode++)
            for (int rep = 0; rep < 3; rep++) {
                const double c = read_flush(h, n, 12, mode); const float t = d2h(h, d, n, s, e0, e1);
                printf("12 threads %s: CPU %.3f ms, then D2H %.3f ms\n", mode == 0 ? "read slices" : mode == 1 ? "read slices, then clflushopt them" : "read + clflushopt line by line", c, t);
            }
        cudaFreeHost(h);
    }
    for (int kind = 0; kind < 2; kind++) {
        const int NB = 48;
        std::vector<void*> bufs(NB);
        for (int i = 0; i < NB; i++) {
            if (kind == 0) { if (cudaHostAlloc(&bufs[i], n, cudaHostAllocDefault) != cudaSuccess) return 1; memset(bufs[i], 0, n); }
            else { bufs[i] = thp_pinned(n); if (!bufs[i]) { printf("huge-page pinned allocation failed\n"); return 1; } }
        }
        const char* name = kind == 0 ? "cudaHostAlloc" : "2 MiB aligned + MADV_HUGEPAGE + cudaHostRegister";
        for (int i = 0; i < 3; i++) d2h(bufs[0], d, n, s, e0, e1);
        float same = 1e9f; for (int i = 0; i < 5; i++) { float t = d2h(bufs[0], d, n, s, e0, e1); if (t < same) same = t; }
        /* round robin over 48 buffers (480 MB, 120 k small pages): every copy goes to a buffer that was a DMA target 47 copies ago */
        float sum = 0; int cnt = 0;
        for (int round = 0; round < 3; round++)
            for (int i = 0; i < NB; i++) { float t = d2h(bufs[i], d, n, s, e0, e1); if (round > 0) { sum += t; cnt++; } }
        printf("%-52s: %.1f MB D2H into the buffer just used %.3f ms, into one of 48 buffers in turn %.3f ms\n", name, n / 1e6, same, sum / cnt);
        for (int i = 0; i < NB; i++) { if (kind == 0) cudaFreeHost(bufs[i]); else { cudaHostUnregister(bufs[i]); free(bufs[i]); } }
    }
    return 0;
}
