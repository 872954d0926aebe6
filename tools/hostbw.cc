/* tools/hostbw.cc -- measurement aid: host memory copy bandwidth vs thread count, plain memcpy and non-temporal stores
 * (sizes the staging copy pool of the host path).  g++ -O2 -pthread -o hostbw.bin hostbw.cc */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

__attribute__((target("avx2"))) static void nt_copy(void* d, const void* s, size_t n)
{
    char* dp = (char*)d; const char* sp = (const char*)s;
    while (n && ((uintptr_t)dp & 31)) { *dp++ = *sp++; n--; }
    size_t v = n / 32;
    for (size_t i = 0; i < v; i++) _mm256_stream_si256((__m256i*)dp + i, _mm256_loadu_si256((const __m256i*)sp + i));
    _mm_sfence();
    memcpy(dp + v * 32, sp + v * 32, n - v * 32);
}
int main()
{
    const size_t bytes = 1ull << 30;
    char* a = (char*)aligned_alloc(4096, bytes); char* b = (char*)aligned_alloc(4096, bytes);
    memset(a, 1, bytes); memset(b, 2, bytes);
    for (int mode = 0; mode < 2; mode++)
        for (int nt : {1, 2, 3, 4, 6, 8, 12, 16}) {
            const size_t piece = 256u << 10;
            auto t0 = std::chrono::steady_clock::now();
            for (int rep = 0; rep < 3; rep++) {
                std::vector<std::thread> th;
                for (int t = 0; t < nt; t++)
                    th.emplace_back([=] {
                        for (size_t off = (size_t)t * piece; off < bytes; off += (size_t)nt * piece) {
                            if (mode) nt_copy(b + off, a + off, piece); else memcpy(b + off, a + off, piece);
                        }
                    });
                for (auto& x : th) x.join();
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%s %2d threads: %.1f GB/s copied\n", mode ? "non-temporal" : "memcpy      ", nt, 3.0 * bytes / dt / 1e9);
        }
    return 0;
}
