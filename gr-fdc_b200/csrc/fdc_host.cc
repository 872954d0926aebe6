/* fdc_host.cc -- host-side geometry and window-table construction (no CUDA in this file).
 * These are restatements of the reference's setup code; all arithmetic keeps the reference's types and
 * evaluation order because bin indices and table values are compared bit for bit.
 * Compile with -ffp-contract=off. */
#include "fdc_host.h"
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <cmath>
#include <complex>
#include <stdexcept>
#include <string>

namespace fdc {

/* python/FrequencyDomainChannelizer.py:37-40 */
static long py_nextpow2(double k)
{
    if (k < 1) throw std::invalid_argument("Cannot evaluate next power 2 of " + std::to_string(k));
    return 1L << (long)std::ceil(std::log2(k));
}

/* python/FrequencyDomainChannelizer.py:322-345, Python 2 semantics (GNU Radio 3.7): int/int floors,
 * round() rounds half away from zero. */
void opt_channelparams(int blocksize, int relinvovl, double freq, double bw, int* f, int* l, int* lout, double* passband,
                       double* stopband)
{
    const double passsamps = (double)blocksize * bw;
    long blocklen = py_nextpow2(passsamps);
    if ((double)blocklen < 1.2 * passsamps) blocklen *= 2;
    double pb = passsamps / (double)blocklen * 1.1, sb = 1.0;
    if (pb >= 1.0) pb = 1.0;
    else if (pb < 0.7) sb = pb + 0.25;
    long freqsamps = (long)std::round(freq * (double)blocksize) % blocksize;
    if (freqsamps < 0) freqsamps += blocksize;                 /* Python % is non-negative */
    freqsamps -= blocklen / 2;
    if (freqsamps < 0) freqsamps = (freqsamps + blocksize) % blocksize;
    if (freqsamps + blocklen > blocksize) freqsamps = blocksize - blocklen;
    *f = (int)freqsamps; *l = (int)blocklen; *lout = (int)(blocklen - blocklen / relinvovl);
    *passband = pb; *stopband = sb;
}

/* lib/windows.h:80-124 : the real mask in double precision */
static void mask_shape(int wintype, int n, int lowsamps, int rampsamps, bool normalize, std::vector<double>& w)
{
    const double v = normalize ? 1.0 : 1.0 / (double)n;
    w.assign((size_t)n, v);
    if (wintype == 1 || wintype == 2) {                         /* HANN, RAMP: zero edge, shaped transition */
        for (int i = 0; i < lowsamps; i++) { w[i] = 0.0; w[n - 1 - i] = 0.0; }
        for (int i = 0; i < rampsamps; i++) {
            double a;
            if (wintype == 2) a = v * (double)(i + 1) / (double)(rampsamps + 1);
            else { const double phi = (double)(i + 1) / (double)(rampsamps + 1) * M_PI; a = v * (-cos(phi) / 2.0 + 0.5); }
            w[lowsamps + i] = a;
            w[n - lowsamps - 1 - i] = w[lowsamps + i];
        }
    } else {                                                    /* RECTANGULAR: edge + half the transition zeroed */
        for (int i = 0; i < lowsamps + rampsamps / 2; i++) { w[i] = 0.0; w[n - 1 - i] = 0.0; }
    }
}

/* lib/windows.h:41-78 with step = 1, normalize = false, as called from lib/phase_shifting_windowing_vcc_impl.cc:62 */
void psw_tables(int blocksize, int relinvovl, float passbw, float stopbw, int wintype, std::vector<std::complex<float> >& out)
{
    if (passbw >= 1.0) { passbw = 1.0; stopbw = 1.0; wintype = 0; }
    else if (stopbw >= 1.0) stopbw = 1.0;
    const int lowsamps = (int)((1.0 - stopbw) * (double)blocksize) / 2;
    const int highsamps = (int)(passbw * (double)blocksize);
    const int rampsamps = (blocksize - 2 * lowsamps - highsamps) / 2;
    std::vector<double> w;
    mask_shape(wintype, blocksize, lowsamps, rampsamps, false, w);
    out.resize((size_t)relinvovl * blocksize);
    int count = 0;
    for (int i = 0; i < relinvovl; i++) {
        const double phi = 2.0 * M_PI * (double)count / (double)relinvovl;
        for (int k = 0; k < blocksize; k++)
            out[(size_t)i * blocksize + k] = (std::complex<float>)std::polar(w[k], phi);
        count = (count + 1) % relinvovl;
    }
}

void psw_check_args(float passbw, float stopbw)
{
    /* messages of lib/phase_shifting_windowing_vcc_impl.cc:46-53 */
    if (passbw <= 0.0) throw std::invalid_argument("PassBw in phase_compensating_windowing_vcc_block must not be < 0");
    if (stopbw <= 0.0) throw std::invalid_argument("StopBw in phase_compensating_windowing_vcc_block must not be < 0");
    if (stopbw < passbw) throw std::invalid_argument("StopBw must not be < PassBw in phase_compensating_windowing_vcc block");
}

int nextpow2_int(double v) { return 1 << (int)std::ceil(std::log2(v)); }

/* 10^(dB/10), lib/PowerActivationChannel_impl.cc:377-381, lib/SegmentDetection_impl.cc:84 */
float db_to_ratio(float db) { return (float)std::pow(10.0, (double)db / 10.0); }


/* ---- copy pool ---------------------------------------------------------------------------------- */
/* Streaming copy: non-temporal stores.  The destination is either a pinned slot the DMA engine reads next or the caller's
 * buffer, never something this thread reads back, and a plain memcpy of such a buffer first reads the destination lines it
 * is about to overwrite (measured on the GPU box, tools/hostbw.cc: 6.2 GB/s per thread with memcpy, 12.5 GB/s with
 * non-temporal stores; 45 against 72 GB/s on all 16 cores). */
/* ---- copies out of a pinned buffer the GPU writes into (D2H destination) ------------------------------------------------
 * A DMA write into lines that several CPU cores hold in their caches is slow: measured on the GPU box (Xeon, 16 cores, one
 * socket) a 10 MB D2H copy takes 0.185 ms into a buffer nobody has read and 1.1-2.0 ms after 12 threads have read it
 * (tools/d2hcache.cu, profiles/r2_d2h_after_cpu_reads.txt).  Flushing the lines after they have been read (clflushopt: they
 * are clean, nothing is written back) restores 0.185 ms at no measurable cost to the copy.  Every copy the library makes
 * out of its own D2H destinations therefore evicts what it has read. */
#if defined(__x86_64__)
#include <immintrin.h>
#include <cpuid.h>
static bool have_clflushopt()
{
    static const bool have = [] { unsigned a, b, c, d; return __get_cpuid_count(7, 0, &a, &b, &c, &d) && (b & (1u << 23)); }();
    return have;
}
void evict_lines(const void* p, size_t bytes)
{
    if (!bytes) return;
    const uintptr_t a = (uintptr_t)p & ~(uintptr_t)63, e = (uintptr_t)p + bytes;
    if (have_clflushopt()) for (uintptr_t q = a; q < e; q += 64) __asm__ volatile("clflushopt %0" : "+m"(*(volatile char*)q));
    else for (uintptr_t q = a; q < e; q += 64) _mm_clflush((const void*)q);
    _mm_sfence();
}
__attribute__((target("avx2"))) static void stream_copy_avx2(char* dp, const char* sp, size_t n)
{
    while (n && ((uintptr_t)dp & 31)) { *dp++ = *sp++; n--; }
    const size_t v = n / 32;
    for (size_t i = 0; i < v; i++) _mm256_stream_si256((__m256i*)dp + i, _mm256_loadu_si256((const __m256i*)sp + i));
    _mm_sfence();
    memcpy(dp + v * 32, sp + v * 32, n - v * 32);
}
static void stream_copy(void* d, const void* s, size_t n, bool evict_src)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && n >= 4096) stream_copy_avx2((char*)d, (const char*)s, n); else memcpy(d, s, n);
    if (evict_src) evict_lines(s, n);
}
#else
void evict_lines(const void*, size_t) {}
static void stream_copy(void* d, const void* s, size_t n, bool) { memcpy(d, s, n); }
#endif
void copy_and_evict(void* dst, const void* src, size_t bytes) { memcpy(dst, src, bytes); evict_lines(src, bytes); }

struct CopyPool::Impl {
    /* one contiguous copy, or (many != 0) a run of small copies described by arrays the submitter keeps alive until wait() */
    struct Piece { char* d; const char* s; size_t n; void* const* md; const void* const* ms; const size_t* mb; size_t many; bool evict; };
    std::vector<std::thread> th;
    std::mutex m; std::condition_variable cv, idle;
    std::deque<Piece> q; size_t inflight; bool stop;
    Impl() : inflight(0), stop(false) {}
    bool run_one(std::unique_lock<std::mutex>& lk)
    {
        if (q.empty()) return false;
        const Piece p = q.front(); q.pop_front();
        lk.unlock();
        if (p.many) for (size_t i = 0; i < p.many; i++) stream_copy(p.md[i], p.ms[i], p.mb[i], p.evict);
        else stream_copy(p.d, p.s, p.n, p.evict);
        lk.lock();
        if (--inflight == 0) idle.notify_all();
        return true;
    }
    void worker()
    {
        std::unique_lock<std::mutex> lk(m);
        while (true) {
            cv.wait(lk, [&] { return stop || !q.empty(); });
            if (stop && q.empty()) return;
            while (run_one(lk)) {}
        }
    }
};
CopyPool::CopyPool(int threads) : d(new Impl)
{
    for (int i = 0; i < threads; i++) d->th.emplace_back([this] { d->worker(); });
}
CopyPool::~CopyPool()
{
    { std::lock_guard<std::mutex> g(d->m); d->stop = true; }
    d->cv.notify_all();
    for (size_t i = 0; i < d->th.size(); i++) d->th[i].join();
    delete d;
}
int CopyPool::threads() const { return (int)d->th.size(); }
/* queues the copy in pieces; nothing runs before wait() (one wake-up per batch, not per copy) */
void CopyPool::submit(void* dst, const void* src, size_t bytes, bool evict_src)
{
    if (!bytes) return;
    const size_t piece = 256u << 10;
    std::lock_guard<std::mutex> g(d->m);
    for (size_t off = 0; off < bytes; off += piece) {
        Impl::Piece p; p.d = (char*)dst + off; p.s = (const char*)src + off; p.n = std::min(piece, bytes - off);
        p.md = 0; p.ms = 0; p.mb = 0; p.many = 0; p.evict = evict_src;
        d->q.push_back(p); d->inflight++;
    }
}
void CopyPool::submit_many(void* const* dst, const void* const* src, const size_t* bytes, size_t n, bool evict_src)
{
    const size_t target = 256u << 10;
    std::lock_guard<std::mutex> g(d->m);
    size_t i = 0;
    while (i < n) {
        size_t j = i, acc = 0;
        while (j < n && (acc < target || j == i)) acc += bytes[j++];
        Impl::Piece p; p.d = 0; p.s = 0; p.n = 0; p.md = dst + i; p.ms = src + i; p.mb = bytes + i; p.many = j - i; p.evict = evict_src;
        d->q.push_back(p); d->inflight++;
        i = j;
    }
}
void CopyPool::wait()
{
    std::unique_lock<std::mutex> lk(d->m);
    if (d->q.size() > 1 && !d->th.empty()) { lk.unlock(); d->cv.notify_all(); lk.lock(); }
    while (d->run_one(lk)) {}                               /* the caller copies too */
    d->idle.wait(lk, [&] { return d->inflight == 0; });
}
CopyPool& copy_pool()
{
    /* leaked on purpose: worker threads must not be joined from a static destructor at interpreter shutdown */
    static CopyPool* pool = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        int n = 0;
        const char* v = getenv("FDC_COPY_THREADS");
        if (v && *v) n = atoi(v);
        else { const unsigned hc = std::thread::hardware_concurrency(); n = hc >= 4 ? (int)(hc * 3 / 4) : 1; }     /* + the caller; measured on the 16-core GPU box: 11 workers 2.8 Gsample/s, 15 workers 3.05 (pinned caller memory: 3.2) */
        if (n < 0) n = 0;
        pool = new CopyPool(n);
    });
    return *pool;
}

}  // namespace fdc
