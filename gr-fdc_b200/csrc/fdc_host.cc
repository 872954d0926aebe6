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
struct CopyPool::Impl {
    struct Piece { char* d; const char* s; size_t n; };
    std::vector<std::thread> th;
    std::mutex m; std::condition_variable cv, idle;
    std::deque<Piece> q; size_t inflight; bool stop;
    Impl() : inflight(0), stop(false) {}
    bool run_one(std::unique_lock<std::mutex>& lk)
    {
        if (q.empty()) return false;
        const Piece p = q.front(); q.pop_front();
        lk.unlock();
        memcpy(p.d, p.s, p.n);
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
            run_one(lk);
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
void CopyPool::submit(void* dst, const void* src, size_t bytes)
{
    if (!bytes) return;
    const size_t piece = 512u << 10;
    if (d->th.empty() || bytes <= piece / 2) { memcpy(dst, src, bytes); return; }       /* small copies are not worth a hand-over */
    {
        std::lock_guard<std::mutex> g(d->m);
        for (size_t off = 0; off < bytes; off += piece) {
            Impl::Piece p; p.d = (char*)dst + off; p.s = (const char*)src + off; p.n = std::min(piece, bytes - off);
            d->q.push_back(p); d->inflight++;
        }
    }
    d->cv.notify_all();
}
void CopyPool::wait()
{
    std::unique_lock<std::mutex> lk(d->m);
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
        else { const unsigned hc = std::thread::hardware_concurrency(); n = hc > 4 ? (int)std::min(12u, hc - 2) - 1 : 1; }
        if (n < 0) n = 0;
        pool = new CopyPool(n);
    });
    return *pool;
}

}  // namespace fdc
