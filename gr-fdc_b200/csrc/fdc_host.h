/* fdc_host.h -- host-side geometry / table builders (see fdc_host.cc). */
#ifndef FDC_HOST_H
#define FDC_HOST_H
#include <complex>
#include <vector>
namespace fdc {
void opt_channelparams(int blocksize, int relinvovl, double freq, double bw, int* f, int* l, int* lout, double* passband,
                       double* stopband);
void psw_tables(int blocksize, int relinvovl, float passbw, float stopbw, int wintype, std::vector<std::complex<float> >& out);
void psw_check_args(float passbw, float stopbw);
int nextpow2_int(double v);
float db_to_ratio(float db);

/* Copy pool of the host path: a few worker threads that move caller memory <-> the library's pinned staging slots in parallel
 * (one thread copies 10-15 GB/s, PCIe moves 25 + 50 GB/s).  submit() cuts a copy into pieces, wait() returns when every
 * submitted piece is done; the calling thread works through the queue as well.  One pool per process. */
class CopyPool {
public:
    explicit CopyPool(int threads);
    ~CopyPool();
    /* evict_src: the source is a pinned buffer the GPU writes into again later -- flush the lines that were read (evict_lines) */
    void submit(void* dst, const void* src, size_t bytes, bool evict_src = false);
    /* n small copies dst[i] <- src[i] (bytes[i] each) as a few tasks of about 256 KiB: a channelizer with thousands of
     * narrow channels hands every chunk back as thousands of few-KiB rows */
    void submit_many(void* const* dst, const void* const* src, const size_t* bytes, size_t n, bool evict_src = false);
    void wait();
    int threads() const;
private:
    struct Impl; Impl* d;
    CopyPool(const CopyPool&); CopyPool& operator=(const CopyPool&);
};
CopyPool& copy_pool();
/* flush [p, p + bytes) from the CPU caches (clflushopt; the lines are clean after a read): a later D2H copy into them runs at
 * PCIe speed instead of waiting for the cores' copies to be invalidated one by one (fdc_host.cc, measured 0.185 against 1.1-2.0 ms
 * per 10 MB) */
void evict_lines(const void* p, size_t bytes);
void copy_and_evict(void* dst, const void* src, size_t bytes);
}
#endif
