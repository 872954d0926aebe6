/* fdc_cabi.cu -- C ABI (include/fdc_cabi.h): library entry points, the fused throughput channelizer context and
 * the copy/multiply/FFT block replacements.  The activity-gated blocks live in fdc_cabi_act.cu. */
#include "fdc_cabi_internal.h"
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <map>
#include <string>

using namespace fdc;

/* ---- error plumbing ---------------------------------------------------------------------------- */
static thread_local std::string g_err;
namespace fdc {
void set_error(const std::string& s) { g_err = s; }
int fail(const std::string& s) { g_err = s; return -1; }
int cuda_fail(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
}
bool require_device()
{
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_err = std::string("fdc_b200: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "count 0") +
                "); this library has no CPU fallback";
        cudaGetLastError();
        return false;
    }
    return true;
}
}  // namespace fdc

extern "C" {

int fdc_api_version(void) { return FDC_API_VERSION; }
const char* fdc_last_error(void) { return g_err.c_str(); }
int fdc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int fdc_set_device(int device)
{
    if (!require_device()) return -1;
    const cudaError_t e = cudaSetDevice(device);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaSetDevice");
}
void* fdc_host_alloc(size_t bytes)
{
    void* p = 0;
    if (!require_device()) return 0;
    const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cuda_fail(e, "cudaHostAlloc"); return 0; }
    return p;
}
void fdc_host_free(void* p) { if (p) cudaFreeHost(p); }
int fdc_copy_threads(void) { return copy_pool().threads(); }
int fdc_host_register(void* p, size_t bytes)
{
    if (!require_device()) return -1;
    if (!p || !bytes) return fail("fdc_host_register: bad arguments");
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaHostRegister");
}
int fdc_host_unregister(void* p)
{
    const cudaError_t e = cudaHostUnregister(p);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaHostUnregister");
}
void fdc_host_evict(const void* p, size_t bytes) { if (p) evict_lines(p, bytes); }
unsigned long long fdc_launch_count(void) { return launch_count(); }
void* fdc_dev_alloc(size_t bytes)
{
    void* p = 0;
    if (!require_device()) return 0;
    const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc"); return 0; }
    return p;
}
void fdc_dev_free(void* p) { if (p) cudaFree(p); }
int fdc_memcpy_h2d(void* dst, const void* src, size_t bytes)
{
    const cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? 0 : cuda_fail(e, "fdc_memcpy_h2d");
}
int fdc_memcpy_d2h(void* dst, const void* src, size_t bytes)
{
    const cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? 0 : cuda_fail(e, "fdc_memcpy_d2h");
}
int fdc_device_synchronize(void)
{
    const cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : cuda_fail(e, "fdc_device_synchronize");
}

int fdc_opt_channelparams(int blocksize, int relinvovl, double freq, double bw, int* f, int* l, int* lout, double* passband,
                          double* stopband)
{
    try { opt_channelparams(blocksize, relinvovl, freq, bw, f, l, lout, passband, stopband); return 0; }
    catch (const std::exception& e) { return fail(e.what()); }
}
int fdc_psw_build_tables(int blocklen, int numphasestates, float passbw, float stopbw, int windowtype, float* out)
{
    try {
        if (blocklen <= 0 || numphasestates <= 0) return fail("blocklen and numphasestates must be > 0");
        psw_check_args(passbw, stopbw);
        std::vector<std::complex<float> > t;
        psw_tables(blocklen, numphasestates, passbw, stopbw, windowtype, t);
        memcpy(out, t.data(), sizeof(std::complex<float>) * t.size());
        return 0;
    } catch (const std::exception& e) { return fail(e.what()); }
}

}  // extern "C"

/* ================================================================================================
 * fused throughput channelizer
 * ================================================================================================ */
struct fdc_chan {
    int dev, N, ovl, hop, nphase, nchan;
    bool big; int N1, N2;
    bool fused_small;                  /* one kernel does K1 + K2 (fdc_k_chanfused.cu): N <= 16384, every channel of the same slice length */
    std::vector<ChanDev> chans;
    std::vector<int> l;
    std::vector<std::pair<int, std::pair<int, int> > > groups;    /* (l, (first index in d_chans, count)) */
    std::vector<int> group_even_f;     /* every slice of the group starts on an even bin (16-byte aligned: TMA bulk copies) */
    long lout_total;
    long blockcount;
    long chunk_blocks;                 /* blocks per K1->K2 round trip: spectrum ring sized to stay in L2 */
    DevBuf d_chans, d_tables, d_hist, d_hist2;
    std::vector<int> launch_order;     /* d_chans[k] describes channel launch_order[k] */
    int nsinks;
    /* channel-sharded sinks by DMA (FDC_SINK_DMA=1, default): the extract kernel writes a chunk's rows into a local staging slab,
     * the copy engines move every owner's part over NVLink while the next chunks are transformed */
    std::vector<void*> sink_bases; std::vector<int> sink_first, sink_count, sink_local;     /* owner k holds channels [first, first + count) */
    std::vector<long> sink_prefix_total;                                                    /* items per block of owner k */
    DevBuf w_out[4]; cudaStream_t cs[4]; cudaEvent_t ev_x[4], ev_c[4]; bool copies_pending[4];
    const float2* tw4;                 /* four-step twiddles (big N) */
    cudaStream_t stream;
    /* device path: chunks alternate between NWORK worker streams, each with its own spectrum / intermediate ring, so
     * that the tail of one chunk's kernels overlaps the head of the next chunk's */
    enum { NWORK = 4 };
    cudaStream_t ws[NWORK];
    DevBuf w_spec[NWORK], w_mid[NWORK];
    cudaEvent_t ev_start, ev_done[NWORK];
    cudaEvent_t ev_hist; bool hist_pending; cudaStream_t hist_stream;     /* the history buffer is written asynchronously at the end of a device call */
    /* host path: NSLOT pipelined chunk slots */
    enum { NSLOT = 4 };
    cudaStream_t hs[NSLOT];
    DevBuf h_in[NSLOT], h_out[NSLOT], h_spec[NSLOT], h_mid[NSLOT];
    PinBuf p_in[NSLOT], p_out[NSLOT];      /* library-owned pinned staging for pageable caller memory */
    std::vector<void*> drain_dst; std::vector<const void*> drain_src; std::vector<size_t> drain_bytes;
    cudaEvent_t h_done[NSLOT];
    long host_chunk;
    /* optional per-kernel timing (fdc_chan_set_profiling): events around K1 and K2 of every chunk */
    bool prof;
    std::vector<cudaEvent_t> prof_ev;      /* triples: before K1, after K1, after K2 */
    std::vector<cudaEvent_t> prof_pool;
    fdc_chan() : nsinks(0), tw4(0), stream(0), ev_start(0), ev_hist(0), hist_pending(false), hist_stream(0), host_chunk(0), prof(false)
    {
        for (int i = 0; i < NSLOT; i++) { hs[i] = 0; h_done[i] = 0; }
        for (int i = 0; i < NWORK; i++) { ws[i] = 0; ev_done[i] = 0; cs[i] = 0; ev_x[i] = 0; ev_c[i] = 0; copies_pending[i] = false; }
    }
    cudaEvent_t ev()
    {
        cudaEvent_t e = 0;
        if (!prof_pool.empty()) { e = prof_pool.back(); prof_pool.pop_back(); }
        else cudaEventCreate(&e);
        prof_ev.push_back(e);
        return e;
    }
};

static long pick_chunk_blocks(int N)
{
    /* spectrum ring of about 64 MiB per worker stream: with the (equally large) four-step intermediate the
     * K1 -> K2 hand-over mostly stays inside the 126 MB L2; never fewer than one wave of CTAs. */
    long c = (64L << 20) / ((long)N * 8);
    if (c < 8) c = 8;
    return c;
}

/* enqueue K1 + K2 for nb blocks; block b reads d_in[b*hop - ovl, b*hop + hop) */
static int chan_enqueue_chunk(fdc_chan* c, const float2* d_in, long nb, float2* d_spec, float2* d_mid,
                              float2* d_out, long call_blocks, long call_blk0, long glob_blk0, cudaStream_t s,
                              const float2* d_hist = 0, long head_blocks = 0, const ExtractParams::Sink* sinks = 0, int nsinks = 0, long head_off = 0,
                              bool need_spec = true)
{
    cudaError_t e;
    if (c->prof) cudaEventRecord(c->ev(), s);
    /* N <= 16384, one slice length, nobody wants the spectrum itself: K1 + K2 in one kernel, the spectrum stays in shared memory */
    if (c->fused_small && !need_spec && (d_out || nsinks) && tuning().fuse_small) {
        FwdParams p; p.in = d_in; p.spec = 0; p.nblocks = nb; p.hop = c->hop; p.ovl = c->ovl; p.N = c->N;
        p.scale = 1.0f / (float)c->N; p.l2pf = tuning().l2pf ? 1 : 0; p.hist = d_hist; p.head_blocks = d_hist ? head_blocks : 0; p.head_off = head_off;
        ExtractParams q; q.spec = 0; q.spec_stride = c->N; q.tables = (const float2*)c->d_tables.p;
        q.chans = (const ChanDev*)c->d_chans.p; q.nsel = c->nchan; q.ny = 0; q.out = d_out; q.tma_ok = 0; q.l2pf = 0; q.bpt = 1; q.nsinks = nsinks;
        for (int k = 0; k < nsinks; k++) q.sink[k] = sinks[k];
        q.nb = nb; q.call_blocks = call_blocks; q.call_blk0 = call_blk0; q.glob_phase0 = (int)(glob_blk0 % c->nphase); q.nphase = c->nphase;
        q.phase_mask = (c->nphase & (c->nphase - 1)) == 0 ? c->nphase - 1 : -1;
        e = launch_chan_fused(p, q, c->groups[0].first, s);
        if (e != cudaSuccess) return cuda_fail(e, "fused channelizer launch");
        if (c->prof) { cudaEventRecord(c->ev(), s); cudaEventRecord(c->ev(), s); }
        return 0;
    }
    if (!c->big) {
        FwdParams p; p.in = d_in; p.spec = d_spec; p.nblocks = nb; p.hop = c->hop; p.ovl = c->ovl; p.N = c->N;
        p.scale = 1.0f / (float)c->N; p.l2pf = tuning().l2pf ? 1 : 0; p.hist = d_hist; p.head_blocks = d_hist ? head_blocks : 0; p.head_off = head_off;
        e = launch_fwd_small(p, s);
    } else {
        BigParams p; p.in = d_in; p.mid = d_mid; p.spec = d_spec; p.tw4 = c->tw4;
        p.nblocks = nb; p.hop = c->hop; p.ovl = c->ovl; p.scale = 1.0f / (float)c->N; p.hist = d_hist; p.head_blocks = d_hist ? head_blocks : 0; p.head_off = head_off;
        e = (tuning().fused && fwd_cluster_supported(c->N)) ? launch_fwd_cluster(p, c->N, s) : launch_fwd_big(p, c->N, s);
    }
    if (e != cudaSuccess) return cuda_fail(e, "forward FFT launch");
    if (c->prof) cudaEventRecord(c->ev(), s);
    if (!d_out && !nsinks) { if (c->prof) cudaEventRecord(c->ev(), s); return 0; }
    for (size_t g = 0; g < c->groups.size(); g++) {
        ExtractParams q; q.spec = d_spec; q.spec_stride = c->N; q.tables = (const float2*)c->d_tables.p;
        q.chans = (const ChanDev*)c->d_chans.p + c->groups[g].second.first;
        q.nsel = c->groups[g].second.second; q.ny = 0; q.out = d_out;
        q.tma_ok = ((uintptr_t)d_spec % 16 == 0) && (c->N % 2 == 0) && c->group_even_f[g];
        q.l2pf = (tuning().l2pf && q.tma_ok) ? 1 : 0; q.bpt = 1; q.nsinks = nsinks;
        for (int k = 0; k < nsinks; k++) q.sink[k] = sinks[k];
        q.nb = nb; q.call_blocks = call_blocks; q.call_blk0 = call_blk0; q.glob_phase0 = (int)(glob_blk0 % c->nphase); q.nphase = c->nphase;
        q.phase_mask = (c->nphase & (c->nphase - 1)) == 0 ? c->nphase - 1 : -1;
        e = launch_extract(q, c->groups[g].first, s);
        if (e != cudaSuccess) return cuda_fail(e, "channel extract launch");
    }
    if (c->prof) cudaEventRecord(c->ev(), s);
    return 0;
}

extern "C" {

fdc_chan* fdc_chan_create(int N, int ovl, int nphase, int nchan, const fdc_chan_desc* ch)
{
    if (!require_device()) return 0;
    if (N < 16 || (N & (N - 1))) { fail("fdc_chan_create: N must be a power of two >= 16"); return 0; }
    if (ovl < 0 || ovl >= N) { fail("fdc_chan_create: need 0 <= ovl < N"); return 0; }
    if (nphase < 1) { fail("fdc_chan_create: nphase must be >= 1"); return 0; }
    if (nchan < 0 || (nchan > 0 && !ch)) { fail("fdc_chan_create: bad channel list"); return 0; }
    fdc_chan* c = new fdc_chan;
    cudaGetDevice(&c->dev);
    c->N = N; c->ovl = ovl; c->hop = N - ovl; c->nphase = nphase; c->nchan = nchan; c->blockcount = 0;
    c->big = false; c->N1 = c->N2 = 0;
    /* one CTA per block up to tuning().fwd_split (exclusive), the two-kernel four-step scheme from there on */
    if (!fwd_small_supported(N) || (N >= tuning().fwd_split && fwd_big_supported(N, 0, 0))) {
        if (!fwd_big_supported(N, &c->N1, &c->N2)) { fail("fdc_chan_create: unsupported FFT length"); delete c; return 0; }
        c->big = true;
        c->tw4 = fourstep_table(c->N1, c->N2);
        if (!c->tw4) { fail("fdc_chan_create: four-step twiddle table allocation failed"); delete c; return 0; }
    }
    /* channels: validate, group by l, pack tables */
    std::vector<float2> tables;
    std::map<int, std::vector<int> > by_l;
    std::map<std::string, long> seen;
    long prefix = 0;
    for (int i = 0; i < nchan; i++) {
        const fdc_chan_desc& d = ch[i];
        if (!tile_len_supported(d.l)) { fail("fdc_chan_create: channel slice length must be a power of two in [2, 16384]"); delete c; return 0; }
        if (d.f < 0 || d.f + d.l > N) { fail("fdc_chan_create: channel slice [f, f+l) outside the spectrum"); delete c; return 0; }
        if (d.lout < 1 || d.lout > d.l) { fail("fdc_chan_create: need 1 <= lout <= l"); delete c; return 0; }
        if (!d.table) { fail("fdc_chan_create: channel table missing"); delete c; return 0; }
        ChanDev cd; memset(&cd, 0, sizeof(cd));
        cd.f = d.f; cd.lout = d.lout; cd.shift = ((d.shift % nphase) + nphase) % nphase;
        cd.lout_prefix = prefix; cd.gain = d.gain;
        prefix += d.lout;
        /* channels whose tables are bit-identical share one copy (equal-bandwidth channel plans have a single table,
         * which then lives in L1/L2 instead of being streamed once per channel) */
        const size_t tbytes = sizeof(float2) * (size_t)nphase * d.l;
        /* a power-of-two gain (the hier block uses l) commutes exactly with every rounding of the chain, so it is folded
         * into the device copy of the table and the kernel stores without a multiply */
        int gexp = 0;
        const bool fold = d.gain > 0.0f && std::frexp(d.gain, &gexp) == 0.5f && gexp > -60 && gexp < 60;
        if (fold) cd.gain = 1.0f;
        std::string key((const char*)d.table, tbytes);
        key.append((const char*)&gexp, fold ? sizeof(gexp) : 0);
        std::map<std::string, long>::iterator hit = seen.find(key);
        if (hit != seen.end()) cd.tab_off = hit->second;
        else {
            cd.tab_off = (long)tables.size();
            seen[key] = cd.tab_off;
            const float2* t = (const float2*)d.table;
            tables.insert(tables.end(), t, t + (size_t)nphase * d.l);
            if (fold) for (size_t k = tables.size() - (size_t)nphase * d.l; k < tables.size(); k++) { tables[k].x *= d.gain; tables[k].y *= d.gain; }
        }
        c->chans.push_back(cd); c->l.push_back(d.l); by_l[d.l].push_back(i);
    }
    c->lout_total = prefix;
    std::vector<int> sel;
    for (std::map<int, std::vector<int> >::iterator it = by_l.begin(); it != by_l.end(); ++it) {
        c->groups.push_back(std::make_pair(it->first, std::make_pair((int)sel.size(), (int)it->second.size())));
        /* neighbours in frequency share a CTA tile: their slices overlap in the spectrum */
        int even = 1;
        for (size_t k = 0; k < it->second.size(); k++) if (c->chans[(size_t)it->second[k]].f & 1) even = 0;
        c->group_even_f.push_back(even);
        std::vector<int> order(it->second);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c->chans[a].f < c->chans[b].f; });
        sel.insert(sel.end(), order.begin(), order.end());
    }
    c->fused_small = !c->big && c->groups.size() == 1 && chan_fused_supported(N, c->groups[0].first);
    c->chunk_blocks = pick_chunk_blocks(N);
    bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->ev_hist, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < fdc_chan::NWORK && ok; i++)
        ok = cudaStreamCreateWithFlags(&c->ws[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
    /* device copy of the channel descriptors in launch order (grouped by l, ascending f) */
    c->launch_order = sel;
    std::vector<ChanDev> ordered(sel.size());
    for (size_t i = 0; i < sel.size(); i++) ordered[i] = c->chans[(size_t)sel[i]];
    ok = ok && c->d_chans.upload(ordered.data(), sizeof(ChanDev) * ordered.size());
    ok = ok && c->d_tables.upload(tables.data(), sizeof(float2) * tables.size());
    ok = ok && c->d_hist.reserve(sizeof(float2) * (size_t)std::max(ovl, 1)) && c->d_hist2.reserve(sizeof(float2) * (size_t)std::max(ovl, 1));
    ok = ok && cudaMemset(c->d_hist.p, 0, sizeof(float2) * (size_t)std::max(ovl, 1)) == cudaSuccess;
    /* warm the twiddle caches so no allocation happens inside a timed region */
    if (ok) {
        if (c->big) { twiddle_table(c->N1); twiddle_table(c->N2); } else twiddle_table(N);
        for (size_t g = 0; g < c->groups.size(); g++) twiddle_table(c->groups[g].first);
    }
    if (!ok) { cuda_fail(cudaGetLastError(), "fdc_chan_create: device allocation"); fdc_chan_destroy(c); return 0; }
    return c;
}

void fdc_chan_destroy(fdc_chan* c)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return;
    cudaDeviceSynchronize();
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_hist) cudaEventDestroy(c->ev_hist);
    for (int i = 0; i < fdc_chan::NWORK; i++) {
        if (c->ws[i]) cudaStreamDestroy(c->ws[i]);
        if (c->ev_done[i]) cudaEventDestroy(c->ev_done[i]);
        if (c->cs[i]) cudaStreamDestroy(c->cs[i]);
        if (c->ev_x[i]) cudaEventDestroy(c->ev_x[i]);
        if (c->ev_c[i]) cudaEventDestroy(c->ev_c[i]);
    }
    for (size_t i = 0; i < c->prof_ev.size(); i++) cudaEventDestroy(c->prof_ev[i]);
    for (size_t i = 0; i < c->prof_pool.size(); i++) cudaEventDestroy(c->prof_pool[i]);
    for (int i = 0; i < fdc_chan::NSLOT; i++) { if (c->hs[i]) cudaStreamDestroy(c->hs[i]); if (c->h_done[i]) cudaEventDestroy(c->h_done[i]); }
    delete c;
}
int fdc_chan_hop(const fdc_chan* c) { return c ? c->hop : -1; }
int fdc_chan_is_fused(const fdc_chan* c) { return (c && c->fused_small && tuning().fuse_small) ? 1 : 0; }
long fdc_chan_blockcount(const fdc_chan* c) { return c ? c->blockcount : -1; }
int fdc_chan_reset(fdc_chan* c)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    cudaDeviceSynchronize();
    c->blockcount = 0;
    const cudaError_t e = cudaMemset(c->d_hist.p, 0, sizeof(float2) * (size_t)std::max(c->ovl, 1));
    return e == cudaSuccess ? 0 : cuda_fail(e, "fdc_chan_reset");
}
int fdc_chan_seek(fdc_chan* c, long first_block)
{
    if (!c || first_block < 0) return fail("fdc_chan_seek: bad arguments");
    c->blockcount = first_block;
    return 0;
}
int fdc_chan_set_history(fdc_chan* c, const void* host)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c || !host) return fail("fdc_chan_set_history: bad arguments");
    if (c->ovl == 0) return 0;
    cudaDeviceSynchronize();
    const cudaError_t e = cudaMemcpy(c->d_hist.p, host, sizeof(float2) * (size_t)c->ovl, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? 0 : cuda_fail(e, "fdc_chan_set_history");
}

/* history for the next call = the last ovl samples of [old history | this call's input] */
static int chan_save_history(fdc_chan* c, const float2* d_in, long nblocks, cudaStream_t s)
{
    if (c->ovl == 0 || nblocks == 0) return 0;
    const long n_new = nblocks * c->hop;
    cudaError_t e;
    if (n_new >= c->ovl) {
        /* into the spare buffer: the kernels of this call read d_hist, and they are ordered before this copy only on `s` */
        e = cudaMemcpyAsync(c->d_hist2.p, d_in + (n_new - c->ovl), sizeof(float2) * (size_t)c->ovl, cudaMemcpyDeviceToDevice, s);
        c->d_hist.swap(c->d_hist2);
    } else {
        float2* h = (float2*)c->d_hist.p; float2* h2 = (float2*)c->d_hist2.p;
        e = cudaMemcpyAsync(h2, h + n_new, sizeof(float2) * (size_t)(c->ovl - n_new), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(h2 + (c->ovl - n_new), d_in, sizeof(float2) * (size_t)n_new, cudaMemcpyDeviceToDevice, s);
        c->d_hist.swap(c->d_hist2);
    }
    return e == cudaSuccess ? 0 : cuda_fail(e, "history save");
}

int fdc_chan_work_device(fdc_chan* c, const void* d_in_v, long nblocks, void* d_out_v, void* d_spectrum_v, void* stream)
{
    return fdc_chan_work_device_slab(c, d_in_v, nblocks, d_out_v, nblocks, 0, d_spectrum_v, stream);
}

static int chan_work_device(fdc_chan* c, const void* d_in_v, long nblocks, void* d_out_v, long slab_blocks, long slab_first_block,
                            void* d_spectrum_v, void* stream, bool use_sinks);

int fdc_chan_work_device_slab(fdc_chan* c, const void* d_in_v, long nblocks, void* d_out_v, long slab_blocks, long slab_first_block,
                              void* d_spectrum_v, void* stream)
{
    return chan_work_device(c, d_in_v, nblocks, d_out_v, slab_blocks, slab_first_block, d_spectrum_v, stream, false);
}

/* Channel-sharded sinks (one process per GPU; FDC/sharded.py ChannelSinks): sink k is the output buffer of the GPU that owns
 * the channels with owner[c] == k, laid out channel-major like the output of fdc_chan_work_device_slab but holding only those
 * channels.  The extract kernel of EVERY rank stores each channel's rows straight into its owner's buffer -- its own memory
 * or peer memory mapped with fdc_ipc_open -- so the exchange is an all-to-all of peer stores that overlaps the butterflies:
 * every GPU receives 1/world of every rank's output instead of one GPU receiving everything. */
int fdc_chan_set_sinks(fdc_chan* c, int nsinks, void* const* d_bases, const int* owner, int local_sink)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    if (nsinks < 1 || nsinks > FDC_MAX_SINKS || !d_bases || !owner) return fail("fdc_chan_set_sinks: need 1 .. 16 sinks");
    std::vector<long> run((size_t)nsinks, 0);
    for (int i = 0; i < c->nchan; i++) {
        if (owner[i] < 0 || owner[i] >= nsinks) return fail("fdc_chan_set_sinks: channel owner out of range");
        c->chans[(size_t)i].owner = owner[i];
        c->chans[(size_t)i].sink_prefix = run[(size_t)owner[i]];
        run[(size_t)owner[i]] += c->chans[(size_t)i].lout;
    }
    for (int k = 0; k < nsinks; k++) if (run[(size_t)k] && !d_bases[k]) return fail("fdc_chan_set_sinks: a sink that owns channels has no buffer");
    c->sink_bases.assign(d_bases, d_bases + nsinks);
    c->sink_first.assign((size_t)nsinks, 0); c->sink_count.assign((size_t)nsinks, 0);
    bool contiguous = true;
    for (int i = 0; i < c->nchan; i++) {
        const int k = owner[i];
        if (c->sink_count[(size_t)k] == 0) c->sink_first[(size_t)k] = i;
        else if (c->sink_first[(size_t)k] + c->sink_count[(size_t)k] != i) contiguous = false;
        c->sink_count[(size_t)k]++;
    }
    if (!contiguous) c->sink_first.clear();            /* scattered ownership: the kernel stores into the sinks itself */
    cudaDeviceSynchronize();
    std::vector<ChanDev> ordered(c->launch_order.size());
    for (size_t i = 0; i < ordered.size(); i++) ordered[i] = c->chans[(size_t)c->launch_order[i]];
    if (!c->d_chans.upload(ordered.data(), sizeof(ChanDev) * ordered.size())) return cuda_fail(cudaGetLastError(), "fdc_chan_set_sinks");
    c->sink_prefix_total = run;
    /* the sink in this GPU's own memory is stored into directly by the kernel; the others are peer memory */
    c->sink_local.assign((size_t)nsinks, 0);
    if (local_sink >= 0 && local_sink < nsinks) c->sink_local[(size_t)local_sink] = 1;
    c->nsinks = nsinks;
    return 0;
}
int fdc_chan_work_device_sinks(fdc_chan* c, const void* d_in, long nblocks, long slab_blocks, long slab_first_block, void* stream)
{
    if (c && c->nsinks < 1) return fail("fdc_chan_work_device_sinks: call fdc_chan_set_sinks first");
    return chan_work_device(c, d_in, nblocks, 0, slab_blocks, slab_first_block, 0, stream, true);
}

static int chan_work_device(fdc_chan* c, const void* d_in_v, long nblocks, void* d_out_v, long slab_blocks, long slab_first_block,
                            void* d_spectrum_v, void* stream, bool use_sinks)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    if (nblocks < 0) return fail("nblocks < 0");
    if (slab_first_block < 0 || slab_first_block + nblocks > slab_blocks) return fail("fdc_chan_work_device_slab: blocks outside the slab");
    if (nblocks == 0) return 0;
    const float2* d_in = (const float2*)d_in_v; float2* d_out = (float2*)d_out_v; float2* d_spectrum = (float2*)d_spectrum_v;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    cudaError_t e = cudaSuccess;
    /* The first nh blocks reach back into the previous call: their loads take the samples before d_in[0] from the history
     * buffer (two-segment loaders, fdc_functors.cuh; lib/overlap_save_impl.cc:70-78 keeps the same history).  No staging
     * copy, no separate launch for them. */
    const long nh = c->ovl ? std::min(nblocks, ((long)c->ovl + c->hop - 1) / c->hop) : 0;
    /* the history may still be being written by an earlier call on another stream */
    if (c->hist_pending && c->hist_stream != s) {
        if ((e = cudaStreamWaitEvent(s, c->ev_hist, 0)) != cudaSuccess) return cuda_fail(e, "stream wait");
    }
    /* worker streams (profiling serialises everything on the caller's stream so that the per-kernel events are clean) */
    int nw = tuning().streams < 1 ? 1 : (tuning().streams > (int)fdc_chan::NWORK ? (int)fdc_chan::NWORK : tuning().streams);
    if (c->prof || nblocks <= c->chunk_blocks) nw = 1;
    cudaStream_t wk[fdc_chan::NWORK];
    for (int i = 0; i < fdc_chan::NWORK; i++) wk[i] = nw == 1 ? s : c->ws[i];
    const long ring = std::min(nblocks, c->chunk_blocks);
    float2* ring_spec[fdc_chan::NWORK]; float2* ring_mid[fdc_chan::NWORK];
    for (int i = 0; i < nw; i++) {
        const size_t rb = sizeof(float2) * (size_t)ring * c->N;
        if (!d_spectrum && !c->w_spec[i].reserve(rb)) return cuda_fail(cudaGetLastError(), "spectrum ring");
        if (c->big && !c->w_mid[i].reserve(rb)) return cuda_fail(cudaGetLastError(), "four-step intermediate");
        ring_spec[i] = (float2*)c->w_spec[i].p; ring_mid[i] = (float2*)c->w_mid[i].p;
    }
    if (nw > 1) {
        if ((e = cudaEventRecord(c->ev_start, s)) != cudaSuccess) return cuda_fail(e, "event record");
        for (int i = 0; i < nw; i++) if ((e = cudaStreamWaitEvent(wk[i], c->ev_start, 0)) != cudaSuccess) return cuda_fail(e, "stream wait");
    }
    /* Sinks.  FDC_SINK_DMA=0: the extract kernel stores every channel's rows into its owner's buffer itself (peer stores over
     * NVLink for remote owners).  FDC_SINK_DMA=1: rows of REMOTE owners go to a compact local staging slab per owner and the copy
     * engines forward them behind the kernels (one 2-D copy per owner and chunk, a copy stream per worker stream); rows of the
     * local owner are stored directly in both modes. */
    const bool sink_dma = use_sinks && tuning().sink_dma;
    ExtractParams::Sink sk[FDC_MAX_SINKS];
    std::vector<long> stage_off((size_t)(use_sinks ? c->nsinks : 0), 0);
    if (sink_dma) {
        long total = 0;
        for (int k = 0; k < c->nsinks; k++) { stage_off[(size_t)k] = total; if (!c->sink_local[(size_t)k]) total += ring * c->sink_prefix_total[(size_t)k]; }
        for (int i = 0; i < nw; i++) {
            if (!c->w_out[i].reserve(sizeof(float2) * (size_t)std::max(1L, total))) return cuda_fail(cudaGetLastError(), "sink staging slab");
            if (!c->cs[i] && ((e = cudaStreamCreateWithFlags(&c->cs[i], cudaStreamNonBlocking)) != cudaSuccess ||
                              (e = cudaEventCreateWithFlags(&c->ev_x[i], cudaEventDisableTiming)) != cudaSuccess ||
                              (e = cudaEventCreateWithFlags(&c->ev_c[i], cudaEventDisableTiming)) != cudaSuccess)) return cuda_fail(e, "sink copy stream");
        }
    }
    int w = 0;
    for (long b0 = 0; b0 < nblocks; w = (w + 1) % nw) {
        const long nb = std::min(c->chunk_blocks, nblocks - b0);
        float2* spec = d_spectrum ? d_spectrum + b0 * c->N : ring_spec[w];
        float2* stage = sink_dma ? (float2*)c->w_out[w].p : 0;
        bool staged = false;
        for (int k = 0; use_sinks && k < c->nsinks; k++) {
            if (sink_dma && !c->sink_local[(size_t)k] && c->sink_prefix_total[(size_t)k]) {
                sk[k].base = stage + stage_off[(size_t)k]; sk[k].blocks = nb; sk[k].blk0 = 0; staged = true;
            } else {
                sk[k].base = (float2*)c->sink_bases[(size_t)k]; sk[k].blocks = slab_blocks; sk[k].blk0 = slab_first_block + b0;
            }
        }
        /* the staging slab of this worker is free once the copies of the chunk that used it last are done */
        if (staged && c->copies_pending[w] && (e = cudaStreamWaitEvent(wk[w], c->ev_c[w], 0)) != cudaSuccess) return cuda_fail(e, "stream wait");
        if (chan_enqueue_chunk(c, d_in + b0 * c->hop, nb, spec, ring_mid[w], d_out, slab_blocks, slab_first_block + b0, c->blockcount + b0, wk[w],
                               (const float2*)c->d_hist.p, std::max(0L, nh - b0), sk, use_sinks ? c->nsinks : 0, b0 * c->hop, d_spectrum != 0)) return -1;
        if (staged) {
            if ((e = cudaEventRecord(c->ev_x[w], wk[w])) != cudaSuccess || (e = cudaStreamWaitEvent(c->cs[w], c->ev_x[w], 0)) != cudaSuccess)
                return cuda_fail(e, "sink copy ordering");
            for (int k = 0; k < c->nsinks && e == cudaSuccess; k++) {
                if (c->sink_local[(size_t)k] || !c->sink_prefix_total[(size_t)k]) continue;
                const float2* src = stage + stage_off[(size_t)k];
                float2* base = (float2*)c->sink_bases[(size_t)k];
                bool uniform = !c->sink_first.empty();
                const int c0 = uniform ? c->sink_first[(size_t)k] : 0, cn = uniform ? c->sink_count[(size_t)k] : 0;
                for (int i = c0 + 1; i < c0 + cn; i++) if (c->chans[(size_t)i].lout != c->chans[(size_t)c0].lout) uniform = false;
                if (uniform) {
                    /* one 2-D copy: a row per channel, nb * lout items wide, into rows [slab_first_block + b0, +nb) of the owner's slabs */
                    const size_t lo = (size_t)c->chans[(size_t)c0].lout;
                    e = cudaMemcpy2DAsync(base + (size_t)(slab_first_block + b0) * lo, sizeof(float2) * (size_t)slab_blocks * lo, src,
                                          sizeof(float2) * (size_t)nb * lo, sizeof(float2) * (size_t)nb * lo, (size_t)cn, cudaMemcpyDefault, c->cs[w]);
                } else {
                    for (int i = 0; i < c->nchan && e == cudaSuccess; i++) {
                        const ChanDev& ch = c->chans[(size_t)i];
                        if (ch.owner != k) continue;
                        e = cudaMemcpyAsync(base + (size_t)slab_blocks * (size_t)ch.sink_prefix + (size_t)(slab_first_block + b0) * (size_t)ch.lout,
                                            src + (size_t)nb * (size_t)ch.sink_prefix, sizeof(float2) * (size_t)nb * (size_t)ch.lout, cudaMemcpyDefault, c->cs[w]);
                    }
                }
            }
            if (e != cudaSuccess) return cuda_fail(e, "sink copy");
            if ((e = cudaEventRecord(c->ev_c[w], c->cs[w])) != cudaSuccess) return cuda_fail(e, "event record");
            c->copies_pending[w] = true;
        }
        b0 += nb;
    }
    if (sink_dma) {
        for (int i = 0; i < nw; i++)
            if (c->copies_pending[i] && (e = cudaStreamWaitEvent(s, c->ev_c[i], 0)) != cudaSuccess) return cuda_fail(e, "stream wait");
    }
    if (nw > 1) {
        for (int i = 0; i < nw; i++) {
            if ((e = cudaEventRecord(c->ev_done[i], wk[i])) != cudaSuccess) return cuda_fail(e, "event record");
            if ((e = cudaStreamWaitEvent(s, c->ev_done[i], 0)) != cudaSuccess) return cuda_fail(e, "stream wait");
        }
    }
    if (chan_save_history(c, d_in, nblocks, s)) return -1;
    if (c->ovl) {
        if ((e = cudaEventRecord(c->ev_hist, s)) != cudaSuccess) return cuda_fail(e, "event record");
        c->hist_pending = true; c->hist_stream = s;
    }
    c->blockcount += nblocks;
    return 0;
}

/* inpveclen > 1 mode of the hier block (python/FrequencyDomainChannelizer.py:284-290): the input items are already
 * fft-shifted, unnormalised spectra; only normalize_input (x 1/N, :216) and the per-channel chains run. */
int fdc_chan_work_spectrum_device(fdc_chan* c, const void* d_spec_in_v, long nblocks, void* d_out_v, void* d_spectrum_v, void* stream)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    if (nblocks < 0) return fail("nblocks < 0");
    if (nblocks == 0) return 0;
    const float2* d_spec_in = (const float2*)d_spec_in_v; float2* d_out = (float2*)d_out_v; float2* d_spectrum = (float2*)d_spectrum_v;
    cudaStream_t s = stream ? (cudaStream_t)stream : c->stream;
    const long ring = std::min(nblocks, c->chunk_blocks);
    if (!d_spectrum && !c->w_spec[0].reserve(sizeof(float2) * (size_t)ring * c->N)) return cuda_fail(cudaGetLastError(), "spectrum ring");
    for (long b0 = 0; b0 < nblocks; b0 += ring) {
        const long nb = std::min(ring, nblocks - b0);
        float2* spec = d_spectrum ? d_spectrum + b0 * c->N : (float2*)c->w_spec[0].p;
        cudaError_t e = launch_scale(d_spec_in + b0 * c->N, spec, nb * c->N, 1.0f / (float)c->N, s);
        if (e != cudaSuccess) return cuda_fail(e, "normalize_input launch");
        if (!d_out) continue;
        for (size_t g = 0; g < c->groups.size(); g++) {
            ExtractParams q; q.spec = spec; q.spec_stride = c->N; q.tables = (const float2*)c->d_tables.p;
            q.chans = (const ChanDev*)c->d_chans.p + c->groups[g].second.first; q.nsel = c->groups[g].second.second; q.ny = 0; q.out = d_out;
            q.tma_ok = ((uintptr_t)spec % 16 == 0) && (c->N % 2 == 0) && c->group_even_f[g];
            q.l2pf = (tuning().l2pf && q.tma_ok) ? 1 : 0; q.bpt = 1; q.nsinks = 0;
            q.nb = nb; q.call_blocks = nblocks; q.call_blk0 = b0; q.glob_phase0 = (int)((c->blockcount + b0) % c->nphase); q.nphase = c->nphase;
            q.phase_mask = (c->nphase & (c->nphase - 1)) == 0 ? c->nphase - 1 : -1;
            e = launch_extract(q, c->groups[g].first, s);
            if (e != cudaSuccess) return cuda_fail(e, "channel extract launch");
        }
    }
    c->blockcount += nblocks;
    return 0;
}

/* ---- peer memory (one process per GPU): the sink rank exports its output buffer, the other ranks map it and let their
 * extract kernels store straight into it over NVLink (fdc_chan_work_device_slab) -- the gather is fused into the kernel */
int fdc_ipc_export(const void* d_ptr, void* handle64)
{
    if (!d_ptr || !handle64) return fail("fdc_ipc_export: bad arguments");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr));
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
    static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &h, sizeof(h));
    return 0;
}
void* fdc_ipc_open(const void* handle64)
{
    if (!handle64) { fail("fdc_ipc_open: bad arguments"); return 0; }
    cudaIpcMemHandle_t h; memcpy(&h, handle64, sizeof(h));
    void* p = 0;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cuda_fail(e, "cudaIpcOpenMemHandle"); return 0; }
    return p;
}
int fdc_ipc_close(void* d_ptr)
{
    const cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
    return e == cudaSuccess ? 0 : cuda_fail(e, "cudaIpcCloseMemHandle");
}

int fdc_chan_set_profiling(fdc_chan* c, int enable)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    cudaDeviceSynchronize();
    c->prof = enable != 0;
    c->prof_pool.insert(c->prof_pool.end(), c->prof_ev.begin(), c->prof_ev.end());
    c->prof_ev.clear();
    return 0;
}
int fdc_chan_get_profile(fdc_chan* c, double* ms_fwd, double* ms_extract, long* chunks)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return cuda_fail(e, "fdc_chan_get_profile");
    double a = 0, b = 0; long n = 0;
    for (size_t i = 0; i + 2 < c->prof_ev.size(); i += 3) {
        float t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t1, c->prof_ev[i], c->prof_ev[i + 1]);
        cudaEventElapsedTime(&t2, c->prof_ev[i + 1], c->prof_ev[i + 2]);
        a += t1; b += t2; n++;
    }
    c->prof_pool.insert(c->prof_pool.end(), c->prof_ev.begin(), c->prof_ev.end());
    c->prof_ev.clear();
    if (ms_fwd) *ms_fwd = a;
    if (ms_extract) *ms_extract = b;
    if (chunks) *chunks = n;
    return 0;
}
int fdc_chan_chunk_blocks(const fdc_chan* c) { return c ? (int)c->chunk_blocks : -1; }
int fdc_chan_set_chunk_blocks(fdc_chan* c, int blocks)
{
    if (!c || blocks < 1) return fail("fdc_chan_set_chunk_blocks: bad arguments");
    cudaDeviceSynchronize();
    c->chunk_blocks = blocks;
    return 0;
}

int fdc_chan_sync(fdc_chan* c)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    cudaError_t e = cudaStreamSynchronize(c->stream);
    for (int i = 0; i < fdc_chan::NWORK && e == cudaSuccess; i++) if (c->ws[i]) e = cudaStreamSynchronize(c->ws[i]);
    for (int i = 0; i < fdc_chan::NSLOT && e == cudaSuccess; i++) if (c->hs[i]) e = cudaStreamSynchronize(c->hs[i]);
    return e == cudaSuccess ? 0 : cuda_fail(e, "fdc_chan_sync");
}

/* Host buffers in, host buffers out (the GNU Radio work() contract: the scheduler owns the buffers, they are valid only during
 * the call and they are ordinary pageable memory, lib/overlap_save_impl.cc:62-81).  The call is cut into chunks that are
 * pipelined over NSLOT slots, each with its own stream, device buffers and -- for pageable caller memory -- a pinned staging
 * pair owned by the library: worker threads copy the caller's samples into the slot's pinned input while the GPU works on the
 * previous slots, the DMA engines move pinned <-> device, and the channel rows of a finished slot are copied from its pinned
 * output into the caller's buffers (copy pool, fdc_host.cc).  Caller memory that is already pinned (fdc_host_alloc or
 * cudaHostRegister) is used in place.  Every chunk after the first carries its own ovl-sample halo, so chunks are independent;
 * the first one reads the halo from the history buffer on the device (two-segment loaders). */
struct HostSlotJob { long b0, nb; bool staged_out; bool busy; };

static bool is_pinned_host(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

/* a finished slot: wait for its D2H and hand the rows to the caller */
static int host_drain_slot(fdc_chan* c, int slot, HostSlotJob& j, long nblocks, void* const* outs)
{
    if (!j.busy) return 0;
    const cudaError_t e = cudaEventSynchronize(c->h_done[slot]);
    if (e != cudaSuccess) return cuda_fail(e, "host path: waiting for a slot");
    if (j.staged_out) {
        const float2* src = (const float2*)c->p_out[slot].p;
        /* one row per channel: few large rows are cut into pieces, thousands of small ones are grouped into tasks */
        std::vector<void*>& vd = c->drain_dst; std::vector<const void*>& vs = c->drain_src; std::vector<size_t>& vb = c->drain_bytes;
        vd.clear(); vs.clear(); vb.clear();
        for (int i = 0; i < c->nchan; i++) {
            if (!outs[i]) continue;
            const long lo = c->chans[i].lout;
            const size_t bytes = sizeof(float2) * (size_t)(j.nb * lo);
            if (bytes >= (128u << 10)) copy_pool().submit((float2*)outs[i] + j.b0 * lo, src + j.nb * c->chans[i].lout_prefix, bytes, true);
            else { vd.push_back((float2*)outs[i] + j.b0 * lo); vs.push_back(src + j.nb * c->chans[i].lout_prefix); vb.push_back(bytes); }
        }
        if (!vd.empty()) copy_pool().submit_many(vd.data(), vs.data(), vb.data(), vd.size(), true);
        copy_pool().wait();
    }
    j.busy = false;
    (void)nblocks;
    return 0;
}

int fdc_chan_work_host(fdc_chan* c, const void* in_v, long nblocks, void* const* outs, void* spectrum_v)
{
    OnDevice on_dev(c ? c->dev : -1);
    if (!c) return fail("null context");
    if (nblocks < 0) return fail("nblocks < 0");
    if (nblocks == 0) return 0;
    const float2* in = (const float2*)in_v; float2* spectrum = (float2*)spectrum_v;
    /* PCIe is the bound here, not the kernels: small chunks (about tuning().host_chunk_mb MiB of input each) keep the copy
     * engines of both directions busy from the first to the last millisecond of the call */
    long chunk = ((long)tuning().host_chunk_mb << 20) / ((long)c->hop * (long)sizeof(float2));
    if (chunk < 4) chunk = 4;
    chunk = std::min(nblocks, std::min(chunk, c->chunk_blocks));
    const bool want_out = outs && c->nchan;
    const int mode = tuning().host_staging;                 /* 0: never stage, 1: stage pageable memory (default), 2: always stage */
    const bool stage_in = mode == 2 || (mode == 1 && !is_pinned_host(in));
    bool stage_out = false;
    if (want_out) {
        stage_out = mode == 2;
        /* one driver query per call, not per channel: the first and the last connected output decide (the ports of a block
         * are allocated the same way) */
        if (mode == 1) {
            const void* first_out = 0; const void* last_out = 0;
            for (int i = 0; i < c->nchan; i++) if (outs[i]) { if (!first_out) first_out = outs[i]; last_out = outs[i]; }
            stage_out = first_out && (!is_pinned_host(first_out) || (last_out != first_out && !is_pinned_host(last_out)));
        }
    }
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < fdc_chan::NSLOT; i++) {
        if (!c->hs[i] && (e = cudaStreamCreateWithFlags(&c->hs[i], cudaStreamNonBlocking)) != cudaSuccess) return cuda_fail(e, "stream");
        if (!c->h_done[i] && (e = cudaEventCreateWithFlags(&c->h_done[i], cudaEventDisableTiming)) != cudaSuccess) return cuda_fail(e, "event");
        bool ok = c->h_in[i].reserve(sizeof(float2) * (size_t)(c->ovl + chunk * c->hop)) &&
                  c->h_out[i].reserve(sizeof(float2) * (size_t)std::max(1L, chunk * c->lout_total)) &&
                  c->h_spec[i].reserve(sizeof(float2) * (size_t)chunk * c->N) &&
                  (!c->big || c->h_mid[i].reserve(sizeof(float2) * (size_t)chunk * c->N)) &&
                  (!stage_in || c->p_in[i].reserve(sizeof(float2) * (size_t)(c->ovl + chunk * c->hop))) &&
                  (!stage_out || c->p_out[i].reserve(sizeof(float2) * (size_t)std::max(1L, chunk * c->lout_total)));
        if (!ok) return cuda_fail(cudaGetLastError(), "host-path staging buffers");
    }
    /* uniform channel layout in the caller's memory -> one 2-D copy per chunk instead of one per channel */
    bool uniform = c->nchan > 1 && want_out && !stage_out;
    if (uniform) {
        for (int i = 0; i < c->nchan && uniform; i++) {
            if (!outs[i] || c->chans[i].lout != c->chans[0].lout) uniform = false;
            else if (i > 0 && (const char*)outs[i] - (const char*)outs[i - 1] != (ptrdiff_t)(sizeof(float2) * nblocks * c->chans[0].lout)) uniform = false;
        }
    }
    const long nh = c->ovl ? ((long)c->ovl + c->hop - 1) / c->hop : 0;
    HostSlotJob jobs[fdc_chan::NSLOT];
    for (int i = 0; i < fdc_chan::NSLOT; i++) { jobs[i].busy = false; jobs[i].staged_out = false; jobs[i].b0 = jobs[i].nb = 0; }
    int slot = 0, last_slot = 0;
    for (long b0 = 0; b0 < nblocks; b0 += chunk, slot = (slot + 1) % fdc_chan::NSLOT) {
        const long nb = std::min(chunk, nblocks - b0);
        cudaStream_t s = c->hs[slot];
        if (host_drain_slot(c, slot, jobs[slot], nblocks, outs)) return -1;         /* the slot's previous chunk */
        float2* d_in = (float2*)c->h_in[slot].p;
        /* samples this chunk uploads: its own hop samples, preceded by as much of the ovl-sample halo as lies inside the caller's
         * buffer (all of it from the second chunk on when ovl <= chunk * hop); the rest comes from the history buffer */
        const long first = std::max(0L, b0 * c->hop - c->ovl), count = (b0 + nb) * c->hop - first;
        const long lead = b0 * c->hop - first;               /* halo samples in front of the chunk's first new sample */
        float2* d_dst = d_in + (c->ovl - lead);
        const float2* h_src = in + first;
        if (stage_in) {
            copy_pool().submit(c->p_in[slot].p, h_src, sizeof(float2) * (size_t)count);
            copy_pool().wait();
            h_src = (const float2*)c->p_in[slot].p;
        }
        if (b0 < nh && c->hist_pending && (e = cudaStreamWaitEvent(s, c->ev_hist, 0)) != cudaSuccess) return cuda_fail(e, "stream wait");
        e = cudaMemcpyAsync(d_dst, h_src, sizeof(float2) * (size_t)count, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return cuda_fail(e, "H2D");
        float2* d_out = want_out ? (float2*)c->h_out[slot].p : 0;
        /* blocks whose window starts before the uploaded samples read the history buffer: only in chunks that begin inside the
         * first nh blocks of the call; their history is the call's history shifted by the samples already seen */
        const long head = std::max(0L, nh - b0);
        if (head > 0 && b0 > 0) {
            /* overlap above 50 % and a chunk boundary inside the head: the loaders index the history relative to the chunk's
             * first sample, so the history this chunk sees is [old history | samples before b0] -- assemble it in front of the
             * uploaded samples instead (rare: ovl > hop and chunk < nh) */
            const long from_hist = c->ovl - lead;
            e = cudaMemcpyAsync(d_in, (const float2*)c->d_hist.p + (c->ovl - from_hist), sizeof(float2) * (size_t)from_hist, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return cuda_fail(e, "history copy");
        }
        if (chan_enqueue_chunk(c, d_in + c->ovl, nb, (float2*)c->h_spec[slot].p, (float2*)c->h_mid[slot].p, d_out, nb, 0,
                               c->blockcount + b0, s, (b0 == 0) ? (const float2*)c->d_hist.p : 0, (b0 == 0) ? head : 0, 0, 0, 0, spectrum != 0)) return -1;
        if (d_out) {
            if (stage_out) {
                e = cudaMemcpyAsync(c->p_out[slot].p, d_out, sizeof(float2) * (size_t)(nb * c->lout_total), cudaMemcpyDeviceToHost, s);
            } else if (uniform) {
                const size_t lo = (size_t)c->chans[0].lout;
                e = cudaMemcpy2DAsync((float2*)outs[0] + b0 * lo, sizeof(float2) * nblocks * lo, d_out, sizeof(float2) * nb * lo,
                                      sizeof(float2) * nb * lo, (size_t)c->nchan, cudaMemcpyDeviceToHost, s);
            } else {
                for (int i = 0; i < c->nchan && e == cudaSuccess; i++) {
                    if (!outs[i]) continue;
                    const long lo = c->chans[i].lout;
                    e = cudaMemcpyAsync((float2*)outs[i] + b0 * lo, d_out + nb * c->chans[i].lout_prefix, sizeof(float2) * (size_t)(nb * lo),
                                        cudaMemcpyDeviceToHost, s);
                }
            }
            if (e != cudaSuccess) return cuda_fail(e, "D2H");
        }
        if (spectrum) {
            e = cudaMemcpyAsync(spectrum + b0 * c->N, c->h_spec[slot].p, sizeof(float2) * (size_t)(nb * c->N), cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) return cuda_fail(e, "D2H spectrum");
        }
        if (b0 + nb >= nblocks && c->ovl) {
            /* history for the next call = the last ovl samples of [old history | this call's input].  A later chunk holds them in
             * its device input [halo | samples]; a call of one chunk is the device path's case (short calls slide the history).
             * Written into the spare buffer and swapped at once: everything enqueued so far has the old pointer. */
            if (b0 > 0) {
                e = cudaMemcpyAsync(c->d_hist2.p, d_in + nb * c->hop, sizeof(float2) * (size_t)c->ovl, cudaMemcpyDeviceToDevice, s);
                if (e != cudaSuccess) return cuda_fail(e, "history save");
                c->d_hist.swap(c->d_hist2);
            } else if (chan_save_history(c, d_in + c->ovl, nb, s)) return -1;
        }
        if ((e = cudaEventRecord(c->h_done[slot], s)) != cudaSuccess) return cuda_fail(e, "event record");
        jobs[slot].b0 = b0; jobs[slot].nb = nb; jobs[slot].staged_out = stage_out && d_out; jobs[slot].busy = true;
        last_slot = slot;
    }
    /* drain in submission order: oldest first */
    for (int k = 1; k <= fdc_chan::NSLOT; k++) {
        const int sl = (last_slot + k) % fdc_chan::NSLOT;
        if (host_drain_slot(c, sl, jobs[sl], nblocks, outs)) return -1;
    }
    c->hist_pending = false;
    c->blockcount += nblocks;
    return 0;
}

}  // extern "C"

/* ================================================================================================
 * copy / multiply / FFT block replacements (host buffers)
 * ================================================================================================ */
struct fdc_overlap_save : DevCtx { int itemsize, outputlen, overlaplen; DevBuf d_in, d_out, d_hist; cudaStream_t s; };
struct fdc_vector_cut : DevCtx { int itemsize, veclen, offset, blocklen; DevBuf d_in, d_out; cudaStream_t s; };
struct fdc_psw : DevCtx { int blocksize, relinvovl, counter, shift; std::vector<std::complex<float> > tables; DevBuf d_tab, d_in, d_out; cudaStream_t s; };
struct fdc_fft : DevCtx { int n, forward, shift; DevBuf d_in, d_out; cudaStream_t s; };

extern "C" {

fdc_overlap_save* fdc_overlap_save_create(int itemsize, int outputlen, int overlaplen)
{
    if (!require_device()) return 0;
    if (itemsize <= 0 || outputlen <= 0 || overlaplen < 0 || overlaplen >= outputlen) { fail("overlap_save: need itemsize > 0 and 0 <= overlaplen < outputlen"); return 0; }
    fdc_overlap_save* b = new fdc_overlap_save;
    b->itemsize = itemsize; b->outputlen = outputlen; b->overlaplen = overlaplen; b->s = 0;
    const size_t hb = (size_t)itemsize * std::max(overlaplen, 1);
    /* history starts as zeros, lib/overlap_save_impl.cc:52 */
    if (cudaStreamCreateWithFlags(&b->s, cudaStreamNonBlocking) != cudaSuccess || !b->d_hist.reserve(hb) || cudaMemset(b->d_hist.p, 0, hb) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "overlap_save create"); fdc_overlap_save_destroy(b); return 0;
    }
    return b;
}
int fdc_overlap_save_work(fdc_overlap_save* b, int n, const void* in, void* out)
{
    OnDevice on_dev(b ? b->dev : -1);
    if (!b || n < 0) return fail("overlap_save work: bad arguments");
    if (n == 0) return 0;
    const long inplen = b->outputlen - b->overlaplen;
    const size_t ib = (size_t)n * inplen * b->itemsize, ob = (size_t)n * b->outputlen * b->itemsize, hb = (size_t)b->overlaplen * b->itemsize;
    if (!b->d_in.reserve(ib) || !b->d_out.reserve(ob)) return cuda_fail(cudaGetLastError(), "overlap_save buffers");
    cudaError_t e = cudaMemcpyAsync(b->d_in.p, in, ib, cudaMemcpyHostToDevice, b->s);
    if (e == cudaSuccess) e = launch_rowcopy(b->d_in.p, b->d_hist.p, (long)hb, b->d_out.p, n, (long)b->outputlen * b->itemsize, inplen * b->itemsize, -(long)hb, b->s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, b->d_out.p, ob, cudaMemcpyDeviceToHost, b->s);
    /* save the last overlaplen input items (lib/overlap_save_impl.cc:78); defined for overlaplen <= inplen */
    if (e == cudaSuccess && hb) {
        if (ib >= hb) e = cudaMemcpyAsync(b->d_hist.p, (const char*)b->d_in.p + (ib - hb), hb, cudaMemcpyDeviceToDevice, b->s);
        else {
            /* more than 50 % overlap and a short call: slide the history (the reference reads out of bounds here) */
            DevBuf tmp; if (!tmp.reserve(hb)) return cuda_fail(cudaGetLastError(), "overlap_save history");
            e = cudaMemcpyAsync(tmp.p, (const char*)b->d_hist.p + ib, hb - ib, cudaMemcpyDeviceToDevice, b->s);
            if (e == cudaSuccess) e = cudaMemcpyAsync((char*)tmp.p + (hb - ib), b->d_in.p, ib, cudaMemcpyDeviceToDevice, b->s);
            if (e == cudaSuccess) e = cudaMemcpyAsync(b->d_hist.p, tmp.p, hb, cudaMemcpyDeviceToDevice, b->s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(b->s);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->s);
    return e == cudaSuccess ? n : cuda_fail(e, "overlap_save work");
}
void fdc_overlap_save_destroy(fdc_overlap_save* b) { if (!b) return; if (b->s) { cudaStreamSynchronize(b->s); cudaStreamDestroy(b->s); } delete b; }

fdc_vector_cut* fdc_vector_cut_create(int itemsize, int veclen, int offset, int blocklen)
{
    if (!require_device()) return 0;
    /* the reference performs no checks in C++ (only GRC does, grc/FDC_vector_cut_vxx.xml:64-68); reading outside the
     * input vector is undefined there and refused here */
    if (itemsize <= 0 || veclen <= 0 || blocklen <= 0 || offset < 0 || offset + blocklen > veclen) { fail("vector_cut_vxx: need 0 <= offset and offset + blocklen <= veclen"); return 0; }
    fdc_vector_cut* b = new fdc_vector_cut;
    b->itemsize = itemsize; b->veclen = veclen; b->offset = offset; b->blocklen = blocklen; b->s = 0;
    if (cudaStreamCreateWithFlags(&b->s, cudaStreamNonBlocking) != cudaSuccess) { cuda_fail(cudaGetLastError(), "vector_cut create"); delete b; return 0; }
    return b;
}
int fdc_vector_cut_work(fdc_vector_cut* b, int n, const void* in, void* out)
{
    OnDevice on_dev(b ? b->dev : -1);
    if (!b || n < 0) return fail("vector_cut work: bad arguments");
    if (n == 0) return 0;
    const size_t ib = (size_t)n * b->veclen * b->itemsize, ob = (size_t)n * b->blocklen * b->itemsize;
    if (!b->d_in.reserve(ib) || !b->d_out.reserve(ob)) return cuda_fail(cudaGetLastError(), "vector_cut buffers");
    cudaError_t e = cudaMemcpyAsync(b->d_in.p, in, ib, cudaMemcpyHostToDevice, b->s);
    if (e == cudaSuccess) e = launch_rowcopy(b->d_in.p, b->d_in.p, 0, b->d_out.p, n, (long)b->blocklen * b->itemsize, (long)b->veclen * b->itemsize, (long)b->offset * b->itemsize, b->s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, b->d_out.p, ob, cudaMemcpyDeviceToHost, b->s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->s);
    return e == cudaSuccess ? n : cuda_fail(e, "vector_cut work");
}
void fdc_vector_cut_destroy(fdc_vector_cut* b) { if (!b) return; if (b->s) { cudaStreamSynchronize(b->s); cudaStreamDestroy(b->s); } delete b; }

fdc_psw* fdc_psw_create(int blocklen, int numphasestates, int shifts, float passbw, float stopbw, int windowtype)
{
    if (!require_device()) return 0;
    fdc_psw* b = 0;
    try {
        psw_check_args(passbw, stopbw);
        if (blocklen <= 0 || numphasestates <= 0) throw std::invalid_argument("blocklen and numphasestates must be > 0");
        b = new fdc_psw; b->s = 0;
        b->blocksize = blocklen; b->relinvovl = numphasestates; b->counter = 0;
        b->shift = ((shifts % numphasestates) + numphasestates) % numphasestates;       /* prevent negative shift, :58 */
        psw_tables(blocklen, numphasestates, passbw, stopbw, windowtype, b->tables);
        if (cudaStreamCreateWithFlags(&b->s, cudaStreamNonBlocking) != cudaSuccess || !b->d_tab.upload(b->tables.data(), sizeof(float2) * b->tables.size()))
            throw std::runtime_error(std::string("psw create: ") + cudaGetErrorString(cudaGetLastError()));
        return b;
    } catch (const std::exception& e) { fail(e.what()); fdc_psw_destroy(b); return 0; }
}
int fdc_psw_work(fdc_psw* b, int n, const void* in, void* out)
{
    OnDevice on_dev(b ? b->dev : -1);
    if (!b || n < 0) return fail("psw work: bad arguments");
    if (n == 0) return 0;
    const size_t bytes = sizeof(float2) * (size_t)n * b->blocksize;
    if (!b->d_in.reserve(bytes) || !b->d_out.reserve(bytes)) return cuda_fail(cudaGetLastError(), "psw buffers");
    cudaError_t e = cudaMemcpyAsync(b->d_in.p, in, bytes, cudaMemcpyHostToDevice, b->s);
    if (e == cudaSuccess) e = launch_psw((const float2*)b->d_in.p, (float2*)b->d_out.p, (const float2*)b->d_tab.p, n, b->blocksize, b->relinvovl, b->counter, b->shift, b->s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, b->d_out.p, bytes, cudaMemcpyDeviceToHost, b->s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->s);
    if (e != cudaSuccess) return cuda_fail(e, "psw work");
    b->counter = (int)((b->counter + (long)(n % b->relinvovl) * b->shift) % b->relinvovl);
    return n;
}
int fdc_psw_state(const fdc_psw* b, int* blocksize, int* relinvovl, int* counter, int* shift)
{
    if (!b) return fail("null block");
    *blocksize = b->blocksize; *relinvovl = b->relinvovl; *counter = b->counter; *shift = b->shift; return 0;
}
int fdc_psw_tables(const fdc_psw* b, float* out)
{
    if (!b) return fail("null block");
    memcpy(out, b->tables.data(), sizeof(std::complex<float>) * b->tables.size()); return 0;
}
void fdc_psw_destroy(fdc_psw* b) { if (!b) return; if (b->s) { cudaStreamSynchronize(b->s); cudaStreamDestroy(b->s); } delete b; }

fdc_fft* fdc_fft_create(int n, int forward, int shift)
{
    if (!require_device()) return 0;
    if (!tile_len_supported(n)) { fail("fft: size must be a power of two in [2, 16384]"); return 0; }
    fdc_fft* b = new fdc_fft; b->n = n; b->forward = forward != 0; b->shift = shift != 0; b->s = 0;
    if (cudaStreamCreateWithFlags(&b->s, cudaStreamNonBlocking) != cudaSuccess) { cuda_fail(cudaGetLastError(), "fft create"); delete b; return 0; }
    twiddle_table(n);
    return b;
}
int fdc_fft_work(fdc_fft* b, long nvec, const void* in, void* out)
{
    OnDevice on_dev(b ? b->dev : -1);
    if (!b || nvec < 0) return fail("fft work: bad arguments");
    if (nvec == 0) return 0;
    const size_t bytes = sizeof(float2) * (size_t)nvec * b->n;
    if (!b->d_in.reserve(bytes) || !b->d_out.reserve(bytes)) return cuda_fail(cudaGetLastError(), "fft buffers");
    cudaError_t e = cudaMemcpyAsync(b->d_in.p, in, bytes, cudaMemcpyHostToDevice, b->s);
    PlainParams p; p.in = (const float2*)b->d_in.p; p.out = (float2*)b->d_out.p; p.nvec = nvec; p.shift = b->shift;
    if (e == cudaSuccess) e = launch_plain_fft(p, b->n, b->forward, b->s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, b->d_out.p, bytes, cudaMemcpyDeviceToHost, b->s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->s);
    return e == cudaSuccess ? 0 : cuda_fail(e, "fft work");
}
void fdc_fft_destroy(fdc_fft* b) { if (!b) return; if (b->s) { cudaStreamSynchronize(b->s); cudaStreamDestroy(b->s); } delete b; }

/* ---- decimated power rows for a waterfall display ---- */
struct fdc_waterfall : DevCtx { int blocklen, width, logmode; DevBuf d_in, d_rows; PinBuf h_rows; cudaStream_t s; };

fdc_waterfall* fdc_waterfall_create(int blocklen, int width, int logmode)
{
    if (!require_device()) return 0;
    if (blocklen < 1 || (blocklen & (blocklen - 1)) || width < 1 || (width & (width - 1))) { fail("waterfall: blocklen and width must be powers of two"); return 0; }
    fdc_waterfall* b = new fdc_waterfall; b->blocklen = blocklen; b->width = width; b->logmode = logmode != 0; b->s = 0;
    if (cudaStreamCreateWithFlags(&b->s, cudaStreamNonBlocking) != cudaSuccess) { cuda_fail(cudaGetLastError(), "waterfall create"); delete b; return 0; }
    return b;
}
int fdc_waterfall_work_device(fdc_waterfall* b, int nblocks, const void* d_spectrum, float* out_host, void* stream)
{
    OnDevice on_dev(b ? b->dev : -1);
    if (!b || nblocks < 0 || (nblocks > 0 && (!d_spectrum || !out_host))) return fail("waterfall work: bad arguments");
    if (nblocks == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : b->s;
    const size_t bytes = sizeof(float) * (size_t)nblocks * b->width;
    if (!b->d_rows.reserve(bytes) || !b->h_rows.reserve(bytes)) return cuda_fail(cudaGetLastError(), "waterfall buffers");
    cudaError_t e = launch_waterfall_rows((const float2*)d_spectrum, b->blocklen, nblocks, b->blocklen, b->width, b->logmode, (float*)b->d_rows.p, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(b->h_rows.p, b->d_rows.p, bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "waterfall work");
    memcpy(out_host, b->h_rows.p, bytes);
    return nblocks;
}
int fdc_waterfall_work_host(fdc_waterfall* b, int nblocks, const void* spectrum, float* out_host)
{
    OnDevice on_dev(b ? b->dev : -1);
    if (!b || nblocks < 0 || (nblocks > 0 && !spectrum)) return fail("waterfall work: bad arguments");
    if (nblocks == 0) return 0;
    const size_t bytes = sizeof(float2) * (size_t)nblocks * b->blocklen;
    if (!b->d_in.reserve(bytes)) return cuda_fail(cudaGetLastError(), "waterfall input staging");
    const cudaError_t e = cudaMemcpyAsync(b->d_in.p, spectrum, bytes, cudaMemcpyHostToDevice, b->s);
    if (e != cudaSuccess) return cuda_fail(e, "waterfall H2D");
    return fdc_waterfall_work_device(b, nblocks, b->d_in.p, out_host, 0);
}
void fdc_waterfall_destroy(fdc_waterfall* b) { if (!b) return; if (b->s) { cudaStreamSynchronize(b->s); cudaStreamDestroy(b->s); } delete b; }

}  // extern "C"
