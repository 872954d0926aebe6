/* fdc_cabi_act.cu -- C ABI of the activity-gated blocks: PowerActivationChannel, SegmentDetection,
 * activity_detection_channelizer_vcm.  Per work() call:
 *   1. K3 on the whole call: power sums / ratio thresholds / compacted edge lists  (GPU)
 *   2. per block, in order: the reference's bookkeeping on those few numbers -> extraction jobs + ordered ops (host)
 *   3. K2 job kernel: window multiply, half swap, backward FFT, overlap discard for ALL jobs of the call (GPU)
 *   4. replay of the ops on the extracted blocks -> PDUs / files in the reference's order (host) */
#include "fdc_cabi_internal.h"
#include "fdc_act_state.h"
#include "fdc_host.h"
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <unordered_map>

using namespace fdc;

namespace {

/* a published message; the payload lives in the engine's message arena (stable until msg_clear) */
struct OutMsg { MsgMeta meta; const cfloat* ptr; size_t n; long logic_samples; OutMsg() : ptr(0), n(0), logic_samples(-1) {} };

static void log_line(int verbose, const std::string& logfile, std::string s)
{
    /* verbose 1 = console, 2 = file (append, one line per call) -- lib/SegmentDetection_impl.cc:659-672 */
    if (verbose == 1) std::cout << s << std::endl;
    else if (verbose == 2) {
        FILE* f = fopen(logfile.c_str(), "a");
        if (!f) std::cerr << "Outputfile not writable: " << logfile << std::endl;
        else { s += "\n"; fwrite(s.c_str(), 1, s.size(), f); fclose(f); }
    }
}
static void init_logfile(int verbose, const std::string& logfile)
{
    if (verbose != 2) return;
    FILE* f = fopen(logfile.c_str(), "w");
    if (!f) std::cerr << "Logfile not writable: " << logfile << std::endl;
    else { fwrite("\n", 1, 1, f); fclose(f); }
}

/* FDC_ACT_TIMING=1: wall-clock share of the phases of a work() call on stderr (measurement aid) */
static bool act_timing() { static const bool on = getenv("FDC_ACT_TIMING") && atoi(getenv("FDC_ACT_TIMING")) > 0; return on; }
/* FDC_ACT_TIMING=1: device-side split of a call's extraction (kernels / D2H), events destroyed on every exit path */
struct ExtractEvents {
    cudaEvent_t ev[3]; bool on;
    explicit ExtractEvents(bool enable) : on(enable) { for (int i = 0; i < 3; i++) { ev[i] = 0; if (on && cudaEventCreate(&ev[i]) != cudaSuccess) on = false; } }
    ~ExtractEvents() { for (int i = 0; i < 3; i++) if (ev[i]) cudaEventDestroy(ev[i]); }
    void mark(int i, cudaStream_t st) { if (on) cudaEventRecord(ev[i], st); }
    float ms(int a, int b) const { float t = 0; if (on) cudaEventElapsedTime(&t, ev[a], ev[b]); return t; }
};
struct PhaseClock {
    std::chrono::steady_clock::time_point t0; double ms[4]; int k;
    PhaseClock() : t0(std::chrono::steady_clock::now()), k(0) { ms[0] = ms[1] = ms[2] = ms[3] = 0; }
    void lap() { const std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now(); if (k < 4) ms[k++] = std::chrono::duration<double, std::milli>(t - t0).count(); t0 = t; }
};

/* device side shared by the three blocks */
struct ActEngine {
    int N; int dev; int verbose; std::string logfile;
    cudaStream_t s;
    DevBuf d_in, d_hist, d_tab, d_jobs, d_out, d_P, d_cnt, d_rr, d_ri, d_fi, d_pw;
    PinBuf h_res, h_misc;
    PinBuf& h_out_buf() { return h_res; }
    /* buffered output blocks of one active channel (the reference's `data` deque of vectors).  Blocks extracted in the current
     * call are only REFERENCED (they sit in the call's result buffer); what is still buffered when the call ends is copied into
     * `data`.  A burst that is published in the call it was extracted in is therefore copied once, into the message arena. */
    struct Pending {
        std::vector<cfloat> data; size_t head; size_t nowned;            /* blocks kept from earlier calls: data[head ...] */
        std::vector<std::pair<const cfloat*, size_t> > segs; size_t seg_head;   /* blocks of this call, in order */
        size_t nblocks;
        Pending() : head(0), nowned(0), seg_head(0), nblocks(0) {}
        void push(const cfloat* p, size_t n) { if (segs.capacity() == 0) segs.reserve(8); segs.push_back(std::make_pair(p, n)); nblocks++; }
        /* the first nb buffered blocks -> out (null: drop them).  With `later`, copies out of the current call's result buffer are
         * only recorded there (destination, source, bytes) and done in one parallel batch at the end of the replay; blocks kept from
         * earlier calls are copied at once (their buffer may be dropped by a later op) */
        struct Deferred { std::vector<void*> dst; std::vector<const void*> src; std::vector<size_t> bytes; };
        void take(size_t nb, size_t blocksamples, cfloat* out, Deferred* later = 0)
        {
            nblocks -= nb;
            const size_t from_owned = std::min(nb, nowned);
            if (from_owned) {
                const size_t n = from_owned * blocksamples;
                if (out) { memcpy(out, data.data() + head, sizeof(cfloat) * n); out += n; }
                head += n; nowned -= from_owned; nb -= from_owned;
                if (nowned == 0) { data.clear(); head = 0; }
            }
            for (; nb > 0 && seg_head < segs.size(); nb--, seg_head++) {
                if (out) {
                    if (later) { later->dst.push_back(out); later->src.push_back(segs[seg_head].first); later->bytes.push_back(sizeof(cfloat) * segs[seg_head].second); }
                    else copy_and_evict(out, segs[seg_head].first, sizeof(cfloat) * segs[seg_head].second);
                    out += segs[seg_head].second;
                }
            }
            if (seg_head == segs.size()) { segs.clear(); seg_head = 0; }
        }
        /* the first nb buffered blocks as ONE run of the current call's result buffer, if they are one (null otherwise) */
        const cfloat* view(size_t nb) const
        {
            if (nowned || nb == 0 || seg_head + nb > segs.size()) return 0;
            for (size_t k = seg_head; k + 1 < seg_head + nb; k++) if (segs[k].first + segs[k].second != segs[k + 1].first) return 0;
            return segs[seg_head].first;
        }
        /* end of the call: the result buffer is about to be reused */
        void keep()
        {
            if (segs.empty()) return;
            if (head > (1u << 20)) { data.erase(data.begin(), data.begin() + head); head = 0; }
            for (size_t k = seg_head; k < segs.size(); k++) {
                data.insert(data.end(), segs[k].first, segs[k].first + segs[k].second); nowned++;
                evict_lines(segs[k].first, sizeof(cfloat) * segs[k].second);         /* the result buffer is a D2H destination again next call */
            }
            segs.clear(); seg_head = 0;
        }
    };
    /* message payloads: fixed-size chunks that are recycled by msg_clear (no allocation, no page faults in steady state);
     * a chunk never moves, so the pointers handed out by msg_get stay valid until msg_clear */
    struct Arena {
        struct Chunk { std::vector<cfloat> buf; size_t used; Chunk(size_t n) : buf(n), used(0) {} };
        std::vector<std::unique_ptr<Chunk> > chunks; size_t cur;
        Arena() : cur(0) {}
        cfloat* alloc(size_t n)
        {
            for (; cur < chunks.size(); cur++)
                if (chunks[cur]->buf.size() - chunks[cur]->used >= n) { cfloat* p = chunks[cur]->buf.data() + chunks[cur]->used; chunks[cur]->used += n; return p; }
            chunks.push_back(std::unique_ptr<Chunk>(new Chunk(std::max(n, (size_t)1 << 20))));       /* 8 MiB */
            cur = chunks.size() - 1; chunks[cur]->used = n;
            return chunks[cur]->buf.data();
        }
        void reset() { for (size_t i = 0; i < chunks.size(); i++) chunks[i]->used = 0; cur = 0; }
    } arena;
    void msg_clear() { msgs.clear(); arena.reset(); }
    /* PDUs that are views of the pinned result buffer and have not been collected yet (no msg_clear since): copy them into the
     * arena before the buffer is overwritten by the next call.  A consumer that drains the messages after every work() call --
     * the GNU Radio wrapper publishes inside work(), the Python blocks dispatch after it -- never pays for this copy. */
    void rescue_views()
    {
        const cfloat* lo = (const cfloat*)h_res.p; const cfloat* hi = lo ? lo + h_res.cap / sizeof(cfloat) : lo;
        for (size_t i = 0; i < msgs.size(); i++) {
            OutMsg& m = msgs[i];
            if (!m.ptr || !m.n || m.ptr < lo || m.ptr >= hi) continue;
            cfloat* keep = arena.alloc(m.n);
            copy_and_evict(keep, m.ptr, sizeof(cfloat) * m.n);
            m.ptr = keep;
        }
    }
    typedef std::unordered_map<long, Pending> PendingMap;          /* a busy segment replays ~10 ops per block: hashed, not ordered */
    PendingMap pending;
    /* direct-mapped front of the map (uids are consecutive integers, a few hundred are live at a time): the replay looks a uid up
     * once per extracted block; map nodes never move, so the cached pointers stay valid until the uid is erased */
    enum { PCACHE = 4096 };
    struct PendingRef { long uid; Pending* p; PendingRef() : uid(-1), p(0) {} };
    std::vector<PendingRef> pcache;
    Pending& pending_of(long uid)
    {
        if (pcache.empty()) pcache.resize(PCACHE);
        PendingRef& r = pcache[(size_t)uid & (PCACHE - 1)];
        if (r.uid != uid) { r.p = &pending[uid]; r.uid = uid; }
        return *r.p;
    }
    void pending_erase(long uid)
    {
        if (!pcache.empty()) { PendingRef& r = pcache[(size_t)uid & (PCACHE - 1)]; if (r.uid == uid) { r.uid = -1; r.p = 0; } }
        pending.erase(uid);
    }
    std::vector<OutMsg> msgs;
    long uid_counter;
    bool logic_only;               /* host-logic hooks: no device, messages carry metadata and sample counts only */

    ActEngine() : N(0), dev(-1), verbose(0), s(0), uid_counter(0), logic_only(false) {}
    ~ActEngine() { OnDevice on_dev(dev); if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } }

    bool init(int blocklen, const std::vector<cfloat>& tab)
    {
        N = blocklen;
        if (cudaGetDevice(&dev) != cudaSuccess) return false;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (!d_hist.reserve(sizeof(float2) * (size_t)N) || cudaMemset(d_hist.p, 0, sizeof(float2) * (size_t)N) != cudaSuccess) return false;
        return d_tab.upload(tab.data(), sizeof(cfloat) * tab.size());
    }

    /* result offsets of jobs [j0, j1) laid out channel by channel (uid), blocks of a channel in order: a burst that is published in the
     * call it was extracted in is then ONE contiguous run of the result buffer and the PDU can point at it (no copy into the arena).
     * Returns the number of samples, -1 when the uids are too sparse for the counting pass (the caller keeps job order). */
    static long channel_layout(const std::vector<ActJob>& jobs, size_t j0, size_t j1, std::vector<long>& dst)
    {
        if (j1 <= j0) return 0;
        long lo = jobs[j0].uid, hi = jobs[j0].uid;
        for (size_t i = j0; i < j1; i++) { lo = std::min(lo, jobs[i].uid); hi = std::max(hi, jobs[i].uid); }
        if (hi - lo >= (long)(4 * (j1 - j0)) + 1024) return -1;
        std::vector<long> run((size_t)(hi - lo + 2), 0);
        for (size_t i = j0; i < j1; i++) run[(size_t)(jobs[i].uid - lo) + 1] += jobs[i].L - jobs[i].skip;
        for (size_t k = 1; k < run.size(); k++) run[k] += run[k - 1];
        for (size_t i = j0; i < j1; i++) { long& r = run[(size_t)(jobs[i].uid - lo)]; dst[i] = r; r += jobs[i].L - jobs[i].skip; }
        return run[run.size() - 1];
    }
    /* run extraction jobs [j0, j1) of a call from the spectrum rows at d_rows (row `row0` of the call is d_rows[0]; the row before
     * it is `d_prev`, the saved history block when that is null); results land in h_out, job j at dst[j] (job order, or by channel);
     * `layout`: offsets of ALL jobs of the call decided beforehand (time-sharded calls, every rank stores into one common run) */
    int extract(const float2* d_rows, const float2* d_prev, const std::vector<ActJob>& jobs, size_t j0, size_t j1, int row0,
                std::vector<long>& dst, long* total_out, cudaStream_t st, float2* d_dst = 0, bool group_by_channel = false,
                const std::vector<long>* layout = 0)
    {
        long total = 0;
        PhaseClock xc;
        if (layout) {
            dst = *layout;
            for (size_t i = j0; i < j1; i++) total += jobs[i].L - jobs[i].skip;
        } else {
            dst.assign(jobs.size(), 0);
            if (group_by_channel) {
                total = channel_layout(jobs, j0, j1, dst);
                if (total < 0) { group_by_channel = false; total = 0; }
            }
            if (!group_by_channel) for (size_t i = j0; i < j1; i++) { dst[i] = total; total += jobs[i].L - jobs[i].skip; }
        }
        *total_out = total;
        if (j1 <= j0 || logic_only) return 0;
        /* group by IFFT length */
        std::vector<int> order(j1 - j0);
        {   /* stable counting sort by log2(L): one pass to count, one to place (a wideband segment has tens of thousands of jobs per call) */
            size_t first[34]; for (int b = 0; b < 34; b++) first[b] = 0;
            for (size_t i = j0; i < j1; i++) {
                if (!tile_len_supported(jobs[i].L)) return fail("activity channel wider than 16384 bins is not supported by the extract kernel");
                first[__builtin_ctz((unsigned)jobs[i].L) + 1]++;
            }
            for (int b = 1; b < 34; b++) first[b] += first[b - 1];
            for (size_t i = j0; i < j1; i++) order[first[__builtin_ctz((unsigned)jobs[i].L)]++] = (int)i;
        }
        std::vector<ExtractJob> ej(order.size());
        for (size_t k = 0; k < order.size(); k++) {
            const ActJob& j = jobs[order[k]];
            if (!tile_len_supported(j.L)) return fail("activity channel wider than 16384 bins is not supported by the extract kernel");
            ej[k].row = j.row - row0; ej[k].start = j.start; ej[k].tab_off = (int)j.tab_off; ej[k].skip = j.skip; ej[k].dst_off = dst[order[k]];
            if (ej[k].row < -1) return fail("activity extract: job refers to a row this shard does not hold");
        }
        if (!d_dst) rescue_views();
        xc.lap();
        /* d_dst: the caller's device buffer (possibly peer memory of the sink rank), results stay on the device */
        if (!d_jobs.upload(ej.data(), sizeof(ExtractJob) * ej.size()) ||
            (!d_dst && (!d_out.reserve(sizeof(float2) * (size_t)total) || !h_out_buf().reserve(sizeof(float2) * (size_t)total))))
            return cuda_fail(cudaGetLastError(), "activity extract buffers");
        xc.lap();
        ExtractEvents ev(act_timing());
        ev.mark(0, st);
        size_t k = 0;
        while (k < order.size()) {
            size_t e = k; const int L = jobs[order[k]].L;
            while (e < order.size() && jobs[order[e]].L == L) e++;
            JobParams p; p.spec = d_rows; p.spec_stride = N; p.hist = d_prev ? d_prev : (const float2*)d_hist.p; p.tables = (const float2*)d_tab.p;
            p.jobs = (const ExtractJob*)d_jobs.p + k; p.out = d_dst ? d_dst : (float2*)d_out.p; p.njobs = (int)(e - k);
            const cudaError_t ce = launch_jobs(p, L, st);
            if (ce != cudaSuccess) return cuda_fail(ce, "activity extract launch");
            k = e;
        }
        xc.lap();
        ev.mark(1, st);
        cudaError_t ce = d_dst ? cudaSuccess : cudaMemcpyAsync(h_out_buf().p, d_out.p, sizeof(float2) * (size_t)total, cudaMemcpyDeviceToHost, st);
        ev.mark(2, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) return cuda_fail(ce, "activity extract D2H");
        xc.lap();
        if (ev.on)
            fprintf(stderr, "  extract: layout + job list %.3f ms, upload + buffers %.3f ms, launches %.3f ms, wait for kernels + D2H %.3f ms (device: kernels %.3f ms, D2H %.3f ms)\n",
                    xc.ms[0], xc.ms[1], xc.ms[2], xc.ms[3], ev.ms(0, 1), ev.ms(1, 2));
        return 0;
    }
    /* replay the ops of a call on the extracted blocks (res + dst[job]) -> PDUs / files in the reference's order */
    void replay(const cfloat* res, const std::vector<long>& dst, const std::vector<ActJob>& jobs, ActOps& ops)
    {
        if (logic_only) {
            for (size_t i = 0; i < ops.size(); i++) {
                const ActOp& o = ops[i];
                if (o.kind == ActOp::PUSH) pending_of(o.uid).nblocks++;
                else if (o.kind == ActOp::DROP) pending_erase(o.uid);
                else {
                    Pending& q = pending_of(o.uid);
                    const size_t ntake = o.ntake < 0 ? q.nblocks : std::min((size_t)o.ntake, q.nblocks);
                    OutMsg m; m.meta = ops.metas[(size_t)o.meta]; m.logic_samples = (long)(ntake * (size_t)o.blocksamples);
                    q.nblocks -= ntake;
                    if (m.meta.publish) msgs.push_back(std::move(m));
                }
            }
            return;
        }
        Pending::Deferred later;
        for (size_t i = 0; i < ops.size(); i++) {
            const ActOp& o = ops[i];
            if (o.kind == ActOp::PUSH) {
                const ActJob& j = jobs[o.job];
                pending_of(o.uid).push(res + dst[o.job], (size_t)(j.L - j.skip));
            } else if (o.kind == ActOp::DROP) {
                pending_erase(o.uid);
            } else {
                Pending& q = pending_of(o.uid);
                const size_t ntake = o.ntake < 0 ? q.nblocks : std::min((size_t)o.ntake, q.nblocks);
                OutMsg m; m.meta = std::move(ops.metas[(size_t)o.meta]);          /* an op list is replayed once: its strings move into the message */
                m.n = ntake * (size_t)o.blocksamples;
                const bool wanted = m.meta.publish || !m.meta.filename.empty();
                /* only results in the engine's own pinned buffer may be handed out as views (shard_assemble replays the caller's memory) */
                const cfloat* direct = (wanted && m.n && res == (const cfloat*)h_res.p) ? q.view(ntake) : 0;
                if (direct) { m.ptr = direct; q.take(ntake, (size_t)o.blocksamples, 0); }       /* a view of the pinned result buffer */
                else {
                    cfloat* dstp = (wanted && m.n) ? arena.alloc(m.n) : 0;
                    m.ptr = dstp;
                    /* a payload that also goes to a file is needed at once; the others are copied in one parallel batch below */
                    q.take(ntake, (size_t)o.blocksamples, dstp, m.meta.filename.empty() ? &later : 0);
                }
                if (!m.meta.filename.empty()) {
                    FILE* fh = fopen(m.meta.filename.c_str(), "wb");
                    if (!fh) std::cerr << "Cannot write to file " << m.meta.filename << std::endl;
                    else { if (m.n) { fwrite(m.ptr, sizeof(cfloat), m.n, fh); if (direct) evict_lines(m.ptr, sizeof(cfloat) * m.n); } fclose(fh); }
                }
                if (!m.meta.logline.empty()) log_line(verbose, logfile, m.meta.logline);
                if (m.meta.publish) msgs.push_back(std::move(m));
            }
        }
        if (!later.dst.empty()) {
            /* tens of MB per call when the results are not channel-contiguous (time-sharded calls): the copy pool instead of one memcpy after the other */
            copy_pool().submit_many(later.dst.data(), later.src.data(), later.bytes.data(), later.dst.size(), true);
            copy_pool().wait();
        }
        for (PendingMap::iterator it = pending.begin(); it != pending.end(); ++it) it->second.keep();
    }
    /* run all extraction jobs of a call and replay the ops */
    int finish(const float2* d_rows, std::vector<ActJob>& jobs, ActOps& ops, cudaStream_t st)
    {
        std::vector<long> dst; long total = 0;
        if (extract(d_rows, 0, jobs, 0, jobs.size(), 0, dst, &total, st, 0, true)) return -1;
        replay((const cfloat*)h_out_buf().p, dst, jobs, ops);
        return 0;
    }

    /* ---- time-sharded call (one process per GPU): every rank measures its own rows, the compact detection records of
     * all ranks are concatenated and EVERY rank runs the same sequential bookkeeping over them (`decide`), each rank extracts the
     * jobs emitted by its own blocks (a contiguous range of the job list), and the sink rank replays the ops on the concatenated
     * results (`assemble`). */
    struct ShardCall {
        std::vector<char> blob; std::vector<ActJob> jobs; ActOps ops; std::vector<long> job_first; bool decided;
        bool by_channel; std::vector<long> layout;     /* shard_layout(1): offsets of all jobs of the call, channel by channel */
        ShardCall() : decided(false), by_channel(false) {}
    } sh;
    template <class T> void blob_put(const T& v) { const char* c = (const char*)&v; sh.blob.insert(sh.blob.end(), c, c + sizeof(T)); }
    void blob_put_edges(const EdgeBlock& e)
    {
        blob_put((int)e.rise.size()); blob_put((int)e.fall.size());
        for (size_t k = 0; k < e.rise.size(); k++) { blob_put(e.rise[k].first); blob_put(e.rise[k].second); }
        for (size_t k = 0; k < e.fall.size(); k++) blob_put(e.fall[k]);
    }
    static bool blob_get_edges(const char*& cur, const char* end, EdgeBlock& e)
    {
        int n[2];
        if (end - cur < (long)sizeof(n)) return false;
        memcpy(n, cur, sizeof(n)); cur += sizeof(n);
        if (n[0] < 0 || n[1] < 0 || end - cur < (long)n[0] * 8 + (long)n[1] * 4) return false;
        e.rise.resize((size_t)n[0]); e.fall.resize((size_t)n[1]);
        for (int k = 0; k < n[0]; k++) { memcpy(&e.rise[(size_t)k].first, cur, 4); memcpy(&e.rise[(size_t)k].second, cur + 4, 4); cur += 8; }
        if (n[1]) memcpy(e.fall.data(), cur, (size_t)n[1] * 4);
        cur += (size_t)n[1] * 4;
        return true;
    }
    void shard_begin(int nblocks_total)
    {
        sh.jobs.clear(); sh.ops.clear(); sh.job_first.assign(1, 0); sh.decided = false; sh.by_channel = false; sh.layout.clear();
        sh.jobs.reserve((size_t)nblocks_total * 16 + 16); sh.ops.reserve((size_t)nblocks_total * 18 + 16);
    }
    bool shard_rows_ok(int first_row, int nrows) const
    { return sh.decided && first_row >= 0 && nrows >= 0 && (size_t)first_row + (size_t)nrows + 1 <= sh.job_first.size(); }
    long shard_samples(int first_row, int nrows) const
    {
        if (!shard_rows_ok(first_row, nrows)) return fail("shard call: rows out of range or no decided call");
        long t = 0;
        for (long j = sh.job_first[(size_t)first_row]; j < sh.job_first[(size_t)(first_row + nrows)]; j++) t += sh.jobs[(size_t)j].L - sh.jobs[(size_t)j].skip;
        return t;
    }
    /* by_channel: the device form of the call keeps ONE run per block instance for all ranks, laid out channel by channel; d_dst of
     * shard_extract_device and d_results of shard_assemble_device are then the start of that run, and bursts that begin and end
     * inside the call become views of the assembled buffer.  Every rank derives the same offsets from the same job list.
     * Returns the samples of the whole call. */
    long shard_layout(int by_channel)
    {
        if (!sh.decided) return fail("shard call: nothing decided");
        sh.layout.assign(sh.jobs.size(), 0);
        long total = by_channel ? channel_layout(sh.jobs, 0, sh.jobs.size(), sh.layout) : -1;
        sh.by_channel = total >= 0;
        if (!sh.by_channel) { total = 0; for (size_t i = 0; i < sh.jobs.size(); i++) { sh.layout[i] = total; total += sh.jobs[i].L - sh.jobs[i].skip; } }
        return total;
    }
    long shard_extract(int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* out_host)
    {
        if (!shard_rows_ok(first_row, nrows)) return fail("shard call: rows out of range or no decided call");
        OnDevice on_dev(dev);
        cudaStream_t st = stream ? (cudaStream_t)stream : s;
        std::vector<long> dst; long total = 0;
        if (extract((const float2*)d_rows, (const float2*)d_prev, sh.jobs, (size_t)sh.job_first[(size_t)first_row],
                    (size_t)sh.job_first[(size_t)(first_row + nrows)], first_row, dst, &total, st)) return -1;
        if (!logic_only && total > 0 && out_host) { copy_pool().submit(out_host, h_out_buf().p, sizeof(float2) * (size_t)total, true); copy_pool().wait(); }
        return total;
    }
    long shard_extract_device(int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* d_dst)
    {
        if (!shard_rows_ok(first_row, nrows) || logic_only || !d_dst) return fail("shard call: rows out of range, no decided call or no destination");
        OnDevice on_dev(dev);
        std::vector<long> dst; long total = 0;
        if (extract((const float2*)d_rows, (const float2*)d_prev, sh.jobs, (size_t)sh.job_first[(size_t)first_row],
                    (size_t)sh.job_first[(size_t)(first_row + nrows)], first_row, dst, &total, stream ? (cudaStream_t)stream : s, (float2*)d_dst, false,
                    sh.layout.size() == sh.jobs.size() && !sh.jobs.empty() ? &sh.layout : 0)) return -1;
        return total;
    }
    int shard_assemble_device(const void* d_results, long nsamples, void* stream)
    {
        if (!sh.decided || logic_only) return fail("shard call: nothing decided");
        OnDevice on_dev(dev);
        cudaStream_t st = stream ? (cudaStream_t)stream : s;
        PhaseClock pc;
        if (nsamples > 0) {
            rescue_views();
            if (!h_out_buf().reserve(sizeof(float2) * (size_t)nsamples)) return cuda_fail(cudaGetLastError(), "assemble buffer");
            cudaError_t ce = cudaMemcpyAsync(h_out_buf().p, d_results, sizeof(float2) * (size_t)nsamples, cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
            if (ce != cudaSuccess) return cuda_fail(ce, "assemble D2H");
        }
        pc.lap();
        static const float2 none = {0.0f, 0.0f};           /* a null result pointer means "drop the call" to shard_assemble */
        const size_t nops = sh.ops.size();
        const int rc = shard_assemble(h_out_buf().p ? h_out_buf().p : (const void*)&none, nsamples, true);
        pc.lap();
        if (act_timing()) fprintf(stderr, "shard_assemble_device: D2H of %ld samples %.3f ms, replay of %zu ops %.3f ms (%zu msgs)\n", nsamples, pc.ms[0], nops, pc.ms[1], msgs.size());
        return rc;
    }
    int shard_assemble(const void* results, long nsamples, bool laid_out = false)
    {
        if (!sh.decided) return fail("shard call: nothing decided");
        if (results || logic_only) {
            std::vector<long> dst(sh.jobs.size(), 0); long total = 0;
            const bool use_layout = laid_out && sh.layout.size() == sh.jobs.size();
            for (size_t i = 0; i < sh.jobs.size(); i++) { dst[i] = use_layout ? sh.layout[i] : total; total += sh.jobs[i].L - sh.jobs[i].skip; }
            if (!logic_only && total != nsamples) return fail("shard assemble: result size does not match the job list");
            replay((const cfloat*)results, dst, sh.jobs, sh.ops);
        }
        sh.jobs.clear(); sh.ops.clear(); sh.layout.clear(); sh.decided = false; sh.by_channel = false;
        return 0;
    }
    int save_hist(const float2* d_rows, int n, cudaStream_t st)
    {
        cudaError_t e = cudaMemcpyAsync(d_hist.p, d_rows + (size_t)(n - 1) * N, sizeof(float2) * (size_t)N, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        return e == cudaSuccess ? 0 : cuda_fail(e, "history save");
    }
    const float2* stage_host(const void* in, int n)
    {
        if (!d_in.reserve(sizeof(float2) * (size_t)n * N)) { cuda_fail(cudaGetLastError(), "input staging"); return 0; }
        const cudaError_t e = cudaMemcpyAsync(d_in.p, in, sizeof(float2) * (size_t)n * N, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { cuda_fail(e, "H2D"); return 0; }
        return (const float2*)d_in.p;
    }
    int msg_get_all(fdc_msg* out) const
    {
        for (size_t i = 0; i < msgs.size(); i++)
            if (msg_get((int)i, out + i)) return -1;
        return (int)msgs.size();
    }
    long msg_copy_data(float* out) const
    {
        /* tens of MB per call on a busy segment (a few large bursts or thousands of short ones): the copy pool's streaming copies
         * instead of one thread's memcpy; payloads that are views of the pinned result buffer are evicted from the CPU caches as
         * they are read (the buffer is the next call's D2H destination, see fdc_host.cc) */
        const cfloat* lo = (const cfloat*)h_res.p; const cfloat* hi = lo ? lo + h_res.cap / sizeof(cfloat) : lo;
        std::vector<void*> vd[2]; std::vector<const void*> vs[2]; std::vector<size_t> vb[2];
        long total = 0;
        for (size_t i = 0; i < msgs.size(); i++) {
            const OutMsg& m = msgs[i];
            if (m.logic_samples >= 0 || m.n == 0) continue;
            if (out) {
                const int view = (m.ptr >= lo && m.ptr < hi) ? 1 : 0;
                if (m.n >= (32u << 10)) copy_pool().submit(out + 2 * total, m.ptr, sizeof(cfloat) * m.n, view != 0);
                else { vd[view].push_back(out + 2 * total); vs[view].push_back(m.ptr); vb[view].push_back(sizeof(cfloat) * m.n); }
            }
            total += (long)m.n;
        }
        if (out) {
            for (int view = 0; view < 2; view++)
                if (!vd[view].empty()) copy_pool().submit_many(vd[view].data(), vs[view].data(), vb[view].data(), vd[view].size(), view != 0);
            copy_pool().wait();
        }
        return total;
    }
    int msg_get(int i, fdc_msg* out) const
    {
        if (i < 0 || i >= (int)msgs.size() || !out) return fail("message index out of range");
        const OutMsg& m = msgs[i];
        memset(out, 0, sizeof(*out));
        strncpy(out->id, m.meta.id.c_str(), sizeof(out->id) - 1);
        out->finalized = m.meta.finalized ? 1 : 0; out->part = m.meta.part; out->rel_cfreq = m.meta.rel_cfreq; out->rel_bw = m.meta.rel_bw;
        out->blockstart = m.meta.blockstart; out->blockend = m.meta.blockend; out->vectorstart = m.meta.vectorstart; out->vectorend = m.meta.vectorend;
        out->nsamples = m.logic_samples >= 0 ? m.logic_samples : (long)m.n;
        out->data = m.logic_samples >= 0 ? (const float*)0 : (const float*)m.ptr;
        return 0;
    }
};

/* detection front end of one segment for a whole call: K3 kernels + D2H of the compacted edge lists */
struct SegDetect {
    enum { CAP = 1024 };
    std::vector<float> last_power;
    int run(ActEngine& e, const float2* d_rows, int n, const SegGeometry& g, float T, int mean, int guard, std::vector<EdgeBlock>& out, cudaStream_t st)
    {
        const int M = (int)g.M;
        out.assign((size_t)n, EdgeBlock());
        last_power.assign((size_t)std::max(M, 0), 0.0f);
        if (M <= 0) return 0;
        const int cap = std::max(1, std::min(M - 1, (int)CAP));
        if (!e.d_P.reserve(sizeof(float) * (size_t)n * M) || !e.d_cnt.reserve(sizeof(int) * 2 * (size_t)n) ||
            !e.d_rr.reserve(sizeof(float) * (size_t)n * cap) || !e.d_ri.reserve(sizeof(int) * (size_t)n * cap) ||
            !e.d_fi.reserve(sizeof(int) * (size_t)n * cap))
            return cuda_fail(cudaGetLastError(), "detection buffers");
        const size_t bytes = (size_t)n * (2 * sizeof(int) + (size_t)cap * 12) + sizeof(float) * (size_t)M;
        if (!e.h_misc.reserve(bytes)) return cuda_fail(cudaGetLastError(), "detection host buffer");
        cudaError_t ce = launch_group_power(d_rows, e.N, n, (int)g.start, (int)g.D, M, mean, (float*)e.d_P.p, st);
        const float invT = 1.0f / T;
        if (ce == cudaSuccess) ce = launch_edges((const float*)e.d_P.p, n, M, T, invT, guard, cap, (int*)e.d_cnt.p, (float*)e.d_rr.p, (int*)e.d_ri.p, (int*)e.d_fi.p, st);
        char* h = (char*)e.h_misc.p;
        int* h_cnt = (int*)h; float* h_rr = (float*)(h_cnt + 2 * (size_t)n); int* h_ri = (int*)(h_rr + (size_t)n * cap); int* h_fi = h_ri + (size_t)n * cap;
        float* h_last = (float*)(h_fi + (size_t)n * cap);
        /* counts first, then only the used columns of the [block][cap] edge lists (a quiet band has a handful of edges per block:
         * kilobytes instead of 12 * cap bytes per block) */
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_cnt, e.d_cnt.p, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_last, (const float*)e.d_P.p + (size_t)(n - 1) * M, sizeof(float) * (size_t)M, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) return cuda_fail(ce, "detection kernels");
        int used = 0;
        for (int b = 0; b < n; b++) {
            const int nr = h_cnt[2 * b], nf = h_cnt[2 * b + 1];
            if (nr <= cap && nf <= cap) used = std::max(used, std::max(nr, nf));
        }
        if (used > 0) {
            const size_t pitch = sizeof(float) * (size_t)cap, width = sizeof(float) * (size_t)used;
            ce = cudaMemcpy2DAsync(h_rr, pitch, e.d_rr.p, pitch, width, (size_t)n, cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaMemcpy2DAsync(h_ri, pitch, e.d_ri.p, pitch, width, (size_t)n, cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaMemcpy2DAsync(h_fi, pitch, e.d_fi.p, pitch, width, (size_t)n, cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
            if (ce != cudaSuccess) return cuda_fail(ce, "edge list D2H");
        }
        memcpy(last_power.data(), h_last, sizeof(float) * (size_t)M);
        std::vector<float> row;
        for (int b = 0; b < n; b++) {
            const int nr = h_cnt[2 * b], nf = h_cnt[2 * b + 1];
            EdgeBlock& eb = out[(size_t)b];
            if (nr <= cap && nf <= cap) {
                for (int k = 0; k < nr; k++) eb.rise.push_back(std::make_pair(h_rr[(size_t)b * cap + k], h_ri[(size_t)b * cap + k]));
                eb.fall.assign(h_fi + (size_t)b * cap, h_fi + (size_t)b * cap + nf);
            } else {
                /* more edges than the compact lists hold (dense noise-only detections): classify this block from its power row,
                 * same IEEE divisions and comparisons as the kernel */
                row.resize((size_t)M);
                ce = cudaMemcpy(row.data(), (const float*)e.d_P.p + (size_t)b * M, sizeof(float) * (size_t)M, cudaMemcpyDeviceToHost);
                if (ce != cudaSuccess) return cuda_fail(ce, "power row D2H");
                for (int i = 0; i + 1 < M; i++) {
                    float den = row[(size_t)i];
                    if (guard && den == 0.0f) den = FLT_MIN;
                    const float r = row[(size_t)i + 1] / den;
                    if (r > T) eb.rise.push_back(std::make_pair(r, i));
                    else if (r < invT) eb.fall.push_back(i);
                }
            }
        }
        return 0;
    }
};

}  // namespace

/* ================================================================================================ PAC */
struct fdc_pac { ActEngine e; PacState st; std::vector<cfloat> tab; };

extern "C" {

static fdc_pac* pac_build(bool logic_only, int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks,
                          int deactivation_delay, int msg, int fileoutput, const char* path, int verbose, int ID)
{
    fdc_pac* b = new fdc_pac;
    b->e.logic_only = logic_only;
    try {
        b->e.verbose = (verbose == 1 || verbose == 2) ? verbose : 0;
        b->e.logfile = std::string("gr-FDC.PowActChan.") + std::to_string(ID) + std::string(".log");
        init_logfile(b->e.verbose, b->e.logfile);
        b->st.init(blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, ID);
        b->st.msg = msg != 0; b->st.fileoutput = fileoutput != 0; b->st.path = path ? path : ""; b->st.verbose = b->e.verbose;
        build_pac_windows(blocklen, relinvovl, b->st.rampsamps, b->tab);
        if (logic_only) { b->e.N = blocklen; return b; }
        if (b->st.extract_width > 16384) throw std::invalid_argument("PowerActivationChannel: channels wider than 16384 bins are not supported by the extract kernel");
        if (!require_device()) throw std::runtime_error(fdc_last_error());
        if (!b->e.init(blocklen, b->tab)) throw std::runtime_error(std::string("PowerActivationChannel: ") + cudaGetErrorString(cudaGetLastError()));
        twiddle_table(b->st.extract_width);
        return b;
    } catch (const std::exception& ex) { const std::string w = ex.what(); fail(w); delete b; return 0; }
}
fdc_pac* fdc_pac_create(int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks, int deactivation_delay,
                        int msg, int fileoutput, const char* path, int verbose, int ID)
{ return pac_build(false, blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, msg, fileoutput, path, verbose, ID); }
fdc_pac* fdc_pac_create_logic(int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks, int deactivation_delay,
                              int msg, int fileoutput, const char* path, int verbose, int ID)
{ return pac_build(true, blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, msg, fileoutput, path, verbose, ID); }
int fdc_pac_logic_work(fdc_pac* b, int n, const float* pwr)
{
    if (!b || !b->e.logic_only || n < 0) return fail("fdc_pac_logic_work: needs a context from fdc_pac_create_logic");
    std::vector<ActJob> jobs; ActOps ops;
    jobs.reserve((size_t)n * 16 + 16); ops.reserve((size_t)n * 18 + 16);     /* no reallocation (and op copies) while the blocks are walked */
    for (int i = 0; i < n; i++) b->st.block(i, pwr[i], b->e.uid_counter, jobs, ops);
    return b->e.finish(0, jobs, ops, 0) ? -1 : n;
}
int fdc_pac_work_device(fdc_pac* b, int n, const void* d_in, void* stream)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("PowerActivationChannel work: bad arguments");
    if (n == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : b->e.s;
    const float2* rows = (const float2*)d_in;
    if (!b->e.d_pw.reserve(sizeof(float) * (size_t)n) || !b->e.h_misc.reserve(sizeof(float) * (size_t)n)) return cuda_fail(cudaGetLastError(), "power buffers");
    cudaError_t ce = launch_band_power(rows, b->e.N, n, b->st.measure_start, b->st.measure_stop, (float*)b->e.d_pw.p, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(b->e.h_misc.p, b->e.d_pw.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return cuda_fail(ce, "band power");
    const float* pw = (const float*)b->e.h_misc.p;
    std::vector<ActJob> jobs; ActOps ops;
    jobs.reserve((size_t)n * 16 + 16); ops.reserve((size_t)n * 18 + 16);     /* no reallocation (and op copies) while the blocks are walked */
    for (int i = 0; i < n; i++) b->st.block(i, pw[i], b->e.uid_counter, jobs, ops);
    if (b->e.finish(rows, jobs, ops, st)) return -1;
    if (b->e.save_hist(rows, n, st)) return -1;
    return n;
}
int fdc_pac_work_host(fdc_pac* b, int n, const void* in)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("PowerActivationChannel work: bad arguments");
    if (n == 0) return 0;
    const float2* rows = b->e.stage_host(in, n);
    return rows ? fdc_pac_work_device(b, n, rows, 0) : -1;
}
int fdc_pac_state(const fdc_pac* b, int* geo, float* f)
{
    if (!b) return fail("null block");
    const PacState& p = b->st;
    const int g[12] = {p.extract_start, p.extract_stop, p.extract_width, p.measure_start, p.measure_stop, p.deltaphase, p.output_len,
                       p.output_ovl_offset, p.active ? 1 : 0, p.count, p.phase, p.blockcount};
    memcpy(geo, g, sizeof(g)); f[0] = p.thresh; f[1] = p.lastpower;
    return 0;
}
int fdc_pac_tables(const fdc_pac* b, float* out)
{
    if (!b) return fail("null block");
    memcpy(out, b->tab.data(), sizeof(cfloat) * b->tab.size()); return 0;
}
int fdc_pac_msg_count(const fdc_pac* b) { return b ? (int)b->e.msgs.size() : -1; }
int fdc_pac_msg_get(const fdc_pac* b, int i, fdc_msg* out) { return b ? b->e.msg_get(i, out) : fail("null block"); }
void fdc_pac_msg_clear(fdc_pac* b) { if (b) b->e.msg_clear(); }
void fdc_pac_destroy(fdc_pac* b) { delete b; }

}  // extern "C"

/* ================================================================================================ SegmentDetection */
struct fdc_segdet { ActEngine e; SegmentState st; SegDetect det; std::vector<cfloat> tab; std::vector<long> offs; float thresh; long blockcount; };

extern "C" {

static fdc_segdet* segdet_build(bool logic_only, int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh,
                                float minchandist, float window_flank_puffer, int maxblocks_to_emit, int channel_deactivation_delay,
                                int messageoutput, int fileoutput, const char* path, int threads, int verbose)
{
    (void)threads;     /* the reference's std::thread-per-carrier mode changes scheduling only; all carriers of a call share one launch here */
    fdc_segdet* b = new fdc_segdet;
    b->e.logic_only = logic_only;
    try {
        b->e.verbose = verbose == 2 ? 2 : (verbose == 1 ? 1 : 0);
        b->e.logfile = std::string("gr-FDC.ActDetChan.ID_") + std::to_string(ID) + std::string(".log");
        init_logfile(b->e.verbose, b->e.logfile);
        /* validation order and texts of lib/SegmentDetection_impl.cc:66-86 */
        if (blocklen < 1 || (blocklen & (blocklen - 1))) throw std::invalid_argument("Blocklen must be Power of 2. ");
        if (relinvovl < 1 || (relinvovl & (relinvovl - 1))) throw std::invalid_argument("Relinvovl must be Power of 2. ");
        if (thresh < 0.0f) throw std::invalid_argument("Threshold is interpreted as dB and must be greater zero to detect channels accordingly. ");
        b->thresh = db_to_ratio(thresh);
        if (window_flank_puffer < 0.0) throw std::invalid_argument("Window flank puffer must not be smaller 0.0. \n");
        SegmentState& s = b->st;
        s.seg_id = ID; s.blocklen = blocklen; s.relinvovl = relinvovl; s.maxblocks = maxblocks_to_emit; s.delay = channel_deactivation_delay;
        s.flank = (double)window_flank_puffer;
        s.g = segdet_geometry(blocklen, seg_start, seg_stop, minchandist);
        s.emit_inside_loop = false; s.msg_output = messageoutput != 0; s.fileoutput = fileoutput != 0; s.path = path ? path : ""; s.verbose = b->e.verbose != 0;
        build_flank_windows(blocklen, relinvovl, s.flank, b->tab, b->offs);
        s.win_offsets = &b->offs;
        b->blockcount = 0;
        if (logic_only) { b->e.N = blocklen; return b; }
        if (!require_device()) throw std::runtime_error(fdc_last_error());
        if (!b->e.init(blocklen, b->tab)) throw std::runtime_error(std::string("SegmentDetection: ") + cudaGetErrorString(cudaGetLastError()));
        return b;
    } catch (const std::exception& ex) { const std::string w = ex.what(); fail(w); delete b; return 0; }
}
fdc_segdet* fdc_segdet_create(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh, float minchandist,
                              float window_flank_puffer, int maxblocks_to_emit, int channel_deactivation_delay, int messageoutput,
                              int fileoutput, const char* path, int threads, int verbose)
{ return segdet_build(false, ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer, maxblocks_to_emit,
                      channel_deactivation_delay, messageoutput, fileoutput, path, threads, verbose); }
fdc_segdet* fdc_segdet_create_logic(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh, float minchandist,
                                    float window_flank_puffer, int maxblocks_to_emit, int channel_deactivation_delay, int messageoutput,
                                    int fileoutput, const char* path, int threads, int verbose)
{ return segdet_build(true, ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer, maxblocks_to_emit,
                      channel_deactivation_delay, messageoutput, fileoutput, path, threads, verbose); }
/* host classification of decimated power rows: the same IEEE divisions and comparisons as k_edges */
static void classify_rows(const float* P, int n, int M, float T, int guard, std::vector<EdgeBlock>& out)
{
    const float invT = 1.0f / T;
    out.assign((size_t)n, EdgeBlock());
    for (int b = 0; b < n; b++)
        for (int i = 0; i + 1 < M; i++) {
            float den = P[(size_t)b * M + i];
            if (guard && den == 0.0f) den = FLT_MIN;
            const float r = P[(size_t)b * M + i + 1] / den;
            if (r > T) out[(size_t)b].rise.push_back(std::make_pair(r, i));
            else if (r < invT) out[(size_t)b].fall.push_back(i);
        }
}
int fdc_segdet_logic_work(fdc_segdet* b, int n, const float* P)
{
    if (!b || !b->e.logic_only || n < 0) return fail("fdc_segdet_logic_work: needs a context from fdc_segdet_create_logic");
    std::vector<EdgeBlock> edges;
    classify_rows(P, n, (int)b->st.g.M, b->thresh, 0, edges);
    std::vector<ActJob> jobs; ActOps ops;
    jobs.reserve((size_t)n * 16 + 16); ops.reserve((size_t)n * 18 + 16);     /* no reallocation (and op copies) while the blocks are walked */
    for (int i = 0; i < n; i++) { b->st.block(i, edges[(size_t)i], b->blockcount, b->e.uid_counter, jobs, ops); b->blockcount++; }
    if (n > 0) b->det.last_power.assign(P + (size_t)(n - 1) * b->st.g.M, P + (size_t)n * b->st.g.M);
    return b->e.finish(0, jobs, ops, 0) ? -1 : n;
}
int fdc_segdet_work_device(fdc_segdet* b, int n, const void* d_in, void* stream)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("SegmentDetection work: bad arguments");
    if (n == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : b->e.s;
    const float2* rows = (const float2*)d_in;
    std::vector<EdgeBlock> edges;
    PhaseClock pc;
    if (b->det.run(b->e, rows, n, b->st.g, b->thresh, 0, 0, edges, st)) return -1;
    pc.lap();
    std::vector<ActJob> jobs; ActOps ops;
    jobs.reserve((size_t)n * 16 + 16); ops.reserve((size_t)n * 18 + 16);     /* no reallocation (and op copies) while the blocks are walked */
    for (int i = 0; i < n; i++) { b->st.block(i, edges[(size_t)i], b->blockcount, b->e.uid_counter, jobs, ops); b->blockcount++; }
    pc.lap();
    std::vector<long> dst; long total = 0;
    if (b->e.extract(rows, 0, jobs, 0, jobs.size(), 0, dst, &total, st, 0, true)) return -1;
    pc.lap();
    b->e.replay((const cfloat*)b->e.h_out_buf().p, dst, jobs, ops);
    pc.lap();
    if (act_timing())
        fprintf(stderr, "SegmentDetection %d blocks: measure %.3f ms, bookkeeping %.3f ms, extract %.3f ms (%zu jobs, %ld samples), replay %.3f ms (%zu msgs)\n",
                n, pc.ms[0], pc.ms[1], pc.ms[2], jobs.size(), total, pc.ms[3], b->e.msgs.size());
    if (b->e.save_hist(rows, n, st)) return -1;
    return n;
}
int fdc_segdet_work_host(fdc_segdet* b, int n, const void* in)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("SegmentDetection work: bad arguments");
    if (n == 0) return 0;
    const float2* rows = b->e.stage_host(in, n);
    return rows ? fdc_segdet_work_device(b, n, rows, 0) : -1;
}
int fdc_segdet_state(const fdc_segdet* b, long* geo, float* f)
{
    if (!b) return fail("null block");
    const SegmentState& s = b->st;
    geo[0] = s.g.start; geo[1] = s.g.stop; geo[2] = s.g.width; geo[3] = s.g.D; geo[4] = s.g.M; geo[5] = b->blockcount;
    geo[6] = (long)s.active.size(); geo[7] = s.chan_counter; f[0] = b->thresh;
    return 0;
}
int fdc_segdet_window(const fdc_segdet* b, int log2w, int phase, float* out)
{
    if (!b || log2w < 0 || log2w + 1 >= (int)b->offs.size() || phase < 0 || phase >= b->st.relinvovl) return fail("no such window");
    memcpy(out, b->tab.data() + b->offs[(size_t)log2w] + ((long)phase << log2w), sizeof(cfloat) << log2w);
    return 0;
}
int fdc_segdet_power(const fdc_segdet* b, float* out)
{
    if (!b) return fail("null block");
    memcpy(out, b->det.last_power.data(), sizeof(float) * b->det.last_power.size()); return 0;
}
int fdc_segdet_active(const fdc_segdet* b, int i, int* out)
{
    if (!b || i < 0 || i >= (int)b->st.active.size()) return fail("active channel index out of range");
    const ActiveChannel& c = b->st.active[(size_t)i];
    const int v[14] = {c.ID, c.detect_start, c.detect_stop, c.extract_start, c.extract_stop, c.extract_width, c.ovlskip, c.outputsamples, c.count,
                       c.phase, c.phaseincrement, c.inactive, c.part, c.ndata};
    memcpy(out, v, sizeof(v)); return 0;
}
int fdc_segdet_msg_count(const fdc_segdet* b) { return b ? (int)b->e.msgs.size() : -1; }
int fdc_segdet_msg_get(const fdc_segdet* b, int i, fdc_msg* out) { return b ? b->e.msg_get(i, out) : fail("null block"); }
void fdc_segdet_msg_clear(fdc_segdet* b) { if (b) b->e.msg_clear(); }
void fdc_segdet_destroy(fdc_segdet* b) { delete b; }

}  // extern "C"

/* ================================================================================================ activity_detection_channelizer_vcm */
struct fdc_actdet {
    ActEngine e; std::vector<SegmentState> segs; std::vector<SegDetect> det; std::vector<cfloat> tab; std::vector<long> offs;
    float thresh; long blockcount;
};

extern "C" {

static fdc_actdet* actdet_build(bool logic_only, int blocklen, const float* segments, int nsegs, float thresh, int relinvovl, int maxblocks,
                                int message, int fileoutput, const char* path, int threads, float minchandist,
                                int channel_deactivation_delay, double window_flank_puffer, int verbose)
{
    fdc_actdet* b = new fdc_actdet;
    b->e.logic_only = logic_only;
    try {
        b->e.verbose = (verbose == 1 || verbose == 2) ? verbose : 0;
        b->e.logfile = "gr-FDC.ActDetChan.log";
        init_logfile(b->e.verbose, b->e.logfile);
        /* validation order and texts of lib/activity_detection_channelizer_vcm_impl.cc:104-140 */
        if (blocklen < 2 || (blocklen & (blocklen - 1))) throw std::invalid_argument("Blocklen invalid. ");
        const int D = actdet_decimation(blocklen, minchandist);
        if (thresh < 0.0f) throw std::invalid_argument("Threshold is interpreted as dB and must be greater zero. ");
        b->thresh = db_to_ratio(thresh);
        if (relinvovl < 1 || (relinvovl & (relinvovl - 1))) throw std::invalid_argument("Relative inverse overlap is invalid, must be >0 and a power of 2. ");
        if (channel_deactivation_delay < 0) throw std::invalid_argument("Channel deactication delay must not be smaller 0. \n");
        if (window_flank_puffer < 0.0) throw std::invalid_argument("Window flank puffer must not be smaller 0.0. \n");
        if (nsegs < 0 || (nsegs > 0 && !segments)) throw std::invalid_argument("Segment is incorrect. ");
        build_flank_windows(blocklen, relinvovl, window_flank_puffer, b->tab, b->offs);
        b->segs.resize((size_t)nsegs); b->det.resize((size_t)nsegs);
        for (int i = 0; i < nsegs; i++) {
            SegmentState& s = b->segs[(size_t)i];
            s.seg_id = i; s.blocklen = blocklen; s.relinvovl = relinvovl; s.maxblocks = maxblocks; s.delay = channel_deactivation_delay;
            s.flank = window_flank_puffer;
            s.g = actdet_geometry(blocklen, segments[2 * i], segments[2 * i + 1], D);
            s.emit_inside_loop = threads == 0;          /* single-thread order, …vcm_impl.cc:306-337; threaded mode emits parts after all carriers */
            s.msg_output = message != 0; s.fileoutput = fileoutput != 0; s.path = path ? path : ""; s.verbose = b->e.verbose != 0;
            s.win_offsets = &b->offs;
        }
        b->blockcount = 1;
        if (logic_only) { b->e.N = blocklen; return b; }
        if (!require_device()) throw std::runtime_error(fdc_last_error());
        if (!b->e.init(blocklen, b->tab)) throw std::runtime_error(std::string("activity_detection_channelizer_vcm: ") + cudaGetErrorString(cudaGetLastError()));
        return b;
    } catch (const std::exception& ex) { const std::string w = ex.what(); fail(w); delete b; return 0; }
}
fdc_actdet* fdc_actdet_create(int blocklen, const float* segments, int nsegs, float thresh, int relinvovl, int maxblocks, int message,
                              int fileoutput, const char* path, int threads, float minchandist, int channel_deactivation_delay,
                              double window_flank_puffer, int verbose)
{ return actdet_build(false, blocklen, segments, nsegs, thresh, relinvovl, maxblocks, message, fileoutput, path, threads, minchandist,
                      channel_deactivation_delay, window_flank_puffer, verbose); }
fdc_actdet* fdc_actdet_create_logic(int blocklen, const float* segments, int nsegs, float thresh, int relinvovl, int maxblocks, int message,
                                    int fileoutput, const char* path, int threads, float minchandist, int channel_deactivation_delay,
                                    double window_flank_puffer, int verbose)
{ return actdet_build(true, blocklen, segments, nsegs, thresh, relinvovl, maxblocks, message, fileoutput, path, threads, minchandist,
                      channel_deactivation_delay, window_flank_puffer, verbose); }
/* P: for every block the decimated (mean) power rows of all segments, concatenated in segment order */
int fdc_actdet_logic_work(fdc_actdet* b, int n, const float* P)
{
    if (!b || !b->e.logic_only || n < 0) return fail("fdc_actdet_logic_work: needs a context from fdc_actdet_create_logic");
    long rowlen = 0;
    for (size_t s = 0; s < b->segs.size(); s++) rowlen += b->segs[s].g.M;
    std::vector<ActJob> jobs; ActOps ops;
    jobs.reserve((size_t)n * 16 + 16); ops.reserve((size_t)n * 18 + 16);     /* no reallocation (and op copies) while the blocks are walked */
    std::vector<EdgeBlock> eb;
    for (int i = 0; i < n; i++) {
        long off = 0;
        for (size_t s = 0; s < b->segs.size(); s++) {
            const int M = (int)b->segs[s].g.M;
            classify_rows(P + (size_t)i * rowlen + off, 1, M, b->thresh, 1, eb);
            b->segs[s].block(i, eb[0], b->blockcount, b->e.uid_counter, jobs, ops);
            b->det[s].last_power.assign(P + (size_t)i * rowlen + off, P + (size_t)i * rowlen + off + M);
            off += M;
        }
        b->blockcount++;
    }
    return b->e.finish(0, jobs, ops, 0) ? -1 : n;
}
int fdc_actdet_work_device(fdc_actdet* b, int n, const void* d_in, void* stream)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("activity_detection_channelizer_vcm work: bad arguments");
    if (n == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : b->e.s;
    const float2* rows = (const float2*)d_in;
    std::vector<std::vector<EdgeBlock> > edges(b->segs.size());
    for (size_t s = 0; s < b->segs.size(); s++)
        if (b->det[s].run(b->e, rows, n, b->segs[s].g, b->thresh, 1, 1, edges[s], st)) return -1;
    std::vector<ActJob> jobs; ActOps ops;
    jobs.reserve((size_t)n * 16 + 16); ops.reserve((size_t)n * 18 + 16);     /* no reallocation (and op copies) while the blocks are walked */
    for (int i = 0; i < n; i++) {
        /* detection in every segment first, then extraction segment by segment (work(), …vcm_impl.cc:553-566); both orders coincide
         * because a segment's detection only touches its own channel list */
        for (size_t s = 0; s < b->segs.size(); s++) b->segs[s].block(i, edges[s][(size_t)i], b->blockcount, b->e.uid_counter, jobs, ops);
        b->blockcount++;
    }
    if (b->e.finish(rows, jobs, ops, st)) return -1;
    if (b->e.save_hist(rows, n, st)) return -1;
    return n;
}
int fdc_actdet_work_host(fdc_actdet* b, int n, const void* in)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("activity_detection_channelizer_vcm work: bad arguments");
    if (n == 0) return 0;
    const float2* rows = b->e.stage_host(in, n);
    return rows ? fdc_actdet_work_device(b, n, rows, 0) : -1;
}
int fdc_actdet_nsegments(const fdc_actdet* b) { return b ? (int)b->segs.size() : -1; }
int fdc_actdet_segment(const fdc_actdet* b, int i, int* out)
{
    if (!b || i < 0 || i >= (int)b->segs.size()) return fail("segment index out of range");
    const SegmentState& s = b->segs[(size_t)i];
    out[0] = s.seg_id; out[1] = (int)s.g.start; out[2] = (int)s.g.stop; out[3] = (int)s.g.width; out[4] = (int)s.g.D; out[5] = (int)s.g.M;
    out[6] = (int)s.active.size();
    return 0;
}
int fdc_actdet_power(const fdc_actdet* b, int seg, float* out)
{
    if (!b || seg < 0 || seg >= (int)b->segs.size()) return fail("segment index out of range");
    memcpy(out, b->det[(size_t)seg].last_power.data(), sizeof(float) * b->det[(size_t)seg].last_power.size()); return 0;
}
int fdc_actdet_msg_count(const fdc_actdet* b) { return b ? (int)b->e.msgs.size() : -1; }
int fdc_actdet_msg_get(const fdc_actdet* b, int i, fdc_msg* out) { return b ? b->e.msg_get(i, out) : fail("null block"); }
void fdc_actdet_msg_clear(fdc_actdet* b) { if (b) b->e.msg_clear(); }
void fdc_actdet_destroy(fdc_actdet* b) { delete b; }

}  // extern "C"

/* ================================================================================================ time-sharded calls
 * the blocks' state crosses shard boundaries, the measurements do not.  Per global call over `nblocks_total` blocks:
 *   rank r:   *_shard_measure(own rows)            -> compact record (band powers / edge lists), a few bytes per block
 *   all:      all-gather of the records (host side plumbing, FDC/sharded.py), concatenated in block order
 *   all:      *_shard_decide(all records)          -> the same job + op lists on every rank (sequential bookkeeping, replicated)
 *   rank r:   *_shard_extract(own rows)            -> samples of the jobs its blocks emitted, in job order
 *   sink:     *_shard_assemble(concatenated)       -> PDUs; other ranks pass NULL and only drop the call */
extern "C" {

long fdc_pac_shard_measure(fdc_pac* b, int n, const void* d_rows, void* stream)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("PowerActivationChannel shard_measure: bad arguments");
    b->e.sh.blob.clear();
    if (n == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : b->e.s;
    if (!b->e.d_pw.reserve(sizeof(float) * (size_t)n) || !b->e.h_misc.reserve(sizeof(float) * (size_t)n)) return cuda_fail(cudaGetLastError(), "power buffers");
    cudaError_t ce = launch_band_power((const float2*)d_rows, b->e.N, n, b->st.measure_start, b->st.measure_stop, (float*)b->e.d_pw.p, st);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(b->e.h_misc.p, b->e.d_pw.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return cuda_fail(ce, "band power");
    b->e.sh.blob.assign((const char*)b->e.h_misc.p, (const char*)b->e.h_misc.p + sizeof(float) * (size_t)n);
    return (long)b->e.sh.blob.size();
}
long fdc_pac_shard_measure_logic(fdc_pac* b, int n, const float* pwr)
{
    if (!b || n < 0 || !b->e.logic_only) return fail("fdc_pac_shard_measure_logic: needs a context from fdc_pac_create_logic");
    b->e.sh.blob.assign((const char*)pwr, (const char*)pwr + sizeof(float) * (size_t)n);
    return (long)b->e.sh.blob.size();
}
long fdc_pac_shard_decide(fdc_pac* b, int n, const void* blob, long bytes)
{
    if (!b || n < 0 || bytes != (long)sizeof(float) * n) return fail("PowerActivationChannel shard_decide: record size does not match the block count");
    b->e.shard_begin(n);
    for (int i = 0; i < n; i++) {
        float pw; memcpy(&pw, (const char*)blob + sizeof(float) * (size_t)i, sizeof(float));
        b->st.block(i, pw, b->e.uid_counter, b->e.sh.jobs, b->e.sh.ops);
        b->e.sh.job_first.push_back((long)b->e.sh.jobs.size());
    }
    b->e.sh.decided = true;
    return (long)b->e.sh.jobs.size();
}

long fdc_segdet_shard_measure(fdc_segdet* b, int n, const void* d_rows, void* stream)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("SegmentDetection shard_measure: bad arguments");
    b->e.sh.blob.clear();
    if (n == 0) return 0;
    std::vector<EdgeBlock> edges;
    if (b->det.run(b->e, (const float2*)d_rows, n, b->st.g, b->thresh, 0, 0, edges, stream ? (cudaStream_t)stream : b->e.s)) return -1;
    for (int i = 0; i < n; i++) b->e.blob_put_edges(edges[(size_t)i]);
    return (long)b->e.sh.blob.size();
}
long fdc_segdet_shard_measure_logic(fdc_segdet* b, int n, const float* P)
{
    if (!b || n < 0 || !b->e.logic_only) return fail("fdc_segdet_shard_measure_logic: needs a context from fdc_segdet_create_logic");
    std::vector<EdgeBlock> edges;
    classify_rows(P, n, (int)b->st.g.M, b->thresh, 0, edges);
    b->e.sh.blob.clear();
    for (int i = 0; i < n; i++) b->e.blob_put_edges(edges[(size_t)i]);
    return (long)b->e.sh.blob.size();
}
long fdc_segdet_shard_decide(fdc_segdet* b, int n, const void* blob, long bytes)
{
    if (!b || n < 0 || bytes < 0) return fail("SegmentDetection shard_decide: bad arguments");
    const char* cur = (const char*)blob; const char* end = cur + bytes;
    b->e.shard_begin(n);
    EdgeBlock eb;
    for (int i = 0; i < n; i++) {
        if (!ActEngine::blob_get_edges(cur, end, eb)) return fail("SegmentDetection shard_decide: truncated detection record");
        b->st.block(i, eb, b->blockcount, b->e.uid_counter, b->e.sh.jobs, b->e.sh.ops); b->blockcount++;
        b->e.sh.job_first.push_back((long)b->e.sh.jobs.size());
    }
    if (cur != end) return fail("SegmentDetection shard_decide: detection record longer than the block count");
    b->e.sh.decided = true;
    return (long)b->e.sh.jobs.size();
}

long fdc_actdet_shard_measure(fdc_actdet* b, int n, const void* d_rows, void* stream)
{
    OnDevice on_dev(b ? b->e.dev : -1);
    if (!b || n < 0 || b->e.logic_only) return fail("activity_detection_channelizer_vcm shard_measure: bad arguments");
    b->e.sh.blob.clear();
    if (n == 0) return 0;
    std::vector<std::vector<EdgeBlock> > edges(b->segs.size());
    for (size_t s = 0; s < b->segs.size(); s++)
        if (b->det[s].run(b->e, (const float2*)d_rows, n, b->segs[s].g, b->thresh, 1, 1, edges[s], stream ? (cudaStream_t)stream : b->e.s)) return -1;
    for (int i = 0; i < n; i++)
        for (size_t s = 0; s < b->segs.size(); s++) b->e.blob_put_edges(edges[s][(size_t)i]);
    return (long)b->e.sh.blob.size();
}
long fdc_actdet_shard_measure_logic(fdc_actdet* b, int n, const float* P)
{
    if (!b || n < 0 || !b->e.logic_only) return fail("fdc_actdet_shard_measure_logic: needs a context from fdc_actdet_create_logic");
    long rowlen = 0;
    for (size_t s = 0; s < b->segs.size(); s++) rowlen += b->segs[s].g.M;
    std::vector<EdgeBlock> eb;
    b->e.sh.blob.clear();
    for (int i = 0; i < n; i++) {
        long off = 0;
        for (size_t s = 0; s < b->segs.size(); s++) {
            classify_rows(P + (size_t)i * rowlen + off, 1, (int)b->segs[s].g.M, b->thresh, 1, eb);
            b->e.blob_put_edges(eb[0]);
            off += b->segs[s].g.M;
        }
    }
    return (long)b->e.sh.blob.size();
}
long fdc_actdet_shard_decide(fdc_actdet* b, int n, const void* blob, long bytes)
{
    if (!b || n < 0 || bytes < 0) return fail("activity_detection_channelizer_vcm shard_decide: bad arguments");
    const char* cur = (const char*)blob; const char* end = cur + bytes;
    b->e.shard_begin(n);
    EdgeBlock eb;
    for (int i = 0; i < n; i++) {
        for (size_t s = 0; s < b->segs.size(); s++) {
            if (!ActEngine::blob_get_edges(cur, end, eb)) return fail("activity_detection_channelizer_vcm shard_decide: truncated detection record");
            b->segs[s].block(i, eb, b->blockcount, b->e.uid_counter, b->e.sh.jobs, b->e.sh.ops);
        }
        b->blockcount++;
        b->e.sh.job_first.push_back((long)b->e.sh.jobs.size());
    }
    if (cur != end) return fail("activity_detection_channelizer_vcm shard_decide: detection record longer than the block count");
    b->e.sh.decided = true;
    return (long)b->e.sh.jobs.size();
}

#define FDC_SHARD_COMMON(X)                                                                                                        \
    int fdc_##X##_shard_blob(const fdc_##X* b, void* out)                                                                          \
    {                                                                                                                              \
        if (!b || !out) return fail("null argument");                                                                              \
        if (!b->e.sh.blob.empty()) memcpy(out, b->e.sh.blob.data(), b->e.sh.blob.size());                                          \
        return 0;                                                                                                                  \
    }                                                                                                                              \
    long fdc_##X##_shard_samples(const fdc_##X* b, int first_row, int nrows)                                                      \
    { return b ? b->e.shard_samples(first_row, nrows) : fail("null block"); }                                                     \
    long fdc_##X##_shard_layout(fdc_##X* b, int by_channel)                                                                       \
    { return b ? b->e.shard_layout(by_channel) : fail("null block"); }                                                            \
    long fdc_##X##_shard_extract(fdc_##X* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* out_host) \
    { return b ? b->e.shard_extract(first_row, nrows, d_rows, d_prev, stream, out_host) : fail("null block"); }                   \
    int fdc_##X##_shard_assemble(fdc_##X* b, const void* results, long nsamples)                                                  \
    { return b ? b->e.shard_assemble(results, nsamples) : fail("null block"); }                                                   \
    long fdc_##X##_shard_extract_device(fdc_##X* b, int first_row, int nrows, const void* d_rows, const void* d_prev, void* stream, void* d_dst) \
    { return b ? b->e.shard_extract_device(first_row, nrows, d_rows, d_prev, stream, d_dst) : fail("null block"); }               \
    int fdc_##X##_shard_assemble_device(fdc_##X* b, const void* d_results, long nsamples, void* stream)                           \
    { return b ? b->e.shard_assemble_device(d_results, nsamples, stream) : fail("null block"); }                                  \
    int fdc_##X##_msg_get_all(const fdc_##X* b, fdc_msg* out)                                                                      \
    { return (b && out) ? b->e.msg_get_all(out) : fail("null argument"); }                                                        \
    long fdc_##X##_msg_copy_data(const fdc_##X* b, float* out)                                                                     \
    { return b ? b->e.msg_copy_data(out) : fail("null block"); }
FDC_SHARD_COMMON(pac)
FDC_SHARD_COMMON(segdet)
FDC_SHARD_COMMON(actdet)
#undef FDC_SHARD_COMMON

}  // extern "C"
