/* fdc_act_state.cc -- see fdc_act_state.h.  Compile with -ffp-contract=off: the geometry below is compared
 * integer for integer with the reference, and its float/double mix follows the reference expression by expression. */
#include "fdc_act_state.h"
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <ctime>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace fdc {

template <class T> static std::string num2str(T v) { std::ostringstream ss; ss << v; return ss.str(); }

const std::string& time_string(time_t raw)
{
    /* lib/SegmentDetection_impl.cc:680-694; the text only changes once a second, so it is formatted once a second
     * (a wideband segment activates hundreds of carriers per block) */
    static thread_local time_t cached_at = (time_t)-1;
    static thread_local std::string cached;
    if (raw != cached_at) {
        struct tm ti; localtime_r(&raw, &ti);
        char p[80];
        strftime(p, sizeof(p), "%Y-%m-%d-%H-%M-%S", &ti);
        cached = p; cached_at = raw;
    }
    return cached;
}
std::string current_time_string() { time_t raw; time(&raw); return time_string(raw); }

/* fmod(fmod(x, y) + 1, y) -- lib/SegmentDetection_impl.cc:700-703 */
static float mod_f(float x, float y) { return (float)fmod(fmod((double)x, (double)y) + 1.0, (double)y); }
/* lib/SegmentDetection_impl.cc:705-708: 1 << (int)ceil(log2(v)) for the integer v >= 1 the caller passes: the smallest power of two
 * >= v (log2 of a power of two is exact, and one more than a power of two already rounds up), computed on the bits */
static long nextpow2_shift(long v) { return v <= 1 ? 1 : 1l << (64 - __builtin_clzl((unsigned long)(v - 1))); }

/* ---- windows ------------------------------------------------------------------------------------ */
void build_flank_windows(int blocklen, int relinvovl, double flank_puffer, std::vector<cfloat>& tab, std::vector<long>& offsets)
{
    const int nsizes = (int)log2((double)blocklen) + 1;
    offsets.assign((size_t)nsizes + 1, 0);
    for (int s = 0; s < nsizes; s++) offsets[s + 1] = offsets[s] + (long)relinvovl * (1L << s);
    tab.assign((size_t)offsets[nsizes], cfloat(0.f, 0.f));
    for (int s = 0; s < nsizes; s++) {
        const int w = 1 << s;
        const int puffersamples = (int)(flank_puffer * (double)w);
        for (int i = 0; i < relinvovl; i++) {
            cfloat* v = tab.data() + offsets[s] + (long)i * w;
            const cfloat ph(std::polar(1.0, 2.0 * M_PI * (double)i / (double)relinvovl));
            for (int k = 0; k < w; k++) v[k] = ph;
            for (int k = 0; k < puffersamples; k++) {
                const float fl = 0.5f - 0.5f * (float)cos(M_PI * (double)k / (double)puffersamples);
                v[k] *= fl;
                v[w - 1 - k] *= fl;
            }
        }
    }
}

void build_pac_windows(int blocklen, int relinvovl, int rampsamps, std::vector<cfloat>& tab)
{
    tab.assign((size_t)relinvovl * blocklen, cfloat(0.f, 0.f));
    for (int i = 0; i < relinvovl; i++) {
        const cfloat ph = std::polar(1.0f, (float)(2.0f * M_PI * (double)i / (double)relinvovl));
        for (int k = 0; k < blocklen; k++) tab[(size_t)i * blocklen + k] = ph;
    }
    /* rising edge on the first rampsamps entries, mirrored onto the end of the length-blocklen vector */
    for (int i = 0; i < rampsamps; i++)
        for (int r = 0; r < relinvovl; r++) {
            cfloat* v = tab.data() + (size_t)r * blocklen;
            v[i] *= (float)sin(0.5 * M_PI * (double)i / (double)(rampsamps + 1));
            v[blocklen - i - 1] = v[i];
        }
}

/* ---- segment geometry --------------------------------------------------------------------------- */
SegGeometry segdet_geometry(int blocklen_i, float start, float stop, float minchandist)
{
    const size_t blocklen = (size_t)blocklen_i;
    minchandist = mod_f(minchandist, 1.0f);
    start = mod_f(start, 1.0f);
    stop = mod_f(stop, 1.0f);
    if (start == stop) throw std::invalid_argument("Start must not be equal to stop. ");
    if (start > stop) { const float t = start; start = stop; stop = t; }
    const double dec = (double)blocklen * (double)minchandist / 2.0;
    const size_t D = dec < 2.0 ? 1 : (size_t)(int)dec;
    size_t width = (size_t)((double)(stop - start) * (double)blocklen);
    if (width % D) width += D - width % D;
    if (width > blocklen) width = blocklen - (blocklen % D);
    const size_t mid = (size_t)((double)(0.5f * (start + stop)) * (double)blocklen);
    size_t s0 = mid < width / 2 ? 0 : mid - width / 2;
    size_t s1 = s0 + width;
    if (s1 > blocklen) { s1 = blocklen; s0 = s1 - blocklen; }          /* sic: lib/SegmentDetection_impl.cc:630-633 */
    SegGeometry g; g.start = (long)s0; g.stop = (long)s1; g.width = (long)width; g.D = (long)D; g.M = (long)(width / D);
    return g;
}
int actdet_decimation(int blocklen, float minchandist)
{
    if (minchandist <= 0.0f || minchandist >= 1.0)
        throw std::invalid_argument(std::string("Minimum channel distance is invalid. Must be in (0,1), is ") + num2str(minchandist));
    const double dec = (double)blocklen * (double)minchandist / 2.0;
    return dec < 2.0 ? 1 : (int)dec;
}
SegGeometry actdet_geometry(int blocklen, float v0, float v1, int D)
{
    if (v0 >= v1 || v0 < 0.0f || v1 > 1.0f) {
        std::string s = "Segment is incorrect. must be of size 2 with each member in (0,1), with v[0]<v[1]. v is [";
        s += num2str(v0) + std::string(", ") + num2str(v1) + std::string(", ") + std::string("]");
        throw std::invalid_argument(s);
    }
    const int mid = (int)std::abs(round(((double)v1 + (double)v0) * 0.5 * (double)blocklen));
    int width = (int)std::abs(round(((double)v1 - (double)v0) * (double)blocklen));
    width = (width % D == 0) ? width : width + D - width % D;
    if (width >= blocklen) {
        /* the reference loops `while(width>=blocklen) width=blocklen-(blocklen%D)` (…vcm_impl.cc:262-263), which never
         * terminates when D divides blocklen; refuse that case instead of hanging */
        if (blocklen % D == 0) throw std::invalid_argument("Segment spans the whole band (the reference does not terminate for this input). ");
        width = blocklen - (blocklen % D);
    }
    int start = mid - width / 2 <= 0 ? 0 : mid - width / 2;
    int stop = start + width;
    if (stop > blocklen) { stop = blocklen; start = blocklen - width; }
    if (start < 0 || stop > blocklen)
        throw std::invalid_argument(std::string("Cannot evaluate start and stop of segment... start=") + num2str(start) + std::string(", stop=") + num2str(stop));
    SegGeometry g; g.start = start; g.stop = stop; g.width = stop - start; g.D = D;
    if (g.width % D)
        throw std::invalid_argument(std::string("Invalid segment width. Not a multiple of channel detection decimation factor. width=") +
                                    num2str(g.width) + std::string(", chan_det_dec_fact=") + num2str(D));
    g.M = g.width / D;
    return g;
}

/* ---- SegmentState --------------------------------------------------------------------------------- */
/* the reference's comparison (descending ratio, `fipair_sort`) as an inlinable functor: std::sort's sequence of comparisons and
 * moves depends only on the comparison results and the length, so the order among equal ratios stays the reference's whatever
 * the second member is (the reference carries the bin, this code the position in the block's rising-edge list) */
struct ratio_desc_t { bool operator()(const std::pair<float, int>& a, const std::pair<float, int>& b) const { return a.first > b.first; } };

/* rising edges by descending ratio.  Few edges: the reference's std::sort call.  Many (a wideband segment has hundreds per block):
 * an LSD radix sort on the float bits -- the ratios are positive (r > T > 0), so their bit patterns order like the values; when
 * all ratios differ the sorted order is unique and equals std::sort's, and when two are equal (the only case in which the
 * reference's order depends on the algorithm) the list is sorted again from its original order with std::sort itself. */
static void sort_by_ratio(std::vector<std::pair<float, int> >& v)
{
    const size_t n = v.size();
    if (n < 96) { std::sort(v.begin(), v.end(), ratio_desc_t()); return; }
    static thread_local std::vector<std::pair<float, int> > tmp, orig;
    orig = v; tmp.resize(n);
    std::pair<float, int>* a = v.data(); std::pair<float, int>* b = tmp.data();
    bool ok = true;
    for (size_t i = 0; i < n && ok; i++) { unsigned u; memcpy(&u, &a[i].first, 4); ok = (u >> 31) == 0 && a[i].first == a[i].first; }   /* positive, not NaN */
    for (int pass = 0; pass < 4 && ok; pass++) {
        size_t cnt[257]; for (int k = 0; k < 257; k++) cnt[k] = 0;
        const int sh = 8 * pass;
        for (size_t i = 0; i < n; i++) { unsigned u; memcpy(&u, &a[i].first, 4); cnt[255 - ((u >> sh) & 255u) + 1]++; }      /* descending */
        for (int k = 1; k < 257; k++) cnt[k] += cnt[k - 1];
        for (size_t i = 0; i < n; i++) { unsigned u; memcpy(&u, &a[i].first, 4); b[cnt[255 - ((u >> sh) & 255u)]++] = a[i]; }
        std::swap(a, b);
    }
    /* four passes: the result is back in v */
    for (size_t i = 1; i < n && ok; i++) ok = v[i - 1].first > v[i].first;
    if (!ok) { v = orig; std::sort(v.begin(), v.end(), ratio_desc_t()); }
}

/* Occupancy of the detection raster by the accepted candidates of the current block: own[p] = 1 + position in the candidate
 * list of the candidate that owns raster point p (0: free), busy = the same as a bit set.  Only the points written for one
 * block are cleared for the next (a wideband segment has tens of thousands of raster points and a few hundred owned ones). */
struct SegmentState::OwnerMap {
    std::vector<int> own; std::vector<unsigned long long> busy; std::vector<std::pair<int, int> > written;
    void begin(size_t points)
    {
        if (own.size() != points) { own.assign(points, 0); busy.assign((points + 63) / 64, 0ull); written.clear(); return; }
        for (size_t k = 0; k < written.size(); k++)
            for (int q = written[k].first; q < written[k].second; q++) { own[(size_t)q] = 0; busy[(size_t)q >> 6] = 0ull; }
        written.clear();
    }
    bool any(int ps, int pe) const               /* one of the points ps .. pe owned? */
    {
        const size_t w0 = (size_t)ps >> 6, w1 = (size_t)pe >> 6;
        const unsigned long long m0 = ~0ull << (ps & 63), m1 = ~0ull >> (63 - (pe & 63));
        if (w0 == w1) return (busy[w0] & m0 & m1) != 0;
        if ((busy[w0] & m0) != 0 || (busy[w1] & m1) != 0) return true;
        for (size_t w = w0 + 1; w < w1; w++) if (busy[w] != 0) return true;
        return false;
    }
    void take(int ps, int pe, int id)
    {
        for (int q = ps; q < pe; q++) { own[(size_t)q] = id; busy[(size_t)q >> 6] |= 1ull << (q & 63); }
        written.push_back(std::make_pair(ps, pe));
    }
};
SegmentState::OwnerMap& SegmentState::owner_map()
{
    static thread_local OwnerMap m;
    return m;
}

void SegmentState::candidates(const EdgeBlock& e, CandList& poss) const
{
    /* lib/SegmentDetection_impl.cc:195-244.  Same order of operations as the reference: rising edges sorted by ratio (ties
     * resolved as the reference's std::sort call resolves them, sort_by_ratio), then walked from the strongest down.  The
     * scratch vectors live for the thread's lifetime: no allocation per block.
     * Everything works on the detection raster (power-bin index p <-> bin start + D * p): a rising edge at power bin r starts a
     * candidate at raster point r, a falling edge at power bin f ends one at raster point f + 1, and the reference's upper_bound
     * over the falling BINS for the first one above the rising bin is the first f >= r -- found for all rising edges at once by
     * one merge of the two lists (both are in ascending bin order) instead of a binary search each.
     * The reference tests a new candidate against every accepted one (quadratic in the number of carriers; a wideband segment
     * has hundreds per block).  The accepted ones are recorded in an occupancy map over the raster points: candidate [a, b) owns
     * the points a .. b-1.  The reference's test "s < b && e >= a" for some accepted [a, b) is "one of the points s .. e is owned"
     * -- same decisions, asked of a bit set of the owned points.  match() uses the map to find WHICH candidate owns a point. */
    static thread_local std::vector<std::pair<float, int> > rise;         /* (ratio, position in e.rise) */
    static thread_local std::vector<int> end_of;                           /* raster end of the candidate rising edge k would start, -1: none */
    const size_t nr = e.rise.size(), nf = e.fall.size();
    rise.resize(nr); end_of.resize(nr);
    bool ascending = true;
    for (size_t k = 1; k < nr && ascending; k++) ascending = e.rise[k - 1].second <= e.rise[k].second;
    for (size_t k = 0, f = 0; k < nr; k++) {
        rise[k] = std::make_pair(e.rise[k].first, (int)k);
        if (ascending) { while (f < nf && e.fall[f] < e.rise[k].second) f++; end_of[k] = f < nf ? e.fall[f] + 1 : -1; }
        else { std::vector<int>::const_iterator it = std::lower_bound(e.fall.begin(), e.fall.end(), e.rise[k].second); end_of[k] = it == e.fall.end() ? -1 : *it + 1; }
    }
    sort_by_ratio(rise);
    OwnerMap& map = owner_map();
    map.begin((size_t)g.M + 2);
    for (size_t r = 0; r < nr; r++) {
        const int pe = end_of[(size_t)rise[r].second];
        if (pe < 0) continue;
        const int ps = e.rise[(size_t)rise[r].second].second;
        if (map.any(ps, pe)) continue;
        const std::array<long, 2> a = {{(long)ps * g.D + g.start, (long)pe * g.D + g.start}};
        poss.push_back(a);
        map.take(ps, pe, (int)poss.size());       /* 1 + position in poss */
    }
}

bool SegmentState::activate(long detect_start, long detect_end, long& uid_counter)
{
    /* lib/SegmentDetection_impl.cc:290-344 */
    const long detect_width = detect_end - detect_start;
    const long extract_mid = detect_start + detect_width / 2;
    const long extract_width = nextpow2_shift((long)ceil((double)detect_width * (1.0 + 2.0 * flank)));
    if (extract_width > blocklen) return false;         /* the reference logs to cerr and skips the carrier */
    if (extract_width > max_extract_width) {
        if (!warned_wide) {
            fprintf(stderr, "fdc_b200 segment %d: carrier [%ld, %ld) needs a %ld-bin slice, wider than the %ld bins the extract kernel transforms; "
                            "carriers this wide are skipped (the reference would extract them)\n", seg_id, detect_start, detect_end, extract_width, max_extract_width);
            warned_wide = true;
        }
        return false;
    }
    long extract_start = extract_mid - extract_width / 2, extract_end = extract_mid + extract_width / 2;
    if (extract_start < 0) { extract_start = 0; extract_end = extract_width; }
    if (extract_end > blocklen) { extract_end = blocklen; extract_start = blocklen - extract_width; }
    ActiveChannel c;
    c.ID = (int)chan_counter++;
    c.detect_start = (int)detect_start; c.detect_stop = (int)detect_end;
    c.ras_lo = (int)std::max(0l, (detect_start - g.start) / g.D - 1); c.ras_hi = (int)std::min((long)g.M, (detect_end - g.start) / g.D - 1);
    c.extract_start = (int)extract_start; c.extract_stop = (int)extract_end; c.extract_width = (int)extract_width;
    c.extract_window = __builtin_ctzl((unsigned long)extract_width);          /* (int)log2(extract_width), a power of two */
    c.ovlskip = (int)(extract_width / relinvovl);
    c.outputsamples = c.extract_width - c.ovlskip;
    c.count = 0; c.phase = 0; c.phaseincrement = (int)(extract_start % relinvovl); c.inactive = -1; c.part = 0;
    time(&c.act_time);                                   /* the id text is built when a PDU is (channel_id) */
    c.uid = uid_counter++; c.ndata = 0;
    active.push_back(c);
    return true;
}

void SegmentState::match(CandList& poss, long& uid_counter)
{
    /* lib/SegmentDetection_impl.cc:246-288 */
    if (poss.empty()) {
        for (size_t k = 0; k < active.size(); k++) active[k].inactive += 1;
        return;
    }
    /* The reference walks, for every active channel in list order, over all remaining candidates and erases the ones that
     * touch it (pc_start < detect_stop && pc_end >= detect_start); what is left is activated in candidate (strength) order.
     * With the occupancy map of candidates(): candidate [a, b) touches the channel iff it owns one of the raster points
     * detect_start - 1 .. detect_stop - 1 (in raster units); "erased" is a flag.  Same result, no quadratic walk. */
    const size_t n = poss.size();
    const std::vector<int>& own = owner_map().own;
    static thread_local std::vector<char> dead;
    dead.assign(n, 0);
    for (size_t k = 0; k < active.size(); k++) {
        ActiveChannel& c = active[k];
        bool inactive = true;
        for (int q = c.ras_lo; q <= c.ras_hi; q++) {
            const int id = own[(size_t)q];
            if (!id || dead[(size_t)id - 1]) continue;
            dead[(size_t)id - 1] = 1;
            c.inactive = 0; inactive = false;
        }
        if (inactive) c.inactive += 1;
    }
    for (size_t i = 0; i < n; i++)
        if (!dead[i]) activate(poss[i][0], poss[i][1], uid_counter);
}

void SegmentState::job(ActiveChannel& c, int row, std::vector<ActJob>& jobs, ActOps& ops)
{
    /* process_channel, lib/SegmentDetection_impl.cc:399-429 */
    ActJob j; j.L = c.extract_width; j.row = row; j.start = c.extract_start;
    j.tab_off = (*win_offsets)[c.extract_window] + (long)c.phase * c.extract_width;
    j.skip = c.ovlskip; j.uid = c.uid;
    ActOp o; o.kind = ActOp::PUSH; o.uid = c.uid; o.job = (int)jobs.size(); o.ntake = 0; o.blocksamples = c.outputsamples; o.meta = -1;
    jobs.push_back(j); ops.push_back(o);
    c.ndata++; c.count++;
    c.phase = (c.phase + c.phaseincrement) % relinvovl;
}

static void append_int(std::string& s, long v)
{
    char d[24]; int n = 0;
    unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
    do { d[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) s.push_back('-');
    while (n) s.push_back(d[--n]);
}
std::string SegmentState::channel_id(const ActiveChannel& c) const
{
    /* "<time of activation>.DETECTED.<segID>.<chanID>", lib/SegmentDetection_impl.cc:674-678 (hundreds per block on a wideband
     * segment: appended by hand, no stream or printf formatting) */
    std::string id;
    id.reserve(48);
    id = time_string(c.act_time);
    id.append(".DETECTED.", 10);
    append_int(id, seg_id);
    id.push_back('.');
    append_int(id, c.ID);
    return id;
}

void SegmentState::meta(MsgMeta& m, const ActiveChannel& c, long blockcount, bool fin) const
{
    m.id = channel_id(c); m.finalized = fin; m.publish = msg_output;
    m.part = fin ? (c.part > 0 ? c.part : -1) : c.part;
    m.rel_bw = (double)c.extract_width / (double)blocklen;
    m.rel_cfreq = (double)(c.extract_start + c.extract_stop) / 2.0 / (double)blocklen;
    m.blockstart = blockcount - c.count; m.blockend = blockcount;
    m.vectorstart = c.extract_start; m.vectorend = c.extract_stop;
    if (fileoutput) m.filename = path + std::string("/") + m.id + (fin ? std::string(".fin") : std::string(".parted.") + std::to_string(c.part));
    if (verbose) {
        m.logline = m.id + (fin ? std::string(".fin: ") : std::string(".part: ")) + std::string("start=") + num2str(c.extract_start) +
                    std::string(", stop=") + num2str(c.extract_stop) + (fin ? std::string("") : std::string(", part=") + num2str(c.part + 1)) +
                    std::string(", blockstart=") + num2str(blockcount - c.count) + std::string(", blockend=") + num2str(blockcount);
    }
}

void SegmentState::emit_final(ActiveChannel& c, long blockcount, ActOps& ops)
{
    /* emit_channel, lib/SegmentDetection_impl.cc:437-482: everything buffered, even nothing */
    ActOp o; o.kind = ActOp::EMIT; o.uid = c.uid; o.job = -1; o.ntake = -1; o.blocksamples = c.outputsamples;
    meta(ops.new_meta(o), c, blockcount, true);
    ops.push_back(o);
    c.ndata = 0;
}

void SegmentState::emit_partial(ActiveChannel& c, long blockcount, ActOps& ops)
{
    /* emit_unfinished_channel, lib/SegmentDetection_impl.cc:484-539 */
    if (maxblocks < 0 || c.ndata < maxblocks) return;
    const int ntx = maxblocks == 0 ? c.ndata : maxblocks;
    if (ntx <= 0) return;
    ActOp o; o.kind = ActOp::EMIT; o.uid = c.uid; o.job = -1; o.ntake = ntx; o.blocksamples = c.outputsamples;
    meta(ops.new_meta(o), c, blockcount, false);
    ops.push_back(o);
    c.ndata -= ntx;
    c.part++;
}

void SegmentState::block(int row, const EdgeBlock& e, long blockcount, long& uid_counter, std::vector<ActJob>& jobs, ActOps& ops)
{
    static thread_local CandList poss;          /* scratch: no allocation per block */
    poss.clear();
    candidates(e, poss);
    match(poss, uid_counter);
    /* process_active_channels_single_thread, lib/SegmentDetection_impl.cc:346-365 */
    for (size_t k = 0; k < active.size(); k++) {
        ActiveChannel& c = active[k];
        if (c.inactive < 0) { job(c, row - 1, jobs, ops); job(c, row, jobs, ops); c.inactive = 0; }
        else if (c.inactive > delay) emit_final(c, blockcount, ops);
        else job(c, row, jobs, ops);
        if (emit_inside_loop && maxblocks >= 0 && c.ndata >= maxblocks) emit_partial(c, blockcount, ops);
    }
    if (!emit_inside_loop && maxblocks >= 0)
        for (size_t k = 0; k < active.size(); k++)
            if (active[k].ndata >= maxblocks) emit_partial(active[k], blockcount, ops);
    /* clear_inactive_channels, lib/SegmentDetection_impl.cc:541-549 */
    size_t keep = 0;
    for (size_t i = 0; i < active.size(); i++) {          /* same survivors in the same order, one pass instead of an erase each */
        if (active[i].inactive > delay) {
            ActOp o; o.kind = ActOp::DROP; o.uid = active[i].uid; o.job = -1; o.ntake = 0; o.blocksamples = 0; o.meta = -1;
            ops.push_back(o);
        } else {
            if (keep != i) active[keep] = std::move(active[i]);
            keep++;
        }
    }
    active.resize(keep);
}

/* ---- PacState ------------------------------------------------------------------------------------- */
static int pac_nextpow2(int k)
{
    if (k <= 0) throw std::invalid_argument(std::string("Can't eval nextpow2 from ") + std::to_string(k) + std::string("\n"));
    return (int)std::pow(2, ceil(log2((double)k)));
}

void PacState::init(int v_blocklen, float cfreq, float bw, int v_relinvovl, float v_thresh, int v_maxblocks, int v_delay, int v_ID)
{
    /* constructor + set_startstop + set_thresh, lib/PowerActivationChannel_impl.cc:42-135, 314-381 */
    ID = v_ID;
    if (v_blocklen <= 0) throw std::invalid_argument(std::string("Blocklen invalid, must be >0, is ") + std::to_string(v_blocklen));
    blocklen = v_blocklen;
    if (v_relinvovl <= 0 || v_relinvovl != pac_nextpow2(v_relinvovl))
        throw std::invalid_argument(std::string("Relinvovl invalid, must be >0 and power of 2, is ") + std::to_string(v_relinvovl));
    relinvovl = v_relinvovl;

    bw = bw > 0.0f ? bw : -bw;
    if (bw > 1.0 || cfreq - bw / 2.0f < 0.0f || cfreq + bw / 2.0f > 1.0f)
        throw std::invalid_argument(std::string("Desired channel is out of band: cfreq=") + std::to_string(cfreq) + std::string(", bw=") + std::to_string(bw));
    extract_width = pac_nextpow2((int)ceil((double)bw * (double)blocklen));
    if (extract_width > blocklen) extract_width = blocklen;
    const int mid = (int)round((double)cfreq * (double)blocklen);
    extract_start = mid - extract_width / 2;
    if (extract_start < 0) extract_start = 0;
    extract_stop = extract_start + extract_width;
    if (extract_stop > blocklen) { extract_stop = blocklen; extract_start = extract_stop - blocklen; }   /* sic, :333-336 */
    measure_start = (int)round((double)(cfreq - bw / 2.0f) * (double)blocklen);
    measure_stop = (int)round((double)(cfreq + bw / 2.0f) * (double)blocklen);
    if (measure_start < extract_start) measure_start = extract_start;
    if (measure_stop > extract_stop) measure_stop = extract_stop;
    rampsamps = ((extract_stop - extract_start) - (measure_stop - measure_start)) / 3;
    deltaphase = extract_start % relinvovl;
    phase = 0;
    output_ovl_offset = extract_width / relinvovl;
    output_len = extract_width - output_ovl_offset;

    if (v_thresh <= 0.0f) throw std::invalid_argument(std::string("Threshold is interpreted as dB and must be >0.0, is ") + std::to_string(v_thresh));
    thresh = (float)pow(10.0, (double)v_thresh / 10.0);
    maxblocks = v_maxblocks;
    deactivation_delay = v_delay <= 0 ? 0 : v_delay;
    lastpower = std::numeric_limits<float>::max();
    active = false;
    blockcount = 1;
    finished_channels = 0; count = 0; part = 0; uid = -1; ndata = 0;
}

void PacState::job(int row, std::vector<ActJob>& jobs, ActOps& ops)
{
    /* process_channel, lib/PowerActivationChannel_impl.cc:260-284 */
    ActJob j; j.L = extract_width; j.row = row; j.start = extract_start; j.tab_off = (long)phase * blocklen; j.skip = output_ovl_offset; j.uid = uid;
    ActOp o; o.kind = ActOp::PUSH; o.uid = uid; o.job = (int)jobs.size(); o.ntake = 0; o.blocksamples = output_len; o.meta = -1;
    jobs.push_back(j); ops.push_back(o);
    ndata++; count++;
    phase = (phase + deltaphase) % relinvovl;
}

void PacState::emit(bool fin, ActOps& ops)
{
    /* emit_data, lib/PowerActivationChannel_impl.cc:212-258 */
    ActOp o; o.kind = ActOp::EMIT; o.uid = uid; o.job = -1; o.ntake = -1; o.blocksamples = output_len;
    MsgMeta& m = ops.new_meta(o);
    m.id = msgID + (fin ? std::string(".fin") : std::string(".part"));
    m.finalized = fin; m.part = part;
    m.rel_cfreq = (double)(extract_start + extract_stop) / 2.0 / (double)blocklen;
    m.rel_bw = (double)extract_width / (double)blocklen;
    m.blockstart = blockcount - count; m.blockend = blockcount; m.vectorstart = -1; m.vectorend = -1;
    m.publish = msg;                           /* no message port: only the file / log side effects remain */
    if (fileoutput) m.filename = path + std::string("/") + msgID + (fin ? std::string(".fin") : (std::string(".parted.") + std::to_string(part)));
    if (verbose)
        m.logline = msgID + (fin ? std::string(".fin") : (std::string(".parted.") + std::to_string(part))) + std::string(": ") + std::string("start=") +
                    std::to_string(extract_start) + std::string(", stop=") + std::to_string(extract_stop) + std::string(", blockstart=") +
                    std::to_string(blockcount - count) + std::string(", blockend=") + std::to_string(blockcount);
    ops.push_back(o);
    ndata = 0;
    part++;
}

void PacState::block(int row, float pwr, long& uid_counter, std::vector<ActJob>& jobs, ActOps& ops)
{
    /* work + measure_power decision, lib/PowerActivationChannel_impl.cc:137-177, 286-306 */
    if (pwr == 0.0f) pwr = std::numeric_limits<float>::min();
    bool toggle = false;
    if ((!active) && pwr / lastpower >= thresh) toggle = true;
    else if (active && lastpower / pwr >= thresh) toggle = true;
    lastpower = pwr;
    if (toggle) {
        if (!active) {
            /* activate: previous and current block, :198-210 */
            part = 0; count = 0; active = true; phase = 0; ndata = 0;
            if (uid >= 0) { ActOp d; d.kind = ActOp::DROP; d.uid = uid; d.job = -1; d.ntake = 0; d.blocksamples = 0; d.meta = -1; ops.push_back(d); }
            uid = uid_counter++;
            msgID = current_time_string() + std::string(".PowActChan.") + std::to_string(ID) + std::string(".") + std::to_string(finished_channels);
            job(row - 1, jobs, ops);
            job(row, jobs, ops);
        } else {
            job(row, jobs, ops);
            active = false;
            emit(true, ops);
            finished_channels++;
        }
    } else if (active) {
        job(row, jobs, ops);
        if (maxblocks == 0 || (maxblocks > 0 && count % maxblocks == 0)) emit(false, ops);
    }
    blockcount++;
}

}  // namespace fdc
