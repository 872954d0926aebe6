/* fdc_act_state.h -- host-side state of the activity-gated blocks (no CUDA here).
 *
 * The GPU does the arithmetic (power sums, ratio thresholds, window multiply + IFFT); what stays on the host is the
 * small, strictly sequential bookkeeping the reference does per block: pairing rising/falling edges, matching
 * candidates against the active-channel list, activation geometry, deactivation counters, PDU metadata.  All of it is
 * integer work on a handful of channels per block and has to be reproduced exactly.
 *
 * Because extraction is batched (one kernel launch per work() call, not one FFT per block), the per-block pass does not
 * touch sample data: it emits extraction JOBS and an ordered list of OPS (push the result of job j onto channel u's
 * buffer / publish a message from the first n buffered blocks of channel u).  After the kernel has run the ops are
 * replayed on the real data, which reproduces the reference's PDU sequence. */
#ifndef FDC_ACT_STATE_H
#define FDC_ACT_STATE_H
#include <array>
#include <complex>
#include <ctime>
#include <deque>
#include <memory>
#include <string>
#include <vector>

namespace fdc {

typedef std::complex<float> cfloat;

struct ActJob {
    int L;            /* extract width (IFFT length) */
    int row;          /* spectrum row of this call, -1 = history block */
    int start;        /* extract_start */
    long tab_off;     /* offset (in complex items) of the phase-selected window in the block's table buffer */
    int skip;         /* leading IFFT outputs dropped */
    long uid;         /* channel instance the result belongs to */
};
struct MsgMeta {
    std::string id;
    bool finalized;
    long part;                    /* -1: key absent */
    double rel_cfreq, rel_bw;
    long blockstart, blockend, vectorstart, vectorend;   /* vector*: -1 = key absent */
    bool publish;                 /* message port connected (msg / messageoutput constructor flag) */
    std::string filename;         /* non-empty: also write the payload there (raw cfp32) */
    std::string logline;
};
struct ActOp {
    enum Kind { PUSH, EMIT, DROP } kind;
    long uid;
    int job;                      /* PUSH: index into the call's job list */
    int ntake;                    /* EMIT: number of buffered blocks to publish, -1 = all */
    int blocksamples;             /* EMIT: samples per buffered block */
    int meta;                     /* EMIT: index into the op list's `metas`, -1 otherwise (ops are one per extracted block: small, trivially copyable) */
};
/* the ordered ops of a call and the metadata of the PDUs they publish */
struct ActOps {
    std::vector<ActOp> v; std::vector<MsgMeta> metas;
    void reserve(size_t n) { v.reserve(n); }
    void clear() { v.clear(); metas.clear(); }
    size_t size() const { return v.size(); }
    const ActOp& operator[](size_t i) const { return v[i]; }
    void push_back(const ActOp& o) { v.push_back(o); }
    MsgMeta& new_meta(ActOp& o) { o.meta = (int)metas.size(); metas.emplace_back(); return metas.back(); }
};

std::string current_time_string();            /* "%Y-%m-%d-%H-%M-%S" */
const std::string& time_string(time_t t);      /* the same for a given time (cached per thread: the text changes once a second) */

/* ---- windows ---------------------------------------------------------------------------------- */
/* lib/SegmentDetection_impl.cc:551-583 == lib/activity_detection_channelizer_vcm_impl.cc:199-228:
 * for every width 2^s <= blocklen, relinvovl phase-rotated rectangles with raised-cosine flanks.
 * Flat layout: offsets[s] + phase * 2^s. */
void build_flank_windows(int blocklen, int relinvovl, double flank_puffer, std::vector<cfloat>& tab, std::vector<long>& offsets);
/* lib/PowerActivationChannel_impl.cc:357-375: relinvovl vectors of length blocklen */
void build_pac_windows(int blocklen, int relinvovl, int rampsamps, std::vector<cfloat>& tab);

/* ---- one detection segment (SegmentDetection, or one `segment` of activity_detection_channelizer_vcm) ---- */
struct ActiveChannel {
    int ID, detect_start, detect_stop, extract_start, extract_stop, extract_width, extract_window, ovlskip, outputsamples;
    int count, phase, phaseincrement, inactive, part;
    time_t act_time;              /* wall clock of the activation: the "<time>" of the channel's message id, formatted when a PDU is built
                                   * (no string in here: the list of a wideband segment is compacted every block) */
    long uid;
    int ndata;                    /* number of buffered blocks (the reference's data.size()) */
    int ras_lo, ras_hi;           /* raster points a candidate must own one of to count as this channel (match()) */
};
struct EdgeBlock {                /* detection result of one block, ascending bin order */
    std::vector<std::pair<float, int> > rise;    /* (ratio, power-bin index i) */
    std::vector<int> fall;                       /* power-bin index i */
};

struct SegGeometry { long start, stop, width, D, M; };
/* lib/SegmentDetection_impl.cc:592-637 (throws std::invalid_argument with the reference's text) */
SegGeometry segdet_geometry(int blocklen, float start, float stop, float minchandist);
/* lib/activity_detection_channelizer_vcm_impl.cc:230-279 */
int actdet_decimation(int blocklen, float minchandist);
SegGeometry actdet_geometry(int blocklen, float v0, float v1, int D);

class SegmentState {
public:
    int seg_id;                   /* ID used in message ids */
    int blocklen, relinvovl, maxblocks, delay;
    double flank;
    SegGeometry g;
    bool emit_inside_loop;        /* activity_detection single-thread order (…vcm_impl.cc:306-337) vs SegmentDetection order */
    std::vector<ActiveChannel> active;        /* in the reference's list order; plain data, compacted in place every block */
    long chan_counter;
    const std::vector<long>* win_offsets;
    std::string path; bool fileoutput; bool verbose; bool msg_output;
    /* widest slice the extract kernel transforms (one CTA: 16384 points).  The reference extracts any width up to blocklen
     * (lib/SegmentDetection_impl.cc:290-309); a carrier that would need more is refused HERE, before any state is touched, and
     * reported once -- it is never half activated (a work() call that fails in the extract after the bookkeeping has
     * advanced would leave the channel active and fail every later call) */
    long max_extract_width; bool warned_wide;

    SegmentState() : seg_id(0), blocklen(0), relinvovl(1), maxblocks(-1), delay(0), flank(0.0), emit_inside_loop(false),
                     chan_counter(0), win_offsets(0), fileoutput(false), verbose(false), msg_output(true),
                     max_extract_width(16384), warned_wide(false) {}
    /* detection + bookkeeping of one block; `blockcount` is the reference's counter value during this block */
    void block(int row, const EdgeBlock& e, long blockcount, long& uid_counter, std::vector<ActJob>& jobs, ActOps& ops);
private:
    typedef std::vector<std::array<long, 2> > CandList;          /* accepted [start, end) candidates of a block, strongest first */
    void candidates(const EdgeBlock& e, CandList& poss) const;
    struct OwnerMap;                                            /* per-thread scratch shared by candidates() and match() */
    static OwnerMap& owner_map();
    std::string channel_id(const ActiveChannel& c) const;
    void match(CandList& poss, long& uid_counter);
    bool activate(long detect_start, long detect_end, long& uid_counter);
    void job(ActiveChannel& c, int row, std::vector<ActJob>& jobs, ActOps& ops);
    void emit_final(ActiveChannel& c, long blockcount, ActOps& ops);
    void emit_partial(ActiveChannel& c, long blockcount, ActOps& ops);
    void meta(MsgMeta& m, const ActiveChannel& c, long blockcount, bool fin) const;
};

/* ---- PowerActivationChannel --------------------------------------------------------------------- */
class PacState {
public:
    int blocklen, relinvovl, extract_start, extract_stop, extract_width, output_len, output_ovl_offset, measure_start,
        measure_stop, maxblocks, deactivation_delay, rampsamps;
    float thresh, lastpower;
    bool active;
    int count, phase, deltaphase, ID, part, finished_channels, blockcount;
    std::string msgID, path; bool msg, fileoutput; int verbose;
    long uid; int ndata;

    /* constructor arguments of lib/PowerActivationChannel_impl.cc:42; throws std::invalid_argument like the reference */
    void init(int v_blocklen, float cfreq, float bw, int v_relinvovl, float v_thresh, int v_maxblocks, int v_delay, int v_ID);
    void block(int row, float pwr, long& uid_counter, std::vector<ActJob>& jobs, ActOps& ops);
private:
    void job(int row, std::vector<ActJob>& jobs, ActOps& ops);
    void emit(bool fin, ActOps& ops);
};

}  // namespace fdc
#endif
