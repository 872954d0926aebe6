/* fdc_gr_blocks.cc -- the six gr::FDC blocks as thin GNU Radio wrappers over libfdc_b200.so (include/fdc_cabi.h).
 *
 * Replaces the bodies of the reference's lib/<block>_impl.cc: the scheduler-facing side (io_signature, names,
 * message port "msgout", sync 1:1 work contract, constructor exceptions) is kept, the arithmetic is the CUDA
 * library's.  GNU Radio hands work() host buffers it owns, so the *_host entry points are used; a flowgraph that
 * wants to stay on the device uses the hier block's fused context (fdc_chan_*) instead.
 * Builds against GNU Radio 3.7/3.8 headers; the CPU test-suite builds it against the oracle's header shim
 * (tests/test_gr_wrappers.py) to run the wrappers and the reference blocks through the same driver. */
#include <FDC/fdc_blocks.h>
#include <gnuradio/io_signature.h>
#include <pmt/pmt.h>
#include <complex>
#include <stdexcept>
#include "fdc_cabi.h"

namespace gr {
namespace FDC {
namespace {

[[noreturn]] void rethrow() { throw std::invalid_argument(fdc_last_error()); }

/* the pmt pair of lib/SegmentDetection_impl.cc:446-460 / lib/PowerActivationChannel_impl.cc:222-232, keys in the
 * reference's insertion order: PowerActivationChannel adds rel_cfreq before rel_bw (:226-227), SegmentDetection and
 * activity_detection_channelizer_vcm rel_bw before rel_cfreq (:454-455).  A GNU Radio 3.7 pmt dict is an association list,
 * so the order shows in printed and serialised PDUs.  PowerActivationChannel messages are the ones without
 * vectorstart / vectorend. */
pmt::pmt_t to_pdu(const fdc_msg& m)
{
    const bool pac = m.vectorstart < 0;
    pmt::pmt_t d = pmt::make_dict();
    d = pmt::dict_add(d, pmt::intern("ID"), pmt::intern(m.id));
    d = pmt::dict_add(d, pmt::intern("finalized"), pmt::from_bool(m.finalized != 0));
    if (m.part >= 0) d = pmt::dict_add(d, pmt::intern("part"), pmt::from_long(m.part));
    if (pac) {
        d = pmt::dict_add(d, pmt::intern("rel_cfreq"), pmt::from_double(m.rel_cfreq));
        d = pmt::dict_add(d, pmt::intern("rel_bw"), pmt::from_double(m.rel_bw));
    } else {
        d = pmt::dict_add(d, pmt::intern("rel_bw"), pmt::from_double(m.rel_bw));
        d = pmt::dict_add(d, pmt::intern("rel_cfreq"), pmt::from_double(m.rel_cfreq));
    }
    d = pmt::dict_add(d, pmt::intern("blockstart"), pmt::from_long(m.blockstart));
    d = pmt::dict_add(d, pmt::intern("blockend"), pmt::from_long(m.blockend));
    if (!pac) {
        d = pmt::dict_add(d, pmt::intern("vectorstart"), pmt::from_long(m.vectorstart));
        d = pmt::dict_add(d, pmt::intern("vectorend"), pmt::from_long(m.vectorend));
    }
    pmt::pmt_t v = pmt::init_c32vector((size_t)m.nsamples, reinterpret_cast<const std::complex<float>*>(m.data));
    fdc_host_evict(m.data, sizeof(float) * 2 * (size_t)m.nsamples);     /* the payload may be a view of the next call's D2H destination (fdc_cabi.h) */
    return pmt::cons(d, v);
}

/* ---- copy / multiply blocks ------------------------------------------------------------------------------ */
class overlap_save_impl : public overlap_save {
    fdc_overlap_save* d_h;
public:
    overlap_save_impl(int itemsize, int outputlen, int overlaplen)
        : gr::sync_block("overlap_save", gr::io_signature::make(1, 1, itemsize * (outputlen - overlaplen)),
                         gr::io_signature::make(1, 1, itemsize * outputlen)),
          d_h(fdc_overlap_save_create(itemsize, outputlen, overlaplen))
    { if (!d_h) rethrow(); }
    ~overlap_save_impl() { fdc_overlap_save_destroy(d_h); }
    int work(int n, gr_vector_const_void_star& in, gr_vector_void_star& out)
    { return fdc_overlap_save_work(d_h, n, in[0], out[0]) < 0 ? -1 : n; }
};

class vector_cut_vxx_impl : public vector_cut_vxx {
    fdc_vector_cut* d_h;
public:
    vector_cut_vxx_impl(int itemsize, int veclen, int offset, int blocklen)
        : gr::sync_block("vector_cut_vxx", gr::io_signature::make(1, 1, itemsize * veclen), gr::io_signature::make(1, 1, itemsize * blocklen)),
          d_h(fdc_vector_cut_create(itemsize, veclen, offset, blocklen))
    { if (!d_h) rethrow(); }
    ~vector_cut_vxx_impl() { fdc_vector_cut_destroy(d_h); }
    int work(int n, gr_vector_const_void_star& in, gr_vector_void_star& out)
    { return fdc_vector_cut_work(d_h, n, in[0], out[0]) < 0 ? -1 : n; }
};

class phase_shifting_windowing_vcc_impl : public phase_shifting_windowing_vcc {
public:
    fdc_psw* d_h;
    phase_shifting_windowing_vcc_impl(int blocklen, int numphasestates, int shifts, float passbw, float stopbw, int windowtype)
        : gr::sync_block("phase_shifting_windowing_vcc", gr::io_signature::make(1, 1, (int)sizeof(gr_complex) * blocklen),
                         gr::io_signature::make(1, 1, (int)sizeof(gr_complex) * blocklen)),
          d_h(fdc_psw_create(blocklen, numphasestates, shifts, passbw, stopbw, windowtype))
    { if (!d_h) rethrow(); }
    ~phase_shifting_windowing_vcc_impl() { fdc_psw_destroy(d_h); }
    int work(int n, gr_vector_const_void_star& in, gr_vector_void_star& out)
    { return fdc_psw_work(d_h, n, in[0], out[0]) < 0 ? -1 : n; }
};

/* ---- activity-gated blocks: vector sink + message source -------------------------------------------------- */
template <class H, int (*COUNT)(const H*), int (*GET)(const H*, int, fdc_msg*), void (*CLEAR)(H*)>
void publish_all(gr::sync_block* blk, H* h)
{
    const int n = COUNT(h);
    for (int i = 0; i < n; i++) {
        fdc_msg m;
        if (GET(h, i, &m) == 0) blk->message_port_pub(pmt::intern("msgout"), to_pdu(m));
    }
    CLEAR(h);
}

class PowerActivationChannel_impl : public PowerActivationChannel {
public:
    fdc_pac* d_h;
    PowerActivationChannel_impl(int blocklen, float cfreq, float bw, int relinvovl, float thresh, int maxblocks, int deactivation_delay,
                                bool msg, bool fileoutput, std::string path, int verbose, int ID)
        : gr::sync_block("PowerActivationChannel", gr::io_signature::make(1, 1, (int)sizeof(gr_complex) * blocklen), gr::io_signature::make(0, 0, 0)),
          d_h(fdc_pac_create(blocklen, cfreq, bw, relinvovl, thresh, maxblocks, deactivation_delay, msg, fileoutput, path.c_str(), verbose, ID))
    { if (!d_h) rethrow(); message_port_register_out(pmt::intern("msgout")); }
    ~PowerActivationChannel_impl() { fdc_pac_destroy(d_h); }
    int work(int n, gr_vector_const_void_star& in, gr_vector_void_star&)
    {
        if (fdc_pac_work_host(d_h, n, in[0]) < 0) return -1;
        publish_all<fdc_pac, fdc_pac_msg_count, fdc_pac_msg_get, fdc_pac_msg_clear>(this, d_h);
        return n;
    }
};

class SegmentDetection_impl : public SegmentDetection {
public:
    fdc_segdet* d_h;
    SegmentDetection_impl(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh, float minchandist,
                          float window_flank_puffer, int maxblocks_to_emit, int channel_deactivation_delay, bool messageoutput,
                          bool fileoutput, std::string path, bool threads, int verbose)
        : gr::sync_block("SegmentDetection", gr::io_signature::make(1, 1, (int)sizeof(gr_complex) * blocklen), gr::io_signature::make(0, 0, 0)),
          d_h(fdc_segdet_create(ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer, maxblocks_to_emit,
                                channel_deactivation_delay, messageoutput, fileoutput, path.c_str(), threads, verbose))
    { if (!d_h) rethrow(); message_port_register_out(pmt::intern("msgout")); }
    ~SegmentDetection_impl() { fdc_segdet_destroy(d_h); }
    int work(int n, gr_vector_const_void_star& in, gr_vector_void_star&)
    {
        if (fdc_segdet_work_host(d_h, n, in[0]) < 0) return -1;
        publish_all<fdc_segdet, fdc_segdet_msg_count, fdc_segdet_msg_get, fdc_segdet_msg_clear>(this, d_h);
        return n;
    }
};

class activity_detection_channelizer_vcm_impl : public activity_detection_channelizer_vcm {
    static std::vector<float> flat(const std::vector<std::vector<float> >& segs)
    {
        std::vector<float> f;
        for (size_t i = 0; i < segs.size(); i++) {
            /* lib/activity_detection_channelizer_vcm_impl.cc:106-111: every segment is a (start, stop) pair */
            if (segs[i].size() != 2) throw std::invalid_argument("Segment does not contain start and stop frequency. ");
            f.push_back(segs[i][0]); f.push_back(segs[i][1]);
        }
        return f;
    }
public:
    fdc_actdet* d_h;
    activity_detection_channelizer_vcm_impl(int blocklen, std::vector<std::vector<float> > segments, float thresh, int relinvovl, int maxblocks,
                                            bool message, bool fileoutput, std::string path, bool threads, float minchandist,
                                            int channel_deactivation_delay, double window_flank_puffer, int verbose)
        : gr::sync_block("activity_detection_channelizer_vcm", gr::io_signature::make(1, 1, (int)sizeof(gr_complex) * blocklen),
                         gr::io_signature::make(0, 0, 0)),
          d_h(0)
    {
        const std::vector<float> f = flat(segments);
        d_h = fdc_actdet_create(blocklen, f.data(), (int)segments.size(), thresh, relinvovl, maxblocks, message, fileoutput, path.c_str(), threads,
                                minchandist, channel_deactivation_delay, window_flank_puffer, verbose);
        if (!d_h) rethrow();
        message_port_register_out(pmt::intern("msgout"));
    }
    ~activity_detection_channelizer_vcm_impl() { fdc_actdet_destroy(d_h); }
    int work(int n, gr_vector_const_void_star& in, gr_vector_void_star&)
    {
        if (fdc_actdet_work_host(d_h, n, in[0]) < 0) return -1;
        publish_all<fdc_actdet, fdc_actdet_msg_count, fdc_actdet_msg_get, fdc_actdet_msg_clear>(this, d_h);
        return n;
    }
};

}  // namespace

/* ---- the public factories ---------------------------------------------------------------------------------- */
overlap_save::sptr overlap_save::make(int itemsize, int outputlen, int overlaplen)
{ return gnuradio::get_initial_sptr(new overlap_save_impl(itemsize, outputlen, overlaplen)); }
vector_cut_vxx::sptr vector_cut_vxx::make(int itemsize, int veclen, int offset, int blocklen)
{ return gnuradio::get_initial_sptr(new vector_cut_vxx_impl(itemsize, veclen, offset, blocklen)); }
phase_shifting_windowing_vcc::sptr phase_shifting_windowing_vcc::make(int blocklen, int numphasestates, int shifts, float passbw, float stopbw,
                                                                      int windowtype)
{ return gnuradio::get_initial_sptr(new phase_shifting_windowing_vcc_impl(blocklen, numphasestates, shifts, passbw, stopbw, windowtype)); }
PowerActivationChannel::sptr PowerActivationChannel::make(int v_blocklen, float v_cfreq, float v_bw, int v_relinvovl, float v_thresh,
                                                          int v_maxblocks, int v_deactivation_delay, bool v_msg, bool v_fileoutput,
                                                          std::string v_path, int verbose, int v_ID)
{
    return gnuradio::get_initial_sptr(new PowerActivationChannel_impl(v_blocklen, v_cfreq, v_bw, v_relinvovl, v_thresh, v_maxblocks,
                                                                      v_deactivation_delay, v_msg, v_fileoutput, v_path, verbose, v_ID));
}
SegmentDetection::sptr SegmentDetection::make(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh, float minchandist,
                                              float window_flank_puffer, int maxblocks_to_emit, int channel_deactivation_delay,
                                              bool messageoutput, bool fileoutput, std::string path, bool threads, int verbose)
{
    return gnuradio::get_initial_sptr(new SegmentDetection_impl(ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist,
                                                                window_flank_puffer, maxblocks_to_emit, channel_deactivation_delay,
                                                                messageoutput, fileoutput, path, threads, verbose));
}
activity_detection_channelizer_vcm::sptr activity_detection_channelizer_vcm::make(
    int v_blocklen, std::vector<std::vector<float> > v_segments, float v_thresh, int v_relinvovl, int v_maxblocks, bool v_message,
    bool v_fileoutput, std::string v_path, bool v_threads, float v_minchandist, int v_channel_deactivation_delay,
    double v_window_flank_puffer, int verbose)
{
    return gnuradio::get_initial_sptr(new activity_detection_channelizer_vcm_impl(v_blocklen, v_segments, v_thresh, v_relinvovl, v_maxblocks,
                                                                                  v_message, v_fileoutput, v_path, v_threads, v_minchandist,
                                                                                  v_channel_deactivation_delay, v_window_flank_puffer, verbose));
}

}  // namespace FDC
}  // namespace gr

#ifdef FDC_GR_TEST_HOOKS
#include FDC_GR_TEST_HOOKS      /* tests/grshim/hooks.inc: raw-pointer factories + state getters for the shared test driver */
#endif
