/* Oracle TU: unmodified /root/reference/lib/SegmentDetection_impl.cc */
#include "ref_common.h"
#define private public
#include "SegmentDetection_impl.cc"
#undef private
using gr::FDC::SegmentDetection_impl;
extern "C" gr::sync_block* ref_segdet_make(int ID, int blocklen, int relinvovl, float seg_start, float seg_stop, float thresh,
                                           float minchandist, float window_flank_puffer, int maxblocks_to_emit,
                                           int channel_deactivation_delay, int messageoutput, int fileoutput, const char* path,
                                           int threads, int verbose)
{ REF_TRY return new SegmentDetection_impl(ID, blocklen, relinvovl, seg_start, seg_stop, thresh, minchandist, window_flank_puffer,
                                            maxblocks_to_emit, channel_deactivation_delay, messageoutput != 0, fileoutput != 0,
                                            std::string(path ? path : ""), threads != 0, verbose); REF_CATCH(0) }
/* geo: d_start, d_stop, d_width, D, M(=power.size()), blockcount, n_active, active_channels_counter ; f: thresh */
extern "C" int ref_segdet_state(gr::sync_block* b, long* geo, float* f)
{
    SegmentDetection_impl* p = dynamic_cast<SegmentDetection_impl*>(b);
    if (!p) return -1;
    geo[0] = (long)p->d_start; geo[1] = (long)p->d_stop; geo[2] = (long)p->d_width; geo[3] = (long)p->d_chan_detection_decimation_factor;
    geo[4] = (long)p->d_power.size(); geo[5] = (long)p->d_blockcount; geo[6] = (long)p->d_active_channels.size();
    geo[7] = (long)p->d_active_channels_counter; f[0] = p->d_thresh; return 0;
}
/* window for width 2^s, phase i: out 2*2^s floats */
extern "C" int ref_segdet_window(gr::sync_block* b, int s, int i, float* out)
{
    SegmentDetection_impl* p = dynamic_cast<SegmentDetection_impl*>(b);
    if (!p || s < 0 || s >= (int)p->d_windows.size() || i < 0 || i >= (int)p->d_windows[s].size()) return -1;
    memcpy(out, p->d_windows[s][i].data(), sizeof(gr_complex) * p->d_windows[s][i].size()); return 0;
}
/* last measured decimated power vector */
extern "C" int ref_segdet_power(gr::sync_block* b, float* out)
{
    SegmentDetection_impl* p = dynamic_cast<SegmentDetection_impl*>(b);
    if (!p) return -1;
    memcpy(out, p->d_power.data(), sizeof(float) * p->d_power.size()); return 0;
}
/* active channel i: ID, detect_start, detect_stop, extract_start, extract_stop, extract_width, ovlskip, outputsamples,
 *                   count, phase, phaseincrement, inactive, part, data.size() */
extern "C" int ref_segdet_active(gr::sync_block* b, int i, int* out)
{
    SegmentDetection_impl* p = dynamic_cast<SegmentDetection_impl*>(b);
    if (!p || i < 0 || i >= (int)p->d_active_channels.size()) return -1;
    const gr::FDC::active_channel& c = p->d_active_channels[i];
    out[0] = c.ID; out[1] = c.detect_start; out[2] = c.detect_stop; out[3] = c.extract_start; out[4] = c.extract_stop;
    out[5] = c.extract_width; out[6] = c.ovlskip; out[7] = c.outputsamples; out[8] = c.count; out[9] = c.phase;
    out[10] = c.phaseincrement; out[11] = c.inactive; out[12] = c.part; out[13] = (int)c.data.size(); return 0;
}
