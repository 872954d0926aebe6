/* Oracle TU: unmodified /root/reference/lib/activity_detection_channelizer_vcm_impl.cc */
#include "ref_common.h"
#define private public
#include "activity_detection_channelizer_vcm_impl.cc"
#undef private
using gr::FDC::activity_detection_channelizer_vcm_impl;
/* segs: nsegs pairs (start, stop) */
extern "C" gr::sync_block* ref_actdet_make(int blocklen, const float* segs, int nsegs, float thresh, int relinvovl, int maxblocks,
                                           int message, int fileoutput, const char* path, int threads, float minchandist,
                                           int channel_deactivation_delay, double window_flank_puffer, int verbose)
{
    REF_TRY
    std::vector<std::vector<float> > v(nsegs);
    for (int i = 0; i < nsegs; i++) { v[i].push_back(segs[2 * i]); v[i].push_back(segs[2 * i + 1]); }
    return new activity_detection_channelizer_vcm_impl(blocklen, v, thresh, relinvovl, maxblocks, message != 0, fileoutput != 0,
                                                       std::string(path ? path : ""), threads != 0, minchandist,
                                                       channel_deactivation_delay, window_flank_puffer, verbose);
    REF_CATCH(0)
}
/* per segment i: ID, start, stop, width, D, M, n_active */
extern "C" int ref_actdet_segment(gr::sync_block* b, int i, int* out)
{
    activity_detection_channelizer_vcm_impl* p = dynamic_cast<activity_detection_channelizer_vcm_impl*>(b);
    if (!p || i < 0 || i >= (int)p->segments.size()) return -1;
    const gr::FDC::segment& s = p->segments[i];
    out[0] = s.ID; out[1] = s.start; out[2] = s.stop; out[3] = s.width; out[4] = s.chan_detection_decimation_factor;
    out[5] = (int)s.power.size(); out[6] = (int)s.active_channels.size(); return 0;
}
extern "C" int ref_actdet_nsegments(gr::sync_block* b)
{
    activity_detection_channelizer_vcm_impl* p = dynamic_cast<activity_detection_channelizer_vcm_impl*>(b);
    return p ? (int)p->segments.size() : -1;
}
extern "C" int ref_actdet_power(gr::sync_block* b, int i, float* out)
{
    activity_detection_channelizer_vcm_impl* p = dynamic_cast<activity_detection_channelizer_vcm_impl*>(b);
    if (!p || i < 0 || i >= (int)p->segments.size()) return -1;
    memcpy(out, p->segments[i].power.data(), sizeof(float) * p->segments[i].power.size()); return 0;
}
